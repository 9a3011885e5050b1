"""Static SASS instruction mix of the loops of one kernel of libkosk_b200.so (no GPU needed): `python tools/sass_loops.py <substring of the
mangled kernel name> [library]`.  Prints the total and, for every backward branch, the opcode histogram of the loop body; bench.py's
executed-operation counts (FMA-heavy issue slots per sharing of the share evaluation) are read off this output."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernel_sass(pattern, lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, keep = None, []
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur and pattern in cur:
            m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
            if m:
                keep.append((int(m.group(1), 16), m.group(2)))
    return keep


def opcode(text):
    t = re.sub(r"^@!?U?P\d+\s+", "", text).split()[0]
    base = t.split(".")[0]
    if base == "IMAD" and (".HI" in t or ".WIDE" in t):
        base += ".HI" if ".HI" in t else ".WIDE"
    if base == "IMAD" and ".MOV" in t:
        base = "IMAD.MOV"
    return base


def main():
    pattern = sys.argv[1]
    lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "mpcith_kyber_kosk_b200", "libkosk_b200.so")
    ins = kernel_sass(pattern, lib)
    print("instructions:", len(ins), dict(collections.Counter(opcode(t) for _, t in ins).most_common(14)))
    for a, t in ins:
        if "BRA" in t:
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                lo = int(m.group(1), 16)
                body = [x for x in ins if lo <= x[0] <= a]
                if len(body) >= 64:
                    print(f"loop {lo:#x}..{a:#x}: {len(body)} instructions", dict(collections.Counter(opcode(x[1]) for x in body).most_common(16)))


if __name__ == "__main__":
    main()
