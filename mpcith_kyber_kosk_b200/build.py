"""Build libkosk_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

nvcc cross-compiles without a GPU, so this runs on the CPU-only build box as well as on a B200 box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkosk_b200.so")
SOURCES = ["kosk_b200.cu"]
HOST_SOURCES = ["wire_host.cpp"]          # host-only code (SIMD byte codec + worker pool of the compact wire format): g++, linked into the same library
HEADERS = ["kosk_common.cuh", "keccak.cuh", "gf_gemm.cuh", "gf_gemm_imma.cuh", "prove_kernels.cuh", "verify_kernels.cuh", "raw_api.cuh", "kem_kernels.cuh", "share_ntt.cuh",
           "wire_kernels.cuh", "wire_host.h", "component_api.cuh"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libkosk_b200.so")


SHIMS = {2: "libkosk_kyber512.so", 3: "libkosk_kyber768.so", 4: "libkosk_kyber1024.so"}     # binary drop-in shims (csrc/dropin_shim.cpp)


def build_shims(force=False):
    """libkosk_kyber{512,768,1024}.so: the reference's own symbol names over libkosk_b200.so (g++ only, no device code)."""
    src = os.path.join(CSRC, "dropin_shim.cpp")
    deps = [src, os.path.join(HERE, "..", "include", "kosk_dropin.hpp"), os.path.join(HERE, "..", "include", "kosk_b200.h")]
    out = []
    for k, name in SHIMS.items():
        lib = os.path.join(HERE, name)
        if force or not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
            r = subprocess.run(["g++", "-std=c++11", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", f"-DKYBER_K={k}", src, "-L" + HERE, "-lkosk_b200",
                                "-Wl,-rpath,$ORIGIN", "-o", lib], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("g++ failed for the drop-in shim:\n" + r.stdout + r.stderr)
        out.append(lib)
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HOST_SOURCES + HEADERS] + [os.path.join(HERE, "..", "include", "kosk_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name, defines):
    """Experiment helper: a second copy of the library with extra -D macros (mpcith_kyber_kosk_b200/libkosk_b200_<name>.so)."""
    lib = os.path.join(HERE, f"libkosk_b200_{name}.so")
    objs = [os.path.join(CSRC, src.rsplit(".", 1)[0] + ".o") for src in HOST_SOURCES]
    cmd = [_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", lib] + [os.path.join(CSRC, s) for s in SOURCES] + objs + ["-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return lib


def build(force=False, verbose=False):
    if not force and not needs_build():
        build_shims()
        return LIB
    objs = []
    for src in HOST_SOURCES:
        obj = os.path.join(CSRC, src.rsplit(".", 1)[0] + ".o")
        r = subprocess.run(["g++", "-O3", "-std=c++17", "-fPIC", "-pthread", "-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
        objs.append(obj)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + objs + ["-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    build_shims(force=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
