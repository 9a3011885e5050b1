"""CPU suite: the C-ABI library builds for sm_100a, loads, exports every symbol include/kosk_b200.h declares,
reports the reference's sizes, and fails loudly without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import mpcith_kyber_kosk_b200 as pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "kosk_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kosk_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/kosk_b200.h but not exported"
    assert sorted(pkg.EXPORTS) == names


def test_sizes_match_reference(built_lib):
    assert [(pkg.pk_bytes(k), pkg.sk_bytes(k), pkg.proof_bytes(k)) for k in (2, 3, 4)] == \
        [(800, 1632, 664340), (1184, 2400, 680980), (1568, 3168, 744148)]
    assert pkg.proof_bytes(5) == 0


def test_bad_arguments(built_lib):
    lib = pkg.load_library()
    h = ctypes.c_void_p()
    assert lib.kosk_b200_create(ctypes.byref(h), 7, 0, 0) == -1
    assert b"kyber_k" in lib.kosk_b200_last_error()


def test_no_cpu_fallback(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.KoskError, match="no CUDA device"):
        pkg.KoskContext(2)


def test_sass_is_sm100a(built_lib):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_dropin_header_compiles_like_the_reference_api(built_lib, tmp_path):
    """include/kosk_dropin.hpp exposes the reference's names and signatures (kosk.hpp:13-24) for every KYBER_K."""
    import subprocess
    src = tmp_path / "t.cpp"
    src.write_text('#include "kosk_dropin.hpp"\n'
                   'static_assert(sizeof(kyber_keypair) == KYBER_PUBLICKEYBYTES + KYBER_SECRETKEYBYTES, "layout");\n'
                   'void (*f)(kyber_keypair *, uint8_t *) = &kyber_verifiable_keygen;\n'
                   'bool (*g)(const uint8_t *, const uint8_t *) = &kyber_kosk_verify;\n'
                   'int main() { return (int)(MPCITH_PROOF_SIZE == 0); }\n')
    libdir = os.path.join(ROOT, "mpcith_kyber_kosk_b200")
    for k, sizes in ((2, 2432), (3, 3584), (4, 4736)):
        exe = tmp_path / f"t{k}"
        subprocess.run(["g++", "-std=c++11", f"-DKYBER_K={k}", "-I" + os.path.join(ROOT, "include"), str(src), "-L" + libdir,
                        "-lkosk_b200", "-Wl,-rpath," + libdir, "-o", str(exe)], check=True)
        assert subprocess.run([str(exe)]).returncode == 0


def test_struct_image_sizes_match_reference_structs(built_lib):
    """kosk_b200_{inst,randomness,range_proof}_bytes == sizeof(mlwe_inst / mpcith_randomness / mpcith_range_proof)."""
    import oracle_lib as O
    lib = pkg.load_library()
    for k in (2, 3, 4):
        S = O.struct_sizes(k)
        assert (lib.kosk_b200_inst_bytes(k), lib.kosk_b200_randomness_bytes(k), lib.kosk_b200_range_proof_bytes(k)) == (S["inst"], S["rand"], S["eta"])
        r = O.ref(k)
        if r is not None:      # the compiler's own sizeof of the reference structs
            assert (r.ref_inst_bytes(), r.ref_randomness_bytes(), r.ref_range_proof_bytes(), r.ref_share_vec_bytes()) == (S["inst"], S["rand"], S["eta"], O.SHARE_VEC_BYTES)
    assert lib.kosk_b200_inst_bytes(9) == 0


def test_dropin_header_exposes_the_struct_level_api(built_lib, tmp_path):
    """kosk_dropin.hpp / include/dropin/*: the reference's struct types (same sizes as the reference's own sizeof, when oracle/_ref is
    there to ask), the struct-level functions of mlwe_prover.hpp:77-99 / mlwe_verifier.hpp:14-15 / kosk.hpp:17-18 and the KEM calls,
    with the reference's signatures, for every KYBER_K (compile + link only: no GPU needed)."""
    import subprocess
    import oracle_lib as O
    src = tmp_path / "t.cpp"
    src.write_text('#include "mlwe_prover.hpp"\n#include "kosk.hpp"\n#include "params.hpp"\n'
                   'using namespace NTL;\n'
                   'void (*f1)(mpcith_randomness *) = &prepare_randomness;\n'
                   'void (*f2)(mpcith_range_proof *) = &prepare_range_proof;\n'
                   'void (*f3)(kyber_keypair *, mlwe_inst *) = &kyber_keygen;\n'
                   'void (*f4)(mpcith_proof *, const mlwe_inst *, const mpcith_randomness *, const mpcith_range_proof *) = &prove;\n'
                   'bool (*f5)(const mpcith_proof *, const mlwe_inst *) = &verify;\n'
                   'void (*f6)(uint8_t *, const mpcith_proof *) = &encode_mpcith_proof;\n'
                   'void (*f7)(mpcith_proof *, const uint8_t *) = &decode_mpcith_proof;\n'
                   'int (*f8)(uint8_t *, uint8_t *, const uint8_t *) = &crypto_kem_enc;\n'
                   'int (*f9)(uint8_t *, const uint8_t *, const uint8_t *) = &crypto_kem_dec;\n'
                   '#include <stdio.h>\n'
                   'int main() { printf("%zu %zu %zu %zu %zu %d %d\\n", sizeof(mlwe_inst), sizeof(mpcith_randomness), sizeof(mpcith_range_proof),\n'
                   '                    sizeof(share_vec), (size_t)MPCITH_PROOF_SIZE, (int)KYBER_CIPHERTEXTBYTES, (int)KYBER_SSBYTES); return 0; }\n')
    libdir = os.path.join(ROOT, "mpcith_kyber_kosk_b200")
    for k in (2, 3, 4):
        exe = tmp_path / f"t{k}"
        subprocess.run(["g++", "-std=c++11", f"-DKYBER_K={k}", "-I" + os.path.join(ROOT, "include", "dropin"), str(src), "-L" + libdir,
                        "-lkosk_b200", "-Wl,-rpath," + libdir, "-o", str(exe)], check=True)
        got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
        S = O.struct_sizes(k)
        assert got == [S["inst"], S["rand"], S["eta"], O.SHARE_VEC_BYTES, pkg.proof_bytes(k), O.CT_BYTES[k], 32]


def test_reference_main_cpp_compiles_against_dropin_headers(built_lib):
    """`make -C oracle dropin-main`: the reference's unmodified main.cpp against include/dropin/ + libkosk_b200.so (the GPU suite runs it)."""
    import subprocess
    if not os.path.exists("/root/reference/main.cpp"):
        pytest.skip("no /root/reference in this checkout")
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "dropin-main"], check=True, capture_output=True)
    for k in (2, 3, 4):
        assert os.path.exists(os.path.join(ROOT, "oracle", "_ref", f"main_dropin_k{k}"))


@pytest.mark.parametrize("k", [2, 3, 4])
def test_shim_exports_the_reference_symbols(built_lib, k):
    """Binary drop-in: every symbol an object compiled against the reference's own headers imports from the reference (main.cpp ->
    oracle/_ref/main_ref_k*.o; the reference's objects = oracle/_ref/libkosk_ref_k*.so) is defined, under the same (mangled) name, by
    libkosk_kyber{512,768,1024}.so; plus the three names SURVEY 8(b) lists."""
    import subprocess
    from mpcith_kyber_kosk_b200 import build
    shim = os.path.join(ROOT, "mpcith_kyber_kosk_b200", build.SHIMS[k])
    assert os.path.exists(shim)

    def syms(path, flag):
        out = subprocess.run(["nm", flag, path], capture_output=True, text=True).stdout
        return {l.split()[-1] for l in out.splitlines() if l.split()}
    have = syms(shim, "-D")
    for name in ("_Z23kyber_verifiable_keygenP13kyber_keypairPh", "_Z17kyber_kosk_verifyPKhS0_", "_Z12kyber_keygenP13kyber_keypairP9mlwe_inst",
                 "_Z5proveP12mpcith_proofPK9mlwe_instPK17mpcith_randomnessPK18mpcith_range_proof", "_Z6verifyPK12mpcith_proofPK9mlwe_inst",
                 "_Z18prepare_randomnessP17mpcith_randomness", "_Z19prepare_range_proofP18mpcith_range_proof",
                 "_Z19encode_mpcith_proofPhPK12mpcith_proof", "_Z19decode_mpcith_proofP12mpcith_proofPKh",
                 f"pqcrystals_kyber{256 * k}_ref_keypair", f"pqcrystals_kyber{256 * k}_ref_enc", f"pqcrystals_kyber{256 * k}_ref_dec"):
        assert name in have, name
    obj = os.path.join(ROOT, "oracle", "_ref", f"main_ref_k{k}.o")
    ref = os.path.join(ROOT, "oracle", "_ref", f"libkosk_ref_k{k}.so")
    if not (os.path.exists(obj) and os.path.exists(ref)):
        pytest.skip("oracle/_ref not built in this checkout (no /root/reference)")
    wanted = {l.split()[-1] for l in subprocess.run(["nm", "-u", obj], capture_output=True, text=True).stdout.splitlines() if l.split()}
    ref_defined = {l.split()[-1] for l in subprocess.run(["nm", "-D", "--defined-only", ref], capture_output=True, text=True).stdout.splitlines() if l.split()}
    from_reference = wanted & ref_defined
    assert len(from_reference) >= 9
    assert from_reference <= have, sorted(from_reference - have)
