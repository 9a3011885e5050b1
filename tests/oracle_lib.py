"""ctypes bindings of the CPU oracles (TEST INFRASTRUCTURE): the plain-C restatement (oracle/kosk_oracle.c)
and, when present, the unmodified reference compiled into oracle/_ref/."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libkosk_oracle.so")


class Layout(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in "k eta F E M".split()] + \
               [(n, ctypes.c_size_t) for n in "pk_bytes sk_bytes proof_bytes".split()] + \
               [(n, ctypes.c_size_t) for n in ("o_f o_Tf o_beta o_gamma o_Tcomm o_I o_s o_e o_t o_NTTs o_NTTe o_NTTAr o_NTTAs "
                                               "o_sr o_er o_seta o_eeta o_ssub o_esub o_zs o_ze o_us o_ue o_comm").split()]


class Trace(ctypes.Structure):
    _fields_ = [("alpha", ctypes.c_uint16 * 78), ("fs1_digest", ctypes.c_uint8 * 32), ("fs2_digest", ctypes.c_uint8 * 32),
                ("I", ctypes.c_uint16 * 150), ("first_share", ctypes.c_uint16 * 1454), ("first_secret", ctypes.c_uint16 * 256),
                ("first_ntt", ctypes.c_uint16 * 256), ("tcomm0", ctypes.c_uint8 * 32), ("view0", ctypes.c_uint8 * 32)]


FIELDS = [f[0] for f in Layout._fields_[8:]]


def build_oracle():
    if not os.path.exists(ORACLE_SO) or any(
            os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(ORACLE_SO)
            for f in ("kosk_oracle.c", "kosk_oracle.h", "ok_keccak.h", "ok_rng.c", "ok_tables.c")):
        subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, capture_output=True)
    return ORACLE_SO


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        lib = ctypes.CDLL(build_oracle())
        vp = ctypes.c_void_p
        lib.kosk_oracle_layout.argtypes = [ctypes.c_int, ctypes.POINTER(Layout)]
        lib.kosk_oracle_verifiable_keygen.argtypes = [ctypes.c_int, vp, ctypes.c_int, vp, vp, vp]
        lib.kosk_oracle_verifiable_keygen.restype = None
        lib.kosk_oracle_verify.argtypes = [ctypes.c_int, vp, vp]
        lib.kosk_oracle_last_trace.restype = ctypes.POINTER(Trace)
        lib.ko_share_ddeg.argtypes = [vp, vp]
        lib.ko_recon_ddeg.argtypes = [vp, vp]
        lib.ko_recon_2ddeg.argtypes = [vp, vp]
        lib.ko_ntt.argtypes = [vp]
        lib.ko_sha3_256.argtypes = [vp, vp, ctypes.c_size_t]
        lib.ko_shake256.argtypes = [vp, ctypes.c_size_t, vp, ctypes.c_size_t]
        lib.ko_shake128.argtypes = [vp, ctypes.c_size_t, vp, ctypes.c_size_t]
        lib.ko_sha3_512.argtypes = [vp, vp, ctypes.c_size_t]
        lib.ko_keccak_f1600.argtypes = [vp]
        lib.ko_randombytes_at.argtypes = [vp, ctypes.c_uint32, vp, ctypes.c_size_t]
        lib.ko_gen_matrix.argtypes = [ctypes.c_int, vp, vp]
        _oracle = lib
    return _oracle


def layout(k):
    L = Layout()
    assert oracle().kosk_oracle_layout(k, ctypes.byref(L)) == 0
    return L


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def oracle_prove(k, seed, rng_mode=0):
    L = layout(k)
    seed = np.frombuffer(bytes(seed), dtype=np.uint8).copy()
    pk, sk, pi = np.zeros(L.pk_bytes, np.uint8), np.zeros(L.sk_bytes, np.uint8), np.zeros(L.proof_bytes, np.uint8)
    oracle().kosk_oracle_verifiable_keygen(k, _p(seed), rng_mode, _p(pk), _p(sk), _p(pi))
    return pk, sk, pi


def oracle_trace():
    return oracle().kosk_oracle_last_trace().contents


def oracle_verify(k, pi, pk):
    pi = np.ascontiguousarray(pi, dtype=np.uint8)
    pk = np.ascontiguousarray(pk, dtype=np.uint8)
    return oracle().kosk_oracle_verify(k, _p(pi), _p(pk)) == 1


def oracle_share(y):
    y = np.ascontiguousarray(y, dtype=np.uint16)
    out = np.zeros(1454, np.uint16)
    oracle().ko_share_ddeg(_p(out), _p(y))
    return out


def oracle_ntt(a):
    a = np.array(a, dtype=np.uint16)
    oracle().ko_ntt(_p(a))
    return a


_refs = {}


def ref(k):
    """The unmodified reference (oracle/_ref), or None when it has not been built (e.g. no /root/reference)."""
    if k not in _refs:
        path = os.path.join(ORACLE_DIR, "_ref", f"libkosk_ref_k{k}.so")
        if not os.path.exists(path):
            _refs[k] = None
        else:
            lib = ctypes.CDLL(path)
            vp = ctypes.c_void_p
            lib.ref_verifiable_keygen.argtypes = [vp, ctypes.c_int, vp, vp, vp]
            lib.ref_verifiable_keygen.restype = None
            lib.ref_kosk_verify.argtypes = [vp, vp]
            for n in ("ref_inst_bytes", "ref_randomness_bytes", "ref_range_proof_bytes", "ref_share_vec_bytes"):
                getattr(lib, n).restype = ctypes.c_size_t
            lib.ref_struct_sequence.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, vp, vp]
            lib.ref_verify_struct.argtypes = [vp, vp]
            if hasattr(lib, "ref_prove_struct_at"):
                lib.ref_prove_struct_at.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp]
                lib.ref_prove_struct_at.restype = None
            lib.ref_ct_bytes.restype = ctypes.c_size_t
            lib.ref_kem_enc_derand.argtypes = [vp, vp, vp, vp]
            lib.ref_kem_enc_derand.restype = None
            lib.ref_kem_dec.argtypes = [vp, vp, vp]
            lib.ref_kem_dec.restype = None
            lib.ref_kem_keypair_derand.argtypes = [vp, vp, vp]
            lib.ref_kem_keypair_derand.restype = None
            lib.ref_kem_keypair_at.argtypes = [vp, ctypes.c_uint32, vp, vp]
            lib.ref_kem_keypair_at.restype = None
            lib.ref_kem_enc_at.argtypes = [vp, ctypes.c_uint32, vp, vp, vp]
            lib.ref_kem_enc_at.restype = None
            _refs[k] = lib
    return _refs[k]


def ref_prove(k, seed, rng_mode=0):
    L = layout(k)
    seed = np.frombuffer(bytes(seed), dtype=np.uint8).copy()
    pk, sk, pi = np.zeros(L.pk_bytes, np.uint8), np.zeros(L.sk_bytes, np.uint8), np.zeros(L.proof_bytes, np.uint8)
    ref(k).ref_verifiable_keygen(_p(seed), rng_mode, _p(pk), _p(sk), _p(pi))
    return pk, sk, pi


def ref_verify(k, pi, pk):
    pi = np.ascontiguousarray(pi, dtype=np.uint8)
    pk = np.ascontiguousarray(pk, dtype=np.uint8)
    return ref(k).ref_kosk_verify(_p(pi), _p(pk)) == 1


# ---- struct-level API (reference main.cpp:16-59): byte images of mpcith_randomness / mpcith_range_proof / mlwe_inst ----
SHARE_VEC_BYTES = 8 + 4 * 1454


def struct_sizes(k):
    F, E = 70 + 2 * k + 1, 2 * (3 if k == 2 else 2) + 1
    return {"inst": (k * k + 3 * k) * 512, "rand": F * (1024 + 2 * SHARE_VEC_BYTES), "eta": 2 * k * E * SHARE_VEC_BYTES,
            "rand_shares_off": 1024 * F, "F": F, "E": E}


def mask_share_vec_len(img, first_off):
    """share_vec.len (ss.hpp:34) is never written by the reference: zero it before comparing images."""
    img = np.array(img, dtype=np.uint8, copy=True)
    for o in range(first_off, img.size, SHARE_VEC_BYTES):
        img[o:o + 8] = 0
    return img


def ref_struct_sequence(k, seed, rng_mode=0):
    """prepare_randomness, prepare_range_proof, kyber_keygen, prove, verify of the unmodified reference, in main.cpp's order."""
    L, S = layout(k), struct_sizes(k)
    lib = ref(k)
    assert (lib.ref_inst_bytes(), lib.ref_randomness_bytes(), lib.ref_range_proof_bytes(), lib.ref_share_vec_bytes()) == \
        (S["inst"], S["rand"], S["eta"], SHARE_VEC_BYTES)
    seed = np.frombuffer(bytes(seed), dtype=np.uint8).copy()
    rnd, eta, inst = np.zeros(S["rand"], np.uint8), np.zeros(S["eta"], np.uint8), np.zeros(S["inst"], np.uint8)
    pk, sk, pi = np.zeros(L.pk_bytes, np.uint8), np.zeros(L.sk_bytes, np.uint8), np.zeros(L.proof_bytes, np.uint8)
    ok = lib.ref_struct_sequence(_p(seed), rng_mode, _p(rnd), _p(eta), _p(inst), _p(pk), _p(sk), _p(pi))
    return {"rand": mask_share_vec_len(rnd, S["rand_shares_off"]), "eta": mask_share_vec_len(eta, 0), "inst": inst,
            "pk": pk, "sk": sk, "pi": pi, "ok": ok == 1}


def ref_prove_struct_at(k, seed, call, inst, rand, eta):
    """The reference's prove() on the given struct images, its randombytes() positioned at call number `call`."""
    L = layout(k)
    seed = np.frombuffer(bytes(seed), dtype=np.uint8).copy()
    inst, rand, eta = (np.ascontiguousarray(x, dtype=np.uint8).copy() for x in (inst, rand, eta))
    pi = np.zeros(L.proof_bytes, np.uint8)
    ref(k).ref_prove_struct_at(_p(seed), int(call), _p(inst), _p(rand), _p(eta), _p(pi))
    return pi


def ref_verify_struct(k, pi, inst):
    pi = np.frombuffer(bytes(pi), dtype=np.uint8).copy()
    inst = np.ascontiguousarray(inst, dtype=np.uint8)
    return ref(k).ref_verify_struct(_p(pi), _p(inst)) == 1


# ---- Kyber KEM of the unmodified reference (kyber/kem.c) ----
CT_BYTES = {2: 768, 3: 1088, 4: 1568}


def _u8(x):
    return np.frombuffer(bytes(x), dtype=np.uint8).copy()


def ref_kem_enc_derand(k, pk, coins):
    pk, coins = _u8(pk), _u8(coins)
    ct, ss = np.zeros(ref(k).ref_ct_bytes(), np.uint8), np.zeros(32, np.uint8)
    ref(k).ref_kem_enc_derand(_p(ct), _p(ss), _p(pk), _p(coins))
    return ct, ss


def ref_kem_dec(k, ct, sk):
    ct, sk = _u8(ct), _u8(sk)
    ss = np.zeros(32, np.uint8)
    ref(k).ref_kem_dec(_p(ss), _p(ct), _p(sk))
    return ss


def ref_kem_keypair_derand(k, coins):
    L = layout(k)
    coins = _u8(coins)
    pk, sk = np.zeros(L.pk_bytes, np.uint8), np.zeros(L.sk_bytes, np.uint8)
    ref(k).ref_kem_keypair_derand(_p(pk), _p(sk), _p(coins))
    return pk, sk


def ref_kem_keypair_at(k, seed, call):
    L = layout(k)
    seed = _u8(seed)
    pk, sk = np.zeros(L.pk_bytes, np.uint8), np.zeros(L.sk_bytes, np.uint8)
    ref(k).ref_kem_keypair_at(_p(seed), call, _p(pk), _p(sk))
    return pk, sk


def ref_kem_enc_at(k, seed, call, pk):
    seed, pk = _u8(seed), _u8(pk)
    ct, ss = np.zeros(ref(k).ref_ct_bytes(), np.uint8), np.zeros(32, np.uint8)
    ref(k).ref_kem_enc_at(_p(seed), call, _p(ct), _p(ss), _p(pk))
    return ct, ss


def seed_of(i, tag=b"kosk-b200"):
    import hashlib
    return hashlib.sha256(tag + b":" + str(i).encode()).digest()
