"""Host-side mirror of the reference's KOSK interface (reference kosk.hpp:13-24) over the C ABI of
libkosk_b200.so (include/kosk_b200.h).  Python is only the binding used by tests and bench.py; all
field/hash arithmetic runs in the CUDA kernels.  There is no CPU fallback: loading fails loudly if the
library is missing, and context creation fails if there is no CUDA device.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkosk_b200.so")

EXPORTS = [
    "kosk_b200_pk_bytes", "kosk_b200_sk_bytes", "kosk_b200_proof_bytes", "kosk_b200_last_error", "kosk_b200_version",
    "kosk_b200_create", "kosk_b200_create_ex", "kosk_b200_lanes", "kosk_b200_destroy", "kosk_b200_verifiable_keygen", "kosk_b200_kosk_verify",
    "kosk_b200_prove_batch", "kosk_b200_prove_batch_async", "kosk_b200_verify_batch", "kosk_b200_prove_batch_device", "kosk_b200_verify_batch_device",
    "kosk_b200_share_eval", "kosk_b200_recon_rows", "kosk_b200_interp_rows", "kosk_b200_sha3_256_rows", "kosk_b200_ntt_rows", "kosk_b200_share_eval_device",
    "kosk_b200_kernel_launches", "kosk_b200_debug_fetch", "kosk_b200_debug_trace", "kosk_b200_sync",
    "kosk_b200_inst_bytes", "kosk_b200_randomness_bytes", "kosk_b200_range_proof_bytes", "kosk_b200_rng_reset", "kosk_b200_rng_calls", "kosk_b200_verifiable_keygen_rng",
    "kosk_b200_prepare_randomness", "kosk_b200_prepare_range_proof", "kosk_b200_keygen", "kosk_b200_prove", "kosk_b200_verify",
    "kosk_b200_ct_bytes", "kosk_b200_kem_enc_derand_batch", "kosk_b200_kem_dec_batch", "kosk_b200_kem_enc_derand_batch_device", "kosk_b200_kem_dec_batch_device",
    "kosk_b200_kem_enc", "kosk_b200_kem_dec", "kosk_b200_kem_keypair_derand_batch", "kosk_b200_kem_keypair",
    "kosk_b200_wire_bytes", "kosk_b200_set_wire", "kosk_b200_wire_info", "kosk_b200_wire_stats", "kosk_b200_prove_batch_packed", "kosk_b200_prove_batch_packed_async", "kosk_b200_verify_batch_packed",
    "kosk_b200_verify_batch_async", "kosk_b200_verify_batch_packed_async",
    "kosk_b200_wire_pack_device", "kosk_b200_wire_unpack_device", "kosk_b200_wire_pack", "kosk_b200_wire_unpack", "kosk_b200_wire_simd",
    "kosk_b200_pool_bytes", "kosk_b200_pool_export", "kosk_b200_pool_import", "kosk_b200_pool_create", "kosk_b200_pool_prove", "kosk_b200_pool_destroy", "kosk_b200_set_strict", "kosk_b200_set_profiling", "kosk_b200_phase_times", "kosk_b200_int_peak",
]

_lib = None


class KoskError(RuntimeError):
    pass


def load_library(path=None):
    """dlopen libkosk_b200.so and declare the prototypes of include/kosk_b200.h."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise KoskError(f"{path} not found: build it with `python -m mpcith_kyber_kosk_b200.build` (no CPU fallback exists)")
    lib = ctypes.CDLL(path)
    vp, sz, i32, u8p = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p
    for n in ("kosk_b200_pk_bytes", "kosk_b200_sk_bytes", "kosk_b200_proof_bytes", "kosk_b200_inst_bytes", "kosk_b200_randomness_bytes", "kosk_b200_range_proof_bytes", "kosk_b200_ct_bytes",
              "kosk_b200_wire_bytes"):
        getattr(lib, n).restype = sz
        getattr(lib, n).argtypes = [i32]
    lib.kosk_b200_last_error.restype = ctypes.c_char_p
    lib.kosk_b200_version.restype = ctypes.c_char_p
    lib.kosk_b200_create.argtypes = [ctypes.POINTER(vp), i32, i32, i32]
    lib.kosk_b200_create_ex.argtypes = [ctypes.POINTER(vp), i32, i32, i32, i32, i32]
    lib.kosk_b200_lanes.argtypes = [vp]
    lib.kosk_b200_destroy.argtypes = [vp]
    lib.kosk_b200_destroy.restype = None
    lib.kosk_b200_verifiable_keygen.argtypes = [vp, u8p, u8p, u8p, u8p]
    lib.kosk_b200_kosk_verify.argtypes = [vp, u8p, u8p]
    lib.kosk_b200_prove_batch.argtypes = [vp, sz, u8p, u8p, u8p, u8p]
    lib.kosk_b200_prove_batch_async.argtypes = [vp, sz, u8p, u8p, u8p, u8p]
    lib.kosk_b200_verify_batch.argtypes = [vp, sz, u8p, u8p, u8p]
    lib.kosk_b200_prove_batch_device.argtypes = [vp, sz, u8p, u8p, u8p, u8p, vp]
    lib.kosk_b200_verify_batch_device.argtypes = [vp, sz, u8p, u8p, u8p, vp]
    lib.kosk_b200_share_eval.argtypes = [vp, sz, u8p, u8p]
    lib.kosk_b200_recon_rows.argtypes = [vp, i32, sz, u8p, u8p]
    lib.kosk_b200_interp_rows.argtypes = [vp, i32, u8p, sz, u8p, u8p]
    lib.kosk_b200_sha3_256_rows.argtypes = [vp, sz, sz, u8p, u8p]
    lib.kosk_b200_ntt_rows.argtypes = [vp, sz, u8p]
    lib.kosk_b200_share_eval_device.argtypes = [vp, sz, u8p, u8p, vp]
    lib.kosk_b200_kernel_launches.argtypes = [vp]
    lib.kosk_b200_kernel_launches.restype = ctypes.c_uint64
    lib.kosk_b200_debug_fetch.argtypes = [vp, ctypes.c_char_p, u8p, sz]
    lib.kosk_b200_sync.argtypes = [vp]
    lib.kosk_b200_debug_trace.argtypes = [vp, u8p, i32]
    lib.kosk_b200_set_profiling.argtypes = [vp, i32]
    lib.kosk_b200_set_strict.argtypes = [vp, i32]
    lib.kosk_b200_pool_create.argtypes = [vp, sz, u8p, ctypes.POINTER(vp)]
    lib.kosk_b200_pool_prove.argtypes = [vp, u8p, u8p, u8p]
    lib.kosk_b200_pool_bytes.argtypes = [vp]
    lib.kosk_b200_pool_bytes.restype = sz
    lib.kosk_b200_pool_export.argtypes = [vp, u8p, sz]
    lib.kosk_b200_pool_import.argtypes = [vp, u8p, sz, ctypes.POINTER(vp)]
    lib.kosk_b200_pool_destroy.argtypes = [vp]
    lib.kosk_b200_pool_destroy.restype = None
    lib.kosk_b200_rng_reset.argtypes = [vp, u8p]
    lib.kosk_b200_rng_calls.argtypes = [vp]
    lib.kosk_b200_rng_calls.restype = ctypes.c_uint32
    lib.kosk_b200_verifiable_keygen_rng.argtypes = [vp, u8p, u8p, u8p]
    lib.kosk_b200_prepare_randomness.argtypes = [vp, u8p]
    lib.kosk_b200_prepare_range_proof.argtypes = [vp, u8p]
    lib.kosk_b200_keygen.argtypes = [vp, u8p, u8p, u8p]
    lib.kosk_b200_prove.argtypes = [vp, u8p, u8p, u8p, u8p]
    lib.kosk_b200_verify.argtypes = [vp, u8p, u8p]
    lib.kosk_b200_kem_enc_derand_batch.argtypes = [vp, sz, u8p, u8p, u8p, u8p]
    lib.kosk_b200_kem_dec_batch.argtypes = [vp, sz, u8p, u8p, u8p]
    lib.kosk_b200_kem_enc_derand_batch_device.argtypes = [vp, sz, u8p, u8p, u8p, u8p, vp]
    lib.kosk_b200_kem_dec_batch_device.argtypes = [vp, sz, u8p, u8p, u8p, vp]
    lib.kosk_b200_kem_keypair_derand_batch.argtypes = [vp, sz, u8p, u8p, u8p]
    lib.kosk_b200_kem_keypair.argtypes = [vp, u8p, u8p]
    lib.kosk_b200_kem_enc.argtypes = [vp, u8p, u8p, u8p]
    lib.kosk_b200_kem_dec.argtypes = [vp, u8p, u8p, u8p]
    lib.kosk_b200_set_wire.argtypes = [vp, i32, i32]
    lib.kosk_b200_wire_info.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(ctypes.c_char_p)]
    lib.kosk_b200_wire_stats.argtypes = [vp, u8p, i32]
    lib.kosk_b200_prove_batch_packed.argtypes = [vp, sz, u8p, u8p, u8p, u8p]
    lib.kosk_b200_prove_batch_packed_async.argtypes = [vp, sz, u8p, u8p, u8p, u8p]
    lib.kosk_b200_verify_batch_packed.argtypes = [vp, sz, u8p, u8p, u8p]
    lib.kosk_b200_verify_batch_async.argtypes = [vp, sz, u8p, u8p, u8p]
    lib.kosk_b200_verify_batch_packed_async.argtypes = [vp, sz, u8p, u8p, u8p]
    lib.kosk_b200_wire_pack_device.argtypes = [vp, sz, u8p, u8p, vp]
    lib.kosk_b200_wire_unpack_device.argtypes = [vp, sz, u8p, u8p, vp]
    lib.kosk_b200_wire_pack.argtypes = [i32, sz, u8p, u8p, i32]
    lib.kosk_b200_wire_unpack.argtypes = [i32, sz, u8p, u8p, i32]
    lib.kosk_b200_wire_simd.restype = ctypes.c_char_p
    lib.kosk_b200_phase_times.argtypes = [vp, u8p, u8p, i32, i32]
    lib.kosk_b200_int_peak.argtypes = [vp, u8p]
    if path == LIB_PATH:
        _lib = lib
    return lib


def pk_bytes(k):
    return load_library().kosk_b200_pk_bytes(k)


def sk_bytes(k):
    return load_library().kosk_b200_sk_bytes(k)


def proof_bytes(k):
    return load_library().kosk_b200_proof_bytes(k)


def wire_bytes(k):
    return load_library().kosk_b200_wire_bytes(k)


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def wire_pack(k, pi, threads=1):
    """Host codec: proofs in the reference layout (pi[n][proof_bytes]) -> compact wire images (include/kosk_b200.h)."""
    lib = load_library()
    pi = np.ascontiguousarray(pi, dtype=np.uint8).reshape(-1, proof_bytes(k))
    out = np.empty((pi.shape[0], wire_bytes(k)), np.uint8)
    rc = lib.kosk_b200_wire_pack(k, pi.shape[0], _ptr(pi), _ptr(out), threads)
    if rc != 0:
        raise KoskError(f"wire_pack failed ({rc}): {lib.kosk_b200_last_error().decode()}")
    return out


def wire_unpack(k, wire, threads=1):
    """Host codec: compact wire images -> proofs in the reference layout."""
    lib = load_library()
    wire = np.ascontiguousarray(wire, dtype=np.uint8).reshape(-1, wire_bytes(k))
    out = np.empty((wire.shape[0], proof_bytes(k)), np.uint8)
    rc = lib.kosk_b200_wire_unpack(k, wire.shape[0], _ptr(wire), _ptr(out), threads)
    if rc != 0:
        raise KoskError(f"wire_unpack failed ({rc}): {lib.kosk_b200_last_error().decode()}")
    return out


def wire_simd():
    return load_library().kosk_b200_wire_simd().decode()


class KoskContext:
    """One (device, KYBER_K) instance of the B200 KOSK core."""

    def __init__(self, kyber_k=2, device=0, max_chunk=0, lanes=0, tensor=False):
        self.lib = load_library()
        self.k = kyber_k
        self._h = ctypes.c_void_p()
        rc = self.lib.kosk_b200_create_ex(ctypes.byref(self._h), kyber_k, device, max_chunk, lanes, 1 if tensor else 0)
        if rc != 0:
            raise KoskError(f"kosk_b200_create failed ({rc}): {self.lib.kosk_b200_last_error().decode()}")
        self.pk_bytes, self.sk_bytes, self.proof_bytes = pk_bytes(kyber_k), sk_bytes(kyber_k), proof_bytes(kyber_k)
        self.wire_bytes = wire_bytes(kyber_k)

    def close(self):
        if self._h:
            self.lib.kosk_b200_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc < 0:
            raise KoskError(f"{what} failed ({rc}): {self.lib.kosk_b200_last_error().decode()}")
        return rc

    # ---- reference-shaped single calls (kosk.hpp:19-23) ----
    def verifiable_keygen(self, seed):
        pk, sk, pi = self.prove_batch(np.frombuffer(bytes(seed), dtype=np.uint8).reshape(1, 32))
        return bytes(pk[0]), bytes(sk[0]), bytes(pi[0])

    def kosk_verify(self, pi, pk):
        a = np.frombuffer(bytes(pi), dtype=np.uint8)
        b = np.frombuffer(bytes(pk), dtype=np.uint8)
        if a.size != self.proof_bytes or b.size != self.pk_bytes:
            raise KoskError("bad proof or pk length")
        return self._check(self.lib.kosk_b200_kosk_verify(self._h, _ptr(a), _ptr(b)), "kosk_verify") == 1

    # ---- struct-level API (reference mlwe_prover.hpp:77-99, mlwe_verifier.hpp:14-15, kosk.hpp:17-18) ----
    # Structs travel as byte images with the reference's layout (include/kosk_b200.h); the context holds the DRBG state.
    def rng_reset(self, seed):
        s = np.frombuffer(bytes(seed), dtype=np.uint8)
        self._check(self.lib.kosk_b200_rng_reset(self._h, _ptr(s)), "rng_reset")

    def rng_calls(self):
        return int(self.lib.kosk_b200_rng_calls(self._h))

    def verifiable_keygen_rng(self):
        """kyber_verifiable_keygen on the context DRBG (continues the call counter instead of starting a fresh seed)."""
        pk, sk, pi = np.empty(self.pk_bytes, np.uint8), np.empty(self.sk_bytes, np.uint8), np.empty(self.proof_bytes, np.uint8)
        self._check(self.lib.kosk_b200_verifiable_keygen_rng(self._h, _ptr(pk), _ptr(sk), _ptr(pi)), "verifiable_keygen_rng")
        return bytes(pk), bytes(sk), bytes(pi)

    def prepare_randomness(self):
        out = np.empty(self.lib.kosk_b200_randomness_bytes(self.k), np.uint8)
        self._check(self.lib.kosk_b200_prepare_randomness(self._h, _ptr(out)), "prepare_randomness")
        return out

    def prepare_range_proof(self):
        out = np.empty(self.lib.kosk_b200_range_proof_bytes(self.k), np.uint8)
        self._check(self.lib.kosk_b200_prepare_range_proof(self._h, _ptr(out)), "prepare_range_proof")
        return out

    def kyber_keygen(self):
        """kyber_keygen (kosk.cpp:4-70): returns (pk, sk, mlwe_inst image)."""
        pk, sk = np.empty(self.pk_bytes, np.uint8), np.empty(self.sk_bytes, np.uint8)
        inst = np.empty(self.lib.kosk_b200_inst_bytes(self.k), np.uint8)
        self._check(self.lib.kosk_b200_keygen(self._h, _ptr(pk), _ptr(sk), _ptr(inst)), "kyber_keygen")
        return bytes(pk), bytes(sk), inst

    def prove(self, inst, rand, eta):
        inst, rand, eta = (np.ascontiguousarray(x, dtype=np.uint8) for x in (inst, rand, eta))
        if (inst.size, rand.size, eta.size) != (self.lib.kosk_b200_inst_bytes(self.k), self.lib.kosk_b200_randomness_bytes(self.k), self.lib.kosk_b200_range_proof_bytes(self.k)):
            raise KoskError("bad struct image length")
        pi = np.empty(self.proof_bytes, np.uint8)
        self._check(self.lib.kosk_b200_prove(self._h, _ptr(pi), _ptr(inst), _ptr(rand), _ptr(eta)), "prove")
        return bytes(pi)

    def verify(self, pi, inst):
        a = np.frombuffer(bytes(pi), dtype=np.uint8)
        inst = np.ascontiguousarray(inst, dtype=np.uint8)
        if a.size != self.proof_bytes or inst.size != self.lib.kosk_b200_inst_bytes(self.k):
            raise KoskError("bad proof or mlwe_inst length")
        return self._check(self.lib.kosk_b200_verify(self._h, _ptr(a), _ptr(inst)), "verify") == 1

    # ---- Kyber KEM on the generated keys (reference kyber/kem.c:76-169) ----
    @property
    def ct_bytes(self):
        return self.lib.kosk_b200_ct_bytes(self.k)

    def kem_keypair_derand_batch(self, coins):
        """crypto_kem_keypair_derand (kem.c:23-33) for coins[n][64]: returns (pk[n], sk[n])."""
        coins = np.ascontiguousarray(coins, dtype=np.uint8).reshape(-1, 64)
        n = coins.shape[0]
        pk, sk = np.empty((n, self.pk_bytes), np.uint8), np.empty((n, self.sk_bytes), np.uint8)
        self._check(self.lib.kosk_b200_kem_keypair_derand_batch(self._h, n, _ptr(coins), _ptr(pk), _ptr(sk)), "kem_keypair_derand_batch")
        return pk, sk

    def crypto_kem_keypair(self):
        """crypto_kem_keypair (kem.c:47-58): the 64 coins are the next call of the context DRBG."""
        pk, sk = np.empty(self.pk_bytes, np.uint8), np.empty(self.sk_bytes, np.uint8)
        self._check(self.lib.kosk_b200_kem_keypair(self._h, _ptr(pk), _ptr(sk)), "kem_keypair")
        return bytes(pk), bytes(sk)

    def kem_enc_derand_batch(self, pk, coins):
        pk = np.ascontiguousarray(pk, dtype=np.uint8).reshape(-1, self.pk_bytes)
        coins = np.ascontiguousarray(coins, dtype=np.uint8).reshape(-1, 32)
        n = pk.shape[0]
        if coins.shape[0] != n:
            raise KoskError("pk / coins batch mismatch")
        ct, ss = np.empty((n, self.ct_bytes), np.uint8), np.empty((n, 32), np.uint8)
        self._check(self.lib.kosk_b200_kem_enc_derand_batch(self._h, n, _ptr(pk), _ptr(coins), _ptr(ct), _ptr(ss)), "kem_enc_derand_batch")
        return ct, ss

    def kem_dec_batch(self, ct, sk):
        ct = np.ascontiguousarray(ct, dtype=np.uint8).reshape(-1, self.ct_bytes)
        sk = np.ascontiguousarray(sk, dtype=np.uint8).reshape(-1, self.sk_bytes)
        n = ct.shape[0]
        if sk.shape[0] != n:
            raise KoskError("ct / sk batch mismatch")
        ss = np.empty((n, 32), np.uint8)
        self._check(self.lib.kosk_b200_kem_dec_batch(self._h, n, _ptr(ct), _ptr(sk), _ptr(ss)), "kem_dec_batch")
        return ss

    def crypto_kem_enc(self, pk):
        """crypto_kem_enc (kem.c:114-122): returns (ct, ss); the coins are the next call of the context DRBG."""
        pk = np.frombuffer(bytes(pk), dtype=np.uint8)
        ct, ss = np.empty(self.ct_bytes, np.uint8), np.empty(32, np.uint8)
        self._check(self.lib.kosk_b200_kem_enc(self._h, _ptr(ct), _ptr(ss), _ptr(pk)), "kem_enc")
        return bytes(ct), bytes(ss)

    def crypto_kem_dec(self, ct, sk):
        ct, sk = np.frombuffer(bytes(ct), dtype=np.uint8), np.frombuffer(bytes(sk), dtype=np.uint8)
        if ct.size != self.ct_bytes or sk.size != self.sk_bytes:
            raise KoskError("bad ct or sk length")
        ss = np.empty(32, np.uint8)
        self._check(self.lib.kosk_b200_kem_dec(self._h, _ptr(ss), _ptr(ct), _ptr(sk)), "kem_dec")
        return bytes(ss)

    def kem_enc_derand_batch_device(self, n, d_pk, d_coins, d_ct, d_ss, stream=0):
        vp = ctypes.c_void_p
        self._check(self.lib.kosk_b200_kem_enc_derand_batch_device(self._h, n, vp(d_pk), vp(d_coins), vp(d_ct), vp(d_ss), vp(stream)), "kem_enc_derand_batch_device")

    def kem_dec_batch_device(self, n, d_ct, d_sk, d_ss, stream=0):
        vp = ctypes.c_void_p
        self._check(self.lib.kosk_b200_kem_dec_batch_device(self._h, n, vp(d_ct), vp(d_sk), vp(d_ss), vp(stream)), "kem_dec_batch_device")

    # ---- batch, host buffers ----
    def prove_batch(self, seeds, out=None):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint8).reshape(-1, 32)
        n = seeds.shape[0]
        if out is None:
            out = (np.empty((n, self.pk_bytes), np.uint8), np.empty((n, self.sk_bytes), np.uint8), np.empty((n, self.proof_bytes), np.uint8))
        pk, sk, pi = out
        self._check(self.lib.kosk_b200_prove_batch(self._h, n, _ptr(seeds), _ptr(pk), _ptr(sk), _ptr(pi)), "prove_batch")
        return pk, sk, pi

    def verify_batch(self, pi, pk):
        pi = np.ascontiguousarray(pi, dtype=np.uint8).reshape(-1, self.proof_bytes)
        pk = np.ascontiguousarray(pk, dtype=np.uint8).reshape(-1, self.pk_bytes)
        n = pi.shape[0]
        if pk.shape[0] != n:
            raise KoskError("pi / pk batch mismatch")
        ok = np.zeros(n, np.uint8)
        self._check(self.lib.kosk_b200_verify_batch(self._h, n, _ptr(pi), _ptr(pk), _ptr(ok)), "verify_batch")
        return ok.astype(bool)

    # ---- compact wire format ----
    def set_wire(self, percent=100, threads=0):
        """Share (0..100 %) of the proofs that the host-buffer batch calls move over the link as 12-bit wire images; the rest as struct bytes."""
        self._check(self.lib.kosk_b200_set_wire(self._h, int(percent), int(threads)), "set_wire")

    def wire_info(self):
        m, t, sname = ctypes.c_int(), ctypes.c_int(), ctypes.c_char_p()
        self._check(self.lib.kosk_b200_wire_info(self._h, ctypes.byref(m), ctypes.byref(t), ctypes.byref(sname)), "wire_info")
        return {"mode": m.value, "threads": t.value, "simd": (sname.value or b"").decode()}

    def wire_stats(self, reset=True):
        out = np.zeros(4, np.uint64)
        self._check(self.lib.kosk_b200_wire_stats(self._h, _ptr(out), 1 if reset else 0), "wire_stats")
        return {"gate_wait_ms": float(out[0]) / 1e6, "slices": int(out[1]), "worker_ms_total": float(out[2]) / 1e6, "proofs": int(out[3])}

    def prove_batch_packed(self, seeds, out=None):
        """prove_batch whose proofs come back as compact wire images: returns (pk, sk, wire[n][wire_bytes])."""
        seeds = np.ascontiguousarray(seeds, dtype=np.uint8).reshape(-1, 32)
        n = seeds.shape[0]
        if out is None:
            out = (np.empty((n, self.pk_bytes), np.uint8), np.empty((n, self.sk_bytes), np.uint8), np.empty((n, self.wire_bytes), np.uint8))
        pk, sk, wire = out
        self._check(self.lib.kosk_b200_prove_batch_packed(self._h, n, _ptr(seeds), _ptr(pk), _ptr(sk), _ptr(wire)), "prove_batch_packed")
        return pk, sk, wire

    def verify_batch_packed(self, wire, pk):
        wire = np.ascontiguousarray(wire, dtype=np.uint8).reshape(-1, self.wire_bytes)
        pk = np.ascontiguousarray(pk, dtype=np.uint8).reshape(-1, self.pk_bytes)
        n = wire.shape[0]
        if pk.shape[0] != n:
            raise KoskError("wire / pk batch mismatch")
        ok = np.zeros(n, np.uint8)
        self._check(self.lib.kosk_b200_verify_batch_packed(self._h, n, _ptr(wire), _ptr(pk), _ptr(ok)), "verify_batch_packed")
        return ok.astype(bool)

    def wire_pack_device(self, n, d_pi, d_wire, stream=0):
        vp = ctypes.c_void_p
        self._check(self.lib.kosk_b200_wire_pack_device(self._h, n, vp(d_pi), vp(d_wire), vp(stream)), "wire_pack_device")

    def wire_unpack_device(self, n, d_wire, d_pi, stream=0):
        vp = ctypes.c_void_p
        self._check(self.lib.kosk_b200_wire_unpack_device(self._h, n, vp(d_wire), vp(d_pi), vp(stream)), "wire_unpack_device")

    # ---- batch, device pointers (integers), asynchronous on `stream` ----
    def prove_batch_device(self, n, d_seeds, d_pk, d_sk, d_pi, stream=0):
        vp = ctypes.c_void_p
        self._check(self.lib.kosk_b200_prove_batch_device(self._h, n, vp(d_seeds), vp(d_pk), vp(d_sk), vp(d_pi), vp(stream)), "prove_batch_device")

    def verify_batch_device(self, n, d_pi, d_pk, d_ok, stream=0):
        vp = ctypes.c_void_p
        self._check(self.lib.kosk_b200_verify_batch_device(self._h, n, vp(d_pi), vp(d_pk), vp(d_ok), vp(stream)), "verify_batch_device")

    def share_eval_device(self, n, d_y, d_planes, stream=0):
        vp = ctypes.c_void_p
        self._check(self.lib.kosk_b200_share_eval_device(self._h, n, vp(d_y), vp(d_planes), vp(stream)), "share_eval_device")

    # ---- components ----
    def share_eval(self, y):
        y = np.ascontiguousarray(y, dtype=np.uint16).reshape(-1, 407)
        out = np.empty((y.shape[0], 1454), np.uint16)
        self._check(self.lib.kosk_b200_share_eval(self._h, y.shape[0], _ptr(y), _ptr(out)), "share_eval")
        return out

    def recon_rows(self, shares, degree2=False):
        """recon_secrets_ddeg / _2ddeg (ss.cpp:37-73) on rows of 407 / 813 party shares -> 256 secrets each."""
        nn = 813 if degree2 else 407
        shares = np.ascontiguousarray(shares, dtype=np.uint16).reshape(-1, nn)
        out = np.empty((shares.shape[0], 256), np.uint16)
        self._check(self.lib.kosk_b200_recon_rows(self._h, 1 if degree2 else 0, shares.shape[0], _ptr(shares), _ptr(out)), "recon_rows")
        return out

    def interp_rows(self, opened, shares, degree2=False):
        """The verifier's interpolation through the first 407 / 813 rest-party nodes, evaluated at 0..406 / 0..255."""
        nn, nt = (813, 256) if degree2 else (407, 407)
        opened = np.ascontiguousarray(opened, dtype=np.uint16).reshape(150)
        shares = np.ascontiguousarray(shares, dtype=np.uint16).reshape(-1, nn)
        out = np.empty((shares.shape[0], nt), np.uint16)
        self._check(self.lib.kosk_b200_interp_rows(self._h, 1 if degree2 else 0, _ptr(opened), shares.shape[0], _ptr(shares), _ptr(out)), "interp_rows")
        return out

    def sha3_256_rows(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.uint8)
        n, ln = rows.shape
        out = np.empty((n, 32), np.uint8)
        self._check(self.lib.kosk_b200_sha3_256_rows(self._h, n, ln, _ptr(rows), _ptr(out)), "sha3_256_rows")
        return out

    def ntt_rows(self, a):
        a = np.array(a, dtype=np.uint16).reshape(-1, 256)
        self._check(self.lib.kosk_b200_ntt_rows(self._h, a.shape[0], _ptr(a)), "ntt_rows")
        return a

    def kernel_launches(self):
        return int(self.lib.kosk_b200_kernel_launches(self._h))

    def debug_fetch(self, what, nbytes, dtype=np.uint8):
        out = np.zeros(nbytes, np.uint8)
        self._check(self.lib.kosk_b200_debug_fetch(self._h, what.encode(), _ptr(out), nbytes), "debug_fetch")
        return out.view(dtype)

    PHASES = ["keygen", "expand", "share1", "commit", "fs1", "eval", "open", "share2", "view", "fs2", "assemble", "verify"]

    # ---- offline / online split ----
    def pool_create(self, seeds):
        """Run the key-independent preprocessing for these seeds and keep it on the device."""
        seeds = np.ascontiguousarray(seeds, dtype=np.uint8).reshape(-1, 32)
        h = ctypes.c_void_p()
        self._check(self.lib.kosk_b200_pool_create(self._h, seeds.shape[0], _ptr(seeds), ctypes.byref(h)), "pool_create")
        return KoskPool(self, h, seeds.shape[0])

    def pool_import(self, image):
        """Rebuild a preprocessing pool from an image written by KoskPool.export()."""
        image = np.ascontiguousarray(image, dtype=np.uint8)
        h = ctypes.c_void_p()
        self._check(self.lib.kosk_b200_pool_import(self._h, _ptr(image), image.size, ctypes.byref(h)), "pool_import")
        n = int(np.frombuffer(image[16:20].tobytes(), np.uint32)[0])
        return KoskPool(self, h, n)

    def set_strict(self, on=True):
        """Hardened verifier; default off = the reference's accept set."""
        self._check(self.lib.kosk_b200_set_strict(self._h, 1 if on else 0), "set_strict")

    def set_profiling(self, on=True):
        self._check(self.lib.kosk_b200_set_profiling(self._h, 1 if on else 0), "set_profiling")

    def phase_times(self, reset=True):
        ms = np.zeros(len(self.PHASES), np.float64)
        calls = np.zeros(len(self.PHASES), np.uint64)
        self._check(self.lib.kosk_b200_phase_times(self._h, _ptr(ms), _ptr(calls), len(self.PHASES), 1 if reset else 0), "phase_times")
        return {n: (float(m), int(c)) for n, m, c in zip(self.PHASES, ms, calls)}

    def int_peak(self):
        out = np.zeros(8, np.float64)
        self._check(self.lib.kosk_b200_int_peak(self._h, _ptr(out)), "int_peak")
        return {"imad": float(out[0]), "lop3": float(out[1]), "shf": float(out[2]), "imma_int8_mac": float(out[3]), "imad_hi": float(out[4])}

    def debug_trace(self, max_triples=4096):
        out = np.zeros(3 * max_triples, np.float64)
        n = self._check(self.lib.kosk_b200_debug_trace(self._h, _ptr(out), max_triples), "debug_trace")
        names = self.PHASES + ["end"]
        return [(int(out[3 * i]), names[int(out[3 * i + 1])] if out[3 * i + 1] >= 0 else "end", float(out[3 * i + 2])) for i in range(n)]

    def sync(self):
        self._check(self.lib.kosk_b200_sync(self._h), "sync")


class KoskPool:
    """Preprocessed (offline) material of n proofs; prove() runs only the online phase."""

    def __init__(self, ctx, handle, n):
        self.ctx, self._h, self.n = ctx, handle, n

    def prove(self):
        c = self.ctx
        pk, sk, pi = np.empty((self.n, c.pk_bytes), np.uint8), np.empty((self.n, c.sk_bytes), np.uint8), np.empty((self.n, c.proof_bytes), np.uint8)
        c._check(c.lib.kosk_b200_pool_prove(self._h, _ptr(pk), _ptr(sk), _ptr(pi)), "pool_prove")
        return pk, sk, pi

    def export(self):
        """Serialise the preprocessed material (contains the seeds: keep it secret)."""
        c = self.ctx
        out = np.empty(c.lib.kosk_b200_pool_bytes(self._h), np.uint8)
        c._check(c.lib.kosk_b200_pool_export(self._h, _ptr(out), out.size), "pool_export")
        return out

    def close(self):
        if self._h:
            self.ctx.lib.kosk_b200_pool_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class kyber_keypair:
    """Mirror of the reference's `kyber_keypair` (kosk.hpp:13-16)."""

    def __init__(self, pk=b"", sk=b""):
        self.pk, self.sk = pk, sk


_default_ctx = {}


def _ctx(k):
    if k not in _default_ctx:
        _default_ctx[k] = KoskContext(k)
    return _default_ctx[k]


def kyber_verifiable_keygen(kyber_k=2, seed=None):
    """kyber_verifiable_keygen (kosk.cpp:72-86): returns (kyber_keypair, pi).  `seed` replaces the global
    randombytes() stream (KOSK counter-mode DRBG); if None, 32 bytes from the OS RNG are used, as the
    reference's randombytes does (kyber/randombytes.c:43-57)."""
    seed = os.urandom(32) if seed is None else bytes(seed)
    pk, sk, pi = _ctx(kyber_k).verifiable_keygen(seed)
    return kyber_keypair(pk, sk), pi


def kyber_kosk_verify(pi, pk, kyber_k=2):
    """kyber_kosk_verify (kosk.cpp:88-117)."""
    return _ctx(kyber_k).kosk_verify(pi, pk)
