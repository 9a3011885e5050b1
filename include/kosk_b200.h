/* kosk_b200.h -- C ABI of the B200-native KOSK prover/verifier core (libkosk_b200.so).
 *
 * Drop-in boundary for the reference's KOSK API (reference kosk.hpp:13-24):
 *   void kyber_verifiable_keygen(kyber_keypair *keypair, uint8_t *pi)   kosk.cpp:72-86
 *   bool kyber_kosk_verify(const uint8_t *pi, const uint8_t *pk)        kosk.cpp:88-117
 * The proof is the raw byte image of struct mpcith_proof (mlwe_prover.hpp:57-75, encode_mpcith_proof
 * mlwe_prover.cpp:540-543); pk/sk are the Kyber byte strings of kyber_keygen (kosk.cpp:57-69).
 * The reference is one process / one K per binary and pulls randomness from a global randombytes()
 * (kyber/randombytes.c:43-57).  This ABI adds what a device library needs and the reference lacks:
 * a context (device, tables, scratch), KYBER_K as a run-time argument, batch entry points, and an
 * explicit 32-byte seed per proof in place of the global RNG:
 *
 *   KOSK counter-mode DRBG: the c-th randombytes(out, len) call made by kyber_verifiable_keygen
 *   (c = 0, 1, ... in the reference's call order) returns
 *   SHAKE256(seed[32] || LE32(c))[0:len].
 *
 * With that definition of randombytes linked into the reference, pk, sk and proof bytes are
 * bit-identical to the reference's.  All functions return 0 on success, a negative KOSK_E_* code on
 * error (kosk_b200_last_error() gives the message); there is no CPU fallback: without a CUDA device
 * kosk_b200_create fails with KOSK_E_CUDA.
 */
#ifndef KOSK_B200_H
#define KOSK_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct kosk_b200_ctx kosk_b200_ctx;

enum { KOSK_OK = 0, KOSK_E_ARG = -1, KOSK_E_CUDA = -2, KOSK_E_NOMEM = -3, KOSK_E_UNSUPPORTED = -4 };

/* sizes: KYBER_PUBLICKEYBYTES / KYBER_SECRETKEYBYTES (kyber/params.h:50-53), MPCITH_PROOF_SIZE (mlwe_prover.hpp:30) */
size_t kosk_b200_pk_bytes(int kyber_k);
size_t kosk_b200_sk_bytes(int kyber_k);
size_t kosk_b200_proof_bytes(int kyber_k);
const char *kosk_b200_last_error(void);
const char *kosk_b200_version(void);

/* Threading: a context serialises its callers (every entry point takes the context's lock); use one context per thread for
 * concurrency.  A context cannot be destroyed while preprocessing pools created from it are alive (kosk_b200_destroy then keeps
 * the context and records an error).
 * Context: one per (device, KYBER_K).  max_chunk = proofs processed per kernel wave (scratch is sized for
 * it; 0 = default).  Replaces the reference's compile-time -DKYBER_K (params.hpp:8-10). */
int kosk_b200_create(kosk_b200_ctx **ctx, int kyber_k, int device, int max_chunk);
/* Same, with the number of pipeline lanes made explicit (0 = default 2).  A batch is split into sub-batches of at
 * most max_chunk proofs that alternate over `lanes` CUDA streams, each with its own scratch: kernels of consecutive
 * sub-batches run back to back, host copies of one sub-batch overlap the kernels of the next. */
int kosk_b200_create_ex(kosk_b200_ctx **ctx, int kyber_k, int device, int max_chunk, int lanes, int flags);
/* flags.  KOSK_F_TENSOR (EXPERIMENTAL, off by default; also KOSK_B200_TENSOR=1 with kosk_b200_create): run the prover's
 * share evaluation on the int8 tensor-core path (limb-split residues, mma.sync) instead of the INT32 pipe.  Same bytes. */
enum { KOSK_F_TENSOR = 1 };
int kosk_b200_lanes(const kosk_b200_ctx *ctx);
void kosk_b200_destroy(kosk_b200_ctx *ctx);

/* kyber_verifiable_keygen (kosk.hpp:19-20) with the RNG seed made explicit.  Host buffers. */
int kosk_b200_verifiable_keygen(kosk_b200_ctx *ctx, const uint8_t seed[32], uint8_t *pk, uint8_t *sk, uint8_t *pi);
/* kyber_kosk_verify (kosk.hpp:22-23): returns 1 = accept, 0 = reject, <0 = error.  Host buffers. */
int kosk_b200_kosk_verify(kosk_b200_ctx *ctx, const uint8_t *pi, const uint8_t *pk);

/* Hardened verifier, a deliberate, switchable deviation from the reference's accept set; default off.
 * When on, kosk_b200_*verify* additionally rejects (i) any field element >= 3329 anywhere in the proof and (ii) rest-party
 * shares of t / s_eta / e_eta that do not lie on the interpolated sharing polynomial (the reference reads only the first
 * 407 of them).  Proofs produced by kosk_b200_*keygen* / the reference prover are unaffected. */
int kosk_b200_set_strict(kosk_b200_ctx *ctx, int on);

/* Alignment: device-resident seeds (d_seeds) must be 8-byte aligned and device-resident proofs (d_pi) 4-byte aligned (cudaMalloc'd
 * buffers are); host buffers may have any alignment.
 * Batch mode, host buffers, densely packed: seeds[n][32], pk[n][pk_bytes], sk[n][sk_bytes], pi[n][proof_bytes].
 * Host<->device copies happen inside (pinned buffers are used asynchronously). */
int kosk_b200_prove_batch(kosk_b200_ctx *ctx, size_t n, const uint8_t *seeds, uint8_t *pk, uint8_t *sk, uint8_t *pi);
int kosk_b200_verify_batch(kosk_b200_ctx *ctx, size_t n, const uint8_t *pi, const uint8_t *pk, uint8_t *ok);
/* Asynchronous form of prove_batch: enqueues H2D of the seeds, the kernels and the D2H of pk/sk/pi on the context's lanes
 * and returns; kosk_b200_sync() waits for everything enqueued.  Consecutive calls (and the sub-batches of one call)
 * alternate over the lanes: their kernels run back to back while the D2H copy of one overlaps the kernels of the next.
 * Host buffers must stay valid (and should be pinned) until kosk_b200_sync returns. */
int kosk_b200_prove_batch_async(kosk_b200_ctx *ctx, size_t n, const uint8_t *seeds, uint8_t *pk, uint8_t *sk, uint8_t *pi);
/* Asynchronous form of verify_batch: the H2D of one call's proofs overlaps the kernels of the previous call on the other lane;
 * ok[] is valid after kosk_b200_sync().  pi / pk / ok must stay valid until then (pi may be reused as soon as the call returns
 * when the wire mode packed it into the context's staging buffer, but do not rely on it). */
int kosk_b200_verify_batch_async(kosk_b200_ctx *ctx, size_t n, const uint8_t *pi, const uint8_t *pk, uint8_t *ok);

/* Compact wire format.  struct mpcith_proof (mlwe_prover.hpp:57-75) stores every GF(3329) element and every
 * party index as a uint16_t; encode_mpcith_proof (mlwe_prover.cpp:540-543) is a memcpy of it.  The wire image holds the same
 * fields in the same order with 12 bits per u16 (two elements per three bytes, little-endian as in kyber/poly.c:124-139) and
 * the two digest arrays (Tcomm, comm) verbatim, every segment 16-byte aligned:
 *     wire = pack12(f_shares .. gamma_shares) | pad | Tcomm | pack12(I .. u_e_2ddeg_shares) | pad | comm
 * 519 136 / 531 616 / 578 992 bytes for Kyber512 / 768 / 1024 instead of 664 340 / 680 980 / 744 148.  Every proof whose u16
 * fields are all < 4096 has a wire image (every proof a prover emits does); decoding is the exact inverse.
 *
 * The host-buffer batch calls above move proofs over PCIe, which bounds them (one B200 proves faster than a Gen5 x16 link
 * carries 664 KB proofs).  In wire mode (default 100; KOSK_B200_WIRE=<percent> or kosk_b200_set_wire) `percent` of the proofs
 * cross the link as wire images: prove packs on the device, copies the compact bytes in slices into a pinned staging buffer and
 * expands them into the caller's pi[] (reference layout, unchanged) on `threads` host worker threads while later slices are
 * still on the link; verify_batch packs on those threads and unpacks on the device.  The remaining proofs cross it as struct
 * bytes straight between the caller's buffer and the device, which costs the host nothing: the link favours 100, a host short
 * of memory bandwidth something lower (bench.py calibrates it), 0 is the round-1 behaviour.  The caller-visible bytes are the
 * same for every setting.  A verify sub-batch containing a proof with a u16 >= 4096 travels as struct bytes (same verdicts).
 * Calls with fewer than 8 proofs always use struct bytes (latency).
 * The *_packed calls hand the compact bytes to / take them from the caller: wire[n][wire_bytes]. */
size_t kosk_b200_wire_bytes(int kyber_k);
int kosk_b200_set_wire(kosk_b200_ctx *ctx, int percent /* 0..100 */, int threads /* 0 = keep / default (KOSK_B200_WIRE_THREADS) */);
int kosk_b200_wire_info(const kosk_b200_ctx *ctx, int *mode, int *threads, const char **simd);
/* instrumentation: out[4] = {ns the gate thread waited for D2H slices, slices, ns summed over the workers spent converting, proofs converted} */
int kosk_b200_wire_stats(kosk_b200_ctx *ctx, uint64_t *out, int reset);
int kosk_b200_prove_batch_packed(kosk_b200_ctx *ctx, size_t n, const uint8_t *seeds, uint8_t *pk, uint8_t *sk, uint8_t *wire);
int kosk_b200_prove_batch_packed_async(kosk_b200_ctx *ctx, size_t n, const uint8_t *seeds, uint8_t *pk, uint8_t *sk, uint8_t *wire);
int kosk_b200_verify_batch_packed(kosk_b200_ctx *ctx, size_t n, const uint8_t *wire, const uint8_t *pk, uint8_t *ok);
int kosk_b200_verify_batch_packed_async(kosk_b200_ctx *ctx, size_t n, const uint8_t *wire, const uint8_t *pk, uint8_t *ok);
/* device-resident conversion, enqueued on `stream`: d_pi 4-byte aligned, d_wire 16-byte aligned */
int kosk_b200_wire_pack_device(kosk_b200_ctx *ctx, size_t n, const uint8_t *d_pi, uint8_t *d_wire, void *stream);
int kosk_b200_wire_unpack_device(kosk_b200_ctx *ctx, size_t n, const uint8_t *d_wire, uint8_t *d_pi, void *stream);
/* host codec (no context, no device; a byte-format conversion like encode/decode_mpcith_proof, mlwe_prover.cpp:540-630):
 * wire_pack fails with KOSK_E_ARG if some u16 of a proof is >= 4096.  threads <= 1: on the calling thread. */
int kosk_b200_wire_pack(int kyber_k, size_t n, const uint8_t *pi, uint8_t *wire, int threads);
int kosk_b200_wire_unpack(int kyber_k, size_t n, const uint8_t *wire, uint8_t *pi, int threads);
const char *kosk_b200_wire_simd(void);      /* "avx512vbmi" | "avx2" | "scalar" */

/* Batch mode, device-resident buffers (same packing), enqueued on `stream` (a cudaStream_t, NULL = default).
 * Asynchronous: the caller synchronises the stream. */
int kosk_b200_prove_batch_device(kosk_b200_ctx *ctx, size_t n, const uint8_t *d_seeds, uint8_t *d_pk, uint8_t *d_sk,
                                 uint8_t *d_pi, void *stream);
int kosk_b200_verify_batch_device(kosk_b200_ctx *ctx, size_t n, const uint8_t *d_pi, const uint8_t *d_pk,
                                  uint8_t *d_ok, void *stream);

/* Offline / online split.  prepare_randomness + prepare_range_proof (mlwe_prover.cpp:4-59) do not depend
 * on the key; the reference times them separately (main.cpp:18-28) but has no way to keep them (its (de)serialiser,
 * mlwe_prover.cpp:61-79, is unused and drops the range-proof half).  pool_create runs that phase for n proofs and keeps
 * the material (the f / NTT_f / eta sharings) on the device; pool_prove runs only keygen + prove() + encode.  For the
 * same seeds the outputs are bit-identical to kosk_b200_prove_batch.  Host buffers, synchronous. */
typedef struct kosk_b200_pool kosk_b200_pool;
int kosk_b200_pool_create(kosk_b200_ctx *ctx, size_t n, const uint8_t *seeds, kosk_b200_pool **pool);
int kosk_b200_pool_prove(kosk_b200_pool *pool, uint8_t *pk, uint8_t *sk, uint8_t *pi);
void kosk_b200_pool_destroy(kosk_b200_pool *pool);
/* Serialiser of the preprocessed material (the working form of mlwe_prover.cpp:61-79): export writes kosk_b200_pool_bytes(pool)
 * bytes (header | seeds | sharing inputs | evaluated f / NTT_f / eta sharings, about 0.7 MB per Kyber512 proof); import rebuilds a
 * pool on any context of the same KYBER_K, and pool_prove on it gives the same bytes as on the exporting pool.  The image contains
 * the seeds: it is as secret as the keys it will produce. */
size_t kosk_b200_pool_bytes(const kosk_b200_pool *pool);
int kosk_b200_pool_export(kosk_b200_pool *pool, void *image, size_t bytes);
int kosk_b200_pool_import(kosk_b200_ctx *ctx, const void *image, size_t bytes, kosk_b200_pool **pool);

/* Struct-level API: the reference's lower-level entry points, used directly by main.cpp:16-59, on byte
 * images of the reference's own structs (x86-64 layout, no padding):
 *   inst  = mlwe_inst          (mlwe_prover.hpp:34-37): int16 A[K][K][256] | t[K][256] | s[K][256] | e[K][256]
 *   rand  = mpcith_randomness  (mlwe_prover.hpp:39-44): u16 f[F][256] | NTT_f[F][256] | share_vec f_shares[F] | NTT_f_shares[F]
 *   eta   = mpcith_range_proof (mlwe_prover.hpp:46-49): share_vec s_eta_shares[K][E] | e_eta_shares[K][E]
 *   share_vec (ss.hpp:33-37)   = size_t len | u16 share_x[1454] | u16 share_y[1454]; len is written as 0 (the reference leaves
 *                                it uninitialised), share_x[p] = p + 256
 *   pi    = mpcith_proof       (mlwe_prover.hpp:57-75) = the proof bytes; encode/decode_mpcith_proof are memcpy
 * Randomness: the reference draws from one global randombytes(); here the context holds the KOSK counter-mode DRBG state
 * (seed, number of calls made so far) and every function below consumes exactly the calls its reference counterpart makes
 * (prepare_randomness 3F, prepare_range_proof 2KE, kyber_keygen 1, prove 3K + 4*eta*K), so any call order gives the bytes the
 * reference gives under the same randombytes definition.  kyber_verifiable_keygen's own order is keygen, prepare_randomness,
 * prepare_range_proof, prove (kosk.cpp:72-86).  prove() assumes rand / eta were produced by the two prepare functions (the
 * clear f / NTT_f vectors are consistent with their sharings, the eta sharings hide the constants -eta..eta).
 * Single instance per call, host buffers, synchronous. */
size_t kosk_b200_inst_bytes(int kyber_k);
size_t kosk_b200_randomness_bytes(int kyber_k);
size_t kosk_b200_range_proof_bytes(int kyber_k);
int kosk_b200_rng_reset(kosk_b200_ctx *ctx, const uint8_t seed[32]);      /* seed the DRBG, call counter = 0 */
uint32_t kosk_b200_rng_calls(const kosk_b200_ctx *ctx);
int kosk_b200_verifiable_keygen_rng(kosk_b200_ctx *ctx, uint8_t *pk, uint8_t *sk, uint8_t *pi);   /* kosk.cpp:72-86 on the context DRBG */
int kosk_b200_prepare_randomness(kosk_b200_ctx *ctx, void *rand);          /* mlwe_prover.cpp:4-39 */
int kosk_b200_prepare_range_proof(kosk_b200_ctx *ctx, void *eta);          /* mlwe_prover.cpp:41-59 */
int kosk_b200_keygen(kosk_b200_ctx *ctx, uint8_t *pk, uint8_t *sk, void *inst /* may be NULL */);   /* kosk.cpp:4-70 */
int kosk_b200_prove(kosk_b200_ctx *ctx, uint8_t *pi, const void *inst, const void *rand, const void *eta);   /* mlwe_prover.cpp:81-538 */
int kosk_b200_verify(kosk_b200_ctx *ctx, const uint8_t *pi, const void *inst);   /* mlwe_verifier.cpp:4-686; 1 accept, 0 reject, <0 error */

/* Kyber KEM on the keys this library generates (the step after keygen in main.cpp:98-113).
 * ct_bytes = KYBER_CIPHERTEXTBYTES (kyber/params.h:53): 768 / 1088 / 1568; shared secrets are KYBER_SSBYTES = 32 bytes.
 *   kem_enc_derand_batch: crypto_kem_enc_derand (kyber/kem.c:76-97) for n independent (pk, coins) pairs, coins[n][32]
 *   kem_dec_batch:        crypto_kem_dec (kyber/kem.c:139-169), implicit rejection included
 *   kem_enc / kem_dec:    the reference's single calls; kem_enc draws its 32 coins as the next randombytes() call of the
 *                         context DRBG (kosk_b200_rng_reset), as crypto_kem_enc does from the global RNG (kem.c:114-122)
 * Host buffers, synchronous; the *_device forms take device pointers and enqueue on `stream`. */
size_t kosk_b200_ct_bytes(int kyber_k);
/* crypto_kem_keypair_derand (kyber/kem.c:23-33) for n coins[n][64] (d | z), and crypto_kem_keypair (kem.c:47-58) with the 64 coins drawn
 * as the next call of the context DRBG.  Same keys as kyber_keygen (kosk.cpp:4-70) for the same d, except sk's last 32 bytes (z). */
int kosk_b200_kem_keypair_derand_batch(kosk_b200_ctx *ctx, size_t n, const uint8_t *coins, uint8_t *pk, uint8_t *sk);
int kosk_b200_kem_keypair(kosk_b200_ctx *ctx, uint8_t *pk, uint8_t *sk);
int kosk_b200_kem_enc_derand_batch(kosk_b200_ctx *ctx, size_t n, const uint8_t *pk, const uint8_t *coins, uint8_t *ct, uint8_t *ss);
int kosk_b200_kem_dec_batch(kosk_b200_ctx *ctx, size_t n, const uint8_t *ct, const uint8_t *sk, uint8_t *ss);
int kosk_b200_kem_enc_derand_batch_device(kosk_b200_ctx *ctx, size_t n, const uint8_t *d_pk, const uint8_t *d_coins, uint8_t *d_ct, uint8_t *d_ss, void *stream);
int kosk_b200_kem_dec_batch_device(kosk_b200_ctx *ctx, size_t n, const uint8_t *d_ct, const uint8_t *d_sk, uint8_t *d_ss, void *stream);
int kosk_b200_kem_enc(kosk_b200_ctx *ctx, uint8_t *ct, uint8_t *ss, const uint8_t *pk);
int kosk_b200_kem_dec(kosk_b200_ctx *ctx, uint8_t *ss, const uint8_t *ct, const uint8_t *sk);

/* Components (BASELINE config 5 microbenches, kernel-level parity tests).
 * share_eval: recompute_share_secrets_ddeg (ss.cpp:76-99) on n rows: y[n][407] -> shares[n][1454]. Host buffers.
 * sha3_256_rows: n independent SHA3-256 over rows of `len` bytes (len even): in[n][len] -> out[n][32].
 * ntt_rows: poly_ntt (kyber/poly.c:261-265) on canonical residues, in place: a[n][256]. */
int kosk_b200_share_eval(kosk_b200_ctx *ctx, size_t n, const uint16_t *y, uint16_t *shares);
int kosk_b200_sha3_256_rows(kosk_b200_ctx *ctx, size_t n, size_t len, const uint8_t *in, uint8_t *out);
int kosk_b200_ntt_rows(kosk_b200_ctx *ctx, size_t n, uint16_t *a);
/* recon_rows: recon_secrets_ddeg (degree2 = 0, ss.cpp:37-54: shares[n][407] of parties 0..406) or recon_secrets_2ddeg (degree2 = 1,
 * ss.cpp:56-73: shares[n][813] of parties 0..812) -> secrets[n][256].
 * interp_rows: the verifier's interpolation through the rest-party nodes (NTL interpolate + eval, mlwe_verifier.cpp:188-224, :510-543):
 * opened[150] = the opened parties; row r holds the shares of the first 407 (813) rest parties in ascending party order; out[n][407]
 * = the interpolant at x = 0..406 (degree2 = 0) or out[n][256] = at x = 0..255 (degree2 = 1).  n <= 4096. */
int kosk_b200_recon_rows(kosk_b200_ctx *ctx, int degree2, size_t n, const uint16_t *shares, uint16_t *secrets);
int kosk_b200_interp_rows(kosk_b200_ctx *ctx, int degree2, const uint16_t *opened, size_t n, const uint16_t *shares, uint16_t *out);
/* device-resident share_eval for benchmarking: d_y[n][416] (zero padded rows), d_planes[n][1456] */
int kosk_b200_share_eval_device(kosk_b200_ctx *ctx, size_t n, const uint16_t *d_y, uint16_t *d_planes, void *stream);

/* Introspection for tests/bench: number of kernels launched by this context so far; copy of an internal
 * buffer of the last prove chunk ("alpha_pow", "I", "tcomm", "views", "Y", "planes") to host memory. */
uint64_t kosk_b200_kernel_launches(const kosk_b200_ctx *ctx);
int kosk_b200_debug_fetch(kosk_b200_ctx *ctx, const char *what, void *out, size_t bytes);
int kosk_b200_sync(kosk_b200_ctx *ctx);
/* debug: (lane, phase, ms since the batch started) triples of the last profiled device-buffer batch; returns the count */
int kosk_b200_debug_trace(kosk_b200_ctx *ctx, double *out, int max_triples);

/* Measurement support.  With profiling on, every prove chunk records CUDA events on its launching stream at the
 * phase boundaries; kosk_b200_phase_times() synchronises and returns accumulated milliseconds and call counts for
 * the KOSK_PH_* phases (order: keygen, expand, share1, commit, fs1, eval, open, share2, view, fs2, assemble, verify).
 * kosk_b200_int_peak() runs issue-rate microbenchmarks and returns, in ops_per_s[5], thread-level ops/s for IMAD, LOP3
 * and SHF (the SM integer-pipe roofline denominators), int8 MAC/s of warp-level mma.sync (for the opt-in tensor path) and
 * thread-level ops/s of IMAD.HI (the high half of a 32 x 32 product, used by the Barrett / Shoup reductions). */
#define KOSK_B200_NPHASE 12
int kosk_b200_set_profiling(kosk_b200_ctx *ctx, int on);
int kosk_b200_phase_times(kosk_b200_ctx *ctx, double *ms, uint64_t *calls, int n, int reset);
int kosk_b200_int_peak(kosk_b200_ctx *ctx, double *ops_per_s);

#ifdef __cplusplus
}
#endif
#endif
