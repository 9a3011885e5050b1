// wire_host.h -- host side of the compact wire format (kosk_common.cuh, WireLayout): the 12-bit <-> u16 codec of proof
// bytes and the worker pool that expands packed proofs into the caller's buffers as their D2H slices land.
// No field or hash arithmetic happens here: this is a byte-format codec (the reference's encode/decode_mpcith_proof,
// mlwe_prover.cpp:540-630, are memcpy's of the same fields).  Compiled by g++ (SIMD intrinsics), linked into libkosk_b200.so.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <atomic>

namespace kosk {

// one proof: wire -> struct mpcith_proof bytes
void wire_unpack_proof(int k, const uint8_t *wire, uint8_t *pi);
// one proof: struct mpcith_proof bytes -> wire; returns 0, or 1 if some u16 of the proof is >= 4096 (not representable; wire is then undefined)
int wire_pack_proof(int k, const uint8_t *pi, uint8_t *wire);
const char *wire_simd_name();          // "avx512vbmi" | "avx2" | "scalar": the code path selected for this CPU

// Worker pool.  A submission converts `n` consecutive proofs in jobs of `per_job`.  It may carry a gate (an opaque event pointer): ONE
// gate thread waits on the gates in submission order (with the wait function given at creation: it polls the CUDA event on the .cu
// side) and only then hands the jobs to the workers, so a slice is expanded while later slices are still on the link and no worker
// ever blocks on the device.  `ctr` (may be null) counts the submission's unfinished jobs: the submitter waits on it before it
// reuses the slice of the staging buffer the jobs read.
struct WirePool;
typedef int (*wire_wait_fn)(void *gate);
WirePool *wire_pool_create(int threads, wire_wait_fn wait);
void wire_pool_destroy(WirePool *p);
int wire_pool_threads(const WirePool *p);
void wire_pool_stats(WirePool *p, uint64_t out[4], int reset);   // ns the gate thread waited, gates, ns the workers spent converting, jobs
// kind 0 = unpack (wire -> pi), 1 = pack (pi -> wire).  `flag` (may be null) is set to 1 by a pack job that met an unrepresentable
// element or when the gate wait failed.
void wire_pool_submit(WirePool *p, void *gate, int kind, int k, size_t n, size_t per_job, const uint8_t *src, uint8_t *dst, std::atomic<int> *ctr, volatile int *flag);
void wire_pool_wait_counter(WirePool *p, std::atomic<int> *ctr);   // until the counter is back to zero
void wire_pool_wait_all(WirePool *p);
void wire_pool_trace(WirePool *p, int kind, uint64_t id);        // measurement only (KOSK_B200_WIRE_TRACE)
// synchronous parallel-for over n proofs on the pool's threads (the caller blocks); returns the OR of the jobs' flags
int wire_pool_run(WirePool *p, int kind, int k, size_t n, const uint8_t *src, uint8_t *dst);

}  // namespace kosk
