// Experiment: can the FMA-heavy pipe (IMAD register-tile MACs, as in k_gf_gemm) and the ALU pipe (LOP3/SHF Keccak rounds, as in
// k_hash_records) be kept busy at the same time on one SM sub-partition?  MODE 0: all warps run the MAC tile; 1: all warps run
// Keccak-f; 2: even warp pairs MAC, odd pairs Keccak (same work per warp as in modes 0/1).  If the pipes overlap, t2 ~ max(t0,t1)/2+;
// if issue / register ports serialise them, t2 ~ (t0 + t1) / 2.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o dual_pipe_bench dual_pipe_bench.cu
#include <cstdint>
#include <cstdio>
#include "../../mpcith_kyber_kosk_b200/csrc/keccak.cuh"
using namespace kosk;

__device__ __forceinline__ void mac_tile(int32_t *out, int iters, int seed)
{
    int32_t acc[8][7];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 7; j++) acc[i][j] = 0;
    int32_t a[8], b[7];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + i * 17 + threadIdx.x;
#pragma unroll
    for (int j = 0; j < 7; j++) b[j] = seed * 3 + j * 29 - threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 7; j++) acc[i][j] += a[i] * b[j];
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = (a[i] >> 1) ^ b[i % 7];       // cheap operand refresh (stands in for the LDS of the tile)
#pragma unroll
            for (int j = 0; j < 7; j++) b[j] += 3;
        }
    }
    int32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 7; j++) s ^= acc[i][j];
    if (s == 0x12345) out[0] = s;
}
__device__ __forceinline__ void kec(uint64_t *out, int perms, int seed)
{
    uint64_t a[25];
#pragma unroll
    for (int i = 0; i < 25; i++) a[i] = (uint64_t)(seed + threadIdx.x) * (2 * i + 1);
#pragma unroll 1
    for (int p = 0; p < perms; p++) keccak_f1600(a);
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < 25; i++) s ^= a[i];
    if (s == 0x12345) out[0] = s;
}
template <int MODE>
__global__ void __launch_bounds__(256, 2) kdual(int32_t *o1, uint64_t *o2, int iters, int perms)
{
    const int w = threadIdx.x >> 5;
    const bool do_mac = MODE == 0 || (MODE == 2 && (w & 4) == 0);
    if (do_mac) mac_tile(o1, iters, blockIdx.x);
    else kec(o2, perms, blockIdx.x);
}
template <int MODE> static float run(int blocks, int iters, int perms, int32_t *o1, uint64_t *o2)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0); kdual<MODE><<<blocks, 256>>>(o1, o2, iters, perms); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    return best;
}
int main()
{
    int32_t *o1; uint64_t *o2; cudaMalloc(&o1, 64); cudaMalloc(&o2, 64);
    const int blocks = 148 * 2 * 4;
    // per-warp work chosen so that a MAC warp and a Keccak warp take about the same time alone
    for (int perms : {32, 53, 80}) {
        const int iters = 1024;    // 1024 * 4 * 56 = 229k IMAD per thread
        float t0 = run<0>(blocks, iters, perms, o1, o2), t1 = run<1>(blocks, iters, perms, o1, o2), t2 = run<2>(blocks, iters, perms, o1, o2);
        printf("{\"iters\": %d, \"perms\": %d, \"mac_only_ms\": %.3f, \"keccak_only_ms\": %.3f, \"half_half_ms\": %.3f, \"serial_model_ms\": %.3f, \"overlap_model_ms\": %.3f, \"err\": \"%s\"}\n",
               iters, perms, t0, t1, t2, 0.5f * (t0 + t1), 0.5f * (t0 > t1 ? t0 : t1), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
