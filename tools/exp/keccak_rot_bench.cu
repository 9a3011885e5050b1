// Experiment: Keccak-f[1600] rho rotations on the FMA pipe (IMAD.HI + IMAD with power-of-two multipliers from the constant
// bank) instead of two SHF funnel shifts on the ALU pipe.  NMOVE = how many of the 24 rho lanes use the IMAD form.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o keccak_rot_bench keccak_rot_bench.cu
#include <cstdint>
#include <cstdio>
#include <vector>
#include "../../mpcith_kyber_kosk_b200/csrc/keccak.cuh"
using namespace kosk;
__constant__ uint32_t c_p2[32];
__device__ __forceinline__ uint64_t rol64m(uint64_t x, int n)
{
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32), rl, rh;
    if (n == 0) return x;
    if (n == 32) return ((uint64_t)lo << 32) | hi;
    if (n > 32) { uint32_t t = lo; lo = hi; hi = t; n -= 32; }
    const uint32_t m = c_p2[n];
    uint32_t t1, t2;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(t1) : "r"(hi), "r"(m));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(rl) : "r"(lo), "r"(m), "r"(t1));
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(t2) : "r"(lo), "r"(m));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(rh) : "r"(hi), "r"(m), "r"(t2));
    return ((uint64_t)rh << 32) | rl;
}
template <int NMOVE, int ROT1>
__device__ __forceinline__ void keccak_v(uint64_t (&a)[25])
{
#pragma unroll 1
    for (int r = 0; r < 24; r++) {
        uint64_t c[5], b[25];
#pragma unroll
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
#pragma unroll
        for (int x = 0; x < 5; x++) {
            uint64_t d = c[(x + 4) % 5] ^ (ROT1 ? rol64m(c[(x + 1) % 5], 1) : rol64(c[(x + 1) % 5], 1));
#pragma unroll
            for (int y = 0; y < 25; y += 5) a[x + y] ^= d;
        }
        int cnt = 0;
#define KOSK_RP(x, y, n) { b[(y) + 5 * ((2 * (x) + 3 * (y)) % 5)] = (n != 0 && cnt < NMOVE) ? rol64m(a[(x) + 5 * (y)], n) : rol64(a[(x) + 5 * (y)], n); if (n != 0) cnt++; }
        KOSK_RP(0, 0, 0)  KOSK_RP(1, 0, 1)  KOSK_RP(2, 0, 62) KOSK_RP(3, 0, 28) KOSK_RP(4, 0, 27)
        KOSK_RP(0, 1, 36) KOSK_RP(1, 1, 44) KOSK_RP(2, 1, 6)  KOSK_RP(3, 1, 55) KOSK_RP(4, 1, 20)
        KOSK_RP(0, 2, 3)  KOSK_RP(1, 2, 10) KOSK_RP(2, 2, 43) KOSK_RP(3, 2, 25) KOSK_RP(4, 2, 39)
        KOSK_RP(0, 3, 41) KOSK_RP(1, 3, 45) KOSK_RP(2, 3, 15) KOSK_RP(3, 3, 21) KOSK_RP(4, 3, 8)
        KOSK_RP(0, 4, 18) KOSK_RP(1, 4, 2)  KOSK_RP(2, 4, 61) KOSK_RP(3, 4, 56) KOSK_RP(4, 4, 14)
#undef KOSK_RP
#pragma unroll
        for (int y = 0; y < 25; y += 5)
#pragma unroll
            for (int x = 0; x < 5; x++) a[x + y] = b[x + y] ^ (~b[(x + 1) % 5 + y] & b[(x + 2) % 5 + y]);
        a[0] ^= c_keccak_rc[r];
    }
}
template <int NMOVE, int ROT1>
__global__ void __launch_bounds__(128) kbench(uint64_t *io, int n)
{
    const size_t t = (size_t)blockIdx.x * 128 + threadIdx.x;
    uint64_t a[25];
    for (int i = 0; i < 25; i++) a[i] = io[i] + t * (2 * i + 1);
    for (int i = 0; i < n; i++) keccak_v<NMOVE, ROT1>(a);
    uint64_t x = 0;
    for (int i = 0; i < 25; i++) x ^= a[i] * (i + 1);
    io[32 + t] = x;
}
template <int NMOVE, int ROT1>
static void run(uint64_t *d, int blocks, int n, std::vector<uint64_t> &ref)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        kbench<NMOVE, ROT1><<<blocks, 128>>>(d, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    std::vector<uint64_t> out((size_t)blocks * 128);
    cudaMemcpy(out.data(), d + 32, out.size() * 8, cudaMemcpyDeviceToHost);
    bool ok = true;
    if (ref.empty()) ref = out; else ok = (ref == out);
    printf("{\"nmove\": %d, \"rot1\": %d, \"ms\": %.4f, \"gperm_per_s\": %.3f, \"match\": %s, \"err\": \"%s\"}\n", NMOVE, ROT1, best,
           (double)blocks * 128 * n / (best * 1e-3) / 1e9, ok ? "true" : "false", cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    uint32_t p2[32]; for (int i = 0; i < 32; i++) p2[i] = 1u << i;
    cudaMemcpyToSymbol(c_p2, p2, sizeof p2);
    const int blocks = 148 * 32, n = 64;
    uint64_t *d; cudaMalloc(&d, (32 + (size_t)blocks * 128) * 8);
    std::vector<uint64_t> h(32); for (int i = 0; i < 32; i++) h[i] = 0x9E3779B97F4A7C15ULL * (i + 1);
    cudaMemcpy(d, h.data(), 32 * 8, cudaMemcpyHostToDevice);
    std::vector<uint64_t> ref;
    run<0, 0>(d, blocks, n, ref);
    run<6, 0>(d, blocks, n, ref);
    run<12, 0>(d, blocks, n, ref);
    run<16, 0>(d, blocks, n, ref);
    run<20, 0>(d, blocks, n, ref);
    run<24, 0>(d, blocks, n, ref);
    run<24, 1>(d, blocks, n, ref);
    run<12, 1>(d, blocks, n, ref);
    return 0;
}
