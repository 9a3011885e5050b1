// idp_bench.cu -- issue rate of the integer dot-product instructions next to IMAD on sm_100a (is IDP.2A / IDP.4A full rate on the FMA-heavy pipe?)
// Question behind it: the pointwise stage of share_ntt.cuh multiplies int16 spectra pairwise; as IMADs every operand first needs an ALU-pipe
// unpack (PRMT / SHF), as IDP.2A (16-bit x 8-bit limbs) it would not.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o idp_bench idp_bench.cu && ./idp_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k_peak(uint32_t *out, int iters, uint32_t seed)
{
    int32_t r[16], a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { r[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u; a[i] = r[i] * 0x85EBCA6Bu + i; }
    const int32_t m = seed | 1u, x = seed * 0x85EBCA6Bu + 3u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if (MODE == 0) r[i] = r[i] * m + x;                                   // IMAD
                else if (MODE == 1) r[i] = __dp2a_lo(a[i], m, r[i]);                   // IDP.2A.LO  acc += a.lo16 * m.b0 + a.hi16 * m.b1
                else if (MODE == 2) r[i] = __dp4a(a[i], m, r[i]);                      // IDP.4A
                else if (MODE == 3) r[i] = __dp2a_hi(a[i], m, r[i]);                   // IDP.2A.HI
                else if (MODE == 4) r[i] += (int32_t)(int16_t)(a[i] & 0xFFFF) * (int32_t)(int16_t)(m & 0xFFFF);   // unpack + IMAD (a loop-invariant: hoisted)
                else if (MODE == 5) { r[i] = __dp2a_lo(a[i], m, r[i]); a[i] = a[i] ^ (~a[(i + 1) & 15] & x); }     // IDP.2A + LOP3 (co-issue with the ALU pipe)
                else if (MODE == 6) { r[i] = r[i] * m + x; a[i] = a[i] ^ (~a[(i + 1) & 15] & x); }                 // IMAD + LOP3
            }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) acc ^= r[i] ^ a[i];
    if (acc == 0x12345678u) out[blockIdx.x] = acc;
}

template <int MODE>
static double run(uint32_t *d, int blocks, int iters)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k_peak<MODE><<<blocks, 256>>>(d, iters, 12345u + rep);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    return (double)blocks * 256.0 * iters * 64.0 / (best * 1e-3);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, iters = 4096;
    uint32_t *d; cudaMalloc(&d, blocks * 4);
    const char *names[] = {"imad", "idp2a_lo", "idp4a", "idp2a_hi", "unpack_imad", "idp2a_plus_lop3", "imad_plus_lop3"};
    double v[7] = {run<0>(d, blocks, iters), run<1>(d, blocks, iters), run<2>(d, blocks, iters), run<3>(d, blocks, iters), run<4>(d, blocks, iters),
                   run<5>(d, blocks, iters), run<6>(d, blocks, iters)};
    printf("{\"bench\": \"idp\", \"sms\": %d", p.multiProcessorCount);
    for (int i = 0; i < 7; i++) printf(", \"%s_tops\": %.3f", names[i], v[i] * 1e-12);
    printf(", \"note\": \"thread-level instruction pairs (modes 5, 6: one FMA-pipe + one ALU-pipe instruction per count) per second\"}\n");
    return 0;
}
