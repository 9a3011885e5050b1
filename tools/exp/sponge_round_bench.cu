// Experiment: per-round latency of the warp-cooperative Keccak-f[1600] (one state word per lane) under variants of the round:
//   V0 the product's round (3 exchange stages: 8 | 4 | 6 shuffles)
//   V1 V0 with iota taken off the critical path (rc folded into the theta inputs of the next round)
//   V2 theta in one exchange stage: every lane fetches the ten words of its two neighbour columns (20 | 6 shuffles)
//   V3 V1 + V2
//   V4 column parity with three dependent shuffles (a ^ a[y+1], then ^ the same two rows further, then ^ a[y+4]): 6 | 4 | 6 shuffles in five stages
//   V5 column parity with two stages (p = a ^ a[y+1]; c = p ^ a[y+2] ^ p[y+3]): 6 | 4 | 6 shuffles in four stages
// Measures ns per permutation for a chain of dependent permutations, with 1 warp per SM and with ~7 warps per SM (the 1024-proof load).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o sponge_round_bench sponge_round_bench.cu
#include <cstdint>
#include <cstdio>
#include <vector>
#include "../../mpcith_kyber_kosk_b200/csrc/keccak.cuh"
using namespace kosk;

struct WK {
    int t, x, y, rho, src_pi, src_c1, src_c2, l5, l10, l15, l20, lm1, lp1, cm[5], cp[5];
    __device__ __forceinline__ void init()
    {
        const int rho_tab[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
        const int lane = threadIdx.x & 31;
        t = lane < 25 ? lane : 0; x = t % 5; y = t / 5;
        int r = 0;
#pragma unroll
        for (int i = 0; i < 25; i++) if (i == t) r = rho_tab[i];
        rho = r;
        src_pi = lane < 25 ? ((3 * y + x) % 5) + 5 * x : lane;
        l5 = lane < 25 ? (t + 5) % 25 : lane; l10 = lane < 25 ? (t + 10) % 25 : lane;
        l15 = lane < 25 ? (t + 15) % 25 : lane; l20 = lane < 25 ? (t + 20) % 25 : lane;
        lm1 = lane < 25 ? (x + 4) % 5 + 5 * y : lane; lp1 = lane < 25 ? (x + 1) % 5 + 5 * y : lane;
        const int x1 = (x + 1) % 5, x2 = (x + 2) % 5;
        src_c1 = lane < 25 ? ((3 * y + x1) % 5) + 5 * x1 : lane;
        src_c2 = lane < 25 ? ((3 * y + x2) % 5) + 5 * x2 : lane;
#pragma unroll
        for (int yy = 0; yy < 5; yy++) { cm[yy] = lane < 25 ? (x + 4) % 5 + 5 * yy : lane; cp[yy] = lane < 25 ? (x + 1) % 5 + 5 * yy : lane; }
    }
    static __device__ __forceinline__ uint64_t shfl(uint64_t v, int src)
    {
        const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
        return ((uint64_t)hi << 32) | lo;
    }
    __device__ __forceinline__ uint64_t rhopichi(uint64_t a) const
    {
        uint32_t lo = (uint32_t)a, hi = (uint32_t)(a >> 32);
        if (rho & 32) { const uint32_t tmp = lo; lo = hi; hi = tmp; }
        const uint32_t nh = __funnelshift_l(lo, hi, rho), nl = __funnelshift_l(hi, lo, rho);
        const uint64_t ar = ((uint64_t)nh << 32) | nl;
        const uint64_t b = shfl(ar, src_pi), b1 = shfl(ar, src_c1), b2 = shfl(ar, src_c2);
        return b ^ (~b1 & b2);
    }
    template <int V, int UNROLL>
    __device__ __forceinline__ uint64_t permute(uint64_t a) const
    {
        const bool lane0 = (threadIdx.x & 31) == 0;
        if (V == 0 || V == 2 || V == 4 || V == 5) {
#pragma unroll UNROLL
            for (int r = 0; r < 24; r++) {
                if (V == 0) {
                    const uint64_t c = a ^ shfl(a, l5) ^ shfl(a, l10) ^ shfl(a, l15) ^ shfl(a, l20);
                    a ^= shfl(c, lm1) ^ rol64(shfl(c, lp1), 1);
                } else if (V == 4) {
                    const uint64_t s1 = a ^ shfl(a, l5), s2 = s1 ^ shfl(s1, l10), c = s2 ^ shfl(a, l20);
                    a ^= shfl(c, lm1) ^ rol64(shfl(c, lp1), 1);
                } else if (V == 5) {
                    const uint64_t p = a ^ shfl(a, l5), q = shfl(a, l10), c = p ^ q ^ shfl(p, l15);
                    a ^= shfl(c, lm1) ^ rol64(shfl(c, lp1), 1);
                } else {
                    const uint64_t m = shfl(a, cm[0]) ^ shfl(a, cm[1]) ^ shfl(a, cm[2]) ^ shfl(a, cm[3]) ^ shfl(a, cm[4]);
                    const uint64_t p = shfl(a, cp[0]) ^ shfl(a, cp[1]) ^ shfl(a, cp[2]) ^ shfl(a, cp[3]) ^ shfl(a, cp[4]);
                    a ^= m ^ rol64(p, 1);
                }
                a = rhopichi(a);
                if (lane0) a ^= c_keccak_rc[r];
            }
            return a;
        } else {
            // iota deferred: `a` lacks rc_prev in lane 0; rcx = rc_prev where this lane's theta inputs see lane 0's word
            uint64_t rc_prev = 0;
#pragma unroll UNROLL
            for (int r = 0; r < 24; r++) {
                // true state = a ^ (lane0 ? rc_prev : 0).  Column 0 parity gains rc_prev; D[x] uses C[x-1], C[x+1].
                const uint64_t own = lane0 ? rc_prev : 0;                       // off the critical path
                const uint64_t dm = (x == 1) ? rc_prev : 0, dp = (x == 4) ? rol64(rc_prev, 1) : 0;   // C[0] feeds D[1] plain and D[4] rotated
                const uint64_t fix = own ^ dm ^ dp;
                if (V == 1) {
                    const uint64_t c = a ^ shfl(a, l5) ^ shfl(a, l10) ^ shfl(a, l15) ^ shfl(a, l20);
                    a ^= shfl(c, lm1) ^ rol64(shfl(c, lp1), 1) ^ fix;
                } else {
                    const uint64_t m = shfl(a, cm[0]) ^ shfl(a, cm[1]) ^ shfl(a, cm[2]) ^ shfl(a, cm[3]) ^ shfl(a, cm[4]);
                    const uint64_t p = shfl(a, cp[0]) ^ shfl(a, cp[1]) ^ shfl(a, cp[2]) ^ shfl(a, cp[3]) ^ shfl(a, cp[4]);
                    a ^= m ^ rol64(p, 1) ^ fix;
                }
                a = rhopichi(a);
                rc_prev = c_keccak_rc[r];
            }
            if (lane0) a ^= rc_prev;
            return a;
        }
    }
};

template <int V, int UNROLL>
__global__ void __launch_bounds__(128, 4) kchain(uint64_t *io, int perms)
{
    WK wk; wk.init();
    const int lane = threadIdx.x & 31, w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint64_t a = io[lane] + (uint64_t)w * 0x9E3779B97F4A7C15ULL * (lane + 1);
#pragma unroll 1
    for (int p = 0; p < perms; p++) { a ^= (uint64_t)p * (lane + 3); a = wk.permute<V, UNROLL>(a); }
    if (lane < 25) io[64 + (size_t)w * 32 + lane] = a;
}
template <int V, int UNROLL>
static void run(uint64_t *d, int blocks, int threads, int perms, std::vector<uint64_t> &ref, const char *tag)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0); kchain<V, UNROLL><<<blocks, threads>>>(d, perms); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    std::vector<uint64_t> out((size_t)blocks * threads);
    cudaMemcpy(out.data(), d + 64, out.size() * 8, cudaMemcpyDeviceToHost);
    bool ok = true;
    if (ref.empty()) ref = out; else ok = (ref == out);
    printf("{\"load\": \"%s\", \"variant\": %d, \"unroll\": %d, \"ns_per_perm\": %.1f, \"cycles_per_round_at_1965\": %.1f, \"match\": %s, \"err\": \"%s\"}\n", tag, V, UNROLL,
           best * 1e6 / perms, best * 1e6 / perms / 24 * 1.965, ok ? "true" : "false", cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    uint64_t *d; cudaMalloc(&d, (64 + 148 * 8 * 128) * 8);
    std::vector<uint64_t> h(64); for (int i = 0; i < 64; i++) h[i] = 0xD1B54A32D192ED03ULL * (i + 1);
    cudaMemcpy(d, h.data(), 64 * 8, cudaMemcpyHostToDevice);
    const int perms = 344;
    {   std::vector<uint64_t> ref;
        run<0, 1>(d, 148, 32, perms, ref, "1 warp/SM"); run<0, 2>(d, 148, 32, perms, ref, "1 warp/SM"); run<0, 3>(d, 148, 32, perms, ref, "1 warp/SM"); run<0, 4>(d, 148, 32, perms, ref, "1 warp/SM");
        run<0, 6>(d, 148, 32, perms, ref, "1 warp/SM"); run<0, 8>(d, 148, 32, perms, ref, "1 warp/SM"); run<0, 24>(d, 148, 32, perms, ref, "1 warp/SM"); run<1, 4>(d, 148, 32, perms, ref, "1 warp/SM"); run<1, 24>(d, 148, 32, perms, ref, "1 warp/SM"); run<4, 24>(d, 148, 32, perms, ref, "1 warp/SM"); run<5, 24>(d, 148, 32, perms, ref, "1 warp/SM"); }
    {   std::vector<uint64_t> ref;
        run<0, 1>(d, 256, 128, perms, ref, "1024 warps"); run<0, 2>(d, 256, 128, perms, ref, "1024 warps"); run<0, 3>(d, 256, 128, perms, ref, "1024 warps"); run<0, 4>(d, 256, 128, perms, ref, "1024 warps");
        run<0, 6>(d, 256, 128, perms, ref, "1024 warps"); run<0, 8>(d, 256, 128, perms, ref, "1024 warps"); run<0, 24>(d, 256, 128, perms, ref, "1024 warps"); run<1, 4>(d, 256, 128, perms, ref, "1024 warps"); run<1, 24>(d, 256, 128, perms, ref, "1024 warps"); run<4, 24>(d, 256, 128, perms, ref, "1024 warps"); run<5, 24>(d, 256, 128, perms, ref, "1024 warps"); }
    {   std::vector<uint64_t> ref;      // one warp per CTA, as the product launches batches of 256 proofs and more
        run<0, 24>(d, 1024, 32, perms, ref, "1024 CTAs x 1 warp"); run<4, 24>(d, 1024, 32, perms, ref, "1024 CTAs x 1 warp"); run<5, 24>(d, 1024, 32, perms, ref, "1024 CTAs x 1 warp");
        run<0, 24>(d, 2048, 32, perms, ref, "2048 CTAs x 1 warp"); run<5, 24>(d, 2048, 32, perms, ref, "2048 CTAs x 1 warp"); run<0, 24>(d, 4096, 32, perms, ref, "4096 CTAs x 1 warp"); run<5, 24>(d, 4096, 32, perms, ref, "4096 CTAs x 1 warp"); }
    return 0;
}
