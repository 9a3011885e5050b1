// warp_slots.cu -- where do the warps of a CTA land?  Prints, for the CTAs resident on SM 0, the hardware warp slot (%warpid) of every warp;
// slot % 4 is the SM sub-partition (scheduler) on NVIDIA GPUs since Volta.  Question behind it: can a fused hash + sponge kernel keep its
// latency-bound sponge warps on a sub-partition of their own (warp 0 of every 4-warp CTA)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_slots warp_slots.cu && ./warp_slots
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int *out, int warps)
{
    unsigned smid, wid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    if ((threadIdx.x & 31) == 0) { out[(blockIdx.x * warps + (threadIdx.x >> 5)) * 2] = smid; out[(blockIdx.x * warps + (threadIdx.x >> 5)) * 2 + 1] = wid; }
    // stay resident so that several CTAs share the SM
    long long t0 = clock64(); while (clock64() - t0 < 2000000) { }
}
int main()
{
    for (int warps : {3, 4, 5}) {
        const int ctas = 148 * 7;
        int *d; cudaMalloc(&d, ctas * warps * 2 * sizeof(int));
        k<<<ctas, 32 * warps>>>(d, warps);
        int *h = new int[ctas * warps * 2];
        cudaMemcpy(h, d, ctas * warps * 2 * sizeof(int), cudaMemcpyDeviceToHost);
        printf("{\"warps_per_cta\": %d, \"sm0\": [", warps);
        bool first = true;
        for (int c = 0; c < ctas; c++) if (h[c * warps * 2] == 0) {
            printf("%s[", first ? "" : ", "); first = false;
            for (int w = 0; w < warps; w++) printf("%s%d", w ? ", " : "", h[(c * warps + w) * 2 + 1]);
            printf("]");
        }
        printf("]}\n");
        cudaFree(d); delete[] h;
    }
    return 0;
}
