// Store-stream floor of k_fused64x8: the same grid (2 CTAs x 148 SMs, 8 warps), the same order of 1 KiB tiles per
// warp (16 pair iterations, two tiles Q rows apart), float4 stores -- without any compute -- against a plain
// sequential fill of the same 1 GiB, for the four store cache policies.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -DSTMODE=0..3 -o store_pattern store_pattern.cu && ./store_pattern
#include <cuda_runtime.h>
#include <cstdio>

#ifndef STMODE
#define STMODE 0
#endif
constexpr int NR = 1024, NA = 64, Q = NR / 8, G = 4, TILE = G * NA, PITERS = (Q / 2) / G;

__device__ __forceinline__ void st4(float4 *p, float4 v)
{
#if STMODE == 0
    __stcs(p, v);      // st.global.cs: streaming, evict first
#elif STMODE == 1
    *p = v;            // default (write back, evict normal)
#elif STMODE == 2
    __stcg(p, v);      // st.global.cg
#else
    __stwt(p, v);      // st.global.wt: write through
#endif
}

__global__ void __launch_bounds__(256) k_pattern(float *map, int n_cpi)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int jp = warp >> 1, qw = (warp & 1) * (Q / 2), row0 = 2 * jp * Q + qw;
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (int cpi = blockIdx.x; cpi < n_cpi; cpi += gridDim.x) {
        float4 *map_w = reinterpret_cast<float4 *>(map + ((long long)cpi * NR + row0) * NA);
        for (int it = 0; it < PITERS; it++) {
            float4 *dst = map_w + it * (TILE / 4);
            st4(dst + lane, v);
            st4(dst + lane + 32, v);
            st4(dst + Q * NA / 4 + lane, v);
            st4(dst + Q * NA / 4 + lane + 32, v);
        }
    }
}

__global__ void k_fill(float4 *p, long long n4)
{
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) st4(p + i, v);
}

int main()
{
    const int n_cpi = 4096;
    const long long n = (long long)n_cpi * NR * NA;
    float *d;
    cudaMalloc(&d, n * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grids[] = {148, 296, 592, 1184, 2368};      // resident CTAs per SM: 1, 2, 4, 8, 8
    for (int mode = 0; mode < 2; mode++)
        for (int g : grids) {
            for (int w = 0; w < 3; w++) { if (mode == 0) k_pattern<<<g, 256>>>(d, n_cpi); else k_fill<<<g, 256>>>((float4 *)d, n / 4); }
            cudaEventRecord(e0);
            for (int r = 0; r < 20; r++) { if (mode == 0) k_pattern<<<g, 256>>>(d, n_cpi); else k_fill<<<g, 256>>>((float4 *)d, n / 4); }
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("STMODE %d %-28s grid %4d x 256: %.1f us per GiB -> %.0f GB/s\n", STMODE,
                   mode == 0 ? "tile pattern of k_fused64x8" : "sequential fill", g, ms / 20 * 1e3, n * 4.0 / (ms / 20 * 1e-3) / 1e9);
        }
    return 0;
}
