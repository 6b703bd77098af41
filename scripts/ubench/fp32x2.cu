// Micro-benchmark: issue/pipe throughput of scalar vs packed FP32 on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2 fp32x2.cu && ./fp32x2
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float2 *out, int iters, float2 seed)
{
    float2 a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = make_float2(seed.x + i + threadIdx.x, seed.y - i);
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, -0.5f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (MODE == 0) { a[i].x = __fmaf_rn(a[i].x, m.x, c.x); a[i].y = __fmaf_rn(a[i].y, m.y, c.y); }   // 2 FFMA
            if (MODE == 1) { a[i] = __ffma2_rn(a[i], m, c); }                                               // 1 FFMA2
            if (MODE == 2) { a[i].x = __fadd_rn(a[i].x, c.x); a[i].y = __fadd_rn(a[i].y, c.y); }             // 2 FADD
            if (MODE == 3) { a[i] = __fadd2_rn(a[i], c); }                                                  // 1 FADD2
            if (MODE == 4) { a[i] = __fmul2_rn(a[i], m); }                                                  // 1 FMUL2
            if (MODE == 5) { a[i] = __ffma2_rn(a[i], m, c); a[i].x = __fmaf_rn(a[i].x, m.y, c.y); }          // 1 FFMA2 + 1 FFMA
        }
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 16; i++) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int flops_per_elem_iter)
{
    const int blocks = 148 * 4, threads = 512, iters = 4096;
    float2 *d;
    cudaMalloc(&d, sizeof(float2) * blocks * threads);
    k<MODE><<<blocks, threads>>>(d, 16, make_float2(1, 2));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, iters, make_float2(1, 2));
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double lane_ops = (double)blocks * threads * iters * 16 * flops_per_elem_iter;   // scalar-equivalent FP ops
    printf("%-28s %8.3f ms  %7.2f T scalar-ops/s  (%.1f ops/clk/SM at 1.965 GHz)\n", name, ms, lane_ops / ms / 1e9,
           lane_ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(d);
}

int main()
{
    run<0>("2x FFMA (scalar)", 2);
    run<1>("1x FFMA2 (packed)", 2);
    run<2>("2x FADD (scalar)", 2);
    run<3>("1x FADD2 (packed)", 2);
    run<4>("1x FMUL2 (packed)", 2);
    run<5>("FFMA2 + FFMA", 3);
    return 0;
}
