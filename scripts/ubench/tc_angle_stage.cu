// Throughput estimate for a tensor-core angle stage WITHOUT shared-memory operand staging (plan for the next round):
// per tile of 128 range bins a 4-warp group reads each thread's own row of range spectra (16 floats) from shared
// memory, splits it into tf32 hi/lo parts, writes them to TENSOR MEMORY with tcgen05.st (TMEM lane = row), one thread
// issues the 3xTF32 MMAs  D[128 x 128] = A[128 x 16] * B[128 x 16]^T  (A from TMEM, B from a swizzled shared tile,
// D columns = [Re bins 0..63 | Im bins 0..63]), the threads read D back with tcgen05.ld, form re^2+im^2 with packed
// FMUL2/FFMA2, keep a running maximum, stage the map rows in shared memory and stream them out coalesced.
// Same output volume as k_fused64x8 on configs[1]: 4096 CPIs x 1024 x 64 floats = 1 GiB.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_angle_stage tc_angle_stage.cu && ./tc_angle_stage
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#ifndef NGROUPS
#define NGROUPS 3
#endif
constexpr int GROUPS = NGROUPS, THREADS = 128 * GROUPS, ROWB = 272;      // staging row: 256 B + 16 B pad (conflict-free STS.128)
constexpr int G_Y = 128 * 64, G_STG = 128 * ROWB, G_SIZE = ((G_Y + G_STG + 1023) / 1024) * 1024;
constexpr int OFF_G = 16384, SMEM = OFF_G + GROUPS * G_SIZE + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
#define LD32(taddr, v)                                                                                                             \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                         \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),   \
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),      \
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),      \
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                              \
                 : "r"(taddr))
#define LD16(taddr, v)                                                                                                             \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"          \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),   \
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                                    \
                 : "r"(taddr))
#define ST32(taddr, v)                                                                                                             \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                                   \
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), \
                   "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),            \
                   "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),            \
                   "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory")

#ifndef MAXREG_THREADS
#define MAXREG_THREADS THREADS
#endif
__global__ void __launch_bounds__(MAXREG_THREADS, 1) k_tc_angle(const float *Bblob, float *map, float *maxout, int n_tiles)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    __shared__ uint64_t mbar_all[GROUPS];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, grp = tid >> 7, r = tid & 127, warp4 = (tid >> 5) & 3, lane = tid & 31;
    unsigned char *gb = base + OFF_G + grp * G_SIZE;
    float4 *ytile = reinterpret_cast<float4 *>(gb);                 // [128 rows][16 floats]
    unsigned char *stg = gb + G_Y;                                   // [128 rows][ROWB]
    for (int e = tid; e < 128 * 32; e += THREADS) reinterpret_cast<float *>(base)[e] = Bblob[e];      // swizzled image
    for (int e = r; e < 128 * 16; e += 128) reinterpret_cast<float *>(ytile)[e] = 0.001f * ((e * 37 + grp) % 101) - 0.05f;
    if (r == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar_all[grp])));
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_slot + (uint32_t)(grp * 160), tmem_a = tmem_d + 128u;
    const uint32_t lane_off = (uint32_t)(warp4 * 32) << 16;
    const uint32_t mbar = smem_u32(&mbar_all[grp]);
    const uint32_t idesc = make_idesc_tf32(128, 128);
    const uint64_t db = make_desc_sw128(smem_u32(base));
    uint32_t parity = 0;
    float best = -1.f;
    for (int tile = blockIdx.x * GROUPS + grp; tile < n_tiles; tile += gridDim.x * GROUPS) {
        // 1-2: this thread's row of range spectra, split into tf32 hi / lo
        uint32_t a[32];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float4 t = ytile[r * 4 + j];
            const float f[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t hi = __float_as_uint(f[q]) & 0xFFFFE000u;
                a[4 * j + q] = hi;
                a[16 + 4 * j + q] = __float_as_uint(f[q] - __uint_as_float(hi));
            }
        }
        // 4: A -> tensor memory (lane = row), then the MMAs by one thread of the group
        ST32(tmem_a + lane_off, a);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
        if (r == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // D = Ahi*Bhi + Alo*Bhi + Ahi*Blo; K = 16 = 2 steps of 8; B tile rows: [Bhi (16) | Blo (16)] tf32
            const uint32_t acol[6] = {0, 8, 16, 24, 0, 8};
            const uint32_t bcol[6] = {0, 8, 0, 8, 16, 24};
#pragma unroll
            for (int i = 0; i < 6; i++) {
                const uint64_t b = db + (uint64_t)((bcol[i] * 4) >> 4);
                const uint32_t acc = i > 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a + acol[i]), "l"(b),
                             "r"(idesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
        }
        {
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
            parity ^= 1;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // 6: epilogue, two halves of 32 bins: Re columns [h*32, +32), Im columns [64 + h*32, +32)
        float4 *srow = reinterpret_cast<float4 *>(stg + r * ROWB);
#ifndef CHUNK16
#define CHUNK16 0
#endif
#pragma unroll
        for (int h = 0; h < (CHUNK16 ? 4 : 2); h++) {
#if CHUNK16
            uint32_t re[16], im[16];                        // 16 bins per step: fits 128 registers per thread
            LD16(tmem_d + lane_off + (uint32_t)(h * 16), re);
            LD16(tmem_d + lane_off + (uint32_t)(64 + h * 16), im);
#else
            uint32_t re[32], im[32];
            LD32(tmem_d + lane_off + (uint32_t)(h * 32), re);
            LD32(tmem_d + lane_off + (uint32_t)(64 + h * 32), im);
#endif
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < (CHUNK16 ? 4 : 8); j++) {
                float2 o[2];
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const float2 rr = make_float2(__uint_as_float(re[4 * j + 2 * q]), __uint_as_float(re[4 * j + 2 * q + 1]));
                    const float2 ii = make_float2(__uint_as_float(im[4 * j + 2 * q]), __uint_as_float(im[4 * j + 2 * q + 1]));
                    o[q] = __ffma2_rn(ii, ii, __fmul2_rn(rr, rr));
                }
                best = fmaxf(best, fmaxf(fmaxf(o[0].x, o[0].y), fmaxf(o[1].x, o[1].y)));
                srow[h * (CHUNK16 ? 4 : 8) + j] = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
        // 7: the tile (128 rows x 256 B, contiguous in the map) leaves coalesced
        float4 *dst = reinterpret_cast<float4 *>(map) + (size_t)tile * 2048;
#pragma unroll 4
        for (int k = 0; k < 16; k++) {
            const int i = k * 128 + r, row = i >> 4, ch = i & 15;
            __stcs(dst + i, *reinterpret_cast<const float4 *>(stg + row * ROWB + ch * 16));
        }
        asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
    }
    if (best > 1e30f) maxout[tid] = best;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512));
}

int main()
{
    const int n_cpi = 4096, n_tiles = n_cpi * 8;
    float *dB, *dmap, *dmax;
    cudaMalloc(&dB, 128 * 32 * 4); cudaMalloc(&dmap, (size_t)n_tiles * 32768); cudaMalloc(&dmax, THREADS * 4);
    float *hB = (float *)malloc(128 * 32 * 4);
    for (int i = 0; i < 128 * 32; i++) { float x = (float)rand() / RAND_MAX - 0.5f; uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&hB[i], &u, 4); }
    cudaMemcpy(dB, hB, 128 * 32 * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_tc_angle, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; w++) k_tc_angle<<<148, THREADS, SMEM>>>(dB, dmap, dmax, n_tiles);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaEventRecord(e0);
    for (int rep = 0; rep < 10; rep++) k_tc_angle<<<148, THREADS, SMEM>>>(dB, dmap, dmax, n_tiles);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("tensor-core angle stage + map store, 4096 CPIs x 1024 x 64: %.1f us (%.0f GB/s of map)\n", ms / 10 * 1e3,
           (double)n_tiles * 32768 / (ms / 10 * 1e-3) / 1e9);
    return 0;
}
