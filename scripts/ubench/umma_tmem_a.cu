// Variant of umma_tf32.cu with the A operand in TENSOR MEMORY: every thread writes its own row (TMEM lane) with
// tcgen05.st, the MMA takes [a_tmem] instead of a shared-memory descriptor -- no operand staging in shared memory.
// Stand-alone check of a hand-written tcgen05 (UMMA) tf32 GEMM tile on sm_100a:
//   D[128 x 128] (fp32, TMEM) = A[128 x 32] * B[128 x 32]^T, A and B K-major in shared memory with the
//   128-byte swizzle, filled by ordinary st.shared; 4 x tcgen05.mma (K = 8 each); tcgen05.ld epilogue.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_tf32 umma_tf32.cu && ./umma_tf32
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address
    d |= (uint64_t)0 << 16;                            // LBO (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                  // SBO: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N)
{
    uint32_t d = 0;
    d |= 1u << 4;                 // c_format = F32
    d |= 2u << 7;                 // a_format = TF32
    d |= 2u << 10;                // b_format = TF32
    d |= 0u << 15;                // a_major = K
    d |= 0u << 16;                // b_major = K
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

__global__ void __launch_bounds__(128, 1) k(const float *A, const float *B, float *D)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *base = (unsigned char *)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    float *sA = (float *)base;                 // 128 rows x 128 B
    float *sB = (float *)(base + 16384);
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // fill A, B with the 128B swizzle: element (r, k) -> (r/8)*1024 + (r%8)*128 + ((k/4) ^ (r%8))*16 + (k%4)*4
    for (int e = tid; e < 128 * 32; e += 128) {
        int r = e / 32, kk = e % 32;
        int off = (r >> 3) * 1024 + (r & 7) * 128 + (((kk >> 2) ^ (r & 7)) << 4) + ((kk & 3) << 2);
        *(float *)((unsigned char *)sB + off) = B[e];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic st.shared -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    {   // A row of this thread -> TMEM lane tid, columns 128..159 (K = 32 tf32 values)
        uint32_t v[32];
        for (int j = 0; j < 32; j++) v[j] = __float_as_uint(A[tid * 32 + j]);
        uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 128u;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                     "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                     ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                       "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
                       "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
                       "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(128, 128);
        const uint64_t db = make_desc_sw128(smem_u32(sB));
        for (int k8 = 0; k8 < 4; k8++) {
            uint64_t b = db + (uint64_t)((k8 * 32) >> 4);
            uint32_t a = tmem + 128u + (uint32_t)(k8 * 8);        // 8 tf32 columns per K step
            uint32_t acc = k8 > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(a), "l"(b), "r"(idesc), "r"(acc)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    // everyone waits for the MMAs
    {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w reads lanes 32w..32w+31 (rows), 4 x 32 columns
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                       "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                       "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; j++) D[(warp * 32 + lane) * 128 + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

int main()
{
    const int M = 128, N = 128, K = 32;
    float *hA = (float *)malloc(M * K * 4), *hB = (float *)malloc(N * K * 4), *hD = (float *)malloc(M * N * 4);
    srand(1);
    auto tf32 = [](float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; };
    for (int i = 0; i < M * K; i++) hA[i] = tf32((float)rand() / RAND_MAX - 0.5f);
    for (int i = 0; i < N * K; i++) hB[i] = tf32((float)rand() / RAND_MAX - 0.5f);
    float *dA, *dB, *dD;
    cudaMalloc(&dA, M * K * 4); cudaMalloc(&dB, N * K * 4); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, N * K * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, M * N * 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * 1024);
    k<<<1, 128, 34 * 1024>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            double s = 0;
            for (int kk = 0; kk < K; kk++) s += (double)hA[i * K + kk] * hB[j * K + kk];
            maxerr = fmax(maxerr, fabs(s - hD[i * N + j]));
            maxref = fmax(maxref, fabs(s));
        }
    printf("max |D - ref| = %.3e (max |ref| = %.3f)  D[0][0..3] = %f %f %f %f\n", maxerr, maxref, hD[0], hD[1], hD[2], hD[3]);
    printf(maxerr < 1e-4 ? "UMMA TF32 TILE (A IN TMEM) OK\n" : "UMMA TF32 TILE (A IN TMEM) MISMATCH\n");
    return 0;
}
