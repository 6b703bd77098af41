// jrc_tc.cuh -- the fused radar chain with the angle DFT on the 5th-generation tensor cores.
//
// BASELINE.json allows tensor cores "only if ncu shows the small-N angle DFT performs better when
// expressed as a complex GEMM".  This opt-in variant (JRC_FUSED_KERNEL=tc) is the experiment: correct (parity
// tests), but measured slower than the SIMT kernel of jrc_fused.cuh (0.26 ms + two helper kernels against
// 0.226 ms per 4096 CPIs) because its A tiles are staged in shared memory; scripts/ubench/tc_angle_stage.cu
// shows the form that does win (A operand in tensor memory).  Here the angle pass
//     M[n][i] = sum_p (-1)^p e^{-j 2 pi p i / Na} y[p][n]
// is the real GEMM  D[128 rows n][2 Na] = A[128][16] * B[2 Na][16]^T  per tile of 128 range bins, issued
// as tcgen05.mma kind::tf32 with the accumulator in TMEM, in 3xTF32 form for float32-class accuracy:
//     A = Ahi + Alo (Ahi = top 19 bits of y, Alo = y - Ahi), B = Bhi + Blo,
//     D = Ahi*Bhi + Alo*Bhi + Ahi*Blo                      (6 MMAs of K = 8 per tile, error ~2^-21)
// Columns 2i / 2i+1 of D are Re / Im of angle bin i, so the epilogue thread that owns TMEM lane n reads
// whole map rows: tcgen05.ld -> re^2+im^2 -> running arg-max -> swizzled staging -> 512-byte stores.
//
// Work unit = (CPI, two slices) = 128 range bins (see jrc_stream.cuh for the slice algebra): the two SIMT
// range passes run one task per thread and write the A tile straight into the 128-byte-swizzled K-major
// layout the UMMA descriptor expects.  One CTA per SM (TMEM kernels are not co-scheduled) made of four
// independent 4-warp groups, each with its own A tile, 128 TMEM columns, mbarrier and named barrier: while one
// group waits for its MMAs the others run their range passes and epilogues.
//
// Around it: k_chan_est (H per CPI, L2 resident) in front, k_tc_finalize (per-CPI 64-bit arg-max key ->
// detection record, noise window read back from the map) behind.
#pragma once
#include "jrc_common.cuh"
#include "jrc_staged.cuh"
#include "jrc_fused.cuh"
#include "jrc_stream.cuh"

namespace jrc {

struct TcParams {
    const c32 *H;                 // [n_cpi][8][64]
    int n_cpi, cpi0;
    float *map;                   // [n_cpi][NR][NA]
    unsigned long long *keys;     // [n_cpi], zeroed; (|.|^2 bits << 32) | (0xFFFFFFFF - range bin n); may be null
    const c32 *tw1g;              // [IR][8]
    const c32 *tw2g;              // [IR][8][8]
    const float *bblob;           // [2*NA rows][32] tf32 [Bhi | Blo], already in the swizzled smem image
    DetDev *dets;                 // finalize only
    EstParams est;                // finalize only
};

template <int IR, int IA>
struct TcGeom {
    static constexpr int NSC = 64, V = 8;
    static constexpr int NR = NSC * IR, NA = V * IA, Q = NR / 8;
    static constexpr int N = 2 * NA;                      // GEMM N: (re, im) per angle bin
    static constexpr int HST = 72;
    // A kernel that allocates TMEM gets ONE CTA per SM from the block scheduler (the occupancy API reports 1
    // whatever the register/shared-memory footprint), so the CTA itself holds GROUPS independent 4-warp groups,
    // each with its own A tile, TMEM columns, mbarrier and named barrier -- the "four CTAs per SM" of the design.
    static constexpr int GROUPS = 512 / N;                // 4 (Na = 64) or 2 (Na = 128): all 512 TMEM columns
    static constexpr int THREADS = 128 * GROUPS;
    static constexpr int TMEM_COLS = 512;
    // shared memory image (bytes), A tiles and the B tile 1024-byte aligned for the 128B swizzle
    static constexpr int OFF_B = 0;                       // N rows x 128 B, shared by the groups
    static constexpr int OFF_G = N * 128;                 // per group: A tile | H tile | B/y tiles
    static constexpr int G_A = 0;                         // 2 x (128 rows x 128 B): tile k+1 is built while the MMAs of
                                                          // tile k run; a consumed tile is the store staging
    static constexpr int G_H = 32768;                     // [8][HST] c32
    static constexpr int G_Y = G_H + 8 * HST * 8;         // 2 slices x [8][HST] c32
    static constexpr int G_SIZE = ((G_Y + 2 * 8 * HST * 8 + 1023) / 1024) * 1024;
    static constexpr int SMEM = OFF_G + GROUPS * G_SIZE + 1024;   // + slack for the 1024-byte alignment
    static_assert(N == 128 || N == 256, "angle zero-pad 64 or 128 (GEMM N of one tcgen05.mma)");
    static_assert(IR >= 2 && (IR & (IR - 1)) == 0, "range interp must be an even power of two");
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major operand, 128-byte swizzle: 8-row atoms of 1024 B (stride byte offset), descriptor version 1
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::tf32, fp32 accumulate, A and B K-major
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
                 "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}

template <int IR, int IA>
__global__ void __launch_bounds__(TcGeom<IR, IA>::THREADS, 1) k_tc64x8(const TcParams P)
{
    using Gm = TcGeom<IR, IA>;
    constexpr int NR = Gm::NR, NA = Gm::NA, Q = Gm::Q, N = Gm::N, HST = Gm::HST;
    constexpr int UPC = IR / 2;            // units (slice pairs) per CPI

    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment for the swizzle atoms, as an OFFSET so that the pointers keep their shared state space
    // (pointer arithmetic through uintptr_t makes ptxas emit generic LD/ST instead of LDS/STS)
    unsigned char *base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    __shared__ uint64_t mbar_all[Gm::GROUPS];
    __shared__ uint32_t tmem_slot;

    const int lane = threadIdx.x & 31;
    const int warp_cta = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int grp = warp_cta >> 2, warp = warp_cta & 3;    // group of 4 warps; warp within the group = TMEM lane quarter
    const int tid = threadIdx.x & 127;                      // thread within the group
    unsigned char *sB = base + Gm::OFF_B;
    unsigned char *gbase = base + Gm::OFF_G + grp * Gm::G_SIZE;
    unsigned char *sA = gbase + Gm::G_A;
    c32 *Hs = reinterpret_cast<c32 *>(gbase + Gm::G_H);
    c32 *Ysl = reinterpret_cast<c32 *>(gbase + Gm::G_Y);
    uint64_t &mbar = mbar_all[grp];
    auto group_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); };

    // ---- one-time setup: B operand image, barrier, TMEM ------------------------------------------
    for (int e = threadIdx.x; e < N * 8; e += Gm::THREADS)        // N rows x 8 chunks of 16 B
        reinterpret_cast<float4 *>(sB)[e] = __ldg(reinterpret_cast<const float4 *>(P.bblob) + e);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp_cta == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(Gm::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;
    const uint32_t tmem = tmem_base + (uint32_t)(grp * N);   // this group's accumulator columns
    const uint32_t idesc = umma_idesc_tf32(128, N);
    const uint64_t adesc0 = umma_desc_sw128(smem_u32(sA)), bdesc = umma_desc_sw128(smem_u32(sB));
    const uint32_t mbar_a = smem_u32(&mbar);
    uint32_t phase = 0;

    // ---- thread roles -----------------------------------------------------------------------------
    // range passes: slice s = tid / 64, channel p = (tid / 8) % 8, lo = tid % 8 (k0 in pass 1, m0 in pass 2)
    const int s_sl = tid >> 6, p_ch = (tid >> 3) & 7, lo = tid & 7;
    c32 *Ys = Ysl + s_sl * 8 * HST;
    // A tile byte offsets of this thread's (row m1*8+lo of slice s_sl, channel p_ch): + m1*1024 per output
    const int a_row = s_sl * 8192 + lo * 128 + ((p_ch & 1) << 3);
    const int a_hi = a_row + (((p_ch >> 1) ^ lo) << 4);
    const int a_lo = a_row + (((4 + (p_ch >> 1)) ^ lo) << 4);
    // epilogue: this thread owns tile row r = tid (TMEM lane tid): slice r/64, rho = r%64
    const int r_sl = tid >> 6, rho = tid & 63;
    const int st_row = lane * 32;                                   // this thread's staging row (floats)
    const int oc = lane & 7, og = lane >> 3;                        // copy-out role: chunk, row-in-quad

    const long long total_units = (long long)P.n_cpi * UPC;
    const long long ustride = (long long)gridDim.x * Gm::GROUPS;
    auto load_H_async = [&](long long u) {
        if (u < total_units) {
            const c32 *Hg = P.H + (u / UPC) * 512;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int e = 2 * (tid + 128 * j);          // 16-byte chunk = 2 complex
                cp_async16(Hs + (e >> 6) * HST + (e & 63), Hg + e);
            }
        }
        cp_async_commit();
    };
    // front end of unit u: H tile (prefetched) -> range pass 1 -> range pass 2 -> A tile `abuf` (hi | lo)
    auto front = [&](long long u, int abuf) {
        const int q0 = (int)(u % UPC) * 2 + s_sl;
        unsigned char *At = sA + abuf * 16384;
        cp_async_wait_all();
        group_sync();
        {
            c32 u8[8];
#pragma unroll
            for (int k1 = 0; k1 < 8; k1++) u8[k1] = Hs[p_ch * HST + lo + 8 * k1];
#pragma unroll
            for (int k1 = 1; k1 < 8; k1++) u8[k1] = cmul_fma(u8[k1], __ldg(P.tw1g + q0 * 8 + k1));
            JRC_FFT8<1>(u8);
#pragma unroll
            for (int m0 = 0; m0 < 8; m0++) Ys[p_ch * HST + skew(lo, m0)] = u8[m0];
        }
        group_sync();
        load_H_async(u + ustride);     // the H tile is free again: prefetch the next unit's channel estimates
        {
            c32 u8[8];
#pragma unroll
            for (int k0 = 0; k0 < 8; k0++) u8[k0] = Ys[p_ch * HST + skew(k0, lo)];
#pragma unroll
            for (int k0 = 1; k0 < 8; k0++) u8[k0] = cmul_fma(u8[k0], __ldg(P.tw2g + (q0 * 8 + k0) * 8 + lo));
            JRC_FFT8<1>(u8);
#pragma unroll
            for (int m1 = 0; m1 < 8; m1++) {
                c32 hi = mk(__uint_as_float(__float_as_uint(u8[m1].x) & 0xFFFFE000u),
                            __uint_as_float(__float_as_uint(u8[m1].y) & 0xFFFFE000u));
                c32 lw = mk(__fsub_rn(u8[m1].x, hi.x), __fsub_rn(u8[m1].y, hi.y));
                *reinterpret_cast<c32 *>(At + a_hi + m1 * 1024) = hi;
                *reinterpret_cast<c32 *>(At + a_lo + m1 * 1024) = lw;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // A tile -> visible to the tensor core
    };
    // angle DFT of tile `abuf` on the tensor core, 3xTF32 (one thread per group issues)
    auto issue_mma = [&](int abuf) {
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t adesc = adesc0 + (uint64_t)(abuf * (16384 >> 4));
            // k-step kk of an operand = +32 bytes inside the swizzle atom = +2 in the descriptor address field
            umma_tf32(tmem, adesc + 0, bdesc + 0, idesc, 0);   // Ahi * Bhi
            umma_tf32(tmem, adesc + 2, bdesc + 2, idesc, 1);
            umma_tf32(tmem, adesc + 4, bdesc + 0, idesc, 1);   // Alo * Bhi
            umma_tf32(tmem, adesc + 6, bdesc + 2, idesc, 1);
            umma_tf32(tmem, adesc + 0, bdesc + 4, idesc, 1);   // Ahi * Blo
            umma_tf32(tmem, adesc + 2, bdesc + 6, idesc, 1);
            umma_commit(mbar_a);
        }
    };
    // epilogue of unit u: one map row per thread out of TMEM; staging in the consumed A tile `abuf`
    auto epilogue = [&](long long u, int abuf) {
        const int cpi = (int)(u / UPC), sg = (int)(u % UPC);
        float *stg = reinterpret_cast<float *>(sA + abuf * 16384) + warp * 1024;   // 32 rows x 32 floats per warp
        const int n_row = Q * (rho >> 3) + IR * (rho & 7) + sg * 2 + r_sl;
        float best = -1.f;
        // tile row warp*32 + 4j + og -> map row Q*((warp&1)*4 + j/2) + IR*(4*(j&1) + og) + q0(slice warp/2)
        float *out_l = P.map + ((long long)cpi * NR + Q * ((warp & 1) * 4) + IR * og + sg * 2 + (warp >> 1)) * NA + oc * 4;
#pragma unroll 1
        for (int half = 0; half < NA / 32; half++) {       // 32 angle bins (64 D columns) per pass
            float mg[32];
            {
                uint32_t v0[32], v1[32];
                const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(half * 64);
                tmem_ld32_nowait(ta, v0);
                tmem_ld32_nowait(ta + 32, v1);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float re0 = __uint_as_float(v0[2 * i]), im0 = __uint_as_float(v0[2 * i + 1]);
                    const float re1 = __uint_as_float(v1[2 * i]), im1 = __uint_as_float(v1[2 * i + 1]);
                    mg[i] = __fmaf_rn(re0, re0, __fmul_rn(im0, im0));
                    mg[16 + i] = __fmaf_rn(re1, re1, __fmul_rn(im1, im1));
                }
            }
            float m = mg[0];
#pragma unroll
            for (int i = 1; i < 32; i++) m = fmaxf(m, mg[i]);
            best = fmaxf(best, m);     // the angle bin is recovered from the map row by k_tc_finalize
            // staging: chunk c of row `lane` goes to physical chunk c ^ (lane & 7) -> conflict-free 16-byte stores
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 8; c++)
                *reinterpret_cast<float4 *>(stg + st_row + ((c ^ (lane & 7)) << 2)) = make_float4(mg[4 * c], mg[4 * c + 1], mg[4 * c + 2], mg[4 * c + 3]);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int rr = 4 * j + og;                       // staging row = TMEM lane within this warp
                const float4 o = *reinterpret_cast<const float4 *>(stg + rr * 32 + ((oc ^ (rr & 7)) << 2));
                __stcs(reinterpret_cast<float4 *>(out_l + (Q * (j >> 1) + IR * 4 * (j & 1)) * NA + half * 32), o);
            }
        }
        if (P.keys) {   // arg-max of the unit -> one atomic per group
            unsigned long long key = best >= 0.f ? pack_key(best, (unsigned)n_row) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                key = other > key ? other : key;
            }
            if (lane == 0 && key) atomicMax(P.keys + cpi, key);
        }
    };

    // ---- software pipeline per group: front(k+1) overlaps the MMAs of k -----------------------------
    long long unit = (long long)blockIdx.x * Gm::GROUPS + grp;
    load_H_async(unit);
    if (unit < total_units) {
        front(unit, 0);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        group_sync();
        issue_mma(0);
        int buf = 0;
        while (true) {
            const long long nxt = unit + ustride;
            const bool has_next = nxt < total_units;
            if (has_next) front(nxt, buf ^ 1);
            mbar_wait(mbar_a, phase);          // D holds unit `unit`; its A tile is consumed
            phase ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            epilogue(unit, buf);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            group_sync();                      // D and the staging tile are free; the next A tile is complete
            if (!has_next) break;
            issue_mma(buf ^ 1);
            unit = nxt;
            buf ^= 1;
        }
    }
    __syncthreads();
    if (warp_cta == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Gm::TMEM_COLS));
}

// One warp per CPI: detection record from the per-CPI key.  The key holds the |.|^2 map value of the
// peak and its linear index; the noise window (lib/range_angle_estimator_impl.cc:197-226) is read back
// from the map the main kernel wrote.
template <int IR, int IA>
__global__ void __launch_bounds__(128) k_tc_finalize(const TcParams P)
{
    using Gm = TcGeom<IR, IA>;
    constexpr int NR = Gm::NR, NA = Gm::NA;
    const int lane = threadIdx.x & 31;
    const int cpi = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (cpi >= P.n_cpi) return;
    const unsigned long long key = P.keys[cpi];
    if (key == 0ull) {
        if (lane == 0) {
            DetDev d; d.range_idx = -1; d.angle_idx = -1; d.peak_power = -1.f;
            d.noise_power = __int_as_float(0x7fc00000); d.snr_db = d.noise_power;
            d.n_noise = 0; d.flags = 0; d.cpi = P.cpi0 + cpi;
            P.dets[cpi] = d;
        }
        return;
    }
    const float peak = __uint_as_float((unsigned)(key >> 32));
    const int nstar = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
    const float *map_c = P.map + (long long)cpi * NR * NA;
    // first angle bin of row nstar that holds the maximum (the key keeps the earliest row among equal values)
    int istar = 0x7fffffff;
    for (int i = lane; i < NA; i += 32)
        if (__ldcg(map_c + (long long)nstar * NA + i) == peak) { istar = i; break; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) istar = min(istar, __shfl_xor_sync(0xffffffffu, istar, o));
    if (istar == 0x7fffffff) istar = 0;      // cannot happen: the key was built from this row
    const NoiseWin w = noise_window(P.est, nstar, istar);
    const int ncols = w.end_a - w.start_a, nrows = w.end_r - w.start_r;
    const int total = (ncols > 0 && nrows > 0) ? nrows * ncols : 0;
    double acc = 0.0;
    for (int j0 = lane; j0 < total; j0 += 32 * 8) {      // 8 independent loads in flight per lane
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int j = j0 + 32 * q;
            v[q] = 0.f;
            if (j < total) {
                const int ir = w.start_r + j / ncols, ia = w.start_a + j % ncols;
                const int r_idx = ((ir % NR) + NR) % NR, a_idx = ((ia % NA) + NA) % NA;
                v[q] = __ldcg(map_c + (long long)r_idx * NA + a_idx);
            }
        }
#pragma unroll
        for (int q = 0; q < 8; q++) acc += (double)v[q];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        DetDev d;
        d.range_idx = nstar; d.angle_idx = istar; d.peak_power = peak; d.n_noise = total;
        d.noise_power = __fdiv_rn((float)acc, (float)total);
        d.snr_db = __fmul_rn(10.f, log10f(__fdiv_rn(d.peak_power, d.noise_power)));
        d.flags = (d.snr_db >= P.est.snr_threshold && d.peak_power >= P.est.power_threshold) ? 1u : 0u;
        d.cpi = P.cpi0 + cpi;
        P.dets[cpi] = d;
    }
}

}  // namespace jrc
