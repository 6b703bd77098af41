// jrc_tcfused.cuh -- the fused radar chain of jrc_fused.cuh with the angle DFT on the 5th-generation tensor cores
// (tcgen05.mma, accumulators and the A operand in tensor memory) and the front end overlapped with the store stream.
//
// BASELINE.json: "Tensor cores are used only if ncu shows the small-N angle DFT performs better when expressed as a
// complex GEMM".  profiles/ (round 1) showed that it does once the A operand lives in TMEM: the SIMT angle pass costs
// ~0.15 ms of issue slots per 4096 CPIs on top of a 0.171 ms store stream, the tensor-core form below leaves the SIMT
// pipes the hi/lo split, |.|^2 and the stores.
//
//   mimo_ofdm_radar conj-MAC (lib/mimo_ofdm_radar_impl.cc:250-274), range zero-pad + fft_vcc IFFT (:312-315;
//   ...radar_sim.grc:940-962)                       -> crew A: 4 warps, SIMT, pruned radix 8 x 8 x IR as in k_fused64x8
//   matrix_transpose + angle zero-pad + fft_vcc FFT + fftshift (lib/matrix_transpose_impl.cc:97-104,
//   ...radar_sim.grc:963-985), complex_to_mag_squared (:637-652)
//                                                   -> crew B: 3 groups of 4 warps, one 128-range-bin tile at a time:
//        M[n][i] = sum_p (-1)^p e^{-j 2 pi p i / Na} y[p][n]   as   D[128 x 2Na] = A[128 x 16] * B[2Na x 16]^T,
//        A row n = (Re y[0][n], Im y[0][n], ..., Im y[7][n]) written by the thread that owns row n to TMEM lane n
//        (tcgen05.st), B = the DFT matrix (shared memory, 128-byte swizzle), 3xTF32:
//        D = Ahi*Bhi + Alo*Bhi + Ahi*Blo (Ahi = top 19 bits) -- six tcgen05.mma of K = 8 per tile, error ~2^-21 --
//        then tcgen05.ld, re^2 + im^2, running maximum, padded staging, 64-byte-per-row coalesced st.global.cs.
//
// One CTA of 512 threads per SM (a kernel that allocates TMEM is not co-scheduled).  The range spectra y of a CPI are
// 64 KiB; two buffers: crew A builds y[k+1] while crew B streams CPI k.  mbarriers: full[b] (A -> B), free[b] (B -> A).
// Layout of a y buffer: row n = 4 float4 (one per channel pair), 16-byte slots XOR-swizzled over two rows so that the
// range passes (lanes along n) and the tile loads (lanes along n) are conflict-free; the range passes run in place.
#pragma once
#include <cstdint>
#include "jrc_common.cuh"
#include "jrc_staged.cuh"
#include "jrc_fused.cuh"

namespace jrc {

struct TcFusedParams {
    PortDev rx, tx;
    const c32 *H;              // [n_cpi][8][64] channel estimates instead of symbols (background path), or nullptr
    int n_cpi, cpi0;
    int T, R, S, n_pre, tx_interleave;
    float *map;                // [n_cpi][NR][64]
    const float *bimg;         // [128][32] tf32 [Bhi | Blo] rows of the angle DFT matrix, 128-byte-swizzled image
    int dbg;                   // measurement switches (JRC_TC_DBG): 1 skip the range passes, 2 skip the map stores
};

template <int IR>
struct TcFusedGeom {
    static constexpr int NSC = 64, V = 8, NA = 64;
    static constexpr int NR = NSC * IR, Q = NR / 8, TILES = NR / 128;
    static constexpr int THREADS = 512, PROD = 128, GROUPS = 3;
    static constexpr int STG_ROW = 272;                                   // a 256-byte map row + 16 B pad: conflict-free STS.128 / LDS.128
    static constexpr int OFF_B = 0;                                       // 16 KiB
    static constexpr int OFF_Y = 16384;                                   // 2 x NR x 64 B
    static constexpr int OFF_STG = OFF_Y + 2 * NR * 64;                   // GROUPS x 64 rows x STG_ROW (half a tile at a time)
    static constexpr int OFF_HS = OFF_STG + GROUPS * 64 * STG_ROW;        // 2 x [4 pairs][64] float4
    static constexpr int OFF_TW = OFF_HS + 2 * 4 * 64 * 16;               // [8][Q] c32
    static constexpr int OFF_IN = OFF_TW + 8 * Q * 8;                     // [(T+R)][S][64] c32
    static size_t smem_bytes(int T, int R, int S) { return (size_t)OFF_IN + (size_t)(T + R) * S * 64 * 8 + 1024; }
    static_assert(TILES >= GROUPS, "every group needs a tile of every CPI");
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr)      // K-major, 128-byte swizzle, 8-row atoms of 1024 B
{
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N)         // kind::tf32, fp32 accumulate, A and B K-major
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t mbar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

#define JRC_TMEM_LD16(taddr, v)                                                                                                    \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"          \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),   \
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                                    \
                 : "r"(taddr))
#define JRC_TMEM_ST32(taddr, v)                                                                                                    \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                                   \
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), \
                   "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),            \
                   "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),            \
                   "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory")

// float4 slot of (row n, channel pair j) in a y buffer: two rows form a 128-byte unit whose eight 16-byte slots are
// XORed with the unit index
__device__ __forceinline__ int yslot(int n, int j) { return ((n >> 1) << 3) + ((((n & 1) << 2) | j) ^ ((n >> 1) & 7)); }

template <int IR>
__global__ void __launch_bounds__(512, 1) k_fused_tc(const TcFusedParams P)
{
    using Gm = TcFusedGeom<IR>;
    constexpr int NR = Gm::NR, Q = Gm::Q, TILES = Gm::TILES, NA = Gm::NA;
    extern __shared__ unsigned char smem_tc[];
    unsigned char *base = smem_tc + ((1024u - (smem_u32(smem_tc) & 1023u)) & 1023u);
    __shared__ uint64_t mbar_full[2], mbar_free[2], mbar_mma[Gm::GROUPS];
    __shared__ uint32_t tmem_slot;
    float4 *ybuf = reinterpret_cast<float4 *>(base + Gm::OFF_Y);            // [2][NR * 4]
    float4 *Hs = reinterpret_cast<float4 *>(base + Gm::OFF_HS);             // [2][4][64]
    c32 *tw2t = reinterpret_cast<c32 *>(base + Gm::OFF_TW);                 // [8][Q] W_NR^{k0 q}
    c32 *inb = reinterpret_cast<c32 *>(base + Gm::OFF_IN);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

    // ---- one-time set-up ----
    for (int e = tid; e < 128 * 32; e += 512) reinterpret_cast<float *>(base + Gm::OFF_B)[e] = P.bimg[e];
    for (int e = tid; e < 8 * Q; e += 512) tw2t[e] = cispi_ratio(2 * (e / Q) * (e % Q), NR);
    if (tid == 0) {
        for (int b = 0; b < 2; b++) { mbar_init(smem_u32(&mbar_full[b]), 1); mbar_init(smem_u32(&mbar_free[b]), Gm::GROUPS); }
        for (int g = 0; g < Gm::GROUPS; g++) mbar_init(smem_u32(&mbar_mma[g]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // the B image is read by the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int n_local = (P.n_cpi - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // CPIs of this CTA: blockIdx.x + i * gridDim.x

    if (warp < 4) {
        // =====================================================================================
        // crew A: conj-MAC + range passes of CPI i into y[i & 1]; one warp per channel pair
        // =====================================================================================
        const int pair = warp;
        const int per_ant = P.S * 64;
        const int q0 = lane % IR;
        c32 tw1[8];
#pragma unroll
        for (int j = 0; j < 8; j++) tw1[j] = cispi_ratio(2 * j * q0, Q);
        const c32 *srx[2], *stx[2];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int p = 2 * pair + c;
            int ch_r, ch_t;
            if (P.tx_interleave) { ch_t = p / P.R; ch_r = p - ch_t * P.R; } else { ch_r = p / P.T; ch_t = p - ch_r * P.T; }
            srx[c] = inb + (P.T + ch_r) * per_ant + lane;
            stx[c] = inb + ch_t * per_ant + lane;
        }
        auto prefetch = [&](int cpi) {
            const int cpa = per_ant >> 1, total = (P.T + P.R) * cpa;
            const int sh = (cpa & (cpa - 1)) ? -1 : 31 - __clz(cpa);
            for (int c = tid; c < total; c += Gm::PROD) {
                const int a = sh >= 0 ? (c >> sh) : c / cpa, w = c - a * cpa;
                const c32 *src = (a < P.T)
                    ? P.tx.base + (long long)cpi * P.tx.cpi_stride + (long long)a * P.tx.ant_stride
                    : P.rx.base + (long long)cpi * P.rx.cpi_stride + (long long)(a - P.T) * P.rx.ant_stride;
                cp_async16(inb + a * per_ant + 2 * w, src + (long long)P.n_pre * 64 + 2 * w);
            }
            cp_async_commit();
        };
        if (!P.H && n_local > 0) prefetch(blockIdx.x);
        for (int i = 0; i < n_local; i++) {
            const int cpi = blockIdx.x + i * gridDim.x, b = i & 1;
            float4 *Hw = Hs + (b * 4 + pair) * 64;
            float4 *y = ybuf + b * (NR * 4);
            // ---- stage 1: channel estimates of this pair, subcarriers lane and lane + 32 ----
            if (P.H) {
                const c32 *Hg = P.H + (long long)cpi * 512 + (2 * pair) * 64 + lane;
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    const c32 h0 = Hg[32 * hh], h1 = Hg[64 + 32 * hh];
                    Hw[32 * hh + lane] = make_float4(h0.x, h0.y, h1.x, h1.y);
                }
            } else {
                cp_async_wait_all();
                named_bar(1, Gm::PROD);                  // the symbols of this CPI are visible to the crew
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    c32 acc0 = mk(0.f, 0.f), acc1 = mk(0.f, 0.f);
                    for (int s = 0; s < P.S; s++) {
                        const c32 x0 = srx[0][s * 64 + 32 * hh], c0 = stx[0][s * 64 + 32 * hh];
                        const c32 x1 = srx[1][s * 64 + 32 * hh], c1 = stx[1][s * 64 + 32 * hh];
                        acc0 = cadd_exact(acc0, cmul_exact(x0, mk(c0.x, -c0.y)));
                        acc1 = cadd_exact(acc1, cmul_exact(x1, mk(c1.x, -c1.y)));
                    }
                    Hw[32 * hh + lane] = make_float4(acc0.x, acc0.y, acc1.x, acc1.y);
                }
                named_bar(1, Gm::PROD);                  // every warp is done with the symbol buffer
                if (i + 1 < n_local) prefetch(cpi + gridDim.x);
            }
            __syncwarp();
            // the buffer must have been streamed out by crew B (CPI i - 2)
            if (i >= 2) mbar_wait(smem_u32(&mbar_free[b]), ((i >> 1) - 1) & 1);
            // ---- stage 2: range pass 1 (pruned: 8 of Q inputs non-zero), tasks (k0, q0) ----
            if (!(P.dbg & 1) || i < 2) {
#pragma unroll
            for (int j = 0; j < (8 * IR) / 32; j++) {
                const int k0 = (lane + 32 * j) / IR;
                c32 u0[8], u1[8];
#pragma unroll
                for (int k1 = 0; k1 < 8; k1++) {
                    const float4 t = Hw[k0 + 8 * k1];
                    u0[k1] = mk(t.x, t.y);
                    u1[k1] = mk(t.z, t.w);
                }
#pragma unroll
                for (int k1 = 1; k1 < 8; k1++) { u0[k1] = cmul_fma(u0[k1], tw1[k1]); u1[k1] = cmul_fma(u1[k1], tw1[k1]); }
                JRC_FFT8<1>(u0);
                JRC_FFT8<1>(u1);
#pragma unroll
                for (int m0 = 0; m0 < 8; m0++)
                    y[yslot(k0 * Q + q0 + IR * m0, pair)] = make_float4(u0[m0].x, u0[m0].y, u1[m0].x, u1[m0].y);
            }
            __syncwarp();
            // ---- stage 3: range pass 2, in place; the result is the A row layout (Re, Im of channels 2 pair, 2 pair + 1) ----
#pragma unroll
            for (int j = 0; j < Q / 32; j++) {
                const int q = lane + 32 * j;
                c32 u0[8], u1[8];
#pragma unroll
                for (int k0 = 0; k0 < 8; k0++) {
                    const float4 t = y[yslot(k0 * Q + q, pair)];
                    u0[k0] = mk(t.x, t.y);
                    u1[k0] = mk(t.z, t.w);
                }
#pragma unroll
                for (int k0 = 1; k0 < 8; k0++) {
                    const c32 w = tw2t[k0 * Q + q];
                    u0[k0] = cmul_fma(u0[k0], w);
                    u1[k0] = cmul_fma(u1[k0], w);
                }
                JRC_FFT8<1>(u0);
                JRC_FFT8<1>(u1);
#pragma unroll
                for (int m1 = 0; m1 < 8; m1++)
                    y[yslot(m1 * Q + q, pair)] = make_float4(u0[m1].x, u0[m1].y, u1[m1].x, u1[m1].y);
            }
            }
            named_bar(1, Gm::PROD);                      // all four channel pairs of y[b] are written
            if (tid == 0) mbar_arrive(smem_u32(&mbar_full[b]));
        }
    } else {
        // =====================================================================================
        // crew B: group g streams the tiles tg = g, g + 3, ... of this CTA's CPIs (tile tg: CPI tg / TILES)
        // =====================================================================================
        const int g = (warp - 4) >> 2, r = tid & 127;                   // r: row of the tile = TMEM lane
        unsigned char *stg = base + Gm::OFF_STG + g * (64 * Gm::STG_ROW);
        const uint32_t tmem_d = tmem_slot + (uint32_t)(g * 160), tmem_a = tmem_d + 128u;
        const uint32_t lane_off = (uint32_t)(((tid >> 5) & 3) * 32) << 16;
        const uint32_t mbar = smem_u32(&mbar_mma[g]);
        const uint32_t idesc = umma_idesc_tf32(128, 2 * NA);
        const uint64_t db = umma_desc_sw128(smem_u32(base + Gm::OFF_B));
        uint32_t mma_parity = 0;
        const int n_tiles = n_local * TILES;
        for (int tg = g; tg < n_tiles; tg += Gm::GROUPS) {
            const int i = tg / TILES, t = tg - i * TILES, b = i & 1;
            const int cpi = blockIdx.x + i * gridDim.x;
            mbar_wait(smem_u32(&mbar_full[b]), (i >> 1) & 1);
            // this thread's row of range spectra, split into tf32 hi / lo
            const float4 *y = ybuf + b * (NR * 4);
            uint32_t a[32];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float4 v = y[yslot(128 * t + r, j)];
                const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t hi = __float_as_uint(f[q]) & 0xFFFFE000u;
                    a[4 * j + q] = hi;
                    a[16 + 4 * j + q] = __float_as_uint(__fsub_rn(f[q], __uint_as_float(hi)));
                }
            }
            JRC_TMEM_ST32(tmem_a + lane_off, a);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            named_bar(2 + g, 128);
            if (r == 0) {
                // the next tile of this group belongs to another CPI: this one's y buffer is no longer needed by the group
                if (tg + Gm::GROUPS >= (i + 1) * TILES) mbar_arrive(smem_u32(&mbar_free[b]));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // D = Ahi*Bhi + Alo*Bhi + Ahi*Blo; K = 16 = 2 steps of 8; B tile rows: [Bhi (16) | Blo (16)] tf32
                const uint32_t acol[6] = {0, 8, 16, 24, 0, 8};
                const uint32_t bcol[6] = {0, 8, 0, 8, 16, 24};
#pragma unroll
                for (int k = 0; k < 6; k++) {
                    const uint64_t bd = db + (uint64_t)((bcol[k] * 4) >> 4);
                    const uint32_t acc = k > 0;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a + acol[k]), "l"(bd),
                                 "r"(idesc), "r"(acc) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
            }
            mbar_wait(mbar, mma_parity);
            mma_parity ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // epilogue, half a tile (64 rows = warps 2 half, 2 half + 1 of the group) at a time: the owners of the rows read
            // their accumulators 16 angle bins at a time -- Re columns [16h, +16), Im columns [64 + 16h, +16) --, form
            // re^2 + im^2 and stage the map rows; then all 128 threads stream the 16 KiB out as whole 256-byte rows
            float4 *dst = reinterpret_cast<float4 *>(P.map + ((long long)cpi * NR + 128 * t) * NA);
#pragma unroll 1
            for (int half = 0; half < 2; half++) {
                if ((r >> 6) == half) {
                    float4 *srow = reinterpret_cast<float4 *>(stg + (r & 63) * Gm::STG_ROW);
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        uint32_t re[16], im[16];
                        JRC_TMEM_LD16(tmem_d + lane_off + (uint32_t)(h * 16), re);
                        JRC_TMEM_LD16(tmem_d + lane_off + (uint32_t)(64 + h * 16), im);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            float2 o[2];
#pragma unroll
                            for (int q = 0; q < 2; q++) {
                                const float2 rr = make_float2(__uint_as_float(re[4 * j + 2 * q]), __uint_as_float(re[4 * j + 2 * q + 1]));
                                const float2 ii = make_float2(__uint_as_float(im[4 * j + 2 * q]), __uint_as_float(im[4 * j + 2 * q + 1]));
                                o[q] = __ffma2_rn(ii, ii, __fmul2_rn(rr, rr));
                            }
                            srow[4 * h + j] = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
                        }
                    }
                }
                named_bar(2 + g, 128);
                // 64 rows x 256 B, contiguous in the map: 16 lanes per row, two rows per store instruction
                float4 *d2 = dst + half * (64 * NA / 4);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int e = k * 128 + r, row = e >> 4, ch = e & 15;
                    const float4 vv = *reinterpret_cast<const float4 *>(stg + row * Gm::STG_ROW + ch * 16);
                    if (!(P.dbg & 2) || vv.x == 123.456f) __stcs(d2 + e, vv);
                }
                named_bar(2 + g, 128);                   // the staging rows are rewritten by the other half / the next tile
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512));
}

}  // namespace jrc
