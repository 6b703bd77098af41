// jrc_wide.cuh -- BASELINE configs[4] (8 TX x 16 RX = 128 virtual channels, 2048 subcarriers, no zero-pads: a 2048 x 128
// map): the chain as TWO kernels that move the symbols, one intermediate and the map -- not the channel estimates
// (2 MiB per CPI) nor the transposed range spectra (2 MiB per CPI) the block-by-block flow writes and re-reads.
//
// The range IFFT (over subcarriers k) and the angle FFT (over virtual channels p) are the two axes of one separable
// 2-D transform of H[p][k], so they may run in either order.  The channel estimate is local in k -- H[.][k] only needs
// the 24 antennas' symbols at subcarrier k (lib/mimo_ofdm_radar_impl.cc:250-274) -- hence:
//
//   k_wide_mac_angle   per (CPI, block of 16 subcarriers): symbols -> shared memory (two TMA tensor tiles on an mbarrier,
//                      double buffered; cp.async when a layout has no tensor map), conj-MAC of all 128 channels (register
//                      tiled: 4 RX x 2 TX per thread), angle FFT + fftshift across the channels for each subcarrier
//                      (matrix_transpose + fft_vcc #B, lib/matrix_transpose_impl.cc:97-104, ...radar_sim.grc:963-985),
//                      result G[cpi][angle bin][k] leaves as one TMA tensor store of whole 128-byte lines.
//   k_wide_range_mag   per (CPI, block of 8 angle bins): range IFFT over k of 2 x 4 rows of G (fft_vcc #A,
//                      ...radar_sim.grc:940-962), |.|^2 (:637-652); a thread ends up with the same range bins of all eight
//                      rows, so map[n][a0..a0+7] leaves as one 32-byte store (a whole sector) straight from registers,
//                      arg-max partials as in k_angle_mag.
//
// G (2 MiB per CPI) is produced and consumed in rounds of CPIs (222: both grids filled exactly).  Measured, it does not
// stay in the L2 at that size and the pair is not bandwidth-bound: HBM per CPI is 2 MiB symbols + 2 x 2 MiB
// G + 1 MiB map against 12 MiB for chan_est -> range -> angle.
// Float32 arithmetic of the same class as the oracle's radix-2 FFTs, not its rounding or order: the criterion is 1e-4 of
// the map peak; detection decisions are settled by jrc_exact.cuh.
#pragma once
#include <cuda.h>             // CUtensorMap (types only: the encoder is looked up at run time, no libcuda link)
#include "jrc_common.cuh"
#include "jrc_tiled.cuh"
#include "jrc_fused.cuh"      // cp_async16

namespace jrc {

struct WideParams {
    PortDev rx, tx;
    const c32 *H;               // channel estimates [n_cpi][128][N] instead of symbols (background path), or nullptr
    int T, R, S, n_pre, tx_interleave, n_cpi;
    c32 *G;                     // [n_cpi][128][N] angle spectra per subcarrier
    float *map;                 // [n_cpi][N][128]
    unsigned long long *keys;
    unsigned *sec;
    const c32 *tw_a;            // [128] w_128^i (forward)
    const c32 *tw_r;            // [N]   W_N^i  (inverse)
    int use_tma;                // symbols arrive as two TMA tiles per unit (tensor maps passed next to this struct)
    int use_tma_store;          // G leaves as two TMA tiles per unit
};

// ---- TMA + mbarrier (sm_100a): the symbol tiles of a unit are two bulk tensor copies issued by one thread ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// shared-memory tile -> tile of a 3-D tensor, tracked by the thread's bulk async-group
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tm, const void *src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(tm), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// tile of a 4-D tensor (coordinates innermost first) -> shared memory, completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *tm, unsigned long long *bar, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

template <int LOG2N>
struct WideGeom {
#ifndef JRC_WIDE_KB
#define JRC_WIDE_KB 16      // subcarriers per k_wide_mac_angle unit: 16 (256 threads, two CTAs per SM) or 32 (512 threads, one)
#endif
    static constexpr int N = 1 << LOG2N, V = 128, KB = JRC_WIDE_KB;
    static constexpr int CTAS_A = KB == 16 ? 2 : 1;                   // resident CTAs per SM of k_wide_mac_angle
    static constexpr int TA = 16 * KB;                                // threads of k_wide_mac_angle: 16 per subcarrier
    static constexpr int AB = 4;                                      // rows per pass of k_wide_range_mag
    static constexpr int UB = 8;                                      // angle bins per unit (two passes): one 32-byte sector of the map
    using GA = TiledGeom<7>;                                          // angle FFT rows: 16 threads per row, 16 rows per CTA
    using GR = TiledGeom<LOG2N>;
    static constexpr int MAX_ANT = 24, MAX_S = 8;
    // k_wide_mac_angle: two symbol buffers [(T+R)][S][KB] + H/FFT rows [KB][RS].  The staging tile of G lives in the symbol
    // buffer the conj-MAC has just consumed (the next prefetch into it waits for the tile's TMA store to have read it).
    static constexpr int BUF = MAX_ANT * MAX_S * KB;                  // c32 per symbol buffer (1024-byte multiple)
    static constexpr size_t SMEM_A = (size_t)2 * BUF * sizeof(c32) + (size_t)KB * GA::RS * sizeof(c32);
    static_assert((size_t)V * (KB + 1) * sizeof(c32) <= (size_t)BUF * sizeof(c32), "staging tile fits a symbol buffer");
    // k_wide_range_mag: AB rows of the transform
    static constexpr int RROW = fpad(N - 1) + 1 + 8;
    static constexpr size_t SMEM_B = (size_t)AB * RROW * sizeof(c32);
};

// ---------------------------------------------------------------------------
// conj-MAC + angle FFT, one (CPI, 16-subcarrier block) at a time
// ---------------------------------------------------------------------------
template <int LOG2N, int S_CT>      // S_CT: number of LTF symbols when known at compile time (0: run-time)
__global__ void __launch_bounds__(WideGeom<LOG2N>::TA, WideGeom<LOG2N>::CTAS_A) k_wide_mac_angle(const WideParams P, const __grid_constant__ CUtensorMap tm_rx,
                                                                            const __grid_constant__ CUtensorMap tm_tx,
                                                                            const __grid_constant__ CUtensorMap tm_g)
{
    using Gm = WideGeom<LOG2N>;
    using GA = typename Gm::GA;
    constexpr int N = Gm::N, V = Gm::V, KB = Gm::KB, RS = GA::RS, TA = Gm::TA, CPR = KB / 2;   // CPR: 16-byte chunks per row
    extern __shared__ __align__(1024) unsigned char smem_wide[];     // (the swizzle of the staging tiles is a function of the address)
    c32 *sym = reinterpret_cast<c32 *>(smem_wide);                        // [2][(T+R)*S][KB]
    c32 *rows = sym + 2 * Gm::BUF;                                        // [KB][RS]: H[.][k] then its angle transform
    // G block, angle bin major.  TMA-store form: two tiles [V][16 subcarriers] of 128-byte rows in the 128-byte swizzle
    // (16-byte chunk j of row a sits at chunk j ^ (a & 7)): the column writes of the angle FFT spread over the banks as
    // with a padded pitch, and the block leaves as two bulk tensor stores.  Otherwise [V][KB+1].
    // (the tile of a unit is the symbol buffer of that unit: stg below)
    const int tid = threadIdx.x;
    const int T = P.T, R = P.R, S = S_CT ? S_CT : P.S, per = (T + R) * S;  // antenna-symbol rows of KB subcarriers each
    constexpr int blocks_per_cpi = N / KB;
    const long long n_units = (long long)P.n_cpi * blocks_per_cpi;

    // angle FFT thread mapping (TiledGeom<7>): row = subcarrier, 16 threads per row
    const int lr_t = tid / GA::TPR, t = tid % GA::TPR;
    DifTw<7> Tw;
    Tw.load(P.tw_a, t, t, true);            // first-pass twiddles carry (-1)^j: output fftshift
    // conj-MAC thread mapping: subcarrier kk, RX block of 4, TX block of 2 (T = 8: 4 blocks) -> 8 channels x S products
    const int kk = tid % KB, rb = (tid / KB) & 3, tb = tid / (4 * KB);

    // Symbols of a unit -> shared memory.  TMA form: the tile [T][S][KB] of the TX packets and the tile [R][S][KB] of the RX
    // packets are one cp.async.bulk.tensor each (4-D tensor maps over (subcarrier, symbol, antenna, CPI), built by the host
    // for this call), issued by thread 0 and counted in bytes on the buffer's mbarrier.  cp.async form (odd strides,
    // JRC_WIDE_TMA=0): chunk i of this thread = 16 bytes (two subcarriers) of antenna-symbol row tid / CPR + (TA / CPR) i,
    // a plan fixed for the whole kernel; only the CPI and the subcarrier block change from unit to unit.
    __shared__ __align__(8) unsigned long long bar[2];
    const bool tma = P.use_tma != 0;
    constexpr int MAXCH = Gm::MAX_ANT * Gm::MAX_S * CPR / TA, RSTEP = TA / CPR;
    const c32 *src0[MAXCH];
    unsigned is_tx = 0;
    if (tma) {
        if (tid == 0) {
            mbar_init(&bar[0], 1);
            mbar_init(&bar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    } else {
#pragma unroll
        for (int i = 0; i < MAXCH; i++) {
            const int row = tid / CPR + RSTEP * i, ch = tid % CPR, ant = row / S, sy = row - ant * S;
            src0[i] = nullptr;
            if (row < per) {
                const bool txr = ant < T;
                src0[i] = (txr ? P.tx.base + (long long)ant * P.tx.ant_stride : P.rx.base + (long long)(ant - T) * P.rx.ant_stride) +
                          (long long)(P.n_pre + sy) * N + 2 * ch;
                is_tx |= txr ? (1u << i) : 0u;
            }
        }
    }
    auto prefetch = [&](long long unit, int buf) {
        const long long cpi = unit / blocks_per_cpi;
        const int k0 = (int)(unit % blocks_per_cpi) * KB;
        if (tma) {
            if (tid == 0) {
                c32 *dst = sym + (size_t)buf * Gm::BUF;
                // (the G tile staged in this buffer two units ago has been read by its store)
                if (P.use_tma_store) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mbar_expect_tx(&bar[buf], (unsigned)(per * KB * sizeof(c32)));
                tma_load_4d(dst, &tm_tx, &bar[buf], 2 * k0, 0, 0, P.tx.cpi_stride ? (int)cpi : 0);
                tma_load_4d(dst + T * S * KB, &tm_rx, &bar[buf], 2 * k0, 0, 0, (int)cpi);
            }
            return;
        }
        const long long otx = cpi * P.tx.cpi_stride + k0, orx = cpi * P.rx.cpi_stride + k0;
        c32 *dst = sym + (size_t)buf * Gm::BUF + (size_t)(tid / CPR) * KB + 2 * (tid % CPR);
#pragma unroll
        for (int i = 0; i < MAXCH; i++)
            if (src0[i]) cp_async16(dst + RSTEP * i * KB, src0[i] + ((is_tx >> i) & 1 ? otx : orx));
        cp_async_commit();
    };
    unsigned phase = 0;          // bit b: parity the next wait on bar[b] expects

    long long unit = blockIdx.x;
    int buf = 0;
    if (!P.H && unit < n_units) prefetch(unit, 0);
    for (; unit < n_units; unit += gridDim.x, buf ^= 1) {
        const int cpi = (int)(unit / blocks_per_cpi), k0 = (int)(unit % blocks_per_cpi) * KB;
        if (!P.H) {
            if (tma) {
                mbar_wait(&bar[buf], (phase >> buf) & 1u);
                phase ^= 1u << buf;
            } else {
                cp_async_wait_all();
            }
            __syncthreads();                                  // symbols of this unit visible; previous unit's stores done
            if (unit + gridDim.x < n_units) prefetch(unit + gridDim.x, buf ^ 1);
            // ---- conj-MAC: H[p][k0 + kk] for r in 4 rb .. +3, t in 2 tb .. +1 ----
            const c32 *sb = sym + (size_t)buf * Gm::BUF + kk;
            for (int r0 = 4 * rb; r0 < R; r0 += 16)
                for (int t0 = 2 * tb; t0 < T; t0 += 8) {
                    c32 acc[4][2];
#pragma unroll
                    for (int i = 0; i < 4; i++) { acc[i][0] = mk(0.f, 0.f); acc[i][1] = mk(0.f, 0.f); }
                    const c32 *pa = sb + (T + r0) * S * KB, *pb = sb + t0 * S * KB;
#pragma unroll 4
                    for (int s = 0; s < S; s++) {
                        c32 a[4], b[2];
#pragma unroll
                        for (int i = 0; i < 4; i++) a[i] = pa[(i * S + s) * KB];
#pragma unroll
                        for (int j = 0; j < 2; j++) b[j] = pb[(j * S + s) * KB];
#pragma unroll
                        for (int i = 0; i < 4; i++)
#pragma unroll
                            for (int j = 0; j < 2; j++) {       // a * conj(b)
                                acc[i][j].x = __fmaf_rn(a[i].x, b[j].x, __fmaf_rn(a[i].y, b[j].y, acc[i][j].x));
                                acc[i][j].y = __fmaf_rn(a[i].y, b[j].x, __fmaf_rn(-a[i].x, b[j].y, acc[i][j].y));
                            }
                    }
                    // channel p of (rx r0 + i, tx t0 + j): p0 + i * pi + j * pj  (lib/mimo_ofdm_radar_impl.cc:262-269)
                    const int pi = P.tx_interleave ? 1 : T, pj = P.tx_interleave ? R : 1, p0 = r0 * pi + t0 * pj;
                    c32 *hrow = rows + kk * RS;
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int j = 0; j < 2; j++) hrow[fpad(p0 + i * pi + j * pj)] = acc[i][j];
                }
        } else {
            __syncthreads();
            for (int e = tid; e < V * KB; e += TA) {
                const int p = e / KB, k = e % KB;
                rows[k * RS + fpad(p)] = P.H[((long long)cpi * V + p) * N + k0 + k];
            }
        }
        __syncthreads();
        c32 *stg = sym + (size_t)buf * Gm::BUF;              // this unit's symbols are consumed: its G tile is staged in their place
        // ---- angle FFT across the 128 channels of subcarrier lr_t (radix 8.8.2, fftshift folded in) ----
        {
            c32 *xrow = rows + lr_t * RS;
            c32 u[8];
#pragma unroll
            for (int m = 0; m < 8; m++) u[m] = xrow[fpad(t + m * GA::TPR)];
            dif_first_full<7, -1>(xrow, t, u, Tw, t & 1);     // in place: a thread writes the eight slots it has read
            __syncthreads();
            c32 o[8];
            dif_passes<7, -1, GA::WARP_SYNC, true>(xrow, t, Tw, o);
            if (P.use_tma_store) {
                const int half = lr_t >> 4, kc = (lr_t & 15) >> 1, ko = lr_t & 1;      // (KB = 16: one tile, half == 0)
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const int a = dif_freq<7>(8 * t + c);
                    stg[half * (V * 16) + a * 16 + (((kc ^ (a & 7)) << 1) | ko)] = o[c];
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic writes -> the async proxy's reads
            } else {
#pragma unroll
                for (int c = 0; c < 8; c++) stg[dif_freq<7>(8 * t + c) * (KB + 1) + lr_t] = o[c];
            }
        }
        __syncthreads();
        if (P.use_tma_store) {
            if (tid == 0) {
                tma_store_3d(&tm_g, stg, 2 * k0, 0, cpi);
                if (KB > 16) tma_store_3d(&tm_g, stg + V * 16, 2 * (k0 + 16), 0, cpi);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            // ---- G[cpi][a][k0 .. k0+31]: two 128-byte lines per angle bin ----
            for (int e = tid; e < V * KB; e += TA) {
                const int a = e / KB, k = e % KB;
                P.G[((long long)cpi * V + a) * N + k0 + k] = stg[a * (KB + 1) + k];
            }
        }
    }
    if (P.use_tma_store && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    cp_async_wait_all();
}

// ---------------------------------------------------------------------------
// range IFFT + |.|^2 + arg-max partials, one (CPI, 8 angle bins) at a time
// ---------------------------------------------------------------------------
template <int LOG2N>
__global__ void __launch_bounds__(TiledGeom<LOG2N>::THREADS, 2) k_wide_range_mag(const WideParams P)
{
    using Gm = WideGeom<LOG2N>;
    using GR = typename Gm::GR;
    constexpr int N = Gm::N, V = Gm::V, AB = Gm::AB, RROW = Gm::RROW, TPR = GR::TPR;
    static_assert(GR::RPC == 1, "one transform per pass of the CTA");
    extern __shared__ __align__(1024) unsigned char smem_wide[];
    c32 *rowbuf = reinterpret_cast<c32 *>(smem_wide);                     // [AB][RROW]
    const int tid = threadIdx.x, lane = tid & 31, t = tid % TPR;
    DifTw<LOG2N> Tw;
    Tw.load(P.tw_r, t, t);
    constexpr int UB = Gm::UB, HALVES = UB / AB;
    const int units_per_cpi = V / UB;
    const long long n_units = (long long)P.n_cpi * units_per_cpi;
    for (long long unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int cpi = (int)(unit / units_per_cpi), a0 = (int)(unit % units_per_cpi) * UB;
        // |.|^2 at this thread's 8 range bins for all UB angle bins: the AB rows of one pass at a time through the row
        // buffer, the values of the earlier pass waiting in registers -- the thread then owns whole 32-byte sectors of the
        // map (a 16-byte store leaves half a sector, which the L2 completes by READING the other half from HBM:
        // measured 2 MiB of fill reads per CPI for the 1 MiB map)
        float v[8][UB];
#pragma unroll
        for (int hf = 0; hf < HALVES; hf++) {
            const c32 *Gu = P.G + ((long long)cpi * V + a0 + hf * AB) * N;
            if (hf) __syncthreads();           // the rows of the previous pass have been read
            {
                // the rows of the NEXT pass (this unit's other half, or the first half of this CTA's next unit) on their way
                // from HBM into the L2 while this pass computes: one 128-byte line per lane and row covers the row's 16 KiB
                const long long nu = unit + gridDim.x;
                const c32 *Gn = hf + 1 < HALVES ? Gu + (long long)AB * N
                                                : (nu < n_units ? P.G + ((nu / units_per_cpi) * V + (nu % units_per_cpi) * UB) * (long long)N : nullptr);
                if (Gn && tid < 128)
#pragma unroll
                    for (int r = 0; r < AB; r++) asm volatile("prefetch.global.L2 [%0];" ::"l"(Gn + (long long)r * N + tid * 16));
            }
            // first pass of the AB rows straight from global memory (L2): 8 loads in flight per row and thread
#pragma unroll
            for (int r = 0; r < AB; r++) {
                c32 u[8];
#pragma unroll
                for (int m = 0; m < 8; m++) u[m] = __ldcg(Gu + (long long)r * N + t + m * TPR);
                dif_first_full<LOG2N, 1>(rowbuf + r * RROW, t, u, Tw);
            }
            __syncthreads();
            // the middle passes of all AB rows share their barriers; the last pass leaves the results in registers
            dif_mid_passes_rows<LOG2N, 1, AB>(rowbuf, RROW, t, Tw);
#pragma unroll
            for (int r = 0; r < AB; r++) {
                c32 o[8];
                dif_last_pass<LOG2N, 1>(rowbuf + r * RROW, t, o);
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const c32 sq = __fmul2_rn(o[c], o[c]);
                    v[c][hf * AB + r] = __fadd_rn(sq.x, sq.y);
                }
            }
        }
        // map[cpi][n][a0 .. a0+7] straight from registers, one 32-byte store per range bin; running maximum per range bin
        float best = -1.f, sec_t = -1.f;
        int best_row = 0;
        float *mp = P.map ? P.map + (long long)cpi * N * V + a0 : nullptr;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int n = dif_freq<LOG2N>(8 * t) + dif_freq<LOG2N>(c);
            if (mp)
                asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(mp + (long long)n * V), "f"(v[c][0]),
                             "f"(v[c][1]), "f"(v[c][2]), "f"(v[c][3]), "f"(v[c][4]), "f"(v[c][5]), "f"(v[c][6]), "f"(v[c][7])
                             : "memory");
            // largest and runner-up of the eight, then whatever loses against the running maximum
            float hi = v[c][0], lo = -1.f;
#pragma unroll
            for (int j = 1; j < UB; j++) {
                lo = fmaxf(lo, fminf(hi, v[c][j]));
                hi = fmaxf(hi, v[c][j]);
            }
            sec_t = fmaxf(sec_t, lo);
            sec_t = fmaxf(sec_t, fminf(hi, best));
            if (hi > best || (hi == best && n < best_row)) { best = hi; best_row = n; }
        }
        if (P.keys) {
            unsigned long long key = best >= 0.f ? pack_key(best, (unsigned)best_row) : 0ull;
            float b2 = sec_t;
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o2);
                const float ob2 = __shfl_xor_sync(0xffffffffu, b2, o2);
                b2 = fmaxf(fmaxf(b2, ob2), (key && other) ? fminf(__uint_as_float((unsigned)(key >> 32)), __uint_as_float((unsigned)(other >> 32))) : -1.f);
                key = other > key ? other : key;
            }
            if (lane == 0 && key) {
                const unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(P.keys + cpi);
                unsigned long long top = cur;
                float loser = -1.f;
                if (key > cur) {
                    const unsigned long long old = atomicMax(P.keys + cpi, key);
                    top = old > key ? old : key;
                    const unsigned long long lo = old > key ? key : old;
                    if (lo) loser = __uint_as_float((unsigned)(lo >> 32));
                } else {
                    loser = __uint_as_float((unsigned)(key >> 32));
                }
                loser = fmaxf(loser, b2);
                if (loser >= __uint_as_float((unsigned)(top >> 32)) * (1.f - 2.f * EPS_AMB)) atomicMax(P.sec + cpi, __float_as_uint(loser));
            }
        }
        __syncthreads();      // the rows are rewritten by the next unit
    }
}

}  // namespace jrc
