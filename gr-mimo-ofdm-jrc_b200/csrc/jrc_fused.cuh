// jrc_fused.cuh -- the fused radar chain for the 64-subcarrier / 8-virtual-channel
// family (the shipped configuration and BASELINE configs[0..1]):
//
//   mimo_ofdm_radar conj-MAC (lib/mimo_ofdm_radar_impl.cc:250-274)
//   -> range zero-pad + fft_vcc IFFT  (:243,:312-315; ...radar_sim.grc:940-962)
//   -> matrix_transpose + angle zero-pad (lib/matrix_transpose_impl.cc:97-104)
//   -> fft_vcc FFT + fftshift          (...radar_sim.grc:963-985)
//   -> complex_to_mag_squared          (...radar_sim.grc:637-652)
//   -> range_angle_estimator           (lib/range_angle_estimator_impl.cc:137-253)
//
// in ONE persistent kernel, one CTA per CPI at a time.  HBM sees the symbols once
// (cp.async prefetch of the next CPI), the |.|^2 map once (whole map rows streamed out of
// a bank-conflict-free per-warp staging tile) and a 32-byte detection record; neither
// zero-pad nor the transpose nor the complex map ever exist in memory.
//
// Index algebra (W_N = e^{+j2pi/N}, w_N = e^{-j2pi/N}, Q = Nr/8, Na = 8*IA):
//   range  y[p][Q*m1 + q] = sum_{k0<8} W_8^{k0 m1} * ( W_Nr^{k0 q} * B[p][k0][q] )
//          B[p][k0][q0 + IR*m0] = sum_{k1<8} W_8^{k1 m0} * ( W_Q^{k1 q0} * H[p][k0 + 8 k1] )
//   angle  M[n][b + IA*a] = sum_{p<8} w_8^{p a} * ( (-1)^p w_Na^{p b} * y[p][n] )
// i.e. three passes of "twiddle 8 inputs, 8-point DFT": the zero-padded inputs of
// both FFTs are pruned away analytically (only the 64 resp. 8 non-zero inputs are
// ever touched) and the output fftshift of the angle FFT is the (-1)^p factor.
// All twiddles are per-thread constants computed once per (persistent) CTA.
#pragma once
#include "jrc_common.cuh"
#include "jrc_staged.cuh"
#include "jrc_exact.cuh"

namespace jrc {

struct FusedParams {
    PortDev rx, tx;            // symbol inputs (unused when FROM_H)
    const c32 *H;              // [n_cpi][8][64] channel estimates (FROM_H: background removal path)
    int n_cpi, cpi0;
    int T, R, S, n_pre, tx_interleave;
    float *map;                // [n_cpi][NR][NA] or nullptr
    DetDev *dets;              // [n_cpi] or nullptr (in-kernel estimator)
    FixCtl *fix_ctl;           // records whose decision k_est_exact has to redo (jrc_exact.cuh)
    int *fix_list;
    EstParams est;
    const c32 *tw_r, *tw_a;    // the staged FFTs' twiddle tables [n/2] (inverse NR, forward NA): in-kernel arg-max resolution
    const int2 *win_tab;       // [NA] k_est_tables
    const double2 *g_tab;      // [NA][8]
};

template <int IR, int IA>
struct FusedGeom {
    static constexpr int NSC = 64, V = 8;
    static constexpr int NR = NSC * IR, NA = V * IA, Q = NR / 8;
    static constexpr int THREADS = 256, WARPS = 8;
    static constexpr int G = 32 / IA;                    // range bins (map rows) per tile
    static constexpr int TILE = G * NA;                  // 256 floats = 1 KiB: G whole map rows
    // angle pass: a thread evaluates rows n and n + Q together (split-complex packed arithmetic), so a
    // warp owns the row pairs {(2j'Q + q, (2j'+1)Q + q)} with j' = warp / 2, q in its half of [0, Q)
    static constexpr int PITERS = (Q / 2) / G;           // pair iterations per warp
    static constexpr int UNROLL = 2;
    static_assert(IA >= 4 && IA <= 32 && (IA & (IA - 1)) == 0, "angle interp must be 4..32, power of two");
    static_assert(IR >= 8 && IR <= 32 && (IR & (IR - 1)) == 0, "range interp must be 8..32, power of two");
    static_assert(PITERS >= UNROLL && PITERS % UNROLL == 0, "map too small");

    static size_t smem_bytes(int T, int R, int S, bool from_h)
    {
        size_t b = (size_t)V * NR * 8 + (size_t)V * NSC * 8 + (size_t)8 * Q * 8 + (size_t)WARPS * 2 * TILE * 4 +
                   64 + 512 + 64 + (size_t)NA * 8 * 8 + (size_t)NA * 8;      // ... + g_tab (float2) + win_tab copies
        if (!from_h) b += (size_t)(T + R) * S * NSC * 8;
        return b;
    }
};

__device__ __forceinline__ void cp_async16(void *sdst, const void *gsrc)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// Angle twiddles of one thread, duplicated into both halves of a register pair.
struct AngleTw { float2 r[8], i[8]; };   // (re,re), (im,im); index 0 unused

// One angle task: rows n0 = 2j'Q + q and n1 = n0 + Q, 8 channel samples each -> twiddle -> forward DFT-8
// -> |.|^2.  re/im/v: .x belongs to row n0, .y to row n1.  Pure _rn intrinsics: re-evaluating a row
// pair reproduces the main loop bit for bit.
// ys holds the range spectra as ys[pr * NR + (4c + j') * Q + q] = (Re y[p][n0], Re y[p][n1], Im y[p][n0],
// Im y[p][n1]) for channel p = 2 pr + c.
__device__ __forceinline__ void angle_pair(const float4 *__restrict__ ys, int NR, int Q, int jq, const AngleTw &tw,
                                           float2 (&re)[8], float2 (&im)[8], float2 (&v)[8])
{
#pragma unroll
    for (int p = 0; p < 8; p++) {
        const float4 t = ys[(p >> 1) * NR + 4 * (p & 1) * Q + jq];
        re[p] = mk(t.x, t.y);
        im[p] = mk(t.z, t.w);
    }
#pragma unroll
    for (int p = 1; p < 8; p++) {
        const float2 r = __ffma2_rn(mk(-im[p].x, -im[p].y), tw.i[p], __fmul2_rn(re[p], tw.r[p]));   // negation: operand modifier
        const float2 i = __ffma2_rn(im[p], tw.r[p], __fmul2_rn(re[p], tw.i[p]));
        re[p] = r;
        im[p] = i;
    }
    fft8s<-1>(re, im);
#pragma unroll
    for (int a = 0; a < 8; a++) v[a] = __ffma2_rn(im[a], im[a], __fmul2_rn(re[a], re[a]));
}

template <int IR, int IA, bool FROM_H, bool WRITE_MAP>
__global__ void __launch_bounds__(256, 2) k_fused64x8(const FusedParams P)
{
    using Gm = FusedGeom<IR, IA>;
    constexpr int NR = Gm::NR, NA = Gm::NA, Q = Gm::Q, G = Gm::G;
    constexpr int PITERS = Gm::PITERS, TILE = Gm::TILE;
    constexpr int T1 = (8 * IR) / 64;          // range pass 1 tasks per lane (channel pair = warp pair)
    constexpr int T2 = Q / 64;                 // range pass 2 tasks per lane
    static_assert(IR >= 8, "the channel-pair front end needs 8*IR >= 64 pass-1 tasks per pair");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *ys = reinterpret_cast<float4 *>(smem_raw);            // [4 pairs][NR][2]  B then y (in place)
    float4 *Hs = ys + 4 * NR;                                     // [4 pairs][64][2]
    c32 *tw2t = reinterpret_cast<c32 *>(Hs + 256);                // [8][Q]   W_Nr^{k0 q}
    float *stg = reinterpret_cast<float *>(tw2t + 8 * Q);         // [8 warps][2][TILE]
    unsigned long long *red = reinterpret_cast<unsigned long long *>(stg + 8 * 2 * TILE);   // [8]
    double *redA = reinterpret_cast<double *>(red + 8);           // [8 warps][8] partial lag sums
    int *sint = reinterpret_cast<int *>(redA + 64);               // [16] scalars
    float2 *gtab = reinterpret_cast<float2 *>(sint + 16);         // [NA][8] window column sums (k_est_tables), float
    int2 *wtab = reinterpret_cast<int2 *>(gtab + NA * 8);         // [NA]    window columns
    c32 *inb = reinterpret_cast<c32 *>(wtab + NA);                // [(T+R)][S][64]

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler

    // ---- per-thread constants (once per persistent CTA) --------------------
    // Stages 1-3 give every warp PAIR one pair of virtual channels (2j, 2j+1; j = warp / 2) and keep the
    // two channels interleaved in shared memory, so every access of the front end and the angle pass's
    // loads are 128 bit wide; the stages only need a 64-thread named barrier between them.
    // range pass 1: task (k0, q0), q0 = lane % IR fixed per lane;  twiddle W_Q^{k1 q0}
    // range pass 2: task q = 32 half + lane + 64 i;                 twiddle W_Nr^{k0 q} from tw2t
    // angle pass:   task (n, b);                                    twiddle (-1)^p w_Na^{p (b + IA*rot)}
    const int q0 = lane % IR;
    const int b = lane % IA, g = lane / IA, rot = g;
    c32 tw1[8];
    AngleTw tw3;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        tw1[j] = cispi_ratio(2 * j * q0, Q);
        const c32 t = cispi_ratio(j * (NA - 2 * (b + IA * rot)), NA);
        tw3.r[j] = mk(t.x, t.x);
        tw3.i[j] = mk(t.y, t.y);
    }
    for (int e = tid; e < 8 * Q; e += 256) tw2t[e] = cispi_ratio(2 * (e / Q) * (e % Q), NR);
    if (P.dets)       // the estimator's tables: the record's last step must not wait for global memory
        for (int e = tid; e < NA * 8; e += 256) {
            const double2 gv = P.g_tab[e];
            gtab[e] = make_float2((float)gv.x, (float)gv.y);
            if (e < NA) wtab[e] = P.win_tab[e];
        }
    const EstParams est = P.est;

    // channels of this warp pair -> (rx antenna, tx antenna)  (lib/mimo_ofdm_radar_impl.cc:262-269)
    const int per_ant = P.S * 64;
    const int pair = warp >> 1, half = warp & 1;
    const c32 *srx[2], *stx[2];
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const int p = 2 * pair + c;
        int ch_r, ch_t;
        if (P.tx_interleave) { ch_t = p / P.R; ch_r = p - ch_t * P.R; } else { ch_r = p / P.T; ch_t = p - ch_r * P.T; }
        srx[c] = inb + (P.T + ch_r) * per_ant + 32 * half + lane;
        stx[c] = inb + ch_t * per_ant + 32 * half + lane;
    }
    float4 *Hw = Hs + pair * 64;       // this pair's channel estimates
    float4 *yw = ys + pair * NR;       // this pair's range spectra
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory"); };

    // staging: slot a' of this thread holds angle bin b + IA*((a'+rot)&7); rotating by
    // the row group makes the 32 lanes of every st.shared hit 32 different banks
    float *wstg = stg + warp * 2 * TILE;   // tile 0: rows n0.., tile 1: rows n0 + Q..
    float *sp[8];
#pragma unroll
    for (int a = 0; a < 8; a++) sp[a] = wstg + g * NA + b + IA * ((a + rot) & 7);

    auto prefetch = [&](int cpi) {
        const int cpa = per_ant >> 1;   // 16-byte chunks per antenna row
        const int total = (P.T + P.R) * cpa;
        const int sh = (cpa & (cpa - 1)) ? -1 : 31 - __clz(cpa);
        for (int c = tid; c < total; c += 256) {
            const int a = sh >= 0 ? (c >> sh) : c / cpa, w = c - a * cpa;
            const c32 *src = (a < P.T)
                ? P.tx.base + (long long)cpi * P.tx.cpi_stride + (long long)a * P.tx.ant_stride
                : P.rx.base + (long long)cpi * P.rx.cpi_stride + (long long)(a - P.T) * P.rx.ant_stride;
            cp_async16(inb + a * per_ant + 2 * w, src + (long long)P.n_pre * 64 + 2 * w);
        }
        cp_async_commit();
    };

    // Detection record of the previous CPI: its last step only needs the shared scratch (not y), so it
    // runs after the next CPI's barrier (A) instead of costing a barrier of its own.
    // Lag d is accumulated by warp pair part(d): 0 -> {0}, 1 -> {1,7}, 2 -> {2,6}, 3 -> {3,4,5}.
    bool pending = false;
    bool amb_prev = false;   // block-uniform: a second group maximum lies within EPS_AMB of the CPI's maximum
    auto finalize = [&]() {
        if (tid < 8) {
            const int d = tid;
            const int part = d == 0 ? 0 : (d == 1 || d == 7) ? 1 : (d == 2 || d == 6) ? 2 : 3;
            const int slot = (d <= 3) ? 0 : (d == 4) ? 2 : (d == 5) ? 4 : 2;   // d=7 -> 2, d=6 -> 2
            const double *r0 = redA + (2 * part) * 8 + slot, *r1 = r0 + 8;
            const double ar = r0[0] + r1[0], ai = d ? r0[1] + r1[1] : 0.0;
            const int imin = sint[1];
            const float2 gd = gtab[imin * 8 + d];
            double contrib = 2.0 * ((double)gd.x * ar - (double)gd.y * ai);
            contrib += __shfl_xor_sync(0xffu, contrib, 4);
            contrib += __shfl_xor_sync(0xffu, contrib, 2);
            contrib += __shfl_xor_sync(0xffu, contrib, 1);
            if (tid == 0) {
                const int2 wa = wtab[imin];
                const int ncols = wa.y - wa.x, nrows = 2 * est.discard_range_idx;
                const int total = (ncols > 0 && nrows > 0) ? nrows * ncols : 0;
                const double s = contrib > 0.0 ? contrib : 0.0;
                DetDev dd;
                dd.range_idx = sint[0]; dd.angle_idx = imin;
                dd.peak_power = __int_as_float(sint[6]);
                dd.n_noise = total;
                dd.noise_power = __fdiv_rn(total > 0 ? (float)s : 0.f, (float)total);
                dd.snr_db = snr_db_fast(dd.peak_power, dd.noise_power);
                dd.flags = (dd.snr_db >= est.snr_threshold && dd.peak_power >= est.power_threshold) ? DET_PASSED : 0u;
                // decisions that FFT rounding could turn are not taken here (jrc_exact.cuh)
                if (amb_prev || sint[5]) dd.flags |= DET_PENDING | DET_AMB;
                if (gate_is_marginal(dd.peak_power, dd.noise_power, dd.snr_db, total, est.snr_threshold, est.power_threshold))
                    dd.flags |= DET_PENDING | DET_GATE;
                dd.cpi = P.cpi0 + sint[7];
                if (dd.flags & DET_PENDING) fix_push(P.fix_ctl, P.fix_list, sint[7]);
                P.dets[sint[7]] = dd;
            }
        }
    };

    int cpi = blockIdx.x;
    if (!FROM_H && cpi < P.n_cpi) {
        prefetch(cpi);
        cp_async_wait_all();
    }
    __syncthreads();

    for (; cpi < P.n_cpi; cpi += gridDim.x) {
        // The symbols of this CPI were published by barrier (E) of the previous one, and stage 1 does not
        // touch y: it overlaps the tail of the previous CPI's estimator.

        // ---- stage 1: channel estimates H[2 pair + {0,1}][k = 32 half + lane] ------
        if (FROM_H) {
            const c32 *Hg = P.H + (long long)cpi * 512 + (2 * pair) * 64 + 32 * half + lane;
            const c32 h0 = Hg[0], h1 = Hg[64];
            Hw[32 * half + lane] = make_float4(h0.x, h0.y, h1.x, h1.y);
        } else {
            c32 acc0 = mk(0.f, 0.f), acc1 = mk(0.f, 0.f);
            for (int s = 0; s < P.S; s++) {
                c32 x0 = srx[0][s * 64], c0 = stx[0][s * 64], x1 = srx[1][s * 64], c1 = stx[1][s * 64];
                acc0 = cadd_exact(acc0, cmul_exact(x0, mk(c0.x, -c0.y)));
                acc1 = cadd_exact(acc1, cmul_exact(x1, mk(c1.x, -c1.y)));
            }
            Hw[32 * half + lane] = make_float4(acc0.x, acc0.y, acc1.x, acc1.y);
        }
        __syncthreads();   // (A) H complete; symbol buffer free; previous CPI's y fully consumed
        if (pending) finalize();
        if (!FROM_H) {
            int nxt = cpi + gridDim.x;
            if (nxt < P.n_cpi) prefetch(nxt);
        }

        // ---- stage 2: range pass 1 (pruned: 8 of Q inputs non-zero) ----------
#pragma unroll
        for (int j = 0; j < T1; j++) {
            const int k0 = (32 * half + lane + 64 * j) / IR;
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
                yw[k0 * Q + q0 + IR * m0] = make_float4(u0[m0].x, u0[m0].y, u1[m0].x, u1[m0].y);
        }
        pair_sync();

        // ---- stage 3: range pass 2, in place ---------------------------------
#pragma unroll
        for (int j = 0; j < T2; j++) {
            const int q = 32 * half + lane + 64 * j;
            c32 u0[8], u1[8];
#pragma unroll
            for (int k0 = 0; k0 < 8; k0++) {
                const float4 t = yw[k0 * Q + q];
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
            // in place per thread, regrouped for the angle pass: rows (2j'Q + q, (2j'+1)Q + q) of one channel
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                yw[jj * Q + q] = make_float4(u0[2 * jj].x, u0[2 * jj + 1].x, u0[2 * jj].y, u0[2 * jj + 1].y);
                yw[(4 + jj) * Q + q] = make_float4(u1[2 * jj].x, u1[2 * jj + 1].x, u1[2 * jj].y, u1[2 * jj + 1].y);
            }
        }
        __syncthreads();   // (D) y[p][n] complete for all channels

        // ---- stage 4: angle pass + |.|^2 + store + running max ---------------
        float best0 = -1.f, best1 = -1.f;   // rows of tile 0 all precede those of tile 1
        float sec = -1.f;                   // largest group maximum of this thread that is not best0 / best1
        int bit0 = 0, bit1 = 0;
        const int jp = warp >> 1;
        const int qw = (warp & 1) * (Q / 2);             // first q of this warp
        const int row0 = 2 * jp * Q + qw;                // first map row of tile 0
        float4 *map_w = WRITE_MAP ? reinterpret_cast<float4 *>(P.map + ((long long)cpi * NR + row0) * NA) : nullptr;
        for (int itb = 0; itb < PITERS; itb += Gm::UNROLL) {
#pragma unroll
            for (int ui = 0; ui < Gm::UNROLL; ui++) {
                const int it = itb + ui;
                float2 re[8], im[8], v[8];
                angle_pair(ys, NR, Q, jp * Q + qw + g + it * G, tw3, re, im, v);
                const float m0 = fmaxf(fmaxf(fmaxf(v[0].x, v[1].x), fmaxf(v[2].x, v[3].x)),
                                       fmaxf(fmaxf(v[4].x, v[5].x), fmaxf(v[6].x, v[7].x)));
                const float m1 = fmaxf(fmaxf(fmaxf(v[0].y, v[1].y), fmaxf(v[2].y, v[3].y)),
                                       fmaxf(fmaxf(v[4].y, v[5].y), fmaxf(v[6].y, v[7].y)));
                sec = fmaxf(sec, fmaxf(fminf(m0, best0), fminf(m1, best1)));
                if (m0 > best0) { best0 = m0; bit0 = it; }
                if (m1 > best1) { best1 = m1; bit1 = it; }
                if (WRITE_MAP) {
                    // conflict-free scalar st.shared of the strided bins, then the warp streams its two
                    // 1 KiB tiles (G whole map rows each, contiguous in HBM) out as 4 x 512 B
#pragma unroll
                    for (int a = 0; a < 8; a++) { sp[a][0] = v[a].x; sp[a][TILE] = v[a].y; }
                    __syncwarp();
                    const float4 *src = reinterpret_cast<const float4 *>(wstg);
                    const float4 o0 = src[lane], o1 = src[lane + 32], o2 = src[lane + 64], o3 = src[lane + 96];
                    __syncwarp();
                    float4 *dst = map_w + it * (TILE / 4);
                    __stcs(dst + lane, o0);
                    __stcs(dst + lane + 32, o1);
                    __stcs(dst + Q * NA / 4 + lane, o2);
                    __stcs(dst + Q * NA / 4 + lane + 32, o3);
                }
            }
        }
        // this thread's candidate: lowest row among its maxima
        const float best = best0 >= best1 ? best0 : best1;
        const int best_row = best0 >= best1 ? row0 + g + bit0 * G : row0 + Q + g + bit1 * G;

        // ---- stage 5: range_angle_estimator ----------------------------------
        pending = false;
        const bool in_kernel_est = P.dets != nullptr;
        if (in_kernel_est) {
            unsigned long long key = best >= 0.f ? pack_key(best, (unsigned)best_row) : 0ull;
            float b2 = fmaxf(sec, fminf(best0, best1));      // runner-up among this thread's group maxima
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                const float o2 = __shfl_xor_sync(0xffffffffu, b2, o);
                const float v1 = __uint_as_float((unsigned)(key >> 32)), v1o = __uint_as_float((unsigned)(other >> 32));
                // keys are distinct unless both are 0: the loser of (key, other) is a runner-up candidate
                b2 = fmaxf(fmaxf(b2, o2), (key && other) ? fminf(v1, v1o) : -1.f);
                key = other > key ? other : key;
            }
            if (lane == 0) { red[warp] = key; sint[8 + warp] = __float_as_int(b2); }
        }
        if (!FROM_H) cp_async_wait_all();
        __syncthreads();   // (E) block maximum; the next CPI's symbols are visible to every thread
        if (in_kernel_est) {
            unsigned long long key = red[0];
            float g2 = __int_as_float(sint[8]);
#pragma unroll
            for (int w = 1; w < 8; w++) {
                const unsigned long long kw = red[w];
                g2 = fmaxf(fmaxf(g2, __int_as_float(sint[8 + w])),
                           (key && kw) ? fminf(__uint_as_float((unsigned)(key >> 32)), __uint_as_float((unsigned)(kw >> 32))) : -1.f);
                key = kw > key ? kw : key;
            }
            if (key == 0ull) {   // NaN-only input: nothing can win the strict '>' scan
                if (tid == 0) {
                    DetDev d; d.range_idx = -1; d.angle_idx = -1; d.peak_power = -1.f;
                    d.noise_power = __int_as_float(0x7fc00000); d.snr_db = d.noise_power;
                    d.n_noise = 0; d.flags = 0; d.cpi = P.cpi0 + cpi;
                    P.dets[cpi] = d;
                }
            } else {
                const float gmax = __uint_as_float((unsigned)(key >> 32));
                const int nstar_fast = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
                const float thr_amb = __fmul_rn(gmax, 1.f - EPS_AMB);
                const bool amb = g2 >= thr_amb;          // block-uniform
                bool resolved = false;
                if (amb) {
                    // Rare (~1e-4 of the CPIs): another map element is within the FFT rounding of the maximum, so the
                    // fast map cannot say which one the reference's scan (:137-151) would keep.  Collect the candidates
                    // (one more angle pass, no stores), evaluate each in the STAGED arithmetic (dit_bin_*: the same float
                    // operations as the radix-2 transforms of the staged path) and let the reference's rule decide.
                    constexpr int MAXC = 24;
                    int *cand = reinterpret_cast<int *>(redA);                       // [0] count, [1 ..] linear map indices
                    c32 *ycand = reinterpret_cast<c32 *>(redA + 16);                 // [8] range spectra of a candidate's row
                    if (tid == 0) cand[0] = 0;
                    __syncthreads();
                    for (int it = 0; it < PITERS; it++) {
                        float2 re[8], im[8], v[8];
                        angle_pair(ys, NR, Q, jp * Q + qw + g + it * G, tw3, re, im, v);
#pragma unroll
                        for (int a = 0; a < 8; a++) {
                            const int i = b + IA * ((a + rot) & 7), r = row0 + g + it * G;
                            if (v[a].x >= thr_amb) { const int sl = atomicAdd(cand, 1); if (sl < MAXC) cand[1 + sl] = r * NA + i; }
                            if (v[a].y >= thr_amb) { const int sl = atomicAdd(cand, 1); if (sl < MAXC) cand[1 + sl] = (r + Q) * NA + i; }
                        }
                    }
                    __syncthreads();
                    const int nc = cand[0];
                    if (nc >= 1 && nc <= MAXC) {
                        unsigned long long kb = 0ull;
                        for (int c = 0; c < nc; c++) {
                            const int lin = cand[1 + c], n = lin / NA, i = lin % NA;
                            {   // warp w: bin n of the range IFFT of channel w
                                const int k0 = (int)(__brev((unsigned)(2 * lane)) >> 26), k1 = (int)(__brev((unsigned)(2 * lane + 1)) >> 26);
                                const float4 h0 = Hs[(warp >> 1) * 64 + k0], h1 = Hs[(warp >> 1) * 64 + k1];
                                const c32 a0 = (warp & 1) ? mk(h0.z, h0.w) : mk(h0.x, h0.y);
                                const c32 a1 = (warp & 1) ? mk(h1.z, h1.w) : mk(h1.x, h1.y);
                                const c32 yv = dit_bin_warp64(a0, a1, 6 + (31 - __clz(IR)), P.tw_r, n);
                                if (lane == 0) ycand[warp] = yv;
                            }
                            __syncthreads();
                            if (tid == 0) {     // bin i of the shifted angle FFT across the 8 channels
                                c32 yy[8];
#pragma unroll
                                for (int p = 0; p < 8; p++) yy[p] = ycand[p];
                                const c32 z = dit_bin8(yy, 3 + (31 - __clz(IA)), P.tw_a, (i + NA / 2) & (NA - 1));
                                const float pw = (float)ref_pow_abs2(z);
                                if (pw == pw) { const unsigned long long k2 = pack_key(pw, (unsigned)lin); kb = k2 > kb ? k2 : kb; }
                            }
                            __syncthreads();
                        }
                        if (tid == 0) {
                            const unsigned lin = 0xFFFFFFFFu - (unsigned)(kb & 0xFFFFFFFFull);
                            sint[0] = (int)(lin / (unsigned)NA); sint[1] = (int)(lin % (unsigned)NA);
                            sint[5] = kb == 0ull;                                       // (NaN candidates only: leave it to k_est_exact)
                            sint[6] = (int)(unsigned)(kb >> 32);
                            sint[7] = cpi;
                            atomicAdd(&P.fix_ctl->n_inkernel, 1);
                        }
                        __syncthreads();
                        resolved = true;
                    }
                }
                amb_prev = amb && !resolved;
                const int nstar = resolved ? sint[0] : nstar_fast;
                // Noise window (lib/range_angle_estimator_impl.cc:197-226) without evaluating its samples:
                //   sum_{r,c} |sum_p y[p][r] w^{p c'}|^2 = ncols*A[0] + 2 Re sum_{d=1..7} g[d] A[d],
                //   A[d] = sum_r sum_q y[q+d][r] conj(y[q][r]),  g[d] = sum_c w^{d c'},  c' = c + Na/2,
                // (w = e^{-j2pi/Na}; the modulo wrap of rows is the index, that of columns the period of w).
                // The window ROWS only depend on the peak's range bin, so every warp starts on them at once:
                // one row per lane (conflict-free loads), the 36 lag products split over the four warp pairs,
                // float within a lane, double from the first reduction on.  g[d] comes from k_est_tables.
                {
                    const int start_r = nstar + NR / 2 - est.discard_range_idx;
                    const int end_r = nstar + NR / 2 + est.discard_range_idx;
                    const int part = warp >> 1;
                    float acc[6];
#pragma unroll
                    for (int k = 0; k < 6; k++) acc[k] = 0.f;
                    for (int ir = start_r + 32 * (warp & 1) + lane; ir < end_r; ir += 64) {
                        const int r_idx = ((ir % NR) + NR) % NR;
                        const int rq = r_idx % Q, rm = r_idx / Q;
                        c32 yv[8];
#pragma unroll
                        for (int p = 0; p < 8; p++) {
                            const float4 t = ys[(p >> 1) * NR + (4 * (p & 1) + (rm >> 1)) * Q + rq];
                            yv[p] = (rm & 1) ? mk(t.y, t.w) : mk(t.x, t.z);
                        }
                        auto lagsum = [&](int d, float &sr, float &si) {
#pragma unroll
                            for (int qq = 0; qq < 8; qq++) {
                                if (qq + d < 8) {
                                    const c32 ya = yv[qq + d], yb = yv[qq];
                                    sr = __fmaf_rn(ya.x, yb.x, __fmaf_rn(ya.y, yb.y, sr));
                                    si = __fmaf_rn(ya.y, yb.x, __fmaf_rn(-ya.x, yb.y, si));
                                }
                            }
                        };
                        if (part == 0) { lagsum(0, acc[0], acc[1]); }
                        else if (part == 1) { lagsum(1, acc[0], acc[1]); lagsum(7, acc[2], acc[3]); }
                        else if (part == 2) { lagsum(2, acc[0], acc[1]); lagsum(6, acc[2], acc[3]); }
                        else { lagsum(3, acc[0], acc[1]); lagsum(4, acc[2], acc[3]); lagsum(5, acc[4], acc[5]); }
                    }
                    // transposing reduction (13 shuffles instead of 60): every step halves the values a
                    // lane carries; lanes 0,4,8 / 16,20,24 end up with the warp totals of acc[0..2] / acc[3..5]
                    double t3[3], t2[2], t1;
                    {
                        const bool hi = lane & 16;
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            const float keep = hi ? acc[3 + k] : acc[k], give = hi ? acc[k] : acc[3 + k];
                            t3[k] = (double)keep + (double)__shfl_xor_sync(0xffffffffu, give, 16);
                        }
                    }
                    {
                        const bool hi = lane & 8;
                        const double k0 = hi ? t3[2] : t3[0], g0 = hi ? t3[0] : t3[2];
                        const double k1 = hi ? 0.0 : t3[1], g1 = hi ? t3[1] : 0.0;
                        t2[0] = k0 + __shfl_xor_sync(0xffffffffu, g0, 8);
                        t2[1] = k1 + __shfl_xor_sync(0xffffffffu, g1, 8);
                    }
                    {
                        const bool hi = lane & 4;
                        const double keep = hi ? t2[1] : t2[0], give = hi ? t2[0] : t2[1];
                        t1 = keep + __shfl_xor_sync(0xffffffffu, give, 4);
                    }
                    t1 += __shfl_xor_sync(0xffffffffu, t1, 2);
                    t1 += __shfl_xor_sync(0xffffffffu, t1, 1);
                    if ((lane & 3) == 0 && (lane & 12) != 12) {
                        // value index: (lane & 16 ? 3 : 0) + (lane & 8 ? 2 : 0) + (lane & 4 ? 1 : 0)
                        redA[warp * 8 + ((lane >> 4) * 3 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1))] = t1;
                    }
                }
                const int sq = nstar % Q, sm1 = nstar / Q;
                if (!resolved && warp == (sm1 >> 1) * 2 + (sq >= Q / 2)) {
                    // the lanes that own row nstar re-evaluate it and pick the first bin == gmax
                    int icand = 0x7fffffff, ncand = 0;
                    c32 zc = mk(0.f, 0.f);
                    if (g == sq % G) {
                        float2 re[8], im[8], v[8];
                        angle_pair(ys, NR, Q, (sm1 >> 1) * Q + sq, tw3, re, im, v);
                        const bool odd = sm1 & 1;
#pragma unroll
                        for (int a = 0; a < 8; a++) {
                            const int i = b + IA * ((a + rot) & 7);
                            const float va = odd ? v[a].y : v[a].x;
                            ncand += va >= thr_amb;
                            if (va == gmax && i < icand) { icand = i; zc = odd ? mk(re[a].y, im[a].y) : mk(re[a].x, im[a].x); }
                        }
                    }
                    int imin = icand;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        imin = min(imin, __shfl_xor_sync(0xffffffffu, imin, o));
                        ncand += __shfl_xor_sync(0xffffffffu, ncand, o);
                    }
                    if (icand == imin) {   // exactly one lane (bins are distinct); imin is always found
                        sint[0] = nstar; sint[1] = imin & (NA - 1);
                        sint[5] = ncand > 1;       // several bins of the winning row within EPS_AMB
                        sint[6] = __float_as_int((float)ref_pow_abs2(zc));
                        sint[7] = cpi;
                    }
                }
                pending = true;   // finished after the next barrier (A) / after the loop
            }
        }
    }
    if (pending) {
        __syncthreads();
        finalize();
    }
}

}  // namespace jrc
