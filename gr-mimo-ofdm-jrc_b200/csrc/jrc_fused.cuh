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
// (cp.async prefetch of the next CPI), the |.|^2 map once (per-warp TMA bulk stores
// from a bank-conflict-free staging tile) and a 32-byte detection record; neither
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

namespace jrc {

struct FusedParams {
    PortDev rx, tx;            // symbol inputs (unused when FROM_H)
    const c32 *H;              // [n_cpi][8][64] channel estimates (FROM_H: background removal path)
    int n_cpi, cpi0;
    int T, R, S, n_pre, tx_interleave;
    float *map;                // [n_cpi][NR][NA] or nullptr
    DetDev *dets;              // [n_cpi] or nullptr (in-kernel estimator)
    unsigned long long *keys;  // [n_cpi] zeroed, or nullptr: per-CPI arg-max key for k_map_finalize instead of dets
    EstParams est;
};

template <int IR, int IA>
struct FusedGeom {
    static constexpr int NSC = 64, V = 8;
    static constexpr int NR = NSC * IR, NA = V * IA, Q = NR / 8;
    static constexpr int THREADS = 256, WARPS = 8;
    static constexpr int G = 32 / IA;                    // range bins per warp iteration
    static constexpr int ROWS_PER_WARP = NR / WARPS;
    static constexpr int ITERS = ROWS_PER_WARP / G;
#ifndef JRC_STORE_MODE
#define JRC_STORE_MODE 0
#endif
    // Map store path (A/B measured on B200, profiles/README.md):
    //   0: each warp stages its 1 KiB tile (G whole map rows) in shared memory and streams it out
    //      with 2 x (LDS.128 + STG.128) per iteration;
    //   1: per-warp cp.async.bulk (UBLKCP) stores of SIT KiB from a ring of NBUF staging tiles;
    //   2: no staging: 8 STG.32 per thread, every warp store writing G x IA/8 full 32-byte sectors.
    static constexpr int STORE = JRC_STORE_MODE;
    static constexpr bool TMA = STORE == 1;
    static constexpr int SIT = TMA ? 2 : 1;              // iterations per store
    static constexpr int NBUF = TMA ? 2 : 1;             // staging buffers per warp
    static constexpr int UNROLL = 4;
    static constexpr int STG_FLOATS = SIT * G * NA;      // floats per staging buffer (SIT KiB)
    static_assert(IA >= 4 && IA <= 32 && (IA & (IA - 1)) == 0, "angle interp must be 4..32, power of two");
    static_assert(IR >= 1 && IR <= 32 && (IR & (IR - 1)) == 0, "range interp must be 1..32, power of two");
    static_assert(ITERS >= UNROLL && ITERS % UNROLL == 0 && UNROLL % (SIT * NBUF) == 0, "map too small for the store pipeline");

    static size_t smem_bytes(int T, int R, int S, bool from_h)
    {
        size_t b = (size_t)V * NR * 8 + (size_t)V * NSC * 8 + (size_t)8 * Q * 8 + (size_t)WARPS * NBUF * STG_FLOATS * 4 +
                   (size_t)NA * 16 + (size_t)NA * 4 + 64 + 1024 + 64;
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
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_s2g(void *gdst, const void *ssrc, unsigned bytes)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// lane-0-only forms as predicated instructions (no BSSY/BSYNC divergence scaffolding in the hot loop)
template <int N>
__device__ __forceinline__ void bulk_wait_read_lane0(int lane)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %0, 0;\n\t@p cp.async.bulk.wait_group.read %1;\n\t}" ::"r"(lane), "n"(N)
                 : "memory");
}
__device__ __forceinline__ void bulk_store_commit_lane0(int lane, void *gdst, const void *ssrc, unsigned bytes)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %0, 0;\n\t"
                 "@p cp.async.bulk.global.shared::cta.bulk_group [%1], [%2], %3;\n\t"
                 "@p cp.async.bulk.commit_group;\n\t}" ::"r"(lane), "l"(gdst), "r"(s), "r"(bytes)
                 : "memory");
}

// one angle row task: 8 channel samples -> twiddle -> forward DFT-8 -> |.|^2.
// Pure _rn intrinsics: re-evaluating a row reproduces the main loop bit for bit.
__device__ __forceinline__ void angle_task(const c32 *__restrict__ ys, int NR, int n, const c32 (&tw)[8],
                                           c32 (&u)[8], float (&v)[8])
{
#pragma unroll
    for (int p = 0; p < 8; p++) u[p] = ys[p * NR + n];
#pragma unroll
    for (int p = 1; p < 8; p++) u[p] = cmul_fma(u[p], tw[p]);
    JRC_FFT8<-1>(u);
#pragma unroll
    for (int a = 0; a < 8; a++) {   // volk_32fc_magnitude_squared_32f: re*re + im*im, each product rounded
        c32 sq = __fmul2_rn(u[a], u[a]);
        v[a] = __fadd_rn(sq.x, sq.y);
    }
}

template <int IR, int IA, bool FROM_H, bool WRITE_MAP>
__global__ void __launch_bounds__(256, 2) k_fused64x8(const FusedParams P)
{
    using Gm = FusedGeom<IR, IA>;
    constexpr int NR = Gm::NR, NA = Gm::NA, Q = Gm::Q, G = Gm::G;
    constexpr int RPW = Gm::ROWS_PER_WARP, ITERS = Gm::ITERS, SIT = Gm::SIT, NBUF = Gm::NBUF;
    constexpr int STGF = Gm::STG_FLOATS;
    constexpr int T1 = (8 * IR + 31) / 32;     // range pass 1 tasks per lane (channel = warp)
    constexpr int T2 = (Q + 31) / 32;          // range pass 2 tasks per lane

    extern __shared__ __align__(128) unsigned char smem_raw[];
    c32 *ys = reinterpret_cast<c32 *>(smem_raw);                  // [8][NR]  B then y (in place)
    c32 *Hs = ys + 8 * NR;                                        // [8][64]
    c32 *tw2t = Hs + 512;                                         // [8][Q]   W_Nr^{k0 q}
    float *stg = reinterpret_cast<float *>(tw2t + 8 * Q);         // [8 warps][NBUF][STGF]
    double2 *tabd = reinterpret_cast<double2 *>(stg + 8 * NBUF * STGF);   // [NA] e^{-j2pi m/NA}
    float *abin = reinterpret_cast<float *>(tabd + NA);                   // [NA] angle_bins copy
    unsigned long long *red = reinterpret_cast<unsigned long long *>(abin + NA);   // [8]
    double2 *redA = reinterpret_cast<double2 *>(red + 8);         // [8 lags] (room for 64)
    int *sint = reinterpret_cast<int *>(redA + 64);               // [16] scalars
    c32 *inb = reinterpret_cast<c32 *>(sint + 16);                // [(T+R)][S][64]

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler

    // ---- per-thread constants (once per persistent CTA) --------------------
    // Stages 1-3 give every warp ONE virtual channel (p = warp), so the channel estimate and both
    // range passes only need warp-level synchronisation.
    // range pass 1: task (k0, q0), q0 = lane % IR fixed per lane;  twiddle W_Q^{k1 q0}
    // range pass 2: task q = lane + 32 j;                           twiddle W_Nr^{k0 q} from tw2t
    // angle pass:   task (n, b);                                    twiddle (-1)^p w_Na^{p (b + IA*rot)}
    const int q0 = lane % IR;
    const int b = lane % IA, g = lane / IA, rot = (Gm::STORE == 2) ? 0 : g;
    c32 tw1[8], tw3[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        tw1[j] = cispi_ratio(2 * j * q0, Q);
        tw3[j] = cispi_ratio(j * (NA - 2 * (b + IA * rot)), NA);
    }
    for (int e = tid; e < 8 * Q; e += 256) tw2t[e] = cispi_ratio(2 * (e / Q) * (e % Q), NR);
    EstParams est = P.est;
    for (int m = tid; m < NA; m += 256) {
        double sn, cs;
        sincospi(-2.0 * (double)m / (double)NA, &sn, &cs);
        tabd[m] = make_double2(cs, sn);
        if (P.dets) abin[m] = P.est.angle_bins[m];
    }
    est.angle_bins = abin;   // the window geometry's binary search runs on shared memory

    // channel of this warp -> (rx antenna, tx antenna)  (lib/mimo_ofdm_radar_impl.cc:262-269)
    const int per_ant = P.S * 64;
    int ch_r, ch_t;
    if (P.tx_interleave) { ch_t = warp / P.R; ch_r = warp - ch_t * P.R; } else { ch_r = warp / P.T; ch_t = warp - ch_r * P.T; }
    const c32 *srx = inb + (P.T + ch_r) * per_ant + lane;
    const c32 *stx = inb + ch_t * per_ant + lane;
    c32 *Hw = Hs + warp * 64;          // this warp's channel estimate
    c32 *yw = ys + warp * NR;          // this warp's range spectrum

    // staging: slot a' of this thread holds angle bin b + IA*((a'+rot)&7); rotating by
    // the row group makes the 32 lanes of every st.shared hit 32 different banks
    float *wstg = stg + warp * NBUF * STGF;
    float *sp[8];
#pragma unroll
    for (int a = 0; a < 8; a++) sp[a] = wstg + g * NA + b + IA * ((a + rot) & 7);

    auto prefetch = [&](int cpi) {
        const int cpa = per_ant >> 1;   // 16-byte chunks per antenna row
        const int total = (P.T + P.R) * cpa;
        for (int c = tid; c < total; c += 256) {
            int a = c / cpa, w = c - a * cpa;
            const c32 *src = (a < P.T)
                ? P.tx.base + (long long)cpi * P.tx.cpi_stride + (long long)a * P.tx.ant_stride
                : P.rx.base + (long long)cpi * P.rx.cpi_stride + (long long)(a - P.T) * P.rx.ant_stride;
            cp_async16(inb + a * per_ant + 2 * w, src + (long long)P.n_pre * 64 + 2 * w);
        }
        cp_async_commit();
    };

    // Detection record of the previous CPI: its last step only needs the shared scratch (not y), so it
    // runs after the next CPI's first barrier instead of costing a barrier of its own.
    bool pending = false;
    auto finalize = [&]() {
        if (tid < 8) {
            const int start_a = sint[4], end_a = sint[5], total = sint[8];
            const int ncols = end_a - start_a;
            const double sr = redA[tid].x, si = redA[tid].y;
            double contrib;
            if (tid == 0) {
                contrib = (double)ncols * sr;
            } else {
                // g[d] = w^{d m0} (1 - w^{d ncols}) / (1 - w^d),  m0 = start_a + Na/2
                const double2 w0 = tabd[(tid * (start_a + NA / 2)) & (NA - 1)];
                const double2 w1 = tabd[(tid * ncols) & (NA - 1)];
                const double2 w2 = tabd[tid];
                const double c0 = w0.x, s0 = w0.y;
                const double nr = 1.0 - w1.x, ni = -w1.y, dr = 1.0 - w2.x, di = -w2.y;
                const double den = dr * dr + di * di;
                const double qr = (nr * dr + ni * di) / den, qi = (ni * dr - nr * di) / den;
                const double gr = c0 * qr - s0 * qi, gi = c0 * qi + s0 * qr;
                contrib = 2.0 * (gr * sr - gi * si);
            }
            contrib += __shfl_xor_sync(0xffu, contrib, 4);
            contrib += __shfl_xor_sync(0xffu, contrib, 2);
            contrib += __shfl_xor_sync(0xffu, contrib, 1);
            if (tid == 0) {
                const double s = contrib > 0.0 ? contrib : 0.0;
                DetDev d;
                d.range_idx = sint[0]; d.angle_idx = sint[1];
                d.peak_power = __int_as_float(sint[6]);
                d.n_noise = total;
                d.noise_power = __fdiv_rn((float)s, (float)total);
                d.snr_db = __fmul_rn(10.f, log10f(__fdiv_rn(d.peak_power, d.noise_power)));
                d.flags = (d.snr_db >= est.snr_threshold && d.peak_power >= est.power_threshold) ? 1u : 0u;
                d.cpi = P.cpi0 + sint[7];
                P.dets[sint[7]] = d;
            }
        }
    };

    int cpi = blockIdx.x;
    if (!FROM_H && cpi < P.n_cpi) prefetch(cpi);

    for (; cpi < P.n_cpi; cpi += gridDim.x) {
        if (!FROM_H) cp_async_wait_all();
        __syncthreads();   // (A) symbols landed; previous CPI fully consumed
        if (pending) finalize();

        // ---- stage 1: channel estimate H[p = warp][k] --------------------------
        if (FROM_H) {
            const c32 *Hg = P.H + (long long)cpi * 512 + warp * 64;
            Hw[lane] = Hg[lane];
            Hw[lane + 32] = Hg[lane + 32];
        } else {
            c32 acc0 = mk(0.f, 0.f), acc1 = mk(0.f, 0.f);
            for (int s = 0; s < P.S; s++) {
                c32 x0 = srx[s * 64], c0 = stx[s * 64], x1 = srx[s * 64 + 32], c1 = stx[s * 64 + 32];
                acc0 = cadd_exact(acc0, cmul_exact(x0, mk(c0.x, -c0.y)));
                acc1 = cadd_exact(acc1, cmul_exact(x1, mk(c1.x, -c1.y)));
            }
            Hw[lane] = acc0;
            Hw[lane + 32] = acc1;
        }
        __syncwarp();

        // ---- stage 2: range pass 1 (pruned: 8 of Q inputs non-zero) ----------
#pragma unroll
        for (int j = 0; j < T1; j++) {
            const int task = lane + 32 * j;
            if (8 * IR >= 32 || task < 8 * IR) {
                const int k0 = task / IR;
                c32 u[8];
#pragma unroll
                for (int k1 = 0; k1 < 8; k1++) u[k1] = Hw[k0 + 8 * k1];
#pragma unroll
                for (int k1 = 1; k1 < 8; k1++) u[k1] = cmul_fma(u[k1], tw1[k1]);
                JRC_FFT8<1>(u);
#pragma unroll
                for (int m0 = 0; m0 < 8; m0++) yw[k0 * Q + q0 + IR * m0] = u[m0];
            }
        }
        __syncwarp();

        // ---- stage 3: range pass 2, in place ---------------------------------
#pragma unroll
        for (int j = 0; j < T2; j++) {
            const int q = lane + 32 * j;
            if (Q >= 32 || q < Q) {
                c32 u[8];
#pragma unroll
                for (int k0 = 0; k0 < 8; k0++) u[k0] = yw[k0 * Q + q];
#pragma unroll
                for (int k0 = 1; k0 < 8; k0++) u[k0] = cmul_fma(u[k0], tw2t[k0 * Q + q]);
                JRC_FFT8<1>(u);
#pragma unroll
                for (int m1 = 0; m1 < 8; m1++) yw[m1 * Q + q] = u[m1];
            }
        }
        __syncthreads();   // (D) y[p][n] complete for all channels; symbol buffer free
        if (!FROM_H) {
            int nxt = cpi + gridDim.x;
            if (nxt < P.n_cpi) prefetch(nxt);
        }

        // ---- stage 4: angle pass + |.|^2 + store + running max ---------------
        float best = -1.f;
        int best_it = 0;
        const int n_base = warp * RPW + g;
        float *map_w = WRITE_MAP ? P.map + ((long long)cpi * NR + warp * RPW) * NA : nullptr;
        for (int it0 = 0; it0 < ITERS; it0 += Gm::UNROLL) {
#pragma unroll
            for (int ui = 0; ui < Gm::UNROLL; ui++) {
                const int it = it0 + ui;
                const int buf = (ui / SIT) % NBUF, sub = ui % SIT;    // compile-time after unrolling
                c32 u[8];
                float v[8];
                angle_task(ys, NR, n_base + it * G, tw3, u, v);
                float m8 = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])),
                                 fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
                if (m8 > best) { best = m8; best_it = it; }
                if (WRITE_MAP && Gm::TMA) {
                    if (sub == 0) {   // the bulk store that last read this buffer must be done with it
                        bulk_wait_read_lane0<NBUF - 1>(lane);
                        __syncwarp();
                    }
#pragma unroll
                    for (int a = 0; a < 8; a++) sp[a][buf * STGF + sub * G * NA] = v[a];
                    if (sub == SIT - 1) {
                        fence_proxy_async_smem();
                        __syncwarp();
                        bulk_store_commit_lane0(lane, map_w + (long long)(it - (SIT - 1)) * G * NA, wstg + buf * STGF,
                                                STGF * 4);
                    }
                } else if (WRITE_MAP && Gm::STORE == 2) {
                    float *dst = map_w + (long long)(it * G + g) * NA + b;
#pragma unroll
                    for (int a = 0; a < 8; a++) __stcs(dst + IA * a, v[a]);
                } else if (WRITE_MAP) {
                    // conflict-free scalar st.shared of the strided bins, then the warp streams its
                    // 1 KiB tile (G whole map rows, contiguous in HBM) out as 2 x 512 B
#pragma unroll
                    for (int a = 0; a < 8; a++) sp[a][0] = v[a];
                    __syncwarp();
                    const float4 o0 = reinterpret_cast<const float4 *>(wstg)[lane];
                    const float4 o1 = reinterpret_cast<const float4 *>(wstg)[lane + 32];
                    __syncwarp();
                    float4 *dst = reinterpret_cast<float4 *>(map_w + (long long)it * G * NA);
                    __stcs(dst + lane, o0);
                    __stcs(dst + lane + 32, o1);
                }
            }
        }

        // ---- stage 5: range_angle_estimator ----------------------------------
        pending = false;
        if (P.keys) {
            // map-backed detection: fold this CTA's maximum into the CPI's 64-bit key (no CTA barrier);
            // k_map_finalize turns key + map into the record after the kernel
            unsigned long long key = best >= 0.f ? pack_key(best, (unsigned)(n_base + best_it * G)) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                key = other > key ? other : key;
            }
            if (lane == 0 && key) atomicMax(P.keys + cpi, key);
        } else if (P.dets) {
            unsigned long long key = best >= 0.f ? pack_key(best, (unsigned)(n_base + best_it * G)) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                key = other > key ? other : key;
            }
            if (lane == 0) red[warp] = key;
            __syncthreads();   // (E)
            key = red[0];
#pragma unroll
            for (int w = 1; w < 8; w++) key = red[w] > key ? red[w] : key;
            if (key == 0ull) {   // NaN-only input: nothing can win the strict '>' scan
                if (tid == 0) {
                    DetDev d; d.range_idx = -1; d.angle_idx = -1; d.peak_power = -1.f;
                    d.noise_power = __int_as_float(0x7fc00000); d.snr_db = d.noise_power;
                    d.n_noise = 0; d.flags = 0; d.cpi = P.cpi0 + cpi;
                    P.dets[cpi] = d;
                }
            } else {
                const float gmax = __uint_as_float((unsigned)(key >> 32));
                const int nstar = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
                if (warp == nstar / RPW) {
                    // Everything that is left runs in the warp that owns row nstar; the other warps go on
                    // to the next CPI's barrier (A).  The lanes that own the row re-evaluate it and pick the
                    // first bin == gmax.
                    int icand = 0x7fffffff;
                    c32 zc = mk(0.f, 0.f);
                    if (g == (nstar % RPW) % G) {
                        c32 u[8]; float v[8];
                        angle_task(ys, NR, nstar, tw3, u, v);
#pragma unroll
                        for (int a = 0; a < 8; a++) {
                            int i = b + IA * ((a + rot) & 7);
                            if (v[a] == gmax && i < icand) { icand = i; zc = u[a]; }
                        }
                    }
                    int imin = icand;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) imin = min(imin, __shfl_xor_sync(0xffffffffu, imin, o));
                    const int src = __ffs(__ballot_sync(0xffffffffu, icand == imin)) - 1;
                    int start_r = 0, end_r = 0, total = 0;
                    if (lane == src) {
                        NoiseWin w = noise_window(est, nstar, imin);
                        const int ncols = w.end_a - w.start_a, nrows = w.end_r - w.start_r;
                        start_r = w.start_r; end_r = w.end_r;
                        total = (ncols > 0 && nrows > 0) ? nrows * ncols : 0;
                        sint[0] = nstar; sint[1] = imin;
                        sint[4] = w.start_a; sint[5] = w.end_a;
                        sint[6] = __float_as_int((float)ref_pow_abs2(zc));
                        sint[7] = cpi;
                        sint[8] = total;
                    }
                    start_r = __shfl_sync(0xffffffffu, start_r, src);
                    end_r = __shfl_sync(0xffffffffu, end_r, src);
                    total = __shfl_sync(0xffffffffu, total, src);
                    // Noise window (lib/range_angle_estimator_impl.cc:197-226) without evaluating its samples:
                    //   sum_{r,c} |sum_p y[p][r] w^{p c'}|^2 = ncols*A[0] + 2 Re sum_{d=1..7} g[d] A[d],
                    //   A[d] = sum_r sum_q y[q+d][r] conj(y[q][r]),  g[d] = sum_c w^{d c'},  c' = c + Na/2,
                    // (w = e^{-j2pi/Na}; the modulo wrap of rows is the index, that of columns the period of w).
                    // One window row per lane (consecutive rows: conflict-free loads), 36 complex MACs per row
                    // instead of 8 per sample; float within a lane, double from the first reduction on.
                    float acc[16];   // [0..7] Re A[d], [8..15] Im A[d]
#pragma unroll
                    for (int d = 0; d < 16; d++) acc[d] = 0.f;
                    if (total > 0) {
                        for (int ir = start_r + lane; ir < end_r; ir += 32) {
                            const int r_idx = ((ir % NR) + NR) % NR;
                            c32 yv[8];
#pragma unroll
                            for (int p = 0; p < 8; p++) yv[p] = ys[p * NR + r_idx];
#pragma unroll
                            for (int d = 0; d < 8; d++) {
#pragma unroll
                                for (int qq = 0; qq + d < 8; qq++) {
                                    const c32 ya = yv[qq + d], yb = yv[qq];
                                    acc[d] = __fmaf_rn(ya.x, yb.x, __fmaf_rn(ya.y, yb.y, acc[d]));
                                    acc[8 + d] = __fmaf_rn(ya.y, yb.x, __fmaf_rn(-ya.x, yb.y, acc[8 + d]));
                                }
                            }
                        }
                    }
                    // transposing reduction: every step halves the values a lane carries; lane l ends up
                    // with the warp total of value l >> 1
                    double s8[8], s4[4], s2[2], s1;
                    {
                        const bool hi = lane & 16;
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const float keep = hi ? acc[8 + k] : acc[k], give = hi ? acc[k] : acc[8 + k];
                            s8[k] = (double)keep + (double)__shfl_xor_sync(0xffffffffu, give, 16);
                        }
                    }
                    {
                        const bool hi = lane & 8;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const double keep = hi ? s8[4 + k] : s8[k], give = hi ? s8[k] : s8[4 + k];
                            s4[k] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
                        }
                    }
                    {
                        const bool hi = lane & 4;
#pragma unroll
                        for (int k = 0; k < 2; k++) {
                            const double keep = hi ? s4[2 + k] : s4[k], give = hi ? s4[k] : s4[2 + k];
                            s2[k] = keep + __shfl_xor_sync(0xffffffffu, give, 4);
                        }
                    }
                    {
                        const bool hi = lane & 2;
                        const double keep = hi ? s2[1] : s2[0], give = hi ? s2[0] : s2[1];
                        s1 = keep + __shfl_xor_sync(0xffffffffu, give, 2);
                    }
                    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
                    if (!(lane & 1)) {
                        const int idx = lane >> 1;   // < 8: Re A[idx], else Im A[idx - 8]
                        reinterpret_cast<double *>(redA)[2 * (idx & 7) + (idx >> 3)] = s1;
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
    if (Gm::TMA) {   // the CTA's shared memory must outlive the bulk stores that read it
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
    }
}

}  // namespace jrc
