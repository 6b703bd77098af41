// jrc_stream.cuh -- "slice-streaming" form of the fused radar chain for the 64-subcarrier /
// 8-virtual-channel family: an opt-in A/B kernel (JRC_FUSED_KERNEL=stream); measured 14.7 M CPI/s on configs[1]
// against 18 M for k_fused64x8 (profiles/README.md), kept because its footprint does not grow with the zero-pad.
//
// Same arithmetic as k_fused64x8 (jrc_fused.cuh: three passes of "twiddle 8 inputs, DFT-8"), but
// organised so that a WARP, not a CTA, is the unit of execution and no CTA barrier exists:
//
//   range  y[p][Q*m1 + q0 + IR*m0] = sum_{k0} W_8^{k0 m1} ( W_Nr^{k0 q} * B[p][k0][m0] ),  q = q0 + IR*m0
//          B[p][k0][m0]            = sum_{k1} W_8^{k1 m0} ( W_Q^{k1 q0} * H[p][k0 + 8 k1] )
//
// For a fixed residue q0 ("slice") the two range passes touch 8x8x8 values and yield 64 complete
// range bins for all 8 channels -- 4.6 KiB of shared memory instead of the whole 64 KiB range
// spectrum.  A work unit is (CPI, SPU consecutive slices); warps pull units round-robin, run
// pass 1 -> pass 2 -> angle pass + |.|^2 + store for each slice with __syncwarp only, and fold
// their running maximum into one 64-bit atomicMax per CPI.  Twenty independent warps per SM sit in
// different phases, so the FP32 pipe, the LSU and the issue slots are shared by all stages instead
// of being claimed by one stage at a time (profiles/README.md, round 1).
//
// Kernels around it: k_chan_est (jrc_staged.cuh) produces H[cpi][8][64] (4 KiB per CPI, L2
// resident), k_stream_finalize turns the per-CPI key into the detection record: it re-runs the
// winning slice bit-identically to recover the angle bin and the complex peak, and reads the noise
// window back from the map this kernel wrote.
#pragma once
#include "jrc_common.cuh"
#include "jrc_staged.cuh"

namespace jrc {

struct StreamParams {
    const c32 *H;                 // [n_cpi][8][64]
    int n_cpi, cpi0;
    float *map;                   // [n_cpi][NR][NA]
    unsigned long long *keys;     // [n_cpi] zero-initialised, or nullptr (no detections wanted)
    const c32 *tw1g;              // [IR][8]      W_Q^{k1 q0}
    const c32 *tw2g;              // [IR][8][8]   W_Nr^{k0 (q0 + IR m0)}  as [q0][k0][m0]
    DetDev *dets;                 // finalize only
    EstParams est;                // finalize only
};

template <int IR, int IA>
struct StreamGeom {
    static constexpr int NSC = 64, V = 8;
    static constexpr int NR = NSC * IR, NA = V * IA, Q = NR / 8;
    static constexpr int G = 32 / IA;                 // range bins per angle iteration
    static constexpr int ITERS = 64 / G;              // angle iterations per slice
    static constexpr int HST = 72;                    // padded row stride (complex) of the per-warp tiles
    static constexpr int WARP_SMEM = 2 * 8 * HST * 8 + 256 * 4;   // H tile + B/y tile + 1 KiB staging
    static_assert(IA >= 4 && IA <= 32 && (IA & (IA - 1)) == 0, "angle interp must be 4..32, power of two");
    static_assert(IR >= 1 && IR <= 64 && (IR & (IR - 1)) == 0, "range interp must be a power of two");
};

// slot of element (a, b) of an 8x8 tile, skewed so that both "fixed a, lanes over b" and
// "fixed b, lanes over a" sweep 8 distinct 8-byte bank pairs
__device__ __forceinline__ int skew(int a, int b) { return b * 8 + ((a + b) & 7); }

// range pass 1 + 2 of one slice q0 for this warp's CPI: Hs -> Ys (y[p][m1][m0] at Ys[p*HST + skew(m1, m0)])
template <int IR, int IA>
__device__ __forceinline__ void range_slice(const c32 *Hs, c32 *Ys, int q0, int lane, const c32 *__restrict__ tw1g,
                                            const c32 *__restrict__ tw2g)
{
    constexpr int HST = StreamGeom<IR, IA>::HST;
    const int lo = lane & 7, ph = lane >> 3;
    c32 tw[8];
#pragma unroll
    for (int k1 = 1; k1 < 8; k1++) tw[k1] = __ldg(tw1g + q0 * 8 + k1);
    // pass 1: task (p, k0 = lo)
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int p = ph + 4 * j;
        c32 u[8];
#pragma unroll
        for (int k1 = 0; k1 < 8; k1++) u[k1] = Hs[p * HST + lo + 8 * k1];
#pragma unroll
        for (int k1 = 1; k1 < 8; k1++) u[k1] = cmul_fma(u[k1], tw[k1]);
        JRC_FFT8<1>(u);
#pragma unroll
        for (int m0 = 0; m0 < 8; m0++) Ys[p * HST + skew(lo, m0)] = u[m0];
    }
    __syncwarp();
    // pass 2: task (p, m0 = lo), in place
#pragma unroll
    for (int k0 = 1; k0 < 8; k0++) tw[k0] = __ldg(tw2g + (q0 * 8 + k0) * 8 + lo);
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int p = ph + 4 * j;
        c32 u[8];
#pragma unroll
        for (int k0 = 0; k0 < 8; k0++) u[k0] = Ys[p * HST + skew(k0, lo)];
#pragma unroll
        for (int k0 = 1; k0 < 8; k0++) u[k0] = cmul_fma(u[k0], tw[k0]);
        JRC_FFT8<1>(u);
#pragma unroll
        for (int m1 = 0; m1 < 8; m1++) Ys[p * HST + skew(m1, lo)] = u[m1];
    }
    __syncwarp();
}

// one angle task of the slice-local row rho (= m1*8 + m0): twiddle, forward DFT-8, |.|^2
template <int IR, int IA>
__device__ __forceinline__ void angle_slice_task(const c32 *Ys, int rho, const c32 (&tw)[8], c32 (&u)[8], float (&v)[8])
{
    constexpr int HST = StreamGeom<IR, IA>::HST;
    const int slot = skew(rho >> 3, rho & 7);
#pragma unroll
    for (int p = 0; p < 8; p++) u[p] = Ys[p * HST + slot];
#pragma unroll
    for (int p = 1; p < 8; p++) u[p] = cmul_fma(u[p], tw[p]);
    JRC_FFT8<-1>(u);
#pragma unroll
    for (int a = 0; a < 8; a++) {   // volk_32fc_magnitude_squared_32f: re*re + im*im, each product rounded
        c32 sq = __fmul2_rn(u[a], u[a]);
        v[a] = __fadd_rn(sq.x, sq.y);
    }
}

template <int IR, int IA>
__device__ __forceinline__ void load_H_tile(c32 *Hs, const c32 *__restrict__ Hg, int lane)
{
    constexpr int HST = StreamGeom<IR, IA>::HST;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const int e = lane + 32 * j;
        Hs[(e >> 6) * HST + (e & 63)] = __ldg(Hg + e);
    }
    __syncwarp();
}

template <int IR, int IA, int SPU, int WPC>
__global__ void __launch_bounds__(WPC * 32, 5) k_stream64x8(const StreamParams P)
{
    using Gm = StreamGeom<IR, IA>;
    constexpr int NR = Gm::NR, NA = Gm::NA, Q = Gm::Q, G = Gm::G, ITERS = Gm::ITERS, HST = Gm::HST;
    constexpr int UPC = IR / SPU;     // units per CPI
    static_assert(IR % SPU == 0, "slices per unit must divide the range interpolation factor");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    unsigned char *wbase = smem_raw + warp * Gm::WARP_SMEM;
    c32 *Hs = reinterpret_cast<c32 *>(wbase);                 // [8][HST]
    c32 *Ys = Hs + 8 * HST;                                   // [8][HST]  B then y of the current slice
    float *wstg = reinterpret_cast<float *>(Ys + 8 * HST);    // [G][NA] staging tile (1 KiB)

    // angle pass lane role: task (row group g, bin residue b); twiddle (-1)^p w_Na^{p (b + IA*rot)}
    const int b = lane % IA, g = lane / IA, rot = g;
    c32 tw3[8];
#pragma unroll
    for (int j = 0; j < 8; j++) tw3[j] = cispi_ratio(j * (NA - 2 * (b + IA * rot)), NA);
    float *sp[8];
#pragma unroll
    for (int a = 0; a < 8; a++) sp[a] = wstg + g * NA + b + IA * ((a + rot) & 7);
    // store role: this lane moves floats [4*lane, 4*lane+4) and [128 + 4*lane, ...) of the staged tile
    const int tr0 = (4 * lane) / NA, tc0 = (4 * lane) % NA;
    const int tr1 = (4 * lane + 128) / NA, tc1 = (4 * lane + 128) % NA;

    const long long total_units = (long long)P.n_cpi * UPC;
    const long long stride = (long long)gridDim.x * WPC;
    for (long long unit = (long long)blockIdx.x * WPC + warp; unit < total_units; unit += stride) {
        const int cpi = (int)(unit / UPC), sg = (int)(unit % UPC);
        load_H_tile<IR, IA>(Hs, P.H + (long long)cpi * 512, lane);
        float *map_c = P.map + (long long)cpi * NR * NA;
        float best = -1.f;
        int best_n = 0;
        for (int s = 0; s < SPU; s++) {
            const int q0 = sg * SPU + s;
            range_slice<IR, IA>(Hs, Ys, q0, lane, P.tw1g, P.tw2g);
#pragma unroll 4
            for (int it = 0; it < ITERS; it++) {
                const int rho = it * G + g;
                c32 u[8];
                float v[8];
                angle_slice_task<IR, IA>(Ys, rho, tw3, u, v);
                float m8 = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])),
                                 fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
                const int n = Q * (rho >> 3) + IR * (rho & 7) + q0;
                // first maximum in row-major order: larger value, or equal value in an earlier row
                if (m8 > best || (m8 == best && n < best_n)) { best = m8; best_n = n; }
#pragma unroll
                for (int a = 0; a < 8; a++) sp[a][0] = v[a];
                __syncwarp();
                const float4 o0 = reinterpret_cast<const float4 *>(wstg)[lane];
                const float4 o1 = reinterpret_cast<const float4 *>(wstg)[lane + 32];
                __syncwarp();
                const int r0 = it * G + tr0, r1 = it * G + tr1;
                const int n0 = Q * (r0 >> 3) + IR * (r0 & 7) + q0, n1 = Q * (r1 >> 3) + IR * (r1 & 7) + q0;
                __stcs(reinterpret_cast<float4 *>(map_c + (long long)n0 * NA + tc0), o0);
                __stcs(reinterpret_cast<float4 *>(map_c + (long long)n1 * NA + tc1), o1);
            }
        }
        if (P.keys) {
            unsigned long long key = best >= 0.f ? pack_key(best, (unsigned)best_n) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                key = other > key ? other : key;
            }
            if (lane == 0 && key) atomicMax(P.keys + cpi, key);
        }
    }
}

// One warp per CPI: detection record from the per-CPI key (lib/range_angle_estimator_impl.cc:152-253).
template <int IR, int IA, int WPC>
__global__ void __launch_bounds__(WPC * 32) k_stream_finalize(const StreamParams P)
{
    using Gm = StreamGeom<IR, IA>;
    constexpr int NR = Gm::NR, NA = Gm::NA, Q = Gm::Q, G = Gm::G, HST = Gm::HST;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int cpi = blockIdx.x * WPC + warp;
    if (cpi >= P.n_cpi) return;
    unsigned char *wbase = smem_raw + warp * Gm::WARP_SMEM;
    c32 *Hs = reinterpret_cast<c32 *>(wbase);
    c32 *Ys = Hs + 8 * HST;

    const unsigned long long key = P.keys[cpi];
    if (key == 0ull) {   // NaN-only input: nothing can win the strict '>' scan
        if (lane == 0) {
            DetDev d; d.range_idx = -1; d.angle_idx = -1; d.peak_power = -1.f;
            d.noise_power = __int_as_float(0x7fc00000); d.snr_db = d.noise_power;
            d.n_noise = 0; d.flags = 0; d.cpi = P.cpi0 + cpi;
            P.dets[cpi] = d;
        }
        return;
    }
    const float gmax = __uint_as_float((unsigned)(key >> 32));
    const int nstar = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
    // row -> slice coordinates: n = Q*m1 + IR*m0 + q0
    const int q = nstar % Q, m1 = nstar / Q, q0 = q % IR, m0 = q / IR;
    const int rho = m1 * 8 + m0;
    const int b = lane % IA, g = lane / IA, rot = g;
    c32 tw3[8];
#pragma unroll
    for (int j = 0; j < 8; j++) tw3[j] = cispi_ratio(j * (NA - 2 * (b + IA * rot)), NA);
    load_H_tile<IR, IA>(Hs, P.H + (long long)cpi * 512, lane);
    range_slice<IR, IA>(Hs, Ys, q0, lane, P.tw1g, P.tw2g);
    int icand = 0x7fffffff;
    c32 zc = mk(0.f, 0.f);
    if (g == rho % G) {   // the lanes that evaluated this row in the main kernel, same twiddles -> same bits
        c32 u[8]; float v[8];
        angle_slice_task<IR, IA>(Ys, rho, tw3, u, v);
#pragma unroll
        for (int a = 0; a < 8; a++) {
            int i = b + IA * ((a + rot) & 7);
            if (v[a] == gmax && i < icand) { icand = i; zc = u[a]; }
        }
    }
    int imin = icand;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) imin = min(imin, __shfl_xor_sync(0xffffffffu, imin, o));
    // peak power with the reference's pow(abs(z),2) evaluation, broadcast from the owning lane
    float pk = (icand == imin && imin != 0x7fffffff) ? (float)ref_pow_abs2(zc) : -1.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pk = fmaxf(pk, __shfl_xor_sync(0xffffffffu, pk, o));
    if (imin == 0x7fffffff) {   // cannot happen (the row is re-evaluated bit-identically); fail loudly
        if (lane == 0) {
            DetDev d; d.range_idx = nstar; d.angle_idx = -1; d.peak_power = gmax;
            d.noise_power = __int_as_float(0x7fc00000); d.snr_db = d.noise_power; d.n_noise = -1; d.flags = 0x80000000u;
            d.cpi = P.cpi0 + cpi;
            P.dets[cpi] = d;
        }
        return;
    }
    // noise window (:197-226) from the |.|^2 map this CPI's warps wrote
    const NoiseWin w = noise_window(P.est, nstar, imin);
    const int ncols = w.end_a - w.start_a, nrows = w.end_r - w.start_r;
    const int total = (ncols > 0 && nrows > 0) ? nrows * ncols : 0;
    const float *map_c = P.map + (long long)cpi * NR * NA;
    double acc = 0.0;
    for (int j = lane; j < total; j += 32) {
        const int ir = w.start_r + j / ncols, ia = w.start_a + j % ncols;
        const int r_idx = ((ir % NR) + NR) % NR, a_idx = ((ia % NA) + NA) % NA;
        acc += (double)__ldcg(map_c + (long long)r_idx * NA + a_idx);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        DetDev d;
        d.range_idx = nstar; d.angle_idx = imin; d.peak_power = pk; d.n_noise = total;
        d.noise_power = __fdiv_rn((float)acc, (float)total);
        d.snr_db = __fmul_rn(10.f, log10f(__fdiv_rn(d.peak_power, d.noise_power)));
        d.flags = (d.snr_db >= P.est.snr_threshold && d.peak_power >= P.est.power_threshold) ? 1u : 0u;
        d.cpi = P.cpi0 + cpi;
        P.dets[cpi] = d;
    }
}

}  // namespace jrc
