// jrc_exact.cuh -- the reference-order resolution of range_angle_estimator decisions
// (lib/range_angle_estimator_impl.cc:137-151 arg-max, :209-227 noise window, :234 gate).
//
// The fused and tiled kernels take their decisions on their own float32 map, which differs from the
// staged (oracle-order) arithmetic by FFT rounding (~3e-7 of the map peak).  Wherever such a difference
// could change a published DECISION they do not decide; they mark the record and k_est_exact redoes the
// estimator for that CPI in the staged kernels' arithmetic (radix2_rows of jrc_staged.cuh, ref_pow_abs2,
// the sequential float += double window sum), so that
//   * peak indices and the gate flag of EVERY record equal the staged path's / the CPU oracle's, and
//   * a record whose decision was redone (DET_EXACT) is bit-identical to the staged path's in every field.
// Marks (DetDev::flags):
//   DET_PENDING  set by the fast kernel, cleared here
//   DET_AMB      a second map element lies within EPS_AMB of the maximum: the arg-max is redone on the
//                candidates (taken from the fast |.|^2 map when there is one, else from a full scan)
//   DET_GATE     snr_db / peak_power lie within the error bound of the fast noise estimate of a threshold:
//                the window sum is redone in the reference's order
// One cluster of CTAs per marked CPI; the marked CPIs of a batch come as a device list (FixCtl).
#pragma once
#include "jrc_common.cuh"
#include "jrc_staged.cuh"
#include <cooperative_groups.h>

namespace jrc {

constexpr unsigned DET_PASSED = 1u, DET_EXACT = 2u;
constexpr unsigned DET_PENDING = 0x80000000u, DET_AMB = 0x40000000u, DET_GATE = 0x20000000u;
// relative power margin inside which two map elements are "the same height": measured map error 3e-7 of the
// peak on either element, plus the two roundings of pow(abs(z),2)
constexpr float EPS_AMB = 4e-6f;

struct FixCtl { int count; int done; int n_marked, n_redone, n_inkernel, pad; };   // the last three: running totals (statistics)

// ---------------------------------------------------------------------------
// ONE output bin of the staged radix-2 DIT FFT (radix2_rows) of a zero-padded input, without the transform.
// With n_in = 2^m non-zero inputs in front of n - n_in zeros, the first log2(n / n_in) butterfly levels only copy
// (u +- w*0 = u exactly), and bin o of the rest is a pairwise reduction of the bit-reversed inputs A[k] = x[rev_m(k)]
// with ONE twiddle and ONE sign per level:
//     level l = log2(n/n_in) + s, s = 0..m-1:   A'[k] = A[2k] +- w_l * A[2k+1],
//     w_l = tw[(o mod 2^l) * n / 2^(l+1)],  sign = bit l of o
// -- the same float operations, in the same order, as the transform performs on the way to that bin (checked bin by
// bin against the CPU oracle's FFT: tests/test_oracle_kat.py::test_single_bin_dit).  n_in - 1 complex MACs instead of
// (n/2) log2 n butterflies: what lets a kernel re-decide an ambiguous arg-max in the reference's arithmetic.
// ---------------------------------------------------------------------------
__device__ __forceinline__ c32 dit_level(c32 u, c32 v, const c32 *__restrict__ tw, int log2n, int l, int o)
{
    const c32 w = tw[(o & ((1 << l) - 1)) << (log2n - l - 1)];
    const c32 t = cmul_exact(v, w);
    return ((o >> l) & 1) ? csub_exact(u, t) : cadd_exact(u, t);
}
// 64 inputs, one warp: lane holds the leaves A[2 lane] = x[rev6(2 lane)] and A[2 lane + 1]; result in lane 0
__device__ __forceinline__ c32 dit_bin_warp64(c32 a0, c32 a1, int log2n, const c32 *__restrict__ tw, int o)
{
    const int l0 = log2n - 6;
    c32 val = dit_level(a0, a1, tw, log2n, l0, o);
#pragma unroll
    for (int s = 1; s < 6; s++) {
        c32 other;
        other.x = __shfl_down_sync(0xffffffffu, val.x, 1 << (s - 1));
        other.y = __shfl_down_sync(0xffffffffu, val.y, 1 << (s - 1));
        val = dit_level(val, other, tw, log2n, l0 + s, o);
    }
    return val;
}
// 8 inputs in natural order, one thread
__device__ __forceinline__ c32 dit_bin8(const c32 (&x)[8], int log2n, const c32 *__restrict__ tw, int o)
{
    const int l0 = log2n - 3;
    c32 A[4], B[2];
    // A[k] = x[rev3(k)]: (0,4) (2,6) (1,5) (3,7)
    A[0] = dit_level(x[0], x[4], tw, log2n, l0, o);
    A[1] = dit_level(x[2], x[6], tw, log2n, l0, o);
    A[2] = dit_level(x[1], x[5], tw, log2n, l0, o);
    A[3] = dit_level(x[3], x[7], tw, log2n, l0, o);
    B[0] = dit_level(A[0], A[1], tw, log2n, l0 + 1, o);
    B[1] = dit_level(A[2], A[3], tw, log2n, l0 + 1, o);
    return dit_level(B[0], B[1], tw, log2n, l0 + 2, o);
}

// 32 * 2^lm inputs, one warp, any power-of-two transform length >= the input count: lane holds the 2^lm consecutive leaves
// A[(lane << lm) + idx] = x[rev(...)] (fetched through load(), which returns zero beyond the real inputs: zeros in front of
// the copy levels go through the same u +- w*0 the transform performs), reduces them with a binary-counter stack, then
// five levels across the lanes.  Result in lane 0.
template <class F>
__device__ __forceinline__ c32 dit_bin_warp(F load, int log2_in, int log2n, const c32 *__restrict__ tw, int o, int lane)
{
    const int l0 = log2n - log2_in, lm = log2_in - 5;
    c32 stack[8];
    for (int idx = 0; idx < (1 << lm); idx++) {
        const unsigned g = ((unsigned)lane << lm) + (unsigned)idx;
        c32 val = load((int)(__brev(g) >> (32 - log2_in)));
        int lev = 0;
        for (int k = idx; k & 1; k >>= 1) { val = dit_level(stack[lev], val, tw, log2n, l0 + lev, o); lev++; }
        stack[lev] = val;
    }
    c32 val = stack[lm];
#pragma unroll
    for (int s = 0; s < 5; s++) {
        c32 other;
        other.x = __shfl_down_sync(0xffffffffu, val.x, 1 << s);
        other.y = __shfl_down_sync(0xffffffffu, val.y, 1 << s);
        val = dit_level(val, other, tw, log2n, l0 + lm + s, o);
    }
    return val;
}

// Is the gate decision (:234) safe on the fast path's values?  noise carries the error of the fast window sum against
// the reference's sequential float accumulation (<= n_noise * 2^-24 relative, the worst case of recursive summation of
// positive terms), the FFT rounding of the window samples (each within 3e-7 of the map PEAK amplitude) and of the peak.
__device__ __forceinline__ bool gate_is_marginal(float peak, float noise, float snr_db, int n_noise, float snr_thr, float pow_thr)
{
    const float rel = (float)n_noise * 5.9604645e-8f + 4e-6f * sqrtf(__fdiv_rn(peak, noise)) + 1e-5f;
    const float margin_db = 4.3429448f * rel + 2e-6f * fabsf(snr_db) + 2e-6f * fabsf(snr_thr);
    const bool near_snr = fabsf(snr_db - snr_thr) <= margin_db;            // NaN / inf: false
    const bool near_pow = fabsf(peak - pow_thr) <= 1e-5f * fabsf(peak);
    return near_snr || near_pow;
}

// fast kernels: append a marked CPI (index within the batch) to the list
__device__ __forceinline__ void fix_push(FixCtl *ctl, int *list, int cpi)
{
    list[atomicAdd(&ctl->count, 1)] = cpi;
    atomicAdd(&ctl->n_marked, 1);
}

struct ExactParams {
    PortDev rx, tx;          // symbols (H == nullptr)
    const c32 *H;            // or channel estimates [n_cpi][V][N] (background path)
    int N, T, R, S, n_pre, tx_interleave, V, Nr, Na, log2Nr, log2Na;
    const c32 *tw_r, *tw_a;  // radix-2 tables [n/2]: inverse Nr, forward Na (get_twiddles)
    EstParams est;
    DetDev *dets;
    const float *map;        // fast |.|^2 map [n_cpi][Nr][Na] or nullptr
    int cpi0;
    const int *list;
    FixCtl *ctl;
    c32 *scratch;            // per cluster: [V][Nr] range spectra, then EXACT_WCAP floats of window samples
    size_t scratch_stride;   // in c32
    int buf_elems;           // dynamic shared memory, in c32 (>= max(Nr, Na))
};

constexpr int EXACT_MAX_CAND = 64;
constexpr int EXACT_WCAP = 1 << 16;          // window samples a cluster exchanges at a time (floats)

// One CLUSTER of CTAs per marked CPI (cluster size 1 for the small maps, 8 for the large ones: a 4096 x 256 map with 32
// channels is ~1 M exact butterflies, 0.95 ms on one CTA).  The CTAs of a cluster split every stage by rank and meet at
// cluster barriers; rank 0 owns the candidate list and the per-CTA arg-max keys (the other ranks write them through
// distributed shared memory) and does the one thing that cannot be split, the sequential window sum.  All decisions that
// steer control flow are read from rank 0 after a barrier, so the barriers are cluster-uniform.
struct ExactShared {
    int ncand;
    int cand[EXACT_MAX_CAND];
    unsigned long long keys[16];
};

__global__ void __launch_bounds__(256) k_est_exact(const ExactParams P)
{
    namespace cg = cooperative_groups;
    extern __shared__ c32 sm[];
    __shared__ float chunk[1024];
    __shared__ unsigned long long s_red[8];
    __shared__ ExactShared S;
    __shared__ int s_rows[EXACT_MAX_CAND], s_wrows[EXACT_MAX_CAND];
    __shared__ float s_noise;
    cg::cluster_group cl = cg::this_cluster();
    const int CL = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    const int cid = (int)blockIdx.x / CL, ncl = (int)gridDim.x / CL;
    ExactShared *S0 = cl.map_shared_rank(&S, 0);          // rank 0's copy (== &S on rank 0)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    asm volatile("griddepcontrol.wait;" ::: "memory");     // launched programmatically behind the kernel that fills the list
    const int count = *reinterpret_cast<volatile int *>(&P.ctl->count);
    const int N = P.N, V = P.V, Nr = P.Nr, Na = P.Na;
    c32 *Y = P.scratch + (size_t)cid * P.scratch_stride;
    float *W = reinterpret_cast<float *>(Y + (size_t)V * Nr);      // [EXACT_WCAP] window samples, reference order

    auto block_max = [&](unsigned long long key) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other > key ? other : key;
        }
        __syncthreads();
        if (lane == 0) s_red[warp] = key;
        __syncthreads();
        key = s_red[0];
#pragma unroll
        for (int w = 1; w < 8; w++) key = s_red[w] > key ? s_red[w] : key;
        return key;
    };
    // angle FFT (fft_vcc #B: forward, shift) of map rows rows[0..nrows) into sm, row lr at sm + lr*Na, natural
    // (shifted) order is read through ang()
    auto angle_rows = [&](const int *rows_idx, int row_first, int nrows, bool consecutive) {
        for (int e = tid; e < nrows * Na; e += 256) {
            const int lr = e / Na, i = e % Na;
            const int n = consecutive ? row_first + lr : rows_idx[lr];
            const unsigned rev = __brev((unsigned)i) >> (32 - P.log2Na);
            sm[lr * Na + rev] = i < V ? Y[(size_t)i * Nr + n] : mk(0.f, 0.f);     // matrix_transpose + angle zero-pad
        }
        __syncthreads();
        radix2_rows(sm, Na, nrows, P.tw_a);
    };
    auto ang = [&](int lr, int i) { return sm[lr * Na + ((i + (Na + 1) / 2) % Na)]; };

    if (rank == 0 && tid == 0) S.ncand = 0;
    cl.sync();
    // single-bin evaluation of candidates (no transform): needs at least a warp's worth of points on either axis
    const int log2N = 31 - __clz(N);
    int log2V = 5;
    while ((1 << log2V) < V) log2V++;
    const bool single_bin_ok = P.map && (1 << log2N) == N && N >= 32 && Nr >= N && Na >= (1 << log2V) && V <= 256;
    __shared__ c32 s_y[256];

    for (int it = cid; it < count; it += ncl) {
        const int c = P.list[it];
        const DetDev d0 = P.dets[c];             // (rank 0 rewrites it only behind the barriers below)
        const int rpa = max(1, P.buf_elems / Na);
        // exact conj-MAC (:250-274) of channel p, subcarrier i
        auto chan_est = [&](int p, int i) {
            if (P.H) return P.H[((size_t)c * V + p) * N + i];
            int r, t;
            if (P.tx_interleave) { t = p / P.R; r = p % P.R; } else { r = p / P.T; t = p % P.T; }
            const c32 *prx = P.rx.base + c * P.rx.cpi_stride + r * P.rx.ant_stride + (long long)P.n_pre * N + i;
            const c32 *ptx = P.tx.base + c * P.tx.cpi_stride + t * P.tx.ant_stride + (long long)P.n_pre * N + i;
            c32 v = mk(0.f, 0.f);
            for (int s = 0; s < P.S; s++) {
                const c32 a = prx[(long long)s * N], b = ptx[(long long)s * N];
                v = cadd_exact(v, cmul_exact(a, mk(b.x, -b.y)));
            }
            return v;
        };
        // ---- range spectra of all channels, staged arithmetic: conj-MAC, zero-pad, fft_vcc #A (made when first needed) ----
        bool have_Y = false;
        auto range_spectra = [&]() {
            if (have_Y) return;
            have_Y = true;
            const int rp = max(1, min(V, P.buf_elems / Nr));
            for (int p0 = 0, q = 0; p0 < V; p0 += rp, q++) {
                if (q % CL != rank) continue;
                const int np = min(rp, V - p0);
                for (int e = tid; e < np * Nr; e += 256) {
                    const int lr = e / Nr, i = e % Nr;
                    sm[lr * Nr + (__brev((unsigned)i) >> (32 - P.log2Nr))] = i < N ? chan_est(p0 + lr, i) : mk(0.f, 0.f);
                }
                __syncthreads();
                radix2_rows(sm, Nr, np, P.tw_r);
                for (int e = tid; e < np * Nr; e += 256) Y[(size_t)(p0 + e / Nr) * Nr + e % Nr] = sm[e];
                __syncthreads();
            }
            __threadfence();
            cl.sync();                           // Y complete
        };

        // ---- arg-max (:137-151): first maximum of (float)pow(abs(z),2) in row-major order ----
        unsigned long long key = 0ull;
        bool full_scan = (d0.flags & DET_AMB) && (P.map == nullptr || d0.range_idx < 0);
        if ((d0.flags & DET_AMB) && !full_scan) {
            // candidates: every element of the fast map within 2*EPS_AMB of its maximum; every rank scans its share
            const float *mc = P.map + (size_t)c * Nr * Na;
            const float thr = __fmul_rn(mc[(size_t)d0.range_idx * Na + d0.angle_idx], 1.f - 2.f * EPS_AMB);
            const long long tot4 = ((long long)Nr * Na) >> 2;
            const long long q0 = tot4 * rank / CL, q1 = tot4 * (rank + 1) / CL;
            for (long long q = q0 + tid; q < q1; q += 256 * 8) {
                float4 v4[8];
#pragma unroll
                for (int u = 0; u < 8; u++)
                    v4[u] = (q + 256 * u < q1) ? __ldcs(reinterpret_cast<const float4 *>(mc) + q + 256 * u) : make_float4(-1.f, -1.f, -1.f, -1.f);
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const float vv[4] = {v4[u].x, v4[u].y, v4[u].z, v4[u].w};
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (vv[k] >= thr) {
                            const int slot = atomicAdd(&S0->ncand, 1);
                            if (slot < EXACT_MAX_CAND) S0->cand[slot] = (int)((q + 256 * u) * 4 + k);
                        }
                }
            }
            cl.sync();
            const int ncand = S0->ncand;
            if (ncand > EXACT_MAX_CAND || ncand == 0) full_scan = true;      // cluster-uniform
        }
        if (full_scan) {
            range_spectra();
            for (int n0 = 0, q = 0; n0 < Nr; n0 += rpa, q++) {
                if (q % CL != rank) continue;
                const int nr = min(rpa, Nr - n0);
                angle_rows(nullptr, n0, nr, true);
                for (int e = tid; e < nr * Na; e += 256) {
                    const int lr = e / Na, i = e % Na;
                    const float pw = (float)ref_pow_abs2(ang(lr, i));
                    if (pw == pw) {
                        const unsigned long long k2 = pack_key(pw, (unsigned)((n0 + lr) * Na + i));
                        key = k2 > key ? k2 : key;
                    }
                }
                __syncthreads();
            }
        } else if (single_bin_ok) {
            // every candidate cell in the staged arithmetic WITHOUT the transforms (dit_bin_warp): per channel one bin of the
            // zero-padded range IFFT (a warp each), then one bin of the angle FFT over the channels.  The candidates are
            // dealt round the ranks.  (The unmarked arg-max is a single candidate: the fast path's own peak.)
            const int nc = (d0.flags & DET_AMB) ? S0->ncand : 1;
            for (int j = rank; j < nc; j += CL) {
                const int lin = (d0.flags & DET_AMB) ? S0->cand[j] : d0.range_idx * Na + d0.angle_idx;
                const int nn = lin / Na, ii = lin % Na;
                for (int p = warp; p < V; p += 8) {
                    const c32 y = dit_bin_warp([&](int i) { return chan_est(p, i); }, log2N, P.log2Nr, P.tw_r, nn, lane);
                    if (lane == 0) s_y[p] = y;
                }
                __syncthreads();
                if (warp == 0) {
                    const c32 z = dit_bin_warp([&](int i) { return i < V ? s_y[i] : mk(0.f, 0.f); }, log2V, P.log2Na, P.tw_a,
                                               (ii + (Na + 1) / 2) % Na, lane);
                    if (lane == 0) {
                        const float pw = (float)ref_pow_abs2(z);
                        if (pw == pw) {
                            const unsigned long long k2 = pack_key(pw, (unsigned)lin);
                            key = k2 > key ? k2 : key;
                        }
                    }
                }
                __syncthreads();
            }
        } else {
            range_spectra();
            if (rank == 0) {
                // one row per candidate
                const int nc = (d0.flags & DET_AMB) ? S.ncand : 1;
                for (int c0 = 0; c0 < nc; c0 += min(rpa, EXACT_MAX_CAND)) {
                    const int nr = min(min(rpa, EXACT_MAX_CAND), nc - c0);
                    if (tid < nr) s_rows[tid] = ((d0.flags & DET_AMB) ? S.cand[c0 + tid] : d0.range_idx * Na + d0.angle_idx) / Na;
                    __syncthreads();
                    angle_rows(s_rows, 0, nr, false);
                    if (tid < nr) {
                        const int lin = (d0.flags & DET_AMB) ? S.cand[c0 + tid] : d0.range_idx * Na + d0.angle_idx;
                        const float pw = (float)ref_pow_abs2(ang(tid, lin % Na));
                        if (pw == pw) key = pack_key(pw, (unsigned)lin);
                    }
                    __syncthreads();
                }
            }
        }
        key = block_max(key);
        if (tid == 0) S0->keys[rank] = key;
        cl.sync();
        key = S0->keys[0];
        for (int r = 1; r < CL; r++) { const unsigned long long k2 = S0->keys[r]; key = k2 > key ? k2 : key; }
        if (key == 0ull) {       // NaN-only map: nothing wins the strict '>' scan
            if (rank == 0 && tid == 0) {
                DetDev d; d.range_idx = -1; d.angle_idx = -1; d.peak_power = -1.f;
                d.noise_power = __int_as_float(0x7fc00000); d.snr_db = d.noise_power;
                d.n_noise = 0; d.flags = DET_EXACT; d.cpi = P.cpi0 + c;
                P.dets[c] = d;
                S.ncand = 0;
            }
            cl.sync();           // rank 0's shared state is reused by the next marked CPI
            continue;
        }
        const unsigned lin = 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull);
        const int nstar = (int)(lin / (unsigned)Na), istar = (int)(lin % (unsigned)Na);
        const float peak = __uint_as_float((unsigned)(key >> 32));
        if (!(d0.flags & DET_GATE) && nstar == d0.range_idx && istar == d0.angle_idx) {
            // the candidates' order is the fast path's: its window, noise estimate and (safe) gate decision stand
            if (rank == 0 && tid == 0) { P.dets[c].flags = d0.flags & DET_PASSED; S.ncand = 0; }
            cl.sync();
            continue;
        }
        range_spectra();         // (a no-op when the arg-max already needed them)

        // ---- noise window (:152-227) in the reference's order ----
        const NoiseWin w = noise_window(P.est, nstar, istar);
        const int ncols = w.end_a - w.start_a, nrows = w.end_r - w.start_r;
        const int total = (ncols > 0 && nrows > 0) ? nrows * ncols : 0;
        if (tid == 0) s_noise = 0.f;
        __syncthreads();
        const int rb = min(rpa, EXACT_MAX_CAND);             // rows per angle_rows call
        if (total > 0 && CL > 1 && ncols <= EXACT_WCAP) {
            // the ranks fill W with the window's sqrt-power samples, rows_seg rows at a time; rank 0 adds them up in order
            const int rows_seg = max(1, EXACT_WCAP / ncols);
            for (int s0 = 0; s0 < nrows; s0 += rows_seg) {
                const int ns = min(rows_seg, nrows - s0);
                for (int r0 = 0, q = 0; r0 < ns; r0 += rb, q++) {
                    if (q % CL != rank) continue;
                    const int nr = min(rb, ns - r0);
                    if (tid < nr) { const int ir = w.start_r + s0 + r0 + tid; s_wrows[tid] = ((ir % Nr) + Nr) % Nr; }
                    __syncthreads();
                    angle_rows(s_wrows, 0, nr, false);
                    for (int e = tid; e < nr * ncols; e += 256) {
                        const int lr = e / ncols, j = e % ncols;
                        const int ia = w.start_a + j, a_idx = ((ia % Na) + Na) % Na;
                        W[(size_t)(r0 + lr) * ncols + j] = ref_abs(ang(lr, a_idx));
                    }
                    __syncthreads();
                }
                __threadfence();
                cl.sync();
                if (rank == 0) {
                    const int cells = ns * ncols;
                    for (int cb = 0; cb < cells; cb += 1024) {
                        const int cnt = min(1024, cells - cb);
                        for (int j = tid; j < cnt; j += 256) chunk[j] = __ldcg(W + cb + j);
                        __syncthreads();
                        if (tid < 32) {
                            const float acc = seq_sum_sq_warp(s_noise, chunk, cnt, tid);        // :217 float += double
                            if (tid == 0) s_noise = acc;
                        }
                        __syncthreads();
                    }
                }
                if (s0 + rows_seg < nrows) cl.sync();        // W is rewritten by the next segment
            }
        } else if (total > 0 && rank == 0) {
            for (int r0 = 0; r0 < nrows; r0 += rb) {
                const int nr = min(rb, nrows - r0);
                if (tid < nr) { const int ir = w.start_r + r0 + tid; s_wrows[tid] = ((ir % Nr) + Nr) % Nr; }
                __syncthreads();
                angle_rows(s_wrows, 0, nr, false);
                for (int lr = 0; lr < nr; lr++)
                    for (int cb = 0; cb < ncols; cb += 1024) {
                        const int cnt = min(1024, ncols - cb);
                        for (int j = tid; j < cnt; j += 256) {
                            const int ia = w.start_a + cb + j, a_idx = ((ia % Na) + Na) % Na;
                            chunk[j] = ref_abs(ang(lr, a_idx));
                        }
                        __syncthreads();
                        if (tid < 32) {
                            const float acc = seq_sum_sq_warp(s_noise, chunk, cnt, tid);        // :217 float += double
                            if (tid == 0) s_noise = acc;
                        }
                        __syncthreads();
                    }
            }
        }
        if (rank == 0 && tid == 0) {
            DetDev d;
            d.range_idx = nstar; d.angle_idx = istar; d.peak_power = peak; d.n_noise = total;
            d.noise_power = __fdiv_rn(s_noise, (float)total);                                   // :226
            d.snr_db = snr_db_of(d.peak_power, d.noise_power);                                  // :227
            d.flags = ((d.snr_db >= P.est.snr_threshold && d.peak_power >= P.est.power_threshold) ? DET_PASSED : 0u) | DET_EXACT;
            d.cpi = P.cpi0 + c;
            P.dets[c] = d;
            atomicAdd(&P.ctl->n_redone, 1);
            S.ncand = 0;
        }
        cl.sync();               // rank 0's shared state and W are reused by the next marked CPI
    }
    // the last CTA to finish re-arms the list for the next batch
    __shared__ bool s_last;
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(&P.ctl->done, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && tid == 0) { P.ctl->count = 0; P.ctl->done = 0; __threadfence(); }
    cl.sync();                   // no CTA leaves while its shared memory may still be read by another rank
}

}  // namespace jrc
