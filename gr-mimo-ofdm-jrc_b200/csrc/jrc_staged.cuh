// jrc_staged.cuh -- one kernel per reference block ("staged" path).
//
// These kernels keep the exact per-block semantics of the reference flowgraph
// (mimo_ofdm_radar -> fft_vcc -> matrix_transpose -> fft_vcc -> mag^2 / estimator)
// and the exact float evaluation order of the CPU oracle, so that they are
// bit-identical to it.  They serve the per-block C-ABI entry points, every
// configuration the fused kernel has no specialisation for, and background removal.
// Paths in comments are relative to the reference tree.
#pragma once
#include "jrc_common.cuh"

namespace jrc {

// ---------------------------------------------------------------------------
// mimo_ofdm_radar conj-MAC  (lib/mimo_ofdm_radar_impl.cc:250-274)
// H[cpi][p][k] = sum_s rx[r][(n_pre+s)*N+k] * conj(tx[t][(n_pre+s)*N+k]), s ascending,
// p = r*T+t (or t*R+r when tx_interleave, :262-269).  One thread per (cpi,p,k).
// ---------------------------------------------------------------------------
struct PortDev { const c32 *base; long long cpi_stride; long long ant_stride; };

__global__ void k_chan_est(PortDev rx, PortDev tx, int n_cpi, int N, int T, int R, int S, int n_pre,
                           int tx_interleave, c32 *__restrict__ H /* [n_cpi][V][N] */)
{
    const int V = T * R;
    const long long total = (long long)n_cpi * V * N;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        int k = (int)(e % N);
        int p = (int)((e / N) % V);
        long long cpi = e / ((long long)N * V);
        int r, t;
        if (tx_interleave) { t = p / R; r = p % R; } else { r = p / T; t = p % T; }
        const c32 *prx = rx.base + cpi * rx.cpi_stride + r * rx.ant_stride + (long long)n_pre * N + k;
        const c32 *ptx = tx.base + cpi * tx.cpi_stride + t * tx.ant_stride + (long long)n_pre * N + k;
        c32 acc = mk(0.f, 0.f);
        for (int s = 0; s < S; s++) {
            c32 a = prx[(long long)s * N];
            c32 b = ptx[(long long)s * N];
            c32 prod = cmul_exact(a, mk(b.x, -b.y));
            acc = cadd_exact(acc, prod);
        }
        H[e] = acc;
    }
}

// Same conj-MAC, register tiled for large arrays: one thread owns subcarrier k of a TT x RR block of
// (tx, rx) antennas, so every symbol sample is loaded once per block instead of once per channel
// (TT + RR loads per TT*RR products).  Per channel the operation order is that of k_chan_est: the
// estimates are bit-identical.
template <int TT, int RR>
__global__ void __launch_bounds__(256) k_chan_est_tile(PortDev rx, PortDev tx, int n_cpi, int N, int T, int R, int S, int n_pre,
                                                       int tx_interleave, c32 *__restrict__ H /* [n_cpi][V][N] */)
{
    const int TB = T / TT, RB = R / RR, V = T * R;
    const long long total = (long long)n_cpi * TB * RB * N;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(e % N);
        const int blk = (int)((e / N) % (TB * RB));
        const long long cpi = e / ((long long)N * TB * RB);
        const int t0 = (blk % TB) * TT, r0 = (blk / TB) * RR;
        const c32 *prx = rx.base + cpi * rx.cpi_stride + r0 * rx.ant_stride + (long long)n_pre * N + k;
        const c32 *ptx = tx.base + cpi * tx.cpi_stride + t0 * tx.ant_stride + (long long)n_pre * N + k;
        c32 acc[RR][TT];
#pragma unroll
        for (int r = 0; r < RR; r++)
#pragma unroll
            for (int t = 0; t < TT; t++) acc[r][t] = mk(0.f, 0.f);
#pragma unroll 2
        for (int s = 0; s < S; s++) {
            c32 a[RR], b[TT];
#pragma unroll
            for (int r = 0; r < RR; r++) a[r] = prx[r * rx.ant_stride + (long long)s * N];
#pragma unroll
            for (int t = 0; t < TT; t++) { c32 v = ptx[t * tx.ant_stride + (long long)s * N]; b[t] = mk(v.x, -v.y); }
#pragma unroll
            for (int r = 0; r < RR; r++)
#pragma unroll
                for (int t = 0; t < TT; t++) acc[r][t] = cadd_exact(acc[r][t], cmul_exact(a[r], b[t]));
        }
#pragma unroll
        for (int r = 0; r < RR; r++)
#pragma unroll
            for (int t = 0; t < TT; t++) {
                const int p = tx_interleave ? (t0 + t) * R + (r0 + r) : (r0 + r) * T + (t0 + t);
                H[(cpi * V + p) * N + k] = acc[r][t];
            }
    }
}

// ---------------------------------------------------------------------------
// background record / removal  (lib/mimo_ofdm_radar_impl.cc:276-300)
// One thread per channel-estimate element; the CPIs of the batch are walked in
// order because the ring buffer evolves frame by frame.  ring: [record_len][VN]
// (slot (head+b)%record_len is the b-th oldest), temp: radar_chan_est_temp.
// size/head are the ring state BEFORE this batch; the host advances them.
// ---------------------------------------------------------------------------
__global__ void k_background(c32 *__restrict__ H /* [n_cpi][VN] in: raw, out: subtracted */,
                             int n_cpi, int VN, c32 *__restrict__ ring, c32 *__restrict__ temp,
                             int record_len, int size0, int head0, int recording, int removal)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= VN) return;
    int size = size0, head = head0;
    c32 tmp = temp[idx];
    for (int c = 0; c < n_cpi; c++) {
        c32 h = H[(long long)c * VN + idx];
        if (recording) tmp = h;                                   // :276-279
        if (removal) {
            c32 mean = mk(0.f, 0.f);
            for (int b = 0; b < size; b++) {                      // :286-289
                c32 e = ring[(long long)((head + b) % record_len) * VN + idx];
                mean.x = __fadd_rn(mean.x, __fdiv_rn(e.x, (float)size));
                mean.y = __fadd_rn(mean.y, __fdiv_rn(e.y, (float)size));
            }
            h = csub_exact(h, mean);                              // :291
            H[(long long)c * VN + idx] = h;
            if (record_len > 0) {                                 // :297-300 push_back
                if (size < record_len) { ring[(long long)((head + size) % record_len) * VN + idx] = tmp; size++; }
                else { ring[(long long)head * VN + idx] = tmp; head = (head + 1) % record_len; }
            }
        }
    }
    temp[idx] = tmp;
}

// zero-padded copy H[cpi][p][0:N] -> out[cpi][p][0:N*interp]  (:243, :312-315)
// (copy, optional: the unpadded rows once more, for the block's capture_radar_data -- :348-370)
__global__ void k_pad_rows(const c32 *__restrict__ H, c32 *__restrict__ out, long long rows, int N, int Nout,
                           c32 *__restrict__ copy)
{
    const long long total = rows * Nout;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        int n = (int)(e % Nout);
        long long row = e / Nout;
        c32 v = mk(0.f, 0.f);
        if (n < N) {
            v = H[row * N + n];
            if (copy) copy[row * N + n] = v;
        }
        out[e] = v;
    }
}

// ---------------------------------------------------------------------------
// gr::fft::fft_vcc  (GNU Radio 3.8 gr-fft semantics): batched radix-2
// DIT in shared memory with the oracle's butterfly order and the same float32
// twiddle table (rounded from double on the host) -> bit-identical to the oracle.
// Row r reads n_in valid samples at in + r*in_stride (the rest of the n-point
// input is zero: this is how the range zero-pad is never materialised) and writes
// n samples at out + r*n.  rows_per_cta rows share a CTA when n is small.
// ---------------------------------------------------------------------------
// butterflies of `rows` n-point transforms that sit bit-reversed in shared memory (row r at sm + r*n); every
// thread of the CTA takes part.  Shared by k_fft_rows and the reference-order estimator (jrc_exact.cuh): the
// float operation order below IS the oracle's.
__device__ __forceinline__ void radix2_rows(c32 *sm, int n, int rows, const c32 *__restrict__ tw)
{
    // n is a power of two: every index below is a shift or a mask (integer divisions were 40 % of this loop's instructions)
    const int log2n = __ffs(n) - 1, hmask = (n >> 1) - 1;
    if (log2n == 0) return;
    const int nbf = rows << (log2n - 1);
    for (int lh = 0; lh < log2n; lh++) {                 // len = 2 << lh, half = 1 << lh, step = n / len
        const int half = 1 << lh, tshift = log2n - lh - 1;
        for (int b = threadIdx.x; b < nbf; b += blockDim.x) {
            const int lr = b >> (log2n - 1), bb = b & hmask;
            const int grp = bb >> lh, j = bb & (half - 1);
            const int i0 = (lr << log2n) + (grp << (lh + 1)) + j, i1 = i0 + half;
            c32 w = tw[j << tshift];
            c32 u = sm[i0];
            c32 v = cmul_exact(sm[i1], w);
            sm[i0] = cadd_exact(u, v);
            sm[i1] = csub_exact(u, v);
        }
        __syncthreads();
    }
}

// Two optional epilogues for the one-frame chain of the fused mode (same values, fewer launches on its critical path):
//   tr_w > 0: the spectra leave transposed, out[i * tr_w + row] -- matrix_transpose (lib/matrix_transpose_impl.cc:97-104)
//             into an array whose padding columns the caller zeroed once;
//   keys:     the rows are ONE map: arg-max key of (float)pow(abs(z),2) over the outputs, exactly k_est_argmax's
//             (lib/range_angle_estimator_impl.cc:137-151); *keys zeroed by the caller.
__global__ void k_fft_rows(const c32 *in, long long in_stride, int n_in,   // in may alias out (in place)
                           c32 *out, int n, int log2n, long long rows, int rows_per_cta,
                           int forward, int shift, const c32 *__restrict__ tw /* [n/2] */,
                           int tr_w, unsigned long long *keys)
{
    extern __shared__ c32 sm[];
    const long long row0 = (long long)blockIdx.x * rows_per_cta;
    const int offset = (n + 1) / 2;
    const int tot = rows_per_cta * n;
    for (int e = threadIdx.x; e < tot; e += blockDim.x) {
        int lr = e / n, i = e % n;
        long long row = row0 + lr;
        c32 v = mk(0.f, 0.f);
        if (row < rows) {
            int src = (!forward && shift) ? (i + offset) % n : i;   // swap input halves
            if (src < n_in) v = in[row * in_stride + src];
        }
        unsigned rev = (log2n == 0) ? 0u : (__brev((unsigned)i) >> (32 - log2n));
        sm[lr * n + rev] = v;
    }
    __syncthreads();
    radix2_rows(sm, n, rows_per_cta, tw);
    unsigned long long best = 0ull;
    for (int e = threadIdx.x; e < tot; e += blockDim.x) {
        int lr = e / n, i = e % n;
        long long row = row0 + lr;
        if (row < rows) {
            int src = (forward && shift) ? (i + offset) % n : i;    // swap output halves
            const c32 v = sm[lr * n + src];
            if (tr_w) out[(long long)i * tr_w + row] = v;
            else out[row * n + i] = v;
            if (keys) {
                const float p = (float)ref_pow_abs2(v);
                if (p == p) {
                    const unsigned long long key = pack_key(p, (unsigned)(row * n + i));
                    best = key > best ? key : best;
                }
            }
        }
    }
    if (keys) {
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        __shared__ unsigned long long wbest[32];
        if ((threadIdx.x & 31) == 0) wbest[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x < 32) {
            best = threadIdx.x < (blockDim.x >> 5) ? wbest[threadIdx.x] : 0ull;
            for (int o = 16; o > 0; o >>= 1) {
                unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                best = other > best ? other : best;
            }
            if (threadIdx.x == 0 && best) atomicMax(keys, best);
        }
    }
}

// ---------------------------------------------------------------------------
// RX OFDM front end for a whole batch: ofdm_cyclic_prefix_remover
// (lib/ofdm_cyclic_prefix_remover_impl.cc:92-95) + fft_vxx(forward, shift) (...radar_sim.grc:898-939) for the n_sym
// symbols behind the n_pre preamble symbols of every RX antenna of every CPI.  Time samples
//   rx[cpi*cpi_stride + r*ant_stride + (n_pre + s)*(n + cp) + cp + i]
// -> DC-centred subcarrier vectors sym[((cpi*R + r)*n_sym + s)*n + k], the packed layout the chain kernels read with
// n_pre = 0.  Same butterflies as k_fft_rows: bit-identical to jrc_ofdm_demod symbol by symbol.
// ---------------------------------------------------------------------------
__global__ void k_ofdm_demod_batch(PortDev rx, int R, int n_pre, int n_sym, int cp, c32 *__restrict__ sym, int n, int log2n,
                                   long long rows, int rows_per_cta, const c32 *__restrict__ tw /* [n/2], forward */)
{
    extern __shared__ c32 sm[];
    const long long row0 = (long long)blockIdx.x * rows_per_cta;
    const int offset = (n + 1) / 2;
    const int tot = rows_per_cta * n;
    for (int e = threadIdx.x; e < tot; e += blockDim.x) {
        int lr = e / n, i = e % n;
        long long row = row0 + lr;
        c32 v = mk(0.f, 0.f);
        if (row < rows) {
            const int s = (int)(row % n_sym);
            const long long ca = row / n_sym;
            const int r = (int)(ca % R);
            const long long cpi = ca / R;
            v = rx.base[cpi * rx.cpi_stride + r * rx.ant_stride + (long long)(n_pre + s) * (n + cp) + cp + i];
        }
        unsigned rev = (log2n == 0) ? 0u : (__brev((unsigned)i) >> (32 - log2n));
        sm[lr * n + rev] = v;
    }
    __syncthreads();
    radix2_rows(sm, n, rows_per_cta, tw);
    for (int e = threadIdx.x; e < tot; e += blockDim.x) {
        int lr = e / n, i = e % n;
        long long row = row0 + lr;
        if (row < rows) sym[row * n + i] = sm[lr * n + (i + offset) % n];      // fft_vxx shift=True: output halves swapped
    }
}

// The 64-subcarrier case of k_ofdm_demod_batch, one WARP per symbol, no shared memory: lane l holds positions l and
// l + 32 of the bit-reversed working array, i.e. the consecutive input samples 2*brev5(l) and 2*brev5(l) + 1 (one
// 16-byte load).  A butterfly of the stages with half < 32 pairs lane l with lane l ^ half: the upper lane forms v*w,
// the two exchange (u or v*w) and each forms its own output with the oracle's operations, in the oracle's order
// (radix2_rows: u + v*w, u - v*w; w = tw[j * 64 / len]); the last stage is local to the lane.  Bit-identical to the
// generic kernel.  Rows must start 16-byte aligned (even strides, even cp).
__global__ void __launch_bounds__(256) k_ofdm_demod64(PortDev rx, int R, int n_pre, int n_sym, int cp, c32 *__restrict__ sym,
                                                      long long rows, const c32 *__restrict__ tw /* [32], forward */)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int src = (int)(__brev((unsigned)lane) >> 27) * 2;
    for (long long row = warp0; row < rows; row += nwarps) {
        const int s = (int)(row % n_sym);
        const long long ca = row / n_sym;
        const int r = (int)(ca % R);
        const long long cpi = ca / R;
        const c32 *x = rx.base + cpi * rx.cpi_stride + r * rx.ant_stride + (long long)(n_pre + s) * (64 + cp) + cp;
        const float4 ld = *reinterpret_cast<const float4 *>(x + src);
        c32 a = mk(ld.x, ld.y), b = mk(ld.z, ld.w);          // positions lane, lane + 32
#pragma unroll
        for (int half = 1; half < 32; half <<= 1) {
            const bool upper = (lane & half) != 0;
            const c32 w = tw[(lane & (half - 1)) * (32 / half)];
            const c32 ta = upper ? cmul_exact(a, w) : a, tb = upper ? cmul_exact(b, w) : b;
            c32 oa, ob;
            oa.x = __shfl_xor_sync(0xffffffffu, ta.x, half); oa.y = __shfl_xor_sync(0xffffffffu, ta.y, half);
            ob.x = __shfl_xor_sync(0xffffffffu, tb.x, half); ob.y = __shfl_xor_sync(0xffffffffu, tb.y, half);
            a = upper ? csub_exact(oa, ta) : cadd_exact(ta, oa);
            b = upper ? csub_exact(ob, tb) : cadd_exact(tb, ob);
        }
        const c32 v = cmul_exact(b, tw[lane]);
        const c32 lo = cadd_exact(a, v), hi = csub_exact(a, v);   // positions lane, lane + 32
        sym[row * 64 + lane + 32] = lo;                           // fft_vxx shift=True: output halves swapped
        sym[row * 64 + lane] = hi;
    }
}

// ---------------------------------------------------------------------------
// matrix_transpose  (lib/matrix_transpose_impl.cc:97-104), batched over mats:
// in [mat][K][L] -> out [mat][L][W], W = output_len*interp, out[l][k] = in[k][l] for
// k < K, zero elsewhere.  32x32 shared-memory tiles keep both sides coalesced.
// ---------------------------------------------------------------------------
__global__ void k_transpose_pad(const c32 *__restrict__ in, c32 *__restrict__ out, int K, int L, int W)
{
    __shared__ c32 tile[32][33];
    const long long mat = blockIdx.z;
    const c32 *src = in + mat * (long long)K * L;
    c32 *dst = out + mat * (long long)L * W;
    const int l0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        int k = k0 + dy, l = l0 + threadIdx.x;
        tile[dy][threadIdx.x] = (k < K && l < L) ? src[(long long)k * L + l] : mk(0.f, 0.f);
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        int l = l0 + dy, k = k0 + threadIdx.x;
        if (l < L && k < W) dst[(long long)l * W + k] = tile[threadIdx.x][dy];
    }
}

// blocks_complex_to_mag_squared: VOLK generic kernel, re*re + im*im, no contraction
__global__ void k_mag_squared(const c32 *__restrict__ in, float *__restrict__ out, long long n)
{
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        c32 z = in[e];
        out[e] = __fadd_rn(__fmul_rn(z.x, z.x), __fmul_rn(z.y, z.y));
    }
}

// ---------------------------------------------------------------------------
// range_angle_estimator  (lib/range_angle_estimator_impl.cc:137-227)
// pass 1: arg-max of (float)pow(abs(z),2), first maximum in row-major order wins
// (strict '>' at :144).  keys[mat] must be zeroed first.
// ---------------------------------------------------------------------------
__global__ void k_est_argmax(const c32 *__restrict__ map, long long per_mat, unsigned long long *keys)
{
    const long long mat = blockIdx.y;
    const c32 *m = map + mat * per_mat;
    unsigned long long best = 0ull;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < per_mat;
         e += (long long)gridDim.x * blockDim.x) {
        float p = (float)ref_pow_abs2(m[e]);
        if (p == p) {   // NaN never wins (curr_power > peak_power is false)
            unsigned long long key = pack_key(p, (unsigned)e);
            best = key > best ? key : best;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    __shared__ unsigned long long wbest[32];
    if ((threadIdx.x & 31) == 0) wbest[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        best = threadIdx.x < (blockDim.x >> 5) ? wbest[threadIdx.x] : 0ull;
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if (threadIdx.x == 0 && best) atomicMax(keys + mat, best);
    }
}

struct EstParams {
    const float *angle_bins;   // device [n_angle]
    int n_angle, n_range;
    int discard_range_idx;     // int(noise_discard_range_m / (range_bins[1]-range_bins[0])), host-computed (:189)
    float noise_discard_angle_deg;
    float snr_threshold, power_threshold;
};

struct DetDev {   // == jrc_det
    int range_idx, angle_idx; float peak_power, noise_power, snr_db; int n_noise; unsigned flags; int cpi;
};

// noise-window geometry (:152-201), shared by the staged and the fused kernels
struct NoiseWin { int start_r, end_r, start_a, end_a; };
__device__ __forceinline__ NoiseWin noise_window(const EstParams &P, int peak_r, int peak_a)
{
    float angle_val = P.angle_bins[peak_a];
    float angle_null = __fadd_rn(angle_val, 90.f);                       // :155
    if (angle_null >= 90.f) angle_null = __fsub_rn(angle_null, 180.f);   // :157-160
    int lo = 0, hi = P.n_angle;                                          // std::lower_bound :163
    while (lo < hi) { int mid = lo + ((hi - lo) >> 1); if (P.angle_bins[mid] < angle_null) lo = mid + 1; else hi = mid; }
    int idx;
    if (lo == 0) idx = 0;                                                // :172
    else if (lo == P.n_angle) idx = lo - 1;       // past-the-end read at :170 -> previous bin
    else {
        double a = P.angle_bins[lo - 1], b = P.angle_bins[lo];
        idx = (fabs((double)angle_null - a) < fabs((double)angle_null - b)) ? lo - 1 : lo;   // :175-180
    }
    if (idx == P.n_angle - 1) idx = P.n_angle - 2;                       // :184-187
    float da_f = __fdiv_rn(P.noise_discard_angle_deg,
                           __fsub_rn(P.angle_bins[(idx + 1) % P.n_angle], P.angle_bins[idx]));   // :190
    int da = (int)da_f;
    if (da <= 0) da = 1;                                                 // :192-195
    NoiseWin w;
    w.start_r = peak_r + P.n_range / 2 - P.discard_range_idx;            // :197
    w.end_r   = peak_r + P.n_range / 2 + P.discard_range_idx;            // :198
    w.start_a = idx - da;                                                // :200
    w.end_a   = idx + da;                                                // :201
    return w;
}

// ---------------------------------------------------------------------------
// target_simulator (lib/target_simulator_impl.cc:326-379) building blocks.
// ---------------------------------------------------------------------------
// volk_32fc_x2_multiply_32fc (:346,:353): out[row][i] = a[arow][i] * b[brow][i], arow = row / a_div, brow = row
__global__ void k_cmul_rows(const c32 *__restrict__ a, int a_div, const c32 *__restrict__ b, c32 *__restrict__ out,
                            int n, int rows)
{
    const long long total = (long long)rows * n;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(e / n), i = (int)(e % n);
        out[e] = cmul_exact(a[(long long)(row / a_div) * n + i], b[e]);
    }
}

// FFT of arbitrary length (the packet lengths of the simulation are not powers of two): direct DFT with
// float64 products and sums in index order and a host-computed cos/sin table -- the arithmetic of the CPU
// restatement's stand-in for FFTW, result rounded to float32.  One thread per output bin, rows on grid.y.
__global__ void __launch_bounds__(128) k_dft_any(const c32 *__restrict__ in, c32 *__restrict__ out, int n,
                                                 const double2 *__restrict__ tab /* [n] e^{-+j2pi k/n} */)
{
    __shared__ c32 tile[128];
    const c32 *x = in + (long long)blockIdx.y * n;
    const int m = blockIdx.x * 128 + threadIdx.x;
    double sr = 0.0, si = 0.0;
    long long idx = 0;
    for (int k0 = 0; k0 < n; k0 += 128) {
        if (k0 + threadIdx.x < n) tile[threadIdx.x] = x[k0 + threadIdx.x];
        __syncthreads();
        const int cnt = n - k0 < 128 ? n - k0 : 128;
        if (m < n) {
            for (int kk = 0; kk < cnt; kk++) {
                const double2 w = tab[idx];
                const double re = (double)tile[kk].x, im = (double)tile[kk].y;
                sr = __dadd_rn(sr, __dsub_rn(__dmul_rn(re, w.x), __dmul_rn(im, w.y)));
                si = __dadd_rn(si, __dadd_rn(__dmul_rn(re, w.y), __dmul_rn(im, w.x)));
                idx += m;
                if (idx >= n) idx -= n;
            }
        }
        __syncthreads();
    }
    if (m < n) out[(long long)blockIdx.y * n + m] = mk((float)sr, (float)si);
}

// per RX antenna: fold the per-target results (:359-367; the reference's memcpy keeps the LAST target, the
// accumulate option sums them), optional random phase per target (:360-363), self coupling (:372-378)
__global__ void k_sim_combine(const c32 *__restrict__ res /* [n_rx][n_targets][n] */, const c32 *__restrict__ in,
                              const c32 *__restrict__ phase /* [n_targets] or null */, int n, int n_targets,
                              int accumulate, int self_coupling, float g, c32 *__restrict__ out /* [n_rx][n] */)
{
    const int l = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        c32 o = mk(0.f, 0.f);
        for (int k = 0; k < n_targets; k++) {
            c32 r = res[((long long)l * n_targets + k) * n + i];
            if (phase) r = cmul_exact(r, phase[k]);
            o = accumulate ? cadd_exact(o, r) : r;
        }
        if (self_coupling) o = mk(__fadd_rn(o.x, __fmul_rn(g, in[i].x)), __fadd_rn(o.y, __fmul_rn(g, in[i].y)));
        out[(long long)l * n + i] = o;
    }
}

// ---------------------------------------------------------------------------
// Batched scene synthesis: what target_simulator (lib/target_simulator_impl.cc:177,188,296-303) + the RX OFDM demodulator
// deliver to the radar block for point targets, evaluated directly in the frequency domain for a whole batch of CPIs:
//     Y[cpi][r][s][k] = sum_t X[t][s][k] * sum_j a_j exp(-j 2 pi tau_{j,t,r} (f_k + fc)),
//     tau = (2 R_j - d_{t,r} sin(az_j)) / c,   d_{t,r} = lambda + (t + T r) lambda / 2     (...radar_sim.grc:105-147)
// plus optional complex Gaussian noise from a counter-based generator.  The delay phase needs float64 (tau * f ~ 1e3
// cycles).  One thread per (cpi, r, k).  Mirrors mimo_ofdm_jrc.synth.rx_symbols (NumPy float64), which the tests compare
// it with; it feeds bench.py's configs[4] sweep, whose 128 GiB of RX symbols cannot be kept anywhere.
// ---------------------------------------------------------------------------
struct SceneParams {
    const c32 *tx;            // [T][S][N]
    const float *range_m, *az_deg, *amp;     // [n_cpi][J]
    int n_cpi, J, T, R, S, N;
    double samp_rate, center_freq;
    float noise_sigma;        // per real component; 0: none
    unsigned long long seed;
    c32 *rx;                  // [n_cpi][R][S][N]
};

__device__ __forceinline__ unsigned long long scene_mix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <int TMAX>
__global__ void __launch_bounds__(256) k_scene_synth(const SceneParams P)
{
    const double c_light = 3e8, lam = c_light / P.center_freq;
    const long long total = (long long)P.n_cpi * P.R * P.N;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(e % P.N), r = (int)((e / P.N) % P.R);
        const long long cpi = e / ((long long)P.N * P.R);
        const double fk = (double)(k - P.N / 2) * P.samp_rate / (double)P.N + P.center_freq;
        c32 hch[TMAX];
#pragma unroll
        for (int t = 0; t < TMAX; t++) hch[t] = mk(0.f, 0.f);
        for (int j = 0; j < P.J; j++) {
            const double rg = (double)P.range_m[cpi * P.J + j], a = (double)P.amp[cpi * P.J + j];
            const double sn = sinpi((double)P.az_deg[cpi * P.J + j] / 180.0);
#pragma unroll
            for (int t = 0; t < TMAX; t++) {
                if (t < P.T) {
                    const double d = lam + (double)(t + P.T * r) * lam * 0.5;
                    const double cyc = (2.0 * rg - d * sn) / c_light * fk;
                    const double fr = cyc - floor(cyc);
                    double s, c;
                    sincospi(-2.0 * fr, &s, &c);
                    hch[t].x += (float)(a * c);
                    hch[t].y += (float)(a * s);
                }
            }
        }
        for (int s = 0; s < P.S; s++) {
            float yr = 0.f, yi = 0.f;
#pragma unroll
            for (int t = 0; t < TMAX; t++) {
                if (t < P.T) {
                    const c32 x = P.tx[((long long)t * P.S + s) * P.N + k];
                    yr += x.x * hch[t].x - x.y * hch[t].y;
                    yi += x.x * hch[t].y + x.y * hch[t].x;
                }
            }
            const long long o = ((cpi * P.R + r) * P.S + s) * P.N + k;
            if (P.noise_sigma > 0.f) {
                const unsigned long long hsh = scene_mix64(P.seed ^ (unsigned long long)o * 0xD1342543DE82EF95ull);
                const float u1 = ((float)(unsigned)(hsh >> 40) + 1.0f) * (1.0f / 16777217.0f);
                const float u2 = (float)(unsigned)((hsh >> 8) & 0xFFFFFFu) * (1.0f / 16777216.0f);
                const float rad = P.noise_sigma * sqrtf(-2.0f * logf(u1));
                float sn2, cs2;
                sincospif(2.0f * u2, &sn2, &cs2);
                yr += rad * cs2;
                yi += rad * sn2;
            }
            P.rx[o] = mk(yr, yi);
        }
    }
}

// blocks_nlog10_ff (...radar_sim.grc:725-745, in front of gui_heatmap_plot): n*log10(max(x, 1e-18)) + k
__global__ void k_nlog10(const float *__restrict__ in, float *__restrict__ out, long long cnt, float n, float k)
{
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < cnt;
         e += (long long)gridDim.x * blockDim.x)
        out[e] = __fadd_rn(__fmul_rn(n, log10f(fmaxf(in[e], 1e-18f))), k);
}

// Per-angle-bin tables for the fused kernel's estimator (the column window only depends on the peak's
// angle bin): win[m] = (start_a, end_a) of noise_window(., m) and the closed-form column sums
//   g[m][d] = sum_{c in window} w^{d (c + Na/2)},  w = e^{-j2pi/Na},  d = 1..7;   g[m][0] = ncols / 2,
// so that the window power is 2 Re sum_d g[d] A[d] with the lag autocorrelations A[d] of the 8 channels.
__global__ void k_est_tables(EstParams P, int2 *__restrict__ win, double2 *__restrict__ g)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= P.n_angle) return;
    const NoiseWin w = noise_window(P, 0, m);
    win[m] = make_int2(w.start_a, w.end_a);
    const int ncols = w.end_a - w.start_a, NA = P.n_angle;
    g[m * 8] = make_double2(0.5 * (double)ncols, 0.0);
    const long long m0 = w.start_a + NA / 2;
    for (int d = 1; d < 8; d++) {
        double s0, c0, s1, c1, s2, c2;
        sincospi(-2.0 * (double)(((d * m0) % NA + NA) % NA) / (double)NA, &s0, &c0);
        sincospi(-2.0 * (double)(((long long)d * ncols % NA + NA) % NA) / (double)NA, &s1, &c1);
        sincospi(-2.0 * (double)d / (double)NA, &s2, &c2);
        const double nr = 1.0 - c1, ni = -s1, dr = 1.0 - c2, di = -s2;
        const double den = dr * dr + di * di;
        const double qr = (nr * dr + ni * di) / den, qi = (ni * dr - nr * di) / den;
        g[m * 8 + d] = make_double2(c0 * qr - s0 * qi, c0 * qi + s0 * qr);
    }
}

// pass 2: one CTA per map.  The window powers are evaluated in parallel, but the
// float accumulation runs in the reference's order on one thread (:211-221) so the
// noise power is bit-identical.  snr/flags are finalised here (snr_db_of: log10 in double, rounded once);
// host-side callers that need the host libm's bit pattern recompute them.
__global__ void k_est_finalize(const c32 *__restrict__ map, long long per_mat, int n_inputs, int vlen,
                               unsigned long long *keys, EstParams P,
                               DetDev *__restrict__ dets, int cpi0, int reset_keys)   // reset_keys: leave keys[] zeroed for the next frame
{
    const long long mat = blockIdx.x;
    const c32 *m = map + mat * per_mat;
    __shared__ float chunk[2][1024];
    __shared__ NoiseWin win;
    __shared__ int s_peak_r, s_peak_a;
    __shared__ float s_noise;
    unsigned long long key = keys[mat];
    if (threadIdx.x == 0) {
        DetDev d;
        d.cpi = cpi0 + (int)mat; d.flags = 0; d.n_noise = 0;
        if (key == 0ull) {   // empty / all-NaN map
            d.range_idx = -1; d.angle_idx = -1; d.peak_power = -1.f;
            d.noise_power = __int_as_float(0x7fc00000); d.snr_db = d.noise_power;
            dets[mat] = d;
            s_peak_r = -1;
        } else {
            unsigned lin = 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull);
            s_peak_r = (int)(lin / (unsigned)vlen); s_peak_a = (int)(lin % (unsigned)vlen);
            win = noise_window(P, s_peak_r, s_peak_a);
        }
        s_noise = 0.f;
    }
    __syncthreads();
    if (s_peak_r < 0) return;
    const int ncols = win.end_a - win.start_a;
    const int nrows = win.end_r - win.start_r;
    const long long total = (ncols > 0 && nrows > 0) ? (long long)nrows * ncols : 0;
    // 1024 cells at a time: warp 0 runs the chain over one block of values while the other warps evaluate the next one
    auto eval_chunk = [&](long long base, float *dst, int t0, int nt) {
        const int cnt = (int)((total - base) < 1024 ? (total - base) : 1024);
        for (int j = t0; j < cnt; j += nt) {
            long long s = base + j;
            int ir = win.start_r + (int)(s / ncols), ia = win.start_a + (int)(s % ncols);
            int r_idx = ((ir % n_inputs) + n_inputs) % n_inputs;
            int a_idx = ((ia % vlen) + vlen) % vlen;
            dst[j] = ref_abs(m[a_idx + (long long)vlen * r_idx]);
        }
    };
    if (total > 0) eval_chunk(0, chunk[0], threadIdx.x, blockDim.x);
    __syncthreads();
    float acc = 0.f;
    int b = 0;
    for (long long base = 0; base < total; base += 1024, b ^= 1) {
        const int cnt = (int)((total - base) < 1024 ? (total - base) : 1024);
        if (threadIdx.x < 32) acc = seq_sum_sq_warp(acc, chunk[b], cnt, threadIdx.x);      // :217 float += double
        else if (base + 1024 < total) eval_chunk(base + 1024, chunk[b ^ 1], threadIdx.x - 32, blockDim.x - 32);
        __syncthreads();
    }
    if (threadIdx.x == 0) s_noise = acc;
    if (threadIdx.x == 0) {
        DetDev d;
        d.range_idx = s_peak_r; d.angle_idx = s_peak_a;
        d.peak_power = __uint_as_float((unsigned)(key >> 32));
        d.n_noise = (int)total;
        d.noise_power = __fdiv_rn(s_noise, (float)d.n_noise);            // :226
        d.snr_db = snr_db_of(d.peak_power, d.noise_power);                            // :227
        d.flags = (d.snr_db >= P.snr_threshold && d.peak_power >= P.power_threshold) ? 1u : 0u;   // :234
        d.cpi = cpi0 + (int)mat;
        dets[mat] = d;
        if (reset_keys) keys[mat] = 0ull;
    }
}


// ---------------------------------------------------------------------------
// fft_peak_detect  (lib/fft_peak_detect_impl.cc:88-95): first maximum of abs(in[p])
// over [protect, n-protect) among samples with pow(abs,2) > 10^(thr/10) (double).
// abs() is monotone with the key, so the arg-max runs on packed (abs, index) keys.
// ---------------------------------------------------------------------------
__global__ void k_peak1d_scan(const c32 *__restrict__ in, int n, int protect, double thr_lin,
                              unsigned long long *key)
{
    unsigned long long best = 0ull;
    for (long long p = (long long)protect + blockIdx.x * (long long)blockDim.x + threadIdx.x;
         p < (long long)n - protect; p += (long long)gridDim.x * blockDim.x) {
        float a = ref_abs(in[p]);
        if (a == a && (double)a * (double)a > thr_lin) {
            unsigned long long k = pack_key(a, (unsigned)p);
            best = k > best ? k : best;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if ((threadIdx.x & 31) == 0 && best) atomicMax(key, best);
}

struct Peak1dDev { int k; float freq, phase, mag; c32 z; };
__global__ void k_peak1d_finalize(const c32 *__restrict__ in, const unsigned long long *key, Peak1dDev *out)
{
    unsigned long long k = *key;
    Peak1dDev o; o.k = -1; o.freq = 0.f; o.phase = 0.f; o.mag = 0.f; o.z = mk(0.f, 0.f);
    if (k) {
        o.k = (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull));
        o.mag = __uint_as_float((unsigned)(k >> 32));
        o.z = in[o.k];
        o.phase = atan2f(o.z.y, o.z.x);   // host callers recompute with libm for bit parity
    }
    *out = o;
}

// ---------------------------------------------------------------------------
// zero_pad  (lib/zero_pad_impl.cc:76-93): copy + N(0,1e-2) complex noise pads.
// Counter-based generator (splitmix64 of seed^index, Box-Muller): the reference
// seeds from std::random_device per call, so only the distribution is defined.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void k_zero_pad(const c32 *__restrict__ in, int n, unsigned pad_front, unsigned pad_tail,
                           unsigned long long seed, c32 *__restrict__ out)
{
    const long long total = (long long)n + pad_front + pad_tail;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        if (e >= pad_front && e < (long long)pad_front + n) { out[e] = in[e - pad_front]; continue; }
        unsigned long long h = mix64(seed ^ (unsigned long long)e * 0xD1342543DE82EF95ull);
        float u1 = ((float)(unsigned)(h >> 40) + 1.0f) * (1.0f / 16777217.0f);
        float u2 = (float)(unsigned)((h >> 8) & 0xFFFFFFu) * (1.0f / 16777216.0f);
        float rad = 1e-2f * sqrtf(-2.0f * logf(u1));
        float s, c;
        sincospif(2.0f * u2, &s, &c);
        out[e] = mk(rad * c, rad * s);
    }
}

}  // namespace jrc
