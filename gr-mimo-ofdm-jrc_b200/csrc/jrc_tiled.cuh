// jrc_tiled.cuh -- the radar chain for configurations without a k_fused64x8 specialisation
// (BASELINE configs[2]: 32 virtual channels, 256 subcarriers, 4096 x 256 map; configs[4]: 128
// virtual channels, 2048 subcarriers, 2048 x 128 map), as two tiled kernels around k_chan_est:
//
//   k_fft8_rows   range IFFT of every channel row      (fft_vcc #A, ...radar_sim.grc:940-962; the
//                 zero-pad of lib/mimo_ofdm_radar_impl.cc:312-315 is implied by n_in < N)
//   k_angle_mag   matrix_transpose + angle zero-pad    (lib/matrix_transpose_impl.cc:97-104)
//                 + fft_vcc #B with fftshift           (...radar_sim.grc:963-985)
//                 + complex_to_mag_squared             (...radar_sim.grc:637-652)
//                 + the arg-max scan of range_angle_estimator (lib/range_angle_estimator_impl.cc:137-151)
//                 as a per-CPI 64-bit key that k_map_finalize turns into the detection record.
//
// The transposed / angle-padded [Nr][Na] matrix, the complex map and the arg-max pass never touch
// HBM (the staged path moves ~11x the algorithmic bytes at configs[2]); HBM sees the symbols, the
// channel estimates, the range spectra [V][Nr] (written and read once) and the |.|^2 map.
//
// FFT: decimation-in-frequency with radix-8 passes in registers (8 points per thread per pass, the
// packed DFT-8 of jrc_common.cuh), remaining factor 4 or 2 in the last pass.  A row lives in shared
// memory between passes (padded: one float2 per 8), twiddles w_N^i come from one table per (N,
// direction).  When the zero-pad leaves at most N/8 non-zero inputs the first pass degenerates to
// "copy the input to 8 places with a twiddle" and is fed straight from HBM; the last pass hands its
// results to the caller in registers (frequency f of position p = dif_freq(p)), so neither the load nor
// the digit-reversing store goes through shared memory.
// Same float32 arithmetic class as the oracle's radix-2 FFT, not its rounding: the parity criterion is
// 1e-4 of the map peak (tests), detections bit-exact whenever the peak is not a rounding-level tie.
#pragma once
#include <type_traits>
#include "jrc_common.cuh"
#include "jrc_staged.cuh"
#include "jrc_exact.cuh"

namespace jrc {

__device__ __host__ __forceinline__ constexpr int fpad(int i) { return i + (i >> 3); }

// frequency whose result the DIF passes (radix 8, ..., 8, then 4 or 2) leave at position p
// (a permutation of bit fields: dif_freq(a | b) = dif_freq(a) | dif_freq(b) for disjoint a, b)
template <int LOG2N>
__device__ __host__ __forceinline__ constexpr int dif_freq(int p)
{
    int f = 0, shift = 0, rem = LOG2N;
    for (; rem >= 3; rem -= 3) { f |= ((p >> (rem - 3)) & 7) << shift; shift += 3; }
    if (rem > 0) f |= (p & ((1 << rem) - 1)) << shift;
    return f;
}

template <int LOG2N>
struct TiledGeom {
    static constexpr int N = 1 << LOG2N;
    static constexpr int TPR = N / 8;                          // threads per row
    static constexpr int THREADS = TPR > 256 ? TPR : 256;
    static constexpr int RPC = THREADS / TPR;                  // rows per CTA
    // row stride in shared memory: when lanes run over rows first (transposing loads), 16 consecutive lanes
    // (min(16, RPC) rows x 16/RPC columns) must hit 16 different 8-byte banks
    static constexpr int RS = ((fpad(N - 1) + 1 + 15) / 16) * 16 + (RPC >= 16 ? 1 : RPC >= 2 ? 16 / RPC : 0);
    static constexpr int NTW = (LOG2N + 2) / 3 - 1;            // passes that apply twiddles (all but the last)
    static constexpr size_t SMEM = (size_t)RPC * RS * sizeof(c32);      // one tile of rows
    static constexpr bool WARP_SYNC = TPR <= 32;
};

// fpad(base + m*s) - fpad(base) for base = (multiple of 8) + j, j < s, s a power of two: a compile-time
// constant (m*s never straddles a multiple of 8 together with j), so the 8 accesses of a pass are one base
// address plus immediates
__device__ __forceinline__ constexpr int fpad_step(int m, int s) { return m * s + ((m * s) >> 3); }

template <bool WARP_SYNC>
__device__ __forceinline__ void row_sync() { if (WARP_SYNC) __syncwarp(); else __syncthreads(); }

// Twiddles of one thread for all passes but the last, loaded once per kernel: pass i works on blocks of
// L = N / 8^i points, thread index j_i within the block, factors w_L^{j_i k} = w_N^{8^i j_i k}, k = 1..7.
template <int LOG2N>
struct DifTw {
    c32 w[TiledGeom<LOG2N>::NTW > 0 ? TiledGeom<LOG2N>::NTW : 1][7];
    // j0: index of this thread in the first pass (the caller's choice when that pass is pruned), t: index
    // within the row for the other passes
    // alt0: the first-pass factors carry the sign (-1)^j0 (input modulation that moves the output by N/2: fftshift)
    __device__ __forceinline__ void load(const c32 *__restrict__ tw, int j0, int t, bool alt0 = false)
    {
#pragma unroll
        for (int i = 0; i < TiledGeom<LOG2N>::NTW; i++) {
            const int log2L = LOG2N - 3 * i;
            const int j = i == 0 ? j0 : (t & ((1 << (log2L - 3)) - 1));
            const int q = j << (3 * i);
#pragma unroll
            for (int k = 1; k < 8; k++) {
                c32 v = __ldg(tw + q * k);
                if (i == 0 && alt0 && (j0 & 1)) v = mk(-v.x, -v.y);
                asm volatile("" : "+f"(v.x), "+f"(v.y));   // keep it in registers: ptxas would re-load it inside the tile loop
                w[i][k - 1] = v;
            }
        }
    }
    __device__ __forceinline__ c32 get(int i, int k) const { return w[i][k - 1]; }
};

// The same factors kept in shared memory (7 per distinct thread index and pass) instead of 14 registers per
// pass: k_angle_mag trades 14 broadcast LDS per pass for a third resident CTA per SM.
template <int LOG2N>
struct DifTwS {
    const c32 *p[TiledGeom<LOG2N>::NTW > 0 ? TiledGeom<LOG2N>::NTW : 1];
    static constexpr int table_entries()      // sum over passes of 7 * (distinct j of the pass)
    {
        int n = 0;
        for (int i = 0; i < TiledGeom<LOG2N>::NTW; i++) n += 7 << (LOG2N - 3 * i - 3);
        return n;
    }
    // fills the table (all threads of the CTA, followed by the caller's __syncthreads) and points at this thread's rows
    __device__ __forceinline__ void init(c32 *table, const c32 *__restrict__ tw, int j0, int t, int tid, int nthreads,
                                         bool alt0 = false)
    {
        int off = 0;
#pragma unroll
        for (int i = 0; i < TiledGeom<LOG2N>::NTW; i++) {
            const int nj = 1 << (LOG2N - 3 * i - 3);
            for (int e = tid; e < 7 * nj; e += nthreads) {
                const int j = e / 7, k = e % 7 + 1;
                c32 v = __ldg(tw + (j << (3 * i)) * k);
                if (i == 0 && alt0 && (j & 1)) v = mk(-v.x, -v.y);
                table[off + e] = v;
            }
            const int j = i == 0 ? j0 : (t & (nj - 1));
            p[i] = table + off + 7 * j;
            off += 7 * nj;
        }
    }
    __device__ __forceinline__ c32 get(int i, int k) const { return p[i][k - 1]; }
};

// First pass when only inputs j < N/8 can be non-zero: u[k] = x_j * w_N^{j k}, written to the 8 places
// the full pass would write.  (r, j) of this thread is the caller's choice (coalesced HBM reads).
template <int LOG2N, class TW>
__device__ __forceinline__ void dif_first_pruned(c32 *x, int j, c32 xj, const TW &T, bool neg0 = false)
{
    constexpr int s = 1 << (LOG2N - 3);
    c32 *xb = x + fpad(j);
    xb[0] = neg0 ? mk(-xj.x, -xj.y) : xj;
#pragma unroll
    for (int k = 1; k < 8; k++) xb[fpad_step(k, s)] = cmul_fma(xj, T.get(0, k));
}

// First pass in general, fed from HBM as well: u[m] = x[j + m N/8] (zero beyond n_in), DFT-8, twiddle, store.
template <int LOG2N, int DIR, class TW>
__device__ __forceinline__ void dif_first_full(c32 *x, int j, c32 (&u)[8], const TW &T, bool neg0 = false)
{
    constexpr int s = 1 << (LOG2N - 3);
    JRC_FFT8<DIR>(u);
    if (neg0) u[0] = mk(-u[0].x, -u[0].y);
#pragma unroll
    for (int k = 1; k < 8; k++) u[k] = cmul_fma(u[k], T.get(0, k));
    c32 *xb = x + fpad(j);
#pragma unroll
    for (int k = 0; k < 8; k++) xb[fpad_step(k, s)] = u[k];
}

// Passes of one row of N = 2^LOG2N points in shared memory (index i at x[fpad(i)]) by N/8 threads,
// t = this thread's index within the row.  DIR = -1 forward, +1 backward.  SKIP_FIRST: the first pass was
// done by dif_first_pruned / dif_first_full.  The last pass leaves 8 results per thread in out[]: out[c] belongs to
// position 8t + c, i.e. frequency dif_freq(8t + c).
// The caller must synchronise the row before this call; all threads of the CTA must call it together
// unless WARP_SYNC.
template <int LOG2N, int DIR, bool WARP_SYNC, bool SKIP_FIRST, class TW>
__device__ __forceinline__ void dif_passes(c32 *x, int t, const TW &T, c32 (&out)[8])
{
    int log2L = SKIP_FIRST ? LOG2N - 3 : LOG2N;
#pragma unroll
    for (int i = SKIP_FIRST ? 1 : 0; i < TiledGeom<LOG2N>::NTW; i++, log2L -= 3) {
        const int s = 1 << (log2L - 3);              // distance of the 8 inputs = L/8
        const int j = t & (s - 1), base = ((t >> (log2L - 3)) << log2L) + j;
        c32 *xb = x + fpad(base);
        c32 u[8];
#pragma unroll
        for (int m = 0; m < 8; m++) u[m] = xb[fpad_step(m, s)];
        JRC_FFT8<DIR>(u);
#pragma unroll
        for (int k = 1; k < 8; k++) u[k] = cmul_fma(u[k], T.get(i, k));
#pragma unroll
        for (int k = 0; k < 8; k++) xb[fpad_step(k, s)] = u[k];
        row_sync<WARP_SYNC>();
    }
    // last pass: 8 consecutive points 8t .. 8t+7
    {
        const c32 *xb = x + 9 * t;                    // fpad(8t + m) = 9t + m
#pragma unroll
        for (int m = 0; m < 8; m++) out[m] = xb[m];
    }
    if (log2L == 3) {
        JRC_FFT8<DIR>(out);
    } else if (log2L == 2) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const c32 a = out[4 * h], b = out[4 * h + 1], c = out[4 * h + 2], d = out[4 * h + 3];
            const c32 s0 = __fadd2_rn(a, c), s1 = __fadd2_rn(a, mk(-c.x, -c.y)), s2 = __fadd2_rn(b, d);
            const c32 s3 = __fadd2_rn(b, mk(-d.x, -d.y));
            const c32 r3 = DIR < 0 ? mk(s3.y, -s3.x) : mk(-s3.y, s3.x);    // -j / +j times (b - d)
            out[4 * h] = __fadd2_rn(s0, s2);
            out[4 * h + 2] = __fadd2_rn(s0, mk(-s2.x, -s2.y));
            out[4 * h + 1] = __fadd2_rn(s1, r3);
            out[4 * h + 3] = __fadd2_rn(s1, mk(-r3.x, -r3.y));
        }
    } else {
#pragma unroll
        for (int h = 0; h < 4; h++) {
            const c32 a = out[2 * h], b = out[2 * h + 1];
            out[2 * h] = __fadd2_rn(a, b);
            out[2 * h + 1] = __fadd2_rn(a, mk(-b.x, -b.y));
        }
    }
}

// The passes between the first and the last one for ROWS independent rows that all threads of the CTA work on (thread t
// of every row): the rows are walked inside each pass, so one barrier per pass serves all of them and a thread has ROWS
// independent load -> DFT -> store chains to overlap.  x0: first row, rs: row pitch.  Ends with a barrier.
template <int LOG2N, int DIR, int ROWS, class TW>
__device__ __forceinline__ void dif_mid_passes_rows(c32 *x0, int rs, int t, const TW &T)
{
    int log2L = LOG2N - 3;
#pragma unroll
    for (int i = 1; i < TiledGeom<LOG2N>::NTW; i++, log2L -= 3) {
        const int s = 1 << (log2L - 3);
        const int j = t & (s - 1), base = ((t >> (log2L - 3)) << log2L) + j;
#pragma unroll
        for (int r = 0; r < ROWS; r++) {
            c32 *xb = x0 + r * rs + fpad(base);
            c32 u[8];
#pragma unroll
            for (int m = 0; m < 8; m++) u[m] = xb[fpad_step(m, s)];
            JRC_FFT8<DIR>(u);
#pragma unroll
            for (int k = 1; k < 8; k++) u[k] = cmul_fma(u[k], T.get(i, k));
#pragma unroll
            for (int k = 0; k < 8; k++) xb[fpad_step(k, s)] = u[k];
        }
        __syncthreads();
    }
}

// The last pass of one row (8 consecutive points 8t .. 8t+7 -> frequencies dif_freq(8t + c)); no synchronisation.
template <int LOG2N, int DIR>
__device__ __forceinline__ void dif_last_pass(const c32 *x, int t, c32 (&out)[8])
{
    constexpr int NTW = TiledGeom<LOG2N>::NTW;
    constexpr int log2L = LOG2N - 3 * NTW;
    const c32 *xb = x + 9 * t;
#pragma unroll
    for (int m = 0; m < 8; m++) out[m] = xb[m];
    if (log2L == 3) {
        JRC_FFT8<DIR>(out);
    } else if (log2L == 2) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const c32 a = out[4 * h], b = out[4 * h + 1], c = out[4 * h + 2], d = out[4 * h + 3];
            const c32 s0 = __fadd2_rn(a, c), s1 = __fadd2_rn(a, mk(-c.x, -c.y)), s2 = __fadd2_rn(b, d);
            const c32 s3 = __fadd2_rn(b, mk(-d.x, -d.y));
            const c32 r3 = DIR < 0 ? mk(s3.y, -s3.x) : mk(-s3.y, s3.x);
            out[4 * h] = __fadd2_rn(s0, s2);
            out[4 * h + 2] = __fadd2_rn(s0, mk(-s2.x, -s2.y));
            out[4 * h + 1] = __fadd2_rn(s1, r3);
            out[4 * h + 3] = __fadd2_rn(s1, mk(-r3.x, -r3.y));
        }
    } else {
#pragma unroll
        for (int h = 0; h < 4; h++) {
            const c32 a = out[2 * h], b = out[2 * h + 1];
            out[2 * h] = __fadd2_rn(a, b);
            out[2 * h + 1] = __fadd2_rn(a, mk(-b.x, -b.y));
        }
    }
}

// ---------------------------------------------------------------------------
// range IFFT: row r reads n_in samples at in + r*in_stride (the rest of the N-point input is zero) and
// writes N samples at out + r*N, natural order.
// ---------------------------------------------------------------------------
template <int LOG2N, int DIR, bool PRUNED>
__global__ void __launch_bounds__(TiledGeom<LOG2N>::THREADS) k_fft8_rows(const c32 *__restrict__ in, long long in_stride, int n_in,
                                                                         c32 *__restrict__ out, long long rows,
                                                                         const c32 *__restrict__ tw)
{
    using Gm = TiledGeom<LOG2N>;
    constexpr int N = Gm::N, RPC = Gm::RPC, RS = Gm::RS, TPR = Gm::TPR;
    extern __shared__ __align__(16) unsigned char smem_raw_t[];
    c32 *sm = reinterpret_cast<c32 *>(smem_raw_t);
    const int tid = threadIdx.x;
    const int lr_t = tid / TPR, t = tid % TPR;
    c32 *xrow = sm + lr_t * RS;
    DifTw<LOG2N> T;
    T.load(tw, t, t);
    c32 xn = mk(0.f, 0.f);     // PRUNED: this thread's input of the next row, fetched one iteration ahead
    if (PRUNED) {
        const long long row = (long long)blockIdx.x * RPC + lr_t;
        if (t < n_in && row < rows) xn = in[row * in_stride + t];
    }
    for (long long row0 = (long long)blockIdx.x * RPC; row0 < rows; row0 += (long long)gridDim.x * RPC) {
        const long long row = row0 + lr_t;
        if (PRUNED) {      // n_in <= N/8: thread (row, j = t) takes its single non-zero input straight from HBM
            const c32 xj = xn;
            const long long nrow = row + (long long)gridDim.x * RPC;
            xn = mk(0.f, 0.f);
            if (t < n_in && nrow < rows) xn = in[nrow * in_stride + t];
            dif_first_pruned<LOG2N>(xrow, t, xj, T);
        } else {           // every thread reads its 8 inputs (stride N/8, coalesced across the row's threads) from HBM
            c32 u[8];
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const int i = t + m * TPR;
                u[m] = (i < n_in && row < rows) ? in[row * in_stride + i] : mk(0.f, 0.f);
            }
            dif_first_full<LOG2N, DIR>(xrow, t, u, T);
        }
        __syncthreads();
        c32 o[8];
        dif_passes<LOG2N, DIR, Gm::WARP_SYNC, true>(xrow, t, T, o);
        if (row < rows) {
            c32 *orow = out + row * N + dif_freq<LOG2N>(8 * t);      // + dif_freq(c): immediates
#pragma unroll
            for (int c = 0; c < 8; c++) orow[dif_freq<LOG2N>(c)] = o[c];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// angle stage: tile = RPC consecutive range bins of one CPI.
//   Y   [n_cpi][V][Nr]   range spectra
//   map [n_cpi][Nr][NA]  |.|^2, angle bin fastest, fftshifted
//   keys[n_cpi]          (map value bits << 32) | (0xFFFFFFFF - range bin): atomicMax, zeroed by the host
//   sec[n_cpi]           bits of the largest group maximum that is NOT the CPI's maximum (every key that loses an
//                        exchange, and every warp's runner-up): k_map_finalize marks the record when it lies within
//                        EPS_AMB of the maximum (jrc_exact.cuh).  Zeroed by the host.
// ---------------------------------------------------------------------------
// Twiddles: registers (2 CTAs per SM) when the first pass is pruned, shared memory (3 CTAs per SM) otherwise --
// measured on configs[2] (registers 0.234 ms vs shared 0.250 ms) and configs[4] (0.160 ms vs 0.134 ms).
template <int LOG2NA, bool PRUNED>
#ifndef JRC_ANGLE_CTAS
#define JRC_ANGLE_CTAS 2     // resident CTAs per SM asked of ptxas for the pruned form (A/B: 3, 4)
#endif
__global__ void __launch_bounds__(256, PRUNED ? JRC_ANGLE_CTAS : 3) k_angle_mag(const c32 *__restrict__ Y, int V, int Nr, int log2_tiles_per_cpi, int n_cpi,
                                                      float *__restrict__ map, unsigned long long *__restrict__ keys,
                                                      unsigned *__restrict__ sec, const c32 *__restrict__ tw)
{
    using Gm = TiledGeom<LOG2NA>;
    constexpr int NA = Gm::N, RPC = Gm::RPC, RS = Gm::RS, TPR = Gm::TPR;
    static_assert(Gm::THREADS == 256, "angle FFT length above 2048 is not supported");
    extern __shared__ __align__(16) unsigned char smem_raw_t[];
    c32 *sm = reinterpret_cast<c32 *>(smem_raw_t);          // two tile buffers: one CTA barrier per tile
    const int tid = threadIdx.x, lane = tid & 31;
    const int lr_t = tid / TPR, t = tid % TPR;
    // The output fftshift is an input modulation: channel p enters as (-1)^p y[p] (its sign folded into the first
    // pass's twiddles), so position 8t + c of the last pass holds angle bin dif_freq(8t) + dif_freq(c) directly.
    const int obin0 = dif_freq<LOG2NA>(8 * t);
    // every CTA takes a contiguous run of tiles, so the running maximum is folded into keys[] once per CPI
    const long long n_tiles = (long long)n_cpi << log2_tiles_per_cpi;
    const long long tile_begin = n_tiles * blockIdx.x / gridDim.x, tile_end = n_tiles * (blockIdx.x + 1) / gridDim.x;
    const int tpc_mask = (1 << log2_tiles_per_cpi) - 1;
    // thread (range bin r fastest, channel / first-pass index pp): coalesced reads, first pass on the fly
    const int pr = tid % RPC, pp = tid / RPC;
    const bool neg0 = pp & 1;
    typename std::conditional<PRUNED, DifTw<LOG2NA>, DifTwS<LOG2NA>>::type T;
    if constexpr (PRUNED) {
        T.load(tw, pp, t, true);
    } else {
        T.init(sm + 2 * RPC * RS, tw, pp, t, tid, 256, true);     // table behind the two tile buffers
        __syncthreads();
    }
    // input pointer of tile ft for this thread, advanced tile by tile (+RPC range bins, next CPI at the wrap)
    long long ft = tile_begin;
    const c32 *fp = Y + ((tile_begin >> log2_tiles_per_cpi) * V + pp) * Nr + ((int)tile_begin & tpc_mask) * RPC + pr;
    const long long cpi_jump = (long long)(V - 1) * Nr;
    auto fetch_next = [&]() {        // PRUNED: the single non-zero input of the next tile not fetched yet
        const c32 v = (ft < tile_end && pp < V) ? *fp : mk(0.f, 0.f);
        ft++;
        fp += RPC;
        if (((int)ft & tpc_mask) == 0) fp += cpi_jump;
        return v;
    };
    auto flush = [&](int cpi, float best, int best_row, float sec_t) {
        if (!keys) return;
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
            // most candidates lose against the running maximum: a plain read filters them before the (same-address,
            // hence serialised) atomic.  Whoever loses an exchange is folded into sec[] when it could still matter
            // (the running maximum only grows).
            const unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(keys + cpi);
            unsigned long long top = cur;
            float loser = -1.f;
            if (key > cur) {
                const unsigned long long old = atomicMax(keys + cpi, key);
                top = old > key ? old : key;
                const unsigned long long lo = old > key ? key : old;
                if (lo) loser = __uint_as_float((unsigned)(lo >> 32));
            } else {
                loser = __uint_as_float((unsigned)(key >> 32));
            }
            loser = fmaxf(loser, b2);
            if (loser >= __uint_as_float((unsigned)(top >> 32)) * (1.f - 2.f * EPS_AMB)) atomicMax(sec + cpi, __float_as_uint(loser));
        }
    };
    constexpr int PF = 2;                           // tiles fetched ahead (one is not enough: measured 0.31 -> 0.23 ms)
    c32 xq[PF];
#pragma unroll
    for (int i = 0; i < PF; i++) xq[i] = PRUNED ? fetch_next() : mk(0.f, 0.f);
    float best = -1.f, sec_t = -1.f;
    int best_row = 0, cur_cpi = (int)(tile_begin >> log2_tiles_per_cpi);
    int buf = 0;
    // map rows are contiguous across tiles and CPIs: row of (tile, lr_t) = tile * RPC + lr_t
    float *mp = map ? map + (tile_begin * RPC + lr_t) * NA + obin0 : nullptr;
    for (long long tile = tile_begin; tile < tile_end; tile++, buf ^= 1) {
        const int cpi = (int)(tile >> log2_tiles_per_cpi), n0 = ((int)tile & tpc_mask) * RPC;
        if (cpi != cur_cpi) {
            flush(cur_cpi, best, best_row, sec_t);
            best = sec_t = -1.f;
            cur_cpi = cpi;
        }
        c32 *rows = sm + buf * (RPC * RS);
        if (PRUNED) {
            const c32 xj = xq[0];
#pragma unroll
            for (int i = 0; i + 1 < PF; i++) xq[i] = xq[i + 1];
            xq[PF - 1] = fetch_next();
            dif_first_pruned<LOG2NA>(rows + pr * RS, pp, xj, T, neg0);
        } else {
            // transposing load: thread (range bin r fastest, j) reads channels j + m NA/8 -- per channel RPC consecutive
            // range bins, contiguous in HBM -- and runs the first pass on them (NA/8 is even: one sign for all eight)
            const c32 *Yc = fp;
            c32 u[8];
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const int p = pp + m * TPR;
                u[m] = p < V ? Yc[(long long)m * TPR * Nr] : mk(0.f, 0.f);   // angle zero-pad
            }
            ft++;
            fp += RPC;
            if (((int)ft & tpc_mask) == 0) fp += cpi_jump;
            dif_first_full<LOG2NA, -1>(rows + pr * RS, pp, u, T, neg0);
        }
        __syncthreads();
        c32 o[8];
        dif_passes<LOG2NA, -1, Gm::WARP_SYNC, true>(rows + lr_t * RS, t, T, o);
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; c++) {       // re*re + im*im, each product rounded (VOLK's generic kernel)
            const c32 sq = __fmul2_rn(o[c], o[c]);
            v[c] = __fadd_rn(sq.x, sq.y);
        }
        if (map) {
#pragma unroll
            for (int c = 0; c < 8; c++) __stcs(mp + dif_freq<LOG2NA>(c), v[c]);
            mp += RPC * NA;
        }
        const float m8 = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
        sec_t = fmaxf(sec_t, fminf(m8, best));
        if (m8 > best) { best = m8; best_row = n0 + lr_t; }     // rows ascend within a CPI: the first maximum stays
    }
    if (tile_begin < tile_end) flush(cur_cpi, best, best_row, sec_t);
}

// ---------------------------------------------------------------------------
// Detection record from a per-CPI arg-max key and the |.|^2 map (used behind the fused kernels when the
// map is written anyway): key = (map value bits << 32) | (0xFFFFFFFF - range bin n), the earliest row
// among equal values.  One CTA per CPI: first bin of row n that holds the value, noise window
// (lib/range_angle_estimator_impl.cc:152-227) read back from the map, SNR gate (:234).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_map_finalize(const float *__restrict__ map, const unsigned long long *__restrict__ keys,
                                                      const unsigned *__restrict__ sec, int n_cpi, int NR, int NA, EstParams P,
                                                      DetDev *__restrict__ dets, int cpi0, FixCtl *fix_ctl, int *fix_list)
{
    extern __shared__ float s_abins[];     // angle_bins copy: the window geometry's binary search stays on chip
    __shared__ int s_istar[4], s_ncand[4];
    __shared__ double s_acc[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cpi = blockIdx.x;            // one CTA per CPI: the window reads are latency bound
    if (cpi >= n_cpi) return;
    // the key, the runner-up and the angle axis are independent loads: one round trip
    const unsigned long long key = __ldcg(keys + cpi);
    const unsigned sec_bits = __ldcg(sec + cpi);
    for (int i = tid; i < NA; i += blockDim.x) s_abins[i] = P.angle_bins[i];
    __syncthreads();
    EstParams est = P;
    est.angle_bins = s_abins;
    if (key == 0ull) {   // NaN-only input: nothing can win the strict '>' scan
        if (tid == 0) {
            DetDev d; d.range_idx = -1; d.angle_idx = -1; d.peak_power = -1.f;
            d.noise_power = __int_as_float(0x7fc00000); d.snr_db = d.noise_power;
            d.n_noise = 0; d.flags = 0; d.cpi = cpi0 + cpi;
            dets[cpi] = d;
        }
        return;
    }
    const float peak = __uint_as_float((unsigned)(key >> 32));
    const int nstar = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
    const float *map_c = map + (long long)cpi * NR * NA;
    const float thr_amb = __fmul_rn(peak, 1.f - EPS_AMB);
    int istar = 0x7fffffff, ncand = 0;
    for (int i = tid; i < NA; i += 128) {
        const float v = __ldcg(map_c + (long long)nstar * NA + i);
        if (v == peak && i < istar) istar = i;
        ncand += v >= thr_amb;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        istar = min(istar, __shfl_xor_sync(0xffffffffu, istar, o));
        ncand += __shfl_xor_sync(0xffffffffu, ncand, o);
    }
    if (lane == 0) { s_istar[warp] = istar; s_ncand[warp] = ncand; }
    __syncthreads();
    istar = min(min(s_istar[0], s_istar[1]), min(s_istar[2], s_istar[3]));
    ncand = s_ncand[0] + s_ncand[1] + s_ncand[2] + s_ncand[3];
    if (istar == 0x7fffffff) istar = 0;      // cannot happen: the key was built from this row
    const NoiseWin w = noise_window(est, nstar, istar);
    const int ncols = w.end_a - w.start_a, nrows = w.end_r - w.start_r;
    const int total = (ncols > 0 && nrows > 0) ? nrows * ncols : 0;
    double acc = 0.0;
    for (int j0 = tid; j0 < total; j0 += 128 * 16) {     // 16 independent loads in flight per lane
        float v[16];
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int j = j0 + 128 * q;
            v[q] = 0.f;
            if (j < total) {
                const int ir = w.start_r + j / ncols, ia = w.start_a + j % ncols;
                const int r_idx = ((ir % NR) + NR) % NR, a_idx = ((ia % NA) + NA) % NA;
                v[q] = __ldcg(map_c + (long long)r_idx * NA + a_idx);
            }
        }
#pragma unroll
        for (int q = 0; q < 16; q++) acc += (double)v[q];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_acc[warp] = acc;
    __syncthreads();
    if (tid == 0) {
        acc = (s_acc[0] + s_acc[1]) + (s_acc[2] + s_acc[3]);
        DetDev d;
        d.range_idx = nstar; d.angle_idx = istar; d.peak_power = peak; d.n_noise = total;
        d.noise_power = __fdiv_rn((float)acc, (float)total);
        d.snr_db = snr_db_fast(d.peak_power, d.noise_power);
        d.flags = (d.snr_db >= P.snr_threshold && d.peak_power >= P.power_threshold) ? DET_PASSED : 0u;
        // decisions that FFT rounding could turn are not taken here (jrc_exact.cuh)
        if (ncand > 1 || __uint_as_float(sec_bits) >= thr_amb) d.flags |= DET_PENDING | DET_AMB;
        if (gate_is_marginal(d.peak_power, d.noise_power, d.snr_db, total, P.snr_threshold, P.power_threshold))
            d.flags |= DET_PENDING | DET_GATE;
        d.cpi = cpi0 + cpi;
        dets[cpi] = d;
        if (d.flags & DET_PENDING) fix_push(fix_ctl, fix_list, cpi);
    }
}


}  // namespace jrc
