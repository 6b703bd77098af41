// jrc_slice.cuh -- BASELINE configs[2] (4 TX x 8 RX = 32 virtual channels, 256 subcarriers, range zero-pad 4096, angle
// zero-pad 256) and its relatives: range IFFT + matrix_transpose + angle FFT + |.|^2 + arg-max in ONE kernel whose only
// HBM traffic is the channel estimates in and the map out.
//
//   fft_vcc #A  (...radar_sim.grc:940-962; zero-pad of lib/mimo_ofdm_radar_impl.cc:312-315)
//   matrix_transpose + angle zero-pad (lib/matrix_transpose_impl.cc:97-104)
//   fft_vcc #B with fftshift (...radar_sim.grc:963-985), complex_to_mag_squared (:637-652)
//   arg-max scan of range_angle_estimator (lib/range_angle_estimator_impl.cc:137-151) as keys[] / sec[]
//
// How the [V][Nr] range spectra stay out of HBM (they are 1 MiB per CPI at configs[2]): the zero-padded range IFFT of
// Nr = N*IR points decimates exactly into IR independent N-point IFFTs,
//     y[p][q + IR*m] = sum_k ( H[p][k] * W_Nr^{k q} ) * W_N^{k m},        q = 0 .. IR-1  ("slice"),
// and a slice PAIR (q, q+1) of all V channels is 2*V*N complex values = 128 KiB: it fits one SM.  A CTA takes a
// (CPI, slice pair) unit, transforms the V channels into shared memory, and runs the angle stage for the 2*N range
// bins of the unit straight from there -- the transpose is an index.  The two slices of a pair are adjacent map
// rows (n, n+1), i.e. 2 KiB runs in HBM.
//
// Arithmetic: every transform handles TWO rows at once in split-complex form -- a float2 holds the real (or imaginary)
// parts of the same element of both rows -- so each packed FFMA2/FADD2/FMUL2 does useful work in both halves and every
// shared-memory access is 128 bits wide (about 113 instructions per map row of 256 bins against 195 for the
// one-row kernels of jrc_tiled.cuh).  Rows of a range transform = the two slices of one channel; rows of an angle
// transform = the two range bins (n, n+1).  One warp per row pair, __syncwarp between passes.
//   range: radix 8.8.4 decimation in frequency, in place in the slice-pair buffer.
//   angle: with only V <= 32 of the 256 inputs non-zero, bin i = b + 8a is a 32-point DFT of the inputs twiddled by w_256^{pb}:
//              M[b + 8a] = sum_{s<8} w_8^{s k2} [ w_32^{s k1} sum_{m<4} w_4^{m k1} ( (-1)^p w_256^{p b} x_p ) ],   p = s + 8m, a = k1 + 4 k2.
//          Lane (bq, s) loads its 4 inputs once, and for b = bq and bq + 4 twiddles them, runs the 4-point DFT in registers and
//          applies w_32^{s k1}; ONE exchange through shared memory (8 float4 out, 8 in, conflict-free both ways); lane (b, k1)
//          finishes with one 8-point DFT over s and owns bins lane + 32 k2 -- every store instruction of the warp is one
//          full 128-byte line.  The exchange is what bounds a shared-memory FFT (the LSU moves 128 B per clock): this form
//          moves 96 wavefronts per row pair (16 in, 64 exchange, 16 out) where three radix-8/8/4 passes move 170, about the time
//          the packed FP32 work needs (~80 clocks).
// Same float32 arithmetic class as the oracle's radix-2 FFTs, not their rounding (map criterion 1e-4 of the peak).
#pragma once
#include "jrc_common.cuh"
#include "jrc_tiled.cuh"

namespace jrc {

struct SliceParams {
    const c32 *H;                 // [n_cpi][V][256] channel estimates
    int V, IR, n_cpi;
    float *map;                   // [n_cpi][256*IR][256]
    unsigned long long *keys;     // [n_cpi] or nullptr (see k_angle_mag)
    unsigned *sec;
    const c32 *tw_range;          // [256*IR] W_Nr^i = e^{+j 2 pi i / Nr}
    const c32 *tw256;             // [256]    w_256^i = e^{-j 2 pi i / 256}
};

struct SliceGeom {
    static constexpr int N = 256, NA = 256, THREADS = 512, WARPS = 16;
    static constexpr int ROW = 288;                       // fpad(255) + 1 float4 of one transform
    static constexpr int PITCH = 289;                     // row pitch of the slice-pair spectra: == 1 mod 8 -> the 32 lanes of the
                                                          // angle stage's transposing read hit 8 x 4 distinct banks per quarter warp
    static constexpr size_t SMEM = (size_t)(32 * PITCH + WARPS * ROW + N) * sizeof(float4);
};

// (re, im) * (wr, wi), two rows at once
__device__ __forceinline__ void cmul2(float2 &re, float2 &im, float2 wr, float2 wi)
{
    const float2 r = __ffma2_rn(mk(-im.x, -im.y), wi, __fmul2_rn(re, wr));
    const float2 i = __ffma2_rn(im, wr, __fmul2_rn(re, wi));
    re = r; im = i;
}
// (re, im) * conj(wr, wi)
__device__ __forceinline__ void cmul2c(float2 &re, float2 &im, float2 wr, float2 wi)
{
    const float2 r = __ffma2_rn(im, wi, __fmul2_rn(re, wr));
    const float2 i = __ffma2_rn(mk(-re.x, -re.y), wi, __fmul2_rn(im, wr));
    re = r; im = i;
}

// 4-point DFT of two rows, natural order in and out
template <int DIR>
__device__ __forceinline__ void fft4s(float2 &r0, float2 &r1, float2 &r2, float2 &r3, float2 &i0, float2 &i1, float2 &i2, float2 &i3)
{
#define JRC_ADD(a, b) __fadd2_rn(a, b)
#define JRC_SUB(a, b) __fadd2_rn(a, mk(-(b).x, -(b).y))
    const float2 s0r = JRC_ADD(r0, r2), s0i = JRC_ADD(i0, i2), s1r = JRC_SUB(r0, r2), s1i = JRC_SUB(i0, i2);
    const float2 s2r = JRC_ADD(r1, r3), s2i = JRC_ADD(i1, i3), s3r = JRC_SUB(r1, r3), s3i = JRC_SUB(i1, i3);
    r0 = JRC_ADD(s0r, s2r); i0 = JRC_ADD(s0i, s2i);
    r2 = JRC_SUB(s0r, s2r); i2 = JRC_SUB(s0i, s2i);
    if (DIR < 0) {   // -j (b - d) = (s3i, -s3r)
        r1 = JRC_ADD(s1r, s3i); i1 = JRC_SUB(s1i, s3r);
        r3 = JRC_SUB(s1r, s3i); i3 = JRC_ADD(s1i, s3r);
    } else {         // +j (b - d) = (-s3i, s3r)
        r1 = JRC_SUB(s1r, s3i); i1 = JRC_ADD(s1i, s3r);
        r3 = JRC_ADD(s1r, s3i); i3 = JRC_SUB(s1i, s3r);
    }
#undef JRC_ADD
#undef JRC_SUB
}

// Passes 2 and 3 of a 256-point two-row transform that sits in X (element i at X[fpad(i)]) after its first pass;
// lane t of the owning warp.  Leaves the 8 results of positions 8t .. 8t+7 in re/im (frequency dif_freq<8>(8t + c)).
// tw2r/tw2i: w_32^{(t & 3) k}, k = 1..7, forward sign (conjugated here when DIR > 0).
template <int DIR>
__device__ __forceinline__ void dif2_passes23(float4 *X, int t, const float2 (&tw2r)[7], const float2 (&tw2i)[7],
                                              float2 (&re)[8], float2 (&im)[8])
{
    {
        float4 *xb = X + fpad(((t >> 2) << 5) + (t & 3));
#pragma unroll
        for (int m = 0; m < 8; m++) { const float4 v = xb[fpad_step(m, 4)]; re[m] = mk(v.x, v.y); im[m] = mk(v.z, v.w); }
        fft8s<DIR>(re, im);
#pragma unroll
        for (int k = 1; k < 8; k++) { if (DIR < 0) cmul2(re[k], im[k], tw2r[k - 1], tw2i[k - 1]); else cmul2c(re[k], im[k], tw2r[k - 1], tw2i[k - 1]); }
#pragma unroll
        for (int k = 0; k < 8; k++) xb[fpad_step(k, 4)] = make_float4(re[k].x, re[k].y, im[k].x, im[k].y);
    }
    __syncwarp();
    {
        const float4 *xb = X + 9 * t;
#pragma unroll
        for (int m = 0; m < 8; m++) { const float4 v = xb[m]; re[m] = mk(v.x, v.y); im[m] = mk(v.z, v.w); }
        fft4s<DIR>(re[0], re[1], re[2], re[3], im[0], im[1], im[2], im[3]);
        fft4s<DIR>(re[4], re[5], re[6], re[7], im[4], im[5], im[6], im[7]);
    }
}

__global__ void __launch_bounds__(SliceGeom::THREADS, 1) k_slice256(const SliceParams P)
{
    using Gm = SliceGeom;
    constexpr int N = Gm::N, NA = Gm::NA, PITCH = Gm::PITCH, ROW = Gm::ROW;
    extern __shared__ __align__(16) unsigned char smem_slice[];
    float4 *Ys = reinterpret_cast<float4 *>(smem_slice);          // [32][PITCH]: (Re y_q, Re y_q+1, Im y_q, Im y_q+1) at fpad(position)
    float4 *Xw = Ys + 32 * PITCH;                                 // [WARPS][ROW] angle work rows
    float4 *Sl = Xw + Gm::WARPS * ROW;                            // [256] slice twiddles (W^{kq}, W^{k(q+1)}) as (re, re, im, im)
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int V = P.V, IR = P.IR, Nr = N * IR;

    float4 *X = Xw + warp * ROW;
    // angle stage: before the exchange lane = (bq, s): inputs p = s + 8m, angle sub-bins b = bq and bq + 4;
    //              after it     lane = (b, k1) = (lane & 7, lane >> 3)
    const int abq = lane >> 3, as8 = lane & 7;
    const int xw0 = ((abq << 2) << 3) + ((as8 + abq) & 7), xw1 = (((abq + 4) << 2) << 3) + ((as8 + abq + 4) & 7);   // U[b][k1][(s + b) & 7]
    const int xr0 = ((((lane & 7) << 2) | (lane >> 3)) << 3);

    const int units_per_cpi = IR >> 1;
    const long long n_units = (long long)P.n_cpi * units_per_cpi;
    for (long long unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int cpi = (int)(unit / units_per_cpi), q0 = 2 * (int)(unit % units_per_cpi);
        // ---- slice twiddles of this unit ----
        for (int k = tid; k < N; k += Gm::THREADS) {
            const c32 a = __ldg(P.tw_range + (((long long)k * q0) % Nr)), b = __ldg(P.tw_range + (((long long)k * (q0 + 1)) % Nr));
            Sl[k] = make_float4(a.x, b.x, a.y, b.y);
        }
        __syncthreads();      // (also: the previous unit's angle stage has finished reading Ys)
        // ---- range stage: channel p -> Ys[p], two slices at once ----
        float2 t1r[7], t1i[7], t2r[7], t2i[7];      // pass 1: w_256^{lane k}, pass 2: w_32^{(lane & 3) k}
#pragma unroll
        for (int k = 1; k < 8; k++) {
            const c32 a = __ldg(P.tw256 + ((lane * k) & 255)), b = __ldg(P.tw256 + ((8 * (lane & 3) * k) & 255));
            t1r[k - 1] = mk(a.x, a.x); t1i[k - 1] = mk(a.y, a.y);
            t2r[k - 1] = mk(b.x, b.x); t2i[k - 1] = mk(b.y, b.y);
        }
        for (int p = warp; p < V; p += Gm::WARPS) {
            float4 *Y = Ys + p * PITCH;
            const c32 *Hp = P.H + ((long long)cpi * V + p) * N + lane;
            float2 re[8], im[8];
            c32 h[8];
#pragma unroll
            for (int m = 0; m < 8; m++) h[m] = __ldcg(Hp + 32 * m);
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const float4 s = Sl[lane + 32 * m];
                const float2 hr = mk(h[m].x, h[m].x), hi = mk(h[m].y, h[m].y), wr = mk(s.x, s.y), wi = mk(s.z, s.w);
                re[m] = __ffma2_rn(mk(-hi.x, -hi.y), wi, __fmul2_rn(hr, wr));
                im[m] = __ffma2_rn(hi, wr, __fmul2_rn(hr, wi));
            }
            fft8s<1>(re, im);
#pragma unroll
            for (int k = 1; k < 8; k++) cmul2c(re[k], im[k], t1r[k - 1], t1i[k - 1]);
            {
                float4 *yb = Y + fpad(lane);
#pragma unroll
                for (int k = 0; k < 8; k++) yb[fpad_step(k, 32)] = make_float4(re[k].x, re[k].y, im[k].x, im[k].y);
            }
            __syncwarp();
            dif2_passes23<1>(Y, lane, t2r, t2i, re, im);
            __syncwarp();
            {
                float4 *yb = Y + 9 * lane;
#pragma unroll
                for (int c = 0; c < 8; c++) yb[c] = make_float4(re[c].x, re[c].y, im[c].x, im[c].y);
            }
        }
        for (int p = V + warp; p < 32; p += Gm::WARPS) {          // angle zero-pad within the 32 lanes
            float4 *yb = Ys + p * PITCH + 9 * lane;
#pragma unroll
            for (int c = 0; c < 8; c++) yb[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
        // ---- angle stage: range positions pos -> map rows (n, n + 1), n = q0 + IR * dif_freq(pos) ----
        float2 a1r[2][4], a1i[2][4], w2r[3], w2i[3];      // (-1)^p w_256^{p b} (the fftshift is the sign), w_32^{s k1}
#pragma unroll
        for (int bi = 0; bi < 2; bi++)
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const c32 a = __ldg(P.tw256 + (((as8 + 8 * m) * (abq + 4 * bi)) & 255));
                const float sg = (as8 & 1) ? -1.f : 1.f;
                a1r[bi][m] = mk(sg * a.x, sg * a.x); a1i[bi][m] = mk(sg * a.y, sg * a.y);
            }
#pragma unroll
        for (int k = 1; k < 4; k++) {
            const c32 a = __ldg(P.tw256 + ((8 * as8 * k) & 255));
            w2r[k - 1] = mk(a.x, a.x); w2i[k - 1] = mk(a.y, a.y);
        }
        float best = -1.f, sec_t = -1.f;
        int best_row = 0;
        for (int pos = warp; pos < N; pos += Gm::WARPS) {
            const int m = (pos >> 5) | (((pos >> 2) & 7) << 3) | ((pos & 3) << 6);       // dif_freq<8>(pos)
            const int n = q0 + IR * m;
            float2 re[8], im[8];
            {
                const float4 *yp = Ys + as8 * PITCH + fpad(pos);
                float4 x[4];
#pragma unroll
                for (int mm = 0; mm < 4; mm++) x[mm] = yp[8 * mm * PITCH];      // 8 distinct rows per quarter warp: one wavefront each
#pragma unroll
                for (int bi = 0; bi < 2; bi++) {
                    float2 *r4 = re + 4 * bi, *i4 = im + 4 * bi;
#pragma unroll
                    for (int mm = 0; mm < 4; mm++) {
                        r4[mm] = mk(x[mm].x, x[mm].y); i4[mm] = mk(x[mm].z, x[mm].w);
                        cmul2(r4[mm], i4[mm], a1r[bi][mm], a1i[bi][mm]);
                    }
                    fft4s<-1>(r4[0], r4[1], r4[2], r4[3], i4[0], i4[1], i4[2], i4[3]);
#pragma unroll
                    for (int k = 1; k < 4; k++) cmul2(r4[k], i4[k], w2r[k - 1], w2i[k - 1]);
                    float4 *ub = X + (bi ? xw1 : xw0);
#pragma unroll
                    for (int k = 0; k < 4; k++) ub[k << 3] = make_float4(r4[k].x, r4[k].y, i4[k].x, i4[k].y);
                }
            }
            __syncwarp();
            {
                const float4 *ub = X + xr0;
#pragma unroll
                for (int sp = 0; sp < 8; sp++) {
                    const float4 u = ub[(sp + lane) & 7];
                    re[sp] = mk(u.x, u.y); im[sp] = mk(u.z, u.w);
                }
            }
            __syncwarp();      // the row buffer is free for the next position
            fft8s<-1>(re, im);
            float2 v[8];
#pragma unroll
            for (int c = 0; c < 8; c++) v[c] = __ffma2_rn(im[c], im[c], __fmul2_rn(re[c], re[c]));
            if (P.map) {
                // v[k2] is angle bin lane + 32 k2: one 128-byte line per store instruction
                float *mp = P.map + ((long long)cpi * Nr + n) * NA + lane;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    __stcs(mp + 32 * c, v[c].x);
                    __stcs(mp + NA + 32 * c, v[c].y);
                }
            }
            const float ma = fmaxf(fmaxf(fmaxf(v[0].x, v[1].x), fmaxf(v[2].x, v[3].x)), fmaxf(fmaxf(v[4].x, v[5].x), fmaxf(v[6].x, v[7].x)));
            const float mb = fmaxf(fmaxf(fmaxf(v[0].y, v[1].y), fmaxf(v[2].y, v[3].y)), fmaxf(fmaxf(v[4].y, v[5].y), fmaxf(v[6].y, v[7].y)));
            // row n precedes row n + 1; positions do not come in row order: an equal value only wins with a lower row
            sec_t = fmaxf(sec_t, fminf(ma, mb));
            const float mm = fmaxf(ma, mb);
            const int mrow = ma >= mb ? n : n + 1;
            sec_t = fmaxf(sec_t, fminf(mm, best));
            if (mm > best || (mm == best && mrow < best_row)) { best = mm; best_row = mrow; }
        }
        // ---- fold this unit's maximum and runner-up into keys[cpi] / sec[cpi] (as k_angle_mag does) ----
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
    }
}

}  // namespace jrc
