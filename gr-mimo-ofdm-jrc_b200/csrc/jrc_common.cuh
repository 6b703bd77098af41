// jrc_common.cuh -- device helpers shared by the staged and the fused kernels.
//
// Arithmetic rule for the whole library: every float operation whose rounding
// matters for parity is written with an explicit _rn intrinsic, so nvcc can
// neither contract (a*b+c -> fma) nor re-associate it.  That makes
//   - the staged kernels bit-identical to the CPU oracle's float pipeline, and
//   - the fused kernel's re-evaluation of a peak bit-identical to its main loop.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace jrc {

typedef float2 c32;

__device__ __forceinline__ c32 mk(float x, float y) { c32 r; r.x = x; r.y = y; return r; }

// std::complex<float> product with every product and sum rounded separately
// (what the reference's generic x86-64 build does, lib/mimo_ofdm_radar_impl.cc:273)
__device__ __forceinline__ c32 cmul_exact(c32 a, c32 b)
{
    return mk(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)),
              __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ c32 cadd_exact(c32 a, c32 b) { return mk(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ c32 csub_exact(c32 a, c32 b) { return mk(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)); }

// fast complex product for the fused kernel: 2 FMUL + 2 FFMA, pinned with intrinsics
__device__ __forceinline__ c32 cmul_fma(c32 a, c32 w)
{
    return mk(__fmaf_rn(-a.y, w.y, __fmul_rn(a.x, w.x)),
              __fmaf_rn(a.y, w.x, __fmul_rn(a.x, w.y)));
}

// std::pow(std::abs(z), 2) exactly as the reference evaluates it
// (lib/range_angle_estimator_impl.cc:141,217): hypotf -> float, squared in double.
// glibc's hypotf is (float)sqrt((double)x*x + (double)y*y); both products are exact
// in double so fma() vs mul+add round identically.
__device__ __forceinline__ float ref_abs(c32 z)
{
    double s = fma((double)z.x, (double)z.x, (double)z.y * (double)z.y);
    return (float)sqrt(s);
}
__device__ __forceinline__ double ref_pow_abs2(c32 z)
{
    double a = (double)ref_abs(z);
    return a * a;
}

// 8-point DFT, DIR = -1 forward (e^{-j2pi/8}), +1 backward; natural order in and out.
// 52 FP32 instructions: 40 FADD + 4 FADD (twiddle pre-sums) + 8 FFMA.
template <int DIR>
__device__ __forceinline__ void fft8(c32 (&u)[8])
{
    const float C = 0.70710678118654752440f;
    c32 a0 = mk(__fadd_rn(u[0].x, u[4].x), __fadd_rn(u[0].y, u[4].y));
    c32 a1 = mk(__fsub_rn(u[0].x, u[4].x), __fsub_rn(u[0].y, u[4].y));
    c32 a2 = mk(__fadd_rn(u[2].x, u[6].x), __fadd_rn(u[2].y, u[6].y));
    c32 a3 = mk(__fsub_rn(u[2].x, u[6].x), __fsub_rn(u[2].y, u[6].y));
    c32 a4 = mk(__fadd_rn(u[1].x, u[5].x), __fadd_rn(u[1].y, u[5].y));
    c32 a5 = mk(__fsub_rn(u[1].x, u[5].x), __fsub_rn(u[1].y, u[5].y));
    c32 a6 = mk(__fadd_rn(u[3].x, u[7].x), __fadd_rn(u[3].y, u[7].y));
    c32 a7 = mk(__fsub_rn(u[3].x, u[7].x), __fsub_rn(u[3].y, u[7].y));
    c32 E0 = mk(__fadd_rn(a0.x, a2.x), __fadd_rn(a0.y, a2.y));
    c32 E2 = mk(__fsub_rn(a0.x, a2.x), __fsub_rn(a0.y, a2.y));
    c32 O0 = mk(__fadd_rn(a4.x, a6.x), __fadd_rn(a4.y, a6.y));
    c32 O2 = mk(__fsub_rn(a4.x, a6.x), __fsub_rn(a4.y, a6.y));
    c32 E1, E3, O1, O3, t1, t3;
    if (DIR < 0) {   // multiply by -j: (x,y) -> (y,-x)
        E1 = mk(__fadd_rn(a1.x, a3.y), __fsub_rn(a1.y, a3.x));
        E3 = mk(__fsub_rn(a1.x, a3.y), __fadd_rn(a1.y, a3.x));
        O1 = mk(__fadd_rn(a5.x, a7.y), __fsub_rn(a5.y, a7.x));
        O3 = mk(__fsub_rn(a5.x, a7.y), __fadd_rn(a5.y, a7.x));
        t1 = mk(__fadd_rn(O1.x, O1.y), __fsub_rn(O1.y, O1.x));      // O1*(1-j)
        t3 = mk(__fsub_rn(O3.y, O3.x), __fsub_rn(-O3.x, O3.y));     // O3*(-1-j)
        u[2] = mk(__fadd_rn(E2.x, O2.y), __fsub_rn(E2.y, O2.x));
        u[6] = mk(__fsub_rn(E2.x, O2.y), __fadd_rn(E2.y, O2.x));
    } else {         // multiply by +j: (x,y) -> (-y,x)
        E1 = mk(__fsub_rn(a1.x, a3.y), __fadd_rn(a1.y, a3.x));
        E3 = mk(__fadd_rn(a1.x, a3.y), __fsub_rn(a1.y, a3.x));
        O1 = mk(__fsub_rn(a5.x, a7.y), __fadd_rn(a5.y, a7.x));
        O3 = mk(__fadd_rn(a5.x, a7.y), __fsub_rn(a5.y, a7.x));
        t1 = mk(__fsub_rn(O1.x, O1.y), __fadd_rn(O1.x, O1.y));      // O1*(1+j)
        t3 = mk(__fsub_rn(-O3.x, O3.y), __fsub_rn(O3.x, O3.y));     // O3*(-1+j)
        u[2] = mk(__fsub_rn(E2.x, O2.y), __fadd_rn(E2.y, O2.x));
        u[6] = mk(__fadd_rn(E2.x, O2.y), __fsub_rn(E2.y, O2.x));
    }
    u[0] = mk(__fadd_rn(E0.x, O0.x), __fadd_rn(E0.y, O0.y));
    u[4] = mk(__fsub_rn(E0.x, O0.x), __fsub_rn(E0.y, O0.y));
    u[1] = mk(__fmaf_rn(C, t1.x, E1.x), __fmaf_rn(C, t1.y, E1.y));
    u[5] = mk(__fmaf_rn(-C, t1.x, E1.x), __fmaf_rn(-C, t1.y, E1.y));
    u[3] = mk(__fmaf_rn(C, t3.x, E3.x), __fmaf_rn(C, t3.y, E3.y));
    u[7] = mk(__fmaf_rn(-C, t3.x, E3.x), __fmaf_rn(-C, t3.y, E3.y));
}


// ---------------------------------------------------------------------------
// Same 8-point DFT on Blackwell's packed FP32 pipe (FADD2 / FFMA2, sm_100+): a complex
// add is ONE instruction on an (re,im) register pair, a - b is fma(b, -1, a) (exactly
// a - b rounded once) and the +-j rotations are fma(swapped, (+-1,-+1), .).  Every lane
// result is bit-identical to fft8<DIR>; 31 instead of 52 issue slots.
// ---------------------------------------------------------------------------
template <int DIR>
__device__ __forceinline__ void fft8p(c32 (&u)[8])
{
    const float C = 0.70710678118654752440f;
    const c32 N1 = mk(-1.f, -1.f);
    // multiply by -j: (x,y) -> (y,-x) = swapped * (1,-1);  by +j: (-y,x) = swapped * (-1,1)
    const c32 RP = (DIR < 0) ? mk(1.f, -1.f) : mk(-1.f, 1.f);
    const c32 RM = (DIR < 0) ? mk(-1.f, 1.f) : mk(1.f, -1.f);
    const c32 CC = mk(C, C), NC = mk(-C, -C);
    c32 a0 = __fadd2_rn(u[0], u[4]);
    c32 a1 = __ffma2_rn(u[4], N1, u[0]);
    c32 a2 = __fadd2_rn(u[2], u[6]);
    c32 a3s = mk(__fsub_rn(u[2].y, u[6].y), __fsub_rn(u[2].x, u[6].x));   // (u2-u6) with re/im swapped
    c32 a4 = __fadd2_rn(u[1], u[5]);
    c32 a5 = __ffma2_rn(u[5], N1, u[1]);
    c32 a6 = __fadd2_rn(u[3], u[7]);
    c32 a7s = mk(__fsub_rn(u[3].y, u[7].y), __fsub_rn(u[3].x, u[7].x));
    c32 E0 = __fadd2_rn(a0, a2);
    c32 E2 = __ffma2_rn(a2, N1, a0);
    c32 O0 = __fadd2_rn(a4, a6);
    c32 O2s = mk(__fsub_rn(a4.y, a6.y), __fsub_rn(a4.x, a6.x));
    c32 E1 = __ffma2_rn(a3s, RP, a1);
    c32 E3 = __ffma2_rn(a3s, RM, a1);
    c32 O1 = __ffma2_rn(a7s, RP, a5);
    c32 O3 = __ffma2_rn(a7s, RM, a5);
    c32 t1, t3;
    if (DIR < 0) {
        t1 = mk(__fadd_rn(O1.x, O1.y), __fsub_rn(O1.y, O1.x));      // O1*(1-j)
        t3 = mk(__fsub_rn(O3.y, O3.x), __fsub_rn(-O3.x, O3.y));     // O3*(-1-j)
    } else {
        t1 = mk(__fsub_rn(O1.x, O1.y), __fadd_rn(O1.x, O1.y));      // O1*(1+j)
        t3 = mk(__fsub_rn(-O3.x, O3.y), __fsub_rn(O3.x, O3.y));     // O3*(-1+j)
    }
    u[0] = __fadd2_rn(E0, O0);
    u[4] = __ffma2_rn(O0, N1, E0);
    u[2] = __ffma2_rn(O2s, RP, E2);
    u[6] = __ffma2_rn(O2s, RM, E2);
    u[1] = __ffma2_rn(t1, CC, E1);
    u[5] = __ffma2_rn(t1, NC, E1);
    u[3] = __ffma2_rn(t3, CC, E3);
    u[7] = __ffma2_rn(t3, NC, E3);
}

// ---------------------------------------------------------------------------
// TWO independent 8-point DFTs per call on the packed pipe, split-complex: re[k] / im[k] hold the real /
// imaginary parts of element k of DFT 0 (.x) and DFT 1 (.y).  No lane of a packed instruction is
// wasted on re/im swaps, so the pair costs the 52 issue slots of ONE scalar fft8<DIR> (26 per DFT),
// and each half is bit-identical to fft8<DIR> of that DFT (a - b is FADD2 with a negated operand; the -(x+y)
// term of the 3*pi/4 twiddle is carried as x+y with the sign moved into the constant).
// ---------------------------------------------------------------------------
template <int DIR>
__device__ __forceinline__ void fft8s(float2 (&re)[8], float2 (&im)[8])
{
    const float C = 0.70710678118654752440f;
    const float2 CC = mk(C, C), NC = mk(-C, -C);
#define JRC_ADD(a, b) __fadd2_rn(a, b)
#define JRC_SUB(a, b) __fadd2_rn(a, mk(-(b).x, -(b).y))   /* FADD2 a, -b: the negation is an operand modifier */
    const float2 a0r = JRC_ADD(re[0], re[4]), a0i = JRC_ADD(im[0], im[4]);
    const float2 a1r = JRC_SUB(re[0], re[4]), a1i = JRC_SUB(im[0], im[4]);
    const float2 a2r = JRC_ADD(re[2], re[6]), a2i = JRC_ADD(im[2], im[6]);
    const float2 a3r = JRC_SUB(re[2], re[6]), a3i = JRC_SUB(im[2], im[6]);
    const float2 a4r = JRC_ADD(re[1], re[5]), a4i = JRC_ADD(im[1], im[5]);
    const float2 a5r = JRC_SUB(re[1], re[5]), a5i = JRC_SUB(im[1], im[5]);
    const float2 a6r = JRC_ADD(re[3], re[7]), a6i = JRC_ADD(im[3], im[7]);
    const float2 a7r = JRC_SUB(re[3], re[7]), a7i = JRC_SUB(im[3], im[7]);
    const float2 E0r = JRC_ADD(a0r, a2r), E0i = JRC_ADD(a0i, a2i);
    const float2 E2r = JRC_SUB(a0r, a2r), E2i = JRC_SUB(a0i, a2i);
    const float2 O0r = JRC_ADD(a4r, a6r), O0i = JRC_ADD(a4i, a6i);
    const float2 O2r = JRC_SUB(a4r, a6r), O2i = JRC_SUB(a4i, a6i);
    float2 E1r, E1i, E3r, E3i, O1r, O1i, O3r, O3i, t1r, t1i, t3r, s3;
    if (DIR < 0) {   // multiply by -j: (x,y) -> (y,-x)
        E1r = JRC_ADD(a1r, a3i); E1i = JRC_SUB(a1i, a3r);
        E3r = JRC_SUB(a1r, a3i); E3i = JRC_ADD(a1i, a3r);
        O1r = JRC_ADD(a5r, a7i); O1i = JRC_SUB(a5i, a7r);
        O3r = JRC_SUB(a5r, a7i); O3i = JRC_ADD(a5i, a7r);
        t1r = JRC_ADD(O1r, O1i); t1i = JRC_SUB(O1i, O1r);     // O1*(1-j)
        t3r = JRC_SUB(O3i, O3r); s3 = JRC_ADD(O3r, O3i);      // O3*(-1-j) = (t3r, -s3)
        re[2] = JRC_ADD(E2r, O2i); im[2] = JRC_SUB(E2i, O2r);
        re[6] = JRC_SUB(E2r, O2i); im[6] = JRC_ADD(E2i, O2r);
        re[1] = __ffma2_rn(CC, t1r, E1r); im[1] = __ffma2_rn(CC, t1i, E1i);
        re[5] = __ffma2_rn(NC, t1r, E1r); im[5] = __ffma2_rn(NC, t1i, E1i);
        re[3] = __ffma2_rn(CC, t3r, E3r); im[3] = __ffma2_rn(NC, s3, E3i);
        re[7] = __ffma2_rn(NC, t3r, E3r); im[7] = __ffma2_rn(CC, s3, E3i);
    } else {         // multiply by +j: (x,y) -> (-y,x)
        E1r = JRC_SUB(a1r, a3i); E1i = JRC_ADD(a1i, a3r);
        E3r = JRC_ADD(a1r, a3i); E3i = JRC_SUB(a1i, a3r);
        O1r = JRC_SUB(a5r, a7i); O1i = JRC_ADD(a5i, a7r);
        O3r = JRC_ADD(a5r, a7i); O3i = JRC_SUB(a5i, a7r);
        t1r = JRC_SUB(O1r, O1i); t1i = JRC_ADD(O1r, O1i);     // O1*(1+j)
        s3 = JRC_ADD(O3r, O3i); t3r = JRC_SUB(O3r, O3i);      // O3*(-1+j) = (-s3, t3r)
        re[2] = JRC_SUB(E2r, O2i); im[2] = JRC_ADD(E2i, O2r);
        re[6] = JRC_ADD(E2r, O2i); im[6] = JRC_SUB(E2i, O2r);
        re[1] = __ffma2_rn(CC, t1r, E1r); im[1] = __ffma2_rn(CC, t1i, E1i);
        re[5] = __ffma2_rn(NC, t1r, E1r); im[5] = __ffma2_rn(NC, t1i, E1i);
        re[3] = __ffma2_rn(NC, s3, E3r); im[3] = __ffma2_rn(CC, t3r, E3i);
        re[7] = __ffma2_rn(CC, s3, E3r); im[7] = __ffma2_rn(NC, t3r, E3i);
    }
    re[0] = JRC_ADD(E0r, O0r); im[0] = JRC_ADD(E0i, O0i);
    re[4] = JRC_SUB(E0r, O0r); im[4] = JRC_SUB(E0i, O0i);
#undef JRC_ADD
#undef JRC_SUB
}

#ifndef JRC_SCALAR_FFT8
#define JRC_FFT8 fft8p
#else
#define JRC_FFT8 fft8
#endif

// e^{j*pi*num/den} with an exactly representable argument (den a power of two)
__device__ __forceinline__ c32 cispi_ratio(int num, int den)
{
    float s, c;
    sincospif((float)num / (float)den, &s, &c);
    return mk(c, s);
}

// 10*log10(peak/noise) as the reference evaluates it (lib/range_angle_estimator_impl.cc:227): float division, log10f, float product.  log10 is taken
// in double and rounded once, which is what a correctly rounded log10f returns.
__device__ __forceinline__ float snr_db_of(float peak, float noise)
{
    return __fmul_rn(10.f, (float)log10((double)__fdiv_rn(peak, noise)));
}

// the fast kernels' form: device log10f (<= 2 ulp).  Their gate decision is only kept when it is further from the
// threshold than that (gate_is_marginal, jrc_exact.cuh), and host-side callers recompute the published value with libm.
__device__ __forceinline__ float snr_db_fast(float peak, float noise)
{
    return __fmul_rn(10.f, log10f(__fdiv_rn(peak, noise)));
}

// acc <- (float)((double)acc + (double)v * (double)v) for v = chunk[0 .. cnt) in order: the reference's `float += double`
// (lib/range_angle_estimator_impl.cc:217), a serial chain of convert, add, convert per window cell (~35 cycles).  One warp
// runs it speculatively 32 cells at a time.  With p = v*v = ph + pl (ph = RN(v*v), pl = fma(v, v, -ph), exact) and
// t = RN(acc + ph), e = its exact rounding error (TwoSum), the exact sum is t + (e + pl): whenever |e + pl| stays clear of
// half an ulp of t and t is not a power of two, the reference's result IS t -- rounding the exact sum to double first
// moves it by 2^-53 relative, 2^-29 of that ulp.  So every lane carries the same chain t <- RN(t + ph_k) (one FADD per
// cell, the ph_k by shuffle), lane k keeps the two values of ITS step and checks it afterwards, off the chain; a block with a
// step inside the margin (or a start from zero, a power of two, NaN, Inf, tiny values) is redone with the reference's own
// expression.  ~2.5e-4 of the steps of a noise sum are; NumPy model against the expression, midpoints included:
// tests/test_oracle_kat.py::test_speculative_sequential_sum.  All 32 lanes call it with the same arguments.
__device__ __forceinline__ float seq_sum_sq_warp(float acc, const float *chunk, int cnt, int lane)
{
    for (int j0 = 0; j0 < cnt; j0 += 32) {
        const int n = min(32, cnt - j0);
        const float a = lane < n ? chunk[j0 + lane] : 0.f;
        const float ph = __fmul_rn(a, a), pl = __fmaf_rn(a, a, -ph);
        float t = acc, my_s = 0.f, my_t = 0.f;
#pragma unroll 8
        for (int k = 0; k < n; k++) {
            const float phk = __shfl_sync(0xffffffffu, ph, k);
            const float tn = __fadd_rn(t, phk);
            if (lane == k) { my_s = t; my_t = tn; }
            t = tn;
        }
        bool safe = true;
        if (lane < n) {
            const float bv = __fsub_rn(my_t, my_s);
            const float e = __fadd_rn(__fsub_rn(my_s, __fsub_rn(my_t, bv)), __fsub_rn(ph, bv));
            const float d = __fadd_rn(e, pl);
            const unsigned tb = __float_as_uint(my_t), ex = (tb >> 23) & 0xffu;
            const float ul = __uint_as_float((tb & 0x7f800000u) - (23u << 23));
            safe = (fabsf(d) <= __fmul_rn(0.49999f, ul)) && (tb & 0x007fffffu) != 0u && ex > 30u && ex < 250u;
        }
        if (__all_sync(0xffffffffu, safe)) { acc = t; continue; }
        float s = acc;
        for (int k = 0; k < n; k++) {
            const double ak = (double)__shfl_sync(0xffffffffu, a, k);
            s = (float)((double)s + ak * ak);
        }
        acc = s;
    }
    return acc;
}

__device__ __forceinline__ unsigned long long pack_key(float v, unsigned idx)
{
    // v >= 0 and not NaN: float bits are monotonic; ties -> the smaller index wins
    return ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}

}  // namespace jrc
