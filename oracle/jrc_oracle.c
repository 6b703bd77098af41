/*
 * jrc_oracle.c -- CPU restatement of the gr-mimo-ofdm-jrc radar hot path.
 * TEST INFRASTRUCTURE ONLY (see jrc_oracle.h for the usage rule and the
 * parity-pinning statement).  Build: oracle/Makefile (-O2 -ffp-contract=off,
 * so that every float product/sum is rounded separately exactly as the
 * reference's x86-64 build without -march does).
 *
 * Paths in comments are relative to /root/reference.
 */
#include "jrc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------ */
/* helpers                                                                   */
/* ------------------------------------------------------------------------ */

/* std::pow(std::abs(z), 2) as evaluated by the reference
 * (lib/range_angle_estimator_impl.cc:141, :217): std::abs(complex<float>)
 * is hypotf -> float; std::pow(float,int) promotes to pow(double,double).
 * The double square of a float is exact, so pow() returns it exactly.     */
static inline double ref_pow_abs2(orc_c32 z)
{
    float a = hypotf(z.re, z.im);
    return pow((double)a, 2.0);
}

/* std::complex<float> operator* (no Annex-G recovery needed for finite data):
 * (a+ib)(c+id) = (ac-bd) + i(ad+bc), every product and sum rounded to float */
static inline orc_c32 cmulf(orc_c32 x, orc_c32 y)
{
    orc_c32 r;
    r.re = x.re * y.re - x.im * y.im;
    r.im = x.re * y.im + x.im * y.re;
    return r;
}

/* ------------------------------------------------------------------------ */
/* mimo_ofdm_radar                                                           */
/* ------------------------------------------------------------------------ */

orc_radar *orc_radar_create(int fft_len, int n_tx, int n_rx, int n_sym, int n_pre,
                            int background_removal, int background_recording,
                            int record_len, int interp_factor, int tx_interleave)
{
    /* ctor: lib/mimo_ofdm_radar_impl.cc:66-120 */
    orc_radar *r = (orc_radar *)calloc(1, sizeof(*r));
    size_t vn = (size_t)n_tx * n_rx * fft_len;
    r->fft_len = fft_len; r->n_tx = n_tx; r->n_rx = n_rx; r->n_sym = n_sym; r->n_pre = n_pre;
    r->interp_factor = interp_factor; r->tx_interleave = tx_interleave;
    r->background_removal = background_removal;
    r->background_recording = background_recording;
    r->record_len = record_len;
    r->chan_est = (orc_c32 *)calloc(vn, sizeof(orc_c32));
    r->chan_temp = (orc_c32 *)calloc(vn, sizeof(orc_c32));  /* resize() -> zeros (:116) */
    r->ring = (orc_c32 *)calloc(vn * (size_t)(record_len > 0 ? record_len : 1), sizeof(orc_c32));
    r->ring_size = 0; r->ring_head = 0;
    return r;
}

void orc_radar_destroy(orc_radar *r)
{
    if (!r) return;
    free(r->chan_est); free(r->chan_temp); free(r->ring); free(r);
}

void orc_radar_set_background_record(orc_radar *r, int on) { r->background_recording = on; }

int orc_radar_work(orc_radar *r, const orc_c32 *const *tx, const orc_c32 *const *rx,
                   size_t tx_skip_items, orc_c32 *out)
{
    const int N = r->fft_len, T = r->n_tx, R = r->n_rx, S = r->n_sym;
    const int V = T * R;
    const size_t vn = (size_t)V * N;
    orc_c32 *H = r->chan_est;

    /* :243-244 */
    memset(out, 0, sizeof(orc_c32) * vn * (size_t)r->interp_factor);
    memset(H, 0, sizeof(orc_c32) * vn);

    /* :250-295 -- loop order kept (sc outer) although it does not matter */
    for (int i_sc = 0; i_sc < N; i_sc++) {
        for (int i_rx = 0; i_rx < R; i_rx++) {
            const orc_c32 *in_rx = rx[i_rx] + (size_t)N * r->n_pre;            /* :254-255 */
            for (int i_tx = 0; i_tx < T; i_tx++) {
                const orc_c32 *in_tx = tx[i_tx] + (size_t)N * r->n_pre
                                       + (size_t)N * tx_skip_items;             /* :258-260 */
                int idx = r->tx_interleave ? i_sc + N * (i_tx * R + i_rx)      /* :262-269 */
                                           : i_sc + N * (i_rx * T + i_tx);
                for (int i_sym = 0; i_sym < S; i_sym++) {                       /* :271-274 */
                    orc_c32 a = in_rx[i_sc + i_sym * N];
                    orc_c32 b = in_tx[i_sc + i_sym * N];
                    orc_c32 cb = { b.re, -b.im };
                    orc_c32 p = cmulf(a, cb);
                    H[idx].re = H[idx].re + p.re;
                    H[idx].im = H[idx].im + p.im;
                }
                if (r->background_recording)                                    /* :276-279 */
                    r->chan_temp[idx] = H[idx];
                if (r->background_removal) {                                    /* :281-292 */
                    int sz = r->ring_size;
                    orc_c32 mean = { 0.0f, 0.0f };
                    for (int b = 0; b < sz; b++) {
                        const orc_c32 *e = r->ring + vn * (size_t)((r->ring_head + b) % r->record_len);
                        mean.re += e[idx].re / (float)sz;
                        mean.im += e[idx].im / (float)sz;
                    }
                    H[idx].re = H[idx].re - mean.re;
                    H[idx].im = H[idx].im - mean.im;
                }
            }
        }
    }
    /* :297-300  boost::circular_buffer::push_back (overwrites the oldest when full;
     * a zero-capacity buffer ignores the push) */
    if (r->background_removal && r->record_len > 0) {
        if (r->ring_size < r->record_len) {
            int slot = (r->ring_head + r->ring_size) % r->record_len;
            memcpy(r->ring + vn * (size_t)slot, r->chan_temp, sizeof(orc_c32) * vn);
            r->ring_size++;
        } else {
            memcpy(r->ring + vn * (size_t)r->ring_head, r->chan_temp, sizeof(orc_c32) * vn);
            r->ring_head = (r->ring_head + 1) % r->record_len;
        }
    }
    /* :312-315 */
    for (int p = 0; p < V; p++)
        memcpy(out + (size_t)p * N * r->interp_factor, H + (size_t)p * N, sizeof(orc_c32) * N);
    return V;                                                                   /* :303,339 */
}

/* ------------------------------------------------------------------------ */
/* fft_vcc                                                                   */
/* ------------------------------------------------------------------------ */

static int is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

/* float32 iterative radix-2 DIT, twiddles rounded from double.  tw: n/2 entries
 * e^{sign*j*2*pi*k/n}.  In-place on x (already bit-reversed by the caller).  */
static void fft_pow2_inplace(orc_c32 *x, int n, const orc_c32 *tw)
{
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int i = 0; i < n; i += len) {
            for (int j = 0; j < half; j++) {
                orc_c32 w = tw[j * step];
                orc_c32 u = x[i + j];
                orc_c32 v = cmulf(x[i + j + half], w);
                x[i + j].re = u.re + v.re;        x[i + j].im = u.im + v.im;
                x[i + j + half].re = u.re - v.re; x[i + j + half].im = u.im - v.im;
            }
        }
    }
}

static void make_twiddles(orc_c32 *tw, int n, int forward)
{
    double sgn = forward ? -1.0 : 1.0;
    for (int k = 0; k < n / 2; k++) {
        double a = sgn * 2.0 * M_PI * (double)k / (double)n;
        tw[k].re = (float)cos(a); tw[k].im = (float)sin(a);
    }
}

static void bitrev_copy(const orc_c32 *in, orc_c32 *out, int n)
{
    int bits = 0; while ((1 << bits) < n) bits++;
    for (int i = 0; i < n; i++) {
        unsigned r = 0, v = (unsigned)i;
        for (int b = 0; b < bits; b++) { r = (r << 1) | (v & 1u); v >>= 1; }
        out[r] = in[i];
    }
}

/* any n: float64 direct DFT, result rounded to float32 (stands in for FFTW3f on
 * the non-power-of-two packet lengths of target_simulator)                  */
static void dft_any(const orc_c32 *in, orc_c32 *out, int n, int forward)
{
    double *c = (double *)malloc(sizeof(double) * 2 * (size_t)n);
    double sgn = forward ? -1.0 : 1.0;
    for (int k = 0; k < n; k++) {
        double a = sgn * 2.0 * M_PI * (double)k / (double)n;
        c[2 * k] = cos(a); c[2 * k + 1] = sin(a);
    }
    for (int m = 0; m < n; m++) {
        double sr = 0.0, si = 0.0;
        size_t idx = 0;
        for (int k = 0; k < n; k++) {
            double wr = c[2 * idx], wi = c[2 * idx + 1];
            sr += (double)in[k].re * wr - (double)in[k].im * wi;
            si += (double)in[k].re * wi + (double)in[k].im * wr;
            idx += (size_t)m; if (idx >= (size_t)n) idx -= (size_t)n;
        }
        out[m].re = (float)sr; out[m].im = (float)si;
    }
    free(c);
}

/* GNU Radio 3.8 gr-fft fft_vcc_fftw::work semantics: window none;
 * !forward && shift -> swap input halves; execute; forward && shift -> swap
 * output halves (out[i] = X[(i + ceil(n/2)) mod n]); no scaling.            */
static void fft_vcc_one(const orc_c32 *in, orc_c32 *out, int n, int forward, int shift,
                        const orc_c32 *tw, orc_c32 *tmp, orc_c32 *tmp2)
{
    const orc_c32 *src = in;
    if (!forward && shift) {
        int offset = (n + 1) / 2;  /* fft_vcc: second half first */
        for (int i = 0; i < n; i++) tmp2[i] = in[(i + offset) % n];
        src = tmp2;
    }
    if (is_pow2(n)) { bitrev_copy(src, tmp, n); fft_pow2_inplace(tmp, n, tw); }
    else            { dft_any(src, tmp, n, forward); }
    if (forward && shift) {
        int offset = (n + 1) / 2;
        for (int i = 0; i < n; i++) out[i] = tmp[(i + offset) % n];
    } else {
        memcpy(out, tmp, sizeof(orc_c32) * (size_t)n);
    }
}

void orc_fft_vcc_batch(const orc_c32 *in, orc_c32 *out, int n, int batch, int forward, int shift)
{
    orc_c32 *tw = (orc_c32 *)malloc(sizeof(orc_c32) * (size_t)(n / 2 + 1));
    orc_c32 *tmp = (orc_c32 *)malloc(sizeof(orc_c32) * (size_t)n);
    orc_c32 *tmp2 = (orc_c32 *)malloc(sizeof(orc_c32) * (size_t)n);
    if (is_pow2(n)) make_twiddles(tw, n, forward);
    for (int b = 0; b < batch; b++)
        fft_vcc_one(in + (size_t)b * n, out + (size_t)b * n, n, forward, shift, tw, tmp, tmp2);
    free(tw); free(tmp); free(tmp2);
}

void orc_fft_vcc(const orc_c32 *in, orc_c32 *out, int n, int forward, int shift)
{
    orc_fft_vcc_batch(in, out, n, 1, forward, shift);
}

/* ------------------------------------------------------------------------ */
/* matrix_transpose (lib/matrix_transpose_impl.cc:97-104)                    */
/* ------------------------------------------------------------------------ */
void orc_matrix_transpose(const orc_c32 *in, int k_items, int input_len,
                          int output_len, int interp, orc_c32 *out)
{
    memset(out, 0, sizeof(orc_c32) * (size_t)interp * output_len * input_len);   /* :97 */
    for (int l = 0; l < input_len; l++)                                          /* :100 */
        for (int k = 0; k < k_items; k++)                                        /* :101 */
            out[(size_t)l * output_len * interp + k] = in[(size_t)k * input_len + l];
}

/* blocks_complex_to_mag_squared -> volk_32fc_magnitude_squared_32f (generic kernel) */
void orc_mag_squared(const orc_c32 *in, float *out, size_t n)
{
    for (size_t i = 0; i < n; i++) out[i] = in[i].re * in[i].re + in[i].im * in[i].im;
}

/* ------------------------------------------------------------------------ */
/* range_angle_estimator (lib/range_angle_estimator_impl.cc:122-283)         */
/* ------------------------------------------------------------------------ */
void orc_range_angle_estimate(const orc_c32 *in, int n_inputs, int vlen,
                              const float *range_bins, int n_range_bins,
                              const float *angle_bins, int n_angle_bins,
                              float noise_discard_range_m, float noise_discard_angle_deg,
                              float snr_threshold, float power_threshold,
                              orc_det *det, orc_est_dbg *dbg)
{
    float peak_power = -1;
    float curr_power;
    int peak_range_idx = -1, peak_angle_idx = -1;

    for (int i_range = 0; i_range < n_inputs; i_range++) {                       /* :137 */
        for (int i_angle = 0; i_angle < vlen; i_angle++) {                       /* :139 */
            curr_power = (float)ref_pow_abs2(in[i_angle + (size_t)vlen * i_range]); /* :141 */
            if (curr_power > peak_power) {                                       /* :144 */
                peak_power = curr_power; peak_range_idx = i_range; peak_angle_idx = i_angle;
            }
        }
    }
    memset(det, 0, sizeof(*det));
    det->range_idx = peak_range_idx; det->angle_idx = peak_angle_idx;
    det->peak_power = peak_power;
    if (peak_range_idx < 0) {        /* all-NaN / empty map: the reference would index [-1] (UB) */
        det->noise_power = NAN; det->snr_db = NAN; return;
    }
    float angle_val = angle_bins[peak_angle_idx];                                /* :152 */
    float range_val = range_bins[peak_range_idx];                                /* :153 */

    float angle_null = angle_val + 90;                                           /* :155 */
    if (angle_null >= 90) angle_null = angle_null - 180;                         /* :157-160 */

    /* std::lower_bound (:163-167): first element >= angle_null */
    int geq = 0;
    { int lo = 0, hi = n_angle_bins;
      while (lo < hi) { int mid = lo + (hi - lo) / 2; if (angle_bins[mid] < angle_null) lo = mid + 1; else hi = mid; }
      geq = lo; }
    int angle_null_idx;
    if (geq == 0) {                                                              /* :172 */
        angle_null_idx = 0;
    } else if (geq == n_angle_bins) {
        /* :169-170 read one float past the vector (UB).  Rule adopted:
         * the garbage is far from angle_null -> the previous bin wins           */
        angle_null_idx = geq - 1;
    } else {
        double a = angle_bins[geq - 1], b = angle_bins[geq];                     /* :169-170 */
        if (fabs(angle_null - a) < fabs(angle_null - b)) angle_null_idx = geq - 1; /* :175 */
        else angle_null_idx = geq;                                               /* :178 */
    }
    if (angle_null_idx == n_angle_bins - 1) angle_null_idx = n_angle_bins - 2;   /* :184-187 */

    int discard_range_idx = (int)(noise_discard_range_m / (range_bins[1] - range_bins[0]));  /* :189 */
    int discard_angle_idx = (int)(noise_discard_angle_deg /
        (angle_bins[(angle_null_idx + 1) % n_angle_bins] - angle_bins[angle_null_idx]));       /* :190 */
    if (discard_angle_idx <= 0) discard_angle_idx = 1;                           /* :192-195 */

    int start_range_idx = peak_range_idx + n_range_bins / 2 - discard_range_idx; /* :197 */
    int end_range_idx   = peak_range_idx + n_range_bins / 2 + discard_range_idx; /* :198 */
    int start_angle_idx = angle_null_idx - discard_angle_idx;                    /* :200 */
    int end_angle_idx   = angle_null_idx + discard_angle_idx;                    /* :201 */

    float noise_power = 0;
    int n_noise_samples = 0;
    for (int i_range = start_range_idx; i_range < end_range_idx; i_range++) {    /* :211 */
        int r_idx = ((i_range % n_inputs) + n_inputs) % n_inputs;
        for (int i_angle = start_angle_idx; i_angle < end_angle_idx; i_angle++) {
            int a_idx = ((i_angle % vlen) + vlen) % vlen;
            /* float += double: the sum is formed in double, then rounded (:217) */
            noise_power = (float)((double)noise_power + ref_pow_abs2(in[a_idx + (size_t)vlen * r_idx]));
            n_noise_samples++;
        }
    }
    noise_power = noise_power / n_noise_samples;                                 /* :226 */
    float snr_est = 10 * log10f(peak_power / noise_power);                       /* :227 */

    det->noise_power = noise_power; det->snr_db = snr_est; det->n_noise = n_noise_samples;
    det->flags = (snr_est >= snr_threshold && peak_power >= power_threshold) ? 1u : 0u;  /* :234 */
    if (dbg) {
        dbg->angle_null_idx = angle_null_idx; dbg->discard_range_idx = discard_range_idx;
        dbg->discard_angle_idx = discard_angle_idx;
        dbg->start_range_idx = start_range_idx; dbg->end_range_idx = end_range_idx;
        dbg->start_angle_idx = start_angle_idx; dbg->end_angle_idx = end_angle_idx;
        dbg->range_val = range_val; dbg->angle_val = angle_val;
    }
}

/* ------------------------------------------------------------------------ */
/* fft_peak_detect (lib/fft_peak_detect_impl.cc:77-111)                      */
/* ------------------------------------------------------------------------ */
void orc_fft_peak_detect(const orc_c32 *in, int n, int samp_rate, float interp_factor,
                         float threshold_db, int samp_protect, orc_peak1d *out)
{
    int k = -1;
    float hold = -1;
    double thr = pow(10, threshold_db / 10.0);                                   /* :89 */
    for (int p = samp_protect; p < n - samp_protect; p++) {                      /* :88 */
        float a = hypotf(in[p].re, in[p].im);
        if (a > hold && pow((double)a, 2.0) > thr) { hold = a; k = p; }          /* :89-92 */
    }
    out->k = k; out->freq = 0.0f; out->phase = 0.0f; out->mag = 0.0f;
    if (k != -1) {                                                               /* :98 */
        if (k <= n / 2)
            out->freq = k / (float)n * (samp_rate * interp_factor);              /* :100 */
        else
            out->freq = -((float)samp_rate * interp_factor)
                        + k * (samp_rate * interp_factor / (float)n);            /* :103 */
        out->phase = atan2f(in[k].im, in[k].re);                                 /* :105 std::arg */
        out->mag = hypotf(in[k].re, in[k].im);                                   /* :106 */
    }
    /* k == -1: the reference writes nothing but still returns 1 item (stale buffer
     * content); the oracle pins zeros and reports k = -1.                       */
}

/* ------------------------------------------------------------------------ */
/* zero_pad (lib/zero_pad_impl.cc:67-94)                                     */
/* ------------------------------------------------------------------------ */
static uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static float gauss(uint64_t *s, float sigma)
{
    double u1 = ((double)(splitmix64(s) >> 11) + 1.0) / 9007199254740993.0;
    double u2 = (double)(splitmix64(s) >> 11) / 9007199254740992.0;
    return (float)(sigma * sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2));
}
void orc_zero_pad(const orc_c32 *in, int n, unsigned pad_front, unsigned pad_tail,
                  uint64_t seed, orc_c32 *out)
{
    /* the reference seeds std::default_random_engine from std::random_device per call
     * (:76-77): only N(0, 1e-2) statistics are defined, not the values             */
    uint64_t s = seed;
    for (unsigned i = 0; i < pad_front; i++) { out[i].re = gauss(&s, 1e-2f); out[i].im = gauss(&s, 1e-2f); }
    memcpy(out + pad_front, in, sizeof(orc_c32) * (size_t)n);                    /* :89 */
    for (unsigned i = 0; i < pad_tail; i++) {
        out[pad_front + n + i].re = gauss(&s, 1e-2f); out[pad_front + n + i].im = gauss(&s, 1e-2f);
    }
}

/* ofdm_cyclic_prefix_remover (lib/ofdm_cyclic_prefix_remover_impl.cc:92-95) */
void orc_cp_remove(const orc_c32 *in, int n_sym, int fft_len, int cp_len, orc_c32 *out)
{
    for (int k = 0; k < n_sym; k++)
        memcpy(out + (size_t)fft_len * k, in + cp_len + (size_t)k * (fft_len + cp_len),
               sizeof(orc_c32) * (size_t)fft_len);
}

/* ------------------------------------------------------------------------ */
/* target_simulator (lib/target_simulator_impl.cc:127-385)                   */
/* ------------------------------------------------------------------------ */
void orc_target_simulator(const orc_c32 *in, int n,
                          const float *range, const float *velocity, const float *rcs,
                          const float *azimuth, int n_targets,
                          const float *position_rx, int n_rx,
                          int samp_rate, float center_freq,
                          int self_coupling, float self_coupling_db,
                          int accumulate, orc_c32 *out)
{
    const float c_light = 3e8f;                              /* target_simulator_impl.h:83 */
    const double FOUR_PI_CUBED_SQRT = 44.54662397465366;     /* :33 */
    float *freq = (float *)malloc(sizeof(float) * (size_t)n);
    orc_c32 *filt_dopp = (orc_c32 *)malloc(sizeof(orc_c32) * (size_t)n);
    orc_c32 *filt_time = (orc_c32 *)malloc(sizeof(orc_c32) * (size_t)n);
    orc_c32 *bt = (orc_c32 *)malloc(sizeof(orc_c32) * (size_t)n);
    orc_c32 *bf = (orc_c32 *)malloc(sizeof(orc_c32) * (size_t)n);
    orc_c32 *res = (orc_c32 *)malloc(sizeof(orc_c32) * (size_t)n);

    for (int i = 0; i < n; i++) {                                                /* :264-270 */
        if (i < n / 2) freq[i] = i * (float)samp_rate / (float)n;
        else           freq[i] = i * (float)samp_rate / (float)n - (float)samp_rate;
    }
    for (int l = 0; l < n_rx; l++) {                                             /* :326 */
        orc_c32 *o = out + (size_t)l * n;
        memset(o, 0, sizeof(orc_c32) * (size_t)n);                               /* :339 */
        for (int k = 0; k < n_targets; k++) {                                    /* :342 */
            float doppler = 2 * velocity[k] * center_freq / c_light;             /* :164 */
            float timeshift = (float)((2.0 * range[k] - position_rx[l]
                               * sin(azimuth[k] * M_PI / 180.0)) / c_light);     /* :177 */
            float scale = (float)(c_light * sqrtf(rcs[k]) / FOUR_PI_CUBED_SQRT
                                  / (range[k] * range[k]) / center_freq);        /* :188 */
            float ph = 0.0f;                                                     /* :280 */
            for (int i = 0; i < n; i++) {                                        /* :281-286 */
                filt_dopp[i].re = cosf(ph) * scale; filt_dopp[i].im = sinf(ph) * scale;
                ph = (float)fmod(ph + 2 * M_PI * doppler / (float)samp_rate, 2 * M_PI);
            }
            for (int i = 0; i < n; i++) {                                        /* :297-303 */
                float pt = (float)fmod(2 * M_PI * (timeshift) * (freq[i] + center_freq), 2 * M_PI);
                filt_time[i].re = cosf(pt) / (float)n;       /* exp(-j*pt)/n */
                filt_time[i].im = -sinf(pt) / (float)n;
            }
            for (int i = 0; i < n; i++) bt[i] = cmulf(in[i], filt_dopp[i]);      /* :346 */
            orc_fft_vcc(bt, bf, n, 1, 0);                                        /* :349-350 */
            for (int i = 0; i < n; i++) bf[i] = cmulf(bf[i], filt_time[i]);      /* :353 */
            orc_fft_vcc(bf, res, n, 0, 0);                                       /* :356-357 */
            if (accumulate) for (int i = 0; i < n; i++) { o[i].re += res[i].re; o[i].im += res[i].im; }
            else memcpy(o, res, sizeof(orc_c32) * (size_t)n);                    /* :366 (overwrites) */
        }
        if (self_coupling) {                                                     /* :372-378 */
            float g = (float)pow(10, self_coupling_db / 20.0);
            for (int i = 0; i < n; i++) { o[i].re += g * in[i].re; o[i].im += g * in[i].im; }
        }
    }
    free(freq); free(filt_dopp); free(filt_time); free(bt); free(bf); free(res);
}

/* ------------------------------------------------------------------------ */
/* whole chain: radar -> IFFT -> transpose -> FFT+shift ->
 * {mag^2, estimator}.  Stateless (background removal off, as in the shipped
 * simulation flowgraph, examples/simulation/radar/...radar_sim.grc:1298-1299) */
/* ------------------------------------------------------------------------ */
void orc_chain_batch(const orc_chain_cfg *cfg, const orc_c32 *rx, const orc_c32 *tx,
                     int tx_shared, int n_cpi, int cpi0, float *map_out, orc_c32 *cmap_out,
                     orc_det *dets)
{
    const int N = cfg->fft_len, T = cfg->n_tx, R = cfg->n_rx, V = T * R;
    const int Nr = N * cfg->interp_range, Na = V * cfg->interp_angle;
    const size_t frame = (size_t)(cfg->n_pre + cfg->n_sym) * N;   /* items per port per CPI */
    orc_radar *rad = orc_radar_create(N, T, R, cfg->n_sym, cfg->n_pre, 0, 0, 1,
                                      cfg->interp_range, cfg->tx_interleave);
    orc_c32 *pad = (orc_c32 *)malloc(sizeof(orc_c32) * (size_t)V * Nr);
    orc_c32 *rng = (orc_c32 *)malloc(sizeof(orc_c32) * (size_t)V * Nr);
    orc_c32 *trn = (orc_c32 *)malloc(sizeof(orc_c32) * (size_t)Nr * Na);
    orc_c32 *cmap_local = cmap_out ? NULL : (orc_c32 *)malloc(sizeof(orc_c32) * (size_t)Nr * Na);
    orc_c32 *twr = (orc_c32 *)malloc(sizeof(orc_c32) * (size_t)(Nr / 2 + 1));
    orc_c32 *twa = (orc_c32 *)malloc(sizeof(orc_c32) * (size_t)(Na / 2 + 1));
    orc_c32 *tmp = (orc_c32 *)malloc(sizeof(orc_c32) * (size_t)(Nr > Na ? Nr : Na));
    orc_c32 *tmp2 = (orc_c32 *)malloc(sizeof(orc_c32) * (size_t)(Nr > Na ? Nr : Na));
    const orc_c32 **txp = (const orc_c32 **)malloc(sizeof(void *) * (size_t)T);
    const orc_c32 **rxp = (const orc_c32 **)malloc(sizeof(void *) * (size_t)R);
    if (is_pow2(Nr)) make_twiddles(twr, Nr, 0);
    if (is_pow2(Na)) make_twiddles(twa, Na, 1);

    for (int c = 0; c < n_cpi; c++) {
        for (int t = 0; t < T; t++) txp[t] = tx + ((tx_shared ? 0 : (size_t)c * T) + t) * frame;
        for (int r = 0; r < R; r++) rxp[r] = rx + ((size_t)c * R + r) * frame;
        orc_radar_work(rad, txp, rxp, 0, pad);
        for (int p = 0; p < V; p++)   /* fft_vcc #A: backward, no shift (...radar_sim.grc:940-962) */
            fft_vcc_one(pad + (size_t)p * Nr, rng + (size_t)p * Nr, Nr, 0, 0, twr, tmp, tmp2);
        orc_matrix_transpose(rng, V, Nr, V, cfg->interp_angle, trn);
        orc_c32 *cm = cmap_out ? cmap_out + (size_t)c * Nr * Na : cmap_local;
        for (int n = 0; n < Nr; n++)  /* fft_vcc #B: forward, shift (...radar_sim.grc:963-985) */
            fft_vcc_one(trn + (size_t)n * Na, cm + (size_t)n * Na, Na, 1, 1, twa, tmp, tmp2);
        if (map_out) orc_mag_squared(cm, map_out + (size_t)c * Nr * Na, (size_t)Nr * Na);
        if (dets) {
            orc_range_angle_estimate(cm, Nr, Na, cfg->range_bins, Nr, cfg->angle_bins, Na,
                                     cfg->noise_discard_range_m, cfg->noise_discard_angle_deg,
                                     cfg->snr_threshold, cfg->power_threshold, &dets[c], NULL);
            dets[c].cpi = cpi0 + c;
        }
    }
    free(pad); free(rng); free(trn); free(cmap_local); free(twr); free(twa); free(tmp); free(tmp2);
    free(txp); free(rxp);
    orc_radar_destroy(rad);
}
