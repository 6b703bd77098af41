// jrc_cuda.cu -- host side of libjrc_cuda.so (the C ABI declared in include/jrc_cuda.h).
// Build: gr-mimo-ofdm-jrc_b200/Makefile (nvcc -gencode arch=compute_100a,code=sm_100a).
// There is deliberately no CPU fallback in this file: every entry point either runs
// CUDA kernels or fails with a status code.
#include "jrc_cuda.h"

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "jrc_fused.cuh"
#include "jrc_tiled.cuh"
#include "jrc_slice.cuh"
#include "jrc_wide.cuh"
#ifdef JRC_WITH_TC      // make TC=1: the tensor-core variant of the fused kernel, an experiment that measures slower (profiles/README.md)
#include "jrc_tcfused.cuh"
#endif
#include "jrc_staged.cuh"
#include "jrc_exact.cuh"

using namespace jrc;

static_assert(sizeof(jrc_det) == 32 && sizeof(DetDev) == 32, "detection record is 32 bytes");
static_assert(sizeof(jrc_c32) == sizeof(c32), "complex layout");

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
static thread_local std::string g_err;
static jrc_status fail(jrc_status st, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return st;
}
#define CU(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(JRC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define ST(expr)                                 \
    do {                                         \
        jrc_status s_ = (expr);                  \
        if (s_ != JRC_OK) return s_;             \
    } while (0)

// NVTX ranges around the stages of a call (visible in Nsight Systems; no-ops without a tool attached)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

extern "C" const char *jrc_last_error(void) { return g_err.c_str(); }
extern "C" int32_t jrc_abi_version(void) { return 2; }

// ---------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------
struct GrowBuf {   // grow-only device / pinned-host buffer
    void *p = nullptr;
    size_t cap = 0;
    bool pinned = false;
    jrc_status need(size_t bytes)
    {
        if (bytes <= cap) return JRC_OK;
        if (p) { if (pinned) cudaFreeHost(p); else cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = pinned ? cudaMallocHost(&p, want) : cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return fail(JRC_ERR_CUDA, "allocation of %zu bytes failed: %s", want, cudaGetErrorString(e)); }
        cap = want;
        return JRC_OK;
    }
    void release()
    {
        if (p) { if (pinned) cudaFreeHost(p); else cudaFree(p); }
        p = nullptr; cap = 0;
    }
};

enum { JRC_STREAM_DEPTH = 4 };
struct jrc_stream_state;
struct jrc_fused_state;

struct jrc_chain {
    jrc_chain_cfg cfg;
    jrc_stream_state *sstate = nullptr;         // jrc_chain_submit / jrc_chain_wait slots, created on first use
    jrc_fused_state *fstate = nullptr;          // jrc_radar_estimate_fused ring, created on first use
    int V = 0, Nr = 0, Na = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    // estimator
    bool est_set = false;
    int est_epoch = 0;                          // bumped by every estimator / threshold change (captured graphs bake them in)
    std::vector<float> range_bins, angle_bins;
    float nd_range_m = 0, nd_angle_deg = 0, snr_thr = 0, pow_thr = 0;
    std::map<std::pair<int, int>, double2 *> dft_tabs;    // (n, forward) -> cos/sin table of k_dft_any
    std::map<std::pair<int, int>, c32 *> twiddles_full;   // (n, forward) -> w_n^i, i < n (tiled kernels)
    float *d_angle_bins = nullptr;
    int2 *d_win_tab = nullptr;                  // k_est_tables: per-angle-bin noise window columns
    double2 *d_g_tab = nullptr;                 //               and their closed-form column sums
    // background state (lib/mimo_ofdm_radar_impl.h:48-54)
    c32 *d_ring = nullptr, *d_temp = nullptr;
    int ring_size = 0, ring_head = 0;
    // scratch
    GrowBuf sH, sY, sC, sKeys, sSec, sDet, sIn[2], sMap[2], sDets[2], sMisc, sMisc2, sStage[8];
    GrowBuf sDemod;                              // jrc_chain_run_batch_time: demodulated symbols of one chunk
    GrowBuf sFix, sExact;                        // marked-CPI list (FixCtl + int[n]) and range-spectra scratch of k_est_exact
    int exact_grid = 0;
    GrowBuf pin_a, pin_b;
    std::map<std::pair<int, int>, c32 *> twiddles;   // (n, forward) -> device table
    int last_path = 0;
    int64_t launches = 0;
    int fused_ctas_per_sm = 0;
    float *d_bimg = nullptr;                    // k_fused_tc: angle DFT matrix, tf32 hi/lo, swizzled shared-memory image
    int tc_mode = 0;                            // JRC_TC: 0 off, 1 tensor-core fused kernel where it applies
    std::map<std::pair<const void *, size_t>, int> occ_cache;   // (kernel, dynamic smem) -> resident CTAs per SM, attribute set
    int zero_copy = 1;                          // latency mode: kernel reads/writes pinned host memory directly (JRC_ZEROCOPY=0 disables)
    std::atomic<int> bg_recording{0};           // set_background_record may come from another thread (GUI / RPC callback)
};

static bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
static int ilog2(int n) { int l = 0; while ((1 << l) < n) l++; return l; }

static jrc_status get_twiddles(jrc_chain *h, int n, int forward, const c32 **out)
{
    auto key = std::make_pair(n, forward ? 1 : 0);
    auto it = h->twiddles.find(key);
    if (it != h->twiddles.end()) { *out = it->second; return JRC_OK; }
    // same table as the CPU oracle: cos/sin evaluated in double, rounded to float
    std::vector<c32> tw((size_t)(n / 2 > 0 ? n / 2 : 1));
    const double sgn = forward ? -1.0 : 1.0;
    for (int k = 0; k < n / 2; k++) {
        double a = sgn * 2.0 * M_PI * (double)k / (double)n;
        tw[k].x = (float)cos(a); tw[k].y = (float)sin(a);
    }
    c32 *d = nullptr;
    CU(cudaMalloc(&d, tw.size() * sizeof(c32)));
    CU(cudaMemcpyAsync(d, tw.data(), tw.size() * sizeof(c32), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->twiddles[key] = d;
    *out = d;
    return JRC_OK;
}

extern "C" void jrc_chain_destroy(jrc_chain *h);
static void stream_state_destroy(jrc_chain *h);
static void fused_state_destroy(jrc_chain *h);
struct GrowBuf;
static jrc_status stream_stats(jrc_chain *h, const std::function<jrc_status(const GrowBuf &)> &add);

static jrc_status chain_init(jrc_chain *h, const jrc_chain_cfg *cfg, int Nr, int Na)
{
    h->cfg = *cfg;
    h->bg_recording = cfg->background_recording;
    h->V = cfg->n_tx * cfg->n_rx; h->Nr = Nr; h->Na = Na;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, cfg->device));
    h->sm_count = prop.multiProcessorCount;
    // one stream per handle; the copy streams and events of the chunked host pipeline are created when it first runs
    // (the utility handles of the per-block calls never need them)
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->pin_a.pinned = h->pin_b.pinned = true;
    if (const char *e = getenv("JRC_ZEROCOPY")) h->zero_copy = atoi(e) != 0;
    if (const char *e = getenv("JRC_TC")) h->tc_mode = atoi(e);
    const size_t vn = (size_t)h->V * cfg->fft_len;
    CU(cudaMalloc(&h->d_temp, vn * sizeof(c32)));
    CU(cudaMemsetAsync(h->d_temp, 0, vn * sizeof(c32), h->stream));
    if (cfg->record_len > 0) {
        CU(cudaMalloc(&h->d_ring, vn * sizeof(c32) * (size_t)cfg->record_len));
        CU(cudaMemsetAsync(h->d_ring, 0, vn * sizeof(c32) * (size_t)cfg->record_len, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    return JRC_OK;
}

extern "C" jrc_status jrc_chain_create(const jrc_chain_cfg *cfg, jrc_chain **out)
{
    if (!cfg || !out) return fail(JRC_ERR_INVALID, "null argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(JRC_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(JRC_ERR_INVALID, "device %d out of range (%d devices)", cfg->device, ndev);
    if (cfg->fft_len < 1) return fail(JRC_ERR_INVALID, "fft_len must be positive");
    if (cfg->n_tx < 1 || cfg->n_rx < 1 || cfg->n_sym < 1 || cfg->n_pre < 0) return fail(JRC_ERR_INVALID, "bad antenna/symbol counts");
    if (cfg->interp_range < 1 || cfg->interp_angle < 1) return fail(JRC_ERR_INVALID, "interp factors must be >= 1");
    // (power-of-two sizes are a requirement of the FFT chains, checked where a chain runs: the per-block calls
    //  -- conj-MAC, transpose, estimator, ... -- take any antenna count, like the reference blocks)
    const long long Nr = (long long)cfg->fft_len * cfg->interp_range, Na = (long long)cfg->n_tx * cfg->n_rx * cfg->interp_angle;
    if (Nr > (1 << 24) || Na > (1 << 24)) return fail(JRC_ERR_INVALID, "Nr=%lld / Na=%lld too large", Nr, Na);
    if (cfg->background_removal && cfg->record_len < 0) return fail(JRC_ERR_INVALID, "record_len < 0");

    CU(cudaSetDevice(cfg->device));
    jrc_chain *h = new jrc_chain();
    jrc_status st = chain_init(h, cfg, (int)Nr, (int)Na);
    if (st != JRC_OK) {              // release whatever was created (the message of the failure is kept)
        const std::string msg = g_err;
        jrc_chain_destroy(h);
        g_err = msg;
        return st;
    }
    *out = h;
    return JRC_OK;
}

extern "C" void jrc_chain_destroy(jrc_chain *h)
{
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    stream_state_destroy(h);
    fused_state_destroy(h);
    for (auto &kv : h->twiddles) cudaFree(kv.second);
    GrowBuf *bufs[] = {&h->sH, &h->sY, &h->sC, &h->sKeys, &h->sSec, &h->sFix, &h->sExact, &h->sDet, &h->sIn[0], &h->sIn[1], &h->sMap[0], &h->sMap[1],
                       &h->sDets[0], &h->sDets[1], &h->sMisc, &h->sMisc2, &h->pin_a, &h->pin_b, &h->sDemod};
    for (GrowBuf *b : bufs) b->release();
    for (GrowBuf &b : h->sStage) b.release();
    for (auto &kv : h->twiddles_full) cudaFree(kv.second);
    for (auto &kv : h->dft_tabs) cudaFree(kv.second);
    if (h->d_angle_bins) cudaFree(h->d_angle_bins);
    if (h->d_win_tab) cudaFree(h->d_win_tab);
    if (h->d_g_tab) cudaFree(h->d_g_tab);
    if (h->d_bimg) cudaFree(h->d_bimg);
    if (h->d_ring) cudaFree(h->d_ring);
    if (h->d_temp) cudaFree(h->d_temp);
    for (int i = 0; i < 2; i++) {
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_comp[i]) cudaEventDestroy(h->ev_comp[i]);
        if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
    }
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
    if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
    delete h;
}

extern "C" void *jrc_chain_stream(jrc_chain *h) { return h ? (void *)h->stream : nullptr; }
extern "C" jrc_status jrc_chain_sync(jrc_chain *h)
{
    if (!h) return fail(JRC_ERR_INVALID, "null handle");
    CU(cudaStreamSynchronize(h->stream));
    return JRC_OK;
}
extern "C" int32_t jrc_chain_last_path(const jrc_chain *h) { return h ? h->last_path : 0; }
extern "C" int64_t jrc_chain_launch_count(const jrc_chain *h) { return h ? h->launches : 0; }

// statistics of the reference-order pass (csrc/jrc_exact.cuh) since the handle was created:
// out[0] records marked by the fast kernels, out[1] records redone by k_est_exact, out[2] arg-max ties resolved inside
// the fused kernel.  Synchronises the handle's stream.
extern "C" jrc_status jrc_chain_exact_stats(jrc_chain *h, int64_t *out)
{
    if (!h || !out) return fail(JRC_ERR_INVALID, "null argument");
    out[0] = out[1] = out[2] = 0;
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->stream));
    auto add = [&](const GrowBuf &b) -> jrc_status {
        if (!b.p) return JRC_OK;
        FixCtl c;
        CU(cudaMemcpy(&c, b.p, sizeof(c), cudaMemcpyDeviceToHost));
        out[0] += c.n_marked; out[1] += c.n_redone; out[2] += c.n_inkernel;
        return JRC_OK;
    };
    ST(add(h->sFix));
    ST(stream_stats(h, add));
    return JRC_OK;
}

extern "C" jrc_status jrc_chain_set_estimator(jrc_chain *h, const float *range_bins, int32_t n_range,
                                               const float *angle_bins, int32_t n_angle,
                                               float noise_discard_range_m, float noise_discard_angle_deg,
                                               float snr_threshold, float power_threshold)
{
    if (!h || !range_bins || !angle_bins) return fail(JRC_ERR_INVALID, "null argument");
    if (n_range < 2 || n_angle < 2) return fail(JRC_ERR_INVALID, "need at least two range and two angle bins");
    CU(cudaSetDevice(h->cfg.device));
    h->range_bins.assign(range_bins, range_bins + n_range);
    h->angle_bins.assign(angle_bins, angle_bins + n_angle);
    h->nd_range_m = noise_discard_range_m; h->nd_angle_deg = noise_discard_angle_deg;
    h->snr_thr = snr_threshold; h->pow_thr = power_threshold;
    h->est_epoch++;
    if (h->d_angle_bins) { cudaFree(h->d_angle_bins); h->d_angle_bins = nullptr; }
    CU(cudaMalloc(&h->d_angle_bins, sizeof(float) * (size_t)n_angle));
    CU(cudaMemcpyAsync(h->d_angle_bins, angle_bins, sizeof(float) * (size_t)n_angle, cudaMemcpyHostToDevice, h->stream));
    if (h->d_win_tab) { cudaFree(h->d_win_tab); h->d_win_tab = nullptr; }
    if (h->d_g_tab) { cudaFree(h->d_g_tab); h->d_g_tab = nullptr; }
    CU(cudaMalloc(&h->d_win_tab, sizeof(int2) * (size_t)n_angle));
    CU(cudaMalloc(&h->d_g_tab, sizeof(double2) * 8 * (size_t)n_angle));
    {
        EstParams TP;
        memset(&TP, 0, sizeof(TP));
        TP.angle_bins = h->d_angle_bins; TP.n_angle = n_angle; TP.noise_discard_angle_deg = noise_discard_angle_deg;
        k_est_tables<<<(n_angle + 127) / 128, 128, 0, h->stream>>>(TP, h->d_win_tab, h->d_g_tab);
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(h->stream));
    h->est_set = true;
    return JRC_OK;
}

extern "C" jrc_status jrc_chain_set_thresholds(jrc_chain *h, float snr_threshold, float power_threshold)
{
    if (!h) return fail(JRC_ERR_INVALID, "null handle");
    h->snr_thr = snr_threshold; h->pow_thr = power_threshold;
    h->est_epoch++;
    return JRC_OK;
}
extern "C" jrc_status jrc_chain_set_background_record(jrc_chain *h, int32_t on)
{
    if (!h) return fail(JRC_ERR_INVALID, "null handle");
    h->bg_recording.store(on ? 1 : 0);      // read once per batch by launch_chan_est
    return JRC_OK;
}
extern "C" jrc_status jrc_chain_reset_background(jrc_chain *h)
{
    if (!h) return fail(JRC_ERR_INVALID, "null handle");
    CU(cudaSetDevice(h->cfg.device));
    h->ring_size = 0; h->ring_head = 0;
    CU(cudaMemsetAsync(h->d_temp, 0, (size_t)h->V * h->cfg.fft_len * sizeof(c32), h->stream));
    return JRC_OK;
}

static jrc_status est_params(const jrc_chain *h, int n_range, int n_angle, EstParams *P)
{
    if (!h->est_set) return fail(JRC_ERR_STATE, "jrc_chain_set_estimator has not been called");
    if ((int)h->range_bins.size() != n_range || (int)h->angle_bins.size() != n_angle)
        return fail(JRC_ERR_INVALID, "estimator bins (%zu x %zu) do not match the map (%d x %d)",
                    h->range_bins.size(), h->angle_bins.size(), n_range, n_angle);
    P->angle_bins = h->d_angle_bins;
    P->n_angle = n_angle; P->n_range = n_range;
    // lib/range_angle_estimator_impl.cc:189 -- float division, truncation to int
    P->discard_range_idx = (int)(h->nd_range_m / (h->range_bins[1] - h->range_bins[0]));
    P->noise_discard_angle_deg = h->nd_angle_deg;
    P->snr_threshold = h->snr_thr; P->power_threshold = h->pow_thr;
    return JRC_OK;
}

static int grid_for(long long n, int threads, int sm_count)
{
    long long b = (n + threads - 1) / threads;
    long long cap = (long long)sm_count * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// ---------------------------------------------------------------------------
// staged kernels: launch helpers (all on h->stream)
// ---------------------------------------------------------------------------
static jrc_status launch_fft_rows(jrc_chain *h, const c32 *in, long long in_stride, int n_in, c32 *out, int n,
                                  long long rows, int forward, int shift, int tr_w = 0, unsigned long long *keys = nullptr)
{
    if (!is_pow2(n) || n > 16384) return fail(JRC_ERR_INVALID, "FFT length %d unsupported (power of two <= 16384)", n);
    if (rows <= 0) return JRC_OK;
    const c32 *tw = nullptr;
    ST(get_twiddles(h, n, forward, &tw));
    int rpc = n >= 512 ? 1 : 512 / n;
    int threads = 256;
    size_t smem = (size_t)rpc * n * sizeof(c32);
    if (smem > 48 * 1024)
        CU(cudaFuncSetAttribute(k_fft_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long ctas = (rows + rpc - 1) / rpc;
    if (ctas > 0x7fffffffLL) return fail(JRC_ERR_INVALID, "too many FFT rows");
    k_fft_rows<<<(unsigned)ctas, threads, smem, h->stream>>>(in, in_stride, n_in, out, n, ilog2(n), rows, rpc, forward, shift, tw, tr_w, keys);
    CU(cudaGetLastError());
    h->launches++;
    return JRC_OK;
}

static jrc_status launch_transpose(jrc_chain *h, const c32 *in, c32 *out, int K, int L, int W, int mats)
{
    if (mats <= 0) return JRC_OK;
    int kmax = K > W ? K : W;
    dim3 grid((L + 31) / 32, (kmax + 31) / 32, mats), block(32, 8);
    k_transpose_pad<<<grid, block, 0, h->stream>>>(in, out, K, L, W);
    CU(cudaGetLastError());
    h->launches++;
    return JRC_OK;
}

// ---------------------------------------------------------------------------
// tiled path (jrc_tiled.cuh): k_chan_est -> k_fft8_rows -> k_angle_mag -> k_map_finalize
// ---------------------------------------------------------------------------
static jrc_status get_twiddles_full(jrc_chain *h, int n, int forward, const c32 **out)
{
    auto key = std::make_pair(n, forward ? 1 : 0);
    auto it = h->twiddles_full.find(key);
    if (it != h->twiddles_full.end()) { *out = it->second; return JRC_OK; }
    std::vector<c32> tw((size_t)n);
    const double sgn = forward ? -1.0 : 1.0;
    for (int k = 0; k < n; k++) {
        double a = sgn * 2.0 * M_PI * (double)k / (double)n;
        tw[k].x = (float)cos(a); tw[k].y = (float)sin(a);
    }
    c32 *d = nullptr;
    CU(cudaMalloc(&d, tw.size() * sizeof(c32)));
    CU(cudaMemcpyAsync(d, tw.data(), tw.size() * sizeof(c32), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->twiddles_full[key] = d;
    *out = d;
    return JRC_OK;
}

template <int LOG2N>
static jrc_status launch_fft8_rows_t(jrc_chain *h, const c32 *in, long long in_stride, int n_in, c32 *out, long long rows,
                                     const c32 *tw)
{
    using Gm = TiledGeom<LOG2N>;
    auto kern = (n_in <= Gm::N / 8) ? k_fft8_rows<LOG2N, 1, true> : k_fft8_rows<LOG2N, 1, false>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Gm::SMEM));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Gm::THREADS, Gm::SMEM));
    if (per_sm < 1) return fail(JRC_ERR_INVALID, "range FFT kernel does not fit");
    long long grid = (rows + Gm::RPC - 1) / Gm::RPC, cap = (long long)h->sm_count * per_sm;
    if (grid > cap) grid = cap;
    kern<<<(unsigned)grid, Gm::THREADS, Gm::SMEM, h->stream>>>(in, in_stride, n_in, out, rows, tw);
    CU(cudaGetLastError());
    h->launches++;
    return JRC_OK;
}

template <int LOG2NA>
static jrc_status launch_angle_mag_t(jrc_chain *h, const c32 *Y, int V, int Nr, int n_cpi, float *map,
                                     unsigned long long *keys, unsigned *sec, const c32 *tw)
{
    using Gm = TiledGeom<LOG2NA>;
    auto kern = (V <= Gm::N / 8) ? k_angle_mag<LOG2NA, true> : k_angle_mag<LOG2NA, false>;
    const size_t smem = 2 * Gm::SMEM + (size_t)DifTwS<LOG2NA>::table_entries() * sizeof(c32);   // double-buffered tiles + twiddles
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
    if (per_sm < 1) return fail(JRC_ERR_INVALID, "angle kernel does not fit");
    long long grid = (long long)n_cpi * (Nr / Gm::RPC), cap = (long long)h->sm_count * per_sm;
    if (grid > cap) grid = cap;
    kern<<<(unsigned)grid, 256, smem, h->stream>>>(Y, V, Nr, ilog2(Nr / Gm::RPC), n_cpi, map, keys, sec, tw);
    CU(cudaGetLastError());
    h->launches++;
    return JRC_OK;
}

static bool tiled_config_ok(const jrc_chain *h)
{
    const int lr = ilog2(h->Nr), la = ilog2(h->Na);
    if (!is_pow2(h->Nr) || !is_pow2(h->Na) || lr < 6 || lr > 13 || la < 6 || la > 11) return false;
    if (h->V > h->Na || h->cfg.fft_len > h->Nr) return false;
    const int rpc = 256 / (h->Na / 8);     // range bins per angle tile
    return h->Nr % rpc == 0;
}

// k_slice256 (jrc_slice.cuh): range IFFT + transpose + angle FFT + |.|^2 + arg-max in one kernel
static bool slice_config_ok(const jrc_chain *h)
{
    static const bool off = getenv("JRC_NO_SLICE") != nullptr;     // A/B switch for measurements
    return !off && h->cfg.fft_len == 256 && h->Na == 256 && h->V <= 32 && h->cfg.interp_range >= 2 && h->cfg.interp_range % 2 == 0 &&
           h->Nr <= 16384;
}

static jrc_status launch_slice(jrc_chain *h, const c32 *H, int n_cpi, float *map, unsigned long long *keys, unsigned *sec)
{
    SliceParams P;
    memset(&P, 0, sizeof(P));
    P.H = H; P.V = h->V; P.IR = h->cfg.interp_range; P.n_cpi = n_cpi; P.map = map; P.keys = keys; P.sec = sec;
    ST(get_twiddles_full(h, h->Nr, 0, &P.tw_range));
    ST(get_twiddles_full(h, 256, 1, &P.tw256));
    CU(cudaFuncSetAttribute(k_slice256, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SliceGeom::SMEM));
    long long grid = (long long)n_cpi * (P.IR / 2);
    if (grid > h->sm_count) grid = h->sm_count;
    k_slice256<<<(unsigned)grid, SliceGeom::THREADS, SliceGeom::SMEM, h->stream>>>(P);
    CU(cudaGetLastError());
    h->launches++;
    return JRC_OK;
}

// k_wide_mac_angle + k_wide_range_mag (jrc_wide.cuh): 128 virtual channels x 2048 subcarriers without zero-pads
static bool wide_config_ok(const jrc_chain *h)
{
    static const bool off = getenv("JRC_NO_WIDE") != nullptr;      // A/B switch for measurements
    const jrc_chain_cfg &c = h->cfg;
    return !off && c.fft_len == 2048 && c.interp_range == 1 && c.interp_angle == 1 && h->V == 128 && c.n_tx % 2 == 0 &&
           c.n_rx % 4 == 0 && c.n_tx <= 8 && c.n_rx <= 16 && c.n_sym <= 8;
}

// 4-D tensor map over the packets of one port: (float index within a symbol, symbol, antenna, CPI); a tile is
// [n_ant][n_sym][KB subcarriers].  The encoder comes from the driver at run time (no link dependency on libcuda).
typedef CUresult (*tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_fn tmap_encoder()
{
    static tmap_encode_fn fn = []() -> tmap_encode_fn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (tmap_encode_fn)p;
    }();
    return fn;
}
static bool port_tensor_map(CUtensorMap *tm, const PortDev &port, int n_pre, int N, int n_sym, int n_ant, int n_cpi, int KB)
{
    tmap_encode_fn enc = tmap_encoder();
    const long long cs = port.cpi_stride ? port.cpi_stride : (long long)n_ant * port.ant_stride;    // shared frame: one "CPI"
    if (!enc || ((uintptr_t)port.base & 15u) || (port.ant_stride & 1) || (cs & 1) || port.ant_stride <= 0 || cs <= 0) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)2 * N, (cuuint64_t)n_sym, (cuuint64_t)n_ant, (cuuint64_t)(port.cpi_stride ? n_cpi : 1)};
    const cuuint64_t strides[3] = {(cuuint64_t)N * sizeof(c32), (cuuint64_t)port.ant_stride * sizeof(c32), (cuuint64_t)cs * sizeof(c32)};
    const cuuint32_t box[4] = {(cuuint32_t)2 * KB, (cuuint32_t)n_sym, (cuuint32_t)n_ant, 1}, es[4] = {1, 1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)(port.base + (long long)n_pre * N), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static jrc_status launch_wide(jrc_chain *h, PortDev rx, PortDev tx, const c32 *H, int n_pre, int n_cpi, c32 *G, float *map,
                              unsigned long long *keys, unsigned *sec)
{
    using Gm = WideGeom<11>;
    const jrc_chain_cfg &c = h->cfg;
    WideParams P;
    memset(&P, 0, sizeof(P));
    P.rx = rx; P.tx = tx; P.H = H; P.T = c.n_tx; P.R = c.n_rx; P.S = c.n_sym; P.n_pre = n_pre; P.tx_interleave = c.tx_interleave;
    P.n_cpi = n_cpi; P.G = G; P.map = map; P.keys = keys; P.sec = sec;
    ST(get_twiddles_full(h, 128, 1, &P.tw_a));
    ST(get_twiddles_full(h, 2048, 0, &P.tw_r));
    static const bool tma_off = getenv("JRC_WIDE_TMA") && atoi(getenv("JRC_WIDE_TMA")) == 0;      // A/B switch
    CUtensorMap tm_rx, tm_tx;
    memset(&tm_rx, 0, sizeof(tm_rx));
    memset(&tm_tx, 0, sizeof(tm_tx));
    P.use_tma = !tma_off && !H && port_tensor_map(&tm_rx, rx, n_pre, Gm::N, c.n_sym, c.n_rx, n_cpi, Gm::KB) &&
                port_tensor_map(&tm_tx, tx, n_pre, Gm::N, c.n_sym, c.n_tx, n_cpi, Gm::KB);
    // G [n_cpi][128][N] as (float index within a row, angle bin, CPI); a store tile is [128][16 subcarriers], 128-byte swizzle
    CUtensorMap tm_g;
    memset(&tm_g, 0, sizeof(tm_g));
    static const bool tmas_off = getenv("JRC_WIDE_TMA_STORE") && atoi(getenv("JRC_WIDE_TMA_STORE")) == 0;
    // (only with TMA loads or no loads at all: a cp.async prefetch by all threads could not wait for the store's reads)
    if (!tma_off && !tmas_off && (P.use_tma || H) && tmap_encoder() && !((uintptr_t)G & 127u)) {
        const cuuint64_t dims[3] = {(cuuint64_t)2 * Gm::N, (cuuint64_t)Gm::V, (cuuint64_t)n_cpi};
        const cuuint64_t strides[2] = {(cuuint64_t)Gm::N * sizeof(c32), (cuuint64_t)Gm::V * Gm::N * sizeof(c32)};
        const cuuint32_t box[3] = {32, (cuuint32_t)Gm::V, 1}, es[3] = {1, 1, 1};
        P.use_tma_store = tmap_encoder()(&tm_g, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)G, dims, strides, box, es,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    auto ka = c.n_sym == 8 ? k_wide_mac_angle<11, 8> : (c.n_sym == 4 ? k_wide_mac_angle<11, 4> : k_wide_mac_angle<11, 0>);
    auto kb = k_wide_range_mag<11>;
    CU(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Gm::SMEM_A));
    CU(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Gm::SMEM_B));
    long long ua = (long long)n_cpi * (Gm::N / Gm::KB), ub = (long long)n_cpi * (Gm::V / Gm::UB);
    const long long cap_a = (long long)Gm::CTAS_A * h->sm_count;
    long long ga = ua < cap_a ? ua : cap_a, gb = ub < 2LL * h->sm_count ? ub : 2LL * h->sm_count;
    ka<<<(unsigned)ga, Gm::TA, Gm::SMEM_A, h->stream>>>(P, tm_rx, tm_tx, tm_g);
    CU(cudaGetLastError());
    kb<<<(unsigned)gb, Gm::GR::THREADS, Gm::SMEM_B, h->stream>>>(P);
    CU(cudaGetLastError());
    h->launches += 2;
    return JRC_OK;
}

static jrc_status launch_fft8_rows(jrc_chain *h, const c32 *in, long long in_stride, int n_in, c32 *out, int n, long long rows)
{
    const c32 *tw = nullptr;
    ST(get_twiddles_full(h, n, 0, &tw));
    switch (ilog2(n)) {
#define JRC_CASE(l) case l: return launch_fft8_rows_t<l>(h, in, in_stride, n_in, out, rows, tw);
        JRC_CASE(6) JRC_CASE(7) JRC_CASE(8) JRC_CASE(9) JRC_CASE(10) JRC_CASE(11) JRC_CASE(12) JRC_CASE(13)
#undef JRC_CASE
    }
    return fail(JRC_ERR_INVALID, "range FFT length %d unsupported by the tiled path", n);
}

static jrc_status launch_angle_mag(jrc_chain *h, const c32 *Y, int V, int Nr, int Na, int n_cpi, float *map,
                                   unsigned long long *keys, unsigned *sec)
{
    const c32 *tw = nullptr;
    ST(get_twiddles_full(h, Na, 1, &tw));
    switch (ilog2(Na)) {
#define JRC_CASE(l) case l: return launch_angle_mag_t<l>(h, Y, V, Nr, n_cpi, map, keys, sec, tw);
        JRC_CASE(6) JRC_CASE(7) JRC_CASE(8) JRC_CASE(9) JRC_CASE(10) JRC_CASE(11)
#undef JRC_CASE
    }
    return fail(JRC_ERR_INVALID, "angle FFT length %d unsupported by the tiled path", Na);
}

// ready_keys: the arg-max keys are already there (k_fft_rows' epilogue, a buffer of the caller's); k_est_finalize zeroes
// them again for the next frame (the caller zeroed them once, before the first)
static jrc_status launch_estimate(jrc_chain *h, const c32 *cmap, int n_inputs, int vlen, int mats, int cpi0, DetDev *dets,
                                  unsigned long long *ready_keys = nullptr)
{
    EstParams P;
    ST(est_params(h, n_inputs, vlen, &P));
    const bool have_keys = ready_keys != nullptr;
    if (!have_keys) ST(h->sKeys.need(sizeof(unsigned long long) * (size_t)mats));
    unsigned long long *keys = have_keys ? ready_keys : (unsigned long long *)h->sKeys.p;
    if (!have_keys) CU(cudaMemsetAsync(keys, 0, sizeof(unsigned long long) * (size_t)mats, h->stream));
    long long per = (long long)n_inputs * vlen;
    int bx = (int)((per + 256 * 8 - 1) / (256 * 8));
    if (bx < 1) bx = 1;
    if (bx > 256) bx = 256;
    for (int m0 = 0; m0 < mats && !have_keys; m0 += 65535) {
        int mc = mats - m0 < 65535 ? mats - m0 : 65535;
        k_est_argmax<<<dim3(bx, mc), 256, 0, h->stream>>>(cmap + (long long)m0 * per, per, keys + m0);
        CU(cudaGetLastError());
        h->launches++;
    }
    k_est_finalize<<<mats, 256, 0, h->stream>>>(cmap, per, n_inputs, vlen, keys, P, dets, cpi0, have_keys ? 1 : 0);
    CU(cudaGetLastError());
    h->launches++;
    return JRC_OK;
}

// channel estimates for a batch -> d_H [n_cpi][V][N]  (+ background ring update)
static jrc_status launch_chan_est(jrc_chain *h, PortDev rx, PortDev tx, int n_cpi, c32 *d_H, int n_pre)
{
    const jrc_chain_cfg &c = h->cfg;
    const int recording = h->bg_recording.load();
    long long total = (long long)n_cpi * h->V * c.fft_len;
    if (c.n_tx % 4 == 0 && c.n_rx % 4 == 0)     // large arrays: 4 x 4 antenna blocks per thread
        k_chan_est_tile<4, 4><<<grid_for(total / 16, 256, h->sm_count), 256, 0, h->stream>>>(
            rx, tx, n_cpi, c.fft_len, c.n_tx, c.n_rx, c.n_sym, n_pre, c.tx_interleave, d_H);
    else
        k_chan_est<<<grid_for(total, 256, h->sm_count), 256, 0, h->stream>>>(rx, tx, n_cpi, c.fft_len, c.n_tx, c.n_rx, c.n_sym,
                                                                           n_pre, c.tx_interleave, d_H);
    CU(cudaGetLastError());
    h->launches++;
    if (c.background_removal || recording) {
        int VN = h->V * c.fft_len;
        k_background<<<(VN + 127) / 128, 128, 0, h->stream>>>(d_H, n_cpi, VN, h->d_ring, h->d_temp, c.record_len,
                                                            h->ring_size, h->ring_head, recording,
                                                            c.background_removal);
        CU(cudaGetLastError());
        h->launches++;
        if (c.background_removal && c.record_len > 0) {   // mirror the device-side ring evolution
            for (int i = 0; i < n_cpi; i++) {
                if (h->ring_size < c.record_len) h->ring_size++;
                else h->ring_head = (h->ring_head + 1) % c.record_len;
            }
        }
    }
    return JRC_OK;
}

// ---------------------------------------------------------------------------
// fused path dispatch
// ---------------------------------------------------------------------------
template <int IR, int IA, bool FROM_H>
static jrc_status launch_fused_t(jrc_chain *h, const FusedParams &P, bool *supported)
{
    using Gm = FusedGeom<IR, IA>;
    size_t smem = Gm::smem_bytes(P.T, P.R, P.S, FROM_H);
    auto kern = P.map ? k_fused64x8<IR, IA, FROM_H, true> : k_fused64x8<IR, IA, FROM_H, false>;
    int per_sm = 0;
    const auto ck = std::make_pair((const void *)kern, smem);
    auto it = h->occ_cache.find(ck);
    if (it != h->occ_cache.end()) per_sm = it->second;
    else {
        int dev_smem = 0;
        CU(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->cfg.device));
        if (smem <= (size_t)dev_smem) {     // (else: many LTF symbols, the symbol buffer does not fit next to the spectra)
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
        }
        h->occ_cache[ck] = per_sm;
    }
    if (per_sm < 1) { *supported = false; return JRC_OK; }
    h->fused_ctas_per_sm = per_sm;
    long long grid = (long long)h->sm_count * per_sm;
    if (grid > P.n_cpi) grid = P.n_cpi;
    kern<<<(unsigned)grid, 256, smem, h->stream>>>(P);
    CU(cudaGetLastError());
    h->launches++;
    return JRC_OK;
}

template <bool FROM_H>
static jrc_status launch_fused(jrc_chain *h, const FusedParams &P, bool *supported)
{
    const int IR = h->cfg.interp_range, IA = h->cfg.interp_angle;
    *supported = true;
#define JRC_FUSED_CASE(ir, ia) if (IR == ir && IA == ia) return launch_fused_t<ir, ia, FROM_H>(h, P, supported);
    JRC_FUSED_CASE(8, 16)    // shipped flowgraph: 512 x 128 map
    JRC_FUSED_CASE(16, 8)    // BASELINE configs[1]: 1024 x 64 map
    JRC_FUSED_CASE(8, 8)
    JRC_FUSED_CASE(16, 16)
#undef JRC_FUSED_CASE
    *supported = false;
    return JRC_OK;
}

// ---------------------------------------------------------------------------
// k_est_exact (jrc_exact.cuh): redoes the marked records of a batch in the staged arithmetic
// ---------------------------------------------------------------------------
static jrc_status fix_buffers(jrc_chain *h, int n_cpi, FixCtl **ctl, int **list)
{
    const size_t need = 32 + sizeof(int) * (size_t)n_cpi;
    if (need > h->sFix.cap) {
        ST(h->sFix.need(need));
        CU(cudaMemsetAsync(h->sFix.p, 0, sizeof(FixCtl), h->stream));      // k_est_exact re-arms it after every batch
    }
    *ctl = (FixCtl *)h->sFix.p;
    *list = (int *)((char *)h->sFix.p + 32);
    return JRC_OK;
}

static jrc_status launch_exact(jrc_chain *h, PortDev rx, PortDev tx, const c32 *H, int n_pre, int cpi0, const float *map,
                               DetDev *dets, const EstParams &est)
{
    const jrc_chain_cfg &c = h->cfg;
    ExactParams P;
    memset(&P, 0, sizeof(P));
    P.rx = rx; P.tx = tx; P.H = H;
    P.N = c.fft_len; P.T = c.n_tx; P.R = c.n_rx; P.S = c.n_sym; P.n_pre = n_pre; P.tx_interleave = c.tx_interleave;
    P.V = h->V; P.Nr = h->Nr; P.Na = h->Na; P.log2Nr = ilog2(h->Nr); P.log2Na = ilog2(h->Na);
    ST(get_twiddles(h, h->Nr, 0, &P.tw_r));
    ST(get_twiddles(h, h->Na, 1, &P.tw_a));
    P.est = est; P.dets = dets; P.map = map; P.cpi0 = cpi0;
    P.ctl = (FixCtl *)h->sFix.p;
    P.list = (const int *)((char *)h->sFix.p + 32);
    const int big = h->Nr > h->Na ? h->Nr : h->Na;
    P.buf_elems = big > 8192 ? big : 8192;
    const size_t smem = (size_t)P.buf_elems * sizeof(c32);
    // large maps: eight CTAs per marked CPI (a thread-block cluster), sixteen clusters; small ones: one CTA each
    const int cl = ((size_t)h->V * h->Nr * sizeof(c32) >= ((size_t)256 << 10)) ? 8 : 1;
    if (!h->exact_grid) {
        CU(cudaFuncSetAttribute(k_est_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->exact_grid = cl > 1 ? 16 * cl : 32;
    }
    P.scratch_stride = (size_t)h->V * h->Nr + EXACT_WCAP / 2;
    ST(h->sExact.need((size_t)(h->exact_grid / cl) * P.scratch_stride * sizeof(c32)));
    P.scratch = (c32 *)h->sExact.p;
    // programmatic dependent launch: the grid is staged while the kernel in front of it drains (it waits for that
    // kernel's memory with griddepcontrol.wait), so an empty marked-CPI list costs no launch gap
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = dim3((unsigned)h->exact_grid); lc.blockDim = dim3(256); lc.dynamicSmemBytes = smem; lc.stream = h->stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = (unsigned)cl; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 2;
    CU(cudaLaunchKernelEx(&lc, k_est_exact, P));
    h->launches++;
    return JRC_OK;
}

#ifdef JRC_WITH_TC
// ---------------------------------------------------------------------------
// tensor-core fused kernel (jrc_tcfused.cuh)
// ---------------------------------------------------------------------------
static float tf32_hi(float x)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}

static jrc_status tc_bimg(jrc_chain *h)
{
    if (h->d_bimg) return JRC_OK;
    const int NA = 64, N = 2 * NA;
    std::vector<float> img((size_t)N * 32, 0.f);
    for (int i = 0; i < NA; i++)
        for (int p = 0; p < 8; p++) {
            // D[p][i] = (-1)^p e^{-j 2 pi p i / NA}: the angle DFT with its output fftshift folded in
            const double a = -2.0 * M_PI * (double)((p * i) % NA) / (double)NA, sg = (p & 1) ? -1.0 : 1.0;
            const double dr = sg * cos(a), di = sg * sin(a);
            const double rows[2][2] = {{dr, -di}, {di, dr}};   // column i (Re): (yr, yi) -> dr, -di;  column NA + i (Im): di, dr
            for (int ri = 0; ri < 2; ri++)
                for (int c = 0; c < 2; c++) {
                    const int j = ri * NA + i, k = 2 * p + c;
                    const float full = (float)rows[ri][c], hi = tf32_hi(full), lo = tf32_hi(full - hi);
                    auto at = [&](int kk) { return (size_t)(j >> 3) * 256 + (j & 7) * 32 + ((((kk >> 2) ^ (j & 7)) << 2) + (kk & 3)); };
                    img[at(k)] = hi;
                    img[at(16 + k)] = lo;
                }
        }
    CU(cudaMalloc(&h->d_bimg, img.size() * sizeof(float)));
    CU(cudaMemcpyAsync(h->d_bimg, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return JRC_OK;
}

static bool tc_config_ok(const jrc_chain *h)
{
    return h->tc_mode && h->cfg.fft_len == 64 && h->V == 8 && h->cfg.interp_angle == 8 &&
           (h->cfg.interp_range == 16 || h->cfg.interp_range == 8);
}

template <int IR>
static jrc_status launch_tc_t(jrc_chain *h, const TcFusedParams &P)
{
    using Gm = TcFusedGeom<IR>;
    const size_t smem = Gm::smem_bytes(P.T, P.R, P.S);
    CU(cudaFuncSetAttribute(k_fused_tc<IR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = h->sm_count < P.n_cpi ? h->sm_count : P.n_cpi;
    k_fused_tc<IR><<<grid, Gm::THREADS, smem, h->stream>>>(P);
    CU(cudaGetLastError());
    h->launches++;
    return JRC_OK;
}

#endif

static bool fused_config_ok(const jrc_chain *h)
{
    const int IR = h->cfg.interp_range, IA = h->cfg.interp_angle;
    if (h->cfg.fft_len != 64 || h->V != 8) return false;
    return (IR == 8 && IA == 16) || (IR == 16 && IA == 8) || (IR == 8 && IA == 8) || (IR == 16 && IA == 16);
}

static bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

// n_pre: symbols to skip in front of every antenna row (cfg.n_pre for GNU Radio packets, 0 for the packed host layout).
// defer_exact: do not launch k_est_exact; the caller looks at the DET_PENDING marks itself (latency path).
static jrc_status run_batch_impl(jrc_chain *h, jrc_port_layout rx, jrc_port_layout tx, int32_t n_cpi, int32_t cpi0,
                                 float *map, jrc_c32 *cmap, jrc_det *dets, int32_t path, int n_pre, bool defer_exact)
{
    if (!h) return fail(JRC_ERR_INVALID, "null handle");
    if (n_cpi < 0) return fail(JRC_ERR_INVALID, "n_cpi < 0");
    if (n_cpi == 0) return JRC_OK;
    if (!rx.base || !tx.base) return fail(JRC_ERR_INVALID, "null input pointer");
    if (dets && !h->est_set) return fail(JRC_ERR_STATE, "dets requested but jrc_chain_set_estimator has not been called");
    CU(cudaSetDevice(h->cfg.device));
    const jrc_chain_cfg &c = h->cfg;
    const int N = c.fft_len, V = h->V, Nr = h->Nr, Na = h->Na;
    if (!is_pow2(N) || !is_pow2(Nr) || !is_pow2(Na) || Nr > 16384 || Na > 16384)
        return fail(JRC_ERR_INVALID, "the FFT chain needs power-of-two fft_len / Nr=%d / Na=%d <= 16384", Nr, Na);
    PortDev drx{(const c32 *)rx.base, rx.cpi_stride, rx.ant_stride};
    PortDev dtx{(const c32 *)tx.base, tx.cpi_stride, tx.ant_stride};
    const bool bg = c.background_removal != 0;
    const bool recording = h->bg_recording.load() != 0;
    NvtxRange nv_batch("jrc_chain_run_batch");

    bool want_fused = (path == JRC_PATH_AUTO || path == JRC_PATH_FUSED) && fused_config_ok(h) && cmap == nullptr;
    if (want_fused && !bg) {
        // cp.async moves 16-byte chunks: every antenna row must start 16-byte aligned
        bool al = aligned16(rx.base) && aligned16(tx.base) && (rx.cpi_stride % 2 == 0) && (rx.ant_stride % 2 == 0) &&
                  (tx.cpi_stride % 2 == 0) && (tx.ant_stride % 2 == 0);
        if (!al) want_fused = false;
    }
    if (map && !aligned16(map)) want_fused = false;
    if (path == JRC_PATH_FUSED && !want_fused)
        return fail(JRC_ERR_INVALID, "fused path requested but this configuration/layout has no fused kernel");

    EstParams EP;
    FixCtl *fix_ctl = nullptr;
    int *fix_list = nullptr;
    if (dets) {
        ST(est_params(h, Nr, Na, &EP));
        ST(fix_buffers(h, n_cpi, &fix_ctl, &fix_list));
    }

#ifdef JRC_WITH_TC
    if (want_fused && map && !dets && tc_config_ok(h) && n_cpi >= 32 && !bg && !recording &&
        (size_t)(c.n_tx + c.n_rx) * c.n_sym * 64 * 8 <= 15360) {
        NvtxRange nv("fused: k_fused_tc");
        ST(tc_bimg(h));
        TcFusedParams TP;
        memset(&TP, 0, sizeof(TP));
        TP.rx = drx; TP.tx = dtx; TP.n_cpi = n_cpi; TP.cpi0 = cpi0;
        TP.T = c.n_tx; TP.R = c.n_rx; TP.S = c.n_sym; TP.n_pre = n_pre; TP.tx_interleave = c.tx_interleave;
        TP.map = map; TP.bimg = h->d_bimg;
        if (const char *e = getenv("JRC_TC_DBG")) TP.dbg = atoi(e);
        if (c.interp_range == 16) ST(launch_tc_t<16>(h, TP)); else ST(launch_tc_t<8>(h, TP));
        h->last_path = JRC_PATH_FUSED;
        return JRC_OK;
    }
#endif
    if (want_fused) {
        NvtxRange nv("fused: k_fused64x8");
        FusedParams P;
        memset(&P, 0, sizeof(P));
        P.rx = drx; P.tx = dtx; P.n_cpi = n_cpi; P.cpi0 = cpi0;
        P.T = c.n_tx; P.R = c.n_rx; P.S = c.n_sym; P.n_pre = n_pre; P.tx_interleave = c.tx_interleave;
        P.map = map; P.dets = (DetDev *)dets;
        if (dets) {
            P.est = EP;
            P.win_tab = h->d_win_tab; P.g_tab = h->d_g_tab;
            P.fix_ctl = fix_ctl; P.fix_list = fix_list;
            ST(get_twiddles(h, Nr, 0, &P.tw_r));
            ST(get_twiddles(h, Na, 1, &P.tw_a));
        }
        bool ok = false;
        const c32 *Hsrc = nullptr;
        if (bg || recording) {
            // background path: raw estimates -> ring update/subtraction -> fused kernel from H
            ST(h->sH.need((size_t)n_cpi * V * N * sizeof(c32)));
            ST(launch_chan_est(h, drx, dtx, n_cpi, (c32 *)h->sH.p, n_pre));
            P.H = Hsrc = (const c32 *)h->sH.p;
            ST(launch_fused<true>(h, P, &ok));
        } else {
            ST(launch_fused<false>(h, P, &ok));
        }
        if (ok) {
            h->last_path = JRC_PATH_FUSED;
            if (dets && !defer_exact) ST(launch_exact(h, drx, dtx, Hsrc, n_pre, cpi0, map, (DetDev *)dets, EP));
            return JRC_OK;
        }
        if (path == JRC_PATH_FUSED)
            return fail(JRC_ERR_INVALID, "fused kernel does not fit this configuration (n_sym = %d symbols in shared memory)", c.n_sym);
    }

    // ---- tiled path: configurations without a fused specialisation -------------
    const bool want_tiled = (path == JRC_PATH_AUTO || path == JRC_PATH_TILED) && cmap == nullptr && tiled_config_ok(h) &&
                            (map || dets);
    if (path == JRC_PATH_TILED && !want_tiled)
        return fail(JRC_ERR_INVALID, "tiled path requested but this configuration has no tiled kernel");
    if (want_tiled) {
        NvtxRange nv("tiled: chan_est + range FFT + angle FFT");
        h->last_path = JRC_PATH_TILED;
        const size_t per = (wide_config_ok(h) ? (size_t)4096 : ((size_t)V * N + (slice_config_ok(h) ? 0 : (size_t)V * Nr)) * sizeof(c32)) +
                           (map ? 0 : (size_t)Nr * Na * sizeof(float));
        int chunk = (int)(((size_t)1 << 30) / per);
        if (chunk < 1) chunk = 1;
        if (chunk > n_cpi) chunk = n_cpi;
        const bool slice = slice_config_ok(h);
        const bool wide = wide_config_ok(h);
        // the wide kernels hand G (2 MiB per CPI) from one to the other in rounds of CPIs.  Measured (n = 148): rounds of
        // 37 / 74 / 148 CPIs -> 350 / 371 / 384 k CPI/s, and 228 k at 8, although only rounds of <= 8 CPIs keep all of G in
        // the L2 (ncu, caches left alone: no write-back of G at 8, 100 % at 27; an access-policy window over G with the
        // persisting carve-out changed neither the traffic nor the time): the kernels are latency- and issue-bound, not
        // bandwidth-bound, and what a round costs is its partial last wave.  37 k CPIs fill both grids exactly
        // (128 x 37 = 16 x 296 units of the first kernel, 16 x 37 = 2 x 296 units of the second); 222 CPIs per round.
        static const int wide_round_env = getenv("JRC_WIDE_ROUND") ? atoi(getenv("JRC_WIDE_ROUND")) : 0;
        const int wide_round = wide_round_env > 0 ? wide_round_env : 222;
        if (wide) {
            const bool need_h = bg || recording;
            if (need_h) ST(h->sH.need((size_t)chunk * V * N * sizeof(c32)));
            ST(h->sY.need((size_t)(chunk < wide_round ? chunk : wide_round) * V * N * sizeof(c32)));
            if (!map) ST(h->sC.need((size_t)chunk * Nr * Na * sizeof(float)));
            if (dets) {
                ST(h->sKeys.need(sizeof(unsigned long long) * (size_t)chunk));
                ST(h->sSec.need(sizeof(unsigned) * (size_t)chunk));
            }
            for (int c0 = 0; c0 < n_cpi; c0 += chunk) {
                const int nc = n_cpi - c0 < chunk ? n_cpi - c0 : chunk;
                PortDev crx = drx, ctx = dtx;
                crx.base += (long long)c0 * rx.cpi_stride;
                ctx.base += (long long)c0 * tx.cpi_stride;
                c32 *dH = need_h ? (c32 *)h->sH.p : nullptr;
                float *dM = map ? map + (size_t)c0 * Nr * Na : (float *)h->sC.p;
                unsigned long long *dK = dets ? (unsigned long long *)h->sKeys.p : nullptr;
                unsigned *dS = dets ? (unsigned *)h->sSec.p : nullptr;
                if (need_h) ST(launch_chan_est(h, crx, ctx, nc, dH, n_pre));      // background ring in front of the transforms
                if (dK) {
                    CU(cudaMemsetAsync(dK, 0, sizeof(unsigned long long) * (size_t)nc, h->stream));
                    CU(cudaMemsetAsync(dS, 0, sizeof(unsigned) * (size_t)nc, h->stream));
                }
                for (int r0 = 0; r0 < nc; r0 += wide_round) {
                    const int nr = nc - r0 < wide_round ? nc - r0 : wide_round;
                    PortDev rrx = crx, rtx = ctx;
                    rrx.base += (long long)r0 * rx.cpi_stride;
                    rtx.base += (long long)r0 * tx.cpi_stride;
                    ST(launch_wide(h, rrx, rtx, dH ? dH + (size_t)r0 * V * N : nullptr, n_pre, nr, (c32 *)h->sY.p,
                                   dM + (size_t)r0 * Nr * Na, dK ? dK + r0 : nullptr, dS ? dS + r0 : nullptr));
                }
                if (dets) {
                    k_map_finalize<<<(unsigned)nc, 128, (size_t)Na * sizeof(float), h->stream>>>(
                        dM, dK, dS, nc, Nr, Na, EP, (DetDev *)dets + c0, cpi0 + c0, fix_ctl, fix_list);
                    CU(cudaGetLastError());
                    h->launches++;
                    ST(launch_exact(h, crx, ctx, dH, n_pre, cpi0 + c0, dM, (DetDev *)dets + c0, EP));
                }
            }
            return JRC_OK;
        }
        ST(h->sH.need((size_t)chunk * V * N * sizeof(c32)));
        if (!slice) ST(h->sY.need((size_t)chunk * V * Nr * sizeof(c32)));
        if (!map) ST(h->sC.need((size_t)chunk * Nr * Na * sizeof(float)));
        if (dets) {
            ST(h->sKeys.need(sizeof(unsigned long long) * (size_t)chunk));
            ST(h->sSec.need(sizeof(unsigned) * (size_t)chunk));
        }
        for (int c0 = 0; c0 < n_cpi; c0 += chunk) {
            const int nc = n_cpi - c0 < chunk ? n_cpi - c0 : chunk;
            PortDev crx = drx, ctx = dtx;
            crx.base += (long long)c0 * rx.cpi_stride;
            ctx.base += (long long)c0 * tx.cpi_stride;
            c32 *dH = (c32 *)h->sH.p, *dY = (c32 *)h->sY.p;
            float *dM = map ? map + (size_t)c0 * Nr * Na : (float *)h->sC.p;
            unsigned long long *dK = dets ? (unsigned long long *)h->sKeys.p : nullptr;
            unsigned *dS = dets ? (unsigned *)h->sSec.p : nullptr;
            ST(launch_chan_est(h, crx, ctx, nc, dH, n_pre));
            if (dK) {
                CU(cudaMemsetAsync(dK, 0, sizeof(unsigned long long) * (size_t)nc, h->stream));
                CU(cudaMemsetAsync(dS, 0, sizeof(unsigned) * (size_t)nc, h->stream));
            }
            if (slice) {
                ST(launch_slice(h, dH, nc, dM, dK, dS));
            } else {
                ST(launch_fft8_rows(h, dH, N, N, dY, Nr, (long long)nc * V));
                ST(launch_angle_mag(h, dY, V, Nr, Na, nc, dM, dK, dS));
            }
            if (dets) {
                k_map_finalize<<<(unsigned)nc, 128, (size_t)Na * sizeof(float), h->stream>>>(
                    dM, dK, dS, nc, Nr, Na, EP, (DetDev *)dets + c0, cpi0 + c0, fix_ctl, fix_list);
                CU(cudaGetLastError());
                h->launches++;
                // (the channel estimates of the chunk are still in dH: background removal included)
                ST(launch_exact(h, crx, ctx, dH, n_pre, cpi0 + c0, dM, (DetDev *)dets + c0, EP));
            }
        }
        return JRC_OK;
    }

    // ---- staged path: one kernel per reference block, chunked over CPIs ------
    NvtxRange nv("staged: one kernel per reference block");
    h->last_path = JRC_PATH_STAGED;
    const size_t per_cpi = ((size_t)V * N + (size_t)V * Nr + (cmap ? 0 : (size_t)Nr * Na)) * sizeof(c32);
    const size_t budget = (size_t)1 << 30;
    int chunk = (int)(budget / per_cpi);
    if (chunk < 1) chunk = 1;
    if (chunk > 32768) chunk = 32768;   // grid.z / grid.y limits of the staged kernels
    if (chunk > n_cpi) chunk = n_cpi;
    ST(h->sH.need((size_t)chunk * V * N * sizeof(c32)));
    ST(h->sY.need((size_t)chunk * V * Nr * sizeof(c32)));
    if (!cmap) ST(h->sC.need((size_t)chunk * Nr * Na * sizeof(c32)));
    for (int c0 = 0; c0 < n_cpi; c0 += chunk) {
        const int nc = n_cpi - c0 < chunk ? n_cpi - c0 : chunk;
        PortDev crx = drx, ctx = dtx;
        crx.base += (long long)c0 * rx.cpi_stride;
        ctx.base += (long long)c0 * tx.cpi_stride;
        c32 *dH = (c32 *)h->sH.p, *dY = (c32 *)h->sY.p;
        c32 *dC = cmap ? (c32 *)cmap + (size_t)c0 * Nr * Na : (c32 *)h->sC.p;
        ST(launch_chan_est(h, crx, ctx, nc, dH, n_pre));
        // fft_vcc #A: backward, no shift; the zero-padded tail is implied by n_in = N
        ST(launch_fft_rows(h, dH, N, N, dY, Nr, (long long)nc * V, 0, 0));
        // matrix_transpose + fft_vcc #B (forward, shift).  The transpose is materialised
        // (exactly the reference's data flow) into dC, then transformed in place.
        ST(launch_transpose(h, dY, dC, V, Nr, Na, nc));
        ST(launch_fft_rows(h, dC, Na, Na, dC, Na, (long long)nc * Nr, 1, 1));
        if (map) {
            long long n = (long long)nc * Nr * Na;
            k_mag_squared<<<grid_for(n, 256, h->sm_count), 256, 0, h->stream>>>(dC, map + (size_t)c0 * Nr * Na, n);
            CU(cudaGetLastError());
            h->launches++;
        }
        if (dets) ST(launch_estimate(h, dC, Nr, Na, nc, cpi0 + c0, (DetDev *)dets + c0));
    }
    return JRC_OK;
}

static bool ptr_is_device(const void *p);

// ---------------------------------------------------------------------------
// range-Doppler-angle cube over a burst of CPIs (no counterpart in the reference, whose chain
// stops at one range-angle map per CPI): complex maps of the burst from the one-kernel-per-block path, then one more
// fft_vcc (forward, shifted) along slow time for every (range, angle) cell, then |.|^2.
// ---------------------------------------------------------------------------
extern "C" jrc_status jrc_chain_run_burst(jrc_chain *h, jrc_port_layout rx, jrc_port_layout tx, int32_t n_burst, float *cube)
{
    if (!h || !cube) return fail(JRC_ERR_INVALID, "null argument");
    if (!is_pow2(n_burst) || n_burst > 16384) return fail(JRC_ERR_INVALID, "the burst length must be a power of two <= 16384");
    if (!ptr_is_device(cube)) return fail(JRC_ERR_INVALID, "cube must be device memory");
    CU(cudaSetDevice(h->cfg.device));
    NvtxRange nv("jrc_chain_run_burst");
    const size_t cells = (size_t)h->Nr * h->Na;
    if (cells * (size_t)n_burst > ((size_t)1 << 31)) return fail(JRC_ERR_INVALID, "cube too large");
    ST(h->sMisc.need(cells * n_burst * sizeof(c32)));        // complex maps [burst][cells]
    ST(h->sMisc2.need(cells * n_burst * sizeof(c32)));       // slow-time rows [cells][burst]
    c32 *cm = (c32 *)h->sMisc.p, *st = (c32 *)h->sMisc2.p;
    ST(run_batch_impl(h, rx, tx, n_burst, 0, nullptr, (jrc_c32 *)cm, nullptr, JRC_PATH_STAGED, h->cfg.n_pre, false));
    if (cells > (size_t)65535 * 32 * 32) return fail(JRC_ERR_INVALID, "map too large for the slow-time transpose");
    ST(launch_transpose(h, cm, st, n_burst, (int)cells, n_burst, 1));            // [burst][cells] -> [cells][burst]
    ST(launch_fft_rows(h, st, n_burst, n_burst, st, n_burst, (long long)cells, 1, 1));
    const long long n = (long long)cells * n_burst;
    k_mag_squared<<<grid_for(n, 256, h->sm_count), 256, 0, h->stream>>>(st, cube, n);
    CU(cudaGetLastError());
    h->launches++;
    return JRC_OK;
}

extern "C" jrc_status jrc_chain_run_batch(jrc_chain *h, jrc_port_layout rx, jrc_port_layout tx, int32_t n_cpi,
                                           int32_t cpi0, float *map, jrc_c32 *cmap, jrc_det *dets, int32_t path)
{
    if (!h) return fail(JRC_ERR_INVALID, "null handle");
    return run_batch_impl(h, rx, tx, n_cpi, cpi0, map, cmap, dets, path, h->cfg.n_pre, false);
}

// ---------------------------------------------------------------------------
// host-buffer chain: pinned double-buffered pipeline
// The chain fed with the RX antennas' raw time samples: the demodulation runs in front of the chain on the same stream,
// chunk by chunk, and hands the symbols over in a scratch small enough to stay in the L2 (4096 CPIs of configs[1]: 16 MiB).
extern "C" jrc_status jrc_chain_run_batch_time(jrc_chain *h, jrc_port_layout rx_time, int32_t cp_len, jrc_port_layout tx,
                                                int32_t n_cpi, int32_t cpi0, float *map, jrc_c32 *cmap, jrc_det *dets, int32_t path)
{
    if (!h) return fail(JRC_ERR_INVALID, "null handle");
    if (n_cpi < 0 || cp_len < 0) return fail(JRC_ERR_INVALID, "n_cpi < 0 or cp_len < 0");
    if (n_cpi == 0) return JRC_OK;
    if (!rx_time.base || !tx.base) return fail(JRC_ERR_INVALID, "null input pointer");
    CU(cudaSetDevice(h->cfg.device));
    const jrc_chain_cfg &c = h->cfg;
    const int N = c.fft_len, R = c.n_rx, S = c.n_sym;
    if (!is_pow2(N) || N > 4096) return fail(JRC_ERR_INVALID, "the OFDM demodulator needs a power-of-two fft_len <= 4096");
    if (c.background_removal || h->bg_recording.load())
        return fail(JRC_ERR_INVALID, "background removal keeps per-frame state: demodulate with jrc_ofdm_demod and use jrc_radar_estimate");
    NvtxRange nv("jrc_chain_run_batch_time");
    const size_t per_cpi = (size_t)R * S * N * sizeof(c32);
    int chunk = (int)(((size_t)32 << 20) / per_cpi);
    if (chunk < 1) chunk = 1;
    if (chunk > n_cpi) chunk = n_cpi;
    ST(h->sDemod.need((size_t)chunk * per_cpi));
    const c32 *tw = nullptr;
    ST(get_twiddles(h, N, 1, &tw));
    const int rpc = N >= 512 ? 1 : 512 / N;
    const size_t smem = (size_t)rpc * N * sizeof(c32);
    const size_t Nr = h->Nr, Na = h->Na;
    for (int c0 = 0; c0 < n_cpi; c0 += chunk) {
        const int nc = n_cpi - c0 < chunk ? n_cpi - c0 : chunk;
        PortDev drx{(const c32 *)rx_time.base + (long long)c0 * rx_time.cpi_stride, rx_time.cpi_stride, rx_time.ant_stride};
        const long long rows = (long long)nc * R * S;
        const bool al16 = aligned16(drx.base) && rx_time.cpi_stride % 2 == 0 && rx_time.ant_stride % 2 == 0 && cp_len % 2 == 0;
        if (N == 64 && al16)
            k_ofdm_demod64<<<grid_for(rows * 32, 256, h->sm_count), 256, 0, h->stream>>>(drx, R, c.n_pre, S, cp_len, (c32 *)h->sDemod.p,
                                                                                       rows, tw);
        else
            k_ofdm_demod_batch<<<(unsigned)((rows + rpc - 1) / rpc), 256, smem, h->stream>>>(drx, R, c.n_pre, S, cp_len,
                                                                                          (c32 *)h->sDemod.p, N, ilog2(N), rows, rpc, tw);
        CU(cudaGetLastError());
        h->launches++;
        jrc_port_layout sym{(const jrc_c32 *)h->sDemod.p, (int64_t)R * S * N, (int64_t)S * N};
        // tx keeps its n_pre symbols in front, the demodulated rx has none: tx's base moves instead of two offsets
        jrc_port_layout ctx = tx;
        ctx.base = tx.base + (long long)c0 * tx.cpi_stride + (long long)c.n_pre * N;
        ST(run_batch_impl(h, sym, ctx, nc, cpi0 + c0, map ? map + (size_t)c0 * Nr * Na : nullptr,
                          cmap ? cmap + (size_t)c0 * Nr * Na : nullptr, dets ? dets + c0 : nullptr, path, 0, false));
    }
    return JRC_OK;
}

// ---------------------------------------------------------------------------
// Pinned ranges created through jrc_pinned_alloc / jrc_host_register: looked up without a driver call (the streaming
// path asks four times per CPI).
struct PinnedRange { uintptr_t base; size_t len; uintptr_t dev; };
static std::mutex g_pin_mu;
static std::vector<PinnedRange> g_pinned;
static bool pinned_lookup(const void *p, void **dev)
{
    const uintptr_t a = (uintptr_t)p;
    std::lock_guard<std::mutex> g(g_pin_mu);
    for (const PinnedRange &r : g_pinned)
        if (a >= r.base && a < r.base + r.len) { if (dev) *dev = (void *)(r.dev + (a - r.base)); return true; }
    return false;
}
static void pinned_add(void *p, size_t len)
{
    void *dev = nullptr;
    if (cudaHostGetDevicePointer(&dev, p, 0) != cudaSuccess) { cudaGetLastError(); dev = p; }
    std::lock_guard<std::mutex> g(g_pin_mu);
    g_pinned.push_back({(uintptr_t)p, len, (uintptr_t)dev});
}
static void pinned_remove(void *p)
{
    std::lock_guard<std::mutex> g(g_pin_mu);
    for (size_t i = 0; i < g_pinned.size(); i++)
        if (g_pinned[i].base == (uintptr_t)p) { g_pinned.erase(g_pinned.begin() + i); return; }
}

static bool host_ptr_is_pinned(const void *p)
{
    if (pinned_lookup(p, nullptr)) return true;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}
// device-side alias of a pinned host pointer (unified addressing: normally the pointer itself), nullptr when the
// memory is not mapped into the device's address space
static void *host_dev_alias(const void *p)
{
    if (!p) return nullptr;
    void *dev = nullptr;
    if (pinned_lookup(p, &dev)) return dev;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}
static bool ptr_is_device(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// The final scalar of lib/range_angle_estimator_impl.cc:227,234 with the HOST libm, like jrc_estimate2d: what a block
// publishes has the bit pattern the reference's own expression gives on this record's peak and noise power.  A record
// whose device-side gate was within the margin was redone in the reference's order before (jrc_exact.cuh).
static void finish_records_host(const jrc_chain *h, jrc_det *d, int n)
{
    for (int i = 0; i < n; i++) {
        if (d[i].range_idx < 0) continue;
        d[i].snr_db = 10 * std::log10(d[i].peak_power / d[i].noise_power);
        const uint32_t pass = (d[i].snr_db >= h->snr_thr && d[i].peak_power >= h->pow_thr) ? JRC_DET_PASSED : 0u;
        d[i].flags = pass | (d[i].flags & JRC_DET_EXACT);
    }
}

// ---------------------------------------------------------------------------
// streaming: jrc_chain_submit / jrc_chain_wait (BASELINE configs[3]).  Up to JRC_STREAM_DEPTH submissions are in
// flight, each on the stream of its slot, so CPI k+1's input transfer and kernel overlap CPI k's output transfer.
// ---------------------------------------------------------------------------
struct StreamSlot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    GrowBuf in, map, dets, pin_in, pin_out, fix, exact;
    bool busy = false, deferred = false;
    int64_t ticket = 0;
    int n_cpi = 0, cpi0 = 0, tx_shared = 0;
    float *map_host = nullptr, *map_stage = nullptr;         // caller buffer / pinned staging copy (nullptr: written in place)
    jrc_det *dets_host = nullptr, *dets_final = nullptr;     // caller buffer / where the records land (pinned)
    const c32 *z_rx = nullptr, *z_tx = nullptr;              // device aliases of the (pinned) inputs, deferred exact pass
    float *z_map = nullptr;
    jrc_det *z_dets = nullptr;
};

struct jrc_stream_state {
    StreamSlot slot[JRC_STREAM_DEPTH];
    int64_t next_ticket = 1;
    bool ready = false;
};

static jrc_status stream_state(jrc_chain *h, jrc_stream_state **out)
{
    if (!h->sstate) h->sstate = new jrc_stream_state();
    jrc_stream_state *S = h->sstate;
    if (!S->ready) {
        for (StreamSlot &sl : S->slot) {
            CU(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
            sl.pin_in.pinned = sl.pin_out.pinned = true;
        }
        S->ready = true;
    }
    *out = S;
    return JRC_OK;
}

static void stream_state_destroy(jrc_chain *h)
{
    if (!h->sstate) return;
    for (StreamSlot &sl : h->sstate->slot) {
        if (sl.stream) { cudaStreamSynchronize(sl.stream); cudaStreamDestroy(sl.stream); }
        if (sl.done) cudaEventDestroy(sl.done);
        GrowBuf *bufs[] = {&sl.in, &sl.map, &sl.dets, &sl.pin_in, &sl.pin_out, &sl.fix, &sl.exact};
        for (GrowBuf *b : bufs) b->release();
    }
    delete h->sstate;
    h->sstate = nullptr;
}

static jrc_status stream_stats(jrc_chain *h, const std::function<jrc_status(const GrowBuf &)> &add)
{
    if (!h->sstate) return JRC_OK;
    for (StreamSlot &sl : h->sstate->slot) {
        if (sl.stream) CU(cudaStreamSynchronize(sl.stream));
        ST(add(sl.fix));
    }
    return JRC_OK;
}

// runs fn with the slot's stream and marked-CPI list in place of the handle's
template <class F>
static jrc_status on_slot(jrc_chain *h, StreamSlot &sl, F fn)
{
    std::swap(h->stream, sl.stream);
    std::swap(h->sFix, sl.fix);
    std::swap(h->sExact, sl.exact);
    jrc_status st = fn();
    std::swap(h->stream, sl.stream);
    std::swap(h->sFix, sl.fix);
    std::swap(h->sExact, sl.exact);
    return st;
}

extern "C" jrc_status jrc_chain_submit(jrc_chain *h, const jrc_c32 *rx_host, const jrc_c32 *tx_host, int32_t tx_shared,
                                        int32_t n_cpi, int32_t cpi0, float *map_host, jrc_det *dets_host, int64_t *ticket)
{
    if (!h || !rx_host || !tx_host || !ticket) return fail(JRC_ERR_INVALID, "null argument");
    if (n_cpi <= 0) return fail(JRC_ERR_INVALID, "n_cpi must be positive");
    if (dets_host && !h->est_set) return fail(JRC_ERR_STATE, "dets requested but jrc_chain_set_estimator has not been called");
    CU(cudaSetDevice(h->cfg.device));
    jrc_stream_state *S = nullptr;
    ST(stream_state(h, &S));
    StreamSlot *slp = nullptr;
    for (StreamSlot &c : S->slot) if (!c.busy) { slp = &c; break; }
    if (!slp) return fail(JRC_ERR_STATE, "%d submissions in flight: jrc_chain_wait first", (int)JRC_STREAM_DEPTH);
    StreamSlot &sl = *slp;
    const jrc_chain_cfg &c = h->cfg;
    NvtxRange nv("jrc_chain_submit");
    // the background ring buffer evolves frame by frame (lib/mimo_ofdm_radar_impl.cc:276-300): with it, submissions
    // run one after the other
    const bool bgp = c.background_removal || h->bg_recording.load();
    if (bgp)
        for (StreamSlot &o : S->slot)
            if (o.busy) CU(cudaStreamSynchronize(o.stream));
    const size_t rx_cpi = (size_t)c.n_rx * c.n_sym * c.fft_len, tx_cpi = (size_t)c.n_tx * c.n_sym * c.fft_len;
    const size_t map_cpi = (size_t)h->Nr * h->Na;
    const size_t txn = tx_shared ? tx_cpi : (size_t)n_cpi * tx_cpi;
    const size_t rx_bytes = (size_t)n_cpi * rx_cpi * sizeof(c32), tx_bytes = txn * sizeof(c32);
    const size_t map_bytes = (size_t)n_cpi * map_cpi * sizeof(float), det_bytes = (size_t)n_cpi * sizeof(jrc_det);

    // inputs: pinned caller memory is used in place, pageable memory goes through the slot's pinned buffer
    const c32 *src_rx = (const c32 *)rx_host, *src_tx = (const c32 *)tx_host;
    if (!(host_ptr_is_pinned(rx_host) && host_ptr_is_pinned(tx_host))) {
        ST(sl.pin_in.need(rx_bytes + tx_bytes));
        memcpy(sl.pin_in.p, rx_host, rx_bytes);
        memcpy((char *)sl.pin_in.p + rx_bytes, tx_host, tx_bytes);
        src_rx = (const c32 *)sl.pin_in.p;
        src_tx = (const c32 *)((char *)sl.pin_in.p + rx_bytes);
    }
    // outputs: likewise
    float *dst_map = map_host;
    jrc_det *dst_dets = dets_host;
    const bool map_direct = !map_host || host_ptr_is_pinned(map_host), dets_direct = !dets_host || host_ptr_is_pinned(dets_host);
    if (!map_direct || !dets_direct) {
        ST(sl.pin_out.need((map_direct ? 0 : map_bytes) + (dets_direct ? 0 : det_bytes)));
        if (!map_direct) dst_map = (float *)sl.pin_out.p;
        if (!dets_direct) dst_dets = (jrc_det *)((char *)sl.pin_out.p + (map_direct ? 0 : map_bytes));
    }
    sl.n_cpi = n_cpi; sl.cpi0 = cpi0; sl.tx_shared = tx_shared;
    sl.map_host = map_host; sl.map_stage = map_direct ? nullptr : dst_map;
    sl.dets_host = dets_host; sl.dets_final = dst_dets;
    sl.deferred = false;

    const c32 *z_rx = (const c32 *)host_dev_alias(src_rx), *z_tx = (const c32 *)host_dev_alias(src_tx);
    float *z_map = (float *)host_dev_alias(dst_map);
    jrc_det *z_dets = (jrc_det *)host_dev_alias(dst_dets);
    const long long ant = (long long)c.n_sym * c.fft_len;
    jrc_status st;
    if (h->zero_copy && n_cpi <= 4 && fused_config_ok(h) && z_rx && z_tx && (!map_host || z_map) && (!dets_host || z_dets)) {
        // A few CPIs: the copies cost more than the kernel.  Pinned host memory is device-accessible (unified
        // addressing): the kernel prefetches the symbols over PCIe itself and streams map and records straight into
        // host memory while it computes -- one launch.  Marked records (jrc_exact.cuh) are looked at in wait().
        jrc_port_layout zrx{(const jrc_c32 *)z_rx, (int64_t)rx_cpi, ant};
        jrc_port_layout ztx{(const jrc_c32 *)z_tx, tx_shared ? 0 : (int64_t)tx_cpi, ant};
        sl.z_rx = z_rx; sl.z_tx = z_tx; sl.z_map = z_map; sl.z_dets = z_dets;
        sl.deferred = dets_host != nullptr && !bgp;
        st = on_slot(h, sl, [&]() { return run_batch_impl(h, zrx, ztx, n_cpi, cpi0, z_map, nullptr, z_dets, JRC_PATH_AUTO, 0, sl.deferred); });
    } else {
        ST(sl.in.need(rx_bytes + tx_bytes));
        if (map_host) ST(sl.map.need(map_bytes));
        if (dets_host) ST(sl.dets.need(det_bytes));
        c32 *d_rx = (c32 *)sl.in.p, *d_tx = (c32 *)((char *)sl.in.p + rx_bytes);
        float *d_map = map_host ? (float *)sl.map.p : nullptr;
        jrc_det *d_dets = dets_host ? (jrc_det *)sl.dets.p : nullptr;
        CU(cudaMemcpyAsync(d_rx, src_rx, rx_bytes, cudaMemcpyHostToDevice, sl.stream));
        CU(cudaMemcpyAsync(d_tx, src_tx, tx_bytes, cudaMemcpyHostToDevice, sl.stream));
        jrc_port_layout lrx{(const jrc_c32 *)d_rx, (int64_t)rx_cpi, ant};
        jrc_port_layout ltx{(const jrc_c32 *)d_tx, tx_shared ? 0 : (int64_t)tx_cpi, ant};
        st = on_slot(h, sl, [&]() { return run_batch_impl(h, lrx, ltx, n_cpi, cpi0, d_map, nullptr, d_dets, JRC_PATH_AUTO, 0, false); });
        if (st == JRC_OK && map_host) CU(cudaMemcpyAsync(dst_map, d_map, map_bytes, cudaMemcpyDeviceToHost, sl.stream));
        if (st == JRC_OK && dets_host) CU(cudaMemcpyAsync(dst_dets, d_dets, det_bytes, cudaMemcpyDeviceToHost, sl.stream));
    }
    if (st != JRC_OK) return st;
    // (no event: the slot's stream carries this submission and nothing else until it is waited for)
    sl.busy = true;
    sl.ticket = S->next_ticket++;
    *ticket = sl.ticket;
    return JRC_OK;
}

// pinned host memory for callers without a CUDA toolchain of their own (GNU Radio blocks, ctypes)
extern "C" jrc_status jrc_pinned_alloc(size_t bytes, void **out)
{
    if (!out) return fail(JRC_ERR_INVALID, "null argument");
    *out = nullptr;
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped));
    pinned_add(*out, bytes ? bytes : 1);
    return JRC_OK;
}
extern "C" jrc_status jrc_pinned_free(void *p)
{
    if (p) { pinned_remove(p); CU(cudaFreeHost(p)); }
    return JRC_OK;
}
// page-locks an existing buffer (a GNU Radio stream buffer, a NumPy array) so that submit / run_host use it in place
extern "C" jrc_status jrc_host_register(void *p, size_t bytes)
{
    if (!p || !bytes) return fail(JRC_ERR_INVALID, "null argument");
    CU(cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    pinned_add(p, bytes);
    return JRC_OK;
}
extern "C" jrc_status jrc_host_unregister(void *p)
{
    if (!p) return fail(JRC_ERR_INVALID, "null argument");
    pinned_remove(p);
    CU(cudaHostUnregister(p));
    return JRC_OK;
}

// ---------------------------------------------------------------------------
// peer memory for the detection records of a multi-GPU job: the host rank allocates the table, the other ranks map it
// (CUDA IPC) and their kernels store the 32-byte records straight into it over NVLink
// ---------------------------------------------------------------------------
extern "C" jrc_status jrc_dev_alloc(int32_t device, size_t bytes, void **out)
{
    if (!out) return fail(JRC_ERR_INVALID, "null argument");
    *out = nullptr;
    CU(cudaSetDevice(device));
    CU(cudaMalloc(out, bytes ? bytes : 1));
    CU(cudaMemset(*out, 0, bytes ? bytes : 1));
    return JRC_OK;
}
extern "C" jrc_status jrc_dev_free(void *p)
{
    if (p) CU(cudaFree(p));
    return JRC_OK;
}
extern "C" jrc_status jrc_dev_copy(void *dst, const void *src, size_t bytes)
{
    if (!dst || !src) return fail(JRC_ERR_INVALID, "null argument");
    CU(cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));
    return JRC_OK;
}
extern "C" jrc_status jrc_chain_copy_async(jrc_chain *h, void *dst, const void *src, size_t bytes)
{
    if (!h || !dst || !src) return fail(JRC_ERR_INVALID, "null argument");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, h->stream));
    return JRC_OK;
}
extern "C" jrc_status jrc_ipc_export(void *dev_ptr, void *handle64)
{
    if (!dev_ptr || !handle64) return fail(JRC_ERR_INVALID, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handle64), dev_ptr));
    return JRC_OK;
}
extern "C" jrc_status jrc_ipc_open(const void *handle64, int32_t device, void **out)
{
    if (!handle64 || !out) return fail(JRC_ERR_INVALID, "null argument");
    *out = nullptr;
    CU(cudaSetDevice(device));
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, handle64, sizeof(hnd));
    CU(cudaIpcOpenMemHandle(out, hnd, cudaIpcMemLazyEnablePeerAccess));
    return JRC_OK;
}
extern "C" jrc_status jrc_ipc_close(void *p)
{
    if (p) CU(cudaIpcCloseMemHandle(p));
    return JRC_OK;
}

static StreamSlot *find_ticket(jrc_chain *h, int64_t ticket)
{
    if (!h->sstate) return nullptr;
    for (StreamSlot &c : h->sstate->slot) if (c.busy && c.ticket == ticket) return &c;
    return nullptr;
}

extern "C" jrc_status jrc_chain_poll(jrc_chain *h, int64_t ticket, int32_t *done)
{
    if (!h || !done) return fail(JRC_ERR_INVALID, "null argument");
    StreamSlot *sl = find_ticket(h, ticket);
    if (!sl) return fail(JRC_ERR_INVALID, "unknown ticket %lld", (long long)ticket);
    cudaError_t e = cudaStreamQuery(sl->stream);
    if (e != cudaSuccess && e != cudaErrorNotReady) return fail(JRC_ERR_CUDA, "cudaStreamQuery: %s", cudaGetErrorString(e));
    *done = e == cudaSuccess;
    return JRC_OK;
}

extern "C" jrc_status jrc_chain_wait(jrc_chain *h, int64_t ticket)
{
    if (!h) return fail(JRC_ERR_INVALID, "null handle");
    StreamSlot *slp = find_ticket(h, ticket);
    if (!slp) return fail(JRC_ERR_INVALID, "unknown ticket %lld", (long long)ticket);
    StreamSlot &sl = *slp;
    CU(cudaSetDevice(h->cfg.device));
    NvtxRange nv("jrc_chain_wait");
    sl.busy = false;                       // whatever happens below, the slot is free again
    CU(cudaStreamSynchronize(sl.stream));
    const jrc_chain_cfg &c = h->cfg;
    if (sl.dets_final) {
        if (sl.deferred) {
            // marked records (jrc_exact.cuh) are rare: the reference-order pass is only launched when one is there
            bool marked = false;
            for (int i = 0; i < sl.n_cpi; i++) marked |= (sl.dets_final[i].flags & DET_PENDING) != 0;
            if (marked) {
                EstParams EP;
                ST(est_params(h, h->Nr, h->Na, &EP));
                const long long ant = (long long)c.n_sym * c.fft_len;
                PortDev prx{sl.z_rx, (long long)c.n_rx * ant, ant};
                PortDev ptx{sl.z_tx, sl.tx_shared ? 0 : (long long)c.n_tx * ant, ant};
                ST(on_slot(h, sl, [&]() { return launch_exact(h, prx, ptx, nullptr, 0, sl.cpi0, sl.z_map, (DetDev *)sl.z_dets, EP); }));
                CU(cudaStreamSynchronize(sl.stream));
            }
        }
        finish_records_host(h, sl.dets_final, sl.n_cpi);
        if (sl.dets_final != sl.dets_host) memcpy(sl.dets_host, sl.dets_final, (size_t)sl.n_cpi * sizeof(jrc_det));
    }
    if (sl.map_stage) memcpy(sl.map_host, sl.map_stage, (size_t)sl.n_cpi * h->Nr * h->Na * sizeof(float));
    return JRC_OK;
}

extern "C" jrc_status jrc_chain_run_host(jrc_chain *h, const jrc_c32 *rx_host, const jrc_c32 *tx_host, int32_t tx_shared,
                                          int32_t n_cpi, int32_t cpi0, float *map_host, jrc_det *dets_host)
{
    if (!h || !rx_host || !tx_host) return fail(JRC_ERR_INVALID, "null argument");
    if (n_cpi <= 0) return n_cpi == 0 ? JRC_OK : fail(JRC_ERR_INVALID, "n_cpi < 0");
    if (dets_host && !h->est_set) return fail(JRC_ERR_STATE, "dets requested but jrc_chain_set_estimator has not been called");
    CU(cudaSetDevice(h->cfg.device));
    const jrc_chain_cfg &c = h->cfg;
    const size_t rx_cpi = (size_t)c.n_rx * c.n_sym * c.fft_len, tx_cpi = (size_t)c.n_tx * c.n_sym * c.fft_len;
    const size_t map_cpi = (size_t)h->Nr * h->Na;
    // chunk so that one slot's map stays <= 128 MiB
    int chunk = (int)(((size_t)128 << 20) / (map_cpi * sizeof(float)));
    if (chunk < 1) chunk = 1;
    if (chunk > n_cpi) chunk = n_cpi;
    if (n_cpi <= chunk) {
        // one chunk (typically one CPI per work() call): submit + wait on one slot
        if (h->sstate)
            for (StreamSlot &sl : h->sstate->slot)
                if (sl.busy) return fail(JRC_ERR_STATE, "jrc_chain_run_host while submissions are in flight");
        int64_t ticket = 0;
        ST(jrc_chain_submit(h, rx_host, tx_host, tx_shared, n_cpi, cpi0, map_host, dets_host, &ticket));
        return jrc_chain_wait(h, ticket);
    }
    const bool direct = host_ptr_is_pinned(rx_host) && host_ptr_is_pinned(tx_host) &&
                        (!map_host || host_ptr_is_pinned(map_host)) && (!dets_host || host_ptr_is_pinned(dets_host));
    const size_t in_bytes = (rx_cpi + (tx_shared ? 0 : tx_cpi)) * sizeof(c32);
    for (int s = 0; s < 2; s++) {
        ST(h->sIn[s].need((size_t)chunk * in_bytes + (tx_shared ? tx_cpi * sizeof(c32) : 0)));
        if (map_host) ST(h->sMap[s].need((size_t)chunk * map_cpi * sizeof(float)));
        if (dets_host) ST(h->sDets[s].need((size_t)chunk * sizeof(jrc_det)));
    }
    if (!direct) {
        ST(h->pin_a.need((size_t)chunk * in_bytes + tx_cpi * sizeof(c32)));
        ST(h->pin_b.need((size_t)chunk * (map_host ? map_cpi * sizeof(float) : 0) + (size_t)chunk * sizeof(jrc_det)));
    }
    if (!h->s_h2d) {
        CU(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            CU(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming));
        }
    }
    jrc_status st = JRC_OK;
    NvtxRange nv_host("jrc_chain_run_host (pipelined chunks)");
    int idx = 0;
    for (int c0 = 0; c0 < n_cpi && st == JRC_OK; c0 += chunk, idx++) {
        const int nc = n_cpi - c0 < chunk ? n_cpi - c0 : chunk;
        const int s = idx & 1;
        c32 *d_rx = (c32 *)h->sIn[s].p;
        c32 *d_tx = d_rx + (size_t)chunk * rx_cpi;
        const c32 *src_rx = (const c32 *)rx_host + (size_t)c0 * rx_cpi;
        const c32 *src_tx = (const c32 *)tx_host + (tx_shared ? 0 : (size_t)c0 * tx_cpi);
        const size_t txn = tx_shared ? tx_cpi : (size_t)nc * tx_cpi;
        cudaError_t e = cudaSuccess;
        // the slot's previous compute must be done before its inputs are overwritten
        if (idx >= 2) e = cudaStreamWaitEvent(h->s_h2d, h->ev_comp[s], 0);
        if (!direct) {
            // pageable caller memory: stage through pinned buffers, chunk-synchronously
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->s_h2d);
            memcpy(h->pin_a.p, src_rx, (size_t)nc * rx_cpi * sizeof(c32));
            memcpy((c32 *)h->pin_a.p + (size_t)nc * rx_cpi, src_tx, txn * sizeof(c32));
            src_rx = (const c32 *)h->pin_a.p;
            src_tx = src_rx + (size_t)nc * rx_cpi;
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_rx, src_rx, (size_t)nc * rx_cpi * sizeof(c32), cudaMemcpyHostToDevice, h->s_h2d);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_tx, src_tx, txn * sizeof(c32), cudaMemcpyHostToDevice, h->s_h2d);
        if (e == cudaSuccess) e = cudaEventRecord(h->ev_in[s], h->s_h2d);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(h->stream, h->ev_in[s], 0);
        // the slot's previous D2H must be done before its outputs are overwritten
        if (e == cudaSuccess && idx >= 2) e = cudaStreamWaitEvent(h->stream, h->ev_out[s], 0);
        if (e != cudaSuccess) { st = fail(JRC_ERR_CUDA, "pipeline H2D: %s", cudaGetErrorString(e)); break; }
        jrc_port_layout lrx{(const jrc_c32 *)d_rx, (int64_t)rx_cpi, (int64_t)c.n_sym * c.fft_len};
        jrc_port_layout ltx{(const jrc_c32 *)d_tx, tx_shared ? 0 : (int64_t)tx_cpi, (int64_t)c.n_sym * c.fft_len};
        float *d_map = map_host ? (float *)h->sMap[s].p : nullptr;
        jrc_det *d_dets = dets_host ? (jrc_det *)h->sDets[s].p : nullptr;
        st = run_batch_impl(h, lrx, ltx, nc, cpi0 + c0, d_map, nullptr, d_dets, JRC_PATH_AUTO, 0, false);
        if (st != JRC_OK) break;
        e = cudaEventRecord(h->ev_comp[s], h->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(h->s_d2h, h->ev_comp[s], 0);
        float *dst_map = map_host ? map_host + (size_t)c0 * map_cpi : nullptr;
        jrc_det *dst_dets = dets_host ? dets_host + c0 : nullptr;
        if (!direct) {
            dst_map = map_host ? (float *)h->pin_b.p : nullptr;
            dst_dets = dets_host ? (jrc_det *)((char *)h->pin_b.p + (map_host ? (size_t)chunk * map_cpi * sizeof(float) : 0)) : nullptr;
        }
        if (e == cudaSuccess && map_host)
            e = cudaMemcpyAsync(dst_map, d_map, (size_t)nc * map_cpi * sizeof(float), cudaMemcpyDeviceToHost, h->s_d2h);
        if (e == cudaSuccess && dets_host)
            e = cudaMemcpyAsync(dst_dets, d_dets, (size_t)nc * sizeof(jrc_det), cudaMemcpyDeviceToHost, h->s_d2h);
        if (e == cudaSuccess) e = cudaEventRecord(h->ev_out[s], h->s_d2h);
        if (e == cudaSuccess && !direct) {
            e = cudaStreamSynchronize(h->s_d2h);
            if (e == cudaSuccess && map_host) memcpy(map_host + (size_t)c0 * map_cpi, dst_map, (size_t)nc * map_cpi * sizeof(float));
            if (e == cudaSuccess && dets_host) { finish_records_host(h, dst_dets, nc); memcpy(dets_host + c0, dst_dets, (size_t)nc * sizeof(jrc_det)); }
        }
        if (e != cudaSuccess) { st = fail(JRC_ERR_CUDA, "pipeline D2H: %s", cudaGetErrorString(e)); break; }
    }
    cudaError_t e1 = cudaStreamSynchronize(h->s_h2d), e2 = cudaStreamSynchronize(h->stream), e3 = cudaStreamSynchronize(h->s_d2h);
    if (st == JRC_OK && (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess))
        st = fail(JRC_ERR_CUDA, "pipeline sync: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
    if (st == JRC_OK && dets_host && direct) finish_records_host(h, dets_host, n_cpi);
    return st;
}

// ---------------------------------------------------------------------------
// per-block stage calls.  Pointers may be host or device.
// ---------------------------------------------------------------------------
struct Staging {   // maps caller pointers onto device memory for the duration of one call
    jrc_chain *h;
    struct Out { void *host; void *dev; size_t bytes; };
    std::vector<Out> outs;
    int slot = 0;      // grow-only staging buffers of the handle: no allocation on the per-work() path
    explicit Staging(jrc_chain *hh) : h(hh) {}
    jrc_status take(size_t bytes, void **d)
    {
        if (slot >= (int)(sizeof(h->sStage) / sizeof(h->sStage[0]))) return fail(JRC_ERR_STATE, "too many staged buffers in one call");
        GrowBuf &b = h->sStage[slot++];
        ST(b.need(bytes ? bytes : 1));
        *d = b.p;
        return JRC_OK;
    }
    jrc_status in(const void *p, size_t bytes, const void **dev)
    {
        if (ptr_is_device(p)) { *dev = p; return JRC_OK; }
        void *d = nullptr;
        ST(take(bytes, &d));
        CU(cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, h->stream));
        *dev = d;
        return JRC_OK;
    }
    jrc_status out(void *p, size_t bytes, void **dev)
    {
        if (ptr_is_device(p)) { *dev = p; return JRC_OK; }
        void *d = nullptr;
        ST(take(bytes, &d));
        outs.push_back({p, d, bytes});
        *dev = d;
        return JRC_OK;
    }
    jrc_status finish()
    {
        for (auto &o : outs) CU(cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        return JRC_OK;
    }
};

// ---- fused mode for an unmodified flowgraph: ring of results keyed by CPI sequence number -------------------
struct FusedEntry {
    std::atomic<int64_t> seq{-1};        // -1 while the device may be writing the entry
    cudaEvent_t done = nullptr;          // everything of the entry is on the host
    cudaEvent_t t_done = nullptr;        // the transposed spectra are (an external event-record node of the entry's graph)
    c32 *T = nullptr;                    // page-locked [Nr][V]: the V data columns of matrix_transpose's [Nr][Na] output.  The
                                         // other Na - V columns are zeros (lib/matrix_transpose_impl.cc:91-104): they never
                                         // cross the bus, the consumer's fetch writes them into its buffer itself
    DetDev *det = nullptr;               // page-locked
    cudaGraphExec_t graph = nullptr;     // the rest of the chain into THIS entry, captured once
    int epoch = -1;
};
struct jrc_fused_state {
    cudaStream_t stream = nullptr, stream2 = nullptr;
    cudaEvent_t h_ready = nullptr, t_ready = nullptr, t_join = nullptr;   // (t_ready / t_join: fork and join inside the captured graph)
    cudaEvent_t last_done = nullptr;     // the previous frame's continuation (it reads the estimate the next call overwrites)
    FusedEntry e[JRC_FUSED_RING];
    int64_t next_seq = 0;
    GrowBuf dT, dDet, dKeys;             // dT: [Nr][V] on the device; dKeys: zero between frames
};

static jrc_status fused_state(jrc_chain *h, jrc_fused_state **out)
{
    if (!h->fstate) {
        jrc_fused_state *F = new jrc_fused_state();
        h->fstate = F;
        CU(cudaStreamCreateWithFlags(&F->stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&F->stream2, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&F->h_ready, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&F->t_ready, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&F->t_join, cudaEventDisableTiming));
        for (FusedEntry &e : F->e) {
            CU(cudaEventCreateWithFlags(&e.done, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&e.t_done, cudaEventDisableTiming));
            CU(cudaMallocHost(&e.T, (size_t)h->Nr * h->V * sizeof(c32)));
            CU(cudaMallocHost(&e.det, sizeof(DetDev)));
        }
    }
    *out = h->fstate;
    return JRC_OK;
}

static void fused_state_destroy(jrc_chain *h)
{
    jrc_fused_state *F = h->fstate;
    if (!F) return;
    if (F->stream) { cudaStreamSynchronize(F->stream); cudaStreamDestroy(F->stream); }
    if (F->stream2) { cudaStreamSynchronize(F->stream2); cudaStreamDestroy(F->stream2); }
    if (F->t_ready) cudaEventDestroy(F->t_ready);
    if (F->t_join) cudaEventDestroy(F->t_join);
    if (F->h_ready) cudaEventDestroy(F->h_ready);
    for (FusedEntry &e : F->e) {
        if (e.graph) cudaGraphExecDestroy(e.graph);
        if (e.done) cudaEventDestroy(e.done);
        if (e.t_done) cudaEventDestroy(e.t_done);
        if (e.T) cudaFreeHost(e.T);
        if (e.det) cudaFreeHost(e.det);
    }
    F->dT.release();
    F->dDet.release();
    F->dKeys.release();
    delete F;
    h->fstate = nullptr;
}

static jrc_status radar_estimate_impl(jrc_chain *h, const jrc_c32 *const *tx, const jrc_c32 *const *rx, size_t tx_skip_items,
                                      jrc_c32 *out, jrc_c32 *chan_est_host, int64_t *cpi_seq)
{
    if (!h || !tx || !rx || !out) return fail(JRC_ERR_INVALID, "null argument");
    CU(cudaSetDevice(h->cfg.device));
    const jrc_chain_cfg &c = h->cfg;
    const int N = c.fft_len, V = h->V, Nr = h->Nr, Na = h->Na;
    jrc_fused_state *F = nullptr;
    FusedEntry *fused_entry_pending = nullptr;
    int64_t fused_seq = -1;
    if (cpi_seq) {
        if (!h->est_set) return fail(JRC_ERR_STATE, "fused mode needs jrc_chain_set_estimator on this handle");
        if (!is_pow2(Nr) || !is_pow2(Na) || Nr > 16384 || Na > 16384)
            return fail(JRC_ERR_INVALID, "the FFT chain needs power-of-two Nr=%d / Na=%d <= 16384", Nr, Na);
        ST(fused_state(h, &F));
        if (F->last_done) CU(cudaStreamWaitEvent(h->stream, F->last_done, 0));
    }
    const size_t frame = (size_t)(c.n_pre + c.n_sym) * N;   // items actually read per port
    Staging sg(h);
    // gather the per-port packets into one device block [T+R][frame]
    ST(h->sMisc.need((size_t)(c.n_tx + c.n_rx) * frame * sizeof(c32)));
    c32 *blk = (c32 *)h->sMisc.p;
    const bool host_in = !ptr_is_device(tx[0]) && !ptr_is_device(rx[0]);
    if (host_in) {
        // stream buffers of the scheduler (pageable): one packed page-locked block instead of a driver-staged copy per
        // port; the estimate kernel reads it in place over PCIe (27 KiB, each sample once) unless JRC_ZEROCOPY=0
        const size_t bytes = (size_t)(c.n_tx + c.n_rx) * frame * sizeof(c32);
        ST(h->pin_a.need(bytes));
        c32 *pk = (c32 *)h->pin_a.p;
        for (int t = 0; t < c.n_tx; t++)
            memcpy(pk + (size_t)t * frame, tx[t] + tx_skip_items * (size_t)N, frame * sizeof(c32));   // lib/mimo_ofdm_radar_impl.cc:260
        for (int r = 0; r < c.n_rx; r++) memcpy(pk + (size_t)(c.n_tx + r) * frame, rx[r], frame * sizeof(c32));
        c32 *alias = h->zero_copy ? (c32 *)host_dev_alias(pk) : nullptr;
        if (alias) blk = alias;
        else CU(cudaMemcpyAsync(blk, pk, bytes, cudaMemcpyHostToDevice, h->stream));
    } else {
        for (int t = 0; t < c.n_tx; t++) {
            const jrc_c32 *src = tx[t] + tx_skip_items * (size_t)N;
            CU(cudaMemcpyAsync(blk + (size_t)t * frame, src, frame * sizeof(c32), cudaMemcpyDefault, h->stream));
        }
        for (int r = 0; r < c.n_rx; r++)
            CU(cudaMemcpyAsync(blk + (size_t)(c.n_tx + r) * frame, rx[r], frame * sizeof(c32), cudaMemcpyDefault, h->stream));
    }
    PortDev dtx{blk, 0, (long long)frame}, drx{blk + (size_t)c.n_tx * frame, 0, (long long)frame};
    ST(h->sH.need((size_t)V * N * sizeof(c32)));
    c32 *dH = (c32 *)h->sH.p;
    ST(launch_chan_est(h, drx, dtx, 1, dH, c.n_pre));
    if (F) {
        // the rest of the chain, in the arithmetic of the separate blocks, behind this call's back: one graph launch on
        // the second stream -- three kernels (range fft_vcc storing its spectra transposed, angle fft_vcc with the arg-max
        // scan as its epilogue, k_est_finalize storing the record into its page-locked slot), one 32 KiB copy, one event
        NvtxRange nv("fused mode: range fft, transpose, angle fft, estimator");
        const int64_t seq = F->next_seq++;
        FusedEntry &e = F->e[seq % JRC_FUSED_RING];
        e.seq.store(-1, std::memory_order_release);
        if (!e.graph || e.epoch != h->est_epoch) {
            const size_t cells = (size_t)Nr * Na, tcells = (size_t)Nr * V;
            ST(h->sC.need(cells * sizeof(c32)));
            ST(F->dT.need(tcells * sizeof(c32)));
            ST(F->dDet.need(sizeof(DetDev)));
            if (!F->dKeys.p) {
                ST(F->dKeys.need(sizeof(unsigned long long)));
                CU(cudaMemsetAsync(F->dKeys.p, 0, sizeof(unsigned long long), h->stream));   // (ordered before the graph by h_ready)
            }
            const c32 *tw = nullptr;
            ST(get_twiddles(h, Nr, 0, &tw));        // (tables are built on the handle's stream, before the capture)
            ST(get_twiddles(h, Na, 1, &tw));
            if (e.graph) { CU(cudaGraphExecDestroy(e.graph)); e.graph = nullptr; }
            c32 *dT = (c32 *)F->dT.p, *dC = (c32 *)h->sC.p;
            unsigned long long *keys = (unsigned long long *)F->dKeys.p;
            DetDev *det_alias = (DetDev *)host_dev_alias(e.det);
            // Two branches behind the range transform: the spectra's copy out with the entry's t_done recorded behind it
            // (an external event-record node: pending from the graph launch on, fired when the copy is through), and angle
            // fft_vcc (it zero-pads its V-sample rows itself, as it does for the separate block) + estimator; joined at the end.
            CU(cudaStreamBeginCapture(F->stream, cudaStreamCaptureModeThreadLocal));
            std::swap(h->stream, F->stream);
            jrc_status st = [&]() -> jrc_status {
                ST(launch_fft_rows(h, dH, N, N, dT, Nr, V, 0, 0, V));                      // dT[n][v] = spectrum of channel v at bin n
                CU(cudaEventRecord(F->t_ready, h->stream));
                CU(cudaStreamWaitEvent(F->stream2, F->t_ready, 0));
                CU(cudaMemcpyAsync(e.T, dT, tcells * sizeof(c32), cudaMemcpyDeviceToHost, F->stream2));
                CU(cudaEventRecordWithFlags(e.t_done, F->stream2, cudaEventRecordExternal));
                CU(cudaEventRecord(F->t_join, F->stream2));
                ST(launch_fft_rows(h, dT, V, V, dC, Na, Nr, 1, 1, 0, keys));
                ST(launch_estimate(h, dC, Nr, Na, 1, 0, det_alias ? det_alias : (DetDev *)F->dDet.p, keys));
                if (!det_alias) CU(cudaMemcpyAsync(e.det, F->dDet.p, sizeof(DetDev), cudaMemcpyDeviceToHost, h->stream));
                CU(cudaStreamWaitEvent(h->stream, F->t_join, 0));
                return JRC_OK;
            }();
            std::swap(h->stream, F->stream);
            cudaGraph_t g = nullptr;
            cudaError_t ce = cudaStreamEndCapture(F->stream, &g);
            if (st != JRC_OK) { if (g) cudaGraphDestroy(g); return st; }
            CU(ce);
            ce = cudaGraphInstantiate(&e.graph, g, 0);
            cudaGraphDestroy(g);
            CU(ce);
            e.epoch = h->est_epoch;
        }
        fused_entry_pending = &e;
        fused_seq = seq;
    }
    // The continuation depends on the estimate only (h_ready is recorded here).  Its launch is the longest host-side step
    // of the call: this call's own small output kernel is queued first and completes while the host is busy with the graph
    // launch, so the synchronize below returns at once.
    if (fused_entry_pending) CU(cudaEventRecord(F->h_ready, h->stream));
    auto launch_continuation = [&]() -> jrc_status {
        if (!fused_entry_pending) return JRC_OK;
        FusedEntry &e = *fused_entry_pending;
        fused_entry_pending = nullptr;
        CU(cudaStreamWaitEvent(F->stream, F->h_ready, 0));
        CU(cudaGraphLaunch(e.graph, F->stream));
        CU(cudaEventRecord(e.done, F->stream));
        h->launches += 5;
        F->last_done = e.done;
        e.seq.store(fused_seq, std::memory_order_release);
        *cpi_seq = fused_seq;
        return JRC_OK;
    };
    const size_t out_items = (size_t)V * Nr, est_items = (size_t)V * N;
    if (!ptr_is_device(out) && !host_ptr_is_pinned(out) && h->zero_copy) {
        // pageable output (a scheduler buffer): only the estimate crosses the bus (V*N samples stored in place into a
        // page-locked block: the packet is the estimate's rows followed by zeros, lib/mimo_ofdm_radar_impl.cc:300-312), the
        // range zero-padding is written by the host straight into the caller's buffer
        ST(h->pin_b.need(est_items * sizeof(c32)));
        c32 *pin = (c32 *)h->pin_b.p, *dpin = (c32 *)host_dev_alias(pin);
        if (dpin) {
            k_pad_rows<<<grid_for((long long)est_items, 256, h->sm_count), 256, 0, h->stream>>>(dH, dpin, V, N, N, nullptr);
            CU(cudaGetLastError());
            h->launches++;
            ST(launch_continuation());
            CU(cudaStreamSynchronize(h->stream));
            for (int v = 0; v < V; v++) {
                memcpy(out + (size_t)v * Nr, pin + (size_t)v * N, (size_t)N * sizeof(c32));
                memset(out + (size_t)v * Nr + N, 0, (size_t)(Nr - N) * sizeof(c32));
            }
            if (chan_est_host) memcpy(chan_est_host, pin, est_items * sizeof(c32));
            return JRC_OK;
        }
    }
    ST(launch_continuation());
    void *dout = nullptr;
    ST(sg.out(out, out_items * sizeof(c32), &dout));
    k_pad_rows<<<grid_for((long long)out_items, 256, h->sm_count), 256, 0, h->stream>>>(dH, (c32 *)dout, V, N, Nr, nullptr);
    CU(cudaGetLastError());
    h->launches++;
    if (chan_est_host) CU(cudaMemcpyAsync(chan_est_host, dH, est_items * sizeof(c32), cudaMemcpyDeviceToHost, h->stream));
    return sg.finish();
}

extern "C" jrc_status jrc_radar_estimate(jrc_chain *h, const jrc_c32 *const *tx, const jrc_c32 *const *rx,
                                          size_t tx_skip_items, jrc_c32 *out, jrc_c32 *chan_est_host)
{
    return radar_estimate_impl(h, tx, rx, tx_skip_items, out, chan_est_host, nullptr);
}

extern "C" jrc_status jrc_radar_estimate_fused(jrc_chain *h, const jrc_c32 *const *tx, const jrc_c32 *const *rx,
                                                size_t tx_skip_items, jrc_c32 *out, jrc_c32 *chan_est_host, int64_t *cpi_seq)
{
    if (!cpi_seq) return fail(JRC_ERR_INVALID, "null argument");
    return radar_estimate_impl(h, tx, rx, tx_skip_items, out, chan_est_host, cpi_seq);
}

// Both fetches run on the downstream blocks' threads while the radar block's thread keeps submitting: an entry is
// valid for a sequence number if it carries that number before AND after the copy.
static jrc_status fused_entry(jrc_chain *h, int64_t cpi_seq, bool transposed_only, FusedEntry **out)
{
    if (!h) return fail(JRC_ERR_INVALID, "null argument");
    jrc_fused_state *F = h->fstate;
    if (!F || cpi_seq < 0) return fail(JRC_ERR_STATE, "CPI %lld is not cached", (long long)cpi_seq);
    FusedEntry &e = F->e[cpi_seq % JRC_FUSED_RING];
    if (e.seq.load(std::memory_order_acquire) != cpi_seq) return fail(JRC_ERR_STATE, "CPI %lld is not cached", (long long)cpi_seq);
    CU(cudaSetDevice(h->cfg.device));
    if (!transposed_only) CU(cudaEventSynchronize(e.done));
    *out = &e;
    return JRC_OK;
}

extern "C" jrc_status jrc_fused_fetch_transposed(jrc_chain *h, int64_t cpi_seq, jrc_c32 *out)
{
    if (!out) return fail(JRC_ERR_INVALID, "null argument");
    FusedEntry *e = nullptr;
    ST(fused_entry(h, cpi_seq, true, &e));
    // row n of the [Nr][Na] array: the V spectra samples, then zeros (lib/matrix_transpose_impl.cc:91-104) -- 1/interp_angle
    // of the array is read (from the entry), the rest is stores into the caller's buffer
    const int V = h->V, Na = h->Na, Nr = h->Nr;
    // The zeros do not depend on the frame: they are written while the spectra are still on their way (the range transform
    // starts when the radar block's estimate is there, i.e. about when that block returns), one memset of the whole array;
    // then the wait, then the short data runs.
    c32 *o = (c32 *)out;
    const c32 *t = e->T;
    if (Na > V) memset(o, 0, (size_t)Nr * Na * sizeof(c32));
    CU(cudaEventSynchronize(e->t_done));
    for (int n = 0; n < Nr; n++, o += Na, t += V)
        for (int v = 0; v < V; v++) o[v] = t[v];
    if (e->seq.load(std::memory_order_acquire) != cpi_seq) return fail(JRC_ERR_STATE, "CPI %lld was overwritten", (long long)cpi_seq);
    return JRC_OK;
}

extern "C" jrc_status jrc_fused_fetch_det(jrc_chain *h, int64_t cpi_seq, float snr_threshold, float power_threshold, jrc_det *det)
{
    if (!det) return fail(JRC_ERR_INVALID, "null argument");
    FusedEntry *e = nullptr;
    ST(fused_entry(h, cpi_seq, false, &e));
    memcpy(det, e->det, sizeof(DetDev));
    if (e->seq.load(std::memory_order_acquire) != cpi_seq) return fail(JRC_ERR_STATE, "CPI %lld was overwritten", (long long)cpi_seq);
    if (det->range_idx >= 0) {
        // final scalar of lib/range_angle_estimator_impl.cc:227,234 with the host libm, like jrc_estimate2d
        det->snr_db = 10 * std::log10(det->peak_power / det->noise_power);
        det->flags = (det->snr_db >= snr_threshold && det->peak_power >= power_threshold) ? JRC_DET_PASSED : 0u;
    }
    return JRC_OK;
}

extern "C" jrc_status jrc_fft_vcc(jrc_chain *h, const jrc_c32 *in, jrc_c32 *out, int32_t n, int32_t batch,
                                   int32_t forward, int32_t shift)
{
    if (!h || !in || !out) return fail(JRC_ERR_INVALID, "null argument");
    if (batch <= 0) return JRC_OK;
    CU(cudaSetDevice(h->cfg.device));
    Staging sg(h);
    const void *din = nullptr; void *dout = nullptr;
    size_t bytes = (size_t)n * batch * sizeof(c32);
    ST(sg.in(in, bytes, &din));
    ST(sg.out(out, bytes, &dout));
    ST(launch_fft_rows(h, (const c32 *)din, n, n, (c32 *)dout, n, batch, forward, shift));
    return sg.finish();
}

extern "C" jrc_status jrc_transpose_pad(jrc_chain *h, const jrc_c32 *in, int32_t k_items, int32_t input_len,
                                         int32_t output_len, int32_t interp, jrc_c32 *out)
{
    if (!h || !in || !out) return fail(JRC_ERR_INVALID, "null argument");
    if (k_items < 0 || input_len < 1 || output_len < 1 || interp < 1) return fail(JRC_ERR_INVALID, "bad sizes");
    // lib/matrix_transpose_impl.cc:82-83: the packet must hold whole columns
    if (k_items > output_len * interp) return fail(JRC_ERR_INVALID, "input_len and output_len do not match to packet length");
    CU(cudaSetDevice(h->cfg.device));
    Staging sg(h);
    const void *din = nullptr; void *dout = nullptr;
    ST(sg.in(in, (size_t)k_items * input_len * sizeof(c32), &din));
    ST(sg.out(out, (size_t)input_len * output_len * interp * sizeof(c32), &dout));
    ST(launch_transpose(h, (const c32 *)din, (c32 *)dout, k_items, input_len, output_len * interp, 1));
    return sg.finish();
}

extern "C" jrc_status jrc_mag_squared(jrc_chain *h, const jrc_c32 *in, float *out, size_t n)
{
    if (!h || !in || !out) return fail(JRC_ERR_INVALID, "null argument");
    if (n == 0) return JRC_OK;
    CU(cudaSetDevice(h->cfg.device));
    Staging sg(h);
    const void *din = nullptr; void *dout = nullptr;
    ST(sg.in(in, n * sizeof(c32), &din));
    ST(sg.out(out, n * sizeof(float), &dout));
    k_mag_squared<<<grid_for((long long)n, 256, h->sm_count), 256, 0, h->stream>>>((const c32 *)din, (float *)dout, (long long)n);
    CU(cudaGetLastError());
    h->launches++;
    return sg.finish();
}

// ---------------------------------------------------------------------------
// target_simulator (lib/target_simulator_impl.cc:127-385)
// ---------------------------------------------------------------------------
static jrc_status fft_any(jrc_chain *h, const c32 *in, c32 *out, int n, long long rows, int forward)
{
    if (is_pow2(n) && n <= 16384)      // the staged radix-2 kernel: bit-identical to the CPU restatement
        return launch_fft_rows(h, in, n, n, out, n, rows, forward, 0);
    auto key = std::make_pair(n, forward ? 1 : 0);
    double2 *tab = nullptr;
    auto it = h->dft_tabs.find(key);
    if (it != h->dft_tabs.end()) tab = it->second;
    else {
        std::vector<double2> t((size_t)n);
        const double sgn = forward ? -1.0 : 1.0;
        for (int k = 0; k < n; k++) {
            double a = sgn * 2.0 * M_PI * (double)k / (double)n;
            t[k].x = cos(a); t[k].y = sin(a);
        }
        CU(cudaMalloc(&tab, t.size() * sizeof(double2)));
        CU(cudaMemcpyAsync(tab, t.data(), t.size() * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        h->dft_tabs[key] = tab;
    }
    dim3 grid((unsigned)((n + 127) / 128), (unsigned)rows);
    k_dft_any<<<grid, 128, 0, h->stream>>>(in, out, n, tab);
    CU(cudaGetLastError());
    h->launches++;
    return JRC_OK;
}

extern "C" jrc_status jrc_target_sim(jrc_chain *h, const jrc_c32 *in, int32_t n, const float *range, const float *velocity,
                                      const float *rcs, const float *azimuth, int32_t n_targets, const float *position_rx,
                                      int32_t n_rx, int32_t samp_rate, float center_freq, int32_t self_coupling,
                                      float self_coupling_db, const jrc_c32 *target_phase, int32_t accumulate, jrc_c32 *out)
{
    if (!h || !in || !out || !range || !velocity || !rcs || !azimuth || !position_rx) return fail(JRC_ERR_INVALID, "null argument");
    if (n < 1 || n_targets < 0 || n_rx < 1 || samp_rate <= 0) return fail(JRC_ERR_INVALID, "bad sizes");
    if ((long long)n_rx * (n_targets > 0 ? n_targets : 1) > 65535) return fail(JRC_ERR_INVALID, "too many (target, rx) pairs");
    CU(cudaSetDevice(h->cfg.device));
    const int K = n_targets, L = n_rx;
    // channel filters exactly as the reference builds them (float arithmetic, :164-188, :264-303)
    const float c_light = 3e8f;
    const double FOUR_PI_CUBED_SQRT = 44.54662397465366;
    std::vector<c32> filt((size_t)(K + K * L) * n + (size_t)K);    // [K] doppler, [L][K] time shift, [K] phase
    std::vector<float> freq((size_t)n);
    for (int i = 0; i < n; i++)
        freq[i] = i < n / 2 ? i * (float)samp_rate / (float)n : i * (float)samp_rate / (float)n - (float)samp_rate;
    for (int k = 0; k < K; k++) {
        const float doppler = 2 * velocity[k] * center_freq / c_light;
        const float scale = (float)(c_light * sqrtf(rcs[k]) / FOUR_PI_CUBED_SQRT / (range[k] * range[k]) / center_freq);
        float ph = 0.0f;
        c32 *fd = filt.data() + (size_t)k * n;
        for (int i = 0; i < n; i++) {
            fd[i].x = cosf(ph) * scale; fd[i].y = sinf(ph) * scale;
            ph = (float)fmod(ph + 2 * M_PI * doppler / (float)samp_rate, 2 * M_PI);
        }
        for (int l = 0; l < L; l++) {
            const float timeshift = (float)((2.0 * range[k] - position_rx[l] * sin(azimuth[k] * M_PI / 180.0)) / c_light);
            c32 *ft = filt.data() + ((size_t)K + (size_t)l * K + k) * n;
            for (int i = 0; i < n; i++) {
                const float pt = (float)fmod(2 * M_PI * (timeshift) * (freq[i] + center_freq), 2 * M_PI);
                ft[i].x = cosf(pt) / (float)n; ft[i].y = -sinf(pt) / (float)n;
            }
        }
    }
    c32 *ph_host = filt.data() + (size_t)(K + K * L) * n;
    for (int k = 0; k < K; k++) ph_host[k] = target_phase ? *(const c32 *)&target_phase[k] : make_float2(1.f, 0.f);
    Staging sg(h);
    const void *din = nullptr; void *dout = nullptr;
    ST(sg.in(in, (size_t)n * sizeof(c32), &din));
    ST(sg.out(out, (size_t)L * n * sizeof(c32), &dout));
    const size_t KL = (size_t)K * L;
    ST(h->sMisc.need(filt.size() * sizeof(c32)));
    ST(h->sMisc2.need(((size_t)2 * K + 2 * KL) * n * sizeof(c32) + 16));
    c32 *d_filt = (c32 *)h->sMisc.p;
    c32 *d_bt = (c32 *)h->sMisc2.p, *d_bf = d_bt + (size_t)K * n, *d_g = d_bf + (size_t)K * n, *d_res = d_g + KL * n;
    CU(cudaMemcpyAsync(d_filt, filt.data(), filt.size() * sizeof(c32), cudaMemcpyHostToDevice, h->stream));
    if (K > 0) {
        // in * doppler filter -> FFT                                            (:346-350), once per target
        k_cmul_rows<<<grid_for((long long)K * n, 256, h->sm_count), 256, 0, h->stream>>>((const c32 *)din, K, d_filt, d_bt, n, K);
        CU(cudaGetLastError());
        h->launches++;
        ST(fft_any(h, d_bt, d_bf, n, K, 1));
        // * time-shift filter of (rx l, target k) -> IFFT                      (:353-357)
        // rows ordered [l][k]: the spectrum of target k is row (row % K) -> use a_div trick per l
        for (int l = 0; l < L; l++) {
            k_cmul_rows<<<grid_for((long long)K * n, 256, h->sm_count), 256, 0, h->stream>>>(
                d_bf, 1, d_filt + ((size_t)K + (size_t)l * K) * n, d_g + (size_t)l * K * n, n, K);
            CU(cudaGetLastError());
            h->launches++;
        }
        ST(fft_any(h, d_g, d_res, n, (long long)KL, 0));
    }
    const float g = (float)pow(10, self_coupling_db / 20.0);
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)L);
    k_sim_combine<<<grid, 256, 0, h->stream>>>(d_res, (const c32 *)din, target_phase ? d_filt + (size_t)(K + K * L) * n : nullptr, n, K,
                                               accumulate, self_coupling, g, (c32 *)dout);
    CU(cudaGetLastError());
    h->launches++;
    return sg.finish();
}

extern "C" jrc_status jrc_scene_synth(jrc_chain *h, const jrc_c32 *tx, int32_t n_cpi, int32_t n_targets, const float *range_m,
                                       const float *az_deg, const float *amp, double samp_rate, double center_freq,
                                       float noise_sigma, uint64_t seed, jrc_c32 *rx_dev)
{
    if (!h || !tx || !range_m || !az_deg || !amp || !rx_dev) return fail(JRC_ERR_INVALID, "null argument");
    if (n_cpi < 0 || n_targets < 0) return fail(JRC_ERR_INVALID, "bad sizes");
    if (n_cpi == 0) return JRC_OK;
    if (!ptr_is_device(rx_dev)) return fail(JRC_ERR_INVALID, "rx must be device memory");
    const jrc_chain_cfg &c = h->cfg;
    if (c.n_tx > 16) return fail(JRC_ERR_INVALID, "at most 16 TX antennas");
    CU(cudaSetDevice(c.device));
    NvtxRange nv("jrc_scene_synth");
    Staging sg(h);
    const void *dtx = nullptr, *dr = nullptr, *da = nullptr, *dm = nullptr;
    const size_t np = (size_t)n_cpi * (n_targets > 0 ? n_targets : 1) * sizeof(float);
    ST(sg.in(tx, (size_t)c.n_tx * c.n_sym * c.fft_len * sizeof(c32), &dtx));
    ST(sg.in(range_m, np, &dr));
    ST(sg.in(az_deg, np, &da));
    ST(sg.in(amp, np, &dm));
    SceneParams P;
    memset(&P, 0, sizeof(P));
    P.tx = (const c32 *)dtx; P.range_m = (const float *)dr; P.az_deg = (const float *)da; P.amp = (const float *)dm;
    P.n_cpi = n_cpi; P.J = n_targets; P.T = c.n_tx; P.R = c.n_rx; P.S = c.n_sym; P.N = c.fft_len;
    P.samp_rate = samp_rate; P.center_freq = center_freq; P.noise_sigma = noise_sigma; P.seed = seed; P.rx = (c32 *)rx_dev;
    const long long total = (long long)n_cpi * c.n_rx * c.fft_len;
    const int grid = grid_for(total, 256, h->sm_count);
    if (c.n_tx <= 4) k_scene_synth<4><<<grid, 256, 0, h->stream>>>(P);
    else if (c.n_tx <= 8) k_scene_synth<8><<<grid, 256, 0, h->stream>>>(P);
    else k_scene_synth<16><<<grid, 256, 0, h->stream>>>(P);
    CU(cudaGetLastError());
    h->launches++;
    return sg.finish();
}

extern "C" jrc_status jrc_nlog10(jrc_chain *h, const float *in, float *out, size_t n_items, float n, float k)
{
    if (!h || !in || !out) return fail(JRC_ERR_INVALID, "null argument");
    if (n_items == 0) return JRC_OK;
    CU(cudaSetDevice(h->cfg.device));
    Staging sg(h);
    const void *din = nullptr; void *dout = nullptr;
    ST(sg.in(in, n_items * sizeof(float), &din));
    ST(sg.out(out, n_items * sizeof(float), &dout));
    k_nlog10<<<grid_for((long long)n_items, 256, h->sm_count), 256, 0, h->stream>>>((const float *)din, (float *)dout,
                                                                                   (long long)n_items, n, k);
    CU(cudaGetLastError());
    h->launches++;
    return sg.finish();
}

extern "C" jrc_status jrc_estimate2d(jrc_chain *h, const jrc_c32 *map, int32_t n_inputs, int32_t vlen, jrc_det *det)
{
    if (!h || !map || !det) return fail(JRC_ERR_INVALID, "null argument");
    if (n_inputs < 1 || vlen < 1) return fail(JRC_ERR_INVALID, "empty map");
    CU(cudaSetDevice(h->cfg.device));
    Staging sg(h);
    const void *din = nullptr;
    ST(sg.in(map, (size_t)n_inputs * vlen * sizeof(c32), &din));
    ST(h->sDet.need(sizeof(DetDev)));
    ST(launch_estimate(h, (const c32 *)din, n_inputs, vlen, 1, 0, (DetDev *)h->sDet.p));
    DetDev d;
    CU(cudaMemcpyAsync(&d, h->sDet.p, sizeof(d), cudaMemcpyDeviceToHost, h->stream));
    ST(sg.finish());
    memcpy(det, &d, sizeof(d));
    if (det->range_idx >= 0) {
        // final scalar of lib/range_angle_estimator_impl.cc:227,234 with the host libm so that
        // the published snr has the reference's bit pattern
        det->snr_db = 10 * std::log10(det->peak_power / det->noise_power);
        det->flags = (det->snr_db >= h->snr_thr && det->peak_power >= h->pow_thr) ? JRC_DET_PASSED : 0u;
    }
    return JRC_OK;
}

extern "C" jrc_status jrc_peak1d(jrc_chain *h, const jrc_c32 *in, int32_t n, int32_t samp_rate, float interp_factor,
                                  float threshold_db, int32_t samp_protect, jrc_peak1d_out *out)
{
    if (!h || !in || !out) return fail(JRC_ERR_INVALID, "null argument");
    if (n < 0) return fail(JRC_ERR_INVALID, "n < 0");
    CU(cudaSetDevice(h->cfg.device));
    out->k = -1; out->freq = 0.f; out->phase = 0.f; out->mag = 0.f;
    if (n == 0) return JRC_OK;
    Staging sg(h);
    const void *din = nullptr;
    ST(sg.in(in, (size_t)n * sizeof(c32), &din));
    ST(h->sKeys.need(sizeof(unsigned long long)));
    ST(h->sDet.need(sizeof(Peak1dDev)));
    CU(cudaMemsetAsync(h->sKeys.p, 0, sizeof(unsigned long long), h->stream));
    const double thr_lin = std::pow(10, threshold_db / 10.0);      // lib/fft_peak_detect_impl.cc:89
    k_peak1d_scan<<<grid_for(n, 256, h->sm_count), 256, 0, h->stream>>>((const c32 *)din, n, samp_protect, thr_lin,
                                                                      (unsigned long long *)h->sKeys.p);
    CU(cudaGetLastError());
    k_peak1d_finalize<<<1, 1, 0, h->stream>>>((const c32 *)din, (const unsigned long long *)h->sKeys.p, (Peak1dDev *)h->sDet.p);
    CU(cudaGetLastError());
    h->launches += 2;
    Peak1dDev d;
    CU(cudaMemcpyAsync(&d, h->sDet.p, sizeof(d), cudaMemcpyDeviceToHost, h->stream));
    ST(sg.finish());
    out->k = d.k;
    if (d.k != -1) {
        // bin -> Hz conversion and arg() are scalar host work, evaluated exactly as :98-107
        const int k = d.k;
        if (k <= n / 2) out->freq = k / (float)n * (samp_rate * interp_factor);
        else out->freq = -((float)samp_rate * interp_factor) + k * (samp_rate * interp_factor / (float)n);
        out->phase = std::atan2(d.z.y, d.z.x);
        out->mag = d.mag;
    }
    return JRC_OK;
}

extern "C" jrc_status jrc_cp_remove(jrc_chain *h, const jrc_c32 *in, int32_t n_sym, int32_t fft_len, int32_t cp_len, jrc_c32 *out)
{
    if (!h || !in || !out) return fail(JRC_ERR_INVALID, "null argument");
    if (n_sym < 0 || fft_len < 1 || cp_len < 0) return fail(JRC_ERR_INVALID, "bad sizes");
    if (n_sym == 0) return JRC_OK;
    CU(cudaSetDevice(h->cfg.device));
    Staging sg(h);
    const void *din = nullptr; void *dout = nullptr;
    ST(sg.in(in, (size_t)n_sym * (fft_len + cp_len) * sizeof(c32), &din));
    ST(sg.out(out, (size_t)n_sym * fft_len * sizeof(c32), &dout));
    // strided copy: symbol k of the packet starts at k*(fft_len+cp_len) + cp_len  (:92-95)
    CU(cudaMemcpy2DAsync(dout, (size_t)fft_len * sizeof(c32), (const c32 *)din + cp_len, (size_t)(fft_len + cp_len) * sizeof(c32),
                         (size_t)fft_len * sizeof(c32), (size_t)n_sym, cudaMemcpyDeviceToDevice, h->stream));
    return sg.finish();
}

extern "C" jrc_status jrc_ofdm_demod(jrc_chain *h, const jrc_c32 *in, int32_t n_sym, int32_t fft_len, int32_t cp_len, jrc_c32 *out)
{
    if (!h || !in || !out) return fail(JRC_ERR_INVALID, "null argument");
    if (n_sym < 0 || fft_len < 1 || cp_len < 0) return fail(JRC_ERR_INVALID, "bad sizes");
    if (n_sym == 0) return JRC_OK;
    CU(cudaSetDevice(h->cfg.device));
    Staging sg(h);
    const void *din = nullptr; void *dout = nullptr;
    ST(sg.in(in, (size_t)n_sym * (fft_len + cp_len) * sizeof(c32), &din));
    ST(sg.out(out, (size_t)n_sym * fft_len * sizeof(c32), &dout));
    // the cyclic-prefix removal is the row stride of the FFT's input: nothing is copied
    ST(launch_fft_rows(h, (const c32 *)din + cp_len, fft_len + cp_len, fft_len, (c32 *)dout, fft_len, n_sym, 1, 1));
    return sg.finish();
}

extern "C" jrc_status jrc_zero_pad(jrc_chain *h, const jrc_c32 *in, int32_t n, uint32_t pad_front, uint32_t pad_tail,
                                    uint64_t seed, jrc_c32 *out)
{
    if (!h || !out || (!in && n > 0)) return fail(JRC_ERR_INVALID, "null argument");
    if (n < 0) return fail(JRC_ERR_INVALID, "n < 0");
    CU(cudaSetDevice(h->cfg.device));
    const size_t total = (size_t)n + pad_front + pad_tail;
    if (total == 0) return JRC_OK;
    Staging sg(h);
    const void *din = nullptr; void *dout = nullptr;
    if (n > 0) ST(sg.in(in, (size_t)n * sizeof(c32), &din));
    ST(sg.out(out, total * sizeof(c32), &dout));
    k_zero_pad<<<grid_for((long long)total, 256, h->sm_count), 256, 0, h->stream>>>((const c32 *)din, n, pad_front, pad_tail,
                                                                                   (unsigned long long)seed, (c32 *)dout);
    CU(cudaGetLastError());
    h->launches++;
    return sg.finish();
}
