/*
 * jrc_cuda.h -- C ABI of libjrc_cuda.so, the B200 (sm_100a) implementation of the
 * gr-mimo-ofdm-jrc radar range-angle hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * The GNU Radio block wrappers (gr-mimo-ofdm-jrc_b200/lib/<block>_impl.cc) call
 * ONLY these entry points from their work() functions; INTEGRATION.md shows the
 * binding a maintainer of the reference adds.  Paths cited below are relative to
 * the reference tree (ceyhunozkaptan/gr-mimo-ofdm-jrc).
 *
 * Conventions
 *   - every function returns jrc_status; jrc_last_error() gives the message of the
 *     last failure on the calling thread.  No exceptions cross this boundary; the
 *     C++ wrappers translate non-zero status to std::runtime_error (the reference's
 *     own error convention, e.g. lib/matrix_transpose_impl.cc:82-83).
 *   - a handle owns one CUDA stream, its device scratch and pinned staging buffers.
 *     Handles are independent; one handle must not be used from two threads at once
 *     (GNU Radio never calls a block's work() concurrently with itself).
 *   - "complex" is interleaved float32 (re, im) == gr_complex == std::complex<float>.
 *   - there is NO CPU fallback: without a usable CUDA device every call fails.
 */
#ifndef JRC_CUDA_H
#define JRC_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define JRC_API
#else
#define JRC_API __attribute__((visibility("default")))
#endif

typedef int32_t jrc_status;
enum {
    JRC_OK = 0,
    JRC_ERR_INVALID = 1,     /* bad argument / unsupported size            */
    JRC_ERR_CUDA = 2,        /* CUDA runtime error (message has the detail) */
    JRC_ERR_NO_DEVICE = 3,   /* no CUDA device: the path refuses to run     */
    JRC_ERR_STATE = 4        /* call order / missing configuration          */
};

typedef struct { float re, im; } jrc_c32;   /* gr_complex */

typedef struct jrc_chain jrc_chain;         /* opaque handle */

/* Construction parameters == the make() arguments of mimo_ofdm_radar
 * (include/mimo_ofdm_jrc/mimo_ofdm_radar.h:48-60) plus the angle zero-pad of
 * matrix_transpose (include/mimo_ofdm_jrc/matrix_transpose.h:48).            */
typedef struct {
    int32_t fft_len;               /* subcarriers (power of two)                 */
    int32_t n_tx, n_rx, n_sym, n_pre;
    int32_t interp_range;          /* mimo_ofdm_radar interp_factor  (Nr = fft_len*interp_range) */
    int32_t interp_angle;          /* matrix_transpose interp_factor (Na = n_tx*n_rx*interp_angle) */
    int32_t tx_interleave;         /* enable_tx_interleave (lib/mimo_ofdm_radar_impl.cc:262-269) */
    int32_t background_removal;    /* lib/mimo_ofdm_radar_impl.cc:281-292         */
    int32_t background_recording;  /* lib/mimo_ofdm_radar_impl.cc:276-279         */
    int32_t record_len;            /* ring-buffer capacity (:115)                 */
    int32_t device;                /* CUDA device ordinal                         */
} jrc_chain_cfg;

/* Detection record: what range_angle_estimator publishes per CPI
 * (lib/range_angle_estimator_impl.cc:137-253), 32 bytes.                     */
typedef struct {
    int32_t  range_idx;     /* peak_range_idx                                 */
    int32_t  angle_idx;     /* peak_angle_idx                                 */
    float    peak_power;    /* (float)pow(abs(z),2) at the peak               */
    float    noise_power;   /* mean of pow(abs(z),2) over the noise window    */
    float    snr_db;        /* 10*log10(peak/noise)                           */
    int32_t  n_noise;       /* n_noise_samples                                */
    uint32_t flags;         /* JRC_DET_PASSED | JRC_DET_EXACT                 */
    int32_t  cpi;           /* sequence number: cpi0 + index in the batch     */
} jrc_det;

#define JRC_DET_PASSED 1u   /* snr >= snr_threshold && peak >= power_threshold (:234): the "params" message goes out */
/* The fused and tiled kernels decide on their own float32 map, which differs from the reference-order (staged)
 * arithmetic by FFT rounding (~3e-7 of the map peak).  A record whose arg-max or gate lies inside that margin is
 * redone in the reference's order (the staged kernels' float operations, the sequential window sum): the peak
 * indices and JRC_DET_PASSED of EVERY record equal the staged path's, and a record carrying JRC_DET_EXACT is
 * bit-identical to it in every field.                                            */
#define JRC_DET_EXACT 2u

/* fft_peak_detect outputs (lib/fft_peak_detect_impl.cc:98-107); k == -1: no
 * sample passed the threshold (the reference then leaves its outputs unwritten). */
typedef struct { int32_t k; float freq, phase, mag; } jrc_peak1d_out;

/* Input layout of a batch of CPIs, in units of complex samples.  Element
 * (cpi, antenna a, LTF symbol s, subcarrier k) of a port lives at
 *     base + cpi*cpi_stride + a*ant_stride + (n_pre + s)*fft_len + k
 * i.e. every antenna row is a GNU Radio packet of fft_len-vectors starting at
 * item 0 of the frame (lib/mimo_ofdm_radar_impl.cc:254-260).  tx.cpi_stride == 0
 * shares one TX frame between all CPIs of the batch.                          */
typedef struct { const jrc_c32 *base; int64_t cpi_stride; int64_t ant_stride; } jrc_port_layout;

/* ---- life cycle ---------------------------------------------------------- */
JRC_API jrc_status jrc_chain_create(const jrc_chain_cfg *cfg, jrc_chain **out);
JRC_API void       jrc_chain_destroy(jrc_chain *h);
JRC_API const char *jrc_last_error(void);
JRC_API int32_t    jrc_abi_version(void);
/* the handle's stream as a cudaStream_t (so callers can order their own work) */
JRC_API void      *jrc_chain_stream(jrc_chain *h);
JRC_API jrc_status jrc_chain_sync(jrc_chain *h);

/* range_angle_estimator::make parameters
 * (include/mimo_ofdm_jrc/range_angle_estimator.h:48-58).  range_bins/angle_bins are
 * HOST arrays of n_range/n_angle floats (copied).                             */
JRC_API jrc_status jrc_chain_set_estimator(jrc_chain *h,
                                           const float *range_bins, int32_t n_range,
                                           const float *angle_bins, int32_t n_angle,
                                           float noise_discard_range_m,
                                           float noise_discard_angle_deg,
                                           float snr_threshold, float power_threshold);
/* GRC callbacks: set_snr_threshold / set_power_threshold
 * (lib/range_angle_estimator_impl.cc:286-287), set_background_record
 * (lib/mimo_ofdm_radar_impl.cc:342-346)                                       */
JRC_API jrc_status jrc_chain_set_thresholds(jrc_chain *h, float snr_threshold, float power_threshold);
JRC_API jrc_status jrc_chain_set_background_record(jrc_chain *h, int32_t on);
/* clears the background ring buffer (a freshly constructed block)            */
JRC_API jrc_status jrc_chain_reset_background(jrc_chain *h);

/* ---- the fused chain (device-resident batch) ------------------------------ *
 * mimo_ofdm_radar -> fft_vcc(IFFT) -> matrix_transpose -> fft_vcc(FFT,shift) ->
 * {complex_to_mag_squared, range_angle_estimator} for n_cpi independent CPIs.
 * All pointers are DEVICE pointers; work is enqueued on the handle's stream and
 * the call returns without synchronising.
 *   map   : [n_cpi][Nr][Na] float32 |.|^2, or NULL
 *   cmap  : [n_cpi][Nr][Na] complex map (what the estimator block would see), or NULL
 *   dets  : [n_cpi] records, or NULL (needs jrc_chain_set_estimator first)
 * path: JRC_PATH_AUTO picks the fused single-kernel path when the configuration
 * has one (fft_len 64, 8 virtual channels, cmap == NULL), else the tiled kernels
 * (power-of-two Nr in 64..8192 and Na in 64..2048, cmap == NULL: radix-8 FFT kernels with
 * the transpose, |.|^2 and arg-max fused into the angle FFT), else the staged
 * kernels (one per reference block, bit-identical to the CPU restatement). */
enum { JRC_PATH_AUTO = 0, JRC_PATH_FUSED = 1, JRC_PATH_STAGED = 2, JRC_PATH_TILED = 3 };
JRC_API jrc_status jrc_chain_run_batch(jrc_chain *h, jrc_port_layout rx, jrc_port_layout tx,
                                       int32_t n_cpi, int32_t cpi0,
                                       float *map, jrc_c32 *cmap, jrc_det *dets, int32_t path);
/* The same chain fed with the RX antennas' raw TIME samples: in the flowgraph
 * ofdm_cyclic_prefix_remover (lib/ofdm_cyclic_prefix_remover_impl.cc:62-99) and fft_vxx(forward, shift)
 * (...radar_sim.grc:898-939) sit between the receiver and the radar block's rx ports.  Sample i of symbol s of
 * antenna a of a CPI lives at
 *     rx_time.base + cpi*cpi_stride + a*ant_stride + s*(fft_len + cp_len) + i,   s < n_pre + n_sym
 * (strides in complex samples; the n_pre preamble symbols are skipped, not transformed).  tx stays in the frequency
 * domain, as in the flowgraph's TX branch.  The demodulation runs in front of the chain on the same stream, in
 * chunks whose symbols stay in the L2, and is bit-identical to jrc_ofdm_demod symbol by symbol; everything else as
 * jrc_chain_run_batch.  Not for handles with background removal (per-frame state).                           */
JRC_API jrc_status jrc_chain_run_batch_time(jrc_chain *h, jrc_port_layout rx_time, int32_t cp_len,
                                            jrc_port_layout tx, int32_t n_cpi, int32_t cpi0,
                                            float *map, jrc_c32 *cmap, jrc_det *dets, int32_t path);
/* Range-Doppler-angle cube of a burst of n_burst consecutive CPIs (power of two): the complex range-angle maps of the
 * burst (one-kernel-per-block path, bit-identical to the CPU restatement) followed by a forward, fftshifted FFT along
 * slow time for every (range, angle) cell and |.|^2.  Device pointers, asynchronous on the handle's stream.
 *   cube : [Nr][Na][n_burst] float32, Doppler bin fastest, zero Doppler at n_burst / 2
 * Not part of the reference (its chain ends at one map per CPI); SURVEY.md 8(f) rank 4.                     */
JRC_API jrc_status jrc_chain_run_burst(jrc_chain *h, jrc_port_layout rx, jrc_port_layout tx, int32_t n_burst, float *cube);

/* which path the last run_batch took (JRC_PATH_FUSED / JRC_PATH_TILED / JRC_PATH_STAGED) and how
 * many kernels it launched                                                    */
JRC_API int32_t    jrc_chain_last_path(const jrc_chain *h);
JRC_API int64_t    jrc_chain_launch_count(const jrc_chain *h);

/* Same chain with HOST buffers (packed layout rx[n_cpi][n_rx][n_sym][fft_len],
 * tx[n_cpi or 1][n_tx][n_sym][fft_len], i.e. n_pre symbols already stripped):
 * pinned double-buffered H2D -> kernels -> D2H, synchronous on return.
 * Up to 4 CPIs in front of the fused kernel take the latency path instead: the
 * kernel reads the (pinned, device-mapped) host symbols and writes map and
 * records straight into host memory -- no copies, one launch (pageable caller
 * buffers go through the handle's pinned staging buffers).
 * map_host / dets_host may be NULL.                                           */
JRC_API jrc_status jrc_chain_run_host(jrc_chain *h, const jrc_c32 *rx_host, const jrc_c32 *tx_host,
                                      int32_t tx_shared, int32_t n_cpi, int32_t cpi0,
                                      float *map_host, jrc_det *dets_host);

/* Running totals of the reference-order pass behind JRC_DET_EXACT since the handle was created: out[0] records the
 * fast kernels marked, out[1] records redone in the reference's order, out[2] arg-max ties settled inside the fused
 * kernel.  Synchronises the handle.                                                                       */
JRC_API jrc_status jrc_chain_exact_stats(jrc_chain *h, int64_t *out);

/* Streaming form of jrc_chain_run_host (BASELINE configs[3], "pinned-host double-buffering"): submit enqueues the
 * chain for n_cpi CPIs and returns at once; up to 4 submissions are in flight, each on its own stream, so the input
 * transfer and kernel of CPI k+1 overlap the output transfer of CPI k.  The buffers must stay valid until
 * jrc_chain_wait(ticket) returns (which also finalises the records); pageable buffers are staged through pinned
 * memory of the handle, pinned ones (cudaHostAlloc / cudaHostRegister, e.g. a registered GNU Radio stream buffer)
 * are read and written in place.  jrc_chain_poll is the non-blocking test.  This is the scheduler hand-off of
 * lib/mimo_ofdm_radar_impl.cc:131-340 made asynchronous: general_work() submits frame k and publishes frame k-1. */
JRC_API jrc_status jrc_chain_submit(jrc_chain *h, const jrc_c32 *rx_host, const jrc_c32 *tx_host,
                                    int32_t tx_shared, int32_t n_cpi, int32_t cpi0,
                                    float *map_host, jrc_det *dets_host, int64_t *ticket);
JRC_API jrc_status jrc_chain_poll(jrc_chain *h, int64_t ticket, int32_t *done);
JRC_API jrc_status jrc_chain_wait(jrc_chain *h, int64_t ticket);

/* Page-locked host memory for callers that do not link the CUDA runtime themselves: allocate (jrc_pinned_alloc) or
 * page-lock an existing buffer (jrc_host_register, e.g. the stream buffers a block sees in start()).  Pinned buffers
 * are what makes jrc_chain_run_host / jrc_chain_submit copy-free.                                         */
JRC_API jrc_status jrc_pinned_alloc(size_t bytes, void **out);
JRC_API jrc_status jrc_pinned_free(void *p);
JRC_API jrc_status jrc_host_register(void *p, size_t bytes);
JRC_API jrc_status jrc_host_unregister(void *p);

/* Multi-GPU detection table without a collective on the critical path: the host rank allocates the table
 * (jrc_dev_alloc) and exports it (jrc_ipc_export, 64-byte handle passed to the other processes by any means); every other
 * rank maps it (jrc_ipc_open) and hands its slice as `dets` to jrc_chain_run_batch: the kernels store the 32-byte records
 * straight into the host rank's memory over NVLink.  After the ranks have synchronised their streams and met at a barrier
 * the table is complete ("NCCL over NVLink used only to gather detections" -- here the gather is the
 * producing kernel's own stores).                                                                               */
JRC_API jrc_status jrc_dev_alloc(int32_t device, size_t bytes, void **out);
JRC_API jrc_status jrc_dev_free(void *p);
JRC_API jrc_status jrc_dev_copy(void *dst, const void *src, size_t bytes);      /* synchronous, any direction */
/* the same copy queued on the handle's stream, behind the batches already submitted: the other way to fill a mapped
 * table -- records in local memory step by step, ONE peer copy of all of them at the drain (a kernel that stores its
 * records over NVLink itself waits for those stores when it ends: measured 1.6 % of a configs[1] step) */
JRC_API jrc_status jrc_chain_copy_async(jrc_chain *h, void *dst, const void *src, size_t bytes);
JRC_API jrc_status jrc_ipc_export(void *dev_ptr, void *handle64);
JRC_API jrc_status jrc_ipc_open(const void *handle64, int32_t device, void **out);
JRC_API jrc_status jrc_ipc_close(void *p);

/* ---- per-block stage calls (exact per-block semantics; pointers may be host
 * or device, detected with cudaPointerGetAttributes; host buffers are staged
 * through the handle's pinned memory and the call is synchronous) ------------ */

/* mimo_ofdm_radar::general_work body (lib/mimo_ofdm_radar_impl.cc:243-315) for ONE
 * frame: tx[n_tx], rx[n_rx] are per-port packet pointers (item 0 of the frame),
 * tx_skip_items = n_tx_samples_discard (:189-197).  out: [V][fft_len*interp_range]
 * zero-padded.  Updates the background ring buffer exactly like the block.
 * chan_est_host (optional, host) receives radar_chan_est [V][fft_len] for
 * capture_radar_data (:348-370).                                              */
JRC_API jrc_status jrc_radar_estimate(jrc_chain *h, const jrc_c32 *const *tx, const jrc_c32 *const *rx,
                                      size_t tx_skip_items, jrc_c32 *out, jrc_c32 *chan_est_host);

/* ---- fused mode for an UNMODIFIED flowgraph: in the shipped graph
 * (examples/simulation/radar/mimo_ofdm_jrc_radar_sim.grc:2165-2232) mimo_ofdm_radar, matrix_transpose and
 * range_angle_estimator are three blocks with two stock fft_vcc blocks between them.  On a handle that carries the
 * WHOLE chain's configuration (interp_angle and jrc_chain_set_estimator), jrc_radar_estimate_fused() is
 * jrc_radar_estimate() -- same outputs, same background ring -- and, without the frame leaving the device, the
 * rest of the chain in the one-kernel-per-block arithmetic (range fft_vcc, transpose, angle fft_vcc, estimator) on
 * a second stream.  The transposed array and the detection record land in a ring of JRC_FUSED_RING
 * entries under *cpi_seq (0, 1, 2, ... per handle); the call returns as soon as its own output is on the host.
 * Of the transposed array only the V data columns are kept (the rest is matrix_transpose's zero padding):
 * jrc_fused_fetch_transposed writes the zeros into the caller's buffer itself.
 * The downstream blocks fetch by sequence number (the jrc_cpi stream tag) from their own threads instead of
 * repeating the work: JRC_ERR_STATE if that CPI is not cached (never made, or overwritten by a newer one). */
#define JRC_FUSED_RING 16
JRC_API jrc_status jrc_radar_estimate_fused(jrc_chain *h, const jrc_c32 *const *tx, const jrc_c32 *const *rx,
                                            size_t tx_skip_items, jrc_c32 *out, jrc_c32 *chan_est_host,
                                            int64_t *cpi_seq);
/* matrix_transpose's output for that CPI: [Nr][Na] complex, zero-padded (lib/matrix_transpose_impl.cc:91-104) */
JRC_API jrc_status jrc_fused_fetch_transposed(jrc_chain *h, int64_t cpi_seq, jrc_c32 *out);
/* range_angle_estimator's record for that CPI (lib/range_angle_estimator_impl.cc:141-234); the gate is evaluated
 * with the thresholds given here (the estimator block's current ones). */
JRC_API jrc_status jrc_fused_fetch_det(jrc_chain *h, int64_t cpi_seq, float snr_threshold, float power_threshold,
                                       jrc_det *det);

/* gr::fft::fft_vcc (GNU Radio 3.8 gr-fft; examples/simulation/radar/
 * mimo_ofdm_jrc_radar_sim.grc:940-985): batch items of length n (power of two,
 * <= 16384), rectangular window, no scaling.                                  */
JRC_API jrc_status jrc_fft_vcc(jrc_chain *h, const jrc_c32 *in, jrc_c32 *out,
                               int32_t n, int32_t batch, int32_t forward, int32_t shift);

/* matrix_transpose::work (lib/matrix_transpose_impl.cc:97-104):
 * in [k_items][input_len] -> out [input_len][output_len*interp] zero-filled   */
JRC_API jrc_status jrc_transpose_pad(jrc_chain *h, const jrc_c32 *in, int32_t k_items,
                                     int32_t input_len, int32_t output_len, int32_t interp,
                                     jrc_c32 *out);

/* blocks_complex_to_mag_squared (...radar_sim.grc:637-652)                    */
JRC_API jrc_status jrc_mag_squared(jrc_chain *h, const jrc_c32 *in, float *out, size_t n);

/* target_simulator::work (lib/target_simulator_impl.cc:201-385) for one packet of n time samples (any n):
 * per RX antenna l and target k  IFFT( FFT(in * doppler_k) * timeshift_{l,k} ), the per-target results
 * folded as the reference does (accumulate = 0: its memcpy keeps the last target, :366) or summed
 * (accumulate = 1), optional per-target phase factor (rndm_phaseshift, :311-320; NULL = none), self
 * coupling (:372-378).  out: [n_rx][n].  The channel filters are built on the host with the reference's
 * float arithmetic (:164-188, :264-303).                                                       */
JRC_API jrc_status jrc_target_sim(jrc_chain *h, const jrc_c32 *in, int32_t n,
                                  const float *range, const float *velocity, const float *rcs,
                                  const float *azimuth, int32_t n_targets,
                                  const float *position_rx, int32_t n_rx,
                                  int32_t samp_rate, float center_freq,
                                  int32_t self_coupling, float self_coupling_db,
                                  const jrc_c32 *target_phase, int32_t accumulate, jrc_c32 *out);

/* Batched point-target scene on the device: what target_simulator (lib/target_simulator_impl.cc:177,188,296-303) followed
 * by the RX OFDM demodulator hands to the radar block, evaluated in the frequency domain for n_cpi CPIs at once
 *     Y[cpi][r][s][k] = sum_t X[t][s][k] sum_j a_j exp(-j 2 pi tau_{j,t,r} (f_k + fc)),  tau = (2 R_j - d_{t,r} sin az_j)/c
 * with the flowgraph's antenna geometry (...radar_sim.grc:105-147) and optional complex Gaussian noise (sigma per real
 * component, counter-based generator keyed by seed).  tx [n_tx][n_sym][fft_len] and the per-CPI target parameters
 * [n_cpi][n_targets] may be host or device memory; rx_dev [n_cpi][n_rx][n_sym][fft_len] is DEVICE memory, ready for
 * jrc_chain_run_batch.  Generator for simulation sweeps whose inputs do not fit anywhere (BASELINE configs[4]).        */
JRC_API jrc_status jrc_scene_synth(jrc_chain *h, const jrc_c32 *tx, int32_t n_cpi, int32_t n_targets,
                                   const float *range_m, const float *az_deg, const float *amp,
                                   double samp_rate, double center_freq, float noise_sigma, uint64_t seed,
                                   jrc_c32 *rx_dev);

/* blocks_nlog10_ff between complex_to_mag_squared and gui_heatmap_plot (...radar_sim.grc:725-745,
 * 2170-2179; bypassed in the simulation flowgraph, active in the USRP one):
 * out = n*log10(max(in, 1e-18)) + k                                            */
JRC_API jrc_status jrc_nlog10(jrc_chain *h, const float *in, float *out, size_t n_items, float n, float k);

/* range_angle_estimator::work (lib/range_angle_estimator_impl.cc:122-253) on one
 * complex map [n_inputs][vlen]; det is a HOST record.                         */
JRC_API jrc_status jrc_estimate2d(jrc_chain *h, const jrc_c32 *map, int32_t n_inputs, int32_t vlen,
                                  jrc_det *det);

/* fft_peak_detect::work (lib/fft_peak_detect_impl.cc:77-111)                  */
JRC_API jrc_status jrc_peak1d(jrc_chain *h, const jrc_c32 *in, int32_t n, int32_t samp_rate,
                              float interp_factor, float threshold_db, int32_t samp_protect,
                              jrc_peak1d_out *out);

/* ofdm_cyclic_prefix_remover::work (lib/ofdm_cyclic_prefix_remover_impl.cc:86-96): n_sym symbols of
 * fft_len + cp_len time samples -> n_sym vectors of fft_len samples (the step in front of the radar path,
 * SURVEY.md 8(f) rank 1).                                                                          */
JRC_API jrc_status jrc_cp_remove(jrc_chain *h, const jrc_c32 *in, int32_t n_sym, int32_t fft_len,
                                 int32_t cp_len, jrc_c32 *out);

/* The RX OFDM demodulator of the flowgraph in one kernel: cyclic-prefix removal followed by
 * fft_vxx(fft_len, forward, shift=True, no window) (...radar_sim.grc:898-939, 2189-2190): time samples in,
 * DC-centred subcarrier vectors out (what the radar block's rx ports receive).                      */
JRC_API jrc_status jrc_ofdm_demod(jrc_chain *h, const jrc_c32 *in, int32_t n_sym, int32_t fft_len,
                                  int32_t cp_len, jrc_c32 *out);

/* zero_pad::work (lib/zero_pad_impl.cc:67-94): out[pad_front + i] = in[i]; the pads
 * are N(0, 1e-2) complex noise from a counter-based generator keyed by seed
 * (the reference draws a fresh std::random_device seed per call).             */
JRC_API jrc_status jrc_zero_pad(jrc_chain *h, const jrc_c32 *in, int32_t n, uint32_t pad_front,
                                uint32_t pad_tail, uint64_t seed, jrc_c32 *out);

#ifdef __cplusplus
}
#endif
#endif /* JRC_CUDA_H */
