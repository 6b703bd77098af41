// radar_chain: the fused B200 chain as one GNU Radio block (see include/mimo_ofdm_jrc/radar_chain.h).
// general_work = mimo_ofdm_radar's frame bookkeeping (lib/mimo_ofdm_radar_impl.cc:131-340 of the reference) + the
// C-ABI chain.  With a pipeline depth above 1 (set_pipeline_depth / JRC_PIPELINE) a call SUBMITS its frame
// (jrc_chain_submit) and publishes the oldest finished one, so the transfers and kernel of CPI k+1 overlap the output
// transfer of CPI k: the |.|^2 packet of a frame leaves one call later, the stream contents are identical.
#include <mimo_ofdm_jrc/radar_chain.h>

#include <gnuradio/io_signature.h>

#include <cstring>
#include <deque>

#include "jrc_host.h"

namespace gr {
namespace mimo_ofdm_jrc {

class radar_chain_impl : public radar_chain
{
    enum { MAX_DEPTH = 4 };
    const int d_fft_len, d_N_tx, d_N_rx, d_N_sym, d_N_pre, d_Nr, d_Na;
    const std::vector<float> d_range_bins, d_angle_bins;
    float d_snr_threshold, d_power_threshold;
    host::stats_log d_log;
    const bool d_debug;
    pmt::pmt_t d_srcid;                 // alias() as a symbol, made on first use (the alias is final once the block is in a graph)
    host::chain_handle d_chain;
    std::mutex d_lock;                      // the GRC callbacks run on another thread than general_work()
    int d_depth = 1;
    // per in-flight frame: packed symbols without the preamble [ant][sym][fft_len] and the record, in pinned memory
    struct slot_t { gr_complex *rx = nullptr, *tx = nullptr; jrc_det *det = nullptr; };
    slot_t d_slot[MAX_DEPTH];
    void *d_pinned = nullptr;
    struct inflight_t { int64_t ticket; int slot; };
    std::deque<inflight_t> d_q;
    int d_next_slot = 0;

    void push_thresholds() { host::check(jrc_chain_set_thresholds(d_chain.get(), d_snr_threshold, d_power_threshold), "RADAR CHAIN"); }

    void publish(const jrc_det &det)
    {
        if (!(det.flags & JRC_DET_PASSED)) return;
        const float range_val = d_range_bins[det.range_idx], angle_val = d_angle_bins[det.angle_idx];
        static const pmt::pmt_t port = pmt::mp("params");
        message_port_pub(port, host::params_message(range_val, angle_val, det.peak_power, det.snr_db));
        if (d_log.record && !d_log.append(det.peak_power, det.snr_db, range_val, angle_val))
            throw std::runtime_error("[RADAR CHAIN] Could not open file!!");
    }

    // retires the oldest in-flight frame: its packet is the next Nr items of the output stream
    int emit_front()
    {
        const inflight_t f = d_q.front();
        host::check(jrc_chain_wait(d_chain.get(), f.ticket), "RADAR CHAIN");
        d_q.pop_front();
        static const pmt::pmt_t len_key = pmt::string_to_symbol("packet_len");
        if (!d_srcid) d_srcid = pmt::string_to_symbol(alias());
        add_item_tag(0, nitems_written(0), len_key, pmt::from_long(d_Nr), d_srcid);
        publish(*d_slot[f.slot].det);
        return d_Nr;
    }

public:
    radar_chain_impl(int fft_len, int N_tx, int N_rx, int N_sym, int N_pre, bool background_removal, bool background_recording,
                     int record_len, int ir, int ia, bool enable_tx_interleave, const std::vector<float> &range_bins,
                     const std::vector<float> &angle_bins, float nd_range_m, float nd_angle_deg, float snr_threshold,
                     float power_threshold, const std::string &stats_path, bool stats_record, bool debug)
        : gr::block("radar_chain", gr::io_signature::make(N_tx + N_rx, N_tx + N_rx, sizeof(gr_complex) * fft_len),
                    gr::io_signature::make(1, 1, sizeof(float) * N_tx * N_rx * ia)),
          d_fft_len(fft_len), d_N_tx(N_tx), d_N_rx(N_rx), d_N_sym(N_sym), d_N_pre(N_pre), d_Nr(fft_len * ir),
          d_Na(N_tx * N_rx * ia), d_range_bins(range_bins), d_angle_bins(angle_bins), d_snr_threshold(snr_threshold),
          d_power_threshold(power_threshold), d_debug(debug)
    {
        jrc_chain_cfg cfg{};
        cfg.fft_len = fft_len; cfg.n_tx = N_tx; cfg.n_rx = N_rx; cfg.n_sym = N_sym; cfg.n_pre = 0;
        cfg.interp_range = ir; cfg.interp_angle = ia; cfg.tx_interleave = enable_tx_interleave;
        cfg.background_removal = background_removal; cfg.background_recording = background_recording;
        cfg.record_len = record_len;
        d_chain.open(cfg, "RADAR CHAIN");
        host::check(jrc_chain_set_estimator(d_chain.get(), d_range_bins.data(), (int)d_range_bins.size(), d_angle_bins.data(),
                                            (int)d_angle_bins.size(), nd_range_m, nd_angle_deg, snr_threshold, power_threshold),
                    "RADAR CHAIN");
        // frame staging buffers in pinned memory: the kernel reads them in place
        const size_t rx_b = sizeof(gr_complex) * N_rx * N_sym * fft_len, tx_b = sizeof(gr_complex) * N_tx * N_sym * fft_len;
        const size_t per = ((rx_b + tx_b + sizeof(jrc_det) + 255) / 256) * 256;
        host::check(jrc_pinned_alloc(per * MAX_DEPTH, &d_pinned), "RADAR CHAIN");
        for (int i = 0; i < MAX_DEPTH; i++) {
            char *b = static_cast<char *>(d_pinned) + per * i;
            d_slot[i].rx = reinterpret_cast<gr_complex *>(b);
            d_slot[i].tx = reinterpret_cast<gr_complex *>(b + rx_b);
            d_slot[i].det = reinterpret_cast<jrc_det *>(b + rx_b + tx_b);
        }
        if (const char *e = std::getenv("JRC_PIPELINE")) set_pipeline_depth(std::atoi(e));
        d_log.path = stats_path; d_log.record = stats_record;
        message_port_register_out(pmt::mp("params"));
        set_tag_propagation_policy(TPP_DONT);
        set_output_multiple(d_Nr);
    }
    ~radar_chain_impl() override
    {
        while (!d_q.empty()) { jrc_chain_wait(d_chain.get(), d_q.front().ticket); d_q.pop_front(); }
        jrc_pinned_free(d_pinned);
    }

    void set_background_record(bool on) override
    {
        std::lock_guard<std::mutex> guard(d_lock);
        host::check(jrc_chain_set_background_record(d_chain.get(), on), "RADAR CHAIN");
    }
    void set_snr_threshold(float v) override { std::lock_guard<std::mutex> guard(d_lock); d_snr_threshold = v; push_thresholds(); }
    void set_power_threshold(float v) override { std::lock_guard<std::mutex> guard(d_lock); d_power_threshold = v; push_thresholds(); }
    void set_stats_record(bool on) override { std::lock_guard<std::mutex> guard(d_lock); d_log.record = on; d_log.header_written = false; }
    void set_pipeline_depth(int depth) override
    {
        std::lock_guard<std::mutex> guard(d_lock);
        d_depth = depth < 1 ? 1 : (depth > MAX_DEPTH ? MAX_DEPTH : depth);
    }

    // one call needs one whole frame on every port (see mimo_ofdm_radar_impl::forecast)
    void forecast(int /*noutput_items*/, gr_vector_int &ninput_items_required) override
    {
        for (auto &n : ninput_items_required) n = d_q.empty() ? d_N_pre + d_N_sym : 0;   // frames in flight can be retired without input
    }

    int general_work(int noutput_items, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items,
                     gr_vector_void_star &output_items) override
    {
        std::lock_guard<std::mutex> guard(d_lock);
        host::frame_plan plan = host::plan_frame(*this, d_N_tx, ninput_items, d_N_pre + d_N_sym);
        if (plan.action == host::frame_plan::NO_RX_TAG) {
            for (size_t i = 0; i < ninput_items.size(); i++) consume((int)i, ninput_items[i]);
        } else if (plan.action == host::frame_plan::DROP_RX) {
            for (int r = 0; r < d_N_rx; r++) consume(d_N_tx + r, (int)plan.rx_packet_len);
        }
        const int queued = (int)d_q.size();
        const bool can_submit = plan.action == host::frame_plan::PROCESS && queued < d_depth &&
                                noutput_items >= (queued + 1) * d_Nr;
        if (can_submit) {
            const int s = d_next_slot;
            d_next_slot = (d_next_slot + 1) % MAX_DEPTH;
            const size_t per_ant = (size_t)d_N_sym * d_fft_len;
            for (int t = 0; t < d_N_tx; t++)
                std::memcpy(d_slot[s].tx + t * per_ant,
                            static_cast<const gr_complex *>(input_items[t]) + (plan.tx_skip_items + d_N_pre) * d_fft_len,
                            per_ant * sizeof(gr_complex));
            for (int r = 0; r < d_N_rx; r++)
                std::memcpy(d_slot[s].rx + r * per_ant,
                            static_cast<const gr_complex *>(input_items[d_N_tx + r]) + (size_t)d_N_pre * d_fft_len,
                            per_ant * sizeof(gr_complex));
            // the packet of this frame follows those of the frames in flight in the output buffer
            float *dst = static_cast<float *>(output_items[0]) + (size_t)queued * d_Nr * d_Na;
            int64_t ticket = 0;
            host::check(jrc_chain_submit(d_chain.get(), reinterpret_cast<const jrc_c32 *>(d_slot[s].rx),
                                         reinterpret_cast<const jrc_c32 *>(d_slot[s].tx), 1, 1, 0, dst, d_slot[s].det, &ticket),
                        "RADAR CHAIN");
            d_q.push_back({ticket, s});
            for (int r = 0; r < d_N_rx; r++) consume(d_N_tx + r, (int)plan.rx_packet_len);
            for (int t = 0; t < d_N_tx; t++) consume(t, (int)(plan.tx_skip_items + plan.tx_packet_len));
        }
        if (d_q.empty()) return 0;
        // retire the oldest frame when it is finished, or when this call could not make progress any other way
        bool retire = !can_submit || (int)d_q.size() >= d_depth;
        if (!retire) {
            int32_t done = 0;
            host::check(jrc_chain_poll(d_chain.get(), d_q.front().ticket, &done), "RADAR CHAIN");
            retire = done != 0;
        }
        if (!retire || noutput_items < d_Nr) return 0;
        return emit_front();
    }
};

radar_chain::sptr radar_chain::make(int fft_len, int N_tx, int N_rx, int N_sym, int N_pre, bool background_removal,
                                    bool background_recording, int record_len, int interp_factor_range,
                                    int interp_factor_angle, bool enable_tx_interleave, std::vector<float> range_bins,
                                    std::vector<float> angle_bins, float noise_discard_range_m, float noise_discard_angle_deg,
                                    float snr_threshold, float power_threshold, const std::string &stats_path,
                                    bool stats_record, const std::string & /*len_tag_key*/, bool debug)
{
    return gnuradio::get_initial_sptr(new radar_chain_impl(fft_len, N_tx, N_rx, N_sym, N_pre, background_removal,
                                                           background_recording, record_len, interp_factor_range,
                                                           interp_factor_angle, enable_tx_interleave, range_bins, angle_bins,
                                                           noise_discard_range_m, noise_discard_angle_deg, snr_threshold,
                                                           power_threshold, stats_path, stats_record, debug));
}

}  // namespace mimo_ofdm_jrc
}  // namespace gr
