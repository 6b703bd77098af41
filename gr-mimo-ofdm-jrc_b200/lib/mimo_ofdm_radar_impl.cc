// mimo_ofdm_radar on B200: the tag/frame bookkeeping stays on the host, the conj-MAC channel
// estimate, background ring buffer and zero-padded output come from jrc_radar_estimate().
// Replaces lib/mimo_ofdm_radar_impl.cc of the reference (general_work :131-340).
#include "mimo_ofdm_radar_impl.h"

#include <gnuradio/io_signature.h>

#include <fstream>

namespace gr {
namespace mimo_ofdm_jrc {

mimo_ofdm_radar::sptr mimo_ofdm_radar::make(int fft_len, int N_tx, int N_rx, int N_sym, int N_pre,
                                            bool background_removal, bool background_recording, int record_len,
                                            int interp_factor, bool enable_tx_interleave,
                                            const std::string &radar_chan_file, const std::string &len_tag_key,
                                            bool debug)
{
    return gnuradio::get_initial_sptr(new mimo_ofdm_radar_impl(fft_len, N_tx, N_rx, N_sym, N_pre, background_removal,
                                                               background_recording, record_len, interp_factor,
                                                               enable_tx_interleave, radar_chan_file, len_tag_key, debug));
}

mimo_ofdm_radar_impl::mimo_ofdm_radar_impl(int fft_len, int N_tx, int N_rx, int N_sym, int N_pre,
                                           bool background_removal, bool background_recording, int record_len,
                                           int interp_factor, bool enable_tx_interleave,
                                           const std::string &radar_chan_file, const std::string & /*len_tag_key*/,
                                           bool debug)
    : gr::block("mimo_ofdm_radar", gr::io_signature::make(N_tx + N_rx, N_tx + N_rx, sizeof(gr_complex) * fft_len),
                gr::io_signature::make(1, 1, sizeof(gr_complex) * fft_len * interp_factor)),
      d_fft_len(fft_len), d_N_tx(N_tx), d_N_rx(N_rx), d_N_sym(N_sym), d_N_pre(N_pre), d_interp_factor(interp_factor),
      d_radar_chan_file(radar_chan_file), d_debug(debug), d_chan_est((size_t)N_tx * N_rx * fft_len)
{
    jrc_chain_cfg &cfg = d_cfg;
    cfg.fft_len = fft_len; cfg.n_tx = N_tx; cfg.n_rx = N_rx; cfg.n_sym = N_sym; cfg.n_pre = N_pre;
    cfg.interp_range = interp_factor; cfg.interp_angle = 1; cfg.tx_interleave = enable_tx_interleave;
    cfg.background_removal = background_removal; cfg.background_recording = background_recording;
    cfg.record_len = record_len;
    d_chain = std::make_shared<host::chain_handle>(cfg, "MIMO OFDM RADAR");
    if (host::fused_session::requested()) {
        host::fused_session::get().register_radar(cfg);
        d_fused_pending = true;
    }
    set_tag_propagation_policy(TPP_DONT);
    set_output_multiple(N_tx * N_rx);   // one call emits all virtual channels of a frame
}

void mimo_ofdm_radar_impl::set_background_record(bool background_recording)
{
    std::lock_guard<std::mutex> guard(d_lock);
    std::cout << "[MIMO OFDM RADAR] Background recording set to  " << background_recording << std::endl;
    d_cfg.background_recording = background_recording;
    host::check(jrc_chain_set_background_record(d_chain->get(), background_recording), "MIMO OFDM RADAR");
}

void mimo_ofdm_radar_impl::capture_radar_data(bool capture_sig)
{
    if (!capture_sig) return;
    std::lock_guard<std::mutex> guard(d_lock);
    // one line per capture: "HH:MM:SS.mmm, N_tx, N_rx, fft_len:" then the estimate, ';' separated,
    // ";\n" terminated (the Eigen IOFormat of lib/mimo_ofdm_radar_impl.cc:352-369)
    std::ofstream f(d_radar_chan_file, std::ofstream::app);
    if (!f.is_open()) throw std::runtime_error("[MIMO OFDM RADAR] Could not open file!!");
    f << host::time_ms_stamp() << ", " << d_N_tx << ", " << d_N_rx << ", " << d_fft_len << ":";
    f << std::setprecision(7);
    for (size_t i = 0; i < d_chan_est.size(); i++) f << (i ? ";" : "") << d_chan_est[i];
    f << ";\n\n";
    f.flush();
    std::cout << "[MIMO OFDM RADAR] Radar image captured!" << std::endl;
}

// One call needs one whole frame on every port.  The reference keeps gr::block's default (1:1) forecast and reads
// past the items it was given when a frame is split over two calls; asking the scheduler for the frame up front makes
// the WAIT branch below (nothing consumed, nothing produced) the exception instead of a way to stall.
void mimo_ofdm_radar_impl::forecast(int /*noutput_items*/, gr_vector_int &ninput_items_required)
{
    for (auto &n : ninput_items_required) n = d_N_pre + d_N_sym;
}

int mimo_ofdm_radar_impl::general_work(int noutput_items, gr_vector_int &ninput_items,
                                       gr_vector_const_void_star &input_items, gr_vector_void_star &output_items)
{
    std::lock_guard<std::mutex> guard(d_lock);
    const int V = d_N_tx * d_N_rx;
    host::frame_plan plan = host::plan_frame(*this, d_N_tx, ninput_items, d_N_pre + d_N_sym);
    switch (plan.action) {
    case host::frame_plan::NO_RX_TAG:     // nothing tagged in sight: flush every port (:216-231)
        for (size_t i = 0; i < ninput_items.size(); i++) consume((int)i, ninput_items[i]);
        return 0;
    case host::frame_plan::WAIT:
        return 0;
    case host::frame_plan::DROP_RX:
        std::cerr << "[MIMO OFDM RADAR] RX frame without a TX frame: dropped" << std::endl;
        for (int r = 0; r < d_N_rx; r++) consume(d_N_tx + r, (int)plan.rx_packet_len);
        return 0;
    case host::frame_plan::PROCESS:
        break;
    }
    if (noutput_items < V) return 0;

    if (d_fused_pending) {
        // first frame: every block of the flowgraph exists by now.  The whole-chain handle replaces this block's own
        // (no background frame has been recorded yet).
        d_fused_pending = false;
        std::string why;
        if (auto chain = host::fused_session::get().open(d_cfg, "MIMO OFDM RADAR", &why)) {
            d_chain = chain;
            d_fused = true;
            std::cout << "[MIMO OFDM RADAR] JRC_FUSED: one device pass per frame for the three radar blocks" << std::endl;
        } else {
            std::cout << "[MIMO OFDM RADAR] JRC_FUSED ignored: " << why << std::endl;
        }
    }

    std::vector<const jrc_c32 *> tx(d_N_tx), rx(d_N_rx);
    for (int t = 0; t < d_N_tx; t++) tx[t] = static_cast<const jrc_c32 *>(input_items[t]);
    for (int r = 0; r < d_N_rx; r++) rx[r] = static_cast<const jrc_c32 *>(input_items[d_N_tx + r]);
    int64_t seq = -1;
    if (d_fused)
        host::check(jrc_radar_estimate_fused(d_chain->get(), tx.data(), rx.data(), (size_t)plan.tx_skip_items,
                                             static_cast<jrc_c32 *>(output_items[0]),
                                             reinterpret_cast<jrc_c32 *>(d_chan_est.data()), &seq),
                    "MIMO OFDM RADAR");
    else
        host::check(jrc_radar_estimate(d_chain->get(), tx.data(), rx.data(), (size_t)plan.tx_skip_items,
                                       static_cast<jrc_c32 *>(output_items[0]),
                                       reinterpret_cast<jrc_c32 *>(d_chan_est.data())),
                    "MIMO OFDM RADAR");

    add_item_tag(0, nitems_written(0), pmt::string_to_symbol("packet_len"), pmt::from_long(V),
                 pmt::string_to_symbol(alias()));
    if (d_fused)
        add_item_tag(0, nitems_written(0), host::fused_session::tag_key(), pmt::from_long((long)seq),
                     pmt::string_to_symbol(alias()));
    for (int r = 0; r < d_N_rx; r++) consume(d_N_tx + r, (int)plan.rx_packet_len);
    for (int t = 0; t < d_N_tx; t++) consume(t, (int)(plan.tx_skip_items + plan.tx_packet_len));
    if (d_debug) std::cout << "[MIMO OFDM RADAR] frame done, tx skip " << plan.tx_skip_items << std::endl;
    return V;
}

}  // namespace mimo_ofdm_jrc
}  // namespace gr
