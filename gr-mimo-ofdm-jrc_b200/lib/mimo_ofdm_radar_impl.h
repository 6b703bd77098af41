#pragma once
#include <mimo_ofdm_jrc/mimo_ofdm_radar.h>
#include "jrc_host.h"

namespace gr {
namespace mimo_ofdm_jrc {

class mimo_ofdm_radar_impl : public mimo_ofdm_radar
{
    const int d_fft_len, d_N_tx, d_N_rx, d_N_sym, d_N_pre, d_interp_factor;
    const std::string d_radar_chan_file;
    const bool d_debug;
    std::shared_ptr<host::chain_handle> d_chain;
    jrc_chain_cfg d_cfg{};
    bool d_fused_pending = false, d_fused = false;   // JRC_FUSED=1: resolved on the first frame (jrc_host.h, fused_session)
    std::vector<gr_complex> d_chan_est;   // last radar_chan_est, host copy for capture_radar_data
    std::mutex d_lock;                    // the GRC callbacks run on another thread than general_work()

public:
    mimo_ofdm_radar_impl(int fft_len, int N_tx, int N_rx, int N_sym, int N_pre, bool background_removal,
                         bool background_recording, int record_len, int interp_factor, bool enable_tx_interleave,
                         const std::string &radar_chan_file, const std::string &len_tag_key, bool debug);
    void set_background_record(bool background_recording) override;
    void capture_radar_data(bool capture_sig) override;
    void forecast(int noutput_items, gr_vector_int &ninput_items_required) override;
    int general_work(int noutput_items, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items,
                     gr_vector_void_star &output_items) override;
};

}  // namespace mimo_ofdm_jrc
}  // namespace gr
