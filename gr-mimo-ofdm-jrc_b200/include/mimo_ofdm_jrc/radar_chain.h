// radar_chain: the whole radar receive chain as ONE block -- what a flowgraph uses to get the fused
// B200 kernel instead of the five-block CPU chain
//   mimo_ofdm_radar -> fft_vxx(IFFT) -> matrix_transpose -> fft_vxx(FFT, shift)
//                   -> {complex_to_mag_squared, range_angle_estimator}
// (examples/simulation/radar/mimo_ofdm_jrc_radar_sim.grc:2165-2232 of the reference).
// Inputs: N_tx + N_rx streams of complex[fft_len] exactly like mimo_ofdm_radar.  Output 0: the |.|^2
// map as float[Na] vectors, one tagged packet of Nr items per CPI (what gui_heatmap_plot consumes).
// Message port "params": the estimator's message.  Not in the reference: its parameters are the
// union of the make() arguments of the blocks it replaces.
#pragma once
#include <gnuradio/block.h>
#include <mimo_ofdm_jrc/api.h>
#include <string>
#include <vector>

namespace gr {
namespace mimo_ofdm_jrc {

class MIMO_OFDM_JRC_API radar_chain : virtual public gr::block
{
public:
    typedef boost::shared_ptr<radar_chain> sptr;
    static sptr make(int fft_len, int N_tx, int N_rx, int N_sym, int N_pre, bool background_removal,
                     bool background_recording, int record_len, int interp_factor_range,
                     int interp_factor_angle, bool enable_tx_interleave, std::vector<float> range_bins,
                     std::vector<float> angle_bins, float noise_discard_range_m, float noise_discard_angle_deg,
                     float snr_threshold, float power_threshold, const std::string &stats_path,
                     bool stats_record, const std::string &len_tag_key = "packet_len", bool debug = false);
    virtual void set_background_record(bool background_record) = 0;
    virtual void set_snr_threshold(float snr_threshold) = 0;
    virtual void set_power_threshold(float power_threshold) = 0;
    virtual void set_stats_record(bool stats_record) = 0;
    // frames in flight (1: every call is synchronous; up to 4: a call submits its frame and publishes the oldest
    // finished one -- jrc_chain_submit / jrc_chain_wait).  Also settable with JRC_PIPELINE=<depth>.
    virtual void set_pipeline_depth(int depth) = 0;
};

}  // namespace mimo_ofdm_jrc
}  // namespace gr
