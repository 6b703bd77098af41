// range_angle_estimator: 2-D peak search on the complex range-angle map, noise window, SNR gate,
// "params" message and CSV log.  Drop-in for include/mimo_ofdm_jrc/range_angle_estimator.h:48-62.
#pragma once
#include <gnuradio/tagged_stream_block.h>
#include <mimo_ofdm_jrc/api.h>
#include <string>
#include <vector>

namespace gr {
namespace mimo_ofdm_jrc {

class MIMO_OFDM_JRC_API range_angle_estimator : virtual public gr::tagged_stream_block
{
public:
    typedef boost::shared_ptr<range_angle_estimator> sptr;
    static sptr make(int vlen, std::vector<float> range_bins, std::vector<float> angle_bins,
                     float noise_discard_range_m, float noise_discard_angle_deg, float snr_threshold,
                     float power_threshold, const std::string &stats_path, bool stats_record,
                     const std::string &len_key = "packet_len", bool debug = false);
    virtual void set_snr_threshold(float snr_threshold) = 0;
    virtual void set_power_threshold(float power_threshold) = 0;
    virtual void set_stats_record(bool stats_record) = 0;
};

}  // namespace mimo_ofdm_jrc
}  // namespace gr
