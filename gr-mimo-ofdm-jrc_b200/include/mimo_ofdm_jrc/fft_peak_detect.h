// fft_peak_detect: 1-D peak of a spectrum packet with guard samples and a dB threshold, outputs
// (freq, phase, mag).  Drop-in for include/mimo_ofdm_jrc/fft_peak_detect.h:48-52.
#pragma once
#include <gnuradio/tagged_stream_block.h>
#include <mimo_ofdm_jrc/api.h>
#include <string>
#include <vector>

namespace gr {
namespace mimo_ofdm_jrc {

class MIMO_OFDM_JRC_API fft_peak_detect : virtual public gr::tagged_stream_block
{
public:
    typedef boost::shared_ptr<fft_peak_detect> sptr;
    static sptr make(int samp_rate, float interp_factor, float threshold, int samp_protect,
                     std::vector<float> max_freq, bool cut_max_freq, const std::string &len_key);
    virtual void set_threshold(float threshold) = 0;
    virtual void set_samp_protect(int samp) = 0;
    virtual void set_max_freq(std::vector<float> freq) = 0;
};

}  // namespace mimo_ofdm_jrc
}  // namespace gr
