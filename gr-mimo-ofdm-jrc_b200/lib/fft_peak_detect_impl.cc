// fft_peak_detect on B200 (jrc_peak1d).  Replaces lib/fft_peak_detect_impl.cc:67-111.
#include <mimo_ofdm_jrc/fft_peak_detect.h>

#include <gnuradio/io_signature.h>

#include "jrc_host.h"

namespace gr {
namespace mimo_ofdm_jrc {

class fft_peak_detect_impl : public fft_peak_detect
{
    int d_samp_rate;
    float d_interp_factor, d_threshold;
    int d_samp_protect;
    std::vector<float> d_max_freq;   // kept for the setter; unused, as in the reference
    bool d_cut_max_freq;
    host::chain_handle d_chain;

protected:
    int calculate_output_stream_length(const gr_vector_int &) override { return 1; }

public:
    fft_peak_detect_impl(int samp_rate, float interp_factor, float threshold, int samp_protect,
                         const std::vector<float> &max_freq, bool cut_max_freq, const std::string &len_key)
        : gr::tagged_stream_block("fft_peak_detect", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                                  gr::io_signature::make3(3, 3, sizeof(float), sizeof(float), sizeof(float)), len_key),
          d_samp_rate(samp_rate), d_interp_factor(interp_factor), d_threshold(threshold), d_samp_protect(samp_protect),
          d_max_freq(max_freq), d_cut_max_freq(cut_max_freq), d_chain(host::utility_cfg(), "FFT PEAK DETECT")
    {
        set_tag_propagation_policy(TPP_DONT);
    }

    void set_threshold(float threshold) override { d_threshold = threshold; }
    void set_samp_protect(int samp) override { d_samp_protect = samp; }
    void set_max_freq(std::vector<float> freq) override { d_max_freq = freq; }

    int work(int, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) override
    {
        jrc_peak1d_out pk;
        host::check(jrc_peak1d(d_chain.get(), static_cast<const jrc_c32 *>(input_items[0]), ninput_items[0], d_samp_rate,
                               d_interp_factor, d_threshold, d_samp_protect, &pk),
                    "FFT PEAK DETECT");
        if (pk.k != -1) {   // no peak: the output item is left as it is, like the reference (:98-110)
            static_cast<float *>(output_items[0])[0] = pk.freq;
            static_cast<float *>(output_items[1])[0] = pk.phase;
            static_cast<float *>(output_items[2])[0] = pk.mag;
        }
        return 1;
    }
};

fft_peak_detect::sptr fft_peak_detect::make(int samp_rate, float interp_factor, float threshold, int samp_protect,
                                            std::vector<float> max_freq, bool cut_max_freq, const std::string &len_key)
{
    return gnuradio::get_initial_sptr(new fft_peak_detect_impl(samp_rate, interp_factor, threshold, samp_protect, max_freq,
                                                               cut_max_freq, len_key));
}

}  // namespace mimo_ofdm_jrc
}  // namespace gr
