// mimo_ofdm_radar: radar channel estimate per TX/RX pair from the MIMO-LTF symbols, background
// removal, range zero-padding.  Drop-in for the reference block of the same name
// (include/mimo_ofdm_jrc/mimo_ofdm_radar.h:48-63): identical make() arguments, ports and setters;
// the arithmetic runs in libjrc_cuda.so (jrc_radar_estimate).
#pragma once
#include <gnuradio/block.h>
#include <mimo_ofdm_jrc/api.h>
#include <string>

namespace gr {
namespace mimo_ofdm_jrc {

class MIMO_OFDM_JRC_API mimo_ofdm_radar : virtual public gr::block
{
public:
    typedef boost::shared_ptr<mimo_ofdm_radar> sptr;
    static sptr make(int fft_len, int N_tx, int N_rx, int N_sym, int N_pre, bool background_removal,
                     bool background_recording, int record_len, int interp_factor, bool enable_tx_interleave,
                     const std::string &radar_chan_file, const std::string &len_tag_key = "packet_len",
                     bool debug = false);
    virtual void set_background_record(bool background_record) = 0;
    virtual void capture_radar_data(bool capture_sig) = 0;
};

}  // namespace mimo_ofdm_jrc
}  // namespace gr
