// target_simulator: static-target radar channel for the simulation flowgraphs, one input packet of time
// samples -> one packet per RX antenna.  Drop-in for include/mimo_ofdm_jrc/target_simulator.h:48-71.
#pragma once
#include <gnuradio/tagged_stream_block.h>
#include <mimo_ofdm_jrc/api.h>
#include <string>
#include <vector>

namespace gr {
namespace mimo_ofdm_jrc {

class MIMO_OFDM_JRC_API target_simulator : virtual public gr::tagged_stream_block
{
public:
    typedef boost::shared_ptr<target_simulator> sptr;
    static sptr make(std::vector<float> range, std::vector<float> velocity, std::vector<float> rcs,
                     std::vector<float> azimuth, std::vector<float> position_rx, int samp_rate, float center_freq,
                     float self_coupling_db, bool rndm_phaseshift = false, bool self_coupling = false,
                     const std::string &len_key = "packet_len", bool debug = false);

    virtual void setup_targets(std::vector<float> range, std::vector<float> velocity, std::vector<float> rcs,
                               std::vector<float> azimuth, std::vector<float> position_rx, int samp_rate,
                               float center_freq, float self_coupling_db, bool rndm_phaseshift, bool self_coupling) = 0;
};

}  // namespace mimo_ofdm_jrc
}  // namespace gr
