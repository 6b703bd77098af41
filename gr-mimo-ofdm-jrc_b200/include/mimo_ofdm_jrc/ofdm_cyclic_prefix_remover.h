// ofdm_cyclic_prefix_remover: time samples of a tagged packet -> fft_len-vectors without the cyclic prefix.
// Drop-in for include/mimo_ofdm_jrc/ofdm_cyclic_prefix_remover.h:48 (the block in front of the radar path).
#pragma once
#include <gnuradio/tagged_stream_block.h>
#include <mimo_ofdm_jrc/api.h>
#include <string>

namespace gr {
namespace mimo_ofdm_jrc {

class MIMO_OFDM_JRC_API ofdm_cyclic_prefix_remover : virtual public gr::tagged_stream_block
{
public:
    typedef boost::shared_ptr<ofdm_cyclic_prefix_remover> sptr;
    static sptr make(int fft_len, int cp_len, std::string len_key = "packet_len");
};

}  // namespace mimo_ofdm_jrc
}  // namespace gr
