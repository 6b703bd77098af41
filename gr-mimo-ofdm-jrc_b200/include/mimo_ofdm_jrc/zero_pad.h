// zero_pad: pads every tagged packet front and tail with N(0, 1e-2) complex noise.
// Drop-in for include/mimo_ofdm_jrc/zero_pad.h:48.
#pragma once
#include <gnuradio/tagged_stream_block.h>
#include <mimo_ofdm_jrc/api.h>

namespace gr {
namespace mimo_ofdm_jrc {

class MIMO_OFDM_JRC_API zero_pad : virtual public gr::tagged_stream_block
{
public:
    typedef boost::shared_ptr<zero_pad> sptr;
    static sptr make(bool debug = false, unsigned int pad_front = 0, unsigned int pad_tail = 0);
};

}  // namespace mimo_ofdm_jrc
}  // namespace gr
