// matrix_transpose: [K][input_len] packet -> [input_len][output_len*interp_factor], zero filled
// (angle zero-pad).  Drop-in for include/mimo_ofdm_jrc/matrix_transpose.h:48 of the reference.
#pragma once
#include <gnuradio/tagged_stream_block.h>
#include <mimo_ofdm_jrc/api.h>
#include <string>

namespace gr {
namespace mimo_ofdm_jrc {

class MIMO_OFDM_JRC_API matrix_transpose : virtual public gr::tagged_stream_block
{
public:
    typedef boost::shared_ptr<matrix_transpose> sptr;
    static sptr make(int input_len, int output_len, int interp_factor, bool debug, std::string len_key = "packet_len");
};

}  // namespace mimo_ofdm_jrc
}  // namespace gr
