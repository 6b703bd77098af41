// zero_pad on B200 (jrc_zero_pad).  Replaces lib/zero_pad_impl.cc:61-94.
#include <mimo_ofdm_jrc/zero_pad.h>

#include <gnuradio/io_signature.h>

#include <random>

#include "jrc_host.h"

namespace gr {
namespace mimo_ofdm_jrc {

class zero_pad_impl : public zero_pad
{
    const bool d_debug;
    const unsigned d_pad_front, d_pad_tail;
    host::chain_handle d_chain;
    std::random_device d_entropy;   // a fresh seed per packet, as the reference draws one per work()

protected:
    int calculate_output_stream_length(const gr_vector_int &ninput_items) override
    {
        return ninput_items[0] + (int)d_pad_front + (int)d_pad_tail;
    }

public:
    zero_pad_impl(bool debug, unsigned pad_front, unsigned pad_tail)
        : gr::tagged_stream_block("zero_pad", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                                  gr::io_signature::make(1, 1, sizeof(gr_complex)), "packet_len"),
          d_debug(debug), d_pad_front(pad_front), d_pad_tail(pad_tail), d_chain(host::utility_cfg(), "ZERO PAD")
    {
        set_tag_propagation_policy(TPP_DONT);
    }

    int work(int, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) override
    {
        const uint64_t seed = ((uint64_t)d_entropy() << 32) ^ (uint64_t)d_entropy();
        host::check(jrc_zero_pad(d_chain.get(), static_cast<const jrc_c32 *>(input_items[0]), ninput_items[0], d_pad_front,
                                 d_pad_tail, seed, static_cast<jrc_c32 *>(output_items[0])),
                    "ZERO PAD");
        return ninput_items[0] + (int)d_pad_front + (int)d_pad_tail;
    }
};

zero_pad::sptr zero_pad::make(bool debug, unsigned int pad_front, unsigned int pad_tail)
{
    return gnuradio::get_initial_sptr(new zero_pad_impl(debug, pad_front, pad_tail));
}

}  // namespace mimo_ofdm_jrc
}  // namespace gr
