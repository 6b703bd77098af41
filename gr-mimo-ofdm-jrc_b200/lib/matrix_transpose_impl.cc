// matrix_transpose on B200 (jrc_transpose_pad).  Replaces lib/matrix_transpose_impl.cc:62-110.
#include <mimo_ofdm_jrc/matrix_transpose.h>

#include <gnuradio/io_signature.h>

#include "jrc_host.h"

namespace gr {
namespace mimo_ofdm_jrc {

class matrix_transpose_impl : public matrix_transpose
{
    const int d_input_len, d_output_len, d_interp_factor;
    const bool d_debug;
    const bool d_fused;                 // JRC_FUSED=1 (jrc_host.h, fused_session)
    host::chain_handle d_chain;

protected:
    int calculate_output_stream_length(const gr_vector_int &) override { return d_input_len; }

public:
    matrix_transpose_impl(int input_len, int output_len, int interp_factor, bool debug, const std::string &len_key)
        : gr::tagged_stream_block("matrix_transpose", gr::io_signature::make(1, 1, sizeof(gr_complex) * input_len),
                                  gr::io_signature::make(1, 1, sizeof(gr_complex) * output_len * interp_factor), len_key),
          d_input_len(input_len), d_output_len(output_len), d_interp_factor(interp_factor), d_debug(debug),
          d_fused(host::fused_session::requested()), d_chain(host::utility_cfg(), "MATRIX TRANSPOSE")
    {
        if (d_fused) host::fused_session::get().register_transpose(input_len, output_len, interp_factor);
        set_relative_rate((double)input_len / (double)output_len);
        set_tag_propagation_policy(TPP_DONT);
    }

    int work(int /*noutput_items*/, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items,
             gr_vector_void_star &output_items) override
    {
        const int k_items = ninput_items[0];
        // the packet must fill whole columns (:82-83)
        if (k_items * float(d_input_len) / float(d_output_len) - k_items * d_input_len / d_output_len != 0)
            throw std::runtime_error("[MATRIX TRANSPOSE] input_len and output_len do not match to packet length");
        // back-pressure: drop this CPI instead of queueing it (:86-89)
        if (pc_output_buffers_full(0) > 0.001) return 0;
        bool served = false;
        if (d_fused) {
            // the radar block already ran this frame through the chain: its transposed array is waiting under the
            // frame's sequence number, and the number travels on with the packet for the estimator
            const int64_t seq = host::fused_session::packet_seq(*this, k_items);
            auto chain = seq >= 0 ? host::fused_session::get().chain() : nullptr;
            if (chain && k_items == d_output_len &&
                jrc_fused_fetch_transposed(chain->get(), seq, static_cast<jrc_c32 *>(output_items[0])) == JRC_OK) {
                add_item_tag(0, nitems_written(0), host::fused_session::tag_key(), pmt::from_long((long)seq),
                             pmt::string_to_symbol(alias()));
                served = true;
            }
        }
        if (!served)
            host::check(jrc_transpose_pad(d_chain.get(), static_cast<const jrc_c32 *>(input_items[0]), k_items, d_input_len,
                                          d_output_len, d_interp_factor, static_cast<jrc_c32 *>(output_items[0])),
                        "MATRIX TRANSPOSE");
        if (d_debug) std::cout << "[MATRIX TRANSPOSE] " << k_items << " x " << d_input_len << " transposed" << std::endl;
        return d_input_len;   // the base class tags the packet with this length
    }
};

matrix_transpose::sptr matrix_transpose::make(int input_len, int output_len, int interp_factor, bool debug, std::string len_key)
{
    return gnuradio::get_initial_sptr(new matrix_transpose_impl(input_len, output_len, interp_factor, debug, len_key));
}

}  // namespace mimo_ofdm_jrc
}  // namespace gr
