// ofdm_cyclic_prefix_remover on B200 (jrc_cp_remove).  Replaces lib/ofdm_cyclic_prefix_remover_impl.cc:62-99.
#include <mimo_ofdm_jrc/ofdm_cyclic_prefix_remover.h>

#include <gnuradio/io_signature.h>

#include "jrc_host.h"

namespace gr {
namespace mimo_ofdm_jrc {

class ofdm_cyclic_prefix_remover_impl : public ofdm_cyclic_prefix_remover
{
    const int d_fft_len, d_cp_len;
    host::chain_handle d_chain;
    std::vector<tag_t> d_tags;

protected:
    int calculate_output_stream_length(const gr_vector_int &ninput_items) override
    {
        return ninput_items[0] / (d_fft_len + d_cp_len);
    }

public:
    ofdm_cyclic_prefix_remover_impl(int fft_len, int cp_len, const std::string &len_key)
        : gr::tagged_stream_block("ofdm_cyclic_prefix_remover", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                                  gr::io_signature::make(1, 1, sizeof(gr_complex) * fft_len), len_key),
          d_fft_len(fft_len), d_cp_len(cp_len), d_chain(host::utility_cfg(), "CP REMOVER")
    {
        set_tag_propagation_policy(TPP_DONT);
    }

    int work(int, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) override
    {
        // the tags sitting on the first sample of the packet move to the first output vector (:79-83)
        get_tags_in_range(d_tags, 0, nitems_read(0), nitems_read(0) + 1);
        for (const tag_t &t : d_tags) add_item_tag(0, nitems_written(0), t.key, t.value, t.srcid);
        const int n_sym = ninput_items[0] / (d_fft_len + d_cp_len);
        host::check(jrc_cp_remove(d_chain.get(), static_cast<const jrc_c32 *>(input_items[0]), n_sym, d_fft_len, d_cp_len,
                                  static_cast<jrc_c32 *>(output_items[0])),
                    "CP REMOVER");
        return n_sym;
    }
};

ofdm_cyclic_prefix_remover::sptr ofdm_cyclic_prefix_remover::make(int fft_len, int cp_len, std::string len_key)
{
    return gnuradio::get_initial_sptr(new ofdm_cyclic_prefix_remover_impl(fft_len, cp_len, len_key));
}

}  // namespace mimo_ofdm_jrc
}  // namespace gr
