// target_simulator on B200 (jrc_target_sim).  Replaces lib/target_simulator_impl.cc:127-385.
#include <mimo_ofdm_jrc/target_simulator.h>

#include <gnuradio/io_signature.h>

#include <cmath>
#include <cstdlib>
#include <ctime>
#include <mutex>

#include "jrc_host.h"

namespace gr {
namespace mimo_ofdm_jrc {

class target_simulator_impl : public target_simulator
{
    std::vector<float> d_range, d_velocity, d_rcs, d_azimuth, d_position_rx;
    int d_samp_rate = 1;
    float d_center_freq = 0.f, d_self_coupling_db = 0.f;
    bool d_rndm_phaseshift = false, d_self_coupling = false;
    host::chain_handle d_chain;
    std::mutex d_lock;                       // setup_targets() arrives from the GUI thread (:138, :207)
    std::vector<gr_complex> d_out, d_phase;
    const pmt::pmt_t d_key = pmt::string_to_symbol("rx_time"), d_srcid = pmt::string_to_symbol("stat_targ_sim");

protected:
    int calculate_output_stream_length(const gr_vector_int &ninput_items) override { return ninput_items[0]; }

public:
    target_simulator_impl(std::vector<float> range, std::vector<float> velocity, std::vector<float> rcs,
                          std::vector<float> azimuth, std::vector<float> position_rx, int samp_rate, float center_freq,
                          float self_coupling_db, bool rndm_phaseshift, bool self_coupling, const std::string &len_key)
        : gr::tagged_stream_block("target_simulator", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                                  gr::io_signature::make((int)position_rx.size(), (int)position_rx.size(), sizeof(gr_complex)),
                                  len_key),
          d_chain(host::utility_cfg(), "TARGET SIM")
    {
        setup_targets(range, velocity, rcs, azimuth, position_rx, samp_rate, center_freq, self_coupling_db,
                      rndm_phaseshift, self_coupling);
    }

    void setup_targets(std::vector<float> range, std::vector<float> velocity, std::vector<float> rcs,
                       std::vector<float> azimuth, std::vector<float> position_rx, int samp_rate, float center_freq,
                       float self_coupling_db, bool rndm_phaseshift, bool self_coupling) override
    {
        std::lock_guard<std::mutex> g(d_lock);
        if (velocity.size() != range.size() || rcs.size() != range.size() || azimuth.size() != range.size())
            throw std::invalid_argument("[TARGET SIM] range, velocity, rcs and azimuth must have the same length");
        d_range = range; d_velocity = velocity; d_rcs = rcs; d_azimuth = azimuth; d_position_rx = position_rx;
        d_samp_rate = samp_rate; d_center_freq = center_freq; d_self_coupling_db = self_coupling_db;
        d_rndm_phaseshift = rndm_phaseshift; d_self_coupling = self_coupling;
        if (d_rndm_phaseshift) std::srand((unsigned)std::time(NULL));    // :196
    }

    int work(int, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) override
    {
        std::lock_guard<std::mutex> g(d_lock);
        const int n = ninput_items[0], L = (int)d_position_rx.size(), K = (int)d_range.size();
        const jrc_c32 *phase = nullptr;
        if (d_rndm_phaseshift) {                                          // :311-320
            d_phase.resize((size_t)K);
            for (int k = 0; k < K; k++)
                d_phase[(size_t)k] = std::exp(gr_complex(0, 2 * (float)M_PI * float((std::rand() % 1000 + 1) / 1000.0)));
            phase = reinterpret_cast<const jrc_c32 *>(d_phase.data());
        }
        d_out.resize((size_t)L * n);
        host::check(jrc_target_sim(d_chain.get(), static_cast<const jrc_c32 *>(input_items[0]), n, d_range.data(),
                                   d_velocity.data(), d_rcs.data(), d_azimuth.data(), K, d_position_rx.data(), L, d_samp_rate,
                                   d_center_freq, d_self_coupling, d_self_coupling_db, phase, /*accumulate*/ 0,
                                   reinterpret_cast<jrc_c32 *>(d_out.data())),
                    "TARGET SIM");
        for (int l = 0; l < L; l++) {
            // rx_time tag at the packet start (:330-336)
            const uint64_t sec = nitems_written(l) / d_samp_rate;
            const double frac = nitems_written(l) / (float)d_samp_rate - sec;
            add_item_tag(l, nitems_written(l), d_key, pmt::make_tuple(pmt::from_uint64(sec), pmt::from_double(frac)), d_srcid);
            std::memcpy(output_items[l], d_out.data() + (size_t)l * n, sizeof(gr_complex) * (size_t)n);
        }
        return n;
    }
};

target_simulator::sptr target_simulator::make(std::vector<float> range, std::vector<float> velocity, std::vector<float> rcs,
                                              std::vector<float> azimuth, std::vector<float> position_rx, int samp_rate,
                                              float center_freq, float self_coupling_db, bool rndm_phaseshift,
                                              bool self_coupling, const std::string &len_key, bool /*debug*/)
{
    return gnuradio::get_initial_sptr(new target_simulator_impl(range, velocity, rcs, azimuth, position_rx, samp_rate, center_freq,
                                                                self_coupling_db, rndm_phaseshift, self_coupling, len_key));
}

}  // namespace mimo_ofdm_jrc
}  // namespace gr
