// range_angle_estimator on B200 (jrc_estimate2d).  Replaces lib/range_angle_estimator_impl.cc:114-302.
#include <mimo_ofdm_jrc/range_angle_estimator.h>

#include <gnuradio/io_signature.h>

#include "jrc_host.h"

namespace gr {
namespace mimo_ofdm_jrc {

class range_angle_estimator_impl : public range_angle_estimator
{
    const int d_vlen;
    const std::vector<float> d_range_bins, d_angle_bins;
    const float d_nd_range_m, d_nd_angle_deg;
    float d_snr_threshold, d_power_threshold;
    host::stats_log d_log;
    const bool d_debug;
    const bool d_fused;                 // JRC_FUSED=1 (jrc_host.h, fused_session)
    host::chain_handle d_chain;

    void push_thresholds() { host::check(jrc_chain_set_thresholds(d_chain.get(), d_snr_threshold, d_power_threshold), "RANGE-ANGLE ESTIMATOR"); }

protected:
    int calculate_output_stream_length(const gr_vector_int &) override { return 0; }

public:
    range_angle_estimator_impl(int vlen, const std::vector<float> &range_bins, const std::vector<float> &angle_bins,
                               float nd_range_m, float nd_angle_deg, float snr_threshold, float power_threshold,
                               const std::string &stats_path, bool stats_record, const std::string &len_key, bool debug)
        : gr::tagged_stream_block("range_angle_estimator", gr::io_signature::make(1, 1, sizeof(gr_complex) * vlen),
                                  gr::io_signature::make(0, 0, 0), len_key),
          d_vlen(vlen), d_range_bins(range_bins), d_angle_bins(angle_bins), d_nd_range_m(nd_range_m),
          d_nd_angle_deg(nd_angle_deg), d_snr_threshold(snr_threshold), d_power_threshold(power_threshold), d_debug(debug),
          d_fused(host::fused_session::requested()), d_chain(host::utility_cfg(), "RANGE-ANGLE ESTIMATOR")
    {
        if (d_fused)
            host::fused_session::get().register_estimator(vlen, range_bins, angle_bins, nd_range_m, nd_angle_deg, snr_threshold,
                                                          power_threshold);
        message_port_register_out(pmt::mp("params"));
        d_log.path = stats_path; d_log.record = stats_record;
        std::ofstream probe(stats_path, std::ofstream::app);
        if (!probe.is_open()) std::cerr << "[RANGE-ANGLE ESTIMATOR] Could not open log file at " << stats_path << std::endl;
        host::check(jrc_chain_set_estimator(d_chain.get(), d_range_bins.data(), (int)d_range_bins.size(), d_angle_bins.data(),
                                            (int)d_angle_bins.size(), nd_range_m, nd_angle_deg, snr_threshold, power_threshold),
                    "RANGE-ANGLE ESTIMATOR");
    }

    void set_snr_threshold(float v) override { d_snr_threshold = v; push_thresholds(); }
    void set_power_threshold(float v) override { d_power_threshold = v; push_thresholds(); }
    void set_stats_record(bool on) override { d_log.record = on; d_log.header_written = false; }

    int work(int, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &) override
    {
        jrc_det det;
        bool served = false;
        if (d_fused) {
            // the record of this frame was made when the radar block ran the chain; the gate uses this block's thresholds
            const int64_t seq = host::fused_session::packet_seq(*this, ninput_items[0]);
            auto chain = seq >= 0 ? host::fused_session::get().chain() : nullptr;
            served = chain && ninput_items[0] == (int)d_range_bins.size() &&
                     jrc_fused_fetch_det(chain->get(), seq, d_snr_threshold, d_power_threshold, &det) == JRC_OK;
        }
        if (!served)
            host::check(jrc_estimate2d(d_chain.get(), static_cast<const jrc_c32 *>(input_items[0]), ninput_items[0], d_vlen, &det),
                        "RANGE-ANGLE ESTIMATOR");
        if (d_debug)
            std::cout << "[RANGE-ANGLE ESTIMATOR] peak (" << det.range_idx << ", " << det.angle_idx << ") power " << det.peak_power
                      << " noise " << det.noise_power << " snr " << det.snr_db << std::endl;
        if (det.flags & JRC_DET_PASSED) {
            const float range_val = d_range_bins[det.range_idx], angle_val = d_angle_bins[det.angle_idx];
            message_port_pub(pmt::mp("params"), host::params_message(range_val, angle_val, det.peak_power, det.snr_db));
            if (d_log.record && !d_log.append(det.peak_power, det.snr_db, range_val, angle_val))
                throw std::runtime_error("[STREAM DECODER] Could not open file!!");   // message kept from the reference (:277)
        }
        return 0;   // sink: the base class consumes the packet
    }
};

range_angle_estimator::sptr range_angle_estimator::make(int vlen, std::vector<float> range_bins, std::vector<float> angle_bins,
                                                        float noise_discard_range_m, float noise_discard_angle_deg,
                                                        float snr_threshold, float power_threshold, const std::string &stats_path,
                                                        bool stats_record, const std::string &len_key, bool debug)
{
    return gnuradio::get_initial_sptr(new range_angle_estimator_impl(vlen, range_bins, angle_bins, noise_discard_range_m,
                                                                     noise_discard_angle_deg, snr_threshold, power_threshold,
                                                                     stats_path, stats_record, len_key, debug));
}

}  // namespace mimo_ofdm_jrc
}  // namespace gr
