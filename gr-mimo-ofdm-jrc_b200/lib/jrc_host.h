// jrc_host.h -- glue shared by the block wrappers: RAII around the C-ABI handle, status -> exception
// (the reference's error convention: work()/constructors throw std::runtime_error), time stamps of
// the two log formats (lib/utils.cc:295-321 of the reference), and the frame bookkeeping of
// mimo_ofdm_radar::general_work (lib/mimo_ofdm_radar_impl.cc:153-241).
#pragma once
#include <jrc_cuda.h>

#include <chrono>
#include <cstdlib>
#include <mutex>
#include <cstdint>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include <gnuradio/block.h>

namespace gr {
namespace mimo_ofdm_jrc {
namespace host {

inline void check(jrc_status st, const char *who)
{
    if (st != JRC_OK) throw std::runtime_error(std::string("[") + who + "] " + jrc_last_error());
}

class chain_handle
{
    jrc_chain *d_h = nullptr;

public:
    chain_handle() {}
    explicit chain_handle(const jrc_chain_cfg &cfg, const char *who) { open(cfg, who); }
    chain_handle(const chain_handle &) = delete;
    chain_handle &operator=(const chain_handle &) = delete;
    ~chain_handle() { jrc_chain_destroy(d_h); }
    // The make() signatures of the reference have no device argument: the CUDA device of a block comes from the
    // environment (JRC_DEVICE=<ordinal>, default 0), one flowgraph process per GPU.
    void open(jrc_chain_cfg cfg, const char *who)
    {
        jrc_chain_destroy(d_h);
        d_h = nullptr;
        if (const char *e = std::getenv("JRC_DEVICE")) cfg.device = std::atoi(e);
        check(jrc_chain_create(&cfg, &d_h), who);
    }
    jrc_chain *get() const { return d_h; }
};

inline jrc_chain_cfg utility_cfg()
{
    jrc_chain_cfg c{};
    c.fft_len = 64; c.n_tx = 1; c.n_rx = 1; c.n_sym = 1; c.interp_range = 1; c.interp_angle = 1;
    return c;
}

// ---- fused mode for an UNMODIFIED flowgraph (JRC_FUSED=1) ---------------------------------------
// The make() signatures stay the reference's, so no block knows the whole chain: mimo_ofdm_radar has the frame
// geometry and the range interpolation, matrix_transpose the angle interpolation, range_angle_estimator the bin axes
// and the noise window.  Each of the three leaves its share here when it is constructed; on its first frame the
// radar block puts them together, opens ONE handle for the whole chain and from then on runs
// jrc_radar_estimate_fused(), tagging its packet with the sequence number ("jrc_cpi").  The stock fft_vcc blocks
// between the three pass the tag along (sync blocks, TPP_ALL_TO_ALL); matrix_transpose and range_angle_estimator
// fetch their result for that number instead of a device round trip of their own, and fall back to their
// per-block call whenever the tag or the cached result is missing.
class fused_session
{
    std::mutex d_mu;
    bool d_have_radar = false, d_have_transpose = false, d_have_estimator = false;
    jrc_chain_cfg d_radar{};
    int d_tr_input_len = 0, d_tr_output_len = 0, d_tr_interp = 0;
    int d_vlen = 0;
    std::vector<float> d_range_bins, d_angle_bins;
    float d_nd_range_m = 0, d_nd_angle_deg = 0, d_snr_thr = 0, d_pow_thr = 0;
    std::weak_ptr<chain_handle> d_chain;

public:
    static fused_session &get() { static fused_session s; return s; }
    static bool requested()
    {
        const char *e = std::getenv("JRC_FUSED");
        return e && std::atoi(e) != 0;
    }
    static const pmt::pmt_t &tag_key() { static const pmt::pmt_t k = pmt::string_to_symbol("jrc_cpi"); return k; }

    void register_radar(const jrc_chain_cfg &cfg) { std::lock_guard<std::mutex> g(d_mu); d_radar = cfg; d_have_radar = true; }
    void register_transpose(int input_len, int output_len, int interp)
    {
        std::lock_guard<std::mutex> g(d_mu);
        d_tr_input_len = input_len; d_tr_output_len = output_len; d_tr_interp = interp; d_have_transpose = true;
    }
    void register_estimator(int vlen, const std::vector<float> &rb, const std::vector<float> &ab, float nd_range_m,
                            float nd_angle_deg, float snr_thr, float pow_thr)
    {
        std::lock_guard<std::mutex> g(d_mu);
        d_vlen = vlen; d_range_bins = rb; d_angle_bins = ab; d_nd_range_m = nd_range_m; d_nd_angle_deg = nd_angle_deg;
        d_snr_thr = snr_thr; d_pow_thr = pow_thr; d_have_estimator = true;
    }

    // the radar block's first frame: one handle for the whole chain, or nullptr and why
    std::shared_ptr<chain_handle> open(const jrc_chain_cfg &radar_cfg, const char *who, std::string *why)
    {
        std::lock_guard<std::mutex> g(d_mu);
        const int Nr = radar_cfg.fft_len * radar_cfg.interp_range, V = radar_cfg.n_tx * radar_cfg.n_rx;
        if (!d_have_transpose || !d_have_estimator) { *why = "no matrix_transpose / range_angle_estimator in this process"; return nullptr; }
        if (d_tr_input_len != Nr || d_tr_output_len != V) { *why = "matrix_transpose sizes do not continue this block's output"; return nullptr; }
        const int Na = V * d_tr_interp;
        if (d_vlen != Na || (int)d_range_bins.size() != Nr || (int)d_angle_bins.size() != Na) {
            *why = "range_angle_estimator sizes do not continue matrix_transpose's output";
            return nullptr;
        }
        jrc_chain_cfg full = radar_cfg;
        full.interp_angle = d_tr_interp;
        auto chain = std::make_shared<chain_handle>(full, who);
        check(jrc_chain_set_estimator(chain->get(), d_range_bins.data(), Nr, d_angle_bins.data(), Na, d_nd_range_m, d_nd_angle_deg,
                                      d_snr_thr, d_pow_thr), who);
        d_chain = chain;
        return chain;
    }
    std::shared_ptr<chain_handle> chain() { std::lock_guard<std::mutex> g(d_mu); return d_chain.lock(); }

    // sequence number on the packet that starts at nitems_read(0), or -1
    static int64_t packet_seq(gr::block &blk, int n_items)
    {
        std::vector<gr::tag_t> tags;
        blk.get_tags_in_range(tags, 0, blk.nitems_read(0), blk.nitems_read(0) + (uint64_t)n_items, tag_key());
        return tags.empty() ? -1 : (int64_t)pmt::to_long(tags[0].value);
    }
};

// "MM-DD-YYYY HH:MM:SS" and "HH:MM:SS.mmm"
inline std::string date_time_stamp()
{
    std::time_t now = std::time(nullptr);
    std::tm tmv = *std::localtime(&now);
    char buf[64];
    std::strftime(buf, sizeof(buf), "%m-%d-%Y %H:%M:%S", &tmv);
    return buf;
}
inline std::string time_ms_stamp()
{
    using namespace std::chrono;
    auto now = system_clock::now();
    auto ms = duration_cast<milliseconds>(now.time_since_epoch()).count() % 1000;
    std::time_t t = system_clock::to_time_t(now);
    std::tm tmv = *std::localtime(&t);
    std::ostringstream os;
    os << std::put_time(&tmv, "%H:%M:%S") << '.' << std::setfill('0') << std::setw(3) << ms;
    return os.str();
}

// What one general_work() call of the radar block has to do with the items in its buffers.
struct frame_plan {
    enum action_t { NO_RX_TAG, WAIT, DROP_RX, PROCESS } action = NO_RX_TAG;
    uint64_t tx_skip_items = 0;     // stale TX frames in front of the matching one
    uint64_t rx_packet_len = 0, tx_packet_len = 0;
};

// RX port 0 (stream N_tx) and TX port 0 carry "packet_len" tags (the key is fixed in the reference,
// :167,:176).  More TX than RX tags in the window -> the oldest TX frames are stale and are skipped
// (:189-197).  Departures from the reference, both where it reads out of bounds: a frame that is
// not completely in the buffers yet is waited for, and an RX frame without any TX tag is dropped.
inline frame_plan plan_frame(gr::block &blk, int n_tx, const gr_vector_int &ninput_items, int frame_items)
{
    static const pmt::pmt_t key = pmt::string_to_symbol("packet_len");
    frame_plan p;
    std::vector<gr::tag_t> rx_tags, tx_tags;
    blk.get_tags_in_range(rx_tags, n_tx, blk.nitems_read(n_tx), blk.nitems_read(n_tx) + ninput_items[n_tx], key);
    if (rx_tags.empty()) return p;
    blk.get_tags_in_range(tx_tags, 0, blk.nitems_read(0), blk.nitems_read(0) + ninput_items[0], key);
    p.rx_packet_len = pmt::to_uint64(rx_tags[0].value);
    if (tx_tags.empty()) { p.action = frame_plan::DROP_RX; return p; }
    size_t off = tx_tags.size() > rx_tags.size() ? tx_tags.size() - rx_tags.size() : 0;
    for (size_t i = 0; i < off; i++) p.tx_skip_items += pmt::to_uint64(tx_tags[i].value);
    p.tx_packet_len = pmt::to_uint64(tx_tags[off].value);
    bool complete = true;
    for (size_t i = 0; i < ninput_items.size(); i++) {
        uint64_t need = (int)i < n_tx ? p.tx_skip_items + (uint64_t)frame_items : (uint64_t)frame_items;
        if ((uint64_t)ninput_items[i] < need) complete = false;
    }
    p.action = complete ? frame_plan::PROCESS : frame_plan::WAIT;
    return p;
}

// publishes the estimator's message and appends the CSV line exactly when and how
// lib/range_angle_estimator_impl.cc:234-279 does
struct stats_log {
    std::string path;
    bool record = false, header_written = false;
    bool append(float power, float snr, float range, float angle)
    {
        std::ofstream f(path, std::ofstream::app);
        if (!f.is_open()) return false;
        if (!header_written) { f << "\n NEW RECORD - " << date_time_stamp() << "\n"; header_written = true; }
        f << time_ms_stamp() << ", \t" << power << ", \t" << snr << ", \t" << range << ", \t" << angle << "\n";
        f.flush();
        return true;
    }
};

inline pmt::pmt_t params_message(float range, float angle, float power, float snr)
{
    auto pack = [](const char *name, float v) {
        return pmt::list2(pmt::string_to_symbol(name), pmt::init_f32vector(1, &v));
    };
    return pmt::list4(pack("range", range), pack("angle", angle), pack("power", power), pack("snr", snr));
}

}  // namespace host
}  // namespace mimo_ofdm_jrc
}  // namespace gr
