// test_blocks.cc -- drives the C++ GNU Radio block wrappers (gr-mimo-ofdm-jrc_b200/lib) through the
// runtime stand-in, one general_work() call at a time, and checks them against the CPU oracle on
// identical inputs.  Needs a GPU (the blocks have no CPU path).  Run by tests/test_cpp_blocks.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <string>
#include <cstdlib>
#include <cstring>
#include <random>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include <mimo_ofdm_jrc/fft_peak_detect.h>
#include <mimo_ofdm_jrc/matrix_transpose.h>
#include <mimo_ofdm_jrc/mimo_ofdm_radar.h>
#include <mimo_ofdm_jrc/ofdm_cyclic_prefix_remover.h>
#include <mimo_ofdm_jrc/radar_log.h>
#include <mimo_ofdm_jrc/target_simulator.h>
#include <mimo_ofdm_jrc/radar_chain.h>
#include <mimo_ofdm_jrc/range_angle_estimator.h>
#include <mimo_ofdm_jrc/zero_pad.h>

#include <jrc_cuda.h>
#include "../../oracle/jrc_oracle.h"

using namespace gr;
using namespace gr::mimo_ofdm_jrc;
typedef std::vector<gr_complex> cvec;

static int g_fail = 0;
#define CHECK(cond, ...)                                                     \
    do {                                                                     \
        if (!(cond)) { g_fail++; std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); } \
    } while (0)

static std::mt19937 rng(12345);
static cvec randvec(size_t n, float scale = 1.f)
{
    std::normal_distribution<float> d(0.f, scale);
    cvec v(n);
    for (auto &z : v) z = gr_complex(d(rng), d(rng));
    return v;
}
static bool same(const gr_complex *a, const orc_c32 *b, size_t n) { return std::memcmp(a, b, n * sizeof(gr_complex)) == 0; }

static std::vector<float> range_bins(int nsc, int ir)
{
    std::vector<float> v(nsc * ir);
    double rmax = 3e8 * nsc / (2 * 125e6);
    for (int i = 0; i < nsc * ir; i++) v[i] = (float)(rmax * i / (nsc * ir - 1));
    return v;
}
static std::vector<float> angle_bins(int Na)
{
    std::vector<float> v(Na);
    for (int i = 0; i < Na; i++) v[i] = (float)(std::asin(2.0 / Na * (i - std::floor(Na / 2.0) + 0.5)) * 180.0 / M_PI);
    return v;
}

// one frame: per-port packets of (pre+S) fft_len-vectors
struct frame_t { std::vector<cvec> tx, rx; };
static frame_t make_frame(int T, int R, int items, int N)
{
    frame_t f;
    for (int t = 0; t < T; t++) f.tx.push_back(randvec((size_t)items * N));
    for (int r = 0; r < R; r++) f.rx.push_back(randvec((size_t)items * N));
    return f;
}

static void test_radar_block()
{
    const int N = 64, T = 4, R = 2, S = 4, pre = 5, IR = 8, V = T * R, items = pre + S + 3;
    for (int interleave = 0; interleave < 2; interleave++) {
        auto blk = mimo_ofdm_radar::make(N, T, R, S, pre, true, true, 3, IR, interleave, "/tmp/jrc_cpp_chan.csv");
        orc_radar *ref = orc_radar_create(N, T, R, S, pre, 1, 1, 3, IR, interleave);
        cvec out((size_t)V * N * IR), refout((size_t)V * N * IR);
        uint64_t rd = 0;
        for (int it = 0; it < 6; it++) {
            frame_t f = make_frame(T, R, items, N);
            std::vector<shim::input_t> in(T + R);
            for (int t = 0; t < T; t++) { in[t].items = f.tx[t].data(); in[t].n_items = items; }
            for (int r = 0; r < R; r++) { in[T + r].items = f.rx[r].data(); in[T + r].n_items = items; }
            in[0].tags.push_back(shim::make_tag(rd, "packet_len", pmt::from_long(items)));
            in[T].tags.push_back(shim::make_tag(rd, "packet_len", pmt::from_long(items)));
            in[T].tags.push_back(shim::make_tag(rd, "rx_time", pmt::make_tuple(pmt::from_uint64(1), pmt::from_double(0.5))));
            auto res = shim::run_once(*blk, in, {{out.data(), 64}});
            std::vector<const orc_c32 *> tp, rp;
            for (auto &v : f.tx) tp.push_back((const orc_c32 *)v.data());
            for (auto &v : f.rx) rp.push_back((const orc_c32 *)v.data());
            orc_radar_work(ref, tp.data(), rp.data(), 0, (orc_c32 *)refout.data());
            CHECK(res.produced == V, "radar produced %d", res.produced);
            CHECK(same(out.data(), (const orc_c32 *)refout.data(), out.size()), "radar output differs (interleave %d frame %d)", interleave, it);
            CHECK(res.out_tags[0].size() == 1 && pmt::to_long(res.out_tags[0][0].value) == V &&
                      res.out_tags[0][0].offset == (uint64_t)it * V && pmt::symbol_to_string(res.out_tags[0][0].srcid) == blk->alias(),
                  "radar output tag");
            for (int p = 0; p < T + R; p++) CHECK(res.consumed[p] == items, "radar consumed[%d]=%d", p, res.consumed[p]);
            rd += items;
            if (it == 3) { blk->set_background_record(false); orc_radar_set_background_record(ref, 0); }
        }
        orc_radar_destroy(ref);
        blk->capture_radar_data(true);
    }
    // stale TX frame in front + no-tag flush
    auto blk = mimo_ofdm_radar::make(N, T, R, S, pre, false, false, 1, IR, false, "/tmp/jrc_cpp_chan.csv");
    orc_radar *ref = orc_radar_create(N, T, R, S, pre, 0, 0, 1, IR, 0);
    frame_t f1 = make_frame(T, R, items, N), f2 = make_frame(T, R, items, N);
    std::vector<cvec> txcat(T);
    for (int t = 0; t < T; t++) { txcat[t] = f1.tx[t]; txcat[t].insert(txcat[t].end(), f2.tx[t].begin(), f2.tx[t].end()); }
    std::vector<shim::input_t> in(T + R);
    for (int t = 0; t < T; t++) { in[t].items = txcat[t].data(); in[t].n_items = 2 * items; }
    for (int r = 0; r < R; r++) { in[T + r].items = f2.rx[r].data(); in[T + r].n_items = items; }
    in[0].tags = {shim::make_tag(0, "packet_len", pmt::from_long(items)), shim::make_tag(items, "packet_len", pmt::from_long(items))};
    in[T].tags = {shim::make_tag(0, "packet_len", pmt::from_long(items))};
    cvec out((size_t)V * N * IR), refout((size_t)V * N * IR);
    auto res = shim::run_once(*blk, in, {{out.data(), 64}});
    std::vector<const orc_c32 *> tp, rp;
    for (auto &v : f2.tx) tp.push_back((const orc_c32 *)v.data());
    for (auto &v : f2.rx) rp.push_back((const orc_c32 *)v.data());
    orc_radar_work(ref, tp.data(), rp.data(), 0, (orc_c32 *)refout.data());
    CHECK(res.produced == V && same(out.data(), (const orc_c32 *)refout.data(), out.size()), "stale TX frame not skipped");
    CHECK(res.consumed[0] == 2 * items && res.consumed[T] == items, "stale TX consume %d %d", res.consumed[0], res.consumed[T]);
    for (auto &i : in) i.tags.clear();
    res = shim::run_once(*blk, in, {{out.data(), 64}});
    CHECK(res.produced == 0 && res.consumed[0] == 2 * items && res.consumed[T] == items, "no-tag flush");
    orc_radar_destroy(ref);
}

static void test_chain_of_blocks()
{
    // radar -> fft_vcc(IFFT) -> matrix_transpose -> fft_vcc(FFT, shift) -> range_angle_estimator,
    // block by block like the shipped flowgraph; the two stock fft_vxx blocks are jrc_fft_vcc calls.
    const int N = 64, T = 4, R = 2, S = 4, pre = 5, IR = 8, IA = 16, V = 8, Nr = N * IR, Na = V * IA, items = pre + S;
    auto rb = range_bins(N, IR); auto ab = angle_bins(Na);
    const float ndr = 2.4f, nda = 2 * 14.4775f;
    auto radar = mimo_ofdm_radar::make(N, T, R, S, pre, false, false, 8, IR, false, "/tmp/jrc_cpp_chan.csv");
    auto transp = matrix_transpose::make(Nr, V, IA, false);
    auto estim = range_angle_estimator::make(Na, rb, ab, ndr, nda, -100.f, 0.f, "/tmp/jrc_cpp_log.csv", true);
    jrc_chain_cfg ucfg{}; ucfg.fft_len = 64; ucfg.n_tx = ucfg.n_rx = ucfg.n_sym = 1; ucfg.interp_range = ucfg.interp_angle = 1;
    jrc_chain *util = nullptr;
    CHECK(jrc_chain_create(&ucfg, &util) == JRC_OK, "%s", jrc_last_error());
    orc_radar *ref = orc_radar_create(N, T, R, S, pre, 0, 0, 8, IR, 0);
    uint64_t rd = 0, rd2 = 0, rd3 = 0;
    for (int it = 0; it < 3; it++) {
        frame_t f = make_frame(T, R, items, N);
        std::vector<shim::input_t> in(T + R);
        for (int t = 0; t < T; t++) { in[t].items = f.tx[t].data(); in[t].n_items = items; }
        for (int r = 0; r < R; r++) { in[T + r].items = f.rx[r].data(); in[T + r].n_items = items; }
        in[0].tags.push_back(shim::make_tag(rd, "packet_len", pmt::from_long(items)));
        in[T].tags.push_back(shim::make_tag(rd, "packet_len", pmt::from_long(items)));
        rd += items;
        cvec pad((size_t)V * Nr), y((size_t)V * Nr), tr((size_t)Nr * Na), cm((size_t)Nr * Na);
        auto r1 = shim::run_once(*radar, in, {{pad.data(), 64}});
        CHECK(r1.produced == V, "radar");
        CHECK(jrc_fft_vcc(util, (const jrc_c32 *)pad.data(), (jrc_c32 *)y.data(), Nr, V, 0, 0) == JRC_OK, "%s", jrc_last_error());
        shim::input_t ti; ti.items = y.data(); ti.n_items = V; ti.tags.push_back(shim::make_tag(rd2, "packet_len", pmt::from_long(V)));
        rd2 += V;
        auto r2 = shim::run_once(*transp, {ti}, {{tr.data(), Nr}});
        CHECK(r2.produced == Nr && r2.consumed[0] == V, "transpose produced %d consumed %d", r2.produced, r2.consumed[0]);
        CHECK(!r2.out_tags[0].empty() && pmt::to_long(r2.out_tags[0].back().value) == Nr, "transpose length tag");
        CHECK(jrc_fft_vcc(util, (const jrc_c32 *)tr.data(), (jrc_c32 *)cm.data(), Na, Nr, 1, 1) == JRC_OK, "%s", jrc_last_error());
        shim::input_t ei; ei.items = cm.data(); ei.n_items = Nr; ei.tags.push_back(shim::make_tag(rd3, "packet_len", pmt::from_long(Nr)));
        rd3 += Nr;
        auto r3 = shim::run_once(*estim, {ei}, {});
        CHECK(r3.produced == 0 && r3.consumed[0] == Nr, "estimator is a sink that consumes the packet");
        // oracle, same data
        std::vector<const orc_c32 *> tp, rp;
        for (auto &v : f.tx) tp.push_back((const orc_c32 *)v.data());
        for (auto &v : f.rx) rp.push_back((const orc_c32 *)v.data());
        cvec opad((size_t)V * Nr), oy((size_t)V * Nr), otr((size_t)Nr * Na), ocm((size_t)Nr * Na);
        orc_radar_work(ref, tp.data(), rp.data(), 0, (orc_c32 *)opad.data());
        orc_fft_vcc_batch((orc_c32 *)opad.data(), (orc_c32 *)oy.data(), Nr, V, 0, 0);
        orc_matrix_transpose((orc_c32 *)oy.data(), V, Nr, V, IA, (orc_c32 *)otr.data());
        orc_fft_vcc_batch((orc_c32 *)otr.data(), (orc_c32 *)ocm.data(), Na, Nr, 1, 1);
        CHECK(same(cm.data(), (const orc_c32 *)ocm.data(), cm.size()), "chain of blocks: complex map differs");
        orc_det od;
        orc_range_angle_estimate((orc_c32 *)ocm.data(), Nr, Na, rb.data(), Nr, ab.data(), Na, ndr, nda, -100.f, 0.f, &od, nullptr);
        auto &msgs = estim->shim_published["params"];
        CHECK((int)msgs.size() == it + 1, "params message count %zu", msgs.size());
        if ((int)msgs.size() == it + 1) {
            auto m = msgs.back();
            auto field = [&](int k, const char *name) {
                auto pr = pmt::nth(k, m);
                CHECK(pmt::symbol_to_string(pmt::nth(0, pr)) == name, "message key %d", k);
                return pmt::f32vector_elements(pmt::nth(1, pr))[0];
            };
            CHECK(field(0, "range") == rb[od.range_idx] && field(1, "angle") == ab[od.angle_idx], "message range/angle");
            CHECK(field(2, "power") == od.peak_power && field(3, "snr") == od.snr_db, "message power/snr %g %g vs %g %g",
                  field(2, "power"), field(3, "snr"), od.peak_power, od.snr_db);
            // the log's consumer (mimo_precoder's radar-aided steering) reads the same detection back
            radar_log_entry le;
            CHECK(radar_log_read_last("/tmp/jrc_cpp_log.csv", le), "radar log unreadable");
            CHECK(std::fabs(le.angle - ab[od.angle_idx]) <= 1e-4f * (1.f + std::fabs(ab[od.angle_idx])) &&
                  std::fabs(le.range - rb[od.range_idx]) <= 1e-4f * (1.f + rb[od.range_idx]), "radar log %g %g", le.range, le.angle);
            auto sv = radar_aided_steering_vector(le.angle, T);
            CHECK((int)sv.size() == T && sv[0] == gr_complex(1.f, 0.f) && std::fabs(std::abs(sv[T - 1]) - 1.f) < 1e-6f, "steering vector");
        }
    }
    // back-pressure: the transpose block drops the CPI but still consumes it
    cvec y((size_t)V * Nr), tr((size_t)Nr * Na);
    shim::input_t ti; ti.items = y.data(); ti.n_items = V; ti.tags.push_back(shim::make_tag(rd2, "packet_len", pmt::from_long(V)));
    transp->shim_output_fullness = 0.5f;
    auto rdrop = shim::run_once(*transp, {ti}, {{tr.data(), Nr}});
    CHECK(rdrop.produced == 0 && rdrop.consumed[0] == V && rdrop.out_tags[0].empty(), "transpose back-pressure drop");
    orc_radar_destroy(ref);
    jrc_chain_destroy(util);
}

static void test_peak_and_pad()
{
    const int n = 40000;
    cvec x = randvec(n);
    x[31000] = gr_complex(40, 9);
    auto pk = fft_peak_detect::make(1000000, 8.0f, 10.0f, 25, {0.f}, false, "packet_len");
    float f = -1, ph = -1, mg = -1;
    shim::input_t in; in.items = x.data(); in.n_items = n; in.tags.push_back(shim::make_tag(0, "packet_len", pmt::from_long(n)));
    auto r = shim::run_once(*pk, {in}, {{&f, 1}, {&ph, 1}, {&mg, 1}});
    orc_peak1d o;
    orc_fft_peak_detect((orc_c32 *)x.data(), n, 1000000, 8.0f, 10.0f, 25, &o);
    CHECK(r.produced == 1 && o.k == 31000 && f == o.freq && ph == o.phase && mg == o.mag, "peak detect %g %g %g vs %g %g %g", f, ph, mg, o.freq, o.phase, o.mag);
    CHECK(r.out_tags.size() == 3 && pmt::to_long(r.out_tags[0][0].value) == 1, "peak detect length tag");
    pk->set_threshold(90.f);
    f = 7.f;
    in.tags[0].offset = n;
    r = shim::run_once(*pk, {in}, {{&f, 1}, {&ph, 1}, {&mg, 1}});
    CHECK(r.produced == 1 && f == 7.f, "no peak leaves the output untouched");

    auto zp = zero_pad::make(false, 7, 240);
    cvec y(720 + 247);
    shim::input_t zi; zi.items = x.data(); zi.n_items = 720; zi.tags.push_back(shim::make_tag(0, "packet_len", pmt::from_long(720)));
    r = shim::run_once(*zp, {zi}, {{y.data(), (int)y.size()}});
    CHECK(r.produced == 967 && std::memcmp(&y[7], x.data(), 720 * sizeof(gr_complex)) == 0, "zero_pad copy");
    CHECK(pmt::to_long(r.out_tags[0][0].value) == 967, "zero_pad length tag");
    double s2 = 0;
    for (int i = 727; i < 967; i++) s2 += std::norm(y[i]);
    CHECK(std::fabs(std::sqrt(s2 / 240 / 2) - 1e-2) < 2e-3, "zero_pad noise sigma %g", std::sqrt(s2 / 240 / 2));

    // ofdm_cyclic_prefix_remover: payload, length tag, and the packet-start tags travel to the first output item
    auto cp = ofdm_cyclic_prefix_remover::make(64, 16, "packet_len");
    cvec t = randvec(9 * 80), u(9 * 64), ou(9 * 64);
    shim::input_t ci; ci.items = t.data(); ci.n_items = 720;
    ci.tags = {shim::make_tag(0, "packet_len", pmt::from_long(720)), shim::make_tag(0, "rx_time", pmt::from_double(1.5))};
    r = shim::run_once(*cp, {ci}, {{u.data(), 9}});
    orc_cp_remove((const orc_c32 *)t.data(), 9, 64, 16, (orc_c32 *)ou.data());
    CHECK(r.produced == 9 && r.consumed[0] == 720 && std::memcmp(u.data(), ou.data(), u.size() * sizeof(gr_complex)) == 0, "cp remover");
    bool has_time = false, has_len = false;
    for (auto &tg : r.out_tags[0]) {
        if (pmt::symbol_to_string(tg.key) == "rx_time" && tg.offset == 0) has_time = true;
        if (pmt::symbol_to_string(tg.key) == "packet_len" && pmt::to_long(tg.value) == 9) has_len = true;
    }
    CHECK(has_time && has_len, "cp remover tags");

    // target_simulator: R output packets identical to the CPU restatement, rx_time tag on every port
    {
        std::vector<float> rg = {7.5f, 30.f}, vel = {3.f, -8.f}, rcs = {1.f, 20.f}, az = {-25.f, 40.f}, pos = {0.f, 0.00625f};
        auto sim = target_simulator::make(rg, vel, rcs, az, pos, 125000000, 24e9f, -10.f, false, true, "packet_len");
        const int n = 1040;
        cvec x = randvec(n), o0(n), o1(n), ro(2 * (size_t)n);
        shim::input_t si; si.items = x.data(); si.n_items = n; si.tags = {shim::make_tag(0, "packet_len", pmt::from_long(n))};
        auto rs = shim::run_once(*sim, {si}, {{o0.data(), n}, {o1.data(), n}});
        orc_target_simulator((const orc_c32 *)x.data(), n, rg.data(), vel.data(), rcs.data(), az.data(), 2, pos.data(), 2, 125000000,
                             24e9f, 1, -10.f, 0, (orc_c32 *)ro.data());
        CHECK(rs.produced == n && rs.consumed[0] == n, "target simulator produced %d", rs.produced);
        CHECK(std::memcmp(o0.data(), ro.data(), n * sizeof(gr_complex)) == 0 &&
              std::memcmp(o1.data(), ro.data() + n, n * sizeof(gr_complex)) == 0, "target simulator output differs");
        bool t0 = false, t1 = false;
        for (auto &tg : rs.out_tags[0]) if (pmt::symbol_to_string(tg.key) == "rx_time" && pmt::symbol_to_string(tg.srcid) == "stat_targ_sim") t0 = true;
        for (auto &tg : rs.out_tags[1]) if (pmt::symbol_to_string(tg.key) == "rx_time") t1 = true;
        CHECK(t0 && t1, "target simulator rx_time tags");
    }
}

static void test_fused_block()
{
    const int N = 64, T = 4, R = 2, S = 4, pre = 5, IR = 8, IA = 16, V = 8, Nr = N * IR, Na = V * IA, items = pre + S;
    auto rb = range_bins(N, IR); auto ab = angle_bins(Na);
    const float ndr = 2.4f, nda = 2 * 14.4775f;
    auto blk = radar_chain::make(N, T, R, S, pre, false, false, 8, IR, IA, false, rb, ab, ndr, nda, -100.f, 0.f, "/tmp/jrc_cpp_log2.csv", false);
    uint64_t rd = 0;
    for (int it = 0; it < 3; it++) {
        frame_t f = make_frame(T, R, items, N);
        // a point-target-like structure so the peak is well separated: rx = tx0 * phase ramp
        for (int r = 0; r < R; r++)
            for (int s = 0; s < items; s++)
                for (int k = 0; k < N; k++) {
                    gr_complex acc = 0;
                    for (int t = 0; t < T; t++)
                        acc += f.tx[t][(size_t)s * N + k] * std::polar(1.0f, (float)(-2 * M_PI * 0.13 * (it + 1) * k + 0.9 * (t + T * r)));
                    f.rx[r][(size_t)s * N + k] = acc;
                }
        std::vector<shim::input_t> in(T + R);
        for (int t = 0; t < T; t++) { in[t].items = f.tx[t].data(); in[t].n_items = items; }
        for (int r = 0; r < R; r++) { in[T + r].items = f.rx[r].data(); in[T + r].n_items = items; }
        in[0].tags.push_back(shim::make_tag(rd, "packet_len", pmt::from_long(items)));
        in[T].tags.push_back(shim::make_tag(rd, "packet_len", pmt::from_long(items)));
        rd += items;
        std::vector<float> map((size_t)Nr * Na);
        auto res = shim::run_once(*blk, in, {{map.data(), Nr}});
        CHECK(res.produced == Nr && pmt::to_long(res.out_tags[0][0].value) == Nr, "fused block produced %d", res.produced);
        // oracle chain on the preamble-stripped frame
        cvec rxp((size_t)R * S * N), txp((size_t)T * S * N);
        for (int t = 0; t < T; t++) std::memcpy(&txp[(size_t)t * S * N], &f.tx[t][(size_t)pre * N], sizeof(gr_complex) * S * N);
        for (int r = 0; r < R; r++) std::memcpy(&rxp[(size_t)r * S * N], &f.rx[r][(size_t)pre * N], sizeof(gr_complex) * S * N);
        orc_chain_cfg oc{N, T, R, S, 0, IR, IA, 0, rb.data(), ab.data(), ndr, nda, -100.f, 0.f};
        std::vector<float> omap((size_t)Nr * Na);
        orc_det od;
        orc_chain_batch(&oc, (orc_c32 *)rxp.data(), (orc_c32 *)txp.data(), 1, 1, 0, omap.data(), nullptr, &od);
        float peak = 0, err = 0;
        for (size_t i = 0; i < omap.size(); i++) { peak = std::max(peak, omap[i]); err = std::max(err, std::fabs(omap[i] - map[i])); }
        CHECK(err <= 1e-4f * peak, "fused block map error %g of peak", err / peak);
        auto &msgs = blk->shim_published["params"];
        CHECK((int)msgs.size() == it + 1, "fused block message");
        if (!msgs.empty()) {
            float rv = pmt::f32vector_elements(pmt::nth(1, pmt::nth(0, msgs.back())))[0];
            float av = pmt::f32vector_elements(pmt::nth(1, pmt::nth(1, msgs.back())))[0];
            CHECK(rv == rb[od.range_idx] && av == ab[od.angle_idx], "fused block peak (%g, %g) vs oracle (%g, %g)", rv, av, rb[od.range_idx], ab[od.angle_idx]);
        }
    }
}

// capture_radar_data(): the CSV line of the drop-in block against the line the REFERENCE block wrote for the same frame
// (tests/golden/c1_capture_line.txt, produced by tests/golden/make_golden.py from the reference's own sources;
// lib/mimo_ofdm_radar_impl.cc:348-377).  The time stamp in front of the first ", " is the only difference allowed.
// JRC_FUSED=1 on the unmodified five-block wiring: the radar block runs the whole chain once per frame, the two
// downstream blocks serve what it cached under the frame's jrc_cpi tag.  Outputs, tags and messages must be the ones
// the separate blocks give; the downstream blocks are handed ZEROED inputs here, so anything they computed themselves
// would show.
// (IR, IA) = (8, 16): the shipped 512x128 map; (16, 8): 1024x64; (4, 1): no angle zero-padding at all -- of the transposed
// array only the data columns are cached, the zero columns are written by the fetch.
static void test_fused_mode(const int IR, const int IA)
{
    const int N = 64, T = 4, R = 2, S = 4, pre = 5, V = 8, Nr = N * IR, Na = V * IA, items = pre + S;
    auto rb = range_bins(N, IR); auto ab = angle_bins(Na);
    const float ndr = 2.4f, nda = 2 * 14.4775f;
    setenv("JRC_FUSED", "1", 1);
    auto radar = mimo_ofdm_radar::make(N, T, R, S, pre, true, true, 3, IR, false, "/tmp/jrc_cpp_chan.csv");
    auto transp = matrix_transpose::make(Nr, V, IA, false);
    auto estim = range_angle_estimator::make(Na, rb, ab, ndr, nda, -100.f, 0.f, "/tmp/jrc_cpp_log.csv", false);
    unsetenv("JRC_FUSED");
    orc_radar *ref = orc_radar_create(N, T, R, S, pre, 1, 1, 3, IR, 0);
    uint64_t rd = 0, rd2 = 0, rd3 = 0;
    const cvec zeros_y((size_t)V * Nr), zeros_cm((size_t)Nr * Na);
    const int n_frames = JRC_FUSED_RING + 4;            // the ring of cached results wraps
    for (int it = 0; it < n_frames; it++) {
        frame_t f = make_frame(T, R, items, N);
        std::vector<shim::input_t> in(T + R);
        for (int t = 0; t < T; t++) { in[t].items = f.tx[t].data(); in[t].n_items = items; }
        for (int r = 0; r < R; r++) { in[T + r].items = f.rx[r].data(); in[T + r].n_items = items; }
        in[0].tags.push_back(shim::make_tag(rd, "packet_len", pmt::from_long(items)));
        in[T].tags.push_back(shim::make_tag(rd, "packet_len", pmt::from_long(items)));
        rd += items;
        cvec pad((size_t)V * Nr), tr((size_t)Nr * Na);
        auto r1 = shim::run_once(*radar, in, {{pad.data(), 64}});
        CHECK(r1.produced == V, "fused mode: radar produced %d", r1.produced);
        CHECK(r1.out_tags[0].size() == 2 && pmt::symbol_to_string(r1.out_tags[0][0].key) == "packet_len" &&
                  pmt::to_long(r1.out_tags[0][0].value) == V && r1.out_tags[0][0].offset == (uint64_t)it * V &&
                  pmt::symbol_to_string(r1.out_tags[0][1].key) == "jrc_cpi" && pmt::to_long(r1.out_tags[0][1].value) == it &&
                  r1.out_tags[0][1].offset == (uint64_t)it * V,
              "fused mode: radar output tags");
        // oracle, block by block
        std::vector<const orc_c32 *> tp, rp;
        for (auto &v : f.tx) tp.push_back((const orc_c32 *)v.data());
        for (auto &v : f.rx) rp.push_back((const orc_c32 *)v.data());
        cvec opad((size_t)V * Nr), oy((size_t)V * Nr), otr((size_t)Nr * Na), ocm((size_t)Nr * Na);
        orc_radar_work(ref, tp.data(), rp.data(), 0, (orc_c32 *)opad.data());
        orc_fft_vcc_batch((orc_c32 *)opad.data(), (orc_c32 *)oy.data(), Nr, V, 0, 0);
        orc_matrix_transpose((orc_c32 *)oy.data(), V, Nr, V, IA, (orc_c32 *)otr.data());
        orc_fft_vcc_batch((orc_c32 *)otr.data(), (orc_c32 *)ocm.data(), Na, Nr, 1, 1);
        orc_det od;
        orc_range_angle_estimate((orc_c32 *)ocm.data(), Nr, Na, rb.data(), Nr, ab.data(), Na, ndr, nda, -100.f, 0.f, &od, nullptr);
        CHECK(same(pad.data(), (const orc_c32 *)opad.data(), pad.size()), "fused mode: radar output differs (frame %d)", it);
        // the stock fft_vcc between the blocks keeps the tags where they are (sync block, TPP_ALL_TO_ALL)
        shim::input_t ti; ti.items = zeros_y.data(); ti.n_items = V;
        ti.tags.push_back(shim::make_tag(rd2, "packet_len", pmt::from_long(V)));
        ti.tags.push_back(shim::make_tag(rd2, "jrc_cpi", pmt::from_long(it)));
        rd2 += V;
        auto r2 = shim::run_once(*transp, {ti}, {{tr.data(), Nr}});
        CHECK(r2.produced == Nr && r2.consumed[0] == V, "fused mode: transpose produced %d consumed %d", r2.produced, r2.consumed[0]);
        CHECK(same(tr.data(), (const orc_c32 *)otr.data(), tr.size()), "fused mode: transposed array differs (frame %d)", it);
        bool have_len = false, have_seq = false;
        for (auto &t : r2.out_tags[0]) {
            if (pmt::symbol_to_string(t.key) == "packet_len") have_len = pmt::to_long(t.value) == Nr && t.offset == (uint64_t)it * Nr;
            if (pmt::symbol_to_string(t.key) == "jrc_cpi") have_seq = pmt::to_long(t.value) == it && t.offset == (uint64_t)it * Nr;
        }
        CHECK(have_len && have_seq && r2.out_tags[0].size() == 2, "fused mode: transpose output tags");
        shim::input_t ei; ei.items = zeros_cm.data(); ei.n_items = Nr;
        ei.tags.push_back(shim::make_tag(rd3, "packet_len", pmt::from_long(Nr)));
        ei.tags.push_back(shim::make_tag(rd3, "jrc_cpi", pmt::from_long(it)));
        rd3 += Nr;
        auto r3 = shim::run_once(*estim, {ei}, {});
        CHECK(r3.produced == 0 && r3.consumed[0] == Nr, "fused mode: estimator consumes the packet");
        auto &msgs = estim->shim_published["params"];
        CHECK((int)msgs.size() == it + 1, "fused mode: params message count %zu", msgs.size());
        if ((int)msgs.size() == it + 1) {
            auto m = msgs.back();
            auto field = [&](int k) { return pmt::f32vector_elements(pmt::nth(1, pmt::nth(k, m)))[0]; };
            CHECK(field(0) == rb[od.range_idx] && field(1) == ab[od.angle_idx] && field(2) == od.peak_power && field(3) == od.snr_db,
                  "fused mode: message differs (frame %d): %g %g %g %g vs %g %g %g %g", it, field(0), field(1), field(2), field(3),
                  rb[od.range_idx], ab[od.angle_idx], od.peak_power, od.snr_db);
        }
        if (it == 5) { radar->set_background_record(false); orc_radar_set_background_record(ref, 0); }
    }
    // a packet whose cached result is gone (frame 0 was overwritten by frame JRC_FUSED_RING) or that carries no
    // jrc_cpi tag goes through the block's own device call: the input decides
    cvec y = randvec((size_t)V * Nr), tr((size_t)Nr * Na), otr((size_t)Nr * Na);
    orc_matrix_transpose((orc_c32 *)y.data(), V, Nr, V, IA, (orc_c32 *)otr.data());
    for (int tagged = 0; tagged < 2; tagged++) {
        shim::input_t ti; ti.items = y.data(); ti.n_items = V;
        ti.tags.push_back(shim::make_tag(rd2, "packet_len", pmt::from_long(V)));
        if (tagged) ti.tags.push_back(shim::make_tag(rd2, "jrc_cpi", pmt::from_long(0)));
        rd2 += V;
        auto r2 = shim::run_once(*transp, {ti}, {{tr.data(), Nr}});
        CHECK(r2.produced == Nr && same(tr.data(), (const orc_c32 *)otr.data(), tr.size()) && r2.out_tags[0].size() == 1,
              "fused mode: fallback of matrix_transpose (tagged %d)", tagged);
    }
    // the estimator's own thresholds gate a cached record
    estim->set_snr_threshold(1000.f);
    shim::input_t ei; ei.items = zeros_cm.data(); ei.n_items = Nr;
    ei.tags.push_back(shim::make_tag(rd3, "packet_len", pmt::from_long(Nr)));
    ei.tags.push_back(shim::make_tag(rd3, "jrc_cpi", pmt::from_long(n_frames - 1)));
    const size_t before = estim->shim_published["params"].size();
    shim::run_once(*estim, {ei}, {});
    CHECK(estim->shim_published["params"].size() == before, "fused mode: threshold change ignored");
    orc_radar_destroy(ref);

    // blocks that do not continue each other: the radar block says so and stays on its own call
    setenv("JRC_FUSED", "1", 1);
    const int IR2 = IR == 4 ? 2 : 4;
    auto radar2 = mimo_ofdm_radar::make(N, T, R, S, pre, false, false, 1, IR2, false, "/tmp/jrc_cpp_chan.csv");
    unsetenv("JRC_FUSED");
    frame_t f = make_frame(T, R, items, N);
    std::vector<shim::input_t> in(T + R);
    for (int t = 0; t < T; t++) { in[t].items = f.tx[t].data(); in[t].n_items = items; }
    for (int r = 0; r < R; r++) { in[T + r].items = f.rx[r].data(); in[T + r].n_items = items; }
    in[0].tags.push_back(shim::make_tag(0, "packet_len", pmt::from_long(items)));
    in[T].tags.push_back(shim::make_tag(0, "packet_len", pmt::from_long(items)));
    cvec pad((size_t)V * N * IR2);
    auto r1 = shim::run_once(*radar2, in, {{pad.data(), 64}});
    CHECK(r1.produced == V && r1.out_tags[0].size() == 1, "fused mode: mismatching blocks must fall back");
}

// JRC_FUSED=1 with the blocks on their own threads, as under the GNU Radio scheduler: the radar block runs ahead on one
// thread (far enough at times that the ring of cached results wraps over frames the consumers have not fetched yet), the
// two downstream blocks follow on another.  Every packet carries the RIGHT input as well, so a fetch that finds its frame
// overwritten falls back to the block's own call and the outputs must still equal the oracle's, frame by frame.
static void test_fused_mode_threads()
{
    const int N = 64, T = 4, R = 2, S = 4, pre = 5, IR = 16, IA = 8, V = 8, Nr = N * IR, Na = V * IA, items = pre + S;
    auto rb = range_bins(N, IR); auto ab = angle_bins(Na);
    const float ndr = 2.4f, nda = 2 * 14.4775f;
    setenv("JRC_FUSED", "1", 1);
    auto radar = mimo_ofdm_radar::make(N, T, R, S, pre, false, false, 1, IR, false, "/tmp/jrc_cpp_chan.csv");
    auto transp = matrix_transpose::make(Nr, V, IA, false);
    auto estim = range_angle_estimator::make(Na, rb, ab, ndr, nda, -100.f, 0.f, "/tmp/jrc_cpp_log.csv", false);
    unsetenv("JRC_FUSED");
    const int n_frames = 120;
    struct job_t { int it; cvec y, cm; };               // what the stock fft_vcc blocks would hand on
    std::deque<job_t> queue;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<cvec> want_tr(n_frames);
    std::vector<orc_det> want_det(n_frames);
    std::atomic<int> bad_tr{0}, bad_msg{0}, served{0};
    std::thread consumer([&]() {
        uint64_t rd2 = 0, rd3 = 0;
        for (int done = 0; done < n_frames; done++) {
            job_t j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !queue.empty(); });
                j = std::move(queue.front());
                queue.pop_front();
            }
            cvec tr((size_t)Nr * Na);
            shim::input_t ti; ti.items = j.y.data(); ti.n_items = V;
            ti.tags.push_back(shim::make_tag(rd2, "packet_len", pmt::from_long(V)));
            ti.tags.push_back(shim::make_tag(rd2, "jrc_cpi", pmt::from_long(j.it)));
            rd2 += V;
            auto r2 = shim::run_once(*transp, {ti}, {{tr.data(), Nr}});
            if (r2.produced != Nr || !same(tr.data(), (const orc_c32 *)want_tr[j.it].data(), tr.size())) bad_tr++;
            if (r2.out_tags[0].size() == 2) served++;
            shim::input_t ei; ei.items = j.cm.data(); ei.n_items = Nr;
            ei.tags.push_back(shim::make_tag(rd3, "packet_len", pmt::from_long(Nr)));
            ei.tags.push_back(shim::make_tag(rd3, "jrc_cpi", pmt::from_long(j.it)));
            rd3 += Nr;
            shim::run_once(*estim, {ei}, {});
            auto &msgs = estim->shim_published["params"];
            const orc_det &od = want_det[j.it];
            if (msgs.size() != 1) { bad_msg++; msgs.clear(); continue; }
            auto field = [&](int k) { return pmt::f32vector_elements(pmt::nth(1, pmt::nth(k, msgs[0])))[0]; };
            if (!(field(0) == rb[od.range_idx] && field(1) == ab[od.angle_idx] && field(2) == od.peak_power && field(3) == od.snr_db)) bad_msg++;
            msgs.clear();
        }
    });
    orc_radar *ref = orc_radar_create(N, T, R, S, pre, 0, 0, 1, IR, 0);
    uint64_t rd = 0;
    int bad_pad = 0;
    for (int it = 0; it < n_frames; it++) {
        frame_t f = make_frame(T, R, items, N);
        std::vector<const orc_c32 *> tp, rp;
        for (auto &v : f.tx) tp.push_back((const orc_c32 *)v.data());
        for (auto &v : f.rx) rp.push_back((const orc_c32 *)v.data());
        cvec opad((size_t)V * Nr), oy((size_t)V * Nr), otr((size_t)Nr * Na), ocm((size_t)Nr * Na);
        orc_radar_work(ref, tp.data(), rp.data(), 0, (orc_c32 *)opad.data());
        orc_fft_vcc_batch((orc_c32 *)opad.data(), (orc_c32 *)oy.data(), Nr, V, 0, 0);
        orc_matrix_transpose((orc_c32 *)oy.data(), V, Nr, V, IA, (orc_c32 *)otr.data());
        orc_fft_vcc_batch((orc_c32 *)otr.data(), (orc_c32 *)ocm.data(), Na, Nr, 1, 1);
        orc_range_angle_estimate((orc_c32 *)ocm.data(), Nr, Na, rb.data(), Nr, ab.data(), Na, ndr, nda, -100.f, 0.f, &want_det[it], nullptr);
        want_tr[it] = otr;
        std::vector<shim::input_t> in(T + R);
        for (int t = 0; t < T; t++) { in[t].items = f.tx[t].data(); in[t].n_items = items; }
        for (int r = 0; r < R; r++) { in[T + r].items = f.rx[r].data(); in[T + r].n_items = items; }
        in[0].tags.push_back(shim::make_tag(rd, "packet_len", pmt::from_long(items)));
        in[T].tags.push_back(shim::make_tag(rd, "packet_len", pmt::from_long(items)));
        rd += items;
        cvec pad((size_t)V * Nr);
        auto r1 = shim::run_once(*radar, in, {{pad.data(), 64}});
        if (r1.produced != V || !same(pad.data(), (const orc_c32 *)opad.data(), pad.size())) bad_pad++;
        {
            std::lock_guard<std::mutex> lk(mu);
            queue.push_back(job_t{it, std::move(oy), std::move(ocm)});
        }
        // hand the frames over in bursts: the radar block gets up to 40 frames ahead of the consumers
        if (it % 40 == 39 || it == n_frames - 1) cv.notify_one();
    }
    cv.notify_one();
    consumer.join();
    orc_radar_destroy(ref);
    CHECK(bad_pad == 0 && bad_tr.load() == 0 && bad_msg.load() == 0, "fused mode on two threads: %d radar outputs, %d transposed arrays, %d messages differ",
          bad_pad, bad_tr.load(), bad_msg.load());
    CHECK(served.load() > 0 && served.load() < n_frames, "fused mode on two threads: %d of %d frames served from the cache (expected some of each kind)",
          served.load(), n_frames);
    std::printf("fused mode on two threads: %d of %d frames served from the cache, the rest through the blocks' own calls\n", served.load(), n_frames);
}

static void test_capture_format(const char *golden_dir)
{
    const int N = 64, T = 4, R = 2, S = 4, V = T * R;
    std::string dir(golden_dir);
    std::ifstream ff(dir + "/c1_frame0.c64", std::ios::binary), gf(dir + "/c1_capture_line.txt");
    CHECK(ff.good() && gf.good(), "golden files not found in %s", golden_dir);
    if (!ff.good() || !gf.good()) return;
    cvec frame((size_t)(T + R) * S * N);
    ff.read(reinterpret_cast<char *>(frame.data()), frame.size() * sizeof(gr_complex));
    std::string golden((std::istreambuf_iterator<char>(gf)), std::istreambuf_iterator<char>());
    const char *path = "/tmp/jrc_cpp_capture.csv";
    std::remove(path);
    auto blk = mimo_ofdm_radar::make(N, T, R, S, 0, false, false, 1, 1, false, path);
    std::vector<shim::input_t> in(T + R);
    for (int t = 0; t < T; t++) { in[t].items = frame.data() + (size_t)t * S * N; in[t].n_items = S; }
    for (int r = 0; r < R; r++) { in[T + r].items = frame.data() + (size_t)(T + r) * S * N; in[T + r].n_items = S; }
    in[0].tags.push_back(shim::make_tag(0, "packet_len", pmt::from_long(S)));
    in[T].tags.push_back(shim::make_tag(0, "packet_len", pmt::from_long(S)));
    cvec out((size_t)V * N);
    auto res = shim::run_once(*blk, in, {{out.data(), V}});
    CHECK(res.produced == V, "capture: radar produced %d", res.produced);
    blk->capture_radar_data(true);
    std::ifstream cf(path);
    std::string line((std::istreambuf_iterator<char>(cf)), std::istreambuf_iterator<char>());
    const size_t cut = line.find(", ");
    CHECK(cut != std::string::npos && cut == 12, "capture: time stamp HH:MM:SS.mmm expected in front (found at %zu)", cut);
    CHECK(cut != std::string::npos && line.substr(cut + 2) == golden, "capture: CSV line differs from the reference block's");
}

int main()
{
    if (const char *g = std::getenv("JRC_GOLDEN_DIR")) test_capture_format(g);
    test_radar_block();
    test_chain_of_blocks();
    test_peak_and_pad();
    test_fused_block();
    test_fused_mode(8, 16);
    test_fused_mode(16, 8);
    test_fused_mode(4, 1);
    test_fused_mode_threads();
    if (g_fail) { std::printf("%d check(s) FAILED\n", g_fail); return 1; }
    std::printf("ALL BLOCK TESTS PASSED\n");
    return 0;
}
