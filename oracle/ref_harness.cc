// ref_harness.cc -- builds the REFERENCE's own block sources (compiled where they lie under
// /root/reference/lib, see build_ref.sh) into oracle/_ref and
//   (1) main(): drives them through the GNU Radio runtime stand-in on seeded inputs and compares
//       every result with the oracle restatement bit for bit  -> pins the oracle;
//   (2) ref_chain_batch(): the reference chain as a C entry point (reference blocks + the oracle's
//       fft_vcc arithmetic for the two stock GNU Radio FFT blocks, which the reference tree does
//       not contain) -> bench.py's CPU baseline of kind "reference".
// TEST INFRASTRUCTURE ONLY.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>

#include <mimo_ofdm_jrc/fft_peak_detect.h>
#include <mimo_ofdm_jrc/matrix_transpose.h>
#include <mimo_ofdm_jrc/mimo_ofdm_radar.h>
#include <mimo_ofdm_jrc/ofdm_cyclic_prefix_remover.h>
#include <mimo_ofdm_jrc/range_angle_estimator.h>
#include <mimo_ofdm_jrc/target_simulator.h>
#include <mimo_ofdm_jrc/zero_pad.h>

#include "jrc_oracle.h"

using namespace gr;
using namespace gr::mimo_ofdm_jrc;
typedef std::vector<gr_complex> cvec;

// ---------------------------------------------------------------------------------------------
// the reference block's own capture_radar_data() dump (lib/mimo_ofdm_radar_impl.cc:348-377) of one frame:
// golden text for the drop-in block's CSV format (tests/golden/make_golden.py)
// ---------------------------------------------------------------------------------------------
extern "C" __attribute__((visibility("default")))
void ref_capture(const orc_c32 *tx, const orc_c32 *rx, int N, int T, int R, int S, int pre, int interleave, const char *path)
{
    const int V = T * R, items = pre + S;
    auto radar = mimo_ofdm_radar::make(N, T, R, S, pre, false, false, 1, 1, interleave, path);
    cvec pad((size_t)V * N);
    std::vector<shim::input_t> in(T + R);
    for (int t = 0; t < T; t++) { in[t].items = tx + (size_t)t * items * N; in[t].n_items = items; }
    for (int r = 0; r < R; r++) { in[T + r].items = rx + (size_t)r * items * N; in[T + r].n_items = items; }
    in[0].tags.push_back(shim::make_tag(0, "packet_len", pmt::from_long(items)));
    in[T].tags.push_back(shim::make_tag(0, "packet_len", pmt::from_long(items)));
    shim::run_once(*radar, in, {{pad.data(), V}});
    radar->capture_radar_data(true);
}

// ---------------------------------------------------------------------------------------------
// reference chain for one batch (C ABI, same arguments as orc_chain_batch)
// ---------------------------------------------------------------------------------------------
static int find_bin(const float *bins, int n, float v)
{
    for (int i = 0; i < n; i++) if (bins[i] == v) return i;
    return -1;
}

extern "C" __attribute__((visibility("default")))
void ref_chain_batch(const orc_chain_cfg *cfg, const orc_c32 *rx, const orc_c32 *tx, int tx_shared, int n_cpi, int cpi0,
                     float *map_out, orc_c32 *cmap_out, orc_det *dets)
{
    const int N = cfg->fft_len, T = cfg->n_tx, R = cfg->n_rx, V = T * R, S = cfg->n_sym, pre = cfg->n_pre;
    const int Nr = N * cfg->interp_range, Na = V * cfg->interp_angle, items = pre + S;
    auto radar = mimo_ofdm_radar::make(N, T, R, S, pre, false, false, 1, cfg->interp_range, cfg->tx_interleave, "/dev/null");
    auto transp = matrix_transpose::make(Nr, V, cfg->interp_angle, false);
    std::vector<float> rb(cfg->range_bins, cfg->range_bins + Nr), ab(cfg->angle_bins, cfg->angle_bins + Na);
    auto estim = range_angle_estimator::make(Na, rb, ab, cfg->noise_discard_range_m, cfg->noise_discard_angle_deg,
                                             cfg->snr_threshold, cfg->power_threshold, "/dev/null", false);
    cvec pad((size_t)V * Nr), y((size_t)V * Nr), tr((size_t)Nr * Na), cm((size_t)Nr * Na);
    uint64_t rd = 0, rd2 = 0, rd3 = 0;
    const size_t frame = (size_t)items * N;
    for (int c = 0; c < n_cpi; c++) {
        std::vector<shim::input_t> in(T + R);
        for (int t = 0; t < T; t++) { in[t].items = tx + ((tx_shared ? 0 : (size_t)c * T) + t) * frame; in[t].n_items = items; }
        for (int r = 0; r < R; r++) { in[T + r].items = rx + ((size_t)c * R + r) * frame; in[T + r].n_items = items; }
        in[0].tags.push_back(shim::make_tag(rd, "packet_len", pmt::from_long(items)));
        in[T].tags.push_back(shim::make_tag(rd, "packet_len", pmt::from_long(items)));
        rd += items;
        shim::run_once(*radar, in, {{pad.data(), V}});
        orc_fft_vcc_batch((const orc_c32 *)pad.data(), (orc_c32 *)y.data(), Nr, V, 0, 0);       // fft_vxx #A
        shim::input_t ti; ti.items = y.data(); ti.n_items = V; ti.tags.push_back(shim::make_tag(rd2, "packet_len", pmt::from_long(V)));
        rd2 += V;
        shim::run_once(*transp, {ti}, {{tr.data(), Nr}});
        orc_c32 *cmp = cmap_out ? cmap_out + (size_t)c * Nr * Na : (orc_c32 *)cm.data();
        orc_fft_vcc_batch((const orc_c32 *)tr.data(), cmp, Na, Nr, 1, 1);                       // fft_vxx #B
        if (map_out) orc_mag_squared(cmp, map_out + (size_t)c * Nr * Na, (size_t)Nr * Na);
        if (dets) {
            shim::input_t ei; ei.items = cmp; ei.n_items = Nr; ei.tags.push_back(shim::make_tag(rd3, "packet_len", pmt::from_long(Nr)));
            rd3 += Nr;
            auto &msgs = estim->shim_published["params"];
            size_t before = msgs.size();
            shim::run_once(*estim, {ei}, {});
            orc_det d;
            std::memset(&d, 0, sizeof(d));
            d.range_idx = d.angle_idx = -1; d.cpi = cpi0 + c;
            if (msgs.size() > before) {
                auto m = msgs.back();
                auto val = [&](int k) { return pmt::f32vector_elements(pmt::nth(1, pmt::nth(k, m)))[0]; };
                d.range_idx = find_bin(rb.data(), Nr, val(0)); d.angle_idx = find_bin(ab.data(), Na, val(1));
                d.peak_power = val(2); d.snr_db = val(3); d.flags = 1;
                msgs.clear();
            }
            dets[c] = d;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pinning tests
// ---------------------------------------------------------------------------------------------
#ifdef REF_HARNESS_MAIN
static int g_fail = 0, g_checks = 0;
#define CHECK(cond, ...)                                                                            \
    do {                                                                                            \
        g_checks++;                                                                                 \
        if (!(cond)) { g_fail++; std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); } \
    } while (0)

static std::mt19937 rng(777);
static cvec randvec(size_t n, float scale = 1.f)
{
    std::normal_distribution<float> d(0.f, scale);
    cvec v(n);
    for (auto &z : v) z = gr_complex(d(rng), d(rng));
    return v;
}
static bool same(const void *a, const void *b, size_t bytes) { return std::memcmp(a, b, bytes) == 0; }
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

static void pin_radar()
{
    struct cfg_t { int N, T, R, S, pre, IR, rec; bool rem, recd, il; };
    const cfg_t cfgs[] = {{64, 4, 2, 4, 5, 8, 8, false, false, false}, {64, 4, 2, 4, 5, 8, 3, true, true, false},
                          {64, 4, 2, 4, 5, 16, 4, true, true, true},   {64, 2, 4, 2, 5, 16, 2, true, false, false},
                          {256, 4, 8, 4, 5, 4, 3, true, true, false},  {32, 2, 2, 3, 1, 2, 1, true, true, true}};
    for (const auto &c : cfgs) {
        const int V = c.T * c.R, items = c.pre + c.S + 2;
        auto blk = mimo_ofdm_radar::make(c.N, c.T, c.R, c.S, c.pre, c.rem, c.recd, c.rec, c.IR, c.il, "/tmp/jrc_ref_chan.csv");
        orc_radar *orc = orc_radar_create(c.N, c.T, c.R, c.S, c.pre, c.rem, c.recd, c.rec, c.IR, c.il);
        cvec out((size_t)V * c.N * c.IR), oout(out.size());
        uint64_t rd = 0;
        for (int it = 0; it < 12; it++) {
            std::vector<cvec> tx, rx;
            for (int t = 0; t < c.T; t++) tx.push_back(randvec((size_t)items * c.N));
            for (int r = 0; r < c.R; r++) rx.push_back(randvec((size_t)items * c.N));
            std::vector<shim::input_t> in(c.T + c.R);
            for (int t = 0; t < c.T; t++) { in[t].items = tx[t].data(); in[t].n_items = items; }
            for (int r = 0; r < c.R; r++) { in[c.T + r].items = rx[r].data(); in[c.T + r].n_items = items; }
            in[0].tags.push_back(shim::make_tag(rd, "packet_len", pmt::from_long(items)));
            in[c.T].tags.push_back(shim::make_tag(rd, "packet_len", pmt::from_long(items)));
            rd += items;
            auto res = shim::run_once(*blk, in, {{out.data(), V}});
            std::vector<const orc_c32 *> tp, rp;
            for (auto &v : tx) tp.push_back((const orc_c32 *)v.data());
            for (auto &v : rx) rp.push_back((const orc_c32 *)v.data());
            int n = orc_radar_work(orc, tp.data(), rp.data(), 0, (orc_c32 *)oout.data());
            CHECK(res.produced == n && n == V, "radar produced %d vs %d", res.produced, n);
            CHECK(same(out.data(), oout.data(), out.size() * sizeof(gr_complex)), "radar N=%d T=%d R=%d frame %d differs", c.N, c.T, c.R, it);
            CHECK(res.consumed[0] == items && res.consumed[c.T] == items, "radar consumed");
            CHECK(res.out_tags[0].size() == 1 && pmt::to_long(res.out_tags[0][0].value) == V && res.out_tags[0][0].offset == (uint64_t)it * V, "radar tag");
            if (it == 6) { blk->set_background_record(!c.recd); orc_radar_set_background_record(orc, !c.recd); }
        }
        // stale TX frame in front of the matching one (lib/mimo_ofdm_radar_impl.cc:189-197,260)
        std::vector<cvec> tx, rx;
        for (int t = 0; t < c.T; t++) tx.push_back(randvec((size_t)2 * items * c.N));
        for (int r = 0; r < c.R; r++) rx.push_back(randvec((size_t)items * c.N));
        std::vector<shim::input_t> in(c.T + c.R);
        for (int t = 0; t < c.T; t++) { in[t].items = tx[t].data(); in[t].n_items = 2 * items; }
        for (int r = 0; r < c.R; r++) { in[c.T + r].items = rx[r].data(); in[c.T + r].n_items = items; }
        in[0].tags = {shim::make_tag(rd, "packet_len", pmt::from_long(items)), shim::make_tag(rd + items, "packet_len", pmt::from_long(items))};
        in[c.T].tags = {shim::make_tag(rd, "packet_len", pmt::from_long(items))};
        auto res = shim::run_once(*blk, in, {{out.data(), V}});
        std::vector<const orc_c32 *> tp, rp;
        for (auto &v : tx) tp.push_back((const orc_c32 *)v.data());
        for (auto &v : rx) rp.push_back((const orc_c32 *)v.data());
        orc_radar_work(orc, tp.data(), rp.data(), (size_t)items, (orc_c32 *)oout.data());
        CHECK(same(out.data(), oout.data(), out.size() * sizeof(gr_complex)), "radar stale-TX skip differs");
        CHECK(res.consumed[0] == 2 * items && res.consumed[c.T] == items, "radar stale-TX consumed %d %d", res.consumed[0], res.consumed[c.T]);
        orc_radar_destroy(orc);
    }
}

static void pin_transpose()
{
    const int shapes[][3] = {{512, 8, 16}, {1024, 8, 8}, {1024, 32, 2}, {64, 4, 4}, {2048, 128, 1}};
    for (auto &s : shapes) {
        const int L = s[0], K = s[1], I = s[2];
        auto blk = matrix_transpose::make(L, K, I, false);
        cvec x = randvec((size_t)K * L), out((size_t)L * K * I, gr_complex(9, 9)), oout(out.size());
        shim::input_t in; in.items = x.data(); in.n_items = K; in.tags.push_back(shim::make_tag(0, "packet_len", pmt::from_long(K)));
        auto res = shim::run_once(*blk, {in}, {{out.data(), L}});
        orc_matrix_transpose((const orc_c32 *)x.data(), K, L, K, I, (orc_c32 *)oout.data());
        CHECK(res.produced == L && res.consumed[0] == K, "transpose produced %d", res.produced);
        CHECK(same(out.data(), oout.data(), out.size() * sizeof(gr_complex)), "transpose %dx%d differs", K, L);
        CHECK(!res.out_tags[0].empty() && pmt::to_long(res.out_tags[0].back().value) == L, "transpose tag");
        blk->shim_output_fullness = 0.01f;
        in.tags[0].offset = K;
        res = shim::run_once(*blk, {in}, {{out.data(), L}});
        CHECK(res.produced == 0 && res.consumed[0] == K, "transpose back-pressure");
    }
}

static void pin_estimator()
{
    const int confs[][3] = {{64, 8, 16}, {64, 16, 8}, {256, 4, 8}};   // Nsc, IR, IA with V = 8 (or 32 for the last)
    for (auto &cf : confs) {
        const int V = cf[0] == 256 ? 32 : 8, Nr = cf[0] * cf[1], Na = V * cf[2];
        auto rb = range_bins(cf[0], cf[1]); auto ab = angle_bins(Na);
        const float ndr = 2 * 1.2f, nda = 2 * (float)(std::asin(2.0 / V) * 180.0 / M_PI);
        for (int trial = 0; trial < 40; trial++) {
            const float snr_thr = trial % 5 == 0 ? 60.f : 15.f, pow_thr = trial % 7 == 0 ? 30.f : 0.f;
            auto blk = range_angle_estimator::make(Na, rb, ab, ndr, nda, snr_thr, pow_thr, "/tmp/jrc_ref_log.csv", trial == 3);
            cvec m = randvec((size_t)Nr * Na, 0.05f);
            std::uniform_int_distribution<int> ur(0, Nr - 1), ua(0, Na - 1);
            int pr = ur(rng), pa = ua(rng);
            if (trial < 8) { pr = (trial & 1) ? Nr - 1 : 0; pa = (trial & 2) ? Na - 1 : ((trial & 4) ? Na / 2 : 0); }
            m[(size_t)pr * Na + pa] = gr_complex(4.f + 0.1f * trial, -3.f);
            if (trial % 3 == 0) m[(size_t)((pr + 5) % Nr) * Na + (pa + 3) % Na] = m[(size_t)pr * Na + pa];   // exact tie
            shim::input_t in; in.items = m.data(); in.n_items = Nr; in.tags.push_back(shim::make_tag(0, "packet_len", pmt::from_long(Nr)));
            auto res = shim::run_once(*blk, {in}, {});
            orc_det od;
            orc_range_angle_estimate((const orc_c32 *)m.data(), Nr, Na, rb.data(), Nr, ab.data(), Na, ndr, nda, snr_thr, pow_thr, &od, nullptr);
            auto &msgs = blk->shim_published["params"];
            CHECK(res.produced == 0 && res.consumed[0] == Nr, "estimator consume");
            CHECK((msgs.size() == 1) == (od.flags == 1), "estimator gate: ref %zu oracle %u (trial %d)", msgs.size(), od.flags, trial);
            if (msgs.size() == 1) {
                auto val = [&](int k) { return pmt::f32vector_elements(pmt::nth(1, pmt::nth(k, msgs[0])))[0]; };
                CHECK(val(0) == rb[od.range_idx] && val(1) == ab[od.angle_idx], "estimator peak bin (trial %d)", trial);
                CHECK(val(2) == od.peak_power, "estimator peak power %.9g vs %.9g", val(2), od.peak_power);
                CHECK(val(3) == od.snr_db, "estimator snr %.9g vs %.9g (trial %d)", val(3), od.snr_db, trial);
                CHECK(pmt::symbol_to_string(pmt::nth(0, pmt::nth(3, msgs[0]))) == "snr", "message keys");
            }
        }
    }
}

static void pin_peak_pad_cp()
{
    for (int trial = 0; trial < 20; trial++) {
        const int n = trial < 2 ? 40000 : 64 + 37 * trial, protect = trial % 4 == 0 ? 0 : 5;
        const float thr = trial % 3 == 0 ? 25.f : 3.f;
        cvec x = randvec(n);
        if (trial % 2) x[n / 3] = x[2 * n / 3] = gr_complex(30, -7);
        auto blk = fft_peak_detect::make(1000000, 8.0f, thr, protect, {0.f}, false, "packet_len");
        float f = 123.f, ph = 123.f, mg = 123.f;
        shim::input_t in; in.items = x.data(); in.n_items = n; in.tags.push_back(shim::make_tag(0, "packet_len", pmt::from_long(n)));
        auto res = shim::run_once(*blk, {in}, {{&f, 1}, {&ph, 1}, {&mg, 1}});
        orc_peak1d o;
        orc_fft_peak_detect((const orc_c32 *)x.data(), n, 1000000, 8.0f, thr, protect, &o);
        CHECK(res.produced == 1, "peak produced");
        if (o.k >= 0) CHECK(f == o.freq && ph == o.phase && mg == o.mag, "peak trial %d: %g %g %g vs %g %g %g", trial, f, ph, mg, o.freq, o.phase, o.mag);
        else CHECK(f == 123.f && ph == 123.f && mg == 123.f, "peak: no detection must leave outputs unwritten");
    }
    auto zp = zero_pad::make(false, 7, 240);
    cvec x = randvec(720), y(967);
    shim::input_t zi; zi.items = x.data(); zi.n_items = 720; zi.tags.push_back(shim::make_tag(0, "packet_len", pmt::from_long(720)));
    auto res = shim::run_once(*zp, {zi}, {{y.data(), 967}});
    cvec oy(967);
    orc_zero_pad((const orc_c32 *)x.data(), 720, 7, 240, 1, (orc_c32 *)oy.data());
    CHECK(res.produced == 967 && same(&y[7], &oy[7], 720 * sizeof(gr_complex)), "zero_pad payload");
    double s2 = 0, o2 = 0;
    for (int i = 727; i < 967; i++) { s2 += std::norm(y[i]); o2 += std::norm(oy[i]); }
    CHECK(std::fabs(std::sqrt(s2 / 480) - 1e-2) < 2e-3 && std::fabs(std::sqrt(o2 / 480) - 1e-2) < 2e-3, "zero_pad sigma %g %g", std::sqrt(s2 / 480), std::sqrt(o2 / 480));
    auto cp = ofdm_cyclic_prefix_remover::make(64, 16, "packet_len");
    cvec t = randvec(9 * 80), u(9 * 64), ou(9 * 64);
    shim::input_t ci; ci.items = t.data(); ci.n_items = 720; ci.tags.push_back(shim::make_tag(0, "packet_len", pmt::from_long(720)));
    res = shim::run_once(*cp, {ci}, {{u.data(), 9}});
    orc_cp_remove((const orc_c32 *)t.data(), 9, 64, 16, (orc_c32 *)ou.data());
    CHECK(res.produced == 9 && same(u.data(), ou.data(), u.size() * sizeof(gr_complex)), "cp remover");
}

static void pin_target_simulator()
{
    const float lam = 3e8f / 24e9f;
    for (int trial = 0; trial < 6; trial++) {
        const int n = trial < 4 ? 960 : 512;
        std::vector<float> rg = {10.f + 7 * trial, 33.f}, vel = {0.f, trial * 3.f}, rcs = {10.f, 3.f}, az = {-30.f + 15 * trial, 20.f};
        if (trial % 2 == 0) { rg.resize(1); vel.resize(1); rcs.resize(1); az.resize(1); }
        std::vector<float> pos = {1 * lam, 3 * lam};
        auto blk = target_simulator::make(rg, vel, rcs, az, pos, 125000000, 24e9f, -20.f, false, trial == 3, "packet_len", false);
        cvec x = randvec(n, 0.3f), o0(n), o1(n), oo((size_t)2 * n);
        shim::input_t in; in.items = x.data(); in.n_items = n; in.tags.push_back(shim::make_tag(0, "packet_len", pmt::from_long(n)));
        auto res = shim::run_once(*blk, {in}, {{o0.data(), n}, {o1.data(), n}});
        orc_target_simulator((const orc_c32 *)x.data(), n, rg.data(), vel.data(), rcs.data(), az.data(), (int)rg.size(), pos.data(), 2,
                             125000000, 24e9f, trial == 3, -20.f, 0, (orc_c32 *)oo.data());
        CHECK(res.produced == n, "simulator produced");
        CHECK(same(o0.data(), &oo[0], n * sizeof(gr_complex)) && same(o1.data(), &oo[n], n * sizeof(gr_complex)), "target_simulator trial %d differs", trial);
        CHECK(res.out_tags[0].size() >= 1 && pmt::symbol_to_string(res.out_tags[0][0].key) == "rx_time", "simulator rx_time tag");
    }
}

static void pin_chain()
{
    // ref_chain_batch (reference blocks) vs orc_chain_batch on a point-target-like batch
    const int N = 64, T = 4, R = 2, S = 4, IR = 8, IA = 16, V = 8, Nr = N * IR, Na = V * IA, n = 6;
    auto rb = range_bins(N, IR); auto ab = angle_bins(Na);
    orc_chain_cfg cfg{N, T, R, S, 0, IR, IA, 0, rb.data(), ab.data(), 2.4f, 28.955f, 15.f, 0.f};
    cvec tx = randvec((size_t)T * S * N), rx((size_t)n * R * S * N);
    for (int c = 0; c < n; c++)
        for (int r = 0; r < R; r++)
            for (int s = 0; s < S; s++)
                for (int k = 0; k < N; k++) {
                    gr_complex acc = 0;
                    for (int t = 0; t < T; t++)
                        acc += tx[((size_t)t * S + s) * N + k] * std::polar(1.0f, (float)(-2 * M_PI * 0.07 * (c + 1) * k + 0.6 * (t + T * r)));
                    rx[(((size_t)c * R + r) * S + s) * N + k] = acc + randvec(1, 0.3f)[0];
                }
    std::vector<float> m1((size_t)n * Nr * Na), m2(m1.size());
    std::vector<orc_det> d1(n), d2(n);
    ref_chain_batch(&cfg, (const orc_c32 *)rx.data(), (const orc_c32 *)tx.data(), 1, n, 0, m1.data(), nullptr, d1.data());
    orc_chain_batch(&cfg, (const orc_c32 *)rx.data(), (const orc_c32 *)tx.data(), 1, n, 0, m2.data(), nullptr, d2.data());
    CHECK(same(m1.data(), m2.data(), m1.size() * sizeof(float)), "chain maps differ");
    for (int c = 0; c < n; c++)
        CHECK(d1[c].range_idx == d2[c].range_idx && d1[c].angle_idx == d2[c].angle_idx && d1[c].peak_power == d2[c].peak_power &&
                  d1[c].snr_db == d2[c].snr_db && d1[c].flags == d2[c].flags, "chain detection %d differs", c);
}

int main()
{
    pin_radar();
    pin_transpose();
    pin_estimator();
    pin_peak_pad_cp();
    pin_target_simulator();
    pin_chain();
    std::printf("%d checks, %d failed\n", g_checks, g_fail);
    if (g_fail) return 1;
    std::printf("ORACLE PINNED AGAINST REFERENCE SOURCES\n");
    return 0;
}
#endif
