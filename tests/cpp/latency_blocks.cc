// BASELINE configs[3]: streaming latency mode -- one CPI per work() call of the fused radar_chain block, driven through
// the runtime stand-in exactly as the scheduler drives a tagged-stream block (length tags, message port):
//   * synchronous calls on pageable stream buffers: p50 / p99 of the per-call wall time;
//   * the submit/wait pipeline (set_pipeline_depth(4)) on a page-locked output ring that advances like the scheduler's
//     circular buffer: sustained CPI/s at one CPI per general_work() and p50 / p99 of the per-CPI latency, frame
//     handed to general_work() -> its packet and message published;
//   * the same CPI through the five separate drop-in blocks.
// Prints one JSON line.
//   build/latency_blocks [n_calls]
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <string>
#include <vector>

#include <jrc_cuda.h>
#include <mimo_ofdm_jrc/matrix_transpose.h>
#include <mimo_ofdm_jrc/mimo_ofdm_radar.h>
#include <mimo_ofdm_jrc/radar_chain.h>
#include <mimo_ofdm_jrc/range_angle_estimator.h>

using namespace gr;
using namespace gr::mimo_ofdm_jrc;
typedef std::vector<gr_complex> cvec;

int main(int argc, char **argv)
{
    const int n_calls = argc > 1 ? std::atoi(argv[1]) : 10000;
    std::mt19937 rng(1);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::string json = "{";       // printed last, on a line of its own (the blocks log to stdout)
    char line[1024];
    const int cfgs[2][2] = {{8, 16}, {16, 8}};     // shipped 512 x 128, configs[1] 1024 x 64
    for (int ci = 0; ci < 2; ci++) {
        const int N = 64, T = 4, R = 2, S = 4, pre = 5, IR = cfgs[ci][0], IA = cfgs[ci][1], V = 8, Nr = N * IR, Na = V * IA, items = pre + S;
        std::vector<float> rb(Nr), ab(Na);
        for (int i = 0; i < Nr; i++) rb[i] = 76.8f * i / (Nr - 1);
        for (int i = 0; i < Na; i++) ab[i] = (float)(std::asin(2.0 * (i - Na / 2) / Na) * 180.0 / M_PI);
        auto blk = radar_chain::make(N, T, R, S, pre, false, false, 8, IR, IA, false, rb, ab, 2.4f, 28.955f, 15.f, 0.f, "/tmp/jrc_lat_log.csv", false);
        std::vector<cvec> tx(T, cvec((size_t)items * N)), rx(R, cvec((size_t)items * N));
        for (auto &v : tx) for (auto &z : v) z = gr_complex(nd(rng) > 0 ? 1.f : -1.f, 0.f);
        for (int r = 0; r < R; r++)
            for (int s = 0; s < items; s++)
                for (int k = 0; k < N; k++) {
                    gr_complex acc = 0;
                    for (int t = 0; t < T; t++) acc += tx[t][(size_t)s * N + k] * std::polar(1.0f, (float)(-2 * M_PI * 0.13 * k + 0.9 * (t + T * r)));
                    rx[r][(size_t)s * N + k] = acc + gr_complex(0.05f * nd(rng), 0.05f * nd(rng));
                }
        std::vector<float> map((size_t)Nr * Na);
        std::vector<double> us;
        us.reserve(n_calls);
        uint64_t rd = 0;
        for (int it = 0; it < n_calls + 200; it++) {
            std::vector<shim::input_t> in(T + R);
            for (int t = 0; t < T; t++) { in[t].items = tx[t].data(); in[t].n_items = items; }
            for (int r = 0; r < R; r++) { in[T + r].items = rx[r].data(); in[T + r].n_items = items; }
            in[0].tags.push_back(shim::make_tag(rd, "packet_len", pmt::from_long(items)));
            in[T].tags.push_back(shim::make_tag(rd, "packet_len", pmt::from_long(items)));
            rd += items;
            auto t0 = std::chrono::steady_clock::now();
            auto res = shim::run_once(*blk, in, {{map.data(), Nr}});
            auto t1 = std::chrono::steady_clock::now();
            if (res.produced != Nr) { std::fprintf(stderr, "radar_chain produced %d items\n", res.produced); return 1; }
            if (it >= 200) us.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
            blk->shim_published["params"].clear();
        }
        // the same CPI through the five separate blocks of the shipped flowgraph (every intermediate crosses PCIe twice)
        std::vector<double> us5;
        {
            auto radar = mimo_ofdm_radar::make(N, T, R, S, pre, false, false, 8, IR, false, "/tmp/jrc_lat_chan.csv");
            auto transp = matrix_transpose::make(Nr, V, IA, false);
            auto estim = range_angle_estimator::make(Na, rb, ab, 2.4f, 28.955f, 15.f, 0.f, "/tmp/jrc_lat_log.csv", false);
            jrc_chain_cfg ucfg{}; ucfg.fft_len = 64; ucfg.n_tx = ucfg.n_rx = ucfg.n_sym = 1; ucfg.interp_range = ucfg.interp_angle = 1;
            jrc_chain *util = nullptr;
            if (jrc_chain_create(&ucfg, &util) != JRC_OK) { std::fprintf(stderr, "%s\n", jrc_last_error()); return 1; }
            cvec pad((size_t)V * Nr), y((size_t)V * Nr), tr((size_t)Nr * Na), cm((size_t)Nr * Na);
            uint64_t r1 = 0, r2 = 0, r3 = 0;
            const int n5 = n_calls / 10 + 20;
            for (int it = 0; it < n5; it++) {
                std::vector<shim::input_t> in(T + R);
                for (int t = 0; t < T; t++) { in[t].items = tx[t].data(); in[t].n_items = items; }
                for (int r = 0; r < R; r++) { in[T + r].items = rx[r].data(); in[T + r].n_items = items; }
                in[0].tags.push_back(shim::make_tag(r1, "packet_len", pmt::from_long(items)));
                in[T].tags.push_back(shim::make_tag(r1, "packet_len", pmt::from_long(items)));
                r1 += items;
                auto t0 = std::chrono::steady_clock::now();
                shim::run_once(*radar, in, {{pad.data(), 64}});
                jrc_fft_vcc(util, (const jrc_c32 *)pad.data(), (jrc_c32 *)y.data(), Nr, V, 0, 0);
                shim::input_t ti; ti.items = y.data(); ti.n_items = V; ti.tags.push_back(shim::make_tag(r2, "packet_len", pmt::from_long(V)));
                r2 += V;
                shim::run_once(*transp, {ti}, {{tr.data(), Nr}});
                jrc_fft_vcc(util, (const jrc_c32 *)tr.data(), (jrc_c32 *)cm.data(), Na, Nr, 1, 1);
                jrc_mag_squared(util, (const jrc_c32 *)cm.data(), map.data(), (size_t)Nr * Na);
                shim::input_t ei; ei.items = cm.data(); ei.n_items = Nr; ei.tags.push_back(shim::make_tag(r3, "packet_len", pmt::from_long(Nr)));
                r3 += Nr;
                shim::run_once(*estim, {ei}, {});
                auto t1 = std::chrono::steady_clock::now();
                if (it >= 20) us5.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
                estim->shim_published["params"].clear();
            }
            jrc_chain_destroy(util);
            std::sort(us5.begin(), us5.end());
        }
        std::vector<double> us3, us3_radar, us3_transp, us3_estim;
        {
            setenv("JRC_FUSED", "1", 1);
            auto radar = mimo_ofdm_radar::make(N, T, R, S, pre, false, false, 8, IR, false, "/tmp/jrc_lat_chan.csv");
            auto transp = matrix_transpose::make(Nr, V, IA, false);
            auto estim = range_angle_estimator::make(Na, rb, ab, 2.4f, 28.955f, 15.f, 0.f, "/tmp/jrc_lat_log.csv", false);
            unsetenv("JRC_FUSED");
            cvec pad((size_t)V * Nr), tr((size_t)Nr * Na);
            uint64_t r1 = 0, r2 = 0, r3 = 0;
            const int n3 = n_calls / 4 + 50;
            for (int it = 0; it < n3; it++) {
                std::vector<shim::input_t> in(T + R);
                for (int t = 0; t < T; t++) { in[t].items = tx[t].data(); in[t].n_items = items; }
                for (int r = 0; r < R; r++) { in[T + r].items = rx[r].data(); in[T + r].n_items = items; }
                in[0].tags.push_back(shim::make_tag(r1, "packet_len", pmt::from_long(items)));
                in[T].tags.push_back(shim::make_tag(r1, "packet_len", pmt::from_long(items)));
                r1 += items;
                auto t0 = std::chrono::steady_clock::now();
                auto o1 = shim::run_once(*radar, in, {{pad.data(), 64}});
                // (fft_vcc #A: tags pass through unchanged)
                shim::input_t ti; ti.items = pad.data(); ti.n_items = V;
                for (auto tg : o1.out_tags[0]) { tg.offset = r2; ti.tags.push_back(tg); }
                r2 += V;
                auto ta = std::chrono::steady_clock::now();
                auto o2 = shim::run_once(*transp, {ti}, {{tr.data(), Nr}});
                auto tb = std::chrono::steady_clock::now();
                // (fft_vcc #B)
                shim::input_t ei; ei.items = tr.data(); ei.n_items = Nr;
                for (auto tg : o2.out_tags[0]) { tg.offset = r3; ei.tags.push_back(tg); }
                r3 += Nr;
                shim::run_once(*estim, {ei}, {});
                auto t1 = std::chrono::steady_clock::now();
                if (o1.out_tags[0].size() != 2 || o2.out_tags[0].size() != 2 || estim->shim_published["params"].size() != 1) {
                    std::fprintf(stderr, "JRC_FUSED: the downstream blocks did not serve the cached frame\n");
                    return 1;
                }
                if (it >= 50) {
                    us3.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
                    us3_radar.push_back(std::chrono::duration<double, std::micro>(ta - t0).count());
                    us3_transp.push_back(std::chrono::duration<double, std::micro>(tb - ta).count());
                    us3_estim.push_back(std::chrono::duration<double, std::micro>(t1 - tb).count());
                }
                estim->shim_published["params"].clear();
            }
            for (auto *v : {&us3, &us3_radar, &us3_transp, &us3_estim}) std::sort(v->begin(), v->end());
        }
        // ---- pipelined: up to 4 frames in flight, page-locked output ring ----
        std::vector<double> lat;
        double sustained = 0.0;
        {
            auto pb = radar_chain::make(N, T, R, S, pre, false, false, 8, IR, IA, false, rb, ab, 2.4f, 28.955f, 15.f, 0.f, "/tmp/jrc_lat_log.csv", false);
            pb->set_pipeline_depth(4);
            const int RING = 32;                        // packets
            void *ringp = nullptr;
            if (jrc_pinned_alloc((size_t)RING * Nr * Na * sizeof(float), &ringp) != JRC_OK) { std::fprintf(stderr, "%s\n", jrc_last_error()); return 1; }
            float *ring = static_cast<float *>(ringp);
            uint64_t rdp = 0, written = 0;
            std::vector<std::chrono::steady_clock::time_point> t_in;
            t_in.reserve(n_calls + 300);
            size_t emitted = 0;
            const int total = n_calls + 200;
            std::chrono::steady_clock::time_point t_start;
            int fed = 0;
            bool offered = false;
            while ((int)emitted < total) {
                std::vector<shim::input_t> in(T + R);
                const bool have = fed < total;
                for (int t = 0; t < T; t++) { in[t].items = tx[t].data(); in[t].n_items = have ? items : 0; }
                for (int r = 0; r < R; r++) { in[T + r].items = rx[r].data(); in[T + r].n_items = have ? items : 0; }
                if (have) {
                    in[0].tags.push_back(shim::make_tag(rdp, "packet_len", pmt::from_long(items)));
                    in[T].tags.push_back(shim::make_tag(rdp, "packet_len", pmt::from_long(items)));
                    if (!offered) { t_in.push_back(std::chrono::steady_clock::now()); offered = true; }
                }
                if (fed == 200 && emitted <= 200 && t_start == std::chrono::steady_clock::time_point()) t_start = std::chrono::steady_clock::now();
                const int pos = (int)(written % RING);
                const int room = std::min(4, RING - pos);              // the ring does not wrap inside one call
                auto res = shim::run_once(*pb, in, {{ring + (size_t)pos * Nr * Na, room * Nr}});
                if (have && res.consumed[T] > 0) { rdp += items; fed++; offered = false; }
                if (res.produced == Nr) {
                    auto now = std::chrono::steady_clock::now();
                    if (emitted >= 200) lat.push_back(std::chrono::duration<double, std::micro>(now - t_in[emitted]).count());
                    emitted++;
                    written++;
                } else if (res.produced != 0) { std::fprintf(stderr, "pipelined radar_chain produced %d items\n", res.produced); return 1; }
                pb->shim_published["params"].clear();
            }
            auto t_stop = std::chrono::steady_clock::now();
            sustained = (double)(total - 200) / std::chrono::duration<double>(t_stop - t_start).count();
            pb.reset();
            jrc_pinned_free(ringp);
            std::sort(lat.begin(), lat.end());
        }
        std::sort(us.begin(), us.end());
        std::snprintf(line, sizeof(line), "%s\"five separate blocks %dx%d, 1 CPI\": {\"calls\": %zu, \"p50_us\": %.1f, \"p99_us\": %.1f}, ", ci ? ", " : "", Nr, Na,
                    us5.size(), us5[us5.size() / 2], us5[(size_t)(us5.size() * 0.99)]);
        json += line;
        std::snprintf(line, sizeof(line), "\"five-block wiring %dx%d with JRC_FUSED=1, the three radar blocks, 1 CPI\": {\"calls\": %zu, \"p50_us\": %.1f, \"p99_us\": %.1f, "
                    "\"p50_us_mimo_ofdm_radar\": %.1f, \"p50_us_matrix_transpose\": %.1f, \"p50_us_range_angle_estimator\": %.1f}, ",
                    Nr, Na, us3.size(), us3[us3.size() / 2], us3[(size_t)(us3.size() * 0.99)], us3_radar[us3_radar.size() / 2],
                    us3_transp[us3_transp.size() / 2], us3_estim[us3_estim.size() / 2]);
        json += line;
        std::snprintf(line, sizeof(line), "\"radar_chain block %dx%d, 1 CPI per work(), pageable buffers\": {\"calls\": %d, \"p50_us\": %.2f, \"p99_us\": %.2f, \"mean_us\": %.2f}, ",
                    Nr, Na, n_calls, us[us.size() / 2], us[(size_t)(us.size() * 0.99)],
                    std::accumulate(us.begin(), us.end(), 0.0) / us.size());
        json += line;
        std::snprintf(line, sizeof(line), "\"radar_chain block %dx%d, 1 CPI per work(), pipeline depth 4, page-locked output ring\": {\"cpis\": %zu, "
                    "\"sustained_cpi_per_s\": %.0f, \"latency_p50_us\": %.2f, \"latency_p99_us\": %.2f}",
                    Nr, Na, lat.size(), sustained, lat[lat.size() / 2], lat[(size_t)(lat.size() * 0.99)]);
        json += line;
    }
    json += "}";
    std::printf("\n%s\n", json.c_str());
    return 0;
}
