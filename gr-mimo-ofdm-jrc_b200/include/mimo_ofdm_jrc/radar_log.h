// radar_log.h -- the consumer side of range_angle_estimator's CSV log.
// The estimator appends "HH:MM:SS.mmm, \t<power>, \t<snr>, \t<range>, \t<angle>\n"
// (lib/range_angle_estimator_impl.cc:264-271); mimo_precoder::compute_radar_aided_steering
// (lib/mimo_precoder_impl.cc:903-983) reads the LAST line, takes the 5th comma-separated field as the
// angle estimate in degrees and steers the TX array towards it.  Host-only helpers, header-only.
#pragma once
#include <cmath>
#include <complex>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace gr {
namespace mimo_ofdm_jrc {

struct radar_log_entry {
    std::string time;
    float power = 0.f, snr = 0.f, range = 0.f, angle = 0.f;
};

// Last record of the log.  Returns false when the file cannot be opened, is empty, or its last line is not
// a record (e.g. the "NEW RECORD" header) -- the precoder then falls back to its channel-estimate file.
inline bool radar_log_read_last(const std::string &path, radar_log_entry &out)
{
    std::ifstream f(path);
    if (!f.is_open()) return false;
    std::string line, last;
    while (std::getline(f, line))
        if (!line.empty()) last = line;
    if (last.empty()) return false;
    std::stringstream ls(last);
    std::string field[5];
    for (int i = 0; i < 4; i++)
        if (!std::getline(ls, field[i], ',')) return false;
    if (!std::getline(ls, field[4])) return false;
    try {
        out.time = field[0];
        out.power = std::stof(field[1]);
        out.snr = std::stof(field[2]);
        out.range = std::stof(field[3]);
        out.angle = std::stof(field[4]);
    } catch (...) {
        return false;
    }
    return true;
}

// Steering vector of the radar-aided precoder for an N_tx half-wavelength array
// (lib/mimo_precoder_impl.cc:952-956): a[i] = exp(j*pi*sin(angle)*i).
inline std::vector<std::complex<float>> radar_aided_steering_vector(float angle_deg, int n_tx)
{
    std::vector<std::complex<float>> a((size_t)n_tx);
    for (int i = 0; i < n_tx; i++)
        a[(size_t)i] = std::exp(std::complex<float>(0.f, (float)(M_PI * std::sin(angle_deg / 180.0 * M_PI) * i)));
    return a;
}

}  // namespace mimo_ofdm_jrc
}  // namespace gr
