// Stand-in for gr::fft::fft_complex (GNU Radio 3.8 gr-fft -> FFTW3f), only for oracle/_ref.
// execute() is the oracle's fft_vcc arithmetic (float32 radix-2 for powers of two, float64 direct
// DFT otherwise), so the reference's target_simulator and the oracle's restatement see the same
// FFT and can be compared bit for bit; the FFT itself is the part of the chain the reference does
// not contain.
#ifndef JRC_REFSHIM_GR_FFT_H
#define JRC_REFSHIM_GR_FFT_H
#include <complex>
#include <vector>
#include "../../../jrc_oracle.h"
namespace gr {
namespace fft {
class fft_complex {
    int d_size; bool d_forward;
    std::vector<std::complex<float>> d_in, d_out;

public:
    fft_complex(int fft_size, bool forward = true, int /*nthreads*/ = 1) : d_size(fft_size), d_forward(forward), d_in(fft_size), d_out(fft_size) {}
    std::complex<float> *get_inbuf() { return d_in.data(); }
    std::complex<float> *get_outbuf() { return d_out.data(); }
    int inbuf_length() const { return d_size; }
    int outbuf_length() const { return d_size; }
    void execute() { orc_fft_vcc((const orc_c32 *)d_in.data(), (orc_c32 *)d_out.data(), d_size, d_forward ? 1 : 0, 0); }
};
inline void *malloc_complex(int size) { return new std::complex<float>[size]; }
inline void free(void *b) { delete[] (std::complex<float> *)b; }
}  // namespace fft
}  // namespace gr
#endif
