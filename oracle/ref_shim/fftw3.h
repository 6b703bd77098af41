/* fftw3.h stand-in: lib/target_simulator_impl.h includes it but only uses gr::fft::fft_complex */
