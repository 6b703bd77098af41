// Stand-in for the three VOLK entry points lib/target_simulator_impl.cc uses (generic kernels).
#ifndef JRC_REFSHIM_VOLK_H
#define JRC_REFSHIM_VOLK_H
#include <complex>
#include <cstdlib>
typedef std::complex<float> lv_32fc_t;
inline size_t volk_get_alignment() { return 64; }
inline void *volk_malloc(size_t size, size_t alignment)
{
    void *p = nullptr;
    if (posix_memalign(&p, alignment < sizeof(void *) ? sizeof(void *) : alignment, size ? size : alignment)) return nullptr;
    return p;
}
inline void volk_free(void *p) { free(p); }
// volk_32fc_x2_multiply_32fc_generic: c[i] = a[i] * b[i] with separately rounded products and sums
inline void volk_32fc_x2_multiply_32fc(lv_32fc_t *c, const lv_32fc_t *a, const lv_32fc_t *b, unsigned int n)
{
    for (unsigned int i = 0; i < n; i++) {
        const float ar = a[i].real(), ai = a[i].imag(), br = b[i].real(), bi = b[i].imag();
        c[i] = lv_32fc_t(ar * br - ai * bi, ar * bi + ai * br);
    }
}
#endif
