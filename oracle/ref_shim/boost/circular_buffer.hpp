// Stand-in for boost::circular_buffer, only for building the REFERENCE sources into oracle/_ref
// (Boost is not installed here).  Semantics of the members lib/mimo_ofdm_radar_impl.cc uses:
// push_back overwrites the oldest element when full and is a no-op at capacity 0; operator[](0)
// is the oldest element.
#ifndef JRC_REFSHIM_CIRCULAR_BUFFER_HPP
#define JRC_REFSHIM_CIRCULAR_BUFFER_HPP
#include <cstddef>
#include <vector>
namespace boost {
template <class T>
class circular_buffer {
    std::vector<T> d_buf;
    size_t d_cap = 0, d_head = 0, d_size = 0;

public:
    circular_buffer() {}
    explicit circular_buffer(size_t cap) { set_capacity(cap); }
    void set_capacity(size_t cap) { d_buf.assign(cap, T()); d_cap = cap; d_head = 0; d_size = 0; }
    size_t capacity() const { return d_cap; }
    size_t size() const { return d_size; }
    bool empty() const { return d_size == 0; }
    bool full() const { return d_size == d_cap; }
    T &operator[](size_t i) { return d_buf[(d_head + i) % d_cap]; }
    const T &operator[](size_t i) const { return d_buf[(d_head + i) % d_cap]; }
    void push_back(const T &v)
    {
        if (d_cap == 0) return;
        if (d_size < d_cap) { d_buf[(d_head + d_size) % d_cap] = v; d_size++; }
        else { d_buf[d_head] = v; d_head = (d_head + 1) % d_cap; }
    }
    void clear() { d_head = 0; d_size = 0; }
};
}  // namespace boost
#endif
