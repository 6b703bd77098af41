// gnuradio/shim_runtime.h -- header-only stand-in for the slice of the GNU Radio 3.8 runtime that
// the radar-path blocks touch: gr::block, gr::tagged_stream_block, io_signature, stream tags,
// message ports, performance counters.  It exists because GNU Radio cannot be installed in the
// build image; the block sources in ../../lib compile unchanged against the real headers.
//
// It is NOT a scheduler: gr::shim::run_once() performs exactly one general_work() call on
// caller-provided buffers and records what the block consumed, produced, tagged and published,
// which is what the block-level parity tests need.  tagged_stream_block::general_work follows
// gnuradio-runtime/lib/tagged_stream_block.cc (maint-3.8): parse the length tag at the head of
// every input, size the output with calculate_output_stream_length(), call work(), consume the
// whole packet on every input (unless WORK_DONE), tag the output with the produced length.
#ifndef JRC_SHIM_GR_RUNTIME_H
#define JRC_SHIM_GR_RUNTIME_H

#include <algorithm>
#include <complex>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include <boost/shared_ptr.hpp>
#include <gnuradio/attributes.h>
#include <pmt/pmt.h>

#define GR_RUNTIME_API
#define GR_M_PI 3.14159265358979323846
#define GR_M_TWOPI (2 * GR_M_PI)

typedef std::complex<float> gr_complex;
typedef std::complex<double> gr_complexd;
typedef std::vector<int> gr_vector_int;
typedef std::vector<unsigned int> gr_vector_uint;
typedef std::vector<float> gr_vector_float;
typedef std::vector<void *> gr_vector_void_star;
typedef std::vector<const void *> gr_vector_const_void_star;

namespace gr {

namespace thread {
typedef std::mutex mutex;
typedef std::unique_lock<std::mutex> scoped_lock;
}  // namespace thread

struct tag_t {
    uint64_t offset = 0;
    pmt::pmt_t key, value, srcid;
    std::vector<long> marked_deleted;
    static bool offset_compare(const tag_t &x, const tag_t &y) { return x.offset < y.offset; }
};

class io_signature {
    int d_min, d_max;
    std::vector<int> d_sizeof;
    io_signature(int mn, int mx, const std::vector<int> &s) : d_min(mn), d_max(mx), d_sizeof(s) {}

public:
    typedef boost::shared_ptr<io_signature> sptr;
    static const int IO_INFINITE = -1;
    static sptr makev(int mn, int mx, const std::vector<int> &s)
    {
        if (mn < 0 || (mx != IO_INFINITE && mx < mn) || s.empty()) throw std::invalid_argument("gr::io_signature");
        return sptr(new io_signature(mn, mx, s));
    }
    static sptr make(int mn, int mx, int s) { return makev(mn, mx, std::vector<int>{s}); }
    static sptr make2(int mn, int mx, int s1, int s2) { return makev(mn, mx, std::vector<int>{s1, s2}); }
    static sptr make3(int mn, int mx, int s1, int s2, int s3) { return makev(mn, mx, std::vector<int>{s1, s2, s3}); }
    int min_streams() const { return d_min; }
    int max_streams() const { return d_max; }
    int sizeof_stream_item(int i) const { return d_sizeof[std::min<size_t>((size_t)std::max(i, 0), d_sizeof.size() - 1)]; }
    std::vector<int> sizeof_stream_items() const { return d_sizeof; }
};

class basic_block : public boost::enable_shared_from_this<basic_block> {
protected:
    std::string d_name, d_alias;
    io_signature::sptr d_input_signature, d_output_signature;
    basic_block(const std::string &name, io_signature::sptr in, io_signature::sptr out)
        : d_name(name), d_alias(name + "0"), d_input_signature(in), d_output_signature(out) {}

public:
    // what message_port_pub carried, per port name (read by the tests)
    std::map<std::string, std::vector<pmt::pmt_t>> shim_published;
    std::vector<std::string> shim_out_ports;

    virtual ~basic_block() {}
    std::string name() const { return d_name; }
    std::string alias() const { return d_alias; }
    void set_block_alias(const std::string &a) { d_alias = a; }
    io_signature::sptr input_signature() const { return d_input_signature; }
    io_signature::sptr output_signature() const { return d_output_signature; }
    void message_port_register_out(pmt::pmt_t port) { shim_out_ports.push_back(pmt::symbol_to_string(port)); }
    void message_port_register_in(pmt::pmt_t) {}
    void message_port_pub(pmt::pmt_t port, pmt::pmt_t msg) { shim_published[pmt::symbol_to_string(port)].push_back(msg); }
};

class block : public basic_block {
public:
    enum { WORK_CALLED_PRODUCE = -2, WORK_DONE = -1 };
    enum tag_propagation_policy_t { TPP_DONT = 0, TPP_ALL_TO_ALL = 1, TPP_ONE_TO_ONE = 2, TPP_CUSTOM = 3 };

    // ---- shim state, driven by gr::shim::run_once -------------------------------------
    struct shim_port {
        uint64_t n_items = 0;            // nitems_read / nitems_written before this call
        std::vector<tag_t> tags;         // inputs: tags on the stream; outputs: tags the block added
    };
    std::vector<shim_port> shim_in, shim_out;
    std::vector<int> shim_consumed;
    float shim_output_fullness = 0.0f;   // what pc_output_buffers_full() reports
    int shim_min_noutput_items = 1, shim_output_multiple = 1;
    double shim_relative_rate = 1.0;

protected:
    tag_propagation_policy_t d_tpp = TPP_ALL_TO_ALL;
    block(const std::string &name, io_signature::sptr in, io_signature::sptr out) : basic_block(name, in, out) {}

public:
    gr::thread::mutex d_setlock;
    typedef boost::shared_ptr<block> sptr;

    virtual void forecast(int noutput_items, gr_vector_int &required)
    {
        for (auto &r : required) r = noutput_items;
    }
    virtual int general_work(int, gr_vector_int &, gr_vector_const_void_star &, gr_vector_void_star &)
    {
        throw std::runtime_error("gr::block::general_work not overridden");
    }
    virtual bool start() { return true; }
    virtual bool stop() { return true; }

    void consume(int which, int how_many)
    {
        if ((size_t)which >= shim_consumed.size()) shim_consumed.resize(which + 1, 0);
        shim_consumed[which] += how_many;
    }
    void consume_each(int how_many) { for (auto &c : shim_consumed) c += how_many; }
    void produce(int, int) {}
    uint64_t nitems_read(unsigned which) { return which < shim_in.size() ? shim_in[which].n_items : 0; }
    uint64_t nitems_written(unsigned which) { return which < shim_out.size() ? shim_out[which].n_items : 0; }

    void add_item_tag(unsigned which, const tag_t &t)
    {
        if (which >= shim_out.size()) shim_out.resize(which + 1);
        shim_out[which].tags.push_back(t);
    }
    void add_item_tag(unsigned which, uint64_t abs_offset, const pmt::pmt_t &key, const pmt::pmt_t &value,
                      const pmt::pmt_t &srcid = pmt::get_PMT_F())
    {
        tag_t t; t.offset = abs_offset; t.key = key; t.value = value; t.srcid = srcid;
        add_item_tag(which, t);
    }
    void remove_item_tag(unsigned which, const tag_t &t)
    {
        if (which >= shim_in.size()) return;
        auto &v = shim_in[which].tags;
        for (size_t i = 0; i < v.size(); i++)
            if (v[i].offset == t.offset && pmt::eqv(v[i].key, t.key)) { v.erase(v.begin() + i); return; }
    }
    void get_tags_in_range(std::vector<tag_t> &v, unsigned which, uint64_t start, uint64_t end)
    {
        v.clear();
        if (which >= shim_in.size()) return;
        for (const auto &t : shim_in[which].tags)
            if (t.offset >= start && t.offset < end) v.push_back(t);
    }
    void get_tags_in_range(std::vector<tag_t> &v, unsigned which, uint64_t start, uint64_t end, const pmt::pmt_t &key)
    {
        v.clear();
        if (which >= shim_in.size()) return;
        for (const auto &t : shim_in[which].tags)
            if (t.offset >= start && t.offset < end && pmt::eqv(t.key, key)) v.push_back(t);
    }
    void get_tags_in_window(std::vector<tag_t> &v, unsigned which, uint64_t rel_start, uint64_t rel_end)
    {
        get_tags_in_range(v, which, nitems_read(which) + rel_start, nitems_read(which) + rel_end);
    }

    void set_tag_propagation_policy(tag_propagation_policy_t p) { d_tpp = p; }
    tag_propagation_policy_t tag_propagation_policy() const { return d_tpp; }
    void set_output_multiple(int m) { shim_output_multiple = m; }
    int output_multiple() const { return shim_output_multiple; }
    void set_relative_rate(double r) { shim_relative_rate = r; }
    double relative_rate() const { return shim_relative_rate; }
    void set_min_noutput_items(int m) { shim_min_noutput_items = m; }
    int min_noutput_items() const { return shim_min_noutput_items; }
    void set_max_noutput_items(int) {}
    void set_min_output_buffer(long) {}
    void set_min_output_buffer(int, long) {}
    void set_max_output_buffer(long) {}
    void set_alignment(int) {}
    void set_history(unsigned) {}

    // performance counters (the reference reads these ad hoc, lib/matrix_transpose_impl.cc:86,105-106)
    float pc_output_buffers_full(int) { return shim_output_fullness; }
    float pc_input_buffers_full(int) { return 0.0f; }
    std::vector<float> pc_output_buffers_full() { return std::vector<float>(1, shim_output_fullness); }
    float pc_work_time_total() { return 0.0f; }
    float pc_work_time() { return 0.0f; }
    float pc_nproduced() { return 0.0f; }
};

class sync_block : public block {
protected:
    sync_block(const std::string &name, io_signature::sptr in, io_signature::sptr out) : block(name, in, out) {}

public:
    virtual int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) = 0;
    int general_work(int noutput_items, gr_vector_int &, gr_vector_const_void_star &in, gr_vector_void_star &out) override
    {
        int r = work(noutput_items, in, out);
        if (r > 0) consume_each(r);
        return r;
    }
};

class tagged_stream_block : public block {
    std::string d_length_tag_key_str;
    gr_vector_int d_n_input_items_reqd;

protected:
    pmt::pmt_t d_length_tag_key;
    tagged_stream_block(const std::string &name, io_signature::sptr in, io_signature::sptr out,
                        const std::string &length_tag_key)
        : block(name, in, out), d_length_tag_key_str(length_tag_key), d_n_input_items_reqd(1, 0),
          d_length_tag_key(pmt::string_to_symbol(length_tag_key)) {}

    virtual void parse_length_tags(const std::vector<std::vector<tag_t>> &tags, gr_vector_int &n_input_items_reqd)
    {
        for (unsigned i = 0; i < tags.size(); i++)
            for (unsigned k = 0; k < tags[i].size(); k++)
                if (pmt::eqv(tags[i][k].key, d_length_tag_key)) {
                    n_input_items_reqd[i] = (int)pmt::to_long(tags[i][k].value);
                    remove_item_tag(i, tags[i][k]);
                }
    }
    virtual int calculate_output_stream_length(const gr_vector_int &ninput_items)
    {
        int n = *std::max_element(ninput_items.begin(), ninput_items.end());
        return (int)(n * relative_rate());
    }
    virtual void update_length_tags(int n_produced, int n_ports)
    {
        for (int i = 0; i < n_ports; i++)
            add_item_tag(i, nitems_written(i), d_length_tag_key, pmt::from_long(n_produced));
    }

public:
    void forecast(int, gr_vector_int &required) override
    {
        for (unsigned i = 0; i < required.size(); i++)
            required[i] = i < d_n_input_items_reqd.size() ? std::max(1, d_n_input_items_reqd[i]) : 1;
    }
    int general_work(int noutput_items, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items,
                     gr_vector_void_star &output_items) override
    {
        if (d_length_tag_key_str.empty()) return work(noutput_items, ninput_items, input_items, output_items);
        if (d_n_input_items_reqd.empty() || d_n_input_items_reqd[0] == 0) {
            std::vector<std::vector<tag_t>> tags(input_items.size());
            for (unsigned i = 0; i < input_items.size(); i++) get_tags_in_range(tags[i], i, nitems_read(i), nitems_read(i) + 1);
            d_n_input_items_reqd.assign(input_items.size(), -1);
            parse_length_tags(tags, d_n_input_items_reqd);
        }
        for (unsigned i = 0; i < input_items.size(); i++) {
            if (d_n_input_items_reqd[i] == -1) throw std::runtime_error("Missing a required length tag on port " + std::to_string(i));
            if (d_n_input_items_reqd[i] > ninput_items[i]) return 0;
        }
        int min_output_size = calculate_output_stream_length(d_n_input_items_reqd);
        if (noutput_items < min_output_size) { set_min_noutput_items(min_output_size); return 0; }
        set_min_noutput_items(1);
        int n_produced = work(noutput_items, d_n_input_items_reqd, input_items, output_items);
        if (n_produced == WORK_DONE) return n_produced;
        for (int i = 0; i < (int)d_n_input_items_reqd.size(); i++) consume(i, d_n_input_items_reqd[i]);
        if (n_produced > 0) update_length_tags(n_produced, (int)output_items.size());
        d_n_input_items_reqd.assign(std::max<size_t>(1, input_items.size()), 0);
        return n_produced;
    }
    virtual int work(int noutput_items, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items,
                     gr_vector_void_star &output_items) = 0;
};

}  // namespace gr

namespace gnuradio {
template <class T>
boost::shared_ptr<T> get_initial_sptr(T *p) { return boost::shared_ptr<T>(p); }
}  // namespace gnuradio

// -------------------------------------------------------------------------------------------
// test driver: one general_work() call
// -------------------------------------------------------------------------------------------
namespace gr {
namespace shim {

struct input_t { const void *items = nullptr; int n_items = 0; std::vector<tag_t> tags; };   // tag offsets absolute
struct output_t { void *items = nullptr; int capacity = 0; };
struct result_t {
    int produced = 0;
    std::vector<int> consumed;
    std::vector<std::vector<tag_t>> out_tags;
};

// Runs blk.general_work once.  nitems_read/nitems_written persist in the block between calls, and a
// tag's offset is relative to the stream, so packets can be fed back to back like the scheduler does.
inline result_t run_once(block &blk, const std::vector<input_t> &in, const std::vector<output_t> &out)
{
    blk.shim_in.resize(in.size());
    blk.shim_out.resize(out.size());
    gr_vector_int ninput(in.size());
    gr_vector_const_void_star in_ptrs(in.size());
    gr_vector_void_star out_ptrs(out.size());
    int noutput = out.empty() ? 0 : out[0].capacity;
    for (size_t i = 0; i < in.size(); i++) {
        ninput[i] = in[i].n_items; in_ptrs[i] = in[i].items;
        blk.shim_in[i].tags = in[i].tags;
    }
    for (size_t i = 0; i < out.size(); i++) {
        out_ptrs[i] = out[i].items; noutput = std::min(noutput, out[i].capacity);
        blk.shim_out[i].tags.clear();
    }
    blk.shim_consumed.assign(in.size(), 0);
    result_t r;
    r.produced = blk.general_work(noutput, ninput, in_ptrs, out_ptrs);
    r.consumed = blk.shim_consumed;
    for (size_t i = 0; i < in.size(); i++) blk.shim_in[i].n_items += (uint64_t)r.consumed[i];
    for (size_t i = 0; i < out.size(); i++) {
        r.out_tags.push_back(blk.shim_out[i].tags);
        if (r.produced > 0) blk.shim_out[i].n_items += (uint64_t)r.produced;
    }
    return r;
}

inline tag_t make_tag(uint64_t offset, const std::string &key, pmt::pmt_t value)
{
    tag_t t; t.offset = offset; t.key = pmt::string_to_symbol(key); t.value = value; t.srcid = pmt::get_PMT_F();
    return t;
}

}  // namespace shim
}  // namespace gr

#endif
