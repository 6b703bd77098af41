// pmt/pmt.h -- the subset of GNU Radio 3.8's polymorphic types the radar-path blocks use.
// Stand-in used ONLY when GNU Radio is not installed (see ../README.md); with a real GNU Radio
// the compiler finds the real <pmt/pmt.h> first and this directory is not on the include path.
#ifndef JRC_SHIM_PMT_H
#define JRC_SHIM_PMT_H
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace pmt {

struct pmt_base {
    enum kind_t { NIL, BOOL, SYMBOL, LONG, UINT64, DOUBLE, TUPLE, F32VEC, PAIR } kind = NIL;
    bool b = false;
    std::string sym;
    long l = 0;
    uint64_t u = 0;
    double d = 0.0;
    std::vector<std::shared_ptr<pmt_base>> items;   // tuple elements / pair {car, cdr}
    std::vector<float> f32;
};
typedef std::shared_ptr<pmt_base> pmt_t;

class wrong_type : public std::invalid_argument {
public:
    explicit wrong_type(const std::string &m) : std::invalid_argument(m) {}
};

inline pmt_t mk_(pmt_base::kind_t k) { auto p = std::make_shared<pmt_base>(); p->kind = k; return p; }
inline pmt_t get_PMT_NIL() { static pmt_t p = mk_(pmt_base::NIL); return p; }
inline pmt_t get_PMT_T() { static pmt_t p = [] { auto q = mk_(pmt_base::BOOL); q->b = true; return q; }(); return p; }
inline pmt_t get_PMT_F() { static pmt_t p = [] { auto q = mk_(pmt_base::BOOL); q->b = false; return q; }(); return p; }
#define PMT_NIL pmt::get_PMT_NIL()
#define PMT_T pmt::get_PMT_T()
#define PMT_F pmt::get_PMT_F()

inline pmt_t string_to_symbol(const std::string &s) { auto p = mk_(pmt_base::SYMBOL); p->sym = s; return p; }
inline pmt_t intern(const std::string &s) { return string_to_symbol(s); }
inline pmt_t mp(const std::string &s) { return string_to_symbol(s); }
inline pmt_t mp(const char *s) { return string_to_symbol(s); }
inline bool is_symbol(const pmt_t &p) { return p && p->kind == pmt_base::SYMBOL; }
inline const std::string symbol_to_string(const pmt_t &p)
{
    if (!is_symbol(p)) throw wrong_type("pmt::symbol_to_string");
    return p->sym;
}
inline pmt_t from_long(long v) { auto p = mk_(pmt_base::LONG); p->l = v; return p; }
inline pmt_t mp(long v) { return from_long(v); }
inline pmt_t mp(int v) { return from_long(v); }
inline pmt_t from_uint64(uint64_t v) { auto p = mk_(pmt_base::UINT64); p->u = v; return p; }
inline pmt_t from_double(double v) { auto p = mk_(pmt_base::DOUBLE); p->d = v; return p; }
inline pmt_t from_float(float v) { return from_double(v); }
inline pmt_t from_bool(bool v) { return v ? get_PMT_T() : get_PMT_F(); }
inline bool is_integer(const pmt_t &p) { return p && p->kind == pmt_base::LONG; }
inline bool is_uint64(const pmt_t &p) { return p && p->kind == pmt_base::UINT64; }
inline bool is_real(const pmt_t &p) { return p && p->kind == pmt_base::DOUBLE; }
inline long to_long(const pmt_t &p)
{
    if (is_integer(p)) return p->l;
    if (is_uint64(p)) return (long)p->u;
    throw wrong_type("pmt::to_long");
}
inline uint64_t to_uint64(const pmt_t &p)
{
    if (is_uint64(p)) return p->u;
    if (is_integer(p) && p->l >= 0) return (uint64_t)p->l;   // GR 3.8 pmt.cc: non-negative integers convert
    throw wrong_type("pmt::to_uint64");
}
inline double to_double(const pmt_t &p)
{
    if (is_real(p)) return p->d;
    if (is_integer(p)) return (double)p->l;
    if (is_uint64(p)) return (double)p->u;
    throw wrong_type("pmt::to_double");
}
inline float to_float(const pmt_t &p) { return (float)to_double(p); }
inline bool to_bool(const pmt_t &p)
{
    if (p && p->kind == pmt_base::BOOL) return p->b;
    throw wrong_type("pmt::to_bool");
}

inline pmt_t make_tuple(const pmt_t &a, const pmt_t &b) { auto p = mk_(pmt_base::TUPLE); p->items = {a, b}; return p; }
inline bool is_tuple(const pmt_t &p) { return p && p->kind == pmt_base::TUPLE; }
inline pmt_t tuple_ref(const pmt_t &p, size_t k)
{
    if (!is_tuple(p) || k >= p->items.size()) throw wrong_type("pmt::tuple_ref");
    return p->items[k];
}

inline pmt_t init_f32vector(size_t k, const float *data) { auto p = mk_(pmt_base::F32VEC); p->f32.assign(data, data + k); return p; }
inline pmt_t init_f32vector(size_t k, const std::vector<float> &data) { return init_f32vector(k, data.data()); }
inline bool is_f32vector(const pmt_t &p) { return p && p->kind == pmt_base::F32VEC; }
inline const std::vector<float> f32vector_elements(const pmt_t &p)
{
    if (!is_f32vector(p)) throw wrong_type("pmt::f32vector_elements");
    return p->f32;
}
inline const float *f32vector_elements(const pmt_t &p, size_t &len)
{
    if (!is_f32vector(p)) throw wrong_type("pmt::f32vector_elements");
    len = p->f32.size();
    return p->f32.data();
}

inline pmt_t cons(const pmt_t &a, const pmt_t &b) { auto p = mk_(pmt_base::PAIR); p->items = {a, b}; return p; }
inline bool is_pair(const pmt_t &p) { return p && p->kind == pmt_base::PAIR; }
inline bool is_null(const pmt_t &p) { return !p || p->kind == pmt_base::NIL; }
inline pmt_t car(const pmt_t &p) { if (!is_pair(p)) throw wrong_type("pmt::car"); return p->items[0]; }
inline pmt_t cdr(const pmt_t &p) { if (!is_pair(p)) throw wrong_type("pmt::cdr"); return p->items[1]; }
inline pmt_t list1(const pmt_t &a) { return cons(a, get_PMT_NIL()); }
inline pmt_t list2(const pmt_t &a, const pmt_t &b) { return cons(a, list1(b)); }
inline pmt_t list3(const pmt_t &a, const pmt_t &b, const pmt_t &c) { return cons(a, list2(b, c)); }
inline pmt_t list4(const pmt_t &a, const pmt_t &b, const pmt_t &c, const pmt_t &d) { return cons(a, list3(b, c, d)); }
inline pmt_t nth(size_t n, pmt_t list)
{
    while (n-- > 0) list = cdr(list);
    return car(list);
}
inline size_t length(pmt_t list)
{
    if (is_tuple(list)) return list->items.size();
    if (is_f32vector(list)) return list->f32.size();
    size_t n = 0;
    while (is_pair(list)) { n++; list = cdr(list); }
    return n;
}

inline bool eqv(const pmt_t &a, const pmt_t &b)
{
    if (a == b) return true;
    if (!a || !b || a->kind != b->kind) return false;
    switch (a->kind) {
    case pmt_base::NIL: return true;
    case pmt_base::BOOL: return a->b == b->b;
    case pmt_base::SYMBOL: return a->sym == b->sym;     // symbols are interned in GNU Radio
    case pmt_base::LONG: return a->l == b->l;
    case pmt_base::UINT64: return a->u == b->u;
    case pmt_base::DOUBLE: return a->d == b->d;
    default: return false;
    }
}
inline bool eq(const pmt_t &a, const pmt_t &b) { return eqv(a, b); }
inline bool equal(const pmt_t &a, const pmt_t &b)
{
    if (eqv(a, b)) return true;
    if (!a || !b || a->kind != b->kind) return false;
    if (a->kind == pmt_base::F32VEC) return a->f32 == b->f32;
    if (a->kind == pmt_base::TUPLE || a->kind == pmt_base::PAIR) {
        if (a->items.size() != b->items.size()) return false;
        for (size_t i = 0; i < a->items.size(); i++) if (!equal(a->items[i], b->items[i])) return false;
        return true;
    }
    return false;
}

}  // namespace pmt
#endif
