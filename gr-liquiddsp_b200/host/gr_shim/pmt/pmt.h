// pmt.h (TEST SHIM) -- the subset of GNU Radio's polymorphic types that the three blocks use.
// GNU Radio is not installed in the build container, so the block sources compile against this
// header for the unit tests; against a real GNU Radio 3.7 tree the genuine <pmt/pmt.h> is used.
#pragma once
#include <complex>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost { using std::shared_ptr; }

namespace pmt {

struct pmt_base;
typedef std::shared_ptr<pmt_base> pmt_t;

struct pmt_base {
    enum kind_t { NIL, SYMBOL, LONG, PAIR, U8VEC, C32VEC, DICT } kind = NIL;
    std::string sym;
    long lval = 0;
    pmt_t car, cdr;
    std::vector<uint8_t> u8;
    std::vector<std::complex<float>> c32;
    std::map<std::string, pmt_t> dict;
};

inline pmt_t make_(pmt_base::kind_t k) { pmt_t p = std::make_shared<pmt_base>(); p->kind = k; return p; }
static const pmt_t PMT_NIL = make_(pmt_base::NIL);

inline pmt_t intern(const std::string &s) { pmt_t p = make_(pmt_base::SYMBOL); p->sym = s; return p; }
inline pmt_t mp(const std::string &s) { return intern(s); }
inline pmt_t mp(const char *s) { return intern(s); }
inline std::string symbol_to_string(const pmt_t &p) { return p->sym; }
inline pmt_t from_long(long v) { pmt_t p = make_(pmt_base::LONG); p->lval = v; return p; }
inline long to_long(const pmt_t &p) { if (p->kind != pmt_base::LONG) throw std::runtime_error("pmt::to_long: wrong type"); return p->lval; }
inline pmt_t cons(const pmt_t &a, const pmt_t &b) { pmt_t p = make_(pmt_base::PAIR); p->car = a; p->cdr = b; return p; }
inline pmt_t car(const pmt_t &p) { if (p->kind != pmt_base::PAIR) throw std::runtime_error("pmt::car: not a pair"); return p->car; }
inline pmt_t cdr(const pmt_t &p) { if (p->kind != pmt_base::PAIR) throw std::runtime_error("pmt::cdr: not a pair"); return p->cdr; }
inline pmt_t init_u8vector(size_t n, const uint8_t *d) { pmt_t p = make_(pmt_base::U8VEC); p->u8.assign(d, d + n); return p; }
inline pmt_t init_u8vector(size_t n, const std::vector<uint8_t> &d) { return init_u8vector(n, d.data()); }
inline pmt_t init_c32vector(size_t n, const std::complex<float> *d) { pmt_t p = make_(pmt_base::C32VEC); p->c32.assign(d, d + n); return p; }
inline pmt_t init_c32vector(size_t n, const std::vector<std::complex<float>> &d) { return init_c32vector(n, d.data()); }
inline std::vector<uint8_t> u8vector_elements(const pmt_t &p) { if (p->kind != pmt_base::U8VEC) throw std::runtime_error("pmt: not a u8vector"); return p->u8; }
inline std::vector<std::complex<float>> c32vector_elements(const pmt_t &p) { if (p->kind != pmt_base::C32VEC) throw std::runtime_error("pmt: not a c32vector"); return p->c32; }
inline size_t length(const pmt_t &p) { return p->kind == pmt_base::U8VEC ? p->u8.size() : p->kind == pmt_base::C32VEC ? p->c32.size() : p->dict.size(); }
inline pmt_t make_dict() { return make_(pmt_base::DICT); }
inline pmt_t dict_add(const pmt_t &d, const pmt_t &k, const pmt_t &v) { pmt_t p = std::make_shared<pmt_base>(*d); p->dict[k->sym] = v; return p; }
inline bool dict_has_key(const pmt_t &d, const pmt_t &k) { return d->kind == pmt_base::DICT && d->dict.count(k->sym) != 0; }
inline pmt_t dict_ref(const pmt_t &d, const pmt_t &k, const pmt_t &dflt) { auto it = d->dict.find(k->sym); return it == d->dict.end() ? dflt : it->second; }
inline bool is_dict(const pmt_t &p) { return p->kind == pmt_base::DICT; }

}  // namespace pmt
