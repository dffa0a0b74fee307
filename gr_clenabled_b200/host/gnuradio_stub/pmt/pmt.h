// The handful of pmt calls the X-engine block makes (lib/clXEngine_impl.cc:294-295,
// 1076-1080, 1202-1203 in the reference): symbols, pairs, c32 vectors, uint64.
#ifndef CLB200_PMT_STUB_H
#define CLB200_PMT_STUB_H
#include <complex>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace pmt {
struct pmt_base {
    enum kind_t { SYMBOL, PAIR, C32VECTOR, UINT64, NIL } kind = NIL;
    std::string sym;
    std::shared_ptr<pmt_base> car, cdr;
    std::vector<std::complex<float>> c32;
    uint64_t u64 = 0;
};
typedef std::shared_ptr<pmt_base> pmt_t;

inline pmt_t string_to_symbol(const std::string &s)
{
    auto p = std::make_shared<pmt_base>();
    p->kind = pmt_base::SYMBOL;
    p->sym = s;
    return p;
}
inline pmt_t intern(const std::string &s) { return string_to_symbol(s); }
inline pmt_t mp(const std::string &s) { return string_to_symbol(s); }
inline std::string symbol_to_string(const pmt_t &p) { return p->sym; }
inline pmt_t cons(const pmt_t &a, const pmt_t &b)
{
    auto p = std::make_shared<pmt_base>();
    p->kind = pmt_base::PAIR;
    p->car = a;
    p->cdr = b;
    return p;
}
inline pmt_t car(const pmt_t &p) { return p->car; }
inline pmt_t cdr(const pmt_t &p) { return p->cdr; }
inline pmt_t init_c32vector(size_t n, const std::complex<float> *data)
{
    auto p = std::make_shared<pmt_base>();
    p->kind = pmt_base::C32VECTOR;
    p->c32.assign(data, data + n);
    return p;
}
inline const std::vector<std::complex<float>> &c32vector_elements(const pmt_t &p) { return p->c32; }
inline pmt_t from_uint64(uint64_t v)
{
    auto p = std::make_shared<pmt_base>();
    p->kind = pmt_base::UINT64;
    p->u64 = v;
    return p;
}
inline uint64_t to_uint64(const pmt_t &p) { return p->u64; }
} // namespace pmt
#endif
