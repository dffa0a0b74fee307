// The handful of pmt calls the X-engine block makes (lib/clXEngine_impl.cc:294-295,
// 1076-1080, 1202-1203 in the reference): symbols, pairs, c32 vectors, uint64 -- and the
// dict / f32vector / s32vector calls of clXCorrelate's "corr" PDU (lib/clXCorrelate_impl.cc:1585-1593).
#ifndef CLB200_PMT_STUB_H
#define CLB200_PMT_STUB_H
#include <complex>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace pmt {
struct pmt_base {
    enum kind_t { SYMBOL, PAIR, C32VECTOR, UINT64, NIL, F32VECTOR, S32VECTOR, DICT } kind = NIL;
    std::string sym;
    std::shared_ptr<pmt_base> car, cdr;
    std::vector<std::complex<float>> c32;
    uint64_t u64 = 0;
    std::vector<float> f32;
    std::vector<int32_t> s32;
    std::vector<std::pair<std::string, std::shared_ptr<pmt_base>>> dict;
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
static const pmt_t PMT_NIL = std::make_shared<pmt_base>();
inline pmt_t init_f32vector(size_t n, const float *data)
{
    auto p = std::make_shared<pmt_base>();
    p->kind = pmt_base::F32VECTOR;
    p->f32.assign(data, data + n);
    return p;
}
inline pmt_t init_s32vector(size_t n, const int32_t *data)
{
    auto p = std::make_shared<pmt_base>();
    p->kind = pmt_base::S32VECTOR;
    p->s32.assign(data, data + n);
    return p;
}
inline const std::vector<float> &f32vector_elements(const pmt_t &p) { return p->f32; }
inline const std::vector<int32_t> &s32vector_elements(const pmt_t &p) { return p->s32; }
inline pmt_t make_dict()
{
    auto p = std::make_shared<pmt_base>();
    p->kind = pmt_base::DICT;
    return p;
}
inline pmt_t dict_add(const pmt_t &d, const pmt_t &key, const pmt_t &value)
{
    auto p = std::make_shared<pmt_base>(*d);
    p->dict.emplace_back(key->sym, value);
    return p;
}
inline pmt_t dict_ref(const pmt_t &d, const pmt_t &key, const pmt_t &not_found)
{
    for (auto &kv : d->dict)
        if (kv.first == key->sym) return kv.second;
    return not_found;
}
} // namespace pmt
#endif
