/* oracle/shim/shim_impl.cc -- TEST INFRASTRUCTURE ONLY.
 * Bodies behind oracle/shim/fftw3.h and gnuradio/sys_paths.h: the transforms lib/fft.cc plans.
 * Powers of two: iterative decimation-in-time radix-2 on float data, twiddles generated in double and
 * rounded once (error ~1e-7 of the output maximum, the level of FFTW's float codelets); other sizes:
 * direct DFT accumulated in double.  Unnormalised, sign as given. */
#include "fftw3.h"
#include <gnuradio/sys_paths.h>

#include <dlfcn.h>
#include <cmath>
#include <complex>
#include <cstring>
#include <string>
#include <vector>

struct clb_shim_fftwf_plan_s {
    int n, sign, kind;            // kind 0: c2c, 1: r2c, 2: c2r
    void *in, *out;
    std::vector<std::complex<float>> tw, work;
    std::vector<int> rev;
};

namespace {
typedef std::complex<float> cf;

void prepare(clb_shim_fftwf_plan_s *p)
{
    const int n = p->n;
    p->work.resize(n);
    if (n & (n - 1)) return;
    p->tw.resize(n / 2 > 0 ? n / 2 : 1);
    for (int k = 0; k < n / 2; k++) {
        const double a = p->sign * 2.0 * M_PI * (double)k / (double)n;
        p->tw[k] = cf((float)cos(a), (float)sin(a));
    }
    p->rev.resize(n);
    int bits = 0;
    while ((1 << bits) < n) bits++;
    for (int i = 0; i < n; i++) {
        int r = 0;
        for (int b = 0; b < bits; b++) r |= ((i >> b) & 1) << (bits - 1 - b);
        p->rev[i] = r;
    }
}

// x (length n) -> p->work
void c2c(clb_shim_fftwf_plan_s *p, const cf *x)
{
    const int n = p->n;
    cf *w = p->work.data();
    if (n & (n - 1)) {
        for (int k = 0; k < n; k++) {
            double sr = 0, si = 0;
            for (int j = 0; j < n; j++) {
                const double a = p->sign * 2.0 * M_PI * (double)(((long)j * k) % n) / (double)n;
                const double c = cos(a), s = sin(a);
                sr += x[j].real() * c - x[j].imag() * s;
                si += x[j].real() * s + x[j].imag() * c;
            }
            w[k] = cf((float)sr, (float)si);
        }
        return;
    }
    for (int i = 0; i < n; i++) w[p->rev[i]] = x[i];
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len / 2, step = n / len;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < half; k++) {
                const cf t = p->tw[k * step];
                const cf b = w[i + k + half];
                const cf v(b.real() * t.real() - b.imag() * t.imag(), b.real() * t.imag() + b.imag() * t.real());
                const cf a = w[i + k];
                w[i + k] = a + v;
                w[i + k + half] = a - v;
            }
    }
}
} // namespace

extern "C" {

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned)
{
    if (n <= 0) return nullptr;
    auto *p = new clb_shim_fftwf_plan_s{n, sign, 0, in, out, {}, {}, {}};
    prepare(p);
    return p;
}
fftwf_plan fftwf_plan_dft_r2c_1d(int n, float *in, fftwf_complex *out, unsigned)
{
    if (n <= 0) return nullptr;
    auto *p = new clb_shim_fftwf_plan_s{n, -1, 1, in, out, {}, {}, {}};
    prepare(p);
    return p;
}
fftwf_plan fftwf_plan_dft_c2r_1d(int n, fftwf_complex *in, float *out, unsigned)
{
    if (n <= 0) return nullptr;
    auto *p = new clb_shim_fftwf_plan_s{n, +1, 2, in, out, {}, {}, {}};
    prepare(p);
    return p;
}
void fftwf_execute(const fftwf_plan p)
{
    const int n = p->n;
    if (p->kind == 0) {
        c2c(p, (const cf *)p->in);
        memcpy(p->out, p->work.data(), sizeof(cf) * n);
    } else if (p->kind == 1) {                       // n real -> n/2+1 complex
        std::vector<cf> x(n);
        for (int i = 0; i < n; i++) x[i] = cf(((const float *)p->in)[i], 0.f);
        c2c(p, x.data());
        memcpy(p->out, p->work.data(), sizeof(cf) * (n / 2 + 1));
    } else {                                         // n/2+1 complex (Hermitian half) -> n real
        std::vector<cf> x(n);
        const cf *h = (const cf *)p->in;
        for (int i = 0; i <= n / 2; i++) x[i] = h[i];
        for (int i = n / 2 + 1; i < n; i++) x[i] = std::conj(h[n - i]);
        c2c(p, x.data());
        for (int i = 0; i < n; i++) ((float *)p->out)[i] = p->work[i].real();
    }
}
void fftwf_destroy_plan(fftwf_plan p) { delete p; }

} // extern "C"

namespace gr {
const char *appdata_path()
{
    static std::string dir = [] {
        Dl_info info;
        std::string d = ".";
        if (dladdr((void *)&fftwf_execute, &info) && info.dli_fname) {
            d = info.dli_fname;
            const size_t k = d.rfind('/');
            d = (k == std::string::npos) ? "." : d.substr(0, k);
        }
        return d;
    }();
    return dir.c_str();
}
} // namespace gr
