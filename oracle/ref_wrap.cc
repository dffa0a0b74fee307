/*
 * oracle/ref_wrap.cc -- TEST INFRASTRUCTURE ONLY.
 * extern "C" entry points over the REFERENCE's own tap-design code
 * (lib/firdes.cc, lib/window.cc), which oracle/Makefile compiles from where it
 * lies under /root/reference into oracle/_ref/libref_firdes.so.  Used only to
 * pin oracle_c.c's window/firdes restatement and to generate tests/golden/.
 * No reference source is copied into this repository.
 */
#include <vector>
#include <cstring>
#include "firdes.h"
#include "window.h"

using gr::clenabled::firdes;
using gr::clenabled::window;

extern "C" {

/* window::build(type, ntaps, beta)  (lib/window.cc) */
__attribute__((visibility("default")))
int ref_window_build(int type, int ntaps, double beta, float *out)
{
    try {
        std::vector<float> w = window::build(static_cast<window::win_type>(type), ntaps, beta);
        std::memcpy(out, w.data(), sizeof(float) * w.size());
        return (int)w.size();
    } catch (...) {
        return -1;
    }
}

/* firdes::low_pass(gain, fs, fc, tw, win, beta)  (lib/firdes.cc:93-135).
 * Returns ntaps; writes at most cap floats. */
__attribute__((visibility("default")))
int ref_firdes_low_pass(double gain, double fs, double fc, double tw, int win, double beta,
                        float *out, int cap)
{
    try {
        std::vector<float> t =
            firdes::low_pass(gain, fs, fc, tw, static_cast<firdes::win_type>(win), beta);
        int n = (int)t.size();
        if (n <= cap) std::memcpy(out, t.data(), sizeof(float) * n);
        return n;
    } catch (...) {
        return -1;
    }
}

__attribute__((visibility("default")))
int ref_firdes_high_pass(double gain, double fs, double fc, double tw, int win, double beta,
                         float *out, int cap)
{
    try {
        std::vector<float> t =
            firdes::high_pass(gain, fs, fc, tw, static_cast<firdes::win_type>(win), beta);
        int n = (int)t.size();
        if (n <= cap) std::memcpy(out, t.data(), sizeof(float) * n);
        return n;
    } catch (...) {
        return -1;
    }
}

__attribute__((visibility("default")))
int ref_firdes_band_pass(double gain, double fs, double f1, double f2, double tw, int win,
                         double beta, float *out, int cap)
{
    try {
        std::vector<float> t =
            firdes::band_pass(gain, fs, f1, f2, tw, static_cast<firdes::win_type>(win), beta);
        int n = (int)t.size();
        if (n <= cap) std::memcpy(out, t.data(), sizeof(float) * n);
        return n;
    } catch (...) {
        return -1;
    }
}

__attribute__((visibility("default")))
int ref_firdes_root_raised_cosine(double gain, double fs, double symrate, double alpha, int ntaps,
                                  float *out, int cap)
{
    try {
        std::vector<float> t = firdes::root_raised_cosine(gain, fs, symrate, alpha, ntaps);
        int n = (int)t.size();
        if (n <= cap) std::memcpy(out, t.data(), sizeof(float) * n);
        return n;
    } catch (...) {
        return -1;
    }
}

} /* extern "C" */
