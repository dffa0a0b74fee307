/* oracle/shim/gnuradio/gr_complex.h -- minimal stand-in for GNU Radio's
 * gr_complex typedef so that the reference's lib/window.{h,cc} and
 * lib/firdes.{h,cc} compile from where they lie. Test infrastructure only. */
#pragma once
#include <complex>
#include <vector>   /* the real header reaches <vector> through gnuradio/types.h; lib/fft_filter.h relies on it */
typedef std::complex<float> gr_complex;
typedef std::complex<double> gr_complexd;
