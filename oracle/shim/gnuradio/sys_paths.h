/* oracle/shim/gnuradio/sys_paths.h -- TEST INFRASTRUCTURE ONLY: where lib/fft.cc keeps its FFTW wisdom
 * file.  Points into oracle/_ref/ (git-ignored) so nothing outside the repository is touched. */
#pragma once
namespace gr {
const char *appdata_path();
}
