/* oracle/shim/boost/filesystem/path.hpp -- TEST INFRASTRUCTURE ONLY: the path join lib/fft.cc:95 does. */
#pragma once
#include <string>
namespace boost {
namespace filesystem {
class path
{
    std::string d;

public:
    path() {}
    path(const std::string &s) : d(s) {}
    path(const char *s) : d(s) {}
    path operator/(const path &o) const { return path(d.empty() || d.back() == '/' ? d + o.d : d + "/" + o.d); }
    const std::string &string() const { return d; }
};
} // namespace filesystem
} // namespace boost
