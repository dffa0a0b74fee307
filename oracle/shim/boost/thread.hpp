/* oracle/shim/boost/thread.hpp -- TEST INFRASTRUCTURE ONLY: boost::mutex + scoped_lock over <mutex>,
 * all that the reference's lib/fft.h uses of Boost.Thread (Boost is not in the image). */
#pragma once
#include <mutex>
namespace boost {
class mutex : public std::mutex
{
public:
    typedef std::unique_lock<std::mutex> scoped_lock;
};
} // namespace boost
