/* oracle/shim/opencl_c.h -- TEST INFRASTRUCTURE ONLY.
 * Just enough of OpenCL C for g++ to compile the kernel source strings that the REFERENCE builds at run time
 * (oracle/ref_kernels.py makes the reference's own buildKernel()/buildProgram() code emit them): address-space
 * qualifiers, the work-item index functions (set by the driver loop that stands in for the NDRange), and the
 * float2 / float4 / uchar2 vector types with the operators those kernels use.  No OpenCL runtime is in the image. */
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

#define __kernel
#define __global
#define __constant const
#define restrict __restrict__

typedef unsigned char uchar;
static thread_local size_t clb_gid[3], clb_gsz[3];
static inline size_t get_global_id(int d) { return clb_gid[d]; }
static inline size_t get_global_size(int d) { return clb_gsz[d]; }

struct float2 {
    float x, y;
    float2() = default;
    float2(float v) : x(v), y(v) {}                 /* `float2 acc = 0;` */
    float2(float a, float b) : x(a), y(b) {}
};
struct float4 {
    float x, y, z, w;
};
struct uchar2 {
    uchar x, y;
};
/* fma(float2, float, float2): the scalar widens to the vector (lib/clPolyphaseChannelizer_impl.cc:165) */
static inline float2 fma(float2 a, float b, float2 c) { return float2(std::fma(a.x, b, c.x), std::fma(a.y, b, c.y)); }
using std::fma;
using std::sqrt;
