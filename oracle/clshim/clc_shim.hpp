// clc_shim.hpp — just enough of OpenCL C 2.0 to compile the reference's three kernel strings
// (Watermark_GPU/kernels/{nvf,me_p3,scaled_neighbors_p3}.hpp) as C++ and run them on the CPU.
//
// TEST INFRASTRUCTURE ONLY (see oracle/wm_oracle.c header).  The kernel text itself is never copied into
// this repository: oracle/clshim/extract.py reads it from /root/reference at build time and writes it,
// with the OpenCL vector-literal syntax `(floatN)(...)` rewritten to a function call, under oracle/_ref/.
//
// Execution model: an NDRange is run work-group by work-group; the work-items of a group are ucontext
// fibers, barrier() yields to the group scheduler, so local-memory phases behave as on a device.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <vector>

namespace clc {

typedef _Float16 half;  // cl_khr_fp16 storage type; conversions round to nearest even like vstore_half
struct int2 { int x, y; };
struct float4 { float x, y, z, w; };
struct float8 { float s[8]; };
inline int2 make_int2(int x, int y) { return {x, y}; }
inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
inline float8 make_float8(float a, float b, float c, float d, float e, float f, float g, float h) { return {{a, b, c, d, e, f, g, h}}; }

// image2d_t(CL_LUMINANCE, CL_FLOAT, width, height) filled by clEnqueueCopyBufferToImage (opencl_utils.cpp:17-22)
struct image2d { const float* data; int width, height; };
typedef const image2d* image2d_t;
typedef int sampler_t;
enum { CLK_NORMALIZED_COORDS_FALSE = 0, CLK_ADDRESS_CLAMP_TO_EDGE = 2, CLK_FILTER_NEAREST = 0x10, CLK_LOCAL_MEM_FENCE = 1 };
inline int get_image_width(image2d_t im) { return im->width; }
inline int get_image_height(image2d_t im) { return im->height; }
inline float4 read_imagef(image2d_t im, sampler_t, int2 c)  // unnormalised, clamp-to-edge, nearest
{
    const int x = std::min(std::max(c.x, 0), im->width - 1), y = std::min(std::max(c.y, 0), im->height - 1);
    const float v = im->data[(size_t)y * im->width + x];
    return {v, 0.0f, 0.0f, 1.0f};  // CL_LUMINANCE: (L, L, L, 1); the kernels read .x only
}
inline void vstore_half8(float8 v, size_t off, half* p) { for (int i = 0; i < 8; i++) p[off * 8 + i] = (half)v.s[i]; }
inline void vstore_half4(float4 v, size_t off, half* p)
{
    p[off * 4 + 0] = (half)v.x; p[off * 4 + 1] = (half)v.y; p[off * 4 + 2] = (half)v.z; p[off * 4 + 3] = (half)v.w;
}

// ---- work-item state and the fiber scheduler ----
struct Item { ucontext_t ctx; bool done; size_t gid[3], lid[3]; };
struct Range { size_t global[3], local[3], group[3]; };
extern thread_local Item* cur;
extern thread_local Range rng;
extern thread_local ucontext_t sched;
extern thread_local const std::function<void()>* body;

inline size_t get_global_id(int d) { return cur->gid[d]; }
inline size_t get_local_id(int d) { return cur->lid[d]; }
inline size_t get_group_id(int d) { return rng.group[d]; }
inline size_t get_local_size(int d) { return rng.local[d]; }
inline size_t get_global_size(int d) { return rng.global[d]; }
inline void barrier(int) { swapcontext(&cur->ctx, &sched); }

void run_ndrange(size_t g0, size_t g1, size_t l0, size_t l1, const std::function<void()>& kernel_body);

}  // namespace clc

// OpenCL C qualifiers
#define __kernel
#define __global
#define __local
#define __constant const
#define __read_only
