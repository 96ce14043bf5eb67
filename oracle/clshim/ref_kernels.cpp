// ref_kernels.cpp — runs the reference's OWN OpenCL kernel source (extracted at build time into oracle/_ref/)
// on the CPU, with the launch geometry and buffer shapes of Watermark.cpp.  TEST INFRASTRUCTURE ONLY.
//   ref_nvf              <- Watermark::computeCustomMask           (Watermark.cpp:96-114)
//   ref_me               <- Watermark::computePredictionErrorMask  (Watermark.cpp:176-196), partials only
//   ref_scaled_neighbors <- Watermark::computeScaledNeighbors      (Watermark.cpp:117-136)
// All image buffers are ArrayFire column-major (rows, cols): element (r, c) at c*rows + r.
#include "clc_shim.hpp"

namespace clc {
thread_local Item* cur = nullptr;
thread_local Range rng;
thread_local ucontext_t sched;
thread_local const std::function<void()>* body = nullptr;

static void fiber_entry()
{
    (*body)();
    cur->done = true;
    swapcontext(&cur->ctx, &sched);
}

void run_ndrange(size_t g0, size_t g1, size_t l0, size_t l1, const std::function<void()>& kernel_body)
{
    const size_t nitems = l0 * l1, STK = 64 * 1024;
    std::vector<Item> items(nitems);
    std::vector<char> stacks(nitems * STK);
    rng.global[0] = g0; rng.global[1] = g1; rng.global[2] = 1;
    rng.local[0] = l0; rng.local[1] = l1; rng.local[2] = 1;
    body = &kernel_body;
    for (size_t gy = 0; gy < g1 / l1; gy++)
        for (size_t gx = 0; gx < g0 / l0; gx++) {
            rng.group[0] = gx; rng.group[1] = gy; rng.group[2] = 0;
            for (size_t ly = 0; ly < l1; ly++)
                for (size_t lx = 0; lx < l0; lx++) {
                    Item& it = items[ly * l0 + lx];
                    it.done = false;
                    it.lid[0] = lx; it.lid[1] = ly; it.lid[2] = 0;
                    it.gid[0] = gx * l0 + lx; it.gid[1] = gy * l1 + ly; it.gid[2] = 0;
                    getcontext(&it.ctx);
                    it.ctx.uc_stack.ss_sp = stacks.data() + (ly * l0 + lx) * STK;
                    it.ctx.uc_stack.ss_size = STK;
                    it.ctx.uc_link = nullptr;
                    makecontext(&it.ctx, fiber_entry, 0);
                }
            bool alive = true;
            while (alive) {  // one pass = one barrier phase
                alive = false;
                for (size_t i = 0; i < nitems; i++) {
                    if (items[i].done) continue;
                    cur = &items[i];
                    swapcontext(&sched, &items[i].ctx);
                    if (!items[i].done) alive = true;
                }
            }
        }
}
}  // namespace clc

using namespace clc;

// the kernel text, verbatim from the reference apart from the `(floatN)(` -> `make_floatN(` rewrite
namespace k_nvf {
#define p 3  // main.cpp:106 builds nvf with -Dp=<settings.ini p>; main.cpp:89 only lets 3 through
#include "nvf.cl.inc"
#undef p
}
// the other window sizes the class accepts (Watermark.cpp:24): the same kernel text with -Dp=5 / 7 / 9
namespace k_nvf5 {
#define p 5
#include "nvf.cl.inc"
#undef p
}
namespace k_nvf7 {
#define p 7
#include "nvf.cl.inc"
#undef p
}
namespace k_nvf9 {
#define p 9
#include "nvf.cl.inc"
#undef p
}
namespace k_me {
#include "me_p3.cl.inc"
}
namespace k_sn {
#include "scaled_neighbors_p3.cl.inc"
}
static const int RxMappings[64] = {
#include "rxmappings.inc"  // Watermark.hpp:29-39, extracted
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) & ~(a - 1); }

extern "C" {

void ref_nvf(const float* img, int rows, int cols, float* out)
{
    const image2d im{img, rows, cols};  // Watermark.cpp:57: Image2D(width = rows, height = cols)
    std::vector<float> local(19 * 19);  // Watermark.cpp:100: (16 + p)^2 floats
    run_ndrange(align_up(rows, 16), align_up(cols, 16), 16, 16,
                [&]() { k_nvf::nvf(&im, out, reinterpret_cast<float(*)[18]>(local.data())); });
}

int ref_nvf_p(const float* img, int rows, int cols, int pw, float* out)
{
    const image2d im{img, rows, cols};
    std::vector<float> local((16 + pw) * (16 + pw));  // Watermark.cpp:100
    const size_t gr = align_up(rows, 16), gc = align_up(cols, 16);
    switch (pw) {
    case 3: run_ndrange(gr, gc, 16, 16, [&]() { k_nvf::nvf(&im, out, reinterpret_cast<float(*)[18]>(local.data())); }); return 0;
    case 5: run_ndrange(gr, gc, 16, 16, [&]() { k_nvf5::nvf(&im, out, reinterpret_cast<float(*)[20]>(local.data())); }); return 0;
    case 7: run_ndrange(gr, gc, 16, 16, [&]() { k_nvf7::nvf(&im, out, reinterpret_cast<float(*)[22]>(local.data())); }); return 0;
    case 9: run_ndrange(gr, gc, 16, 16, [&]() { k_nvf9::nvf(&im, out, reinterpret_cast<float(*)[24]>(local.data())); }); return 0;
    default: return -1;
    }
}

void ref_scaled_neighbors(const float* img, int rows, int cols, const float* coeffs, float* out)
{
    const image2d im{img, rows, cols};
    std::vector<float> local(324);  // Watermark.cpp:127
    run_ndrange(align_up(rows, 16), align_up(cols, 16), 16, 16,
                [&]() { k_sn::scaled_neighbors_p3(&im, out, coeffs, reinterpret_cast<float(*)[18]>(local.data())); });
}

// RxPartial: rows * align64(cols) floats, rxPartial: rows * align64(cols) / 8 floats (Watermark.cpp:178-179)
void ref_me(const float* img, int rows, int cols, float* RxPartial, float* rxPartial)
{
    const image2d im{img, rows, cols};
    std::vector<half> local(2304);  // Watermark.cpp:189
    run_ndrange(align_up(cols, 64), rows, 64, 1,  // Watermark.cpp:190
                [&]() { k_me::me(&im, RxPartial, rxPartial, RxMappings, reinterpret_cast<half(*)[36]>(local.data())); });
}

}  // extern "C"
