// The reference's timing protocol (main.cpp:167-223) in C++ over the façade: one image, 2 warm-ups, `loops`
// synchronous calls per op, mean -> FPS.  usage: latency <rows> <cols> <loops> [graphs=1]
// build: g++ -O2 -std=c++20 -Iinclude -Iwatermarking-gpu_b200/csrc tools/latency.cpp -Lwatermarking-gpu_b200 -lwm_b200 -Wl,-rpath,$PWD/watermarking-gpu_b200 -o tools/latency_cpp
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "Watermark.hpp"

int main(int argc, char** argv)
{
    const dim_t rows = argc > 1 ? atoll(argv[1]) : 1080, cols = argc > 2 ? atoll(argv[2]) : 1920;
    const int loops = argc > 3 ? atoi(argv[3]) : 1000, graphs = argc > 4 ? atoi(argv[4]) : 1;
    std::mt19937 g(1);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> w((size_t)(rows * cols)), img((size_t)(rows * cols));
    for (auto& v : w) v = nd(g);
    // smooth-ish synthetic image (column-major), natural-image-like neighbour correlation
    for (dim_t c = 0; c < cols; c++)
        for (dim_t r = 0; r < rows; r++)
            img[(size_t)(c * rows + r)] = 128.f + 60.f * std::sin(0.013f * r) * std::cos(0.017f * c) + 25.f * std::sin(0.11f * (r + 2 * c)) + 3.f * nd(g);
    Watermark wm(rows, cols, w.data(), 3, 40.0f);
    wm_set_option(wm.handle(), WM_OPT_CUDA_GRAPHS, graphs);
    const wm::Image image(rows, cols, img.data());
    wm::Image out(rows, cols);
    float a = 0, corr = 0;
    auto timeit = [&](const char* name, auto&& fn) {
        fn(); fn();
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < loops; i++) fn();
        const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / loops;
        std::printf("%-11s %8.1f us/call %9.1f FPS\n", name, us, 1e6 / us);
        return us;
    };
    double tot = 0;
    tot += timeit("embed NVF", [&] { wm_embed(wm.handle(), image.desc(), image.desc(), out.desc(), WM_MASK_NVF, &a); });
    tot += timeit("detect NVF", [&] { wm_detect(wm.handle(), out.desc(), WM_MASK_NVF, &corr); });
    std::printf("   a=%.5f corr=%.5f\n", a, corr);
    tot += timeit("embed ME", [&] { wm_embed(wm.handle(), image.desc(), image.desc(), out.desc(), WM_MASK_ME, &a); });
    tot += timeit("detect ME", [&] { wm_detect(wm.handle(), out.desc(), WM_MASK_ME, &corr); });
    std::printf("   a=%.5f corr=%.5f\n", a, corr);
    std::printf("all four ops: %.1f us -> %.1f frames/s (%lldx%lld, graphs=%d)\n", tot, 1e6 / tot, rows, cols, graphs);
    return 0;
}
