// C++ check of the Watermark façade (csrc/Watermark.hpp) and the per-frame driver (csrc/videoprocessingcontext.hpp)
// — written the way the reference's testForImage / testForVideo use the class (main.cpp:140-316).
// usage: test_facade <gray.f32 row-major> <w.dat> <rows> <cols> <frames.u8> <nframes> <linesize>
// prints one line per result: "<name> <value>" (the pytest wrapper compares them with the CPU oracle).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include "videoprocessingcontext.hpp"

template <typename T>
static std::vector<T> slurp(const char* path, size_t n)
{
    std::vector<T> v(n);
    std::ifstream f(path, std::ios::binary);
    if (!f.read(reinterpret_cast<char*>(v.data()), n * sizeof(T))) { std::cerr << "cannot read " << path << "\n"; std::exit(2); }
    return v;
}

int main(int argc, char** argv)
{
    if (argc != 8) { std::cerr << "usage\n"; return 2; }
    const dim_t rows = atoll(argv[3]), cols = atoll(argv[4]);
    const int nframes = atoi(argv[6]), linesize = atoi(argv[7]);
    try {
        // wrong p / wrong W size behave like the reference (Watermark.cpp:24-25,70-71)
        try { Watermark bad(rows, cols, argv[2], 4, 40.0f); std::cout << "bad_p no-throw\n"; }
        catch (const std::runtime_error& e) { std::cout << "bad_p " << (std::string(e.what()).find("Wrong p parameter: 4") == 0 ? 1 : 0) << "\n"; }
        try { Watermark bad(rows + 1, cols, argv[2], 3, 40.0f); std::cout << "bad_w no-throw\n"; }
        catch (const std::runtime_error& e) { std::cout << "bad_w " << (std::string(e.what()).find("W file total elements != image dimensions") != std::string::npos ? 1 : 0) << "\n"; }

        Watermark watermarkObj(rows, cols, argv[2], 3, 40.0f);
        // af::array is column-major: transpose the row-major test image on the host
        const std::vector<float> rm = slurp<float>(argv[1], (size_t)(rows * cols));
        std::vector<float> cm((size_t)(rows * cols));
        for (dim_t r = 0; r < rows; r++)
            for (dim_t c = 0; c < cols; c++) cm[(size_t)(c * rows + r)] = rm[(size_t)(r * cols + c)];
        const wm::Image image(rows, cols, cm.data());
        float a = 0.0f;
        // main.cpp:175-222
        const wm::Image wNVF = watermarkObj.makeWatermark(image, image, a, MASK_TYPE::NVF);
        std::printf("a_nvf %.9g\n", a);
        const wm::Image wME = watermarkObj.makeWatermark(image, image, a, MASK_TYPE::ME);
        std::printf("a_me %.9g\n", a);
        std::printf("corr_nvf %.9g\n", watermarkObj.detectWatermark(wNVF, MASK_TYPE::NVF));
        std::printf("corr_me %.9g\n", watermarkObj.detectWatermark(wME, MASK_TYPE::ME));
        // copies share W and work independently (Watermark.cpp:30-51)
        Watermark copy(watermarkObj);
        std::printf("corr_me_copy %.9g\n", copy.detectWatermark(wME, MASK_TYPE::ME));
        Watermark assigned(rows, cols, argv[2], 3, 30.0f);
        assigned = watermarkObj;
        std::printf("corr_me_assigned %.9g\n", assigned.detectWatermark(wME, MASK_TYPE::ME));
        // checksum of the ME-watermarked image (truncated to u8 like the reference's save path, main.cpp:236)
        std::vector<float> host((size_t)(rows * cols));
        wME.host(host.data());
        unsigned long long sum = 0;
        for (float v : host) sum += (unsigned)(unsigned char)v;
        std::printf("sum_u8_me %llu\n", sum);

        // video: frames with row padding, interval 2 (main.cpp:343-410)
        std::vector<unsigned char> frames = slurp<unsigned char>(argv[5], (size_t)nframes * rows * linesize);
        std::vector<unsigned char> uv((size_t)(rows / 2) * (cols / 2), 128);
        unsigned char* pinned = static_cast<unsigned char*>(wm_host_alloc_pinned(rows * cols));
        const VideoProcessingContext ctx(nullptr, nullptr, 0, &watermarkObj, (int)rows, (int)cols, 2, pinned);
        std::vector<unsigned char> marked((size_t)nframes * rows * cols);
        int next = 0;
        FILE* sink = std::tmpfile();
        const int n = processFrames(ctx, [&](wm::VideoFrame& f) {
                if (next >= nframes) return false;
                f.data[0] = frames.data() + (size_t)next * rows * linesize; f.linesize[0] = linesize;
                f.data[1] = f.data[2] = uv.data(); f.linesize[1] = f.linesize[2] = (int)cols / 2;
                f.height = (int)rows;
                next++;
                return true; },
            [&](wm::VideoFrame* f, int& count) { embedWatermarkFrame(ctx, count, f, sink); });
        std::printf("frames %d\n", n);
        // read back what went down the "pipe": Y, U, V per frame
        std::rewind(sink);
        const size_t ysz = (size_t)rows * cols, uvsz = uv.size();
        std::vector<unsigned char> skip(2 * uvsz);
        for (int i = 0; i < n; i++) {
            if (std::fread(marked.data() + i * ysz, 1, ysz, sink) != ysz || std::fread(skip.data(), 1, 2 * uvsz, sink) != 2 * uvsz) { std::cerr << "pipe short\n"; return 3; }
        }
        std::fclose(sink);
        for (int i = 0; i < n; i++) {
            unsigned long long s = 0;
            for (size_t k = 0; k < ysz; k++) s += marked[i * ysz + k];
            std::printf("frame_sum_%d %llu\n", i, s);
        }
        int count = 0;
        for (int i = 0; i < n; i++) {
            wm::VideoFrame f{};
            f.data[0] = marked.data() + i * ysz; f.linesize[0] = (int)cols; f.height = (int)rows;
            const float c = detectFrameWatermark(ctx, count, &f, false);
            std::printf("frame_corr_%d %.9g\n", i, c);
        }
        // the batched in-memory form over two contexts (two "GPUs": both on device 0 here) must reproduce the per-frame results
        {
            const Watermark second(watermarkObj);
            std::vector<unsigned char> marked2((size_t)nframes * rows * cols);
            std::vector<float> a2((size_t)nframes), c2((size_t)nframes);
            processFramesInMemory({&watermarkObj, &second}, (int)rows, (int)cols, 2, linesize, WM_VIDEO_EMBED, frames.data(), marked2.data(), 0, nframes, a2.data());
            processFramesInMemory({&watermarkObj, &second}, (int)rows, (int)cols, 2, (int)cols, WM_VIDEO_DETECT, marked2.data(), nullptr, 0, nframes, c2.data());
            std::printf("batched_equal_pixels %d\n", (int)(marked2 == marked));
            for (int i = 0; i < n; i++) std::printf("batched_corr_%d %.9g\n", i, c2[(size_t)i] == c2[(size_t)i] ? c2[(size_t)i] : 0.0f);
        }
        wm_host_free_pinned(pinned);
    } catch (const std::exception& e) {
        std::cerr << "exception: " << e.what();
        return 1;
    }
    return 0;
}
