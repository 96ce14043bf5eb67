// Watermark.hpp — C++ façade with the reference's class surface (Watermark_GPU/Watermark.hpp:26-71) over the
// C ABI of include/wm_b200.h.  Header-only; link against libwm_b200.so.
//
// What changes for a caller of the reference class:
//   * af::array arguments become wm::Image — a ref-counted column-major device array with the same
//     (rows, cols[, 3]) f32 convention.  ArrayFire arrays enter through device-pointer interop only:
//         wm::Image::view(arr.device<float>(), arr.dims(0), arr.dims(1), arr.dims(2))   // then arr.unlock()
//   * the `programs` vector of JIT-built cl::Program objects (main.cpp:99-108) is gone: kernels are compiled
//     ahead of time for sm_100a.
//   * everything else — constructor arguments, copy semantics (copies share W and own their workspace, move is
//     deleted), reinitialize, makeWatermark's float& watermarkStrength, detectWatermark's float result, the
//     std::runtime_error messages, the "unsolvable system" fall-backs — is as in Watermark.cpp.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>

#include "wm_b200.h"

using dim_t = long long;  // ArrayFire's dim_t

enum MASK_TYPE  // Watermark.hpp:10-14
{
    ME,
    NVF
};

struct dim2  // Watermark.hpp:16-20
{
    dim_t rows;
    dim_t cols;
};

namespace wm {

// Column-major (rows, cols[, channels]) device array, f32 or u8.  Owning instances free their memory when the
// last copy goes away (like af::array); view() wraps foreign device memory (ArrayFire interop) without owning it.
class Image {
public:
    Image() = default;
    Image(dim_t rows, dim_t cols, int channels = 1, int dtype = WM_F32, int layout = WM_COL_MAJOR)
    {
        init(rows, cols, channels, dtype, layout);
        const int64_t bytes = rows * cols * channels * (dtype == WM_F32 ? 4 : 1);
        void* p = wm_dev_alloc(nullptr, bytes);
        if (!p) throw std::runtime_error("wm::Image: device allocation of " + std::to_string(bytes) + " bytes failed\n");
        mem_ = std::shared_ptr<void>(p, [](void* q) { wm_dev_free(nullptr, q); });
        d_.data = p;
    }
    // like af::array(rows, cols, hostPtr): uploads a column-major host buffer
    Image(dim_t rows, dim_t cols, const float* host, int channels = 1) : Image(rows, cols, channels, WM_F32)
    {
        if (wm_dev_upload(nullptr, d_.data, host, bytes())) throw std::runtime_error("wm::Image: upload failed\n");
    }
    static Image view(void* device_ptr, dim_t rows, dim_t cols, int channels = 1, int dtype = WM_F32, int layout = WM_COL_MAJOR,
                      dim_t ld = 0)
    {
        Image im;
        im.init(rows, cols, channels, dtype, layout);
        im.d_.data = device_ptr;
        im.d_.ld = ld;
        return im;
    }
    dim_t dims(int i) const { return i == 0 ? d_.rows : (i == 1 ? d_.cols : (i == 2 ? d_.channels : 1)); }
    dim_t elements() const { return d_.data ? d_.rows * d_.cols * d_.channels : 0; }
    bool isempty() const { return d_.data == nullptr; }
    int64_t bytes() const { return elements() * (d_.dtype == WM_F32 ? 4 : 1); }
    template <typename T> T* device() const { return static_cast<T*>(d_.data); }
    void host(void* dst) const
    {
        if (wm_dev_download(nullptr, dst, d_.data, bytes())) throw std::runtime_error("wm::Image: download failed\n");
    }
    const wm_image* desc() const { return &d_; }
    wm_image* desc() { return &d_; }

private:
    void init(dim_t rows, dim_t cols, int channels, int dtype, int layout)
    {
        d_ = wm_image{};
        d_.rows = rows; d_.cols = cols; d_.channels = channels; d_.dtype = dtype; d_.layout = layout;
    }
    wm_image d_{};
    std::shared_ptr<void> mem_;
};

}  // namespace wm

/*!
 *  \brief  Functions for watermark computation and detection (B200 path).  Mirrors the reference class.
 */
class Watermark {
public:
    // Watermark.cpp:21-27 (the cl::Program vector is dropped; `device` / `stream` are optional extras)
    Watermark(const dim_t rows, const dim_t cols, const std::string& randomMatrixPath, const int p, const float psnr,
              const int device = 0, void* cudaStream = nullptr)
        : dims({rows, cols}), p(p), psnr(psnr)
    {
        if (p != 3 && p != 5 && p != 7 && p != 9)
            throw std::runtime_error(std::string("Wrong p parameter: ") + std::to_string(p) + "!\n");
        check(wm_create_from_file(&ctx, rows, cols, randomMatrixPath.c_str(), p, psnr, device, cudaStream), nullptr);
    }
    // variant taking W from host memory (row-major rows x cols, the file's layout)
    Watermark(const dim_t rows, const dim_t cols, const float* randomMatrixRowMajor, const int p, const float psnr,
              const int device = 0, void* cudaStream = nullptr)
        : dims({rows, cols}), p(p), psnr(psnr)
    {
        check(wm_create(&ctx, rows, cols, randomMatrixRowMajor, p, psnr, device, cudaStream), nullptr);
    }
    // copy constructor (Watermark.cpp:30-34): shares W, owns a new workspace
    Watermark(const Watermark& other) : dims(other.dims), p(other.p), psnr(other.psnr) { check(wm_clone(other.ctx, &ctx), nullptr); }
    Watermark(Watermark&& other) noexcept = delete;
    Watermark& operator=(Watermark&& other) noexcept = delete;
    // copy assignment (Watermark.cpp:37-51)
    Watermark& operator=(const Watermark& other)
    {
        if (this != &other) {
            wm_ctx* fresh = nullptr;
            check(wm_clone(other.ctx, &fresh), nullptr);
            wm_destroy(ctx);
            ctx = fresh;
            dims = other.dims; p = other.p; psnr = other.psnr;
        }
        return *this;
    }
    ~Watermark() { wm_destroy(ctx); }

    // Watermark.cpp:78-85
    void reinitialize(const std::string& randomMatrixPath, const dim_t rows, const dim_t cols)
    {
        check(wm_reinitialize_from_file(ctx, rows, cols, randomMatrixPath.c_str()), ctx);
        dims = {rows, cols};
    }

    // Watermark.cpp:156-172.  inputImage: gray; outputImage: gray or RGB (the array the watermark is added to).
    // Unsolvable system: returns outputImage unchanged and leaves watermarkStrength untouched (Watermark.cpp:164-165).
    wm::Image makeWatermark(const wm::Image& inputImage, const wm::Image& outputImage, float& watermarkStrength, MASK_TYPE maskType) const
    {
        wm::Image out(outputImage.dims(0), outputImage.dims(1), (int)outputImage.dims(2), outputImage.desc()->dtype,
                      outputImage.desc()->layout);
        const int rc = check(wm_embed(ctx, inputImage.desc(), outputImage.desc(), out.desc(), (int)maskType, &watermarkStrength), ctx);
        if (rc == WM_SINGULAR) return outputImage;
        return out;
    }

    // Watermark.cpp:234-250.  Unsolvable system: 0.0f (Watermark.cpp:246-247).
    float detectWatermark(const wm::Image& watermarkedImage, MASK_TYPE maskType) const
    {
        float corr = 0.0f;
        check(wm_detect(ctx, watermarkedImage.desc(), (int)maskType, &corr), ctx);
        return corr;
    }

    wm_ctx* handle() const { return ctx; }  // for the batched / video entry points of the C ABI

private:
    static int check(int rc, const wm_ctx* c)
    {
        if (rc < 0) throw std::runtime_error(std::string(wm_last_error(c)) + "\n");
        return rc;
    }
    dim2 dims;
    int p;
    float psnr;
    wm_ctx* ctx = nullptr;
};
