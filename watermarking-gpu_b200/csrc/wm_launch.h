// wm_launch.h — launch dispatchers, one translation unit per kernel family so the build compiles them in parallel.
#pragma once
#include "wm_kernels.cuh"

namespace wm {
void launch_sweep(int dtype, int acc, bool tma, dim3 grid, cudaStream_t st, const CUtensorMap& tmI, const SweepArgs& a);
void launch_stats(int dtype, int mask, bool tr, bool tma, dim3 grid, cudaStream_t st, const CUtensorMap& tmI, const CUtensorMap& tmW, const EmbedArgs& a);
void launch_apply(int in_dtype, int out_dtype, int mask, bool tr, bool tma, dim3 grid, cudaStream_t st, const CUtensorMap& tmI,
                  const CUtensorMap& tmW, const EmbedArgs& a);
void launch_apply_ts(int dtype, int mask, bool tr, dim3 grid, cudaStream_t st, const CUtensorMap& tmI, const CUtensorMap& tmW, const CUtensorMap& tmO,
                     const EmbedArgs& a);
void launch_detect(int dtype, int mask, bool tr, bool tma, dim3 grid, cudaStream_t st, const CUtensorMap& tmZ, const CUtensorMap& tmW, const DetectArgs& a);
void launch_nvfp(int dtype, int pw, bool tr, dim3 grid, cudaStream_t st, const NvfpArgs& a);
void launch_plane(int dtype, int what_errseq, bool tr, dim3 grid, cudaStream_t st, const PlaneArgs& a);
void launch_mask_from_errseq(float* plane, long long n, unsigned* scratch, cudaStream_t st);
void launch_transpose(const float* src, float* dst, int rows, int cols, cudaStream_t st);
void launch_rgb2gray(const float* r, const float* g, const float* b, float* gray, long long ld_in, long long ld_out, int L, int P,
                     float wr, float wg, float wb, int blocks, cudaStream_t st);
}  // namespace wm

// dtype / mask codes shared with include/wm_b200.h
#ifndef WM_B200_H
enum { WM_MASK_ME = 0, WM_MASK_NVF = 1 };
enum { WM_F32 = 0, WM_U8 = 1 };
#endif

#define WM_LAUNCH_T(KERNEL, THREADS, SMEM, ...)                        \
    do {                                                               \
        static bool done_[64] = {false};                               \
        int dev_ = 0;                                                  \
        cudaGetDevice(&dev_);                                          \
        if (!done_[dev_ & 63]) { cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM); done_[dev_ & 63] = true; } \
        KERNEL<<<grid, THREADS, SMEM, st>>>(__VA_ARGS__);              \
    } while (0)
#define WM_LAUNCH(KERNEL, SMEM, ...)                                   \
    do {                                                               \
        static bool done_[64] = {false};                               \
        int dev_ = 0;                                                  \
        cudaGetDevice(&dev_);                                          \
        if (!done_[dev_ & 63]) { cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM); done_[dev_ & 63] = true; } \
        KERNEL<<<grid, NT, SMEM, st>>>(__VA_ARGS__);                   \
    } while (0)
