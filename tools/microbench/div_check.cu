// Exhaustive check on the device of the branch-free division used by the NVF mask: q = var / (1 + var) for EVERY float
// var in [-2^-6, 2^17) (the naive variance of 0..255 pixels lies in about [-0.01, 16257]), against IEEE __fdiv_rn.
// Variants: 0-2 Newton refinements of rcp.approx, 1-2 residual corrections of the quotient.  Prints mismatch counts.
#include <cstdio>
#include <cuda_runtime.h>
template <int YREF, int QREF>
__device__ __forceinline__ float div_safe(float n, float d)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(d));
    if (YREF >= 1) y = __fmaf_rn(__fmaf_rn(-d, y, 1.0f), y, y);
    if (YREF >= 2) y = __fmaf_rn(__fmaf_rn(-d, y, 1.0f), y, y);
    float q = __fmul_rn(n, y);
    if (QREF >= 1) q = __fmaf_rn(__fmaf_rn(-d, q, n), y, q);
    if (QREF >= 2) q = __fmaf_rn(__fmaf_rn(-d, q, n), y, q);
    return q;
}
__global__ void check(unsigned lo, unsigned hi, int negative, unsigned long long* bad)
{
    unsigned long long b[5] = {0, 0, 0, 0, 0};
    for (unsigned long long u = lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; u < hi; u += (unsigned long long)gridDim.x * blockDim.x) {
        float var = __uint_as_float((unsigned)u | (negative ? 0x80000000u : 0u));
        const float d = __fadd_rn(1.0f, var);
        const float t = __fdiv_rn(var, d);
        if (div_safe<2, 2>(var, d) != t) b[0]++;
        if (div_safe<1, 2>(var, d) != t) b[1]++;
        if (div_safe<1, 1>(var, d) != t) b[2]++;
        if (div_safe<0, 2>(var, d) != t) b[3]++;
        if (div_safe<0, 1>(var, d) != t) b[4]++;
    }
    for (int i = 0; i < 5; i++) if (b[i]) atomicAdd(bad + i, b[i]);
}
int main()
{
    unsigned long long *d, h[5] = {0, 0, 0, 0, 0};
    cudaMalloc(&d, 40);
    cudaMemset(d, 0, 40);
    const unsigned hi_pos = 0x48000000u;  // 2^17
    const unsigned hi_neg = 0x3c800000u;  // 2^-6
    check<<<148 * 8, 256>>>(0u, hi_pos, 0, d);        // all non-negative floats (incl. denormals) below 2^17
    check<<<148 * 8, 256>>>(0u, hi_neg, 1, d);        // all negative floats above -2^-6
    cudaMemcpy(h, d, 40, cudaMemcpyDeviceToHost);
    printf("values checked: %llu ; mismatches vs __fdiv_rn for (reciprocal refinements, quotient refinements): "
           "(2,2) %llu  (1,2) %llu  (1,1) %llu  (0,2) %llu  (0,1) %llu\n",
           (unsigned long long)hi_pos + hi_neg, h[0], h[1], h[2], h[3], h[4]);
    return cudaGetLastError() != cudaSuccess;
}
