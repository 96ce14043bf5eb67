// Issue-rate microbenchmark for the instructions the sweep / apply kernels lean on (sm_100a):
// FFMA, FFMA2 (fma.rn.f32x2), FMUL+F2FP+FHADD chain, HMUL2, FHADD (add.rn.f32.f16).
// Prints warp-instructions per clock per SM for each.  Build: nvcc -arch=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define UNROLL 16
template <int OP>
__global__ void __launch_bounds__(256) k(float* out, float seed)
{
    float a[UNROLL], b[UNROLL];
    for (int i = 0; i < UNROLL; i++) { a[i] = seed + i + threadIdx.x; b[i] = seed * 0.5f + i; }
    unsigned long long ua[UNROLL / 2];
    for (int i = 0; i < UNROLL / 2; i++) ua[i] = ((unsigned long long)__float_as_uint(a[2 * i]) << 32) | __float_as_uint(a[2 * i + 1]);
    unsigned long long sc = ((unsigned long long)__float_as_uint(1.0001f) << 32) | __float_as_uint(0.9999f);
    unsigned hh = 0x3c003c00u;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < UNROLL; i++) {
            if (OP == 0) a[i] = __fmaf_rn(a[i], 1.0001f, b[i]);                                      // FFMA
            if (OP == 1 && i < UNROLL / 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(ua[i]) : "l"(sc));  // FFMA2
            if (OP == 2) asm volatile("{ .reg .b16 h; mov.b32 {h, _}, %1; add.rn.f32.f16 %0, h, %0; }" : "+f"(a[i]) : "r"(hh));  // FHADD
            if (OP == 3) { unsigned x = __float_as_uint(a[i]); asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(x) : "r"(hh)); a[i] = __uint_as_float(x); }  // HMUL2
            if (OP == 4 && (i & 1) == 0) { unsigned h; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(a[i]), "f"(a[i + 1])); b[i] = __uint_as_float(h); }  // F2FP
            if (OP == 5) a[i] = __fadd_rn(a[i], b[i]);                                               // FADD
            if (OP == 6 && i < UNROLL / 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(ua[i]) : "l"(sc));      // FADD2
            if (OP == 7 && i < UNROLL / 2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(ua[i]) : "l"(sc));      // FMUL2
            if (OP == 8) { unsigned x = __float_as_uint(a[i]); asm volatile("cvt.rn.f32.u8 %0, %1;" : "=f"(a[i]) : "r"(x & 0xffu)); }              // I2F.U8 (+LOP)
            if (OP == 9) { unsigned x = __float_as_uint(a[i]); x = __byte_perm(x, 0x4b000000u, 0x7440); a[i] = __fadd_rn(__uint_as_float(x), -8388608.0f); }  // PRMT + FADD: byte -> float
            if (OP == 10) { unsigned x = __float_as_uint(a[i]); asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(x) : "r"(0x00010001u)); a[i] = __uint_as_float(x); }  // HMUL2, subnormal operand
            if (OP == 11) { unsigned x = __float_as_uint(a[i]); asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(hh), "r"(0x80008000u)); a[i] = __uint_as_float(x); }  // HFMA2
            if (OP == 12) { unsigned x = __float_as_uint(a[i]); if (i & 1) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(hh), "r"(0x80008000u)); else asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(x) : "r"(hh)); a[i] = __uint_as_float(x); }  // HMUL2 / HFMA2 alternating
            if (OP == 13) { unsigned x = __float_as_uint(a[i]); x = __byte_perm(x, __float_as_uint(b[i]), 0x5432); a[i] = __uint_as_float(x); }  // PRMT
        }
    }
    float s = 0;
    for (int i = 0; i < UNROLL; i++) s += a[i] + b[i];
    for (int i = 0; i < UNROLL / 2; i++) s += __uint_as_float((unsigned)ua[i]) + __uint_as_float((unsigned)(ua[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char* name, int per_iter, float* d, int sms, int clk_khz)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 4;
    k<OP><<<blocks, 256>>>(d, 1.0f);
    cudaEventRecord(e0);
    k<OP><<<blocks, 256>>>(d, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double winst = (double)blocks * 8 * ITERS * per_iter;
    const double clocks = ms * 1e-3 * clk_khz * 1e3;
    printf("%-8s %8.3f ms  %.2f warp-instr/clk/SM (at %d MHz)\n", name, ms, winst / clocks / sms, clk_khz / 1000);
}
int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* d;
    cudaMalloc(&d, p.multiProcessorCount * 4 * 256 * 4);
    run<0>("FFMA", UNROLL, d, p.multiProcessorCount, clk);
    run<1>("FFMA2", UNROLL / 2, d, p.multiProcessorCount, clk);
    run<5>("FADD", UNROLL, d, p.multiProcessorCount, clk);
    run<6>("FADD2", UNROLL / 2, d, p.multiProcessorCount, clk);
    run<2>("FHADD", UNROLL, d, p.multiProcessorCount, clk);
    run<3>("HMUL2", UNROLL, d, p.multiProcessorCount, clk);
    run<4>("F2FP", UNROLL / 2, d, p.multiProcessorCount, clk);
    run<7>("FMUL2", UNROLL / 2, d, p.multiProcessorCount, clk);
    run<8>("I2F.U8", UNROLL, d, p.multiProcessorCount, clk);
    run<9>("PRMT+FADD", UNROLL, d, p.multiProcessorCount, clk);
    run<10>("HMUL2sub", UNROLL, d, p.multiProcessorCount, clk);
    run<11>("HFMA2", UNROLL, d, p.multiProcessorCount, clk);
    run<12>("HMUL2/HFMA2", UNROLL, d, p.multiProcessorCount, clk);
    run<13>("PRMT", UNROLL, d, p.multiProcessorCount, clk);
    return 0;
}
