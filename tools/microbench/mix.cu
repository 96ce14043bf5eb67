// Pipe-mix microbenchmark for the Rx sweep's inner unit (sm_100a).  One "unit" = 8 rounded products (2 pixels x 4 lags):
//   8 FMUL + 4 F2FP.PACK + accumulation (8 FHADD | 1 HMMA | none), or for u8 frames 4 HMUL2 + accumulation.
// All operands live in registers (static indexing only), every instruction is an asm volatile so nothing is hoisted or
// folded; 6 independent units per loop iteration give the ILP the real kernel has.  Prints clocks per unit per scheduler
// at 24 warps/SM (the sweep's occupancy).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mix mix.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define NU 6
#define FMUL(d, a, b) asm volatile("mul.rn.f32 %0, %1, %2; // %3" : "=f"(d) : "f"(a), "f"(b), "n"(__COUNTER__))
#define PACK(h, lo, hi) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2; // %3" : "=r"(h) : "f"(hi), "f"(lo), "n"(__COUNTER__))
#define HMUL2(h, a, b) asm volatile("mul.rn.f16x2 %0, %1, %2; // %3" : "=r"(h) : "r"(a), "r"(b), "n"(__COUNTER__))
#define FHADD2(c0, c1, h) asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tadd.rn.f32.f16 %0, lo, %0;\n\tadd.rn.f32.f16 %1, hi, %1;\n\t}" : "+f"(c0), "+f"(c1) : "r"(h))
#define HMMA(c, h0, h1, h2, h3) asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};" \
    : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(h0), "r"(h1), "r"(h2), "r"(h3), "r"(b0), "r"(b1))
#define XORACC(s, h) asm volatile("xor.b32 %0, %0, %1;" : "+r"(s) : "r"(h))

// MUL: 0 none, 1 = 8 FMUL + 4 F2FP, 2 = 4 HMUL2, 3 = 8 FMUL only, 4 = 4 F2FP only;  ACC: 0 none, 1 = 8 FHADD, 2 = 1 HMMA, 3 = 4 XOR (sink)
template <int MUL, int ACC>
__global__ void __launch_bounds__(256, 3) k(float* out, const float* in)
{
    float x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = in[threadIdx.x + i]; y[i] = in[threadIdx.x + 8 + i]; }
    unsigned hx[4], hy[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { PACK(hx[i], x[2 * i], x[2 * i + 1]); PACK(hy[i], y[2 * i], y[2 * i + 1]); }
    unsigned hyy[NU][4];  // distinct partner operands, so that no two HMUL2 are the same expression
#pragma unroll
    for (int u = 0; u < NU; u++)
#pragma unroll
        for (int i = 0; i < 4; i++) hyy[u][i] = hy[i] + 0x00010001u * (unsigned)(u + 1);
    float c[NU][8];
#pragma unroll
    for (int u = 0; u < NU; u++)
#pragma unroll
        for (int i = 0; i < 8; i++) c[u][i] = 0.f;
    const unsigned b0 = (threadIdx.x & 4) ? 0x3c003c00u : 0u, b1 = (threadIdx.x & 8) ? 0x3c003c00u : 0u;
    unsigned sink = 0;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        // operands change every iteration (12 cheap instructions per 6 units), so no product can be hoisted out of the loop
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(1e-3f));
#pragma unroll
        for (int i = 0; i < 4; i++) asm volatile("xor.b32 %0, %0, %1;" : "+r"(hx[i]) : "r"(it & 1));
#pragma unroll
        for (int u = 0; u < NU; u++) {
            unsigned h[4] = {hx[0], hx[1], hx[2], hx[3]};
            float p[8];
            if (MUL == 1 || MUL == 3) {
#pragma unroll
                for (int i = 0; i < 8; i++) FMUL(p[i], x[i], y[(i + u) & 7]);
            }
            if (MUL == 1) {
#pragma unroll
                for (int i = 0; i < 4; i++) PACK(h[i], p[2 * i], p[2 * i + 1]);
            }
            if (MUL == 3) {
#pragma unroll
                for (int i = 0; i < 4; i++) h[i] = __float_as_uint(p[2 * i]) ^ __float_as_uint(p[2 * i + 1]);
            }
            if (MUL == 4) {
#pragma unroll
                for (int i = 0; i < 4; i++) PACK(h[i], x[2 * i], y[(2 * i + u) & 7]);
            }
            if (MUL == 2) {
#pragma unroll
                for (int i = 0; i < 4; i++) HMUL2(h[i], hx[i], hyy[u][i]);
            }
            if (ACC == 1) {
#pragma unroll
                for (int i = 0; i < 4; i++) FHADD2(c[u][2 * i], c[u][2 * i + 1], h[i]);
            }
            if (ACC == 2) HMMA(c[u], h[0], h[1], h[2], h[3]);
            if (ACC == 3) {
#pragma unroll
                for (int i = 0; i < 4; i++) XORACC(sink, h[i]);
            }
        }
    }
    float s = __uint_as_float(sink);
#pragma unroll
    for (int u = 0; u < NU; u++)
#pragma unroll
        for (int i = 0; i < 8; i++) s += c[u][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MUL, int ACC>
void run(const char* name, float* d, const float* in, int sms, int clk_khz)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 3;
    for (int w = 0; w < 3; w++) k<MUL, ACC><<<blocks, 256>>>(d, in);
    cudaEventRecord(e0);
    k<MUL, ACC><<<blocks, 256>>>(d, in);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double units = (double)blocks * 8 * ITERS * NU;  // warp-units
    const double clocks = ms * 1e-3 * clk_khz * 1e3;
    printf("%-28s %8.3f ms  %6.2f clk per unit per scheduler (%s)\n", name, ms, 4.0 * clocks * sms / units, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float *d, *in;
    cudaMalloc(&d, p.multiProcessorCount * 3 * 256 * 4);
    cudaMalloc(&in, 512 * 4);
    float h[512];
    for (int i = 0; i < 512; i++) h[i] = 1.0f + i * 0.37f;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    const int s = p.multiProcessorCount;
    // spin the clocks up
    for (int w = 0; w < 20; w++) k<1, 1><<<s * 3, 256>>>(d, in);
    cudaDeviceSynchronize();
    run<3, 3>("8 FMUL (+4 LOP +4 XOR)", d, in, s, clk);
    run<4, 3>("4 F2FP (+4 XOR)", d, in, s, clk);
    run<0, 2>("1 HMMA", d, in, s, clk);
    run<0, 1>("8 FHADD", d, in, s, clk);
    run<2, 3>("4 HMUL2 (+4 XOR)", d, in, s, clk);
    run<1, 3>("8 FMUL + 4 F2FP (+4 XOR)", d, in, s, clk);
    run<1, 1>("8 FMUL + 4 F2FP + 8 FHADD", d, in, s, clk);
    run<1, 2>("8 FMUL + 4 F2FP + 1 HMMA", d, in, s, clk);
    run<2, 1>("4 HMUL2 + 8 FHADD", d, in, s, clk);
    run<2, 2>("4 HMUL2 + 1 HMMA", d, in, s, clk);
    return 0;
}
