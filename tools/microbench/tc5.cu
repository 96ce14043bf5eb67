// tcgen05 accumulation microbenchmark for the Rx sweep's inner unit (sm_100a).
//
// Question: the sweep is issue-bound; per 8 fp16-rounded products it spends 8 FMUL + 4 F2FP.PACK plus
// 8 FHADD slots (or one legacy HMMA that blocks dispatch for about as long).  Can the accumulation leave the SM's issue slots
// altogether?  Here every producer thread stores its packed halves straight into TENSOR MEMORY (tcgen05.st, lane = thread) as
// the A operand of tcgen05.mma.kind::f16 (M = 128, N = 16, K = 16), B = 0/1 selector matrices in shared memory that map
// product slot k -> lag column n, D (128 x 16 f32) in tensor memory.  One elected thread of a ninth warp issues the MMAs.
//
// Protocol per "half tile" (52 packed registers = 104 products per thread, the sweep's 2 lines x 4 pixels x 13 lags):
//   producers: wait empty[g][s] -> 52 registers by tcgen05.st (x32 + x16 + x4) -> tcgen05.wait::st -> fence -> arrive full[g][s]
//   MMA warp : wait full[g][s] -> fence -> 7 x tcgen05.mma (K blocks of 16 halves; 8 pad halves stay zero) -> tcgen05.commit -> empty[g][s]
// g = warpgroup (warps 0-3 / 4-7 own the same 128 TMEM lanes, different columns), s = double buffer.
//
// Modes: MUL 1 = 8 FMUL + 4 F2FP per unit, 2 = 4 HMUL2 (u8 frames), 5 = 4 FMUL2 (mul.rn.f32x2) + 4 F2FP, 0 = verification pattern
//        ACC 1 = 8 FHADD, 2 = 1 legacy HMMA, 3 = XOR sink (no accumulation), 4 = tcgen05
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc5 tc5.cu ; run: ./tc5 [mode ...]   (each mode may be run in
// its own process: a protocol bug traps instead of hanging).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

#define NU 13           // units (8 products = 4 packed registers) per half tile
#define NREG (4 * NU)   // 52
#define KBLK 7          // K blocks of 16 halves (8 registers): 56 registers, the last 4 stay zero
#define NPROD 256       // producer threads
#define NTHR 288        // + the MMA warp
#define TCOLS 256       // TMEM columns per CTA (two CTAs per SM)
#define D_COL 0
#define A_COL(g, s) (32 + ((g) * 2 + (s)) * 56)

#define FMUL(d, a, b) asm volatile("mul.rn.f32 %0, %1, %2; // %3" : "=f"(d) : "f"(a), "f"(b), "n"(U))
#define FMUL2(d, a, b) asm volatile("mul.rn.f32x2 %0, %1, %2; // %3" : "=l"(d) : "l"(a), "l"(b), "n"(U))
#define PACK(h, lo, hi) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2; // %3" : "=r"(h) : "f"(hi), "f"(lo), "n"(U))
#define HMUL2(h, a, b) asm volatile("mul.rn.f16x2 %0, %1, %2; // %3" : "=r"(h) : "r"(a), "r"(b), "n"(U))
#define FHADD2(c0, c1, h) asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tadd.rn.f32.f16 %0, lo, %0;\n\tadd.rn.f32.f16 %1, hi, %1;\n\t}" : "+f"(c0), "+f"(c1) : "r"(h))
#define HMMA(c, h0, h1, h2, h3) asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};" \
    : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(h0), "r"(h1), "r"(h2), "r"(h3), "r"(b0), "r"(b1))

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ int g_err;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    long long spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (!ok && ++spins > 20000000LL) { g_err = 1; __trap(); }  // a protocol bug must not hang the box
    } while (!ok);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem desc]
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_st4(uint32_t taddr, const unsigned* r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const unsigned* r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const unsigned* r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
                 "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, unsigned* r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
}
// selector: product slot (register i of the half tile, half h) -> lag column.  The verification pattern uses an asymmetric map.
__host__ __device__ inline int lag_of(int i, int h) { return (i * 5 + h * 3) % 13; }
// verification pattern: small integers, so every sum is exact
__host__ __device__ inline int pat(int tid, int it, int i, int h) { return (tid * 7 + it * 3 + i * 5 + h) & 31; }

// smem descriptor of one 16 (n) x 16 (k) fp16 selector, K-major, no swizzle: core matrix = 8 n-rows x 16 bytes (8 k), 128 B;
// k-group stride `lbo`, n-group stride `sbo` (bytes)
__device__ __forceinline__ uint64_t make_bdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version of sm_100
    return d;
}

// one unit = 8 rounded products -> 4 packed registers h[4U .. 4U+3] (+ accumulation for the register-accumulating modes)
template <int MUL, int ACC, int U>
__device__ __forceinline__ void unit(unsigned (&h)[NREG], const float (&x)[8], const float (&y)[16], const unsigned (&hx)[4], const unsigned (&hy)[16],
                                     const unsigned long long (&xx)[4], const unsigned long long (&yy)[16], float (&c)[2][8], unsigned& sink,
                                     unsigned b0, unsigned b1, int tid, int it)
{
    constexpr int u = U;
    unsigned* hu = h + 4 * u;
    if (MUL == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const __half2 v = __floats2half2_rn((float)pat(tid, it, 4 * u + i, 0), (float)pat(tid, it, 4 * u + i, 1));
            hu[i] = *reinterpret_cast<const unsigned*>(&v);
        }
    }
    if (MUL == 1) {
        float p[8];
#pragma unroll
        for (int i = 0; i < 8; i++) FMUL(p[i], x[i], y[(u >> 3) * 8 + ((i + u) & 7)]);  // distinct operand pairs per unit: ptxas merges identical products
#pragma unroll
        for (int i = 0; i < 4; i++) PACK(hu[i], p[2 * i], p[2 * i + 1]);
    }
    if (MUL == 5) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            unsigned long long p2;
            FMUL2(p2, xx[i], yy[(u >> 2) * 4 + ((i + u) & 3)]);
            float lo, hi;
            asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p2));
            PACK(hu[i], lo, hi);
        }
    }
    if (MUL == 2) {
#pragma unroll
        for (int i = 0; i < 4; i++) HMUL2(hu[i], hx[i], hy[(u >> 2) * 4 + ((i + u) & 3)]);
    }
    if (ACC == 1) {
#pragma unroll
        for (int i = 0; i < 4; i++) FHADD2(c[u & 1][2 * i], c[u & 1][2 * i + 1], hu[i]);
    }
    if (ACC == 2) HMMA(c[u & 1], hu[0], hu[1], hu[2], hu[3]);
    if (ACC == 3) {
#pragma unroll
        for (int i = 0; i < 4; i++) asm volatile("xor.b32 %0, %0, %1;" : "+r"(sink) : "r"(hu[i]));
    }
}

template <int MUL, int ACC>
__global__ void __launch_bounds__(NTHR, 2) k(float* out, const float* in, int iters, int swap_lbo, float* dout, int nmma)
{
    __shared__ __align__(128) __half Bsel[KBLK][2][2][8][8];  // [k block][n group][k group][n row][k]
    __shared__ __align__(8) uint64_t full[2][2], empty[2][2], done;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (ACC == 4) {
        for (int e = tid; e < KBLK * 256; e += NTHR) {
            const int b = e >> 8, n = (e >> 4) & 15, kk = e & 15;
            const int i = b * 8 + (kk >> 1), h = kk & 1;
            const float v = (i < NREG && lag_of(i, h) == n) ? 1.0f : 0.0f;
            Bsel[b][n >> 3][kk >> 3][n & 7][kk & 7] = __float2half(v);
        }
        if (tid == 0) {
            for (int g = 0; g < 2; g++)
                for (int s = 0; s < 2; s++) { mbar_init(&full[g][s], 4); mbar_init(&empty[g][s], 1); }
            mbar_init(&done, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (w == 8) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TCOLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // Bsel is read by the tensor core (async proxy)
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    const uint32_t tbase = ACC == 4 ? tmem_base_s : 0;
    if (ACC == 4 && w == 8) {
        // ---------------- MMA warp ----------------
        const uint32_t idesc = (1u << 4) | (2u << 17) | (8u << 24);  // D f32, A/B f16 K-major, N = 16, M = 128
        uint64_t bdesc[KBLK];
#pragma unroll
        for (int b = 0; b < KBLK; b++) bdesc[b] = make_bdesc(smem_u32(&Bsel[b][0][0][0][0]), swap_lbo ? 256 : 128, swap_lbo ? 128 : 256);
        // wait for the producers to have zeroed D and the pad columns
        mbar_wait(&done, 0);
        tc_fence_after();
        for (int it = 0; it < iters; it++) {
            const int s = it & 1;
            const unsigned ph = (it >> 1) & 1;
#pragma unroll
            for (int g = 0; g < 2; g++) {
                mbar_wait(&full[g][s], ph);
                tc_fence_after();
                if (lane == 0) {
#pragma unroll
                    for (int b = 0; b < KBLK; b++)
                        if (b < nmma) tc_mma_ts(tbase + D_COL, tbase + A_COL(g, s) + 8 * b, bdesc[b], idesc, 1u);
                    if (nmma > 0) tc_commit(&empty[g][s]); else mbar_arrive(&empty[g][s]);
                }
                __syncwarp();
            }
        }
        // all MMAs done -> done barrier (phase 1)
        if (lane == 0) tc_commit(&done);
        __syncwarp();
    } else if (w < 8) {
        // ---------------- producers ----------------
        const int g = w >> 2;
        const uint32_t lane_base = (uint32_t)((w & 3) * 32) << 16;
        if (ACC == 4) {
            unsigned z[16];
#pragma unroll
            for (int i = 0; i < 16; i++) z[i] = 0u;
            if (g == 0) { tc_st16(tbase + lane_base + D_COL, z); tc_st16(tbase + lane_base + D_COL + 16, z); }
            tc_st4(tbase + lane_base + A_COL(g, 0) + 52, z);
            tc_st4(tbase + lane_base + A_COL(g, 1) + 52, z);
            tc_wait_st();
            tc_fence_before();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid == 0) mbar_arrive(&done);  // phase 0 of `done`: TMEM initialised
        }
        float x[8], y[16];
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = in[tid + i];
#pragma unroll
        for (int i = 0; i < 16; i++) y[i] = in[tid + 8 + i];
        constexpr int U = 100;  // tag of the set-up packs
        unsigned hx[4], hy[16];
#pragma unroll
        for (int i = 0; i < 4; i++) PACK(hx[i], x[2 * i], x[2 * i + 1]);
#pragma unroll
        for (int i = 0; i < 16; i++) PACK(hy[i], y[i], in[tid + 32 + i]);
        unsigned long long xx[4], yy[16];
#pragma unroll
        for (int i = 0; i < 4; i++) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(xx[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
#pragma unroll
        for (int i = 0; i < 16; i++) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(yy[i]) : "f"(y[i]), "f"(in[tid + 64 + i]));
        float c[2][8];
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) c[u][i] = 0.f;
        const unsigned b0 = (tid & 4) ? 0x3c003c00u : 0u, b1 = (tid & 8) ? 0x3c003c00u : 0u;
        unsigned sink = 0;
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
            const int s = it & 1;
            // operands change every iteration so that no product can be hoisted (the real kernel loads a new window here)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(1e-3f));
#pragma unroll
            for (int i = 0; i < 4; i++) asm volatile("xor.b32 %0, %0, %1;" : "+r"(hx[i]) : "r"(it & 1));
            if (MUL == 5) {
#pragma unroll
                for (int i = 0; i < 4; i++) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(xx[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
            }
            unsigned h[NREG];
#define UNIT(Uv) unit<MUL, ACC, Uv>(h, x, y, hx, hy, xx, yy, c, sink, b0, b1, tid, it);
            UNIT(0) UNIT(1) UNIT(2) UNIT(3) UNIT(4) UNIT(5) UNIT(6) UNIT(7) UNIT(8) UNIT(9) UNIT(10) UNIT(11) UNIT(12)
#undef UNIT
            if (ACC == 4) {
                if (it >= 2) mbar_wait(&empty[g][s], ((it >> 1) - 1) & 1);
                tc_fence_after();
                const uint32_t a = tbase + lane_base + A_COL(g, s);
                tc_st32(a, h);
                tc_st16(a + 32, h + 32);
                tc_st4(a + 48, h + 48);
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[g][s]);
            }
        }
        float sacc = __uint_as_float(sink);
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) sacc += c[u][i];
        out[blockIdx.x * NPROD + tid] = sacc;
        if (ACC == 4) {
            mbar_wait(&done, 1);
            tc_fence_after();
            if (g == 0) {
                unsigned d[16];
                tc_ld16(tbase + lane_base + D_COL, d);
                tc_wait_ld();
                if (dout)
#pragma unroll
                    for (int n = 0; n < 16; n++) dout[((size_t)blockIdx.x * 128 + (w & 3) * 32 + lane) * 16 + n] = __uint_as_float(d[n]);
            }
        }
    }
    if (ACC == 4) {
        tc_fence_before();
        __syncthreads();
        if (w == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(TCOLS) : "memory");
    }
}

static int g_sms, g_clk, g_nmma = KBLK;
template <int MUL, int ACC>
double run(const char* name, float* d, const float* in, float* dout)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = g_sms * 2, iters = 2048;
    for (int w = 0; w < 2; w++) k<MUL, ACC><<<blocks, NTHR>>>(d, in, iters, 0, nullptr, g_nmma);
    cudaEventRecord(e0);
    k<MUL, ACC><<<blocks, NTHR>>>(d, in, iters, 0, nullptr, g_nmma);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double units = (double)blocks * 8 * iters * NU;  // warp-units
    const double clocks = ms * 1e-3 * g_clk * 1e3;
    const double cpu = 4.0 * clocks * g_sms / units;
    printf("%-34s %8.3f ms  %6.2f clk per unit per scheduler (%s)\n", name, ms, cpu, cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
    (void)dout;
    return cpu;
}

// numerical check of the TMEM-A path: D[lane][n] must equal the sum of the pattern's products mapped to lag n
static int verify(float* d, const float* in, float* dout, int swap)
{
    const int blocks = 4, iters = 6;
    cudaMemset(dout, 0, sizeof(float) * blocks * 128 * 16);
    k<0, 4><<<blocks, NTHR>>>(d, in, iters, swap, dout, KBLK);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("verify(swap=%d): %s\n", swap, cudaGetErrorString(e)); return -1; }
    std::vector<float> h(blocks * 128 * 16);
    cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int lane = 0; lane < 128; lane++) {
        double exp[16] = {0};
        for (int g = 0; g < 2; g++) {
            const int tid = g * 128 + lane;
            for (int it = 0; it < iters; it++)
                for (int i = 0; i < NREG; i++)
                    for (int hh = 0; hh < 2; hh++) exp[lag_of(i, hh)] += pat(tid, it, i, hh);
        }
        for (int n = 0; n < 16; n++)
            if (h[(size_t)lane * 16 + n] != (float)exp[n]) { if (bad < 4) printf("  lane %d col %d: got %g want %g\n", lane, n, h[(size_t)lane * 16 + n], exp[n]); bad++; }
    }
    printf("verify(swap_lbo=%d): %s (%d mismatches of 2048)\n", swap, bad ? "MISMATCH" : "exact", bad);
    return bad;
}

int main(int argc, char** argv)
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    cudaDeviceGetAttribute(&g_clk, cudaDevAttrClockRate, 0);
    g_sms = p.multiProcessorCount;
    float *d, *in, *dout;
    cudaMalloc(&d, (size_t)g_sms * 2 * NPROD * 4);
    cudaMalloc(&in, 1024 * 4);
    cudaMalloc(&dout, sizeof(float) * 4 * 128 * 16);
    float h[1024];
    for (int i = 0; i < 1024; i++) h[i] = 1.0f + i * 0.37f;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    const char* mode = argc > 1 ? argv[1] : "all";
    if (!strcmp(mode, "verify")) return verify(d, in, dout, argc > 2 ? atoi(argv[2]) : 0) != 0;
    for (int w = 0; w < 10; w++) k<1, 1><<<g_sms * 2, NTHR>>>(d, in, 2048, 0, nullptr, KBLK);  // spin the clocks up
    cudaDeviceSynchronize();
    const bool all = !strcmp(mode, "all");
    if (all || !strcmp(mode, "base")) {
        run<1, 3>("f32: 8 FMUL + 4 F2FP (+4 XOR)", d, in, dout);
        run<5, 3>("f32: 4 FMUL2 + 4 F2FP (+4 XOR)", d, in, dout);
        run<2, 3>("u8 : 4 HMUL2 (+4 XOR)", d, in, dout);
        run<1, 1>("f32: 8 FMUL + 4 F2FP + 8 FHADD", d, in, dout);
        run<1, 2>("f32: 8 FMUL + 4 F2FP + 1 HMMA", d, in, dout);
        run<5, 2>("f32: 4 FMUL2 + 4 F2FP + 1 HMMA", d, in, dout);
        run<2, 1>("u8 : 4 HMUL2 + 8 FHADD", d, in, dout);
        run<2, 2>("u8 : 4 HMUL2 + 1 HMMA", d, in, dout);
        run<5, 1>("f32: 4 FMUL2 + 4 F2FP + 8 FHADD", d, in, dout);
    }
    if (all || !strcmp(mode, "tc")) {
        run<1, 4>("f32: 8 FMUL + 4 F2FP + tcgen05", d, in, dout);
        run<5, 4>("f32: 4 FMUL2 + 4 F2FP + tcgen05", d, in, dout);
        run<2, 4>("u8 : 4 HMUL2 + tcgen05", d, in, dout);
        run<0, 4>("pattern only + tcgen05", d, in, dout);
        // attribution: the same protocol with fewer tensor-core instructions per slot hand-over
        g_nmma = 1; run<0, 4>("pattern only + tcgen05, 1 MMA/slot", d, in, dout);
        g_nmma = 0; run<0, 4>("pattern only + tcgen05.st, no MMA", d, in, dout);
        g_nmma = KBLK;
    }
    return 0;
}
