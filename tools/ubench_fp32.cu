// FP32 issue-rate microbenchmark for sm_100a: which instruction forms reach 128 lane-ops/clk/SM?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp32 ubench_fp32.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

struct Coef { float c[32]; };

#define PACK(lo, hi, out) asm("mov.b64 %0, {%1, %2};" : "=l"(out) : "f"(lo), "f"(hi))
#define UNPACK(in, lo, hi) asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(in))

template <int MODE>
__global__ void __launch_bounds__(1024) k(const __grid_constant__ Coef P, float *out, int iters, float seed)
{
    constexpr int NCH = 16;
    float a[NCH];
    unsigned long long a2[NCH / 2];
    float x = seed + threadIdx.x * 1e-6f, y = seed * 0.5f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) a[i] = x + i;
#pragma unroll
    for (int i = 0; i < NCH / 2; ++i) PACK(a[2 * i], a[2 * i + 1], a2[i]);
    unsigned long long x2, y2, z2, t2[NCH / 2];
    PACK(x, x, x2);
    PACK(y, y, y2);
    PACK(P.c[30] * 0.f - 0.f, P.c[31] * 0.f - 0.f, z2);   // opaque to the compiler
#pragma unroll
    for (int i = 0; i < NCH / 2; ++i) t2[i] = 0;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
            if (MODE == 0) {  // FFMA R,R,R,R
#pragma unroll
                for (int i = 0; i < NCH; ++i) a[i] = __fmaf_rn(a[i], x, y);
            } else if (MODE == 1) {  // FFMA R,R,UR,R  (coefficient from constant bank)
#pragma unroll
                for (int i = 0; i < NCH; ++i) a[i] = __fmaf_rn(x, P.c[(i + rep * 4) & 31], a[i]);
            } else if (MODE == 2) {  // FMUL R,R,UR ; FADD R,R,R  (parity pair)
#pragma unroll
                for (int i = 0; i < NCH; ++i) a[i] = __fadd_rn(__fmul_rn(a[i], P.c[(i + rep * 4) & 31]), y);
            } else if (MODE == 3) {  // FFMA2
#pragma unroll
                for (int i = 0; i < NCH / 2; ++i)
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(a2[i]) : "l"(a2[i]), "l"(x2), "l"(y2));
            } else if (MODE == 4) {  // FMUL2 + FADD2
#pragma unroll
                for (int i = 0; i < NCH / 2; ++i) {
                    unsigned long long t;
                    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(x2), "l"(y2));
                    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(a2[i]) : "l"(t), "l"(a2[i]));
                }
            } else if (MODE == 5) {  // FADD R,R,R only
#pragma unroll
                for (int i = 0; i < NCH; ++i) a[i] = __fadd_rn(a[i], x);
            } else if (MODE == 6) {  // FMUL R,R,R only
#pragma unroll
                for (int i = 0; i < NCH; ++i) a[i] = __fmul_rn(a[i], x);
            } else if (MODE == 7) {  // FFMA2 with coefficient pair from constant bank
#pragma unroll
                for (int i = 0; i < NCH / 2; ++i) {
                    unsigned long long c2;
                    PACK(P.c[(2 * i + rep * 4) & 31], P.c[(2 * i + 1 + rep * 4) & 31], c2);
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(a2[i]) : "l"(x2), "l"(c2), "l"(a2[i]));
                }
            } else if (MODE == 9) {  // 2x FMUL R,R,UR + FADD2  (parity with packed accumulate)
#pragma unroll
                for (int i = 0; i < NCH / 2; ++i) {
                    float lo, hi;
                    UNPACK(a2[i], lo, hi);
                    float e0 = __fmul_rn(lo, P.c[(2 * i + rep * 4) & 31]), e1 = __fmul_rn(hi, P.c[(2 * i + 1 + rep * 4) & 31]);
                    unsigned long long t;
                    PACK(e0, e1, t);
                    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(a2[i]) : "l"(t), "l"(y2));
                }
            } else if (MODE == 10) {  // FADD2 only
#pragma unroll
                for (int i = 0; i < NCH / 2; ++i)
                    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(a2[i]) : "l"(x2), "l"(a2[i]));
            } else if (MODE == 11) {  // exact packed product: FFMA2(x.F32, c2 from const, Z = (-0,-0) opaque) + FADD2 accumulate
#pragma unroll
                for (int i = 0; i < NCH / 2; ++i) {
                    unsigned long long c2, t;
                    PACK(P.c[(2 * i + rep * 4) & 31], P.c[(2 * i + 1 + rep * 4) & 31], c2);
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(t) : "l"(x2), "l"(c2), "l"(z2));
                    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(a2[i]) : "l"(t), "l"(a2[i]));
                }
            } else if (MODE == 12) {  // same, but the product feeds a DIFFERENT accumulator pair's add one step later (more ILP)
#pragma unroll
                for (int i = 0; i < NCH / 2; ++i) {
                    unsigned long long c2;
                    PACK(P.c[(2 * i + rep * 4) & 31], P.c[(2 * i + 1 + rep * 4) & 31], c2);
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(t2[i]) : "l"(x2), "l"(c2), "l"(z2));
                }
#pragma unroll
                for (int i = 0; i < NCH / 2; ++i)
                    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(a2[i]) : "l"(t2[i]), "l"(a2[i]));
            } else if (MODE == 8) {  // FMNMX + FADD mix (alu + fma pipes)
#pragma unroll
                for (int i = 0; i < NCH; i += 2) { a[i] = fmaxf(a[i], a[i + 1]); a[i + 1] = __fadd_rn(a[i + 1], y); }
            }
        }
    }
    float s = 0.f;
    if (MODE == 3 || MODE == 4 || MODE == 7 || MODE == 9 || MODE == 10 || MODE == 11 || MODE == 12) {
#pragma unroll
        for (int i = 0; i < NCH / 2; ++i) { float lo, hi; UNPACK(a2[i], lo, hi); s += lo + hi; }
    } else {
#pragma unroll
        for (int i = 0; i < NCH; ++i) s += a[i];
    }
    if (s == 12345.678f) out[0] = s;
}

template <int MODE> void run(const char *name, int lane_ops_per_inner, int threads, int blocks_per_sm)
{
    int dev; cudaGetDevice(&dev);
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    Coef P; for (int i = 0; i < 32; ++i) P.c[i] = 1.0f + 1e-7f * i;
    float *out; cudaMalloc(&out, 4);
    int iters = 20000;
    dim3 grid(pr.multiProcessorCount * blocks_per_sm);
    k<MODE><<<grid, threads>>>(P, out, 100, 1.0f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<grid, threads>>>(P, out, iters, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
    double ops = (double)grid.x * threads * iters * 8.0 * lane_ops_per_inner;  // lane-level FP instructions (FFMA counts 1)
    double per_sm_clk = ops / (ms * 1e-3) / pr.multiProcessorCount / (clk_khz * 1e3);
    printf("%-34s thr/blk=%4d blk/SM=%d  %8.3f ms  %7.1f lane-ops/clk/SM (at nominal %d MHz)  %.2f Tlane-op/s\n", name, threads,
           blocks_per_sm, ms, per_sm_clk, clk_khz / 1000, ops / (ms * 1e-3) / 1e12);
    cudaFree(out);
}

int main()
{
    for (int cfg = 0; cfg < 3; ++cfg) {
        int threads = cfg == 0 ? 128 : (cfg == 1 ? 256 : 1024), bps = 1;
        run<0>("FFMA R,R,R,R", 16, threads, bps);
        run<1>("FFMA R,R,UR,R", 16, threads, bps);
        run<2>("FMUL R,R,UR + FADD", 32, threads, bps);
        run<3>("FFMA2 (f32x2)", 16, threads, bps);
        run<4>("FMUL2 + FADD2", 32, threads, bps);
        run<5>("FADD R,R,R", 16, threads, bps);
        run<6>("FMUL R,R,R", 16, threads, bps);
        run<7>("FFMA2 coef pair from const", 16, threads, bps);
        run<8>("FMNMX + FADD mix", 16, threads, bps);
        run<9>("2xFMUL R,R,UR + FADD2", 32, threads, bps);
        run<10>("FADD2", 16, threads, bps);
        run<11>("FFMA2(x,c2,-0) + FADD2 same acc", 32, threads, bps);
        run<12>("FFMA2(x,c2,-0) x8 then FADD2 x8", 32, threads, bps);
    }
    return 0;
}
