// TMEM as a per-thread scratchpad: tcgen05.st / tcgen05.ld (32x32b.x4) round-trip correctness + throughput.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_tmem ubench_tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_st4(uint32_t taddr, float a, float b, float c, float d)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 :: "r"(taddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float &a, float &b, float &c, float &d)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(float *out, int iters, int *errors)
{
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(dst) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_base_s;
    // this warp's lane quarter + its 256-column half
    const uint32_t my = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
    float acc = 0.f;
    int bad = 0;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int c = 0; c < 60; ++c) {
                float v = (float)(threadIdx.x * 1000 + c * 4 + it);
                tmem_st4(my + c * 4, v, v + 1.f, v + 2.f, v + 3.f);
            }
            tmem_wait_st();
        }
        if (MODE == 0 || MODE == 1) {
#pragma unroll
            for (int c = 0; c < 60; c += 6) {
                float r[24];
#pragma unroll
                for (int j = 0; j < 6; ++j) tmem_ld4(my + (c + j) * 4, r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    if (MODE == 0) {
                        float v = (float)(threadIdx.x * 1000 + (c + j) * 4 + it);
                        bad += (r[4 * j] != v) + (r[4 * j + 1] != v + 1.f) + (r[4 * j + 2] != v + 2.f) + (r[4 * j + 3] != v + 3.f);
                    }
                    acc += r[4 * j] + r[4 * j + 1] + r[4 * j + 2] + r[4 * j + 3];
                }
            }
        }
    }
    if (bad) atomicAdd(errors, bad);
    if (acc == 123.456f) out[0] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(base) : "memory");
}

template <int MODE> void run(const char *name, int iters)
{
    int dev; cudaGetDevice(&dev);
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    float *out; int *err; cudaMalloc(&out, 4); cudaMalloc(&err, 4); cudaMemset(err, 0, 4);
    k<MODE><<<pr.multiProcessorCount, 256>>>(out, 10, err);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<pr.multiProcessorCount, 256>>>(out, iters, err);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int herr; cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
    double bytes = (double)256 * 240 * 4 * iters * ((MODE == 0) ? 2 : 1);   // per SM
    printf("%-28s %s  %.3f ms  %.1f B/clk/SM (nominal clk)  errors=%d\n", name, cudaGetErrorString(e), ms,
           bytes / (ms * 1e-3 * clk_khz * 1e3), herr);
}

int main()
{
    run<0>("st+ld round trip (verify)", 2000);
    run<1>("ld only", 4000);
    run<2>("st only", 4000);
    return 0;
}
