// Kernel tables of the library, one translation unit per kernel family (k_*.cu) so that the families compile in parallel
// and a change to one kernel rebuilds one object.  tmpc_api.cu owns contexts, model images and launches; it gets function
// pointers, launch shapes and shared-memory sizes from the look-ups below.
#pragma once
#include <cstddef>

namespace tmpc_dispatch {

struct KernelInfo {
    const void *fn;
    size_t smem;
    int block;
    size_t model_bytes;
    int model_kind;  // 0: tmpc::Model<T,...> (generic kernel)   1: tmpc::ModelF32<...> (packed fp32 kernel)
                     // 2: tmpc::ModelWarp (warp-per-instance kernel; pointers into the ctx's device model image)
                     // 3: tmpc::ModelRT<T> (run-time-shape kernel; pointers into the ctx's device model image + scratch)
    int per_block;   // instances resident per block (threads for the thread-per-instance kernels, warps for kind 2)
    size_t scratch_per_block = 0;   // bytes of SolveArgs::scratch per block (kind 2, eight-warp four-slot kernel)
};

// fp32 12/4/10 production kernel (tmpc_kernel_f32.cuh).  variant 2: g, v in tensor memory, 256 instances / SM (default);
// 1: all state in shared memory, 128 / SM; 3: as 2 with the model image staged into shared memory by a TMA bulk copy instead
// of living in the constant bank (cold PARITY solves; everything else falls back to 2).  pattern: 0 dense, 1 quadrotor.
// per_instance_bounds: the IB instances (variant 2 only; boxes read from the lane's scratch rows)
bool lookup_f32(int policy, bool warm, int pattern, bool const_bounds, int variant, KernelInfo &out, bool per_instance_bounds = false);

// pre-pass of that kernel for per-instance reference trajectories: (ModelF32, Xref, stride, batch, out [batch][12])
const void *lookup_f32_pn_seed(int policy);
// the same kernel running a whole closed loop per claimed instance (SolveArgs::roll_steps MPC steps, state kept on the lane between
// them): warm PARITY solves with a shared box on the tensor-memory variant
bool lookup_f32_roll(int pattern, bool const_bounds, KernelInfo &out);

// generic shared-memory kernel (tmpc_kernel.cuh): fp64 shapes and the development variants of the fp32 shapes.
// variant 0: default for the shape / dtype; 1: all state in shared memory; 2: horizon loops unrolled (4/1/10 fp32 only)
bool lookup_generic(int nx, int nu, int N, int dtype, int policy, bool warm, int variant, KernelInfo &out);
// per-instance systems instances of the generic kernel; global_coeffs: re-read the coefficients from global memory
bool lookup_sys(int nx, int nu, int N, int dtype, int policy, bool warm, bool global_coeffs, KernelInfo &out);
// fp32 12/4/10 per-instance systems, row-pair kernels.  variant 0: two lanes per instance (tmpc_kernel_sysp.cuh, the default for
// that shape); 1: one thread per instance (tmpc_kernel_sys.cuh)
bool lookup_sys_pairs(int nx, int nu, int N, int dtype, int policy, bool warm, bool const_bounds, int variant, KernelInfo &out);

// fp64 12/4/10, one shared model, two lanes per instance (tmpc_kernel_f64p.cuh): the default for that shape / dtype
bool lookup_f64p(int nx, int nu, int N, int dtype, int policy, bool warm, KernelInfo &out);

// register-resident single-input kernel, fp32 4/1/10 (tmpc_kernel_small.cuh); block = 256, 384 or 512
bool lookup_small(int block, int policy, bool warm, KernelInfo &out);
// the same shape running a whole closed loop per claimed instance in registers (SolveArgs::roll_steps MPC steps); block = 256 or 384
bool lookup_small_roll(int block, KernelInfo &out);
// fp32 32/8/50: variant 0 = four instances per warp (tmpc_kernel_warp4.cuh, default); 1 = one instance per warp, g, v in tensor
// memory (tmpc_kernel_warp.cuh, 16 instances / SM); 2 = one instance per warp, all state in shared memory (12 / SM);
// 3 = four instances per warp, eight warps (32 instances / SM; g, v of two slots per warp in L2-resident scratch)
bool lookup_warp(int variant, int policy, bool warm, KernelInfo &out);

}  // namespace tmpc_dispatch
