// Instantiations of the register-resident 4/1/10 kernel (tmpc_kernel_small.cuh) and of the warp-per-instance 32/8/50
// kernel (tmpc_kernel_warp.cuh) behind tmpc_dispatch::lookup_small / lookup_warp.
#include "tmpc.h"
#include "tmpc_dispatch.hpp"
#include "tmpc_kernel_small.cuh"
#include "tmpc_kernel_warp.cuh"
#include "tmpc_kernel_warp4.cuh"

namespace tmpc_dispatch {
namespace {

template <int NX, int NH, int BLOCK, bool FAST, bool WARM>
KernelInfo make_info_small()
{
    KernelInfo k;
    k.fn = (const void *)&tmpc::admm_kernel_small<NX, NH, BLOCK, FAST, WARM>;
    k.smem = 0;   // the whole per-instance state lives in registers
    k.block = BLOCK;
    k.model_bytes = sizeof(tmpc::Model<float, NX, 1, NH>);
    k.model_kind = 0;
    k.per_block = BLOCK;
    return k;
}

template <int NX, int NH, int BLOCK>
bool pick_small(int policy, bool warm, KernelInfo &out)
{
    if (policy == TMPC_ORDER_PARITY) out = warm ? make_info_small<NX, NH, BLOCK, false, true>() : make_info_small<NX, NH, BLOCK, false, false>();
    else out = warm ? make_info_small<NX, NH, BLOCK, true, true>() : make_info_small<NX, NH, BLOCK, true, false>();
    return true;
}

template <int NH, int WARPS, bool FAST, bool WARM, bool TM>
KernelInfo make_info_warp()
{
    KernelInfo k;
    k.fn = (const void *)&tmpc::admm_kernel_warp<NH, WARPS, FAST, WARM, TM>;
    k.smem = tmpc::WarpSmem<NH, TM>::total_bytes(WARPS);
    k.block = WARPS * 32;
    k.model_bytes = sizeof(tmpc::ModelWarp);
    k.model_kind = 2;
    k.per_block = WARPS;
    return k;
}

template <int NH, int WARPS, bool TM>
bool pick_warp(int policy, bool warm, KernelInfo &out)
{
    static_assert(tmpc::WarpSmem<NH, TM>::total_bytes(WARPS) <= 232448, "per-block shared memory limit of sm_100");
    if (policy == TMPC_ORDER_PARITY) out = warm ? make_info_warp<NH, WARPS, false, true, TM>() : make_info_warp<NH, WARPS, false, false, TM>();
    else out = warm ? make_info_warp<NH, WARPS, true, true, TM>() : make_info_warp<NH, WARPS, true, false, TM>();
    return true;
}

}  // namespace

// fused closed loop at 4/1/10 (admm_kernel_small_roll): PARITY, warm buffers; v of the reference in shared memory (160 B per lane)
template <int BLOCK> KernelInfo make_info_small_roll()
{
    KernelInfo k;
    k.fn = (const void *)&tmpc::admm_kernel_small_roll<4, 10, BLOCK>;
    k.smem = tmpc::SVec<float, 4, 10, BLOCK>::BYTES;
    k.block = BLOCK;
    k.model_bytes = sizeof(tmpc::Model<float, 4, 1, 10>);
    k.model_kind = 0;
    k.per_block = BLOCK;
    return k;
}
bool lookup_small_roll(int block, KernelInfo &out)
{
    out = block == 256 ? make_info_small_roll<256>() : block == 512 ? make_info_small_roll<512>() : make_info_small_roll<384>();
    return true;
}

bool lookup_small(int block, int policy, bool warm, KernelInfo &out)
{
    if (block == 256) return pick_small<4, 10, 256>(policy, warm, out);
    if (block == 512) return pick_small<4, 10, 512>(policy, warm, out);
    return pick_small<4, 10, 384>(policy, warm, out);
}

template <int NH, bool FAST, bool WARM>
KernelInfo make_info_warp4()
{
    static_assert(tmpc::Warp4Smem<NH>::total_bytes(4) <= 232448, "per-block shared memory limit of sm_100");
    KernelInfo k;
    k.fn = (const void *)&tmpc::admm_kernel_warp4<NH, FAST, WARM>;
    k.smem = tmpc::Warp4Smem<NH>::total_bytes(4);
    k.block = 128;
    k.model_bytes = sizeof(tmpc::ModelWarp);
    k.model_kind = 2;
    k.per_block = 16;   // 4 warps x 4 slots
    return k;
}

bool lookup_warp(int variant, int policy, bool warm, KernelInfo &out)
{
    if (variant == 1) return pick_warp<50, 16, true>(policy, warm, out);
    if (variant == 2) return pick_warp<50, 12, false>(policy, warm, out);
    if (policy == TMPC_ORDER_PARITY) out = warm ? make_info_warp4<50, false, true>() : make_info_warp4<50, false, false>();
    else out = warm ? make_info_warp4<50, true, true>() : make_info_warp4<50, true, false>();
    return true;
}

}  // namespace tmpc_dispatch
