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

template <int NH, bool FAST, bool WARM, int WARPS>
KernelInfo make_info_warp4()
{
    static_assert(tmpc::Warp4Smem<NH>::total_bytes(WARPS) <= 232448, "per-block shared memory limit of sm_100");
    KernelInfo k;
    k.fn = (const void *)&tmpc::admm_kernel_warp4<NH, FAST, WARM, WARPS>;
    k.smem = tmpc::Warp4Smem<NH>::total_bytes(WARPS);
    k.block = WARPS * 32;
    k.model_bytes = sizeof(tmpc::ModelWarp);
    k.model_kind = 2;
    k.per_block = WARPS * 4;   // four slots per warp
    // eight warps: g, v of two slots per warp in the L2-resident scratch (rows of 128 B)
    k.scratch_per_block = WARPS == 8 ? size_t(WARPS) * 2 * 2 * NH * 128 : 0;
    return k;
}
template <int WARPS> void pick_warp4(int policy, bool warm, KernelInfo &out)
{
    if (policy == TMPC_ORDER_PARITY) out = warm ? make_info_warp4<50, false, true, WARPS>() : make_info_warp4<50, false, false, WARPS>();
    else out = warm ? make_info_warp4<50, true, true, WARPS>() : make_info_warp4<50, true, false, WARPS>();
}

bool lookup_warp(int variant, int policy, bool warm, KernelInfo &out)
{
    if (variant == 1) return pick_warp<50, 16, true>(policy, warm, out);
    if (variant == 2) return pick_warp<50, 12, false>(policy, warm, out);
    if (variant == 3) pick_warp4<8>(policy, warm, out);   // two warps per scheduler
    else pick_warp4<4>(policy, warm, out);
    return true;
}

}  // namespace tmpc_dispatch
