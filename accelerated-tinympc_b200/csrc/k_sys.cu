// Instantiations of the row-pair per-instance-systems kernel (tmpc_kernel_sys.cuh) behind tmpc_dispatch::lookup_sys_pairs.
#include "tmpc.h"
#include "tmpc_dispatch.hpp"
#include "tmpc_kernel_sys.cuh"
#include "tmpc_kernel_sysp.cuh"

namespace tmpc_dispatch {
namespace {

template <int NH, bool FAST, bool WARM, bool CB> KernelInfo make_info_sys()
{
    KernelInfo k;
    k.fn = (const void *)&tmpc::admm_kernel_sys<NH, FAST, WARM, CB>;
    k.smem = tmpc::sysk::SysSmem<NH>::BYTES;
    k.block = 128;
    k.model_bytes = sizeof(tmpc::Model<float, 12, 4, NH>);
    k.model_kind = 0;
    k.per_block = 128;
    return k;
}

}  // namespace

template <int NH, bool FAST, bool WARM, bool CB> KernelInfo make_info_sysp()
{
    KernelInfo k;
    k.fn = (const void *)&tmpc::admm_kernel_sysp<NH, FAST, WARM, CB>;
    k.smem = tmpc::sysk::SysPSmem<NH>::BYTES;
    k.block = 256;
    k.model_bytes = sizeof(tmpc::Model<float, 12, 4, NH>);
    k.model_kind = 0;
    k.per_block = 128;   // two lanes per instance
    return k;
}

template <bool CB> KernelInfo pick_lanepairs(int policy, bool warm)
{
    if (policy == TMPC_ORDER_PARITY) return warm ? make_info_sysp<10, false, true, CB>() : make_info_sysp<10, false, false, CB>();
    return warm ? make_info_sysp<10, true, true, CB>() : make_info_sysp<10, true, false, CB>();
}

template <bool CB> KernelInfo pick_pairs(int policy, bool warm)
{
    if (policy == TMPC_ORDER_PARITY) return warm ? make_info_sys<10, false, true, CB>() : make_info_sys<10, false, false, CB>();
    return warm ? make_info_sys<10, true, true, CB>() : make_info_sys<10, true, false, CB>();
}

bool lookup_sys_pairs(int nx, int nu, int N, int dtype, int policy, bool warm, bool const_bounds, int variant, KernelInfo &out)
{
    if (!(nx == 12 && nu == 4 && N == 10 && dtype == TMPC_F32)) return false;
    if (variant == 1) out = const_bounds ? pick_pairs<true>(policy, warm) : pick_pairs<false>(policy, warm);
    else out = const_bounds ? pick_lanepairs<true>(policy, warm) : pick_lanepairs<false>(policy, warm);
    return true;
}

}  // namespace tmpc_dispatch
