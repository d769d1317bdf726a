// Instantiations of the two-lanes-per-instance fp64 12/4/10 kernel (tmpc_kernel_f64p.cuh) behind tmpc_dispatch::lookup_f64p.
#include "tmpc.h"
#include "tmpc_dispatch.hpp"
#include "tmpc_kernel_f64p.cuh"

namespace tmpc_dispatch {
namespace {

template <int NH, bool FAST, bool WARM> KernelInfo make_info_f64p()
{
    KernelInfo k;
    k.fn = (const void *)&tmpc::admm_kernel_f64p<NH, FAST, WARM>;
    k.smem = tmpc::f64p::Smem<NH>::BYTES;
    k.block = 256;
    k.model_bytes = sizeof(tmpc::Model<double, 12, 4, NH>);
    k.model_kind = 0;
    k.per_block = 128;   // two lanes per instance
    return k;
}

}  // namespace

bool lookup_f64p(int nx, int nu, int N, int dtype, int policy, bool warm, KernelInfo &out)
{
    if (!(nx == 12 && nu == 4 && N == 10 && dtype == TMPC_F64)) return false;
    if (policy == TMPC_ORDER_PARITY) out = warm ? make_info_f64p<10, false, true>() : make_info_f64p<10, false, false>();
    else out = warm ? make_info_f64p<10, true, true>() : make_info_f64p<10, true, false>();
    return true;
}

}  // namespace tmpc_dispatch
