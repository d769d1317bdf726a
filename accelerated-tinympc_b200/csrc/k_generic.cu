// Instantiations of the generic shared-memory kernel (tmpc_kernel.cuh): fp64 shapes, development variants of the fp32
// shapes, per-instance systems -- behind tmpc_dispatch::lookup_generic / lookup_sys.
#include "tmpc.h"
#include "tmpc_dispatch.hpp"
#include "tmpc_kernel.cuh"

namespace tmpc_dispatch {
namespace {

template <class T, int NX, int NU, int NH, int BLOCK, bool FAST, bool WARM, bool UNROLL>
KernelInfo make_info()
{
    KernelInfo k;
    k.fn = (const void *)&tmpc::admm_kernel<T, NX, NU, NH, BLOCK, FAST, WARM, UNROLL>;
    k.smem = tmpc::SmemLayout<T, NX, NU, NH, BLOCK>::BYTES;
    k.block = BLOCK;
    k.model_bytes = sizeof(tmpc::Model<T, NX, NU, NH>);
    k.model_kind = 0;
    k.per_block = BLOCK;
    return k;
}

template <class T, int NX, int NU, int NH, int BLOCK, bool FAST, bool WARM, int SYS>
KernelInfo make_info_g()
{
    KernelInfo k;
    k.fn = (const void *)&tmpc::admm_kernel<T, NX, NU, NH, BLOCK, FAST, WARM, false, SYS>;
    k.smem = SYS == 3 ? tmpc::SmemLayout<T, NX, NU, NH, BLOCK>::BYTES_NOGV + 16                 // g, v in TMEM
                      : tmpc::SmemLayout<T, NX, NU, NH, BLOCK>::BYTES + (SYS == 2 ? 16 : 0);   // + the TMEM base slot
    k.block = BLOCK;
    k.model_bytes = sizeof(tmpc::Model<T, NX, NU, NH>);
    k.model_kind = 0;
    k.per_block = BLOCK;
    return k;
}

template <class T, int NX, int NU, int NH, int BLOCK, bool UNROLL>
bool pick(int policy, bool warm, KernelInfo &out)
{
    if (policy == TMPC_ORDER_PARITY)
        out = warm ? make_info<T, NX, NU, NH, BLOCK, false, true, UNROLL>()
                   : make_info<T, NX, NU, NH, BLOCK, false, false, UNROLL>();
    else
        out = warm ? make_info<T, NX, NU, NH, BLOCK, true, true, UNROLL>()
                   : make_info<T, NX, NU, NH, BLOCK, true, false, UNROLL>();
    return true;
}

// per-instance-systems instances of the generic kernel (SYS = 1: coefficients from the global block; 2: TMEM-resident)
template <class T, int NX, int NU, int NH, int BLOCK, int SYS>
bool pick_sys(int policy, bool warm, KernelInfo &out)
{
    if (policy == TMPC_ORDER_PARITY)
        out = warm ? make_info_g<T, NX, NU, NH, BLOCK, false, true, SYS>() : make_info_g<T, NX, NU, NH, BLOCK, false, false, SYS>();
    else
        out = warm ? make_info_g<T, NX, NU, NH, BLOCK, true, true, SYS>() : make_info_g<T, NX, NU, NH, BLOCK, true, false, SYS>();
    return true;
}

}  // namespace

// Thread-per-instance needs the per-instance state to fit shared memory:
//   quadrotor 12/4/10: 360 scalars -> 128 threads (f32) / 64 threads (f64) per SM
//   cartpole   4/1/10: 111 scalars -> 512 threads (f32) / 128 (f64)
bool lookup_generic(int nx, int nu, int N, int dtype, int policy, bool warm, int variant, KernelInfo &out)
{
    if (nx == 12 && nu == 4 && N == 10) {
        if (dtype == TMPC_F32) return pick<float, 12, 4, 10, 128, false>(policy, warm, out);
        // double: g, v in tensor memory -> 128 instances / SM (variant 1: all state in shared memory, 64 / SM)
        if (variant == 1) return pick<double, 12, 4, 10, 64, false>(policy, warm, out);
        return pick_sys<double, 12, 4, 10, 128, 3>(policy, warm, out);
    }
    if (nx == 4 && nu == 1 && N == 10) {
        if (dtype == TMPC_F32)
            return variant == 2 ? pick<float, 4, 1, 10, 512, true>(policy, warm, out) : pick<float, 4, 1, 10, 512, false>(policy, warm, out);
        return pick<double, 4, 1, 10, 128, false>(policy, warm, out);
    }
    return false;
}

bool lookup_sys(int nx, int nu, int N, int dtype, int policy, bool warm, bool global_coeffs, KernelInfo &out)
{
    if (nx == 12 && nu == 4 && N == 10) {
        if (dtype == TMPC_F32)
            return global_coeffs ? pick_sys<float, 12, 4, 10, 128, 1>(policy, warm, out) : pick_sys<float, 12, 4, 10, 128, 2>(policy, warm, out);
        return pick_sys<double, 12, 4, 10, 64, 1>(policy, warm, out);
    }
    if (nx == 4 && nu == 1 && N == 10)
        return dtype == TMPC_F32 ? pick_sys<float, 4, 1, 10, 512, 1>(policy, warm, out) : pick_sys<double, 4, 1, 10, 128, 1>(policy, warm, out);
    return false;
}

}  // namespace tmpc_dispatch
