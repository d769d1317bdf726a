// Instantiations of the fp32 12/4/10 production kernel (tmpc_kernel_f32.cuh) behind tmpc_dispatch::lookup_f32.
#include "tmpc.h"
#include "tmpc_dispatch.hpp"
#include "tmpc_kernel_f32.cuh"

namespace tmpc_dispatch {
namespace {

template <int NX, int NU, int NH, int BLOCK, bool FAST, bool WARM, bool TM, class PAT, bool CB, bool IB = false, bool CSM = false, bool ROLL = false>
KernelInfo make_info_f32()
{
    KernelInfo k;
    k.fn = (const void *)&tmpc::admm_kernel_f32<NX, NU, NH, BLOCK, FAST, WARM, TM, PAT, CB, IB, CSM, ROLL>;
    k.smem = CSM ? tmpc::SmemLayoutF32<NX, NU, NH, BLOCK, TM>::BYTES_CSM : ROLL ? tmpc::SmemLayoutF32<NX, NU, NH, BLOCK, TM>::BYTES_ROLL : tmpc::SmemLayoutF32<NX, NU, NH, BLOCK, TM>::BYTES;
    k.block = BLOCK;
    k.model_bytes = sizeof(tmpc::ModelF32<NX, NU, NH>);
    k.model_kind = 1;
    k.per_block = BLOCK;
    return k;
}

template <int NX, int NU, int NH, int BLOCK, bool TM, class PAT = tmpc::PatDense<NX>, bool CB = false, bool IB = false>
bool pick_f32(int policy, bool warm, KernelInfo &out)
{
    if (policy == TMPC_ORDER_PARITY)
        out = warm ? make_info_f32<NX, NU, NH, BLOCK, false, true, TM, PAT, CB, IB>() : make_info_f32<NX, NU, NH, BLOCK, false, false, TM, PAT, CB, IB>();
    else
        out = warm ? make_info_f32<NX, NU, NH, BLOCK, true, true, TM, PAT, CB, IB>() : make_info_f32<NX, NU, NH, BLOCK, true, false, TM, PAT, CB, IB>();
    return true;
}

}  // namespace

bool lookup_f32(int policy, bool warm, int pattern, bool cb, int variant, KernelInfo &out, bool ib)
{
    if (variant == 3 && !ib && !warm && policy == TMPC_ORDER_PARITY && cb) {
        // model image staged into shared memory by TMA (the A/B of tmpc_kernel_f32.cuh CSM); cold PARITY solves with constant bounds
        out = pattern == tmpc::PatQuadrotor::id ? make_info_f32<12, 4, 10, 256, false, false, true, tmpc::PatQuadrotor, true, false, true>()
                                                : make_info_f32<12, 4, 10, 256, false, false, true, tmpc::PatDense<12>, true, false, true>();
        return true;
    }
    if (variant == 3) variant = 2;
    if (ib) {
        if (variant != 2) return false;
        return pattern == tmpc::PatQuadrotor::id ? pick_f32<12, 4, 10, 256, true, tmpc::PatQuadrotor, false, true>(policy, warm, out)
                                                 : pick_f32<12, 4, 10, 256, true, tmpc::PatDense<12>, false, true>(policy, warm, out);
    }
    // model-structure specialisation (tmpc_kernel_f32.cuh PatQuadrotor): chosen by build_model when the actual matrices conform
    if (variant == 2 && pattern == tmpc::PatQuadrotor::id)
        return cb ? pick_f32<12, 4, 10, 256, true, tmpc::PatQuadrotor, true>(policy, warm, out)
                  : pick_f32<12, 4, 10, 256, true, tmpc::PatQuadrotor, false>(policy, warm, out);
    if (variant == 2)                                                              // g,v in TMEM: 256 instances / SM
        return cb ? pick_f32<12, 4, 10, 256, true, tmpc::PatDense<12>, true>(policy, warm, out)
                  : pick_f32<12, 4, 10, 256, true, tmpc::PatDense<12>, false>(policy, warm, out);
    return pick_f32<12, 4, 10, 128, false>(policy, warm, out);                     // all state in shared memory
}

const void *lookup_f32_pn_seed(int policy)
{
    return policy == TMPC_ORDER_PARITY ? (const void *)&tmpc::pn_seed_kernel<12, 4, 10, false> : (const void *)&tmpc::pn_seed_kernel<12, 4, 10, true>;
}

// fused closed loop (ROLL instances): warm, tensor-memory variant, shared box, PARITY order
bool lookup_f32_roll(int pattern, bool cb, KernelInfo &out)
{
    if (pattern == tmpc::PatQuadrotor::id)
        out = cb ? make_info_f32<12, 4, 10, 256, false, true, true, tmpc::PatQuadrotor, true, false, false, true>()
                 : make_info_f32<12, 4, 10, 256, false, true, true, tmpc::PatQuadrotor, false, false, false, true>();
    else
        out = cb ? make_info_f32<12, 4, 10, 256, false, true, true, tmpc::PatDense<12>, true, false, false, true>()
                 : make_info_f32<12, 4, 10, 256, false, true, true, tmpc::PatDense<12>, false, false, false, true>();
    return true;
}

}  // namespace tmpc_dispatch
