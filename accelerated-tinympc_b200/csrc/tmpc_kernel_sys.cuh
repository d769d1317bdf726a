// Per-instance SYSTEMS, fp32 12/4/N: the ADMM loop of tiny_solve (/root/reference/src/tinympc/admm.cpp:111-152) when every
// instance brings its own model and cache.  Second generation of the TMEM-resident kernel (the first one
// is the SYS == 2 instance of tmpc_kernel.cuh, kept as TMPC_KERNEL=sys_rows).
//
// What bounds this kernel (profiles/r01_ncu_systems_kernel.md): each lane needs its own 496 coefficients every horizon
// stage, they fill the lane's 512 tensor-memory columns, so a CTA is 128 threads = ONE warp per scheduler and nothing but
// instruction-level parallelism hides latency.  The first kernel fetched one coefficient ROW at a time and ran one dot
// product per row: a 12-long dependent FADD chain per row, scalar FMUL + FADD per MAC, issue slots 50 % busy.  Here:
//
//  * the tensor-memory image holds every matrix in the major in which PAIRS OF OUTPUT ROWS are adjacent
//    (Kinf, Adyn, Bdyn, Quu_inv, AmBKt column-major; Bdyn and Kinf row-major for the two transposed products), so one
//    tcgen05.ld.x16 returns eight coefficient register pairs and a MAC pair is FFMA2 (exact packed product, below) +
//    FADD2: half the issue slots of FMUL + FADD;
//  * a mat-vec is a COLUMN sweep: the 6 (or 2) row-pair chains advance together, products are formed as their
//    coefficients land, in the summation order of the reference build for each product (sequential / half-split tree /
//    vectorised redux -- tmpc_kernel.cuh Orders<float, 12, 4>), so results stay bit-identical;
//  * the coefficient stream is a ring of four 16-column buffers, three tcgen05.ld in flight ahead of the one being
//    consumed, running across stage boundaries and from the forward into the backward sweep;
//  * work.Q and rho live in registers; the refill rewrites the image sixty-four columns per tensor-memory round trip.
//
// Exact packed product: fma(a, b, -0) rounds the exact product once and adding -0 changes neither value nor sign, so
// FFMA2 with an addend pair (-0, -0) the compiler cannot see through IS the packed multiply (tmpc_kernel_f32.cuh prod2).
#pragma once
#include "tmpc_kernel.cuh"

namespace tmpc {
namespace sysk {

__device__ __forceinline__ float2 mk2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 ng2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 ad2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sb2(float2 a, float2 b) { return __fadd2_rn(a, ng2(b)); }
__device__ __forceinline__ float2 pr2(float2 c, float x, float2 Z) { return __ffma2_rn(c, make_float2(x, x), Z); }
__device__ __forceinline__ float2 pp2(float2 a, float2 b, float2 Z) { return __ffma2_rn(a, b, Z); }
__device__ __forceinline__ float2 fm2(float2 c, float x, float2 acc) { return __ffma2_rn(c, make_float2(x, x), acc); }
__device__ __forceinline__ float2 ml2(float2 c, float x) { return __fmul2_rn(c, make_float2(x, x)); }

__device__ __forceinline__ void wait16(float *r)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]),
                   "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]), "+f"(r[14]), "+f"(r[15]) :: "memory");
}

// Tensor-memory image of one system = the order in which the sweeps STREAM it, in 16-column groups ("units"): the forward
// sweep reads units 0..14 of a stage, the backward sweep units 15..30.  The products with only two row-pair chains (Kinf x;
// Bdyn^T p -> Quu_inv s) are interleaved with the six-chain ones (Adyn x; AmBKt p) so that the scheduler always has independent
// chains to issue from.  A unit is 16 consecutive coefficients of one matrix in the major that makes row pairs adjacent.
enum { U_K = 0, U_A = 1, U_B = 2, U_BR = 3, U_QI = 4, U_M = 5, U_KR = 6 };
struct Unit { int kind, idx; };
struct TmMap {
    static constexpr int FWD_UNITS = 15, BWD_UNITS = 16, GROUPS = FWD_UNITS + BWD_UNITS, BWD_COL = 16 * FWD_UNITS;
    // forward:  A0 K0 A1 K1 A2 K2 A3 .. A8 B0 B1 B2       (Kinf, Adyn, Bdyn column-major)
    __host__ __device__ static constexpr Unit fwd(int u)
    {
        return u < 6 ? ((u & 1) ? Unit{U_K, u >> 1} : Unit{U_A, u >> 1}) : u < 12 ? Unit{U_A, u - 3} : Unit{U_B, u - 12};
    }
    // backward: BR0 M0 BR1 M1 BR2 M2 QI M3 .. M8 KR0 KR1 KR2   (Bdyn row-major, AmBKt, Quu_inv column-major, Kinf row-major)
    __host__ __device__ static constexpr Unit bwd(int u)
    {
        return u < 6 ? ((u & 1) ? Unit{U_M, u >> 1} : Unit{U_BR, u >> 1}) : u == 6 ? Unit{U_QI, 0} : u < 13 ? Unit{U_M, u - 4} : Unit{U_KR, u - 13};
    }
    static constexpr int FWD_K_DONE = 5, BWD_BR_DONE = 4, BWD_QI_DONE = 6;   // the unit after which Kinf x / B^T p / Quu_inv s are complete
    // block offset (SysBlock<12,4>) of tensor-memory group g
    __host__ __device__ static constexpr int src(int g)
    {
        using SB = SysBlock<12, 4>;
        const Unit t = g < FWD_UNITS ? fwd(g) : bwd(g - FWD_UNITS);
        return (t.kind == U_K ? SB::K : t.kind == U_A ? SB::A : t.kind == U_B ? SB::B : t.kind == U_BR ? SB::Brm
              : t.kind == U_QI ? SB::Qi : t.kind == U_M ? SB::M : SB::Krm) + 16 * t.idx;
    }
};

// shared memory: the per-instance state of tmpc_kernel.cuh SmemLayout, the TMEM base slot, one mbarrier per warp and, per warp,
// NSLOT staging slots into which the blocks of newly claimed systems are bulk-copied (TMA) at refill
template <int NH> struct SysSmem {
    using L = SmemLayout<float, 12, 4, NH, 128>;
    static constexpr int NSLOT = 3;
    static constexpr int BLKB = SysBlock<12, 4>::STRIDE * 4;   // bytes of one system block
    static constexpr int SLOTB = BLKB + 16;                    // slot pitch: + 16 B so that the slots start in different banks
    static constexpr size_t TMSLOT = L::BYTES, BARS = TMSLOT + 16, STAGE = BARS + 4 * 8;
    static constexpr size_t BYTES = STAGE + size_t(4) * NSLOT * SLOTB;
    static_assert(BLKB % 16 == 0 && STAGE % 16 == 0, "bulk copies move 16-byte aligned multiples of 16 bytes");
    static_assert(BYTES <= 232448, "shared memory of one SM");
};

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

}  // namespace sysk

#ifndef TMPC_SYS_LOOKAHEAD
#define TMPC_SYS_LOOKAHEAD 1024   // claims ahead of the work counter whose system blocks are prefetched into L2
#endif

// CB: the box bounds are identical at every stage (tmpc_set_model checks the actual rows): fixed constant-bank operands instead
// of stage-indexed loads
template <int NH, bool FAST, bool WARM, bool CB>
__global__ void __launch_bounds__(128, 1)
admm_kernel_sys(const __grid_constant__ Model<float, 12, 4, NH> P, const __grid_constant__ SolveArgs<float> a)
{
    using namespace sysk;
    constexpr int NX = 12, NU = 4, BLOCK = 128, HX = NX / 2, HU = NU / 2;
    using SB = SysBlock<NX, NU>;
    using O = Orders<float, NX, NU>;
    using L = SmemLayout<float, NX, NU, NH, BLOCK>;
    using TM = TmMap;
    static_assert(O::Kx == ORD_SEQ && O::Ax == ORD_SEQ && O::Bu == ORD_SEQ && O::Btp == ORD_VECREDUX && O::Qs == ORD_SEQ &&
                  O::Mp == ORD_TREE && O::Ktr == ORD_VECREDUX, "the sweeps below spell out these orders");
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31;
    constexpr unsigned FULLM = 0xffffffffu;
    constexpr int XROW = NX * NH, UROW = NU * (NH - 1);

    unsigned char *sp = smem;
    typename L::SU sd(sp, tid); sp += L::SU::BYTES;
    typename L::SU sy(sp, tid); sp += L::SU::BYTES;
    typename L::SU sz(sp, tid); sp += L::SU::BYTES;
    typename L::SX sg(sp, tid); sp += L::SX::BYTES;
    typename L::SX sv(sp, tid); sp += L::SX::BYTES;
    typename L::SP spn(sp, tid);
    using SS = SysSmem<NH>;
    const int warp = tid >> 5;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(smem + SS::BARS + warp * 8);
    unsigned char *stage = smem + SS::STAGE + (size_t)warp * SS::NSLOT * SS::SLOTB;
    uint32_t bar_phase = 0;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t tcol;
    {
        uint32_t *slot = reinterpret_cast<uint32_t *>(smem + SS::TMSLOT);
        if (tid < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tcol = *slot + ((uint32_t)(((tid >> 5) & 3) * 32) << 16);
    }
    // (-0, -0), opaque to the compiler (a batch is never negative)
    const float nzs = __int_as_float((int)(0x80000000u ^ (unsigned)(a.batch < 0)));
    const float2 Z = mk2(nzs, nzs);

    const float *blk = a.sys;   // idle lanes keep a valid block (instance 0)
    float rho_l = P.rho, nrho_l = P.nrho;
    float2 Qd2[HX];
#pragma unroll
    for (int j = 0; j < HX; ++j) Qd2[j] = mk2(0.f, 0.f);
    long long inst = -1;
    int it = 0;
    int phase = PH_FREE;
    bool exhausted = false;
    int deferred = 0;       // warp-uniform: trips for which a refill of too few lanes has been postponed
    // refill policy (tuning: SolveArgs::test_flags bits 8-12 = fewest free lanes that refill at once, 16-19 = most trips they wait)
    const int refill_min = ((a.test_flags >> 8) & 31) ? ((a.test_flags >> 8) & 31) : 2;
    const int defer_max = ((a.test_flags >> 16) & 15) ? ((a.test_flags >> 16) & 15) : 1;
    float x0[NX];
    float res[4] = {0.f, 0.f, 0.f, 0.f};
    unsigned long long n_iter = 0, n_solved = 0, n_trips = 0, n_inst = 0;
#pragma unroll
    for (int j = 0; j < NX; ++j) x0[j] = 0.f;
    float cb[4][16];   // ring of coefficient groups: unit u of a sweep stage lands in cb[u & 3]

    for (;;) {
        // ------------------------------------------------------------------ lane refill
        const bool need = (phase == PH_FREE) && !exhausted;
        unsigned m = __ballot_sync(FULLM, need);
        {   // too few free lanes wait (a bounded number of trips) for more: the refill section runs for the whole warp
            const bool others_busy = __ballot_sync(FULLM, phase != PH_FREE) != 0;
            if (m && __popc(m) < refill_min && deferred < defer_max && others_busy) { ++deferred; m = 0; }
            else deferred = 0;
        }
        if (m) {
            const int leader = __ffs(m) - 1;
            const int cnt = __popc(m);
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(a.counter, (unsigned long long)cnt);
            base = __shfl_sync(FULLM, base, leader);
            // the systems that will be claimed TMPC_SYS_LOOKAHEAD claims from now: their blocks are prefetched into L2 at the end
            long long pf_inst = -1;
            if (!a.gate && (int)lane < cnt) {
                const long long pp = (long long)base + TMPC_SYS_LOOKAHEAD + lane;
                if (pp < a.batch) pf_inst = claimed_instance(a, pp);
            }
            bool fill = false;
            if (need) {
                const long long idx = (long long)base + __popc(m & ((1u << lane) - 1u));
                const long long ci = idx < a.batch ? claim_instance(a, idx) : -1;
                if (ci >= 0) {
                    fill = true;
                    inst = ci;
                    phase = PH_RUN;
                    it = 0;
                    res[0] = res[1] = res[2] = res[3] = 0.f;
                    blk = a.sys + inst * SB::STRIDE;
                    gload<float, NX>(a.x0 + inst * NX, x0);
                    if (WARM && a.wd) {
#pragma unroll 1
                        for (int i = 0; i < NH - 1; ++i) {
                            float t[NU];
                            gload<float, NU>(a.wd + inst * UROW + i * NU, t); sd.store(i, t);
                            gload<float, NU>(a.wy + inst * UROW + i * NU, t); sy.store(i, t);
                            gload<float, NU>(a.wz + inst * UROW + i * NU, t); sz.store(i, t);
                        }
#pragma unroll 1
                        for (int i = 0; i < NH; ++i) {
                            float t[NX];
                            gload<float, NX>(a.wg + inst * XROW + i * NX, t); sg.store(i, t);
                            gload<float, NX>(a.wv + inst * XROW + i * NX, t); sv.store(i, t);
                        }
                    } else {
                        float zu[NU], zx[NX];
#pragma unroll
                        for (int j = 0; j < NU; ++j) zu[j] = 0.f;
#pragma unroll
                        for (int j = 0; j < NX; ++j) zx[j] = 0.f;
#pragma unroll 1
                        for (int i = 0; i < NH - 1; ++i) { sd.store(i, zu); sy.store(i, zu); sz.store(i, zu); }
#pragma unroll 1
                        for (int i = 0; i < NH; ++i) { sg.store(i, zx); sv.store(i, zx); }
                    }
                } else {
                    exhausted = true;
                }
            }
            // The new systems' blocks, NSLOT lanes per pass: one TMA bulk copy per block into the warp's staging slots (one
            // global-memory latency for everything the lane needs), then the tensor-memory image is rewritten from there.
            unsigned fm = __ballot_sync(FULLM, fill);
            while (fm) {
                const int rank = __popc(fm & ((1u << lane) - 1u));
                const bool mine = ((fm >> lane) & 1u) && rank < SS::NSLOT;
                const int nmine = min(__popc(fm), SS::NSLOT);
                const float *sl = reinterpret_cast<const float *>(stage + (mine ? rank : 0) * SS::SLOTB);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the slots were read through the generic proxy
                __syncwarp();
                if (lane == 0)
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((uint32_t)(nmine * SS::BLKB)) : "memory");
                __syncwarp();
                if (mine)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"((uint32_t)__cvta_generic_to_shared(sl)), "l"(blk), "r"((uint32_t)SS::BLKB), "r"(bar) : "memory");
                {
                    uint32_t done = 0;
                    while (!done)
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                     : "=r"(done) : "r"(bar), "r"(bar_phase) : "memory");
                    bar_phase ^= 1u;
                }
                if (mine) {
                    rho_l = sl[SB::RHO];
                    nrho_l = -rho_l;
                    {
                        float qd[NX];
#pragma unroll
                        for (int w = 0; w < NX / 4; ++w) {
                            const float4 t = *reinterpret_cast<const float4 *>(sl + SB::Qd + 4 * w);
                            qd[4 * w] = t.x; qd[4 * w + 1] = t.y; qd[4 * w + 2] = t.z; qd[4 * w + 3] = t.w;
                        }
#pragma unroll
                        for (int j = 0; j < HX; ++j) Qd2[j] = mk2(qd[2 * j], qd[2 * j + 1]);
                    }
                    {   // p_N seed: -(Xref_{N-1}^T * Pinf)   (admm.cpp:83)
                        float xr[NX], pn[NX];
                        gload<float, NX>(a.Xref + inst * a.xref_stride + (NH - 1) * NX, xr);
#pragma unroll
                        for (int j = 0; j < NX; ++j) {
                            float c[NX];
#pragma unroll
                            for (int w = 0; w < NX / 4; ++w) {
                                const float4 t = *reinterpret_cast<const float4 *>(sl + SB::Pf + j * NX + 4 * w);
                                c[4 * w] = t.x; c[4 * w + 1] = t.y; c[4 * w + 2] = t.z; c[4 * w + 3] = t.w;
                            }
                            pn[j] = -dot<float, O::XtP, NX, FAST>([&](int k) { return c[k]; }, [&](int k) { return xr[k]; });
                        }
                        spn.store(0, pn);
                    }
                }
                // tcgen05 is warp-collective: every lane rewrites its columns, lanes that are not being filled with what they hold.
                // Four 16-column groups per tensor-memory round trip.
#pragma unroll
                for (int b = 0; b < (TM::GROUPS + 3) / 4; ++b) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (4 * b + q < TM::GROUPS) tm_ld16(tcol + 16 * (4 * b + q), cb[q]);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (4 * b + q < TM::GROUPS) wait16(cb[q]);
                    if (mine) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (4 * b + q < TM::GROUPS) {
                                const float4 *src = reinterpret_cast<const float4 *>(sl + TM::src(4 * b + q));
#pragma unroll
                                for (int w = 0; w < 4; ++w) {
                                    const float4 t = src[w];
                                    cb[q][4 * w] = t.x; cb[q][4 * w + 1] = t.y; cb[q][4 * w + 2] = t.z; cb[q][4 * w + 3] = t.w;
                                }
                            }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (4 * b + q < TM::GROUPS) tm_st16(tcol + 16 * (4 * b + q), cb[q]);
                }
                tm_wait_st();
#pragma unroll
                for (int k = 0; k < SS::NSLOT; ++k) fm &= fm - 1u;   // the lanes this pass served
            }
            // L2 prefetch of the blocks (30 lines each), x0 and Xref rows of the systems claimed TMPC_SYS_LOOKAHEAD claims from now
            for (int q = 0; q < cnt; ++q) {
                const long long pi = __shfl_sync(FULLM, pf_inst, q);
                if (pi >= 0) {
                    if (lane < (unsigned)(SS::BLKB / 128)) prefetch_l2(reinterpret_cast<const char *>(a.sys + pi * SB::STRIDE) + lane * 128);
                    else if (lane == 30) prefetch_l2(a.x0 + pi * NX);
                    else if (a.xref_stride) prefetch_l2(a.Xref + pi * a.xref_stride + (NH - 1) * NX);
                }
            }
        }
        if (__all_sync(FULLM, phase == PH_FREE)) break;
        ++n_trips;

        const bool emit = (phase == PH_EMIT);
        if (phase == PH_RUN) ++it;

        // the forward sweep's first three coefficient groups
        tm_ld16(tcol + 0, cb[0]);
        tm_ld16(tcol + 16, cb[1]);
        tm_ld16(tcol + 32, cb[2]);

        // ------------------------------------------------------------------ forward sweep
        // forward_pass (admm.cpp:27-37) + update_slack (:45-61) + update_dual (:67-71) + residual maxima (:95-98)
        float pri_x = 0.f, dua_x = 0.f, pri_u = 0.f, dua_u = 0.f;
        {
            float2 x2[HX];
#pragma unroll
            for (int j = 0; j < HX; ++j) x2[j] = mk2(x0[2 * j], x0[2 * j + 1]);
            float *xo = (emit && a.x) ? a.x + inst * XROW : nullptr;
            float *uo = (emit && a.u) ? a.u + inst * UROW : nullptr;
            float *go = (WARM && emit && a.wg) ? a.wg + inst * XROW : nullptr;
            float *yo = (WARM && emit && a.wy) ? a.wy + inst * UROW : nullptr;
            auto xs = [&](int k) -> float { return (k & 1) ? x2[k >> 1].y : x2[k >> 1].x; };

            // slack / dual / residuals of rows (2j, 2j+1) of the state part of stage i (:48, :59, :70, :95, :96).  The element-wise
            // work is spread over the coefficient units (one row pair per unit) so that its min / max / constant-bank
            // instructions issue in the shadow of the packed FMA-pipe instructions instead of in a block of their own.
            float g[NX], v[NX], d[NU], y[NU], z[NU];
            auto state_pair = [&](int i, int j) {
                const int bi = CB ? 0 : i * NX;
                const float2 gj = mk2(g[2 * j], g[2 * j + 1]), vj = mk2(v[2 * j], v[2 * j + 1]);
                const float2 xg = ad2(x2[j], gj);
                float2 vn;
                vn.x = fminf(P.xmax[bi + 2 * j], fmaxf(P.xmin[bi + 2 * j], xg.x));
                vn.y = fminf(P.xmax[bi + 2 * j + 1], fmaxf(P.xmin[bi + 2 * j + 1], xg.y));
                const float2 rp = sb2(x2[j], vn), rd = sb2(vj, vn);
                pri_x = fmaxf(pri_x, fmaxf(fabsf(rp.x), fabsf(rp.y)));
                dua_x = fmaxf(dua_x, fmaxf(fabsf(rd.x), fabsf(rd.y)));
                const float2 gn = sb2(xg, vn);   // (g + x) - vnew: the sum is the one above (addition commutes bit for bit)
                g[2 * j] = gn.x; g[2 * j + 1] = gn.y;
                v[2 * j] = vn.x; v[2 * j + 1] = vn.y;
            };
            auto state_out = [&](int i) {
                sg.store(i, g);
                sv.store(i, v);
                if (xo) {
                    float xx[NX];
#pragma unroll
                    for (int j = 0; j < HX; ++j) { xx[2 * j] = x2[j].x; xx[2 * j + 1] = x2[j].y; }
                    gstore<float, NX>(xo + i * NX, xx);
                }
            };
            // a stage's state is fetched late in the stage before it (the tensor-memory asm statements are memory barriers to the
            // compiler: it cannot hoist these loads itself), into the registers that stage has finished with
            sg.load(0, g);
            sv.load(0, v);
            sd.load(0, d);
            sy.load(0, y);
            sz.load(0, z);

#pragma unroll 1
            for (int i = 0; i < NH - 1; ++i) {
                float2 kx2[HU], u2[HU], ax2[HX], bu2[HX];
                auto us = [&](int k) -> float { return (k & 1) ? u2[k >> 1].y : u2[k >> 1].x; };
                // what the last three slots of this stage fetch: the next stage's first groups, or the backward sweep's
                const uint32_t nxt = tcol + (i < NH - 2 ? 0u : (uint32_t)TM::BWD_COL);
                if (WARM && go) gstore<float, NX>(go + i * NX, g);
                if (WARM && yo) gstore<float, NU>(yo + i * NU, y);
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    if (u < TM::FWD_UNITS) wait16(cb[u & 3]);
                    if (u + 3 < TM::FWD_UNITS) tm_ld16(tcol + 16 * (u + 3), cb[(u + 3) & 3]);
                    else if (u + 3 >= 16) tm_ld16(nxt + 16 * (u + 3 - 16), cb[(u + 3) & 3]);
                    if (u >= 1 && u <= HX) state_pair(i, u - 1);
                    if (u == HX + 1) state_out(i);
                    if (u == 12) {   // the next stage's state
                        sg.load(i + 1, g);
                        sv.load(i + 1, v);
                        if (i + 1 < NH - 1) { sd.load(i + 1, d); sy.load(i + 1, y); sz.load(i + 1, z); }
                    }
                    if (u < TM::FWD_UNITS) {
                        const Unit t = TM::fwd(u);
#pragma unroll
                        for (int e = 0; e < 16; e += 2) {
                            const float2 c = mk2(cb[u & 3][e], cb[u & 3][e + 1]);
                            const int E = 16 * t.idx + e;
                            if (t.kind == U_K) {                               // Kinf x, rows (2h, 2h+1), column k     (:31)
                                const int k = E / NU, h = (E % NU) / 2;
                                if (k == 0) kx2[h] = FAST ? ml2(c, xs(0)) : pr2(c, xs(0), Z);
                                else kx2[h] = FAST ? fm2(c, xs(k), kx2[h]) : ad2(pr2(c, xs(k), Z), kx2[h]);
                            } else if (t.kind == U_A) {                        // Adyn x                                (:35)
                                const int k = E / NX, j = (E % NX) / 2;
                                if (k == 0) ax2[j] = FAST ? ml2(c, xs(0)) : pr2(c, xs(0), Z);
                                else ax2[j] = FAST ? fm2(c, xs(k), ax2[j]) : ad2(pr2(c, xs(k), Z), ax2[j]);
                            } else {                                           // Bdyn u                                (:35)
                                const int k = E / NX, j = (E % NX) / 2;
                                if constexpr (FAST) ax2[j] = fm2(c, us(k), ax2[j]);
                                else if (k == 0) bu2[j] = pr2(c, us(0), Z);
                                else bu2[j] = ad2(pr2(c, us(k), Z), bu2[j]);
                            }
                        }
                    }
                    if (u == TM::FWD_K_DONE) {   // Kinf x complete: input, slack, dual, residuals of stage i
                        const int bi = CB ? 0 : i * NU;
#pragma unroll
                        for (int h = 0; h < HU; ++h) {
                            const float2 dh = mk2(d[2 * h], d[2 * h + 1]), yh = mk2(y[2 * h], y[2 * h + 1]), zh = mk2(z[2 * h], z[2 * h + 1]);
                            u2[h] = sb2(ng2(kx2[h]), dh);                                                        // :31
                            const float2 uy = ad2(u2[h], yh);                                                    // :47
                            float2 zn;
                            zn.x = fminf(P.umax[bi + 2 * h], fmaxf(P.umin[bi + 2 * h], uy.x));                   // :53
                            zn.y = fminf(P.umax[bi + 2 * h + 1], fmaxf(P.umin[bi + 2 * h + 1], uy.y));
                            const float2 rp = sb2(u2[h], zn), rd = sb2(zh, zn);
                            pri_u = fmaxf(pri_u, fmaxf(fabsf(rp.x), fabsf(rp.y)));                               // :97
                            dua_u = fmaxf(dua_u, fmaxf(fabsf(rd.x), fabsf(rd.y)));                               // :98
                            const float2 yn = sb2(uy, zn);                                                       // :69  (y + u) - znew
                            y[2 * h] = yn.x; y[2 * h + 1] = yn.y;
                            z[2 * h] = zn.x; z[2 * h + 1] = zn.y;
                        }
                        sy.store(i, y);
                        sz.store(i, z);
                        if (uo || (emit && a.u0 && i == 0)) {
                            float uu[NU];
#pragma unroll
                            for (int h = 0; h < HU; ++h) { uu[2 * h] = u2[h].x; uu[2 * h + 1] = u2[h].y; }
                            if (uo) gstore<float, NU>(uo + i * NU, uu);
                            if (emit && a.u0 && i == 0) gstore<float, NU>(a.u0 + inst * NU, uu);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < HX; ++j) x2[j] = FAST ? ax2[j] : ad2(ax2[j], bu2[j]);                       // :35
            }
            if (WARM && go) gstore<float, NX>(go + (NH - 1) * NX, g);
#pragma unroll
            for (int j = 0; j < HX; ++j) state_pair(NH - 1, j);
            state_out(NH - 1);
        }

        // ------------------------------------------------------------------ termination (admm.cpp:91-109, :135-138)
        bool final_bwd = false;  // WARM: max_iter exit still runs the backward pass of its last iteration
        if (phase == PH_RUN) {
            const bool chk = (it % P.check_term) == 0;
            if (chk) {
                res[0] = pri_x;
                res[1] = __fmul_rn(dua_x, rho_l);
                res[2] = pri_u;
                res[3] = __fmul_rn(dua_u, rho_l);
            }
            const bool conv = chk && res[0] < P.pri_tol && res[2] < P.pri_tol && res[1] < P.dua_tol && res[3] < P.dua_tol;
            if (conv || it >= P.max_iter) {
                if (a.iter) a.iter[inst] = it;
                if (a.status) a.status[inst] = conv ? 1 : 11;
                if (a.resid) {
                    a.resid[inst * 4 + 0] = res[0];
                    a.resid[inst * 4 + 1] = res[1];
                    a.resid[inst * 4 + 2] = res[2];
                    a.resid[inst * 4 + 3] = res[3];
                }
                n_iter += (unsigned)it;
                n_solved += conv ? 1u : 0u;
                ++n_inst;
                final_bwd = !conv;
                phase = PH_EMIT;
            }
        } else if (phase == PH_EMIT) {
            phase = PH_FREE;  // trajectory was written by this trip's forward sweep
            if (a.done) { __threadfence(); atomicAdd(a.done + (inst >> a.done_shift), 1u); }
        }

        // ------------------------------------------------------------------ backward sweep
        // update_linear_cost (admm.cpp:77-85) recomputed per stage + backward_pass_grad (:15-22)
        const bool cont = (phase == PH_RUN);
        const bool wout = WARM && (cont || final_bwd) && a.wd;
        if (__any_sync(FULLM, cont || wout)) {
            float2 p2[HX];
            auto ps = [&](int k) -> float { return (k & 1) ? p2[k >> 1].y : p2[k >> 1].x; };
            const float *xr_base = a.Xref + (inst < 0 ? 0 : inst) * a.xref_stride;
            float *wdo = wout ? a.wd + inst * UROW : nullptr;
            float *wvo = wout ? a.wv + inst * XROW : nullptr;
            float *wzo = wout ? a.wz + inst * UROW : nullptr;
            const float2 rho2 = mk2(rho_l, rho_l), nrho2 = mk2(nrho_l, nrho_l);
            {
                float v[NX], g[NX], pn[NX];
                sv.load(NH - 1, v);
                sg.load(NH - 1, g);
                spn.load(0, pn);
                if (WARM && wvo) gstore<float, NX>(wvo + (NH - 1) * NX, v);
#pragma unroll
                for (int j = 0; j < HX; ++j) {
                    const float2 dv = sb2(mk2(v[2 * j], v[2 * j + 1]), mk2(g[2 * j], g[2 * j + 1]));
                    const float2 pj = mk2(pn[2 * j], pn[2 * j + 1]);
                    if constexpr (FAST) p2[j] = __ffma2_rn(nrho2, dv, pj);
                    else p2[j] = sb2(pj, pp2(rho2, dv, Z));                                                      // :84
                }
            }
            // a stage's state and reference row are fetched late in the stage before it (see the forward sweep)
            float z[NU], y[NU], v[NX], g[NX], xr[NX];
            sz.load(NH - 2, z);
            sy.load(NH - 2, y);
            sv.load(NH - 2, v);
            sg.load(NH - 2, g);
            gload<float, NX>(xr_base + (NH - 2) * NX, xr);
#pragma unroll 1
            for (int i = NH - 2; i >= 0; --i) {
                float2 r2[HU], q2[HX], s2[HU], d2[HU], mp2[HX], kr2[HX];
                float2 e0[4][HU], e1[4][HU];          // B^T p: vectorised redux, lane L = k % 4
                float2 t0[HX], t1[HX], ta[HX], tl[HX];  // AmBKt p: half-split tree over 12 = ((3 + 3) + (3 + 3)), 3 = e + (e + e)
                float2 k0[HX], k1[HX];                 // Kinf^T r: (e0 + e2) + (e1 + e3)
                auto rs = [&](int k) -> float { return (k & 1) ? r2[k >> 1].y : r2[k >> 1].x; };
                auto ss = [&](int k) -> float { return (k & 1) ? s2[k >> 1].y : s2[k >> 1].x; };
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    wait16(cb[u & 3]);
                    if (u + 3 < 16) tm_ld16(tcol + TM::BWD_COL + 16 * (u + 3), cb[(u + 3) & 3]);
                    else if (i > 0) tm_ld16(tcol + TM::BWD_COL + 16 * (u + 3 - 16), cb[(u + 3) & 3]);
                    if (u == 12 && i > 0) {   // the next stage's state (this one's was consumed at u == 1)
                        sz.load(i - 1, z);
                        sy.load(i - 1, y);
                        sv.load(i - 1, v);
                        sg.load(i - 1, g);
                        gload<float, NX>(xr_base + (i - 1) * NX, xr);
                    }
                    if (u == 1) {   // r_i, q_i from (z, y, v, g, Xref)
                        if (WARM && wvo) { gstore<float, NX>(wvo + i * NX, v); gstore<float, NU>(wzo + i * NU, z); }
#pragma unroll
                        for (int h = 0; h < HU; ++h)
                            r2[h] = pp2(nrho2, sb2(mk2(z[2 * h], z[2 * h + 1]), mk2(y[2 * h], y[2 * h + 1])), Z);   // :80
#pragma unroll
                        for (int j = 0; j < HX; ++j) {
                            const float2 dv = sb2(mk2(v[2 * j], v[2 * j + 1]), mk2(g[2 * j], g[2 * j + 1]));
                            const float2 cq = ng2(pp2(mk2(xr[2 * j], xr[2 * j + 1]), Qd2[j], Z));                   // :81
                            if constexpr (FAST) q2[j] = __ffma2_rn(nrho2, dv, cq);
                            else q2[j] = sb2(cq, pp2(rho2, dv, Z));                                              // :82
                        }
                    }
                    const Unit t = TM::bwd(u);
#pragma unroll
                    for (int e = 0; e < 16; e += 2) {
                        const float2 c = mk2(cb[u & 3][e], cb[u & 3][e + 1]);
                        const int E = 16 * t.idx + e;
                        if (t.kind == U_BR) {                              // Bdyn^T p, rows (2h, 2h+1), term k       (:19)
                            const int k = E / NU, h = (E % NU) / 2;
                            if constexpr (FAST) {
                                s2[h] = k == 0 ? ml2(c, ps(0)) : fm2(c, ps(k), s2[h]);
                            } else {
                                const int Lk = k % 4, qk = k / 4;
                                const float2 pr = pr2(c, ps(k), Z);
                                if (qk == 0) e0[Lk][h] = pr;
                                else if (qk == 1) e1[Lk][h] = pr;
                                else e0[Lk][h] = ad2(e0[Lk][h], ad2(e1[Lk][h], pr));
                            }
                        } else if (t.kind == U_QI) {                       // Quu_inv s, column k                      (:19)
                            const int k = E / NU, h = (E % NU) / 2;
                            if (k == 0) d2[h] = FAST ? ml2(c, ss(0)) : pr2(c, ss(0), Z);
                            else d2[h] = FAST ? fm2(c, ss(k), d2[h]) : ad2(pr2(c, ss(k), Z), d2[h]);
                        } else if (t.kind == U_M) {                        // AmBKt p, rows (2j, 2j+1), term k         (:20)
                            const int k = E / NX, j = (E % NX) / 2;
                            if constexpr (FAST) {
                                mp2[j] = k == 0 ? ml2(c, ps(0)) : fm2(c, ps(k), mp2[j]);
                            } else {
                                const int mk = k % 3, qk = k / 3;
                                const float2 pr = pr2(c, ps(k), Z);
                                if (mk == 0) t0[j] = pr;
                                else if (mk == 1) t1[j] = pr;
                                else {
                                    const float2 tt = ad2(t0[j], ad2(t1[j], pr));
                                    if (qk == 0) ta[j] = tt;
                                    else if (qk == 1) tl[j] = ad2(ta[j], tt);
                                    else if (qk == 2) ta[j] = tt;
                                    else mp2[j] = ad2(tl[j], ad2(ta[j], tt));
                                }
                            }
                        } else {                                           // Kinf^T r, rows (2j, 2j+1), term k        (:20)
                            const int k = E / NX, j = (E % NX) / 2;
                            if constexpr (FAST) {
                                kr2[j] = k == 0 ? ml2(c, rs(0)) : fm2(c, rs(k), kr2[j]);
                            } else {
                                const float2 pr = pr2(c, rs(k), Z);
                                if (k == 0) k0[j] = pr;
                                else if (k == 1) k1[j] = pr;
                                else if (k == 2) k0[j] = ad2(k0[j], pr);
                                else kr2[j] = ad2(k0[j], ad2(k1[j], pr));
                            }
                        }
                    }
                    if (u == TM::BWD_BR_DONE) {   // B^T p complete
#pragma unroll
                        for (int h = 0; h < HU; ++h) {
                            if constexpr (!FAST) s2[h] = ad2(ad2(e0[0][h], e0[2][h]), ad2(e0[1][h], e0[3][h]));
                            s2[h] = ad2(s2[h], r2[h]);
                        }
                    }
                    if (u == TM::BWD_QI_DONE) {   // d_i = Quu_inv (B^T p + r)
                        float d[NU];
#pragma unroll
                        for (int h = 0; h < HU; ++h) { d[2 * h] = d2[h].x; d[2 * h + 1] = d2[h].y; }
                        sd.store(i, d, cont);
                        if (WARM && wdo) gstore<float, NU>(wdo + i * NU, d);
                    }
                }
#pragma unroll
                for (int j = 0; j < HX; ++j) p2[j] = sb2(ad2(q2[j], mp2[j]), kr2[j]);                            // :20
            }
        } else {
            // nobody sweeps backward: drain the three groups the forward sweep fetched ahead for it
            wait16(cb[0]);
            wait16(cb[1]);
            wait16(cb[2]);
        }
    }

    // ---------------------------------------------------------------------- statistics
    if (a.stats) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_iter += __shfl_down_sync(FULLM, n_iter, o);
            n_solved += __shfl_down_sync(FULLM, n_solved, o);
            n_trips += __shfl_down_sync(FULLM, n_trips, o);
            n_inst += __shfl_down_sync(FULLM, n_inst, o);
        }
        if (lane == 0) {
            atomicAdd(a.stats + 0, n_iter);
            atomicAdd(a.stats + 1, n_solved);
            atomicAdd(a.stats + 2, n_trips);
            atomicAdd(a.stats + 3, n_inst);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tcol) : "memory");
}

}  // namespace tmpc
