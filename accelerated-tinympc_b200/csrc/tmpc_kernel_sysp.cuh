// Per-instance SYSTEMS, fp32 12/4/N -- TWO LANES PER INSTANCE (third generation; tmpc_kernel_sys.cuh is the second).
//
// Why.  A lane needs its instance's 496 coefficients every horizon stage and they only fit in tensor memory, which gives 128
// instances per SM.  With a thread per instance that is 128 threads = one warp per scheduler, and a single warp cannot fill
// the FMA pipe: a packed FFMA2 / FADD2 occupies the warp's issue slot for two cycles, every other instruction for one, and
// every stall is exposed (profiles/r02_ncu_systems_kernel.md: issue = instructions + packed instructions + stalls, pipe 52 %).
// Here the SAME 128 instances per SM are worked by 256 threads: lanes (2t, 2t+1) of a warp share instance t of that warp,
// each owning half of the OUTPUT ROWS of every product (x rows 6h..6h+5, u rows 2h..2h+1, h = lane & 1).  Two warps per
// scheduler: one warp's packed instruction issues in the shadow of the other's, stalls overlap.
//
//  * tensor memory: warps w and w+4 address the same 32 TMEM lanes, so a thread owns 256 of its lane's 512 columns -- exactly
//    the 248 coefficients of its half (row pairs adjacent, streamed in the order the sweeps consume them) + 8 of padding;
//  * every mat-vec is a column sweep over the thread's own row pairs in the reference build's summation order (Orders<float,
//    12, 4>: sequential / half-split tree / vectorised redux), exact packed products (fma(a, b, -0)): bit-identical results;
//  * a product needs the whole input vector: after each stage the two lanes exchange their halves of x (p), u, s, r with
//    warp shuffles (36 per stage pair); the four residual maxima are combined across the pair before the termination test;
//  * state (d, y, z, g, v, p_N seed): the thread's own rows in shared memory, 8-byte chunks laid out [chunk][thread];
//  * refill: one TMA bulk copy of the new system's block per instance into a staging slot, from which both lanes rewrite
//    their tensor-memory columns; the blocks that will be claimed ~1000 claims later are prefetched into L2.
#pragma once
#include "tmpc_kernel_sys.cuh"

namespace tmpc {
namespace sysk {

// Shared-memory accesses by 32-bit shared-space address (the generic-pointer form makes the compiler rebuild the shared window base
// -- S2UR SR_CgaCtaId + ULEA, a scoreboard wait each time -- inside the sweeps).  volatile + "memory": they stay where they are written,
// like the tensor-memory statements around them.
__device__ __forceinline__ float4 lds128(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float2 lds64(uint32_t a)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float2 lo, float2 hi)
{
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(a), "f"(lo.x), "f"(lo.y), "f"(hi.x), "f"(hi.y) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t a, float2 v) { asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"(a), "f"(v.x), "f"(v.y) : "memory"); }

// The thread's own rows of the per-instance state, [chunk][thread] so that every warp access is conflict-free:
//   g, v : 12 floats per stage = three 16-byte chunks (g0..g3 | g4 g5 v0 v1 | v2..v5)
//   d, y : one 16-byte chunk per stage;  z : one 8-byte chunk per stage;  p_N seed: three 8-byte chunks
template <int NH, int BLOCK> struct PairState {
    static constexpr uint32_t GV = 0, DY = GV + 3u * NH * BLOCK * 16, ZZ = DY + (NH - 1) * BLOCK * 16u, PN = ZZ + (NH - 1) * BLOCK * 8u,
                              BYTES = PN + 3u * BLOCK * 8;
    uint32_t gv, dy, zz, pn;
    __device__ __forceinline__ PairState(unsigned char *smem, int tid)
    {
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(smem);
        gv = b + GV + tid * 16; dy = b + DY + tid * 16; zz = b + ZZ + tid * 8; pn = b + PN + tid * 8;
        // opaque to the compiler: otherwise it re-derives the shared window base (S2UR SR_CgaCtaId, a scoreboard wait) inside the sweeps
        asm volatile("" : "+r"(gv), "+r"(dy), "+r"(zz), "+r"(pn));
    }
    __device__ __forceinline__ void load_gv(int i, float2 (&o)[6]) const
    {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float4 t = lds128(gv + (uint32_t)(i * 3 + c) * (BLOCK * 16));
            o[2 * c] = make_float2(t.x, t.y); o[2 * c + 1] = make_float2(t.z, t.w);
        }
    }
    __device__ __forceinline__ void store_gv(int i, const float2 (&o)[6]) const
    {
#pragma unroll
        for (int c = 0; c < 3; ++c) sts128(gv + (uint32_t)(i * 3 + c) * (BLOCK * 16), o[2 * c], o[2 * c + 1]);
    }
    __device__ __forceinline__ void load_dyz(int i, float2 (&o)[3]) const
    {
        const float4 t = lds128(dy + (uint32_t)i * (BLOCK * 16));
        o[0] = make_float2(t.x, t.y); o[1] = make_float2(t.z, t.w);
        o[2] = lds64(zz + (uint32_t)i * (BLOCK * 8));
    }
    __device__ __forceinline__ void store_dyz(int i, const float2 (&o)[3]) const
    {
        sts128(dy + (uint32_t)i * (BLOCK * 16), o[0], o[1]);
        sts64(zz + (uint32_t)i * (BLOCK * 8), o[2]);
    }
    __device__ __forceinline__ void store_yz(int i, float2 y, float2 z) const
    {
        sts64(dy + (uint32_t)i * (BLOCK * 16) + 8, y);
        sts64(zz + (uint32_t)i * (BLOCK * 8), z);
    }
    __device__ __forceinline__ void store_d(int i, float2 d, bool pred) const { if (pred) sts64(dy + (uint32_t)i * (BLOCK * 16), d); }
    __device__ __forceinline__ void load_pn(float2 (&o)[3]) const
    {
#pragma unroll
        for (int c = 0; c < 3; ++c) o[c] = lds64(pn + (uint32_t)c * (BLOCK * 8));
    }
    __device__ __forceinline__ void store_pn(const float2 (&o)[3]) const
    {
#pragma unroll
        for (int c = 0; c < 3; ++c) sts64(pn + (uint32_t)c * (BLOCK * 8), o[c]);
    }
};

// A thread's half of the tensor-memory image as a stream of coefficient PAIRS (rows (r, r+1) of one column / term k), eight
// pairs per 16-column unit: forward units 0..7 (pairs 0..63, 60..63 padding), backward units 8..15.
//   forward:  for k = 0..11: A(k, 0..2) K(k);  then B(k, j) for k = 0..3, j = 0..2
//   backward: for k = 0..11: M(k, 0..2) BR(k); then QI(0..3); then KR(k, j) for k = 0..3, j = 0..2
// The two-chain products (Kinf x; B^T p) ride along with the three-chain ones so that the scheduler always has independent work.
struct PairMap {
    using SB = SysBlock<12, 4>;
    static constexpr int UNITS = 8;   // per sweep
    __host__ __device__ static constexpr Unit fwd(int q)
    {
        return q < 48 ? ((q & 3) < 3 ? Unit{U_A, (q >> 2) * 3 + (q & 3)} : Unit{U_K, q >> 2}) : q < 60 ? Unit{U_B, q - 48} : Unit{-1, 0};
    }
    __host__ __device__ static constexpr Unit bwd(int q)
    {
        return q < 48 ? ((q & 3) < 3 ? Unit{U_M, (q >> 2) * 3 + (q & 3)} : Unit{U_BR, q >> 2}) : q < 52 ? Unit{U_QI, q - 48} : Unit{U_KR, q - 52};
    }
    static constexpr int FWD_K_DONE = 5, BWD_BR_DONE = 5, BWD_QI_DONE = 6;   // the unit that completes Kinf x / B^T p / Quu_inv s
    // block offset of the pair for the half h = 0, and what h = 1 adds
    __host__ __device__ static constexpr int c0(Unit t)
    {
        return t.kind == U_K ? SB::K + 4 * t.idx : t.kind == U_BR ? SB::Brm + 4 * t.idx : t.kind == U_QI ? SB::Qi + 4 * t.idx
             : (t.kind == U_A ? SB::A : t.kind == U_B ? SB::B : t.kind == U_M ? SB::M : SB::Krm) + 12 * (t.idx / 3) + 2 * (t.idx % 3);
    }
    __host__ __device__ static constexpr int c1(Unit t) { return (t.kind == U_K || t.kind == U_BR || t.kind == U_QI) ? 2 : 6; }
};

// What this kernel reads of a system block, as three contiguous ranges (the row-major copies Arm, Qirm, Mrm of the first
// kernels are skipped), and where an offset of the block lands in a staging slot
struct SlotMap {
    using SB = SysBlock<12, 4>;
    static constexpr int R0 = SB::Krm, L0 = 48;                       // Kinf row-major
    static constexpr int R1 = SB::Brm, L1 = SB::B + 48 - SB::Brm;     // Bdyn row-major, Bdyn column-major
    static constexpr int R2 = SB::K, L2 = SB::STRIDE - SB::K;         // Kinf, Adyn, Quu_inv, AmBKt column-major, Pinf, Q, rho
    static constexpr int LEN = L0 + L1 + L2;
    static_assert(SB::Arm == R0 + L0 && SB::Qirm == R1 + L1 && SB::A == SB::K + 48, "SysBlock layout this map was written for");
    __host__ __device__ static constexpr int at(int off) { return off < R0 + L0 ? off - R0 : off < R1 + L1 ? off - R1 + L0 : off - R2 + L0 + L1; }
};

template <int NH> struct SysPSmem {
    static constexpr int BLOCK = 256, WARPS = 8, NSLOT = 2;
    using ST = PairState<NH, BLOCK>;
    static constexpr int BLKB = SlotMap::LEN * 4;   // bytes staged per system
    static constexpr int SLOTB = BLKB + 16;
    static constexpr size_t STATE = ST::BYTES;
    static constexpr size_t TMSLOT = STATE, BARS = TMSLOT + 16, STAGE = BARS + WARPS * 8;
    static constexpr size_t BYTES = STAGE + size_t(WARPS) * NSLOT * SLOTB;
    static_assert(BLKB % 16 == 0 && STAGE % 16 == 0, "bulk copies move 16-byte aligned multiples of 16 bytes");
};

}  // namespace sysk

#ifndef TMPC_SPEC_FACTOR_SYS
#define TMPC_SPEC_FACTOR_SYS 2.0f
#endif

template <int NH, bool FAST, bool WARM, bool CB>
__global__ void __launch_bounds__(256, 1)
admm_kernel_sysp(const __grid_constant__ Model<float, 12, 4, NH> P, const __grid_constant__ SolveArgs<float> a)
{
    using namespace sysk;
    constexpr int NX = 12, NU = 4, OX = 3;   // OX: own row pairs of an nx-vector (own rows of an nu-vector: one pair)
    using SB = SysBlock<NX, NU>;
    using O = Orders<float, NX, NU>;
    using SS = SysPSmem<NH>;
    using PM = PairMap;
    using SL = SlotMap;
    static_assert(O::Kx == ORD_SEQ && O::Ax == ORD_SEQ && O::Bu == ORD_SEQ && O::Btp == ORD_VECREDUX && O::Qs == ORD_SEQ &&
                  O::Mp == ORD_TREE && O::Ktr == ORD_VECREDUX, "the sweeps below spell out these orders");
    static_assert(SS::BYTES <= 232448, "shared memory of one SM");
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31;
    const int warp = tid >> 5;
    const int h = lane & 1;                       // which half of the rows this lane owns
    const unsigned pe = lane & ~1u;   // the pair's even lane
    constexpr unsigned FULLM = 0xffffffffu, EVEN = 0x55555555u;
    constexpr int XROW = NX * NH, UROW = NU * (NH - 1);

    const typename SS::ST st(smem, tid);
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
        // warps w and w + 4 own the same TMEM lanes: each takes 256 of the 512 columns
        tcol = *slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
    }
    const float nzs = __int_as_float((int)(0x80000000u ^ (unsigned)(a.batch < 0)));   // -0, opaque to the compiler
    const float2 Z = mk2(nzs, nzs);

    // whole vectors from the pair's halves (even lane: low rows, odd lane: high rows): one butterfly shuffle per float of the
    // partner's half, then selects (a shuffle costs several issue cycles, a select one)
    const bool odd = h != 0;
    auto px = [&](float vv) -> float { return __shfl_xor_sync(FULLM, vv, 1); };
    auto gather12 = [&](const float2 (&own)[OX], float (&full)[NX]) {
#pragma unroll
        for (int t = 0; t < OX; ++t) {
            const float ox = px(own[t].x), oy = px(own[t].y);
            full[2 * t] = odd ? ox : own[t].x; full[2 * t + 1] = odd ? oy : own[t].y;
            full[6 + 2 * t] = odd ? own[t].x : ox; full[6 + 2 * t + 1] = odd ? own[t].y : oy;
        }
    };
    auto gather4 = [&](float2 own, float (&full)[NU]) {
        const float ox = px(own.x), oy = px(own.y);
        full[0] = odd ? ox : own.x; full[1] = odd ? oy : own.y; full[2] = odd ? own.x : ox; full[3] = odd ? own.y : oy;
    };

    // CB: the (stage-invariant) box of the lane's own rows lives in registers
    float2 cxlo[OX], cxhi[OX], culo, cuhi;
#pragma unroll
    for (int j = 0; j < OX; ++j) {
        cxlo[j] = mk2(P.xmin[6 * h + 2 * j], P.xmin[6 * h + 2 * j + 1]);
        cxhi[j] = mk2(P.xmax[6 * h + 2 * j], P.xmax[6 * h + 2 * j + 1]);
    }
    culo = mk2(P.umin[2 * h], P.umin[2 * h + 1]);
    cuhi = mk2(P.umax[2 * h], P.umax[2 * h + 1]);

    const float *blk = a.sys;
    float rho_l = P.rho, nrho_l = P.nrho;
    float2 Qd2[OX], x0o[OX];
#pragma unroll
    for (int j = 0; j < OX; ++j) { Qd2[j] = mk2(0.f, 0.f); x0o[j] = mk2(0.f, 0.f); }
    long long inst = -1;
    int it = 0;
    int phase = PH_FREE;
    bool exhausted = false;
    int deferred = 0;
    const int refill_min = ((a.test_flags >> 8) & 31) ? ((a.test_flags >> 8) & 31) : 2;
    const int defer_max = ((a.test_flags >> 16) & 15) ? ((a.test_flags >> 16) & 15) : 1;
    float res[4] = {0.f, 0.f, 0.f, 0.f};
    // cold solves: a trip that is likely to be the instance's last (the previous check left every residual within
    // TMPC_SPEC_FACTOR_SYS of its tolerance, or it is the last allowed iteration) stores x / u as it goes, so that an instance
    // that does terminate there needs no emission trip.  A wrong guess costs the stores; they are overwritten later.
    bool spec = false;
    bool counted = true;   // the completion counter (SolveArgs::done) has been bumped for `inst`
    unsigned long long n_iter = 0, n_solved = 0, n_trips = 0, n_inst = 0;
    float cb[4][16];   // ring of coefficient units: unit u of a sweep stage lands in cb[u & 3]

    for (;;) {
        // ------------------------------------------------------------------ refill (per PAIR; both lanes hold the same phase / inst)
        const bool need = (phase == PH_FREE) && !exhausted;
        unsigned m = __ballot_sync(FULLM, need) & EVEN;
        {
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
            long long pf_inst = -1;
            if (!a.gate && (int)lane < cnt) {
                const long long pp = (long long)base + TMPC_SYS_LOOKAHEAD + lane;
                if (pp < a.batch) pf_inst = claimed_instance(a, pp);
            }
            bool fill = false;
            if (need && ((m >> pe) & 1u)) {
                const long long idx = (long long)base + __popc(m & ((1u << pe) - 1u));
                const long long ci = idx < a.batch ? claim_instance(a, idx) : -1;
                if (ci >= 0) {
                    fill = true;
                    inst = ci;
                    phase = PH_RUN;
                    it = 0;
                    spec = !WARM && P.max_iter == 1;
                    counted = false;
                    res[0] = res[1] = res[2] = res[3] = 0.f;
                    blk = a.sys + inst * SB::STRIDE;
                    {
                        const float2 *xp = reinterpret_cast<const float2 *>(a.x0 + inst * NX + 6 * h);
#pragma unroll
                        for (int j = 0; j < OX; ++j) x0o[j] = __ldg(xp + j);
                    }
                    if (WARM && a.wd) {
#pragma unroll 1
                        for (int i = 0; i < NH - 1; ++i) {
                            float2 t[3];
                            t[0] = __ldg(reinterpret_cast<const float2 *>(a.wd + inst * UROW + i * NU + 2 * h));
                            t[1] = __ldg(reinterpret_cast<const float2 *>(a.wy + inst * UROW + i * NU + 2 * h));
                            t[2] = __ldg(reinterpret_cast<const float2 *>(a.wz + inst * UROW + i * NU + 2 * h));
                            st.store_dyz(i, t);
                        }
#pragma unroll 1
                        for (int i = 0; i < NH; ++i) {
                            float2 t[6];
                            const float2 *gp = reinterpret_cast<const float2 *>(a.wg + inst * XROW + i * NX + 6 * h);
                            const float2 *vp = reinterpret_cast<const float2 *>(a.wv + inst * XROW + i * NX + 6 * h);
#pragma unroll
                            for (int j = 0; j < OX; ++j) { t[j] = __ldg(gp + j); t[3 + j] = __ldg(vp + j); }
                            st.store_gv(i, t);
                        }
                    } else {
                        float2 zu[3], zx[6];
#pragma unroll
                        for (int j = 0; j < 3; ++j) zu[j] = mk2(0.f, 0.f);
#pragma unroll
                        for (int j = 0; j < 6; ++j) zx[j] = mk2(0.f, 0.f);
#pragma unroll 1
                        for (int i = 0; i < NH - 1; ++i) st.store_dyz(i, zu);
#pragma unroll 1
                        for (int i = 0; i < NH; ++i) st.store_gv(i, zx);
                    }
                } else {
                    exhausted = true;
                }
            }
            // the new systems' blocks, NSLOT instances per pass: TMA bulk copy into the warp's staging slots, then both lanes of
            // each pair rewrite their tensor-memory columns from there
            unsigned fm = __ballot_sync(FULLM, fill) & EVEN;
            while (fm) {
                const int rank = __popc(fm & ((1u << pe) - 1u));
                const bool mine = ((fm >> pe) & 1u) && rank < SS::NSLOT;
                const int nmine = min(__popc(fm), SS::NSLOT);
                const float *sl = reinterpret_cast<const float *>(stage + (mine ? rank : 0) * SS::SLOTB);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0)
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((uint32_t)(nmine * SS::BLKB)) : "memory");
                __syncwarp();
                if (mine && h == 0) {
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sl);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"(dst), "l"(blk + SL::R0), "r"((uint32_t)(SL::L0 * 4)), "r"(bar) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"(dst + SL::L0 * 4), "l"(blk + SL::R1), "r"((uint32_t)(SL::L1 * 4)), "r"(bar) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"(dst + (SL::L0 + SL::L1) * 4), "l"(blk + SL::R2), "r"((uint32_t)(SL::L2 * 4)), "r"(bar) : "memory");
                }
                {
                    uint32_t done = 0;
                    while (!done)
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                     : "=r"(done) : "r"(bar), "r"(bar_phase) : "memory");
                    bar_phase ^= 1u;
                }
                if (mine) {
                    rho_l = sl[SL::at(SB::RHO)];
                    nrho_l = -rho_l;
#pragma unroll
                    for (int j = 0; j < OX; ++j) Qd2[j] = *reinterpret_cast<const float2 *>(sl + SL::at(SB::Qd) + 6 * h + 2 * j);
                    {   // own rows of the p_N seed: -(Xref_{N-1}^T * Pinf)   (admm.cpp:83)
                        float xr[NX];
                        gload<float, NX>(a.Xref + inst * a.xref_stride + (NH - 1) * NX, xr);
                        float2 pn[OX];
#pragma unroll
                        for (int j = 0; j < 2 * OX; ++j) {
                            float c[NX];
                            const float4 *cp = reinterpret_cast<const float4 *>(sl + SL::at(SB::Pf) + (6 * h + j) * NX);
#pragma unroll
                            for (int w = 0; w < NX / 4; ++w) {
                                const float4 t = cp[w];
                                c[4 * w] = t.x; c[4 * w + 1] = t.y; c[4 * w + 2] = t.z; c[4 * w + 3] = t.w;
                            }
                            const float vv = -dot<float, O::XtP, NX, FAST>([&](int k) { return c[k]; }, [&](int k) { return xr[k]; });
                            if (j & 1) pn[j >> 1].y = vv; else pn[j >> 1].x = vv;
                        }
                        st.store_pn(pn);
                    }
                }
                // tcgen05 is warp-collective: every lane rewrites its 256 columns, lanes that are not being filled with what they hold
#pragma unroll
                for (int b = 0; b < 4; ++b) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) tm_ld16(tcol + 16 * (4 * b + q), cb[q]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) wait16(cb[q]);
                    if (mine) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const int unit = 4 * b + q;
                                const Unit t = unit < PM::UNITS ? PM::fwd(8 * unit + e) : PM::bwd(8 * (unit - PM::UNITS) + e);
                                float2 cv = mk2(0.f, 0.f);
                                if (t.kind >= 0) cv = *reinterpret_cast<const float2 *>(sl + SL::at(PM::c0(t)) + h * PM::c1(t));
                                cb[q][2 * e] = cv.x; cb[q][2 * e + 1] = cv.y;
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) tm_st16(tcol + 16 * (4 * b + q), cb[q]);
                }
                tm_wait_st();
#pragma unroll
                for (int k = 0; k < SS::NSLOT; ++k) fm &= fm - 1u;
            }
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

        tm_ld16(tcol + 0, cb[0]);
        tm_ld16(tcol + 16, cb[1]);
        tm_ld16(tcol + 32, cb[2]);

        // ------------------------------------------------------------------ forward sweep
        // forward_pass (admm.cpp:27-37) + update_slack (:45-61) + update_dual (:67-71) + residual maxima (:95-98)
        float pri_x = 0.f, dua_x = 0.f, pri_u = 0.f, dua_u = 0.f;
        {
            float2 xo2[OX];      // own rows of x_i
            float xs[NX];        // all of x_i
#pragma unroll
            for (int j = 0; j < OX; ++j) xo2[j] = x0o[j];
            gather12(xo2, xs);
            const bool wr = emit || (spec && phase == PH_RUN);
            float *xo = (wr && a.x) ? a.x + inst * XROW + 6 * h : nullptr;
            float *uo = (wr && a.u) ? a.u + inst * UROW + 2 * h : nullptr;
            float *u0o = (wr && a.u0) ? a.u0 + inst * NU + 2 * h : nullptr;
            float *go = (WARM && emit && a.wg) ? a.wg + inst * XROW + 6 * h : nullptr;
            float *yo = (WARM && emit && a.wy) ? a.wy + inst * UROW + 2 * h : nullptr;
            float2 gv[6], dyz[3];   // own rows: g (0..2), v (3..5); d, y, z
            const int bxo = 6 * h, buo = 2 * h;
            auto state_pair = [&](int i, int j) {   // rows (6h + 2j, +1)   (:48, :59, :70, :95, :96)
                const int bi = i * NX + bxo + 2 * j;
                const float2 xg = ad2(xo2[j], gv[j]);
                float2 vn;
                if constexpr (CB) {
                    vn.x = fminf(cxhi[j].x, fmaxf(cxlo[j].x, xg.x));
                    vn.y = fminf(cxhi[j].y, fmaxf(cxlo[j].y, xg.y));
                } else {
                    vn.x = fminf(P.xmax[bi], fmaxf(P.xmin[bi], xg.x));
                    vn.y = fminf(P.xmax[bi + 1], fmaxf(P.xmin[bi + 1], xg.y));
                }
                const float2 rp = sb2(xo2[j], vn), rd = sb2(gv[3 + j], vn);
                pri_x = fmaxf(pri_x, fmaxf(fabsf(rp.x), fabsf(rp.y)));
                dua_x = fmaxf(dua_x, fmaxf(fabsf(rd.x), fabsf(rd.y)));
                gv[j] = sb2(xg, vn);
                gv[3 + j] = vn;
            };
            auto state_out = [&](int i) {
                st.store_gv(i, gv);
                if (xo) {
#pragma unroll
                    for (int j = 0; j < OX; ++j) reinterpret_cast<float2 *>(xo + i * NX)[j] = xo2[j];
                }
            };
            st.load_gv(0, gv);
            st.load_dyz(0, dyz);

#pragma unroll 1
            for (int i = 0; i < NH - 1; ++i) {
                float2 kx2, u2o, ax2[OX], bu2[OX];
                float us[NU];
                const uint32_t nxt = tcol + (i < NH - 2 ? 0u : 128u);
                if (WARM && go) {
#pragma unroll
                    for (int j = 0; j < OX; ++j) reinterpret_cast<float2 *>(go + i * NX)[j] = gv[j];
                }
                if (WARM && yo) *reinterpret_cast<float2 *>(yo + i * NU) = dyz[1];
#pragma unroll
                for (int u = 0; u < PM::UNITS; ++u) {
                    wait16(cb[u & 3]);
                    if (u + 3 < PM::UNITS) tm_ld16(tcol + 16 * (u + 3), cb[(u + 3) & 3]);
                    else tm_ld16(nxt + 16 * (u + 3 - PM::UNITS), cb[(u + 3) & 3]);
                    if (u >= 1 && u <= OX) state_pair(i, u - 1);
                    if (u == OX + 1) state_out(i);
                    if (u == 7) {   // the next stage's state, into the registers this stage has finished with
                        st.load_gv(i + 1, gv);
                        if (i + 1 < NH - 1) st.load_dyz(i + 1, dyz);
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float2 c = mk2(cb[u & 3][2 * e], cb[u & 3][2 * e + 1]);
                        const Unit t = PM::fwd(8 * u + e);
                        if (t.kind == U_K) {                               // Kinf x, own rows, column k              (:31)
                            const int k = t.idx;
                            if (k == 0) kx2 = FAST ? ml2(c, xs[0]) : pr2(c, xs[0], Z);
                            else kx2 = FAST ? fm2(c, xs[k], kx2) : ad2(pr2(c, xs[k], Z), kx2);
                        } else if (t.kind == U_A) {                        // Adyn x                                  (:35)
                            const int k = t.idx / 3, j = t.idx % 3;
                            if (k == 0) ax2[j] = FAST ? ml2(c, xs[0]) : pr2(c, xs[0], Z);
                            else ax2[j] = FAST ? fm2(c, xs[k], ax2[j]) : ad2(pr2(c, xs[k], Z), ax2[j]);
                        } else if (t.kind == U_B) {                        // Bdyn u                                  (:35)
                            const int k = t.idx / 3, j = t.idx % 3;
                            if constexpr (FAST) ax2[j] = fm2(c, us[k], ax2[j]);
                            else if (k == 0) bu2[j] = pr2(c, us[0], Z);
                            else bu2[j] = ad2(pr2(c, us[k], Z), bu2[j]);
                        }
                    }
                    if (u == PM::FWD_K_DONE) {   // Kinf x complete: own rows of the input, slack, dual, residuals
                        const int bi = i * NU + buo;
                        u2o = sb2(ng2(kx2), dyz[0]);                                                             // :31
                        const float2 uy = ad2(u2o, dyz[1]);                                                      // :47
                        float2 zn;
                        if constexpr (CB) {
                            zn.x = fminf(cuhi.x, fmaxf(culo.x, uy.x));                                           // :53
                            zn.y = fminf(cuhi.y, fmaxf(culo.y, uy.y));
                        } else {
                            zn.x = fminf(P.umax[bi], fmaxf(P.umin[bi], uy.x));
                            zn.y = fminf(P.umax[bi + 1], fmaxf(P.umin[bi + 1], uy.y));
                        }
                        const float2 rp = sb2(u2o, zn), rd = sb2(dyz[2], zn);
                        pri_u = fmaxf(pri_u, fmaxf(fabsf(rp.x), fabsf(rp.y)));                                   // :97
                        dua_u = fmaxf(dua_u, fmaxf(fabsf(rd.x), fabsf(rd.y)));                                   // :98
                        st.store_yz(i, sb2(uy, zn), zn);                                                         // :69
                        if (uo) *reinterpret_cast<float2 *>(uo + i * NU) = u2o;
                        if (u0o && i == 0) *reinterpret_cast<float2 *>(u0o) = u2o;
                        gather4(u2o, us);
                    }
                }
#pragma unroll
                for (int j = 0; j < OX; ++j) xo2[j] = FAST ? ax2[j] : ad2(ax2[j], bu2[j]);                       // :35
                gather12(xo2, xs);
            }
            if (WARM && go) {
#pragma unroll
                for (int j = 0; j < OX; ++j) reinterpret_cast<float2 *>(go + (NH - 1) * NX)[j] = gv[j];
            }
#pragma unroll
            for (int j = 0; j < OX; ++j) state_pair(NH - 1, j);
            state_out(NH - 1);
        }
        // the pair's residual maxima
        pri_x = fmaxf(pri_x, __shfl_xor_sync(FULLM, pri_x, 1));
        dua_x = fmaxf(dua_x, __shfl_xor_sync(FULLM, dua_x, 1));
        pri_u = fmaxf(pri_u, __shfl_xor_sync(FULLM, pri_u, 1));
        dua_u = fmaxf(dua_u, __shfl_xor_sync(FULLM, dua_u, 1));

        // ------------------------------------------------------------------ termination (admm.cpp:91-109, :135-138)
        bool final_bwd = false;
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
                if (h == 0) {
                    if (a.iter) a.iter[inst] = it;
                    if (a.status) a.status[inst] = conv ? 1 : 11;
                    if (a.resid) *reinterpret_cast<float4 *>(a.resid + inst * 4) = make_float4(res[0], res[1], res[2], res[3]);
                    n_iter += (unsigned)it;
                    n_solved += conv ? 1u : 0u;
                    ++n_inst;
                }
                final_bwd = !conv;
                phase = spec ? PH_FREE : PH_EMIT;   // (spec: this trip's forward sweep has already written the trajectory)
            } else if constexpr (!WARM) {
                constexpr float SF = TMPC_SPEC_FACTOR_SYS;
                const bool next_chk = ((it + 1) % P.check_term) == 0;
                spec = (it + 1 >= P.max_iter) ||
                       (next_chk && res[0] < SF * P.pri_tol && res[2] < SF * P.pri_tol && res[1] < SF * P.dua_tol && res[3] < SF * P.dua_tol);
            }
        } else if (phase == PH_EMIT) {
            phase = PH_FREE;
        }
        if (a.done) {   // every output of the instance is written (by both lanes of its pair): count it once
            const bool fin = phase == PH_FREE && inst >= 0 && !counted;
            const unsigned cm = __ballot_sync(FULLM, fin);
            if (fin) {
                counted = true;
                __threadfence();
                __syncwarp(cm);
                if (h == 0) atomicAdd(a.done + (inst >> a.done_shift), 1u);
            }
        }

        // ------------------------------------------------------------------ backward sweep
        // update_linear_cost (admm.cpp:77-85) recomputed per stage + backward_pass_grad (:15-22)
        const bool cont = (phase == PH_RUN);
        const bool wout = WARM && (cont || final_bwd) && a.wd;
        if (__any_sync(FULLM, cont || wout)) {
            float2 po2[OX];   // own rows of p
            float ps[NX];     // all of p
            const float *xr_base = a.Xref + (inst < 0 ? 0 : inst) * a.xref_stride + 6 * h;
            float *wdo = wout ? a.wd + inst * UROW + 2 * h : nullptr;
            float *wvo = wout ? a.wv + inst * XROW + 6 * h : nullptr;
            float *wzo = wout ? a.wz + inst * UROW + 2 * h : nullptr;
            const float2 rho2 = mk2(rho_l, rho_l), nrho2 = mk2(nrho_l, nrho_l);
            float2 gv[6], dyz[3], xr2[OX];
            {
                float2 pn[OX];
                st.load_gv(NH - 1, gv);
                st.load_pn(pn);
                if (WARM && wvo) {
#pragma unroll
                    for (int j = 0; j < OX; ++j) reinterpret_cast<float2 *>(wvo + (NH - 1) * NX)[j] = gv[3 + j];
                }
#pragma unroll
                for (int j = 0; j < OX; ++j) {
                    const float2 dv = sb2(gv[3 + j], gv[j]);
                    if constexpr (FAST) po2[j] = __ffma2_rn(nrho2, dv, pn[j]);
                    else po2[j] = sb2(pn[j], pp2(rho2, dv, Z));                                                  // :84
                }
                gather12(po2, ps);
            }
            st.load_gv(NH - 2, gv);
            st.load_dyz(NH - 2, dyz);
#pragma unroll
            for (int j = 0; j < OX; ++j) xr2[j] = __ldg(reinterpret_cast<const float2 *>(xr_base + (NH - 2) * NX) + j);
#pragma unroll 1
            for (int i = NH - 2; i >= 0; --i) {
                float2 r2o, q2[OX], s2o, d2o, mp2[OX], kr2[OX];
                float2 e0[4], e1[4];                     // B^T p: vectorised redux, lane L = k % 4
                float2 t0[OX], t1[OX], ta[OX], tl[OX];   // AmBKt p: half-split tree over 12 = ((3 + 3) + (3 + 3)), 3 = e + (e + e)
                float2 k0[OX], k1[OX];                   // Kinf^T r: (e0 + e2) + (e1 + e3)
                float rs[NU], ss[NU];
#pragma unroll
                for (int u = 0; u < PM::UNITS; ++u) {
                    wait16(cb[u & 3]);
                    if (u + 3 < PM::UNITS) tm_ld16(tcol + 128 + 16 * (u + 3), cb[(u + 3) & 3]);
                    else if (i > 0) tm_ld16(tcol + 128 + 16 * (u + 3 - PM::UNITS), cb[(u + 3) & 3]);
                    if (u == 1) {   // own rows of r_i, q_i from (z, y, v, g, Xref)
                        if (WARM && wvo) {
#pragma unroll
                            for (int j = 0; j < OX; ++j) reinterpret_cast<float2 *>(wvo + i * NX)[j] = gv[3 + j];
                            *reinterpret_cast<float2 *>(wzo + i * NU) = dyz[2];
                        }
                        r2o = pp2(nrho2, sb2(dyz[2], dyz[1]), Z);                                                // :80
                        gather4(r2o, rs);
#pragma unroll
                        for (int j = 0; j < OX; ++j) {
                            const float2 dv = sb2(gv[3 + j], gv[j]);
                            const float2 cq = ng2(pp2(xr2[j], Qd2[j], Z));                                       // :81
                            if constexpr (FAST) q2[j] = __ffma2_rn(nrho2, dv, cq);
                            else q2[j] = sb2(cq, pp2(rho2, dv, Z));                                              // :82
                        }
                    }
                    if (u == 7 && i > 0) {   // the next stage's state (this one's was consumed at u == 1)
                        st.load_gv(i - 1, gv);
                        st.load_dyz(i - 1, dyz);
#pragma unroll
                        for (int j = 0; j < OX; ++j) xr2[j] = __ldg(reinterpret_cast<const float2 *>(xr_base + (i - 1) * NX) + j);
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float2 c = mk2(cb[u & 3][2 * e], cb[u & 3][2 * e + 1]);
                        const Unit t = PM::bwd(8 * u + e);
                        if (t.kind == U_BR) {                              // Bdyn^T p, own rows, term k               (:19)
                            const int k = t.idx;
                            if constexpr (FAST) {
                                s2o = k == 0 ? ml2(c, ps[0]) : fm2(c, ps[k], s2o);
                            } else {
                                const int Lk = k % 4, qk = k / 4;
                                const float2 pr = pr2(c, ps[k], Z);
                                if (qk == 0) e0[Lk] = pr;
                                else if (qk == 1) e1[Lk] = pr;
                                else e0[Lk] = ad2(e0[Lk], ad2(e1[Lk], pr));
                            }
                        } else if (t.kind == U_QI) {                       // Quu_inv s, own rows, column k            (:19)
                            const int k = t.idx;
                            if (k == 0) d2o = FAST ? ml2(c, ss[0]) : pr2(c, ss[0], Z);
                            else d2o = FAST ? fm2(c, ss[k], d2o) : ad2(pr2(c, ss[k], Z), d2o);
                        } else if (t.kind == U_M) {                        // AmBKt p, own rows, term k                (:20)
                            const int k = t.idx / 3, j = t.idx % 3;
                            if constexpr (FAST) {
                                mp2[j] = k == 0 ? ml2(c, ps[0]) : fm2(c, ps[k], mp2[j]);
                            } else {
                                const int mk = k % 3, qk = k / 3;
                                const float2 pr = pr2(c, ps[k], Z);
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
                        } else {                                           // Kinf^T r, own rows, term k               (:20)
                            const int k = t.idx / 3, j = t.idx % 3;
                            if constexpr (FAST) {
                                kr2[j] = k == 0 ? ml2(c, rs[0]) : fm2(c, rs[k], kr2[j]);
                            } else {
                                const float2 pr = pr2(c, rs[k], Z);
                                if (k == 0) k0[j] = pr;
                                else if (k == 1) k1[j] = pr;
                                else if (k == 2) k0[j] = ad2(k0[j], pr);
                                else kr2[j] = ad2(k0[j], ad2(k1[j], pr));
                            }
                        }
                    }
                    if (u == PM::BWD_BR_DONE) {   // B^T p complete: own rows of s, then all of s
                        if constexpr (!FAST) s2o = ad2(ad2(e0[0], e0[2]), ad2(e0[1], e0[3]));
                        s2o = ad2(s2o, r2o);
                        gather4(s2o, ss);
                    }
                    if (u == PM::BWD_QI_DONE) {   // own rows of d_i = Quu_inv (B^T p + r)
                        st.store_d(i, d2o, cont);
                        if (WARM && wdo) *reinterpret_cast<float2 *>(wdo + i * NU) = d2o;
                    }
                }
#pragma unroll
                for (int j = 0; j < OX; ++j) po2[j] = sb2(ad2(q2[j], mp2[j]), kr2[j]);                           // :20
                gather12(po2, ps);
            }
        } else {
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
            atomicAdd(a.stats + 2, n_trips >> 1);   // lane-trips counted per instance (two lanes each)
            atomicAdd(a.stats + 3, n_inst);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tcol) : "memory");   // (warp 0: tcol is the allocation's base)
}

}  // namespace tmpc
