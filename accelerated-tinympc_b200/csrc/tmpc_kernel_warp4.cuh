// Warp-per-FOUR-instances ADMM kernel for the large shape (nx = 32, nu = 8; BASELINE config 5: N = 50) on sm_100a.
//
// Successor of tmpc_kernel_warp.cuh (one instance per warp; kept as TMPC_KERNEL=warp1), same arithmetic, same orders, same
// results bit for bit.  ncu on that kernel (profiles/r01_ncu_summary.md): FMA pipe 58 % busy at 23 % of the FP32 roofline,
// because a fifth of the pipe's work was REDUNDANT -- with one instance per warp the eight rows of Kinf x and Quu_inv s are
// strict 32- and 8-long sequential chains, so the four lanes of a row group all computed the same chain -- and 17.4 k
// warp-instructions per iteration went through one-instance-wide broadcasts, syncs and address arithmetic.
//
// Here one warp carries FOUR instances ("slots") at once:
//   * lane j owns row j of every nx-vector (x, p, g, v, q) of all four slots; the row's coefficients (Adyn, AmBKt, Bdyn,
//     Kinf^T) sit in registers once and serve four products each;
//   * lane (ur = lane >> 2, sl = lane & 3) owns nu-row ur of slot sl: the Kinf x / Quu_inv s chains and all nu-sized
//     element-wise work (u, d, y, z, r, slack, dual) use all 32 lanes on DISTINCT (slot, row) pairs -- nothing is computed twice;
//   * the accumulations of two slots advance together as FADD2 (pairs across slots: each slot's own order is untouched);
//   * one broadcast buffer / one __syncwarp serves four instances; slots run different iteration counts and are refilled
//     individually from the global work counter (a slot that ends skips one backward sweep, nothing else is wasted).
// g and v (73 % of the state) stay in TENSOR MEMORY: 4 slots x 2 x NH columns of the warp's own lane quadrant (4 warps per
// CTA, one per quadrant: 400 of 512 columns); d, y, z, p_N, the broadcast buffers, the coefficient images and the bound rows
// are in shared memory (16 instances per SM, as before).  Latency is hidden by the four independent slots of the one warp
// each scheduler runs, not by warp switching.
#pragma once
#include "tmpc_kernel_warp.cuh"

namespace tmpc {

template <int NH> struct Warp4Smem {
    static constexpr int NS = 4;                                                                     // slots per warp
    static constexpr int D = 0, Y = D + (NH - 1) * WNU, Z = Y + (NH - 1) * WNU, PN = Z + (NH - 1) * WNU, SLOT = PN + WNX;   // floats per slot
    // per warp: XB[slot][32] broadcast of x / p; XT[slot][sl][8] = p(4 j + sl): the row-major GEMV's packets of SIMD lane sl;
    // UB[slot][8] broadcast of u / r; SB[slot][8] broadcast of s
    static constexpr int XB = NS * SLOT, XT = XB + NS * WNX, UB = XT + NS * WNX, SB = UB + NS * WNU, FLOATS = SB + NS * WNU;
    static_assert(SLOT % 4 == 0 && FLOATS % 4 == 0, "16-byte aligned regions");
    using SH = WarpSmem<NH, true>;                                                                   // CTA-shared staging: coefficient images, bounds, TMEM slot
    static constexpr size_t total_bytes(int warps) { return size_t(SH::SH_FLOATS) * 4 + size_t(FLOATS) * 4 * warps; }
};

// WARPS = 8 (TMPC_KERNEL=warp4x2): TWO warps per scheduler.  Tensor memory decides the residency of the four-warp kernel (400 of a lane
// quadrant's 512 columns per warp); shared memory has room for eight warps (197 KB).  So each warp keeps g, v of slots 0 and 1 in
// tensor memory (200 columns; two warps per quadrant) and g, v of slots 2 and 3 in its rows of the L2-resident scratch
// (SolveArgs::scratch: [warp][2 slots][g | v][NH] rows of 128 B, one float per lane, read and written by the owning lane only, one
// stage ahead of their use like the tensor-memory columns).  32 instances per SM.
template <int NH, bool FAST, bool WARM, int WARPS = 4>
__global__ void __launch_bounds__(WARPS * 32, 1)
admm_kernel_warp4(const __grid_constant__ ModelWarp P, const __grid_constant__ SolveArgs<float> a)
{
    using S = Warp4Smem<NH>;
    using SH = typename S::SH;
    constexpr int NS = S::NS;
    static_assert(WARPS == 4 || WARPS == 8, "one or two warps per scheduler");
    constexpr int TS = WARPS == 4 ? NS : 2;   // slots whose g, v live in tensor memory
    static_assert((WARPS / 4) * TS * 2 * NH <= 512, "slots x (g, v) x NH columns of the warps that share a TMEM lane quadrant");
    constexpr unsigned FULLM = 0xffffffffu;
    constexpr int XROW = WNX * NH, UROW = WNU * (NH - 1);
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ur = lane >> 2;            // the nu-row this lane owns ...
    const int sl = lane & 3;             // ... of slot sl; also the SIMD lane of the reference's packet in the row-major GEMV
    float *shr = reinterpret_cast<float *>(smem);
    float *ws = shr + SH::SH_FLOATS + size_t(warp) * S::FLOATS;
    float *xb = ws + S::XB, *xt = ws + S::XT, *ub = ws + S::UB, *sb = ws + S::SB;
    float *my = ws + sl * S::SLOT;       // the slot whose nu-rows this lane owns
    float *myd = my + S::D, *myy = my + S::Y, *myz = my + S::Z;
    const float Qd = __ldg(P.Qd + lane);
    const float2 Z = P.nz2;
    (void)Z;
    // rho / -rho in registers for good: as constant-bank operands ptxas re-loads them (LDCU) at the top of every backward stage
    // and the one warp of the scheduler waits out the load
    float rho = P.rho, nrho = P.nrho;
    asm volatile("" : "+f"(rho), "+f"(nrho));
    const bool warm = WARM && a.wd;
    unsigned long long n_iter = 0, n_solved = 0, n_inst = 0, n_trips = 0;

    // ---- CTA set-up: tensor memory, coefficient images and bound rows into shared memory
    uint32_t *slot_addr = reinterpret_cast<uint32_t *>(shr + SH::SH_SLOT);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(slot_addr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int k = threadIdx.x; k < WARP_FWD4 * 32; k += WARPS * 32) reinterpret_cast<float4 *>(shr + SH::SH_FWD)[k] = __ldg(P.fwd4 + k);
    for (int k = threadIdx.x; k < WARP_BWD4 * 32; k += WARPS * 32) reinterpret_cast<float4 *>(shr + SH::SH_BWD)[k] = __ldg(P.bwd4 + k);
    for (int k = threadIdx.x; k < NH * WNX; k += WARPS * 32) { shr[SH::SH_XMIN + k] = __ldg(P.xmin + k); shr[SH::SH_XMAX + k] = __ldg(P.xmax + k); }
    for (int k = threadIdx.x; k < (NH - 1) * WNU; k += WARPS * 32) { shr[SH::SH_UMIN + k] = __ldg(P.umin + k); shr[SH::SH_UMAX + k] = __ldg(P.umax + k); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // this warp's lane quadrant (warp % 4) and its column range in it
    const uint32_t tbase = *slot_addr + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * TS * 2 * NH);
    float *scr = nullptr;                                                  // this lane's column of the warp's scratch rows
    if constexpr (TS < NS) scr = reinterpret_cast<float *>(a.scratch) + ((size_t)blockIdx.x * WARPS + warp) * ((NS - TS) * 2 * NH) * 32 + lane;
    const float4 *fwd4 = reinterpret_cast<const float4 *>(shr + SH::SH_FWD), *bwd4 = reinterpret_cast<const float4 *>(shr + SH::SH_BWD);
    const float *bxmin = shr + SH::SH_XMIN, *bxmax = shr + SH::SH_XMAX, *bumin = shr + SH::SH_UMIN, *bumax = shr + SH::SH_UMAX;
    // slot e, stage i: g at column e*2*NH + i, v at + NH
    auto gcol = [&](int e, int i) -> uint32_t { return tbase + (uint32_t)(e * 2 * NH + i); };
    auto vcol = [&](int e, int i) -> uint32_t { return tbase + (uint32_t)(e * 2 * NH + NH + i); };
    // g / v of (slot e, stage i): tensor-memory column or scratch row (e is a compile-time constant wherever these are called)
    auto ldg_ = [&](int e, int i, float &r) { if (e < TS) tm_ld1(gcol(e, i), r); else r = __ldcg(scr + ((e - TS) * 2 * NH + i) * 32); };
    auto ldv_ = [&](int e, int i, float &r) { if (e < TS) tm_ld1(vcol(e, i), r); else r = __ldcg(scr + ((e - TS) * 2 * NH + NH + i) * 32); };
    auto stg_ = [&](int e, int i, float r) { if (e < TS) tm_st1(gcol(e, i), r); else __stcg(scr + ((e - TS) * 2 * NH + i) * 32, r); };
    auto stv_ = [&](int e, int i, float r) { if (e < TS) tm_st1(vcol(e, i), r); else __stcg(scr + ((e - TS) * 2 * NH + NH + i) * 32, r); };

    // ---- per-slot state (warp-uniform unless noted)
    long long inst[NS];
    int it[NS];
    bool run[NS];
    float x0[NS];                        // per lane: row `lane` of the slot's initial state
    float res[NS][4];
#pragma unroll
    for (int e = 0; e < NS; ++e) { inst[e] = 0; it[e] = 0; run[e] = false; x0[e] = 0.f; res[e][0] = res[e][1] = res[e][2] = res[e][3] = 0.f; }
    bool exhausted = false;

    auto finish = [&](int e, bool conv) {
        // outputs of a slot that ended: y, g as the reference leaves them (this iteration's; d, v, z were mirrored by the backward
        // sweeps), iteration count, status, residuals
        float *sy = ws + e * S::SLOT + S::Y;
        if (warm) {
            float *gy = a.wy + inst[e] * UROW, *gg = a.wg + inst[e] * XROW;
            __syncwarp();
            for (int k = lane; k < UROW; k += 32) gy[k] = sy[k];
#pragma unroll 5
            for (int i = 0; i < NH; ++i) {
                float g;
                ldg_(e, i, g);
                asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(g) :: "memory");
                gg[i * WNX + lane] = g;
            }
        }
        if (lane == 0) {
            if (a.iter) a.iter[inst[e]] = it[e];
            if (a.status) a.status[inst[e]] = conv ? 1 : 11;
            if (a.resid) *reinterpret_cast<float4 *>(a.resid + inst[e] * 4) = make_float4(res[e][0], res[e][1], res[e][2], res[e][3]);
        }
        n_iter += (unsigned)it[e]; n_solved += conv ? 1u : 0u; ++n_inst;
        if (a.done) {
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(a.done + (inst[e] >> a.done_shift), 1u);
        }
        run[e] = false;
    };

    for (;;) {
        // ------------------------------------------------------------------ refill free slots
#pragma unroll
        for (int e = 0; e < NS; ++e) {
            if (run[e] || exhausted) continue;
            long long idx;
            {
                unsigned long long b = 0;
                if (lane == 0) b = atomicAdd(a.counter, 1ull);
                idx = (long long)__shfl_sync(FULLM, b, 0);
            }
            if (idx >= a.batch) { exhausted = true; continue; }
            const long long ni = claim_instance(a, idx);
            if (ni < 0) { exhausted = true; continue; }
            inst[e] = ni; it[e] = 0; run[e] = true;
            res[e][0] = res[e][1] = res[e][2] = res[e][3] = 0.f;
            float *se = ws + e * S::SLOT;
            const float *xref = a.Xref + ni * a.xref_stride;
            x0[e] = __ldg(a.x0 + ni * WNX + lane);
            {   // p_N seed: -(Xref_{N-1}^T Pinf)   (admm.cpp:83)
                const float *xl = xref + (NH - 1) * WNX;
                se[S::PN + lane] = -dot<float, ORD_VECREDUX, WNX, FAST>([&](int k) { return __ldg(P.Pt + k * WNX + lane); }, [&](int k) { return __ldg(xl + k); });
            }
            if (warm) {
                const float *gd = a.wd + ni * UROW, *gy = a.wy + ni * UROW, *gz = a.wz + ni * UROW;
                const float *gg = a.wg + ni * XROW, *gv = a.wv + ni * XROW;
                for (int k = lane; k < UROW; k += 32) { se[S::D + k] = gd[k]; se[S::Y + k] = gy[k]; se[S::Z + k] = gz[k]; }
#pragma unroll 5
                for (int i = 0; i < NH; ++i) { stg_(e, i, gg[i * WNX + lane]); stv_(e, i, gv[i * WNX + lane]); }
            } else {
                for (int k = lane; k < UROW; k += 32) { se[S::D + k] = 0.f; se[S::Y + k] = 0.f; se[S::Z + k] = 0.f; }
#pragma unroll 5
                for (int i = 0; i < NH; ++i) { stg_(e, i, 0.f); stv_(e, i, 0.f); }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        if (!(run[0] || run[1] || run[2] || run[3])) break;
        __syncwarp();
        ++n_trips;
#pragma unroll
        for (int e = 0; e < NS; ++e) if (run[e]) ++it[e];
        const bool xw[NS] = {run[0] && a.x, run[1] && a.x, run[2] && a.x, run[3] && a.x};
        const bool urun = (sl == 0 ? run[0] : sl == 1 ? run[1] : sl == 2 ? run[2] : run[3]);        // per lane: my nu-slot is live
        const long long uinst = (sl == 0 ? inst[0] : sl == 1 ? inst[1] : sl == 2 ? inst[2] : inst[3]);

        // ------------------------------------------------------------------ forward sweep
        // forward_pass (admm.cpp:27-37) + update_slack (:45-61) + update_dual (:67-71) + residual maxima (:95-98)
        float pri_x[NS], dua_x[NS], pri_u = 0.f, dua_u = 0.f;
#pragma unroll
        for (int e = 0; e < NS; ++e) pri_x[e] = dua_x[e] = 0.f;
        {
            // Coefficient rows are NOT held in registers across the sweep (72 registers: with four slots of state on top ptxas was
            // re-materialising addresses with S2R inside the stage loop): each 16-byte chunk of Adyn(lane,:) / Kinf(ur,:) is read from
            // the CTA's shared-memory image next to the x chunks it multiplies, one chunk ahead of its use.
            float Bc[WNU];
#pragma unroll
            for (int g4 = 0; g4 < 2; ++g4) {
                const float4 t = fwd4[(16 + g4) * 32 + lane];
                Bc[4 * g4] = t.x; Bc[4 * g4 + 1] = t.y; Bc[4 * g4 + 2] = t.z; Bc[4 * g4 + 3] = t.w;
            }
            const float4 *xb4 = reinterpret_cast<const float4 *>(xb);
            const float4 *cA4 = fwd4 + lane, *cK4 = fwd4 + 8 * 32 + lane;
            struct Chunk { float4 cA, cK, v0, v1, v2, v3, vk; };
            auto ld_chunk = [&](int g4, Chunk &c) {
                c.cA = cA4[g4 * 32]; c.cK = cK4[g4 * 32];
                c.v0 = xb4[g4]; c.v1 = xb4[8 + g4]; c.v2 = xb4[16 + g4]; c.v3 = xb4[24 + g4]; c.vk = xb4[sl * 8 + g4];
            };
            float x[NS], g[NS], v[NS];
#pragma unroll
            for (int e = 0; e < NS; ++e) { x[e] = x0[e]; ldg_(e, 0, g[e]); ldv_(e, 0, v[e]); }
            float xmn = bxmin[lane], xmx = bxmax[lane], umn = bumin[ur], umx = bumax[ur];
            float d = myd[ur], y = myy[ur], z = myz[ur];
            float *uo = (urun && a.u) ? a.u + uinst * UROW + ur : nullptr;
            float *xo[NS];     // always a valid row address (a.x may be null: then never stored through)
#pragma unroll
            for (int e = 0; e < NS; ++e) xo[e] = a.x + inst[e] * XROW + lane;
#pragma unroll 1
            for (int i = 0; i < NH - 1; ++i) {
#pragma unroll
                for (int e = 0; e < NS; ++e) xb[e * WNX + lane] = x[e];
                __syncwarp();
                asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(g[0]), "+f"(g[1]), "+f"(g[2]), "+f"(g[3]), "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]) :: "memory");
#pragma unroll
                for (int e = 0; e < NS; ++e) if (xw[e]) xo[e][i * WNX] = x[e];
                // Adyn(lane,:) x_i of every slot (pairs of slots advance as one FADD2 chain) and Kinf(ur,:) x_i of slot sl
                float2 a01, a23;
                float ka;
                // four sequential-chain steps (one 16-byte chunk of the row) for the five chains this lane advances
                auto do_chunk = [&](const Chunk &c, auto first_tag) {
                    constexpr bool FIRST = decltype(first_tag)::value;
                    const float A4[4] = {c.cA.x, c.cA.y, c.cA.z, c.cA.w}, K4[4] = {c.cK.x, c.cK.y, c.cK.z, c.cK.w};
                    const float s0[4] = {c.v0.x, c.v0.y, c.v0.z, c.v0.w}, s1[4] = {c.v1.x, c.v1.y, c.v1.z, c.v1.w}, s2[4] = {c.v2.x, c.v2.y, c.v2.z, c.v2.w},
                                s3[4] = {c.v3.x, c.v3.y, c.v3.z, c.v3.w}, sk[4] = {c.vk.x, c.vk.y, c.vk.z, c.vk.w};
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        if constexpr (FAST) {
                            if (FIRST && t == 0) {
                                a01 = __fmul2_rn(f2(A4[0], A4[0]), f2(s0[0], s1[0])); a23 = __fmul2_rn(f2(A4[0], A4[0]), f2(s2[0], s3[0]));
                                ka = __fmul_rn(K4[0], sk[0]);
                            } else {
                                a01 = __ffma2_rn(f2(A4[t], A4[t]), f2(s0[t], s1[t]), a01); a23 = __ffma2_rn(f2(A4[t], A4[t]), f2(s2[t], s3[t]), a23);
                                ka = __fmaf_rn(K4[t], sk[t], ka);
                            }
                        } else {
                            const float2 e01 = f2(__fmul_rn(A4[t], s0[t]), __fmul_rn(A4[t], s1[t])), e23 = f2(__fmul_rn(A4[t], s2[t]), __fmul_rn(A4[t], s3[t]));
                            const float ek = __fmul_rn(K4[t], sk[t]);
                            if (FIRST && t == 0) { a01 = e01; a23 = e23; ka = ek; }
                            else { a01 = add2(e01, a01); a23 = add2(e23, a23); ka = __fadd_rn(ek, ka); }
                        }
                    }
                };
                {
                    // straight-line, three chunk buffers, every chunk loaded TWO chunks (~80 issue slots) ahead of its use; the empty
                    // asm statements pin the loads where they are written (ptxas otherwise sinks them next to their first use and the
                    // single warp of the scheduler eats the whole shared-memory latency: 20 % of the stall samples of the first version)
#define TMPC_PIN() asm volatile("" ::: "memory")
                    Chunk ca, cb, cc;
                    ld_chunk(0, ca); ld_chunk(1, cb); ld_chunk(2, cc); TMPC_PIN();
                    do_chunk(ca, std::true_type());  ld_chunk(3, ca); TMPC_PIN();
                    do_chunk(cb, std::false_type()); ld_chunk(4, cb); TMPC_PIN();
                    do_chunk(cc, std::false_type()); ld_chunk(5, cc); TMPC_PIN();
                    do_chunk(ca, std::false_type()); ld_chunk(6, ca); TMPC_PIN();
                    do_chunk(cb, std::false_type()); ld_chunk(7, cb); TMPC_PIN();
                    do_chunk(cc, std::false_type()); TMPC_PIN();
                    do_chunk(ca, std::false_type()); TMPC_PIN();
                    do_chunk(cb, std::false_type());
#undef TMPC_PIN
                }
                const float ax[NS] = {a01.x, a01.y, a23.x, a23.y};
                const float u = __fsub_rn(-ka, d);                                                     // :31  (slot sl, row ur)
                {   // input slack / dual / residuals of (slot sl, row ur)
                    float t = __fadd_rn(u, y);                                                         // :47
                    t = fminf(umx, fmaxf(umn, t));                                                     // :53
                    pri_u = fmaxf(pri_u, fabsf(__fsub_rn(u, t)));                                      // :97
                    dua_u = fmaxf(dua_u, fabsf(__fsub_rn(z, t)));                                      // :98
                    const float yn = __fsub_rn(__fadd_rn(y, u), t);                                    // :69
                    myy[i * WNU + ur] = yn;
                    myz[i * WNU + ur] = t;
                    ub[sl * WNU + ur] = u;
                    if (uo) uo[i * WNU] = u;
                    if (i == 0 && urun && a.u0) a.u0[uinst * WNU + ur] = u;
                }
                // state slack / dual / residuals of row `lane` of every slot, two slots per packed instruction
#pragma unroll
                for (int h = 0; h < NS; h += 2) {
                    const float2 x2 = f2(x[h], x[h + 1]), g2 = f2(g[h], g[h + 1]), v2 = f2(v[h], v[h + 1]);
                    float2 t = add2(x2, g2);                                                           // :48
                    t.x = fminf(xmx, fmaxf(xmn, t.x)); t.y = fminf(xmx, fmaxf(xmn, t.y));              // :59
                    const float2 rp = sub2(x2, t), rd = sub2(v2, t);
                    pri_x[h] = fmaxf(pri_x[h], fabsf(rp.x)); pri_x[h + 1] = fmaxf(pri_x[h + 1], fabsf(rp.y));   // :95
                    dua_x[h] = fmaxf(dua_x[h], fabsf(rd.x)); dua_x[h + 1] = fmaxf(dua_x[h + 1], fabsf(rd.y));   // :96
                    const float2 gn = sub2(add2(g2, x2), t);                                           // :70
                    stg_(h, i, gn.x); stv_(h, i, t.x);
                    stg_(h + 1, i, gn.y); stv_(h + 1, i, t.y);
                }
                __syncwarp();
                {   // operands of stage i+1 (the nu-rows of the last stage do not exist: re-read stage i's)
                    const int in = i + 1, iu = (in < NH - 1) ? in : i;
                    xmn = bxmin[in * WNX + lane]; xmx = bxmax[in * WNX + lane];
                    umn = bumin[iu * WNU + ur]; umx = bumax[iu * WNU + ur];
#pragma unroll
                    for (int e = 0; e < NS; ++e) { ldg_(e, in, g[e]); ldv_(e, in, v[e]); }
                    d = myd[iu * WNU + ur]; y = myy[iu * WNU + ur]; z = myz[iu * WNU + ur];
                }
                // x_{i+1} = Adyn x_i + Bdyn u_i                                                         :35
                float2 b01, b23;
#pragma unroll
                for (int g4 = 0; g4 < 2; ++g4) {
                    const float4 v0 = reinterpret_cast<const float4 *>(ub)[g4], v1 = reinterpret_cast<const float4 *>(ub + WNU)[g4],
                                 v2 = reinterpret_cast<const float4 *>(ub + 2 * WNU)[g4], v3 = reinterpret_cast<const float4 *>(ub + 3 * WNU)[g4];
                    const float s0[4] = {v0.x, v0.y, v0.z, v0.w}, s1[4] = {v1.x, v1.y, v1.z, v1.w}, s2[4] = {v2.x, v2.y, v2.z, v2.w},
                                s3[4] = {v3.x, v3.y, v3.z, v3.w};
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int k = 4 * g4 + t;
                        if constexpr (FAST) {
                            if (k == 0) { b01 = a01; b23 = a23; }
                            b01 = __ffma2_rn(f2(Bc[k], Bc[k]), f2(s0[t], s1[t]), b01); b23 = __ffma2_rn(f2(Bc[k], Bc[k]), f2(s2[t], s3[t]), b23);
                        } else {
                            const float2 e01 = f2(__fmul_rn(Bc[k], s0[t]), __fmul_rn(Bc[k], s1[t])), e23 = f2(__fmul_rn(Bc[k], s2[t]), __fmul_rn(Bc[k], s3[t]));
                            if (k == 0) { b01 = e01; b23 = e23; }
                            else { b01 = add2(e01, b01); b23 = add2(e23, b23); }
                        }
                    }
                }
                if constexpr (FAST) { x[0] = b01.x; x[1] = b01.y; x[2] = b23.x; x[3] = b23.y; }
                else {
                    const float2 n01 = add2(f2(ax[0], ax[1]), b01), n23 = add2(f2(ax[2], ax[3]), b23);
                    x[0] = n01.x; x[1] = n01.y; x[2] = n23.x; x[3] = n23.y;
                }
            }
            {   // last stage: state slack / dual only
                constexpr int i = NH - 1;
                asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(g[0]), "+f"(g[1]), "+f"(g[2]), "+f"(g[3]), "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]) :: "memory");
#pragma unroll
                for (int e = 0; e < NS; ++e) {
                    if (xw[e]) xo[e][i * WNX] = x[e];
                    float t = __fadd_rn(x[e], g[e]);
                    t = fminf(xmx, fmaxf(xmn, t));
                    pri_x[e] = fmaxf(pri_x[e], fabsf(__fsub_rn(x[e], t)));
                    dua_x[e] = fmaxf(dua_x[e], fabsf(__fsub_rn(v[e], t)));
                    stg_(e, i, __fsub_rn(__fadd_rn(g[e], x[e]), t)); stv_(e, i, t);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
        }

        // ------------------------------------------------------------------ termination (admm.cpp:91-109, :135-138), slot by slot
        bool cont[NS], fbw[NS];     // the slot goes on / ends at max_iter but still owes its warm state the last backward sweep
        {
            // input residuals live in the lanes of the slot's (sl) group: mask the others out of the warp maximum
            float mpu[NS], mdu[NS];
#pragma unroll
            for (int e = 0; e < NS; ++e) {
                mpu[e] = warp_max_nonneg(sl == e ? pri_u : 0.f);
                mdu[e] = warp_max_nonneg(sl == e ? dua_u : 0.f);
            }
#pragma unroll
            for (int e = 0; e < NS; ++e) {
                cont[e] = false; fbw[e] = false;
                if (!run[e]) continue;
                const bool chk = (it[e] % P.check_term) == 0;
                if (chk) {
                    res[e][0] = warp_max_nonneg(pri_x[e]);
                    res[e][1] = __fmul_rn(warp_max_nonneg(dua_x[e]), P.rho);
                    res[e][2] = mpu[e];
                    res[e][3] = __fmul_rn(mdu[e], P.rho);
                }
                const bool conv = chk && res[e][0] < P.pri_tol && res[e][2] < P.pri_tol && res[e][1] < P.dua_tol && res[e][3] < P.dua_tol;
                const bool last = it[e] >= P.max_iter;
                if (conv) finish(e, true);
                else if (last && !warm) finish(e, false);   // the final backward pass only matters for the warm state it leaves behind
                else if (last) fbw[e] = true;
                else cont[e] = true;
            }
        }
        const bool any_bwd = cont[0] || cont[1] || cont[2] || cont[3] || fbw[0] || fbw[1] || fbw[2] || fbw[3];
        if (!any_bwd) continue;

        // ------------------------------------------------------------------ backward sweep
        // update_linear_cost (admm.cpp:77-85) recomputed per stage + backward_pass_grad (:15-22)
        __syncwarp();   // the forward sweep's last reads of ub precede this sweep's writes
        {
            float BTc[WNU], Qic[WNU], KTc[WNU];     // (the AmBKt row is read chunk by chunk next to the p chunks it multiplies)
#pragma unroll
            for (int g4 = 0; g4 < 2; ++g4) {
                const float4 t = bwd4[(8 + g4) * 32 + lane], s = bwd4[(10 + g4) * 32 + lane], w = bwd4[(12 + g4) * 32 + lane];
                BTc[4 * g4] = t.x; BTc[4 * g4 + 1] = t.y; BTc[4 * g4 + 2] = t.z; BTc[4 * g4 + 3] = t.w;
                Qic[4 * g4] = s.x; Qic[4 * g4 + 1] = s.y; Qic[4 * g4 + 2] = s.z; Qic[4 * g4 + 3] = s.w;
                KTc[4 * g4] = w.x; KTc[4 * g4 + 1] = w.y; KTc[4 * g4 + 2] = w.z; KTc[4 * g4 + 3] = w.w;
            }
            // warm-state mirror (v = vnew, z = znew, d of this sweep) of the slots that run this sweep
            const bool mir[NS] = {warm && (cont[0] || fbw[0]), warm && (cont[1] || fbw[1]), warm && (cont[2] || fbw[2]), warm && (cont[3] || fbw[3])};
            const bool umir = (sl == 0 ? mir[0] : sl == 1 ? mir[1] : sl == 2 ? mir[2] : mir[3]);
            const bool ucont = (sl == 0 ? cont[0] : sl == 1 ? cont[1] : sl == 2 ? cont[2] : cont[3]);
            float *wvo[NS];     // row addresses (only stored through when mir[e])
#pragma unroll
            for (int e = 0; e < NS; ++e) wvo[e] = a.wv + inst[e] * XROW + lane;
            float *wdo = umir ? a.wd + uinst * UROW + ur : nullptr, *wzo = umir ? a.wz + uinst * UROW + ur : nullptr;
            const float *xrf[NS];
#pragma unroll
            for (int e = 0; e < NS; ++e) xrf[e] = a.Xref + inst[e] * a.xref_stride;
            float p[NS];
            {
                float g[NS], v[NS];
#pragma unroll
                for (int e = 0; e < NS; ++e) { ldg_(e, NH - 1, g[e]); ldv_(e, NH - 1, v[e]); }
                asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(g[0]), "+f"(g[1]), "+f"(g[2]), "+f"(g[3]), "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]) :: "memory");
#pragma unroll
                for (int e = 0; e < NS; ++e) {
                    const float pn = ws[e * S::SLOT + S::PN + lane];
                    if (mir[e]) wvo[e][(NH - 1) * WNX] = v[e];
                    const float dvg = __fsub_rn(v[e], g[e]);
                    if constexpr (FAST) p[e] = __fmaf_rn(nrho, dvg, pn);
                    else p[e] = __fsub_rn(pn, __fmul_rn(rho, dvg));                                  // :84
                }
            }
            // stage operands one stage ahead
            float z = myz[(NH - 2) * WNU + ur], y = myy[(NH - 2) * WNU + ur];
            float v[NS], g[NS], xr[NS];
#pragma unroll
            for (int e = 0; e < NS; ++e) { ldg_(e, NH - 2, g[e]); ldv_(e, NH - 2, v[e]); xr[e] = __ldg(xrf[e] + (NH - 2) * WNX + lane); }
#pragma unroll 1
            for (int i = NH - 2; i >= 0; --i) {
                const float r = __fmul_rn(nrho, __fsub_rn(z, y));                                    // :80  (slot sl, row ur)
                asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(g[0]), "+f"(g[1]), "+f"(g[2]), "+f"(g[3]), "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]) :: "memory");
#pragma unroll
                for (int e = 0; e < NS; ++e) {
                    xb[e * WNX + lane] = p[e];
                    xt[e * WNX + sl * 8 + ur] = p[e];       // p(4 j + sl) with j = ur = lane >> 2: packet j of SIMD lane sl
                    if (mir[e]) wvo[e][i * WNX] = v[e];
                }
                ub[sl * WNU + ur] = r;
                if (wzo) wzo[i * WNU] = z;
                __syncwarp();
                // Bdyn^T p_{i+1}: row ur, SIMD lane sl accumulates packets j = 0..7 of e(4j + sl) sequentially, every slot
                float bp[NS];
                {
                    float2 c01, c23;
#pragma unroll
                    for (int g4 = 0; g4 < 2; ++g4) {
                        const float4 v0 = reinterpret_cast<const float4 *>(xt + sl * 8)[g4], v1 = reinterpret_cast<const float4 *>(xt + WNX + sl * 8)[g4],
                                     v2 = reinterpret_cast<const float4 *>(xt + 2 * WNX + sl * 8)[g4], v3 = reinterpret_cast<const float4 *>(xt + 3 * WNX + sl * 8)[g4];
                        const float s0[4] = {v0.x, v0.y, v0.z, v0.w}, s1[4] = {v1.x, v1.y, v1.z, v1.w}, s2[4] = {v2.x, v2.y, v2.z, v2.w},
                                    s3[4] = {v3.x, v3.y, v3.z, v3.w};
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const int j = 4 * g4 + t;
                            if constexpr (FAST) {
                                if (j == 0) { c01 = __fmul2_rn(f2(BTc[0], BTc[0]), f2(s0[0], s1[0])); c23 = __fmul2_rn(f2(BTc[0], BTc[0]), f2(s2[0], s3[0])); }
                                else { c01 = __ffma2_rn(f2(BTc[j], BTc[j]), f2(s0[t], s1[t]), c01); c23 = __ffma2_rn(f2(BTc[j], BTc[j]), f2(s2[t], s3[t]), c23); }
                            } else {
                                const float2 e01 = f2(__fmul_rn(BTc[j], s0[t]), __fmul_rn(BTc[j], s1[t])), e23 = f2(__fmul_rn(BTc[j], s2[t]), __fmul_rn(BTc[j], s3[t]));
                                if (j == 0) { c01 = e01; c23 = e23; }
                                else { c01 = add2(e01, c01); c23 = add2(e23, c23); }
                            }
                        }
                    }
                    // predux across the four SIMD lanes: (l0 + l2) + (l1 + l3)
                    float2 o01 = f2(__shfl_xor_sync(FULLM, c01.x, 2), __shfl_xor_sync(FULLM, c01.y, 2)), o23 = f2(__shfl_xor_sync(FULLM, c23.x, 2), __shfl_xor_sync(FULLM, c23.y, 2));
                    c01 = add2(c01, o01); c23 = add2(c23, o23);
                    o01 = f2(__shfl_xor_sync(FULLM, c01.x, 1), __shfl_xor_sync(FULLM, c01.y, 1)); o23 = f2(__shfl_xor_sync(FULLM, c23.x, 1), __shfl_xor_sync(FULLM, c23.y, 1));
                    c01 = add2(c01, o01); c23 = add2(c23, o23);
                    bp[0] = c01.x; bp[1] = c01.y; bp[2] = c23.x; bp[3] = c23.y;
                }
                // s = Bdyn^T p + r for (slot sl, row ur): the r of that pair is this lane's own
                const float bps = (sl == 0 ? bp[0] : sl == 1 ? bp[1] : sl == 2 ? bp[2] : bp[3]);
                sb[sl * WNU + ur] = __fadd_rn(bps, r);
                // AmBKt p_{i+1} (row lane, scalar half-split tree over 32) and Kinf^T r_i (row lane) of every slot
                float2 m01, m23;
                {
                    float2 q01[8], q23[8];     // sums of 4 consecutive terms: ((e0+e1)+(e2+e3)) per 16-byte chunk
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4) {
                        const float4 v0 = reinterpret_cast<const float4 *>(xb)[g4], v1 = reinterpret_cast<const float4 *>(xb + WNX)[g4],
                                     v2 = reinterpret_cast<const float4 *>(xb + 2 * WNX)[g4], v3 = reinterpret_cast<const float4 *>(xb + 3 * WNX)[g4];
                        const float s0[4] = {v0.x, v0.y, v0.z, v0.w}, s1[4] = {v1.x, v1.y, v1.z, v1.w}, s2[4] = {v2.x, v2.y, v2.z, v2.w},
                                    s3[4] = {v3.x, v3.y, v3.z, v3.w};
                        const float4 mc = bwd4[g4 * 32 + lane];
                        const float Mc[4] = {mc.x, mc.y, mc.z, mc.w};
                        if constexpr (FAST) {   // one FMA chain per slot
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const int k = 4 * g4 + t;
                                if (k == 0) { m01 = __fmul2_rn(f2(Mc[0], Mc[0]), f2(s0[0], s1[0])); m23 = __fmul2_rn(f2(Mc[0], Mc[0]), f2(s2[0], s3[0])); }
                                else { m01 = __ffma2_rn(f2(Mc[t], Mc[t]), f2(s0[t], s1[t]), m01); m23 = __ffma2_rn(f2(Mc[t], Mc[t]), f2(s2[t], s3[t]), m23); }
                            }
                        } else {
                            float2 e01[4], e23[4];
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                e01[t] = f2(__fmul_rn(Mc[t], s0[t]), __fmul_rn(Mc[t], s1[t]));
                                e23[t] = f2(__fmul_rn(Mc[t], s2[t]), __fmul_rn(Mc[t], s3[t]));
                            }
                            q01[g4] = add2(add2(e01[0], e01[1]), add2(e01[2], e01[3]));
                            q23[g4] = add2(add2(e23[0], e23[1]), add2(e23[2], e23[3]));
                        }
                    }
                    if constexpr (!FAST) {
                        // tree(0,32) = (tree(0,8) + tree(8,8)) + (tree(16,8) + tree(24,8)), tree(8 terms) = chunk + chunk
                        m01 = add2(add2(add2(q01[0], q01[1]), add2(q01[2], q01[3])), add2(add2(q01[4], q01[5]), add2(q01[6], q01[7])));
                        m23 = add2(add2(add2(q23[0], q23[1]), add2(q23[2], q23[3])), add2(add2(q23[4], q23[5]), add2(q23[6], q23[7])));
                    }
                }
                float2 k01, k23;
                {
                    float2 e01[8], e23[8];
#pragma unroll
                    for (int g4 = 0; g4 < 2; ++g4) {
                        const float4 v0 = reinterpret_cast<const float4 *>(ub)[g4], v1 = reinterpret_cast<const float4 *>(ub + WNU)[g4],
                                     v2 = reinterpret_cast<const float4 *>(ub + 2 * WNU)[g4], v3 = reinterpret_cast<const float4 *>(ub + 3 * WNU)[g4];
                        const float s0[4] = {v0.x, v0.y, v0.z, v0.w}, s1[4] = {v1.x, v1.y, v1.z, v1.w}, s2[4] = {v2.x, v2.y, v2.z, v2.w},
                                    s3[4] = {v3.x, v3.y, v3.z, v3.w};
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const int k = 4 * g4 + t;
                            e01[k] = f2(__fmul_rn(KTc[k], s0[t]), __fmul_rn(KTc[k], s1[t]));
                            e23[k] = f2(__fmul_rn(KTc[k], s2[t]), __fmul_rn(KTc[k], s3[t]));
                        }
                    }
                    if constexpr (FAST) {
                        k01 = e01[0]; k23 = e23[0];
#pragma unroll
                        for (int k = 1; k < 8; ++k) { k01 = add2(k01, e01[k]); k23 = add2(k23, e23[k]); }
                    } else {
                        // vectorised redux over 8: SIMD lane l sums e(l) + e(4 + l); then (l0 + l2) + (l1 + l3)
                        k01 = add2(add2(add2(e01[0], e01[4]), add2(e01[2], e01[6])), add2(add2(e01[1], e01[5]), add2(e01[3], e01[7])));
                        k23 = add2(add2(add2(e23[0], e23[4]), add2(e23[2], e23[6])), add2(add2(e23[1], e23[5]), add2(e23[3], e23[7])));
                    }
                }
                float q[NS];
#pragma unroll
                for (int e = 0; e < NS; ++e) {
                    const float cq = -__fmul_rn(xr[e], Qd);                                            // :81
                    const float dvg = __fsub_rn(v[e], g[e]);
                    if constexpr (FAST) q[e] = __fmaf_rn(nrho, dvg, cq);
                    else q[e] = __fsub_rn(cq, __fmul_rn(rho, dvg));                                  // :82
                }
                __syncwarp();
                {   // operands of stage i-1
                    const int ip = (i > 0) ? i - 1 : 0;
                    z = myz[ip * WNU + ur]; y = myy[ip * WNU + ur];
#pragma unroll
                    for (int e = 0; e < NS; ++e) { ldg_(e, ip, g[e]); ldv_(e, ip, v[e]); xr[e] = __ldg(xrf[e] + ip * WNX + lane); }
                }
                // d_i = Quu_inv (Bdyn^T p_{i+1} + r_i)                                                  :19  (slot sl, row ur)
                {
                    const float4 s0 = reinterpret_cast<const float4 *>(sb + sl * WNU)[0], s1 = reinterpret_cast<const float4 *>(sb + sl * WNU)[1];
                    const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                    float d;
                    if constexpr (FAST) d = dot<float, ORD_SEQ, WNU, true>([&](int k) { return Qic[k]; }, [&](int k) { return ss[k]; });
                    else {
                        float ed[WNU];
                        prod_pairs<WNU>(Qic, ss, ed, Z);
                        d = sum_seq<WNU>(ed);
                    }
                    if (ucont) myd[i * WNU + ur] = d;
                    if (wdo) wdo[i * WNU] = d;
                }
                {   // p_i = (q_i + AmBKt p_{i+1}) - Kinf^T r_i                                          :20
                    const float2 n01 = sub2(add2(f2(q[0], q[1]), m01), k01), n23 = sub2(add2(f2(q[2], q[3]), m23), k23);
                    p[0] = n01.x; p[1] = n01.y; p[2] = n23.x; p[3] = n23.y;
                }
                __syncwarp();   // xb / xt / ub / sb are rewritten by the next stage
            }
        }
#pragma unroll
        for (int e = 0; e < NS; ++e) if (fbw[e]) finish(e, false);
    }
    if (a.stats && lane == 0) {
        atomicAdd(a.stats + 0, n_iter);
        atomicAdd(a.stats + 1, n_solved);
        atomicAdd(a.stats + 2, n_trips * NS);   // slot-trips (a refilled slot idles through one backward sweep)
        atomicAdd(a.stats + 3, n_inst);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        const uint32_t base = *reinterpret_cast<uint32_t *>(shr + SH::SH_SLOT);
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(base) : "memory");
    }
}

}  // namespace tmpc
