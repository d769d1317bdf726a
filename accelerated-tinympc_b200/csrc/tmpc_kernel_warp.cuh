// Warp-per-instance ADMM kernel for the large shape (nx = 32, nu = 8; BASELINE config 5: N = 50) on sm_100a.
//
// At 32/8/50 one instance carries 17.5 KB of state that survives an iteration (d, y, z: 3 x 392 floats; g, v:
// 2 x 1600), so a thread-per-instance layout would leave 12 threads per SM.  Here ONE WARP owns one instance:
//
//   * lane j owns row j of every nx-vector (x_i, p_i, g_i, v_i, q_i): nx == 32 == warp width, so each horizon
//     stage of the state is one conflict-free 128-byte shared-memory row and one coalesced 128-byte global row;
//     the nu-vectors (u, d, y, z, r, s) are replicated 4x: lanes 4r..4r+3 carry row r;
//   * mat-vec COEFFICIENTS live in registers for the duration of a sweep: lane j holds row j of Adyn / AmBKt /
//     Bdyn / Kinf^T and row (lane >> 2) of Kinf / Quu_inv (forward sweep: 72 registers, backward sweep: 56),
//     fetched once per sweep as coalesced float4 loads from a lane-major image built on the host;
//   * the VECTOR operand is broadcast through a 128-byte shared-memory buffer (one STS.32 + eight LDS.128 per
//     32-long product instead of 32 shuffles);
//   * Bdyn^T p is Eigen's row-major GEMV at this size (4 SIMD lanes x 8 sequential packets per output row, then
//     predux): its 8 x 4 partial chains map one-to-one onto the 32 lanes, the predux is two xor-shuffles;
//   * all per-instance state stays in shared memory across iterations (12 warps = 12 instances per SM fill the
//     227 KB); warps claim instances from a global counter, so early exit needs no lane-level bookkeeping at all;
//   * the four residual maxima are reduced across the warp with redux.sync (max over the uint image of |.| >= 0).
//
// x and u are written to the output on every iteration (coalesced rows that stay L2-resident until the instance
// ends: 7.9 KB per ~15 k issued instructions), so termination needs no emission pass.  With WARM the backward
// sweep mirrors d, v(=vnew), z(=znew) to the caller's buffers exactly like the thread-per-instance kernels, which
// reproduces the reference's "v/z one iteration behind on an early exit" workspace (SURVEY 8a note W).
//
// PARITY orders at this shape (oracle/tinympc_oracle.c select_orders, pinned to the compiled reference):
//   Kinf x, Adyn x, Bdyn u, Quu_inv s: sequential;  Bdyn^T p: row-major GEMV;  AmBKt p: scalar tree over 32;
//   Kinf^T r: vectorised redux over 8;  Xref^T Pinf: vectorised redux over 32.
#pragma once
#include "tmpc_kernel_f32.cuh"

namespace tmpc {

constexpr int WNX = 32, WNU = 8;

// lane-major coefficient image (device memory), float4 = 4 consecutive k for one lane
//   fwd4[(0..7)*32 + lane]   Adyn(lane, 4g..4g+3)
//   fwd4[(8..15)*32 + lane]  Kinf(lane>>2, 4g..4g+3)
//   fwd4[(16..17)*32 + lane] Bdyn(lane, 4g..4g+3)
//   bwd4[(0..7)*32 + lane]   AmBKt(lane, 4g..4g+3)
//   bwd4[(8..9)*32 + lane]   Bdyn^T(lane>>2, 4j + (lane&3)), j = 4g..4g+3     (GEMV partial chain of SIMD lane lane&3)
//   bwd4[(10..11)*32 + lane] Quu_inv(lane>>2, 4g..4g+3)
//   bwd4[(12..13)*32 + lane] Kinf^T(lane, 4g..4g+3) = Kinf(4g.., lane)
struct ModelWarp {
    const float4 *fwd4;
    const float4 *bwd4;
    const float *Pt;            // [32][32]: Pt[k*32 + j] = Pinf(k, j)
    const float *Qd;            // [32]
    const float *xmin, *xmax;   // [NH][32]   (+-inf when the bound is disabled)
    const float *umin, *umax;   // [NH-1][8]
    float rho, nrho, pri_tol, dua_tol;
    int max_iter, check_term;
    float2 nz2;                 // (-0, -0): addend of the exactly rounded packed product (prod2, tmpc_kernel_f32.cuh)
};
constexpr int WARP_FWD4 = 18, WARP_BWD4 = 14;

// K individually rounded products c[k]*x[k].  TMPC_WARP_PROD2 = 1 forms them two per FFMA2 (prod2: here the coefficient
// pairs sit in registers, so the packed product halves the multiply instructions of every mat-vec).  Measured on B200
// (32,768 instances, PARITY, cold): 5.11e7 it/s with scalar FMUL, 4.97e7 it/s packed -- FFMA2 occupies the FMA pipe for two
// cycles and lengthens the dependent chains, and this kernel is latency-bound, not issue-bound.  Default: scalar.
#ifndef TMPC_WARP_PROD2
#define TMPC_WARP_PROD2 0
#endif
#ifndef TMPC_WARP_TREE2
#define TMPC_WARP_TREE2 0   // AmBKt p: the scalar tree's adds issued pairwise as FADD2 (bit-identical, 16 instead of 31
                            // instructions).  Measured on B200 (32,768 instances, PARITY): cold 5.01e7 it/s with it, 5.22e7
                            // without; warm 4.24e7 vs 4.48e7 -- like the packed products, FADD2 holds the FMA pipe for two
                            // cycles and the register pairing costs moves.  Default: scalar tree.
#endif
template <int K> __device__ __forceinline__ void prod_pairs(const float (&c)[K], const float (&x)[K], float (&e)[K], const float2 Z)
{
    static_assert(K % 2 == 0, "pairs");
#pragma unroll
    for (int m = 0; m < K / 2; ++m) {
        if constexpr (TMPC_WARP_PROD2) {
            const float2 r = prod2(f2(c[2 * m], c[2 * m + 1]), f2(x[2 * m], x[2 * m + 1]), Z);
            e[2 * m] = r.x; e[2 * m + 1] = r.y;
        } else {
            e[2 * m] = __fmul_rn(c[2 * m], x[2 * m]); e[2 * m + 1] = __fmul_rn(c[2 * m + 1], x[2 * m + 1]);
        }
    }
}
// Half-split tree over 32 terms (Redux.h redux_novec_unroller) with the adds of each level issued in pairs: the two
// operands of every add are the same as in the scalar tree, so the result is bit-identical; 16 instructions instead of 31.
__device__ __forceinline__ float tree32_pairs(const float (&e)[32])
{
    float2 a[8];   // level 1: (e0+e1, e2+e3), (e4+e5, e6+e7), ...
#pragma unroll
    for (int m = 0; m < 8; ++m) a[m] = add2(f2(e[4 * m], e[4 * m + 2]), f2(e[4 * m + 1], e[4 * m + 3]));
    float2 b[4];   // level 2: ((e0+e1)+(e2+e3), (e4+e5)+(e6+e7)), ...
#pragma unroll
    for (int m = 0; m < 4; ++m) b[m] = add2(f2(a[2 * m].x, a[2 * m + 1].x), f2(a[2 * m].y, a[2 * m + 1].y));
    // level 3: sums of 8: (b0.x+b0.y, b1.x+b1.y), (b2.x+b2.y, b3.x+b3.y)
    const float2 c0 = add2(f2(b[0].x, b[1].x), f2(b[0].y, b[1].y)), c1 = add2(f2(b[2].x, b[3].x), f2(b[2].y, b[3].y));
    // level 4: sums of 16: (c0.x+c0.y, c1.x+c1.y)
    const float2 d = add2(f2(c0.x, c1.x), f2(c0.y, c1.y));
    return __fadd_rn(d.x, d.y);
}

template <int K> __device__ __forceinline__ float sum_seq(const float (&e)[K])
{
    float acc = e[0];
#pragma unroll
    for (int k = 1; k < K; ++k) acc = __fadd_rn(e[k], acc);
    return acc;
}

// TM = g and v (2 x NH x 32 floats, 73 % of the per-instance state) live in TENSOR MEMORY instead of shared memory:
// lane j's TMEM lane holds row j, one column per horizon stage (tcgen05.ld/st.32x32b.x1), 2*NH columns per instance,
// so the 4 warps that share a TMEM lane quadrant hold 4 instances in 400 of its 512 columns.  Shared memory then keeps
// only d, y, z and the broadcast buffers (5 KB per instance): 16 warps = 16 instances per SM instead of 12, and room
// for the bound rows and the coefficient images (read with LDS instead of LDG).
template <int NH, bool TM = false> struct WarpSmem {
    static constexpr int G = 0, V = G + (TM ? 0 : NH * WNX), D = V + (TM ? 0 : NH * WNX), Y = D + (NH - 1) * WNU, Z = Y + (NH - 1) * WNU,
                         PN = Z + (NH - 1) * WNU, XB = PN + WNX, UB = XB + WNX, SB = UB + WNU, FLOATS = SB + WNU;
    static constexpr size_t BYTES = size_t(FLOATS) * 4;
    static_assert(FLOATS % 4 == 0, "per-warp regions stay 16-byte aligned");
    // CTA-shared staging (TM only): fwd4 | bwd4 | xmin | xmax | umin | umax, then the TMEM base slot
    static constexpr int SH_FWD = 0, SH_BWD = SH_FWD + WARP_FWD4 * 32 * 4, SH_XMIN = SH_BWD + WARP_BWD4 * 32 * 4, SH_XMAX = SH_XMIN + NH * WNX,
                         SH_UMIN = SH_XMAX + NH * WNX, SH_UMAX = SH_UMIN + (NH - 1) * WNU, SH_SLOT = SH_UMAX + (NH - 1) * WNU,
                         SH_FLOATS = TM ? SH_SLOT + 4 : 0;
    static constexpr size_t total_bytes(int warps) { return size_t(SH_FLOATS) * 4 + BYTES * warps; }
};

__device__ __forceinline__ void tm_ld1(uint32_t a, float &r) { asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=f"(r) : "r"(a) : "memory"); }
__device__ __forceinline__ void tm_st1(uint32_t a, float r) { asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(a), "f"(r) : "memory"); }
__device__ __forceinline__ void tm_wait_ld2(float &a, float &b) { asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(a), "+f"(b) :: "memory"); }

__device__ __forceinline__ float warp_max_nonneg(float v)
{
    return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(v)));
}

template <int NH, int WARPS, bool FAST, bool WARM, bool TM = false>
__global__ void __launch_bounds__(WARPS * 32, 1)
admm_kernel_warp(const __grid_constant__ ModelWarp P, const __grid_constant__ SolveArgs<float> a)
{
    using S = WarpSmem<NH, TM>;
    static_assert(!TM || (WARPS % 4 == 0 && (WARPS / 4) * 2 * NH <= 512), "TMEM: WARPS/4 instances per lane quadrant, 2*NH columns each");
    constexpr unsigned FULLM = 0xffffffffu;
    constexpr int XROW = WNX * NH, UROW = WNU * (NH - 1);
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ur = lane >> 2;            // the nu-row this lane carries (replicated over 4 lanes)
    const int sl = lane & 3;             // SIMD lane of the reference's packet in the row-major GEMV
    const bool uw = (sl == 0);           // the lane of each group that writes nu-vectors
    float *shr = reinterpret_cast<float *>(smem);                                   // CTA-shared staging (TM)
    float *ws = shr + S::SH_FLOATS + size_t(warp) * S::FLOATS;
    float *sg = ws + S::G, *sv = ws + S::V, *sd = ws + S::D, *sy = ws + S::Y, *sz = ws + S::Z, *spn = ws + S::PN;
    float *xb = ws + S::XB, *ub = ws + S::UB, *sb = ws + S::SB;
    const float4 *xb4 = reinterpret_cast<const float4 *>(xb);
    const float4 *ub4 = reinterpret_cast<const float4 *>(ub);
    const float4 *sb4 = reinterpret_cast<const float4 *>(sb);
    const float Qd = __ldg(P.Qd + lane);
    const float2 Z = P.nz2;
    const bool warm = WARM && a.wd;
    unsigned long long n_iter = 0, n_solved = 0, n_inst = 0;

    // model sources: global memory (read-only path), or their shared-memory copies when TM leaves the room
    const float4 *fwd4 = P.fwd4, *bwd4 = P.bwd4;
    const float *bxmin = P.xmin, *bxmax = P.xmax, *bumin = P.umin, *bumax = P.umax;
    uint32_t gcol = 0, vcol = 0;   // TMEM addresses of this warp's g / v columns (stage i at +i)
    if constexpr (TM) {
        uint32_t *slot = reinterpret_cast<uint32_t *>(shr + S::SH_SLOT);
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        for (int k = threadIdx.x; k < WARP_FWD4 * 32; k += WARPS * 32) reinterpret_cast<float4 *>(shr + S::SH_FWD)[k] = __ldg(P.fwd4 + k);
        for (int k = threadIdx.x; k < WARP_BWD4 * 32; k += WARPS * 32) reinterpret_cast<float4 *>(shr + S::SH_BWD)[k] = __ldg(P.bwd4 + k);
        for (int k = threadIdx.x; k < NH * WNX; k += WARPS * 32) { shr[S::SH_XMIN + k] = __ldg(P.xmin + k); shr[S::SH_XMAX + k] = __ldg(P.xmax + k); }
        for (int k = threadIdx.x; k < (NH - 1) * WNU; k += WARPS * 32) { shr[S::SH_UMIN + k] = __ldg(P.umin + k); shr[S::SH_UMAX + k] = __ldg(P.umax + k); }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t base = *slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 2 * NH);
        gcol = base; vcol = base + NH;
        fwd4 = reinterpret_cast<const float4 *>(shr + S::SH_FWD); bwd4 = reinterpret_cast<const float4 *>(shr + S::SH_BWD);
        bxmin = shr + S::SH_XMIN; bxmax = shr + S::SH_XMAX; bumin = shr + S::SH_UMIN; bumax = shr + S::SH_UMAX;
    }
    // state rows g_i[lane], v_i[lane]: shared memory, or TMEM columns (loads complete at gv_wait)
    auto gv_load = [&](int i, float &g, float &v) {
        if constexpr (TM) { tm_ld1(gcol + i, g); tm_ld1(vcol + i, v); }
        else { g = sg[i * WNX + lane]; v = sv[i * WNX + lane]; }
    };
    auto gv_wait = [&](float &g, float &v) { if constexpr (TM) tm_wait_ld2(g, v); };
    auto gv_store = [&](int i, float g, float v) {
        if constexpr (TM) { tm_st1(gcol + i, g); tm_st1(vcol + i, v); }
        else { sg[i * WNX + lane] = g; sv[i * WNX + lane] = v; }
    };
    auto gv_fence = [&]() { if constexpr (TM) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); };
    auto ld4 = [&](const float4 *p) -> float4 { if constexpr (TM) return *p; else return __ldg(p); };
    auto ld1 = [&](const float *p) -> float { if constexpr (TM) return *p; else return __ldg(p); };

    for (;;) {
        long long inst;
        {
            unsigned long long b = 0;
            if (lane == 0) b = atomicAdd(a.counter, 1ull);
            inst = (long long)__shfl_sync(FULLM, b, 0);
        }
        if (inst >= a.batch) break;
        inst = claim_instance(a, inst);
        if (inst < 0) break;
        const float *xref = a.Xref + inst * a.xref_stride;
        const float x0 = __ldg(a.x0 + inst * WNX + lane);
        // p_N seed: -(Xref_{N-1}^T Pinf)   (admm.cpp:83)
        {
            const float *xl = xref + (NH - 1) * WNX;
            const float pn = -dot<float, ORD_VECREDUX, WNX, FAST>([&](int k) { return __ldg(P.Pt + k * WNX + lane); },
                                                                   [&](int k) { return __ldg(xl + k); });
            spn[lane] = pn;
        }
        if (warm) {
            const float *gd = a.wd + inst * UROW, *gy = a.wy + inst * UROW, *gz = a.wz + inst * UROW;
            const float *gg = a.wg + inst * XROW, *gv = a.wv + inst * XROW;
            for (int k = lane; k < UROW; k += 32) { sd[k] = gd[k]; sy[k] = gy[k]; sz[k] = gz[k]; }
#pragma unroll 2
            for (int i = 0; i < NH; ++i) gv_store(i, gg[i * WNX + lane], gv[i * WNX + lane]);
        } else {
            for (int k = lane; k < UROW; k += 32) { sd[k] = 0.f; sy[k] = 0.f; sz[k] = 0.f; }
#pragma unroll 2
            for (int i = 0; i < NH; ++i) gv_store(i, 0.f, 0.f);
        }
        gv_fence();
        __syncwarp();

        float *xo = a.x ? a.x + inst * XROW : nullptr;
        float *uo = a.u ? a.u + inst * UROW : nullptr;
        int it = 0;
        bool conv = false;
        float res[4] = {0.f, 0.f, 0.f, 0.f};

        for (;;) {
            ++it;
            // -------------------------------------------------------------- forward sweep
            // forward_pass (admm.cpp:27-37) + update_slack (:45-61) + update_dual (:67-71) + residual maxima (:95-98)
            float pri_x = 0.f, dua_x = 0.f, pri_u = 0.f, dua_u = 0.f;
            {
                float Ac[WNX], Kc[WNX], Bc[WNU];
#pragma unroll
                for (int g4 = 0; g4 < 8; ++g4) {
                    const float4 t = ld4(fwd4 + g4 * 32 + lane), s = ld4(fwd4 + (8 + g4) * 32 + lane);
                    Ac[4 * g4] = t.x; Ac[4 * g4 + 1] = t.y; Ac[4 * g4 + 2] = t.z; Ac[4 * g4 + 3] = t.w;
                    Kc[4 * g4] = s.x; Kc[4 * g4 + 1] = s.y; Kc[4 * g4 + 2] = s.z; Kc[4 * g4 + 3] = s.w;
                }
#pragma unroll
                for (int g4 = 0; g4 < 2; ++g4) {
                    const float4 t = ld4(fwd4 + (16 + g4) * 32 + lane);
                    Bc[4 * g4] = t.x; Bc[4 * g4 + 1] = t.y; Bc[4 * g4 + 2] = t.z; Bc[4 * g4 + 3] = t.w;
                }
                float x = x0;
                // stage operands are fetched one stage ahead (shared-memory state and the global bound rows), so
                // their latency overlaps the previous stage's Bdyn u chain instead of heading the dependent chain
                float xmn = ld1(bxmin + lane), xmx = ld1(bxmax + lane);
                float umn = ld1(bumin + ur), umx = ld1(bumax + ur);
                float g, v, d = sd[ur], y = sy[ur], z = sz[ur];
                gv_load(0, g, v);
#pragma unroll 1
                for (int i = 0; i < NH - 1; ++i) {
                    xb[lane] = x;
                    __syncwarp();
                    gv_wait(g, v);
                    if (xo) xo[i * WNX + lane] = x;
                    // [Kinf(ur,:) ; Adyn(lane,:)] x_i : two sequential chains advancing together as one float2
                    float2 ka;
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4) {
                        const float4 xv = xb4[g4];
                        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const int k = 4 * g4 + t;
                            if constexpr (FAST) {
                                if (k == 0) ka = __fmul2_rn(f2(Kc[0], Ac[0]), f2(xs[0], xs[0]));
                                else ka = __ffma2_rn(f2(Kc[k], Ac[k]), f2(xs[t], xs[t]), ka);
                            } else {
                                const float2 e = TMPC_WARP_PROD2 ? prod2(f2(Kc[k], Ac[k]), f2(xs[t], xs[t]), Z)
                                                                 : f2(__fmul_rn(Kc[k], xs[t]), __fmul_rn(Ac[k], xs[t]));
                                if (k == 0) ka = e;
                                else ka = add2(e, ka);
                            }
                        }
                    }
                    const float u = __fsub_rn(-ka.x, d);                                               // :31
                    // slack / dual / residuals for (x_i[lane], u_i[ur]) as one float2
                    const float2 xu = f2(x, u), gy = f2(g, y), vz = f2(v, z);
                    float2 t = add2(xu, gy);                                                           // :47-48
                    t.x = fminf(xmx, fmaxf(xmn, t.x));                                                 // :59
                    t.y = fminf(umx, fmaxf(umn, t.y));                                                 // :53
                    const float2 rp = sub2(xu, t), rd = sub2(vz, t);
                    pri_x = fmaxf(pri_x, fabsf(rp.x)); pri_u = fmaxf(pri_u, fabsf(rp.y));              // :95,:97
                    dua_x = fmaxf(dua_x, fabsf(rd.x)); dua_u = fmaxf(dua_u, fabsf(rd.y));              // :96,:98
                    const float2 gyn = sub2(add2(gy, xu), t);                                          // :69-70
                    gv_store(i, gyn.x, t.x);
                    if (uw) {
                        sy[i * WNU + ur] = gyn.y;
                        sz[i * WNU + ur] = t.y;
                        ub[ur] = u;
                        if (uo) uo[i * WNU + ur] = u;
                        if (i == 0 && a.u0) a.u0[inst * WNU + ur] = u;
                    }
                    __syncwarp();
                    {   // operands of stage i+1 (the nu-rows of the last stage do not exist: re-read stage i's)
                        const int in = i + 1, iu = (in < NH - 1) ? in : i;
                        xmn = ld1(bxmin + in * WNX + lane); xmx = ld1(bxmax + in * WNX + lane);
                        umn = ld1(bumin + iu * WNU + ur); umx = ld1(bumax + iu * WNU + ur);
                        gv_load(in, g, v);
                        d = sd[iu * WNU + ur]; y = sy[iu * WNU + ur]; z = sz[iu * WNU + ur];
                    }
                    // x_{i+1} = Adyn x_i + Bdyn u_i                                                     :35
                    const float4 u0 = ub4[0], u1 = ub4[1];
                    const float us[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
                    float bu;
                    if constexpr (FAST) {
                        bu = ka.y;
#pragma unroll
                        for (int k = 0; k < WNU; ++k) bu = __fmaf_rn(Bc[k], us[k], bu);
                        x = bu;
                    } else {
                        float eb[WNU];
                        prod_pairs<WNU>(Bc, us, eb, Z);
                        bu = sum_seq<WNU>(eb);
                        x = __fadd_rn(ka.y, bu);
                    }
                }
                {   // last stage: state slack / dual only
                    constexpr int i = NH - 1;
                    gv_wait(g, v);
                    if (xo) xo[i * WNX + lane] = x;
                    float t = __fadd_rn(x, g);
                    t = fminf(xmx, fmaxf(xmn, t));
                    pri_x = fmaxf(pri_x, fabsf(__fsub_rn(x, t)));
                    dua_x = fmaxf(dua_x, fabsf(__fsub_rn(v, t)));
                    gv_store(i, __fsub_rn(__fadd_rn(g, x), t), t);
                    gv_fence();
                }
            }
            // -------------------------------------------------------------- termination (admm.cpp:91-109, :135-138)
            const bool chk = (it % P.check_term) == 0;
            if (chk) {
                res[0] = warp_max_nonneg(pri_x);
                res[1] = __fmul_rn(warp_max_nonneg(dua_x), P.rho);
                res[2] = warp_max_nonneg(pri_u);
                res[3] = __fmul_rn(warp_max_nonneg(dua_u), P.rho);
            }
            conv = chk && res[0] < P.pri_tol && res[2] < P.pri_tol && res[1] < P.dua_tol && res[3] < P.dua_tol;
            if (conv) break;
            const bool last = it >= P.max_iter;
            if (last && !warm) break;   // the final backward pass only matters for the warm state it leaves behind

            // -------------------------------------------------------------- backward sweep
            // update_linear_cost (admm.cpp:77-85) recomputed per stage + backward_pass_grad (:15-22)
            __syncwarp();   // the forward sweep's last reads of ub precede this sweep's writes
            {
                float Mc[WNX], BTc[WNU], Qic[WNU], KTc[WNU];
#pragma unroll
                for (int g4 = 0; g4 < 8; ++g4) {
                    const float4 t = ld4(bwd4 + g4 * 32 + lane);
                    Mc[4 * g4] = t.x; Mc[4 * g4 + 1] = t.y; Mc[4 * g4 + 2] = t.z; Mc[4 * g4 + 3] = t.w;
                }
#pragma unroll
                for (int g4 = 0; g4 < 2; ++g4) {
                    const float4 t = ld4(bwd4 + (8 + g4) * 32 + lane), s = ld4(bwd4 + (10 + g4) * 32 + lane),
                                 w = ld4(bwd4 + (12 + g4) * 32 + lane);
                    BTc[4 * g4] = t.x; BTc[4 * g4 + 1] = t.y; BTc[4 * g4 + 2] = t.z; BTc[4 * g4 + 3] = t.w;
                    Qic[4 * g4] = s.x; Qic[4 * g4 + 1] = s.y; Qic[4 * g4 + 2] = s.z; Qic[4 * g4 + 3] = s.w;
                    KTc[4 * g4] = w.x; KTc[4 * g4 + 1] = w.y; KTc[4 * g4 + 2] = w.z; KTc[4 * g4 + 3] = w.w;
                }
                float *wdo = warm ? a.wd + inst * UROW : nullptr;
                float *wvo = warm ? a.wv + inst * XROW : nullptr;
                float *wzo = warm ? a.wz + inst * UROW : nullptr;
                float p;
                {
                    float v, g;
                    gv_load(NH - 1, g, v);
                    const float pn = spn[lane];
                    gv_wait(g, v);
                    if (WARM && wvo) wvo[(NH - 1) * WNX + lane] = v;
                    const float dvg = __fsub_rn(v, g);
                    if constexpr (FAST) p = __fmaf_rn(P.nrho, dvg, pn);
                    else p = __fsub_rn(pn, __fmul_rn(P.rho, dvg));                                     // :84
                }
                // stage operands one stage ahead, as in the forward sweep
                float z = sz[(NH - 2) * WNU + ur], y = sy[(NH - 2) * WNU + ur];
                float v, g;
                gv_load(NH - 2, g, v);
                float xr = __ldg(xref + (NH - 2) * WNX + lane);
#pragma unroll 1
                for (int i = NH - 2; i >= 0; --i) {
                    const float r = __fmul_rn(P.nrho, __fsub_rn(z, y));                                // :80
                    gv_wait(g, v);
                    xb[lane] = p;
                    if (uw) ub[ur] = r;
                    if (WARM && wvo) {
                        wvo[i * WNX + lane] = v;
                        if (uw) wzo[i * WNU + ur] = z;
                    }
                    __syncwarp();
                    // Bdyn^T p_{i+1}: row ur, SIMD lane sl accumulates packets j = 0..7 of e(4j + sl) sequentially
                    float bp;
                    if constexpr (FAST) {
                        bp = __fmul_rn(BTc[0], xb[sl]);
#pragma unroll
                        for (int j = 1; j < 8; ++j) bp = __fmaf_rn(BTc[j], xb[4 * j + sl], bp);
                    } else {
                        float pj[8], ej[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) pj[j] = xb[4 * j + sl];
                        prod_pairs<8>(BTc, pj, ej, Z);
                        bp = sum_seq<8>(ej);
                    }
                    bp = __fadd_rn(bp, __shfl_xor_sync(FULLM, bp, 2));       // (l0+l2), (l1+l3)
                    bp = __fadd_rn(bp, __shfl_xor_sync(FULLM, bp, 1));       // (l0+l2)+(l1+l3)
                    const float s = __fadd_rn(bp, r);
                    if (uw) sb[ur] = s;
                    // AmBKt p_{i+1} (row lane) and Kinf^T r_i (row lane) while the s vector settles
                    float pv[WNX];
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4) {
                        const float4 t = xb4[g4];
                        pv[4 * g4] = t.x; pv[4 * g4 + 1] = t.y; pv[4 * g4 + 2] = t.z; pv[4 * g4 + 3] = t.w;
                    }
                    float mp;
                    if constexpr (FAST) mp = dot<float, ORD_TREE, WNX, true>([&](int k) { return Mc[k]; }, [&](int k) { return pv[k]; });
                    else {
                        float em[WNX];
                        prod_pairs<WNX>(Mc, pv, em, Z);
                        if constexpr (TMPC_WARP_TREE2) mp = tree32_pairs(em);
                        else mp = red_tree<float, 0, WNX>([&](int k) { return em[k]; });
                    }
                    const float4 r0 = ub4[0], r1 = ub4[1];
                    const float rs[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
                    float kr;
                    if constexpr (FAST) kr = dot<float, ORD_VECREDUX, WNU, true>([&](int k) { return KTc[k]; }, [&](int k) { return rs[k]; });
                    else {
                        float ek[WNU];
                        prod_pairs<WNU>(KTc, rs, ek, Z);
                        const float2 l01 = add2(f2(ek[0], ek[1]), f2(ek[4], ek[5])), l23 = add2(f2(ek[2], ek[3]), f2(ek[6], ek[7]));
                        const float2 t = add2(l01, l23);                     // (l0+l2, l1+l3)
                        kr = __fadd_rn(t.x, t.y);
                    }
                    const float cq = -__fmul_rn(xr, Qd);                                               // :81
                    const float dvg = __fsub_rn(v, g);
                    float q;
                    if constexpr (FAST) q = __fmaf_rn(P.nrho, dvg, cq);
                    else q = __fsub_rn(cq, __fmul_rn(P.rho, dvg));                                     // :82
                    __syncwarp();
                    {   // operands of stage i-1
                        const int ip = (i > 0) ? i - 1 : 0;
                        z = sz[ip * WNU + ur]; y = sy[ip * WNU + ur];
                        gv_load(ip, g, v);
                        xr = __ldg(xref + ip * WNX + lane);
                    }
                    // d_i = Quu_inv (Bdyn^T p_{i+1} + r_i)                                              :19
                    const float4 s0 = sb4[0], s1 = sb4[1];
                    const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                    float d;
                    if constexpr (FAST) d = dot<float, ORD_SEQ, WNU, true>([&](int k) { return Qic[k]; }, [&](int k) { return ss[k]; });
                    else {
                        float ed[WNU];
                        prod_pairs<WNU>(Qic, ss, ed, Z);
                        d = sum_seq<WNU>(ed);
                    }
                    if (uw) {
                        if (!last) sd[i * WNU + ur] = d;
                        if (WARM && wdo) wdo[i * WNU + ur] = d;
                    }
                    p = __fsub_rn(__fadd_rn(q, mp), kr);                                               // :20
                    __syncwarp();   // xb / ub / sb are rewritten by the next stage
                }
            }
            if (last) break;
        }

        // ------------------------------------------------------------------ outputs
        if (warm) {   // y, g as the reference leaves them (this iteration's); d, v, z were mirrored by the backward sweeps
            float *gy = a.wy + inst * UROW, *gg = a.wg + inst * XROW;
            __syncwarp();
            for (int k = lane; k < UROW; k += 32) gy[k] = sy[k];
#pragma unroll 2
            for (int i = 0; i < NH; ++i) {
                float g, v;
                gv_load(i, g, v);
                gv_wait(g, v);
                gg[i * WNX + lane] = g;
            }
        }
        if (lane == 0) {
            if (a.iter) a.iter[inst] = it;
            if (a.status) a.status[inst] = conv ? 1 : 11;
            if (a.resid) *reinterpret_cast<float4 *>(a.resid + inst * 4) = make_float4(res[0], res[1], res[2], res[3]);
        }
        n_iter += (unsigned)it; n_solved += conv ? 1u : 0u; ++n_inst;
        if (a.done) {
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(a.done + (inst >> a.done_shift), 1u);
        }
        __syncwarp();   // the next instance overwrites this warp's shared-memory state
    }
    if (a.stats && lane == 0) {
        atomicAdd(a.stats + 0, n_iter);
        atomicAdd(a.stats + 1, n_solved);
        atomicAdd(a.stats + 2, n_iter);   // lane-trips == iterations: no emission trip, no idle lanes
        atomicAdd(a.stats + 3, n_inst);
    }
    if constexpr (TM) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) {
            const uint32_t base = *reinterpret_cast<uint32_t *>(shr + S::SH_SLOT);
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(base) : "memory");
        }
    }
}

}  // namespace tmpc
