// fp32 kernel for SMALL single-input shapes (nu = 1, nx a multiple of 4; compiled: cartpole 4/1/10, config 4):
// same algorithm and bit-exact results as tmpc_kernel.cuh, organised for a shape whose whole ADMM state is a
// hundred scalars.
//
//  * the generic kernel keeps {d,y,z,g,v} in shared memory and walks the horizon with a run-time stage index; at
//    4/1/10 an iteration is only ~1400 FP instructions, so the LDS/STS round trips, the address arithmetic and the
//    bound look-ups are a third of everything it issues (ncu: issue slots 88 % busy, FMA pipe 55 %).  Here the state
//    (111 scalars + the p_N seed) lives in REGISTERS for the life of the instance and both sweeps are fully unrolled
//    over the horizon: every state access is a register, every coefficient and bound a fixed constant-bank operand;
//  * element-wise work and the nx-row accumulations run on row pairs as FADD2 (products stay scalar FMUL, rounded
//    individually: the evaluation order of the reference is kept, see Orders<> in tmpc_kernel.cuh);
//  * termination is predicted one trip ahead as in the 12/4/10 kernel, so most instances need no emission trip.
#pragma once
#include "tmpc_kernel_f32.cuh"

namespace tmpc {

template <int NX, int NH, int BLOCK, bool FAST, bool WARM>
__global__ void __launch_bounds__(BLOCK, 1)
admm_kernel_small(const __grid_constant__ Model<float, NX, 1, NH> P, const __grid_constant__ SolveArgs<float> a)
{
    static_assert(NX % 4 == 0, "16-byte rows of x");
    using O = Orders<float, NX, 1>;
    static_assert(O::Ax == ORD_SEQ && O::Mp == ORD_SEQ && O::Bu == ORD_SEQ && O::Qs == ORD_SEQ, "row sweeps are sequential at nu = 1");
    constexpr int H = NX / 2;
    constexpr unsigned FULLM = 0xffffffffu;
    constexpr int XROW = NX * NH, UROW = NH - 1;
    const float2 Z = make_float2(-0.f, -0.f);   // only reaches prod2 when a TMPC_PROD2_* switch is on
    const unsigned lane = threadIdx.x & 31;

    // ---- per-instance state, registers
    float2 g[NH][H], v[NH][H], pn[H];
    float d[NH - 1], y[NH - 1], z[NH - 1];
    float x0[NX];
#pragma unroll
    for (int i = 0; i < NH; ++i)
#pragma unroll
        for (int j = 0; j < H; ++j) g[i][j] = v[i][j] = f2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NH - 1; ++i) d[i] = y[i] = z[i] = 0.f;
#pragma unroll
    for (int j = 0; j < H; ++j) pn[j] = f2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < NX; ++j) x0[j] = 0.f;

    long long inst = -1;
    int it = 0;
    int phase = PH_FREE;
    bool exhausted = false;
    bool spec = false;
    float res[4] = {0.f, 0.f, 0.f, 0.f};
    unsigned long long n_iter = 0, n_solved = 0, n_trips = 0, n_inst = 0;

    for (;;) {
        // ------------------------------------------------------------------ lane refill
        const bool need = (phase == PH_FREE) && !exhausted;
        const unsigned m = __ballot_sync(FULLM, need);
        if (m) {
            const int leader = __ffs(m) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(a.counter, (unsigned long long)__popc(m));
            base = __shfl_sync(FULLM, base, leader);
            if (need) {
                const long long idx = (long long)base + __popc(m & ((1u << lane) - 1u));
                const long long ci = idx < a.batch ? claim_instance(a, idx) : -1;
                if (ci >= 0) {
                    inst = ci; phase = PH_RUN; it = 0;
                    spec = (P.max_iter <= 1);
                    res[0] = res[1] = res[2] = res[3] = 0.f;
                    gload<float, NX>(a.x0 + inst * NX, x0);
                    float xr[NX];
                    gload<float, NX>(a.Xref + inst * a.xref_stride + (NH - 1) * NX, xr);
#pragma unroll
                    for (int j = 0; j < NX; ++j) {   // p_N seed: -(Xref_{N-1}^T Pinf)  (admm.cpp:83)
                        const float t = -dot<float, O::XtP, NX, FAST>([&](int k) { return P.Pf[k + j * NX]; }, [&](int k) { return xr[k]; });
                        if (j & 1) pn[j / 2].y = t; else pn[j / 2].x = t;
                    }
                    if (WARM && a.wd) {
#pragma unroll
                        for (int i = 0; i < NH - 1; ++i) {
                            d[i] = __ldg(a.wd + inst * UROW + i);
                            y[i] = __ldg(a.wy + inst * UROW + i);
                            z[i] = __ldg(a.wz + inst * UROW + i);
                        }
#pragma unroll
                        for (int i = 0; i < NH; ++i) {
                            float tg[NX], tv[NX];
                            gload<float, NX>(a.wg + inst * XROW + i * NX, tg);
                            gload<float, NX>(a.wv + inst * XROW + i * NX, tv);
#pragma unroll
                            for (int j = 0; j < H; ++j) { g[i][j] = f2(tg[2 * j], tg[2 * j + 1]); v[i][j] = f2(tv[2 * j], tv[2 * j + 1]); }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < NH - 1; ++i) d[i] = y[i] = z[i] = 0.f;
#pragma unroll
                        for (int i = 0; i < NH; ++i)
#pragma unroll
                            for (int j = 0; j < H; ++j) g[i][j] = v[i][j] = f2(0.f, 0.f);
                    }
                } else {
                    exhausted = true;
                }
            }
        }
        if (__all_sync(FULLM, phase == PH_FREE)) break;
        ++n_trips;

        const bool emit = (phase == PH_EMIT);
        if (phase == PH_RUN) ++it;

        // ------------------------------------------------------------------ forward sweep
        // forward_pass (admm.cpp:27-37) + update_slack (:45-61) + update_dual (:67-71) + residual maxima (:95-98)
        float pri_x = 0.f, dua_x = 0.f, pri_u = 0.f, dua_u = 0.f;
        {
            float x[NX];
#pragma unroll
            for (int j = 0; j < NX; ++j) x[j] = x0[j];
            const bool wr = emit || (spec && phase == PH_RUN);
            float *xo = (wr && a.x) ? a.x + inst * XROW : nullptr;
            float *uo = (wr && a.u) ? a.u + inst * UROW : nullptr;
            float *go = (WARM && wr && a.wg) ? a.wg + inst * XROW : nullptr;
            float *yo = (WARM && wr && a.wy) ? a.wy + inst * UROW : nullptr;

#pragma unroll
            for (int i = 0; i < NH; ++i) {
                float2 xn[H];
                if (i < NH - 1) {
                    // u_i = -(Kinf x_i) - d_i                                                        :31
                    const float kx = dot<float, O::Kx, NX, FAST>([&](int k) { return P.K[k]; }, [&](int k) { return x[k]; });
                    const float u = __fsub_rn(-kx, d[i]);
                    if (WARM && yo && emit) yo[i] = y[i];
                    float zn = __fadd_rn(u, y[i]);                                                   // :47
                    zn = fminf(P.umax[i], fmaxf(P.umin[i], zn));                                     // :53
                    pri_u = fmaxf(pri_u, fabsf(__fsub_rn(u, zn)));                                   // :97
                    dua_u = fmaxf(dua_u, fabsf(__fsub_rn(z[i], zn)));                                // :98
                    y[i] = __fsub_rn(__fadd_rn(y[i], u), zn);                                        // :69
                    z[i] = zn;
                    if (WARM && yo && !emit) yo[i] = y[i];
                    if (uo) uo[i] = u;
                    if (i == 0 && wr && a.u0) a.u0[inst] = u;
                    // x_{i+1} = A x_i + B u_i                                                        :35
                    float2 ax[H];
                    matvec2<ORD_SEQ, NX, NX, NX, 0, FAST>(P.A, x, ax, Z);
#pragma unroll
                    for (int j = 0; j < H; ++j) {
                        const float2 b2 = f2(P.B[2 * j], P.B[2 * j + 1]);
                        if constexpr (FAST) xn[j] = __ffma2_rn(b2, f2(u, u), ax[j]);
                        else xn[j] = add2(ax[j], f2(__fmul_rn(b2.x, u), __fmul_rn(b2.y, u)));
                    }
                }
                // state slack / dual / residuals of stage i
                if (WARM && go && emit) {
                    float t[NX];
#pragma unroll
                    for (int j = 0; j < H; ++j) { t[2 * j] = g[i][j].x; t[2 * j + 1] = g[i][j].y; }
                    gstore<float, NX>(go + i * NX, t);
                }
#pragma unroll
                for (int j = 0; j < H; ++j) {
                    const float2 x2 = f2(x[2 * j], x[2 * j + 1]);
                    float2 t = add2(x2, g[i][j]);                                                    // :48
                    t.x = fminf(P.xmax[i * NX + 2 * j], fmaxf(P.xmin[i * NX + 2 * j], t.x));         // :59
                    t.y = fminf(P.xmax[i * NX + 2 * j + 1], fmaxf(P.xmin[i * NX + 2 * j + 1], t.y));
                    const float2 rp = sub2(x2, t), rd = sub2(v[i][j], t);
                    pri_x = fmaxf(pri_x, fmaxf(fabsf(rp.x), fabsf(rp.y)));                           // :95
                    dua_x = fmaxf(dua_x, fmaxf(fabsf(rd.x), fabsf(rd.y)));                           // :96
                    g[i][j] = sub2(add2(g[i][j], x2), t);                                            // :70
                    v[i][j] = t;
                }
                if (WARM && go && !emit) {
                    float t[NX];
#pragma unroll
                    for (int j = 0; j < H; ++j) { t[2 * j] = g[i][j].x; t[2 * j + 1] = g[i][j].y; }
                    gstore<float, NX>(go + i * NX, t);
                }
                if (xo) gstore<float, NX>(xo + i * NX, x);
                if (i < NH - 1) {
#pragma unroll
                    for (int j = 0; j < H; ++j) { x[2 * j] = xn[j].x; x[2 * j + 1] = xn[j].y; }
                }
            }
        }

        // ------------------------------------------------------------------ termination (admm.cpp:91-109, :135-138)
        bool final_bwd = false;
        bool finished = false;
        if (phase == PH_RUN) {
            const bool chk = (it % P.check_term) == 0;
            if (chk) {
                res[0] = pri_x; res[1] = __fmul_rn(dua_x, P.rho); res[2] = pri_u; res[3] = __fmul_rn(dua_u, P.rho);
            }
            const bool conv = chk && res[0] < P.pri_tol && res[2] < P.pri_tol && res[1] < P.dua_tol && res[3] < P.dua_tol;
            if (conv || it >= P.max_iter) {
                if (a.iter) a.iter[inst] = it;
                if (a.status) a.status[inst] = conv ? 1 : 11;
                if (a.resid) *reinterpret_cast<float4 *>(a.resid + inst * 4) = make_float4(res[0], res[1], res[2], res[3]);
                n_iter += (unsigned)it; n_solved += conv ? 1u : 0u; ++n_inst;
                final_bwd = !conv;
                if (spec) { phase = PH_FREE; finished = true; }   // x,u of this very trip are already in the output
                else phase = PH_EMIT;
                spec = false;
            } else {
                constexpr float SF = TMPC_SPEC_FACTOR;
                const bool next_chk = ((it + 1) % P.check_term) == 0;
                spec = (it + 1 >= P.max_iter) ||
                       (next_chk && res[0] < SF * P.pri_tol && res[2] < SF * P.pri_tol && res[1] < SF * P.dua_tol && res[3] < SF * P.dua_tol);
            }
        } else if (phase == PH_EMIT) {
            phase = PH_FREE;
            finished = true;
        }

        // ------------------------------------------------------------------ backward sweep
        // update_linear_cost (admm.cpp:77-85) recomputed per stage + backward_pass_grad (:15-22)
        const bool cont = (phase == PH_RUN);
        const bool wout = WARM && (cont || final_bwd) && a.wd;
        if (__any_sync(FULLM, cont || wout)) {
            float p[NX];
            const float *xr_base = a.Xref + (inst < 0 ? 0 : inst) * a.xref_stride;
            float *wdo = wout ? a.wd + inst * UROW : nullptr;
            float *wvo = wout ? a.wv + inst * XROW : nullptr;
            float *wzo = wout ? a.wz + inst * UROW : nullptr;
            auto store_v = [&](int i) {
                float t[NX];
#pragma unroll
                for (int j = 0; j < H; ++j) { t[2 * j] = v[i][j].x; t[2 * j + 1] = v[i][j].y; }
                gstore<float, NX>(wvo + i * NX, t);
            };
            if (WARM && wvo) store_v(NH - 1);
#pragma unroll
            for (int j = 0; j < H; ++j) {
                const float2 dvg = sub2(v[NH - 1][j], g[NH - 1][j]);
                float2 t;
                if constexpr (FAST) t = __ffma2_rn(f2(P.nrho, P.nrho), dvg, pn[j]);
                else t = sub2(pn[j], f2(__fmul_rn(P.rho, dvg.x), __fmul_rn(P.rho, dvg.y)));          // :84
                p[2 * j] = t.x; p[2 * j + 1] = t.y;
            }
#pragma unroll
            for (int i = NH - 2; i >= 0; --i) {
                float xr[NX];
                gload<float, NX>(xr_base + i * NX, xr);
                if (WARM && wvo) { store_v(i); wzo[i] = z[i]; }
                const float r = __fmul_rn(P.nrho, __fsub_rn(z[i], y[i]));                            // :80
                // d_i = Quu_inv (B^T p_{i+1} + r_i)                                                  :19
                const float bp = dot<float, O::Btp, NX, FAST>([&](int k) { return P.B[k]; }, [&](int k) { return p[k]; });
                const float dn = __fmul_rn(P.Qi[0], __fadd_rn(bp, r));
                if (cont) d[i] = dn;
                if (WARM && wdo) wdo[i] = dn;
                // p_i = q_i + AmBKt p_{i+1} - Kinf^T r_i                                             :20
                float2 mp[H];
                matvec2<ORD_SEQ, NX, NX, NX, 0, FAST>(P.M, p, mp, Z);
#pragma unroll
                for (int j = 0; j < H; ++j) {
                    const float2 cq = neg2(f2(__fmul_rn(xr[2 * j], P.Qd[2 * j]), __fmul_rn(xr[2 * j + 1], P.Qd[2 * j + 1])));   // :81
                    const float2 dvg = sub2(v[i][j], g[i][j]);
                    float2 q;
                    if constexpr (FAST) q = __ffma2_rn(f2(P.nrho, P.nrho), dvg, cq);
                    else q = sub2(cq, f2(__fmul_rn(P.rho, dvg.x), __fmul_rn(P.rho, dvg.y)));         // :82
                    const float2 kr = f2(__fmul_rn(P.K[2 * j], r), __fmul_rn(P.K[2 * j + 1], r));
                    const float2 t = sub2(add2(q, mp[j]), kr);
                    p[2 * j] = t.x; p[2 * j + 1] = t.y;
                }
            }
        }
        if (finished && a.done) { __threadfence(); atomicAdd(a.done + (inst >> a.done_shift), 1u); }
    }

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
}

// Fused closed loop (SolveArgs::roll_steps MPC steps per claimed instance, the examples' loop codegen_cartpole.cpp:75-122 /
// quadrotor_hovering.cpp:90-114 with reset duals and a fixed reference): the lane keeps the instance in REGISTERS from the first
// step to the last -- measurement <- Adyn x0 + Bdyn u(:,0) (stage 1 of the solve's own forward sweep), y = g = 0, warm d / v / z.
// At this shape a closed-loop step takes 1-4 iterations, so the per-step path is bound by moving 1.1 KB of state through HBM for
// ~3,500 FLOP; here that traffic is paid once per roll_steps steps.
// The reference leaves v, z one iteration behind after an early exit (admm.cpp:135-138: the exit comes before v = vnew, z = znew).
// The kernel keeps BOTH generations on chip -- the forward sweep writes vn, zn (registers); v (shared memory, one 16-byte row per
// stage and lane) and z advance only when the iteration does not end the step by convergence -- so the hand-over to the next step is
// the reference's state with no special case.
// The sweeps are fully unrolled (the state is indexed statically), so the loop body is ~40 KB of code that every warp streams once
// per trip: code size is a first-order cost here.  The forward sweep therefore carries NO output code; the last step's x, u are
// written by a separate emission pass (the forward pass alone, evaluated again) in the trip that ends the rollout.
// PARITY order, warm buffers required (they receive the workspace the loop leaves: d, v, z, y, g of the last step).
template <int NX, int NH, int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1)
admm_kernel_small_roll(const __grid_constant__ Model<float, NX, 1, NH> P, const __grid_constant__ SolveArgs<float> a)
{
    static_assert(NX % 4 == 0, "16-byte rows of x");
    constexpr bool FAST = false;
    using O = Orders<float, NX, 1>;
    static_assert(O::Ax == ORD_SEQ && O::Mp == ORD_SEQ && O::Bu == ORD_SEQ && O::Qs == ORD_SEQ, "row sweeps are sequential at nu = 1");
    constexpr int H = NX / 2;
    constexpr unsigned FULLM = 0xffffffffu;
    constexpr int XROW = NX * NH, UROW = NH - 1;
    const float2 Z = make_float2(-0.f, -0.f);
    const unsigned lane = threadIdx.x & 31;
    const int S = a.roll_steps;

    extern __shared__ __align__(16) unsigned char smem[];
    const SVec<float, NX, NH, BLOCK> sv(smem, threadIdx.x);   // v as the reference holds it (one iteration behind vn until it advances)
    float2 g[NH][H], vn[NH][H], pn[H];
    float d[NH - 1], y[NH - 1], z[NH - 1], zn[NH - 1];
    float x0[NX], x1[NX];
    float u0r = 0.f;
#pragma unroll
    for (int i = 0; i < NH; ++i)
#pragma unroll
        for (int j = 0; j < H; ++j) g[i][j] = vn[i][j] = f2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NH - 1; ++i) d[i] = y[i] = z[i] = zn[i] = 0.f;
#pragma unroll
    for (int j = 0; j < H; ++j) pn[j] = f2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < NX; ++j) x0[j] = x1[j] = 0.f;

    long long inst = -1;
    int it = 0, rs = 0;
    // tracking a reference table: the window of MPC step `step` of this lane's instance (SolveArgs::roll_table)
    const bool tab = a.roll_table != nullptr;
    auto window = [&](int step) -> const float * {
        long long w0 = (long long)(a.roll_start ? __ldg(a.roll_start + (inst < 0 ? 0 : inst)) : 0) + a.roll_step0 + step;
        if (w0 > a.roll_rows - NH) w0 = a.roll_rows - NH;
        return a.roll_table + w0 * NX;
    };
    auto seed_pn = [&](const float *xl) {   // p_N seed: -(Xref_{N-1}^T Pinf)  (admm.cpp:83)
        float xr[NX];
        gload<float, NX>(xl, xr);
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            const float t = -dot<float, O::XtP, NX, FAST>([&](int k) { return P.Pf[k + j * NX]; }, [&](int k) { return xr[k]; });
            if (j & 1) pn[j / 2].y = t; else pn[j / 2].x = t;
        }
    };
    int phase = PH_FREE;
    bool exhausted = false;
    float res[4] = {0.f, 0.f, 0.f, 0.f};
    unsigned n_iter = 0, n_solved = 0, n_trips = 0, n_inst = 0;

    for (;;) {
        // ------------------------------------------------------------------ lane refill
        // Refills are WARP-SYNCHRONOUS: the warp claims 32 instances when its last lane has finished.  A lane idles for the trip or two
        // by which its rollout was shorter than the longest of the warp, but the refill's memory latency (claim, 250 B of rows per lane,
        // p_N seed) is paid once per rollout instead of in nearly every trip, and the lanes' last steps -- the only trips that run the
        // emission pass -- coincide.  (test_flags bit 4: every lane refills on its own, for the A/B: 29.1 vs 16.7 ms)
        const bool need = (phase == PH_FREE) && !exhausted;
        unsigned m = __ballot_sync(FULLM, need);
        if (!(a.test_flags & 16) && __any_sync(FULLM, phase != PH_FREE)) m = 0;
        if (m) {
            const int leader = __ffs(m) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(a.counter, (unsigned long long)__popc(m));
            base = __shfl_sync(FULLM, base, leader);
            if (need) {
                const long long idx = (long long)base + __popc(m & ((1u << lane) - 1u));
                const long long ci = idx < a.batch ? claim_instance(a, idx) : -1;
                if (ci >= 0) {
                    inst = ci; phase = PH_RUN; it = 0; rs = 0;
                    res[0] = res[1] = res[2] = res[3] = 0.f;
                    gload<float, NX>(a.x0 + inst * NX, x0);
                    seed_pn((tab ? window(0) : a.Xref + inst * a.xref_stride) + (NH - 1) * NX);
                    // warm d, v, z of the caller's workspace; the duals start from zero (the loop resets them every step)
#pragma unroll
                    for (int i = 0; i < NH - 1; ++i) {
                        d[i] = __ldg(a.wd + inst * UROW + i);
                        z[i] = __ldg(a.wz + inst * UROW + i);
                        y[i] = 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < NH; ++i) {
                        float tv[NX];
                        gload<float, NX>(a.wv + inst * XROW + i * NX, tv);
                        sv.store(i, tv);
#pragma unroll
                        for (int j = 0; j < H; ++j) g[i][j] = f2(0.f, 0.f);
                    }
                } else {
                    exhausted = true;
                }
            }
        }
        if (__all_sync(FULLM, phase == PH_FREE)) break;
        ++n_trips;

        if (phase == PH_RUN) ++it;
        const bool last = rs >= S - 1;   // the step whose outputs and workspace the caller gets

        // ------------------------------------------------------------------ forward sweep
        // forward_pass (admm.cpp:27-37) + update_slack (:45-61) + update_dual (:67-71) + residual maxima (:95-98).  No output code in
        // here: the trajectories of the last step are written by the emission pass below, in the trip that ends it
        float pri_x = 0.f, dua_x = 0.f, pri_u = 0.f, dua_u = 0.f;
        {
            float x[NX];
#pragma unroll
            for (int j = 0; j < NX; ++j) x[j] = x0[j];
#pragma unroll
            for (int i = 0; i < NH; ++i) {
                float2 xn[H];
                if (i < NH - 1) {
                    // u_i = -(Kinf x_i) - d_i                                                        :31
                    const float kx = dot<float, O::Kx, NX, FAST>([&](int k) { return P.K[k]; }, [&](int k) { return x[k]; });
                    const float u = __fsub_rn(-kx, d[i]);
                    float t = __fadd_rn(u, y[i]);                                                    // :47
                    t = fminf(P.umax[i], fmaxf(P.umin[i], t));                                       // :53
                    pri_u = fmaxf(pri_u, fabsf(__fsub_rn(u, t)));                                    // :97
                    dua_u = fmaxf(dua_u, fabsf(__fsub_rn(z[i], t)));                                 // :98
                    y[i] = __fsub_rn(__fadd_rn(y[i], u), t);                                         // :69
                    zn[i] = t;
                    if (i == 0) u0r = u;
                    // x_{i+1} = A x_i + B u_i                                                        :35
                    float2 ax[H];
                    matvec2<ORD_SEQ, NX, NX, NX, 0, FAST>(P.A, x, ax, Z);
#pragma unroll
                    for (int j = 0; j < H; ++j) {
                        const float2 b2 = f2(P.B[2 * j], P.B[2 * j + 1]);
                        xn[j] = add2(ax[j], f2(__fmul_rn(b2.x, u), __fmul_rn(b2.y, u)));
                    }
                    if (i == 0) {   // = the plant's next state (codegen_cartpole.cpp:117: the same expression on the same operands)
#pragma unroll
                        for (int j = 0; j < H; ++j) { x1[2 * j] = xn[j].x; x1[2 * j + 1] = xn[j].y; }
                    }
                }
                // state slack / dual / residuals of stage i
                float vo[NX];
                sv.load(i, vo);
#pragma unroll
                for (int j = 0; j < H; ++j) {
                    const float2 x2 = f2(x[2 * j], x[2 * j + 1]);
                    float2 t = add2(x2, g[i][j]);                                                    // :48
                    t.x = fminf(P.xmax[i * NX + 2 * j], fmaxf(P.xmin[i * NX + 2 * j], t.x));         // :59
                    t.y = fminf(P.xmax[i * NX + 2 * j + 1], fmaxf(P.xmin[i * NX + 2 * j + 1], t.y));
                    const float2 rp = sub2(x2, t), rd = sub2(f2(vo[2 * j], vo[2 * j + 1]), t);
                    pri_x = fmaxf(pri_x, fmaxf(fabsf(rp.x), fabsf(rp.y)));                           // :95
                    dua_x = fmaxf(dua_x, fmaxf(fabsf(rd.x), fabsf(rd.y)));                           // :96
                    g[i][j] = sub2(add2(g[i][j], x2), t);                                            // :70
                    vn[i][j] = t;
                }
                if (i < NH - 1) {
#pragma unroll
                    for (int j = 0; j < H; ++j) { x[2 * j] = xn[j].x; x[2 * j + 1] = xn[j].y; }
                }
            }
        }

        // ------------------------------------------------------------------ termination (admm.cpp:91-109, :135-138)
        bool final_bwd = false;   // last step, max_iter exit: this trip's backward sweep still runs and writes the workspace
        bool step_end = false;    // a step before the last ended in this trip: the next one starts after the backward section
        bool finished = false;
        bool fin = false;         // the last step of this lane's instance ended in this trip
        bool advance = false;     // v = vnew, z = znew (admm.cpp:141-142): every iteration that does not end its step by convergence
        if (phase == PH_RUN) {
            const bool chk = (it % P.check_term) == 0;
            if (chk) {
                res[0] = pri_x; res[1] = __fmul_rn(dua_x, P.rho); res[2] = pri_u; res[3] = __fmul_rn(dua_u, P.rho);
            }
            const bool conv = chk && res[0] < P.pri_tol && res[2] < P.pri_tol && res[1] < P.dua_tol && res[3] < P.dua_tol;
            advance = !conv;
            if (conv || it >= P.max_iter) {
                n_iter += (unsigned)it; n_solved += conv ? 1u : 0u; ++n_inst;
                if (!last) {
                    const long long h = (long long)rs * a.batch + inst;
                    if (a.roll_iter) a.roll_iter[h] = it;
                    if (a.roll_status) a.roll_status[h] = conv ? 1 : 11;
                    if (a.roll_u0) a.roll_u0[h] = u0r;
                    if (a.roll_x) gstore<float, NX>(a.roll_x + h * NX, x1);
                    step_end = true;
                } else {
                    if (a.iter) a.iter[inst] = it;
                    if (a.status) a.status[inst] = conv ? 1 : 11;
                    if (a.resid) *reinterpret_cast<float4 *>(a.resid + inst * 4) = make_float4(res[0], res[1], res[2], res[3]);
                    final_bwd = !conv;
                    fin = true;
                    if (conv) {
                        // the workspace an early exit leaves: d, v, z as they entered this iteration (y, g go out with the trajectories)
#pragma unroll
                        for (int i = 0; i < NH - 1; ++i) { a.wd[inst * UROW + i] = d[i]; a.wz[inst * UROW + i] = z[i]; }
#pragma unroll
                        for (int i = 0; i < NH; ++i) {
                            float t[NX];
                            sv.load(i, t);
                            gstore<float, NX>(a.wv + inst * XROW + i * NX, t);
                        }
                    }
                    phase = PH_FREE; finished = true;
                }
            }
        }
        // ------------------------------------------------------------------ emission: the rollout of this lane ends in this trip
        // x, u of the solve = its last forward pass (admm.cpp:27-37 with the d that pass used: BEFORE the backward sweep of a max_iter
        // exit), evaluated again without the slack / dual work; y, g as the terminating iteration left them
        if (__any_sync(FULLM, fin)) {
            if (fin) {
                float *xo = a.x ? a.x + inst * XROW : nullptr;
                float *uo = a.u ? a.u + inst * UROW : nullptr;
                float x[NX];
#pragma unroll
                for (int j = 0; j < NX; ++j) x[j] = x0[j];
#pragma unroll
                for (int i = 0; i < NH; ++i) {
                    if (xo) gstore<float, NX>(xo + i * NX, x);
                    {
                        float t[NX];
#pragma unroll
                        for (int j = 0; j < H; ++j) { t[2 * j] = g[i][j].x; t[2 * j + 1] = g[i][j].y; }
                        gstore<float, NX>(a.wg + inst * XROW + i * NX, t);
                    }
                    if (i < NH - 1) {
                        const float kx = dot<float, O::Kx, NX, FAST>([&](int k) { return P.K[k]; }, [&](int k) { return x[k]; });
                        const float u = __fsub_rn(-kx, d[i]);
                        if (uo) uo[i] = u;
                        if (i == 0 && a.u0) a.u0[inst] = u;
                        a.wy[inst * UROW + i] = y[i];
                        float2 ax[H];
                        matvec2<ORD_SEQ, NX, NX, NX, 0, FAST>(P.A, x, ax, Z);
#pragma unroll
                        for (int j = 0; j < H; ++j) {
                            const float2 b2 = f2(P.B[2 * j], P.B[2 * j + 1]);
                            const float2 xn = add2(ax[j], f2(__fmul_rn(b2.x, u), __fmul_rn(b2.y, u)));
                            x[2 * j] = xn.x; x[2 * j + 1] = xn.y;
                        }
                    }
                }
            }
        }
        if (advance) {
#pragma unroll
            for (int i = 0; i < NH - 1; ++i) z[i] = zn[i];
#pragma unroll
            for (int i = 0; i < NH; ++i) {
                float t[NX];
#pragma unroll
                for (int j = 0; j < H; ++j) { t[2 * j] = vn[i][j].x; t[2 * j + 1] = vn[i][j].y; }
                sv.store(i, t);
            }
        }

        // ------------------------------------------------------------------ backward sweep
        // update_linear_cost (admm.cpp:77-85) recomputed per stage + backward_pass_grad (:15-22)
        const bool cont = advance;                 // (a running lane whose iteration did not converge: max_iter exits included)
        const bool wout = final_bwd;
        if (__any_sync(FULLM, cont)) {
            float p[NX];
            const float *xr_base = tab ? window(rs) : a.Xref + (inst < 0 ? 0 : inst) * a.xref_stride;
            float *wdo = wout ? a.wd + inst * UROW : nullptr;
            float *wvo = wout ? a.wv + inst * XROW : nullptr;
            float *wzo = wout ? a.wz + inst * UROW : nullptr;
            auto store_v = [&](int i) {
                float t[NX];
#pragma unroll
                for (int j = 0; j < H; ++j) { t[2 * j] = vn[i][j].x; t[2 * j + 1] = vn[i][j].y; }   // (= v: the lane has advanced)
                gstore<float, NX>(wvo + i * NX, t);
            };
            if (wvo) store_v(NH - 1);
#pragma unroll
            for (int j = 0; j < H; ++j) {
                const float2 dvg = sub2(vn[NH - 1][j], g[NH - 1][j]);
                const float2 t = sub2(pn[j], f2(__fmul_rn(P.rho, dvg.x), __fmul_rn(P.rho, dvg.y)));  // :84
                p[2 * j] = t.x; p[2 * j + 1] = t.y;
            }
#pragma unroll
            for (int i = NH - 2; i >= 0; --i) {
                float xr[NX];
                gload<float, NX>(xr_base + i * NX, xr);
                if (wvo) { store_v(i); wzo[i] = z[i]; }
                const float r = __fmul_rn(P.nrho, __fsub_rn(z[i], y[i]));                            // :80
                // d_i = Quu_inv (B^T p_{i+1} + r_i)                                                  :19
                const float bp = dot<float, O::Btp, NX, FAST>([&](int k) { return P.B[k]; }, [&](int k) { return p[k]; });
                const float dn = __fmul_rn(P.Qi[0], __fadd_rn(bp, r));
                if (cont) d[i] = dn;
                if (wdo) wdo[i] = dn;
                // p_i = q_i + AmBKt p_{i+1} - Kinf^T r_i                                             :20
                float2 mp[H];
                matvec2<ORD_SEQ, NX, NX, NX, 0, FAST>(P.M, p, mp, Z);
#pragma unroll
                for (int j = 0; j < H; ++j) {
                    const float2 cq = neg2(f2(__fmul_rn(xr[2 * j], P.Qd[2 * j]), __fmul_rn(xr[2 * j + 1], P.Qd[2 * j + 1])));   // :81
                    const float2 dvg = sub2(vn[i][j], g[i][j]);
                    const float2 q = sub2(cq, f2(__fmul_rn(P.rho, dvg.x), __fmul_rn(P.rho, dvg.y)));  // :82
                    const float2 kr = f2(__fmul_rn(P.K[2 * j], r), __fmul_rn(P.K[2 * j + 1], r));
                    const float2 t = sub2(add2(q, mp[j]), kr);
                    p[2 * j] = t.x; p[2 * j + 1] = t.y;
                }
            }
        }
        // ------------------------------------------------------------------ next MPC step on this lane
        if (step_end) {
#pragma unroll
            for (int j = 0; j < NX; ++j) x0[j] = x1[j];                                              // codegen_cartpole.cpp:117
#pragma unroll
            for (int i = 0; i < NH - 1; ++i) y[i] = 0.f;                                             // :100
#pragma unroll
            for (int i = 0; i < NH; ++i)
#pragma unroll
                for (int j = 0; j < H; ++j) g[i][j] = f2(0.f, 0.f);                                  // :101
            ++rs; it = 0;
            if (tab) seed_pn(window(rs) + (NH - 1) * NX);   // the window moves on
            res[0] = res[1] = res[2] = res[3] = 0.f;
            // the measurement the LAST step starts from: the caller's plant step reads it after the launch
            if (rs >= S - 1) gstore<float, NX>(const_cast<float *>(a.x0) + inst * NX, x0);
        }
        if (finished && a.done) { __threadfence(); atomicAdd(a.done + (inst >> a.done_shift), 1u); }
    }

    if (a.stats) {
        unsigned long long t_iter = n_iter, t_solved = n_solved, t_trips = n_trips, t_inst = n_inst;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            t_iter += __shfl_down_sync(FULLM, t_iter, o);
            t_solved += __shfl_down_sync(FULLM, t_solved, o);
            t_trips += __shfl_down_sync(FULLM, t_trips, o);
            t_inst += __shfl_down_sync(FULLM, t_inst, o);
        }
        if (lane == 0) {
            atomicAdd(a.stats + 0, t_iter);
            atomicAdd(a.stats + 1, t_solved);
            atomicAdd(a.stats + 2, t_trips);
            atomicAdd(a.stats + 3, t_inst);
        }
    }
}

}  // namespace tmpc
