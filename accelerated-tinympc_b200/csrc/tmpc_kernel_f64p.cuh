// fp64 12/4/N (the reference's shipped tinytype), one shared model: TWO LANES PER INSTANCE.
//
// The thread-per-instance double kernel (tmpc_kernel.cuh, SYS == 3) holds 128 instances per SM -- g and v of an instance are
// 480 tensor-memory cells -- with 128 threads = one warp per scheduler, 255 registers per thread, and its coefficients come
// through the uniform path (a double cannot be a constant-bank operand of DMUL: LDCU.64 + two moves per coefficient, on whose
// scoreboard DMUL waits; profiles/r02_ncu_fp64.md: issue slots 25 % busy, FP64 pipe 22 %).  Here, as in the per-instance-systems
// kernel (tmpc_kernel_sysp.cuh), the same 128 instances per SM are worked by 256 threads: lanes (2t, 2t+1) of a warp share an
// instance, each owning half of the output rows of every product (x rows 6h..6h+5, u rows 2h..2h+1, h = lane & 1):
//  * two warps per scheduler, half the registers per thread;
//  * the model lives in SHARED memory in the major in which a lane's own rows are adjacent (Kinf, Adyn, Bdyn, Quu_inv, AmBKt
//    column-major; Bdyn and Kinf row-major for the transposed products): one LDS.128 = the coefficients of two rows, straight
//    into the registers DMUL reads;
//  * own rows of g, v in tensor memory (warps w and w+4 share TMEM lanes: 256 columns per thread, 240 used), of d, y, z and the
//    p_N seed in shared memory;
//  * the halves of x (p), u, s, r are exchanged with butterfly shuffles; the residual maxima are combined across the pair.
// Every product keeps the summation order of the reference's double build (Orders<double, 12, 4>): bit-identical results.
#pragma once
#include "tmpc_kernel.cuh"

namespace tmpc {
namespace f64p {

__device__ __forceinline__ double2 ldsd2(uint32_t a)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void stsd2(uint32_t a, double2 v) { asm volatile("st.shared.v2.f64 [%0], {%1,%2};" :: "r"(a), "d"(v.x), "d"(v.y) : "memory"); }
// model coefficients: read-only after the prologue (no "memory" clobber: arithmetic and other loads may move around them;
// volatile: never hoisted out of the sweeps or above the prologue's barrier)
__device__ __forceinline__ double2 ldc2(uint32_t a)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}

// There is no double-precision min / max instruction: fmin / fmax become a compare plus half a dozen moves and selects each, and a
// forward stage has 32 of them (a third of its instructions).  For finite data (the only data parity is claimed for) a compare
// and a select give the same VALUE (only the sign of a zero result can differ, which no later value depends on).
__device__ __forceinline__ double maxabs(double m, double t) { const double at = fabs(t); return at > m ? at : m; }   // m >= 0
__device__ __forceinline__ double clampd(double v, double lo, double hi)
{
    const double r = lo > v ? lo : v;
    return hi < r ? hi : r;
}

// model image in shared memory (doubles)
struct MdlMap {
    static constexpr int K = 0, A = K + 48, B = A + 144, BR = B + 48, QI = BR + 48, M = QI + 16, KR = M + 144, LEN = KR + 48;
};

template <int NH> struct Smem {
    static constexpr int BLOCK = 256;
    static constexpr uint32_t DYZ = 0, PN = DYZ + 3u * (NH - 1) * BLOCK * 16, MDL = PN + 3u * BLOCK * 16, TMSLOT = MDL + MdlMap::LEN * 8,
                              BYTES = TMSLOT + 16;
};

}  // namespace f64p

#ifndef TMPC_SPEC_FACTOR_F64
#define TMPC_SPEC_FACTOR_F64 2.0
#endif

template <int NH, bool FAST, bool WARM>
__global__ void __launch_bounds__(256, 1)
admm_kernel_f64p(const __grid_constant__ Model<double, 12, 4, NH> P, const __grid_constant__ SolveArgs<double> a)
{
    using namespace f64p;
    using N = Num<double>;
    constexpr int NX = 12, NU = 4, OX = 6, OU = 2;   // own rows
    using O = Orders<double, NX, NU>;
    using SS = Smem<NH>;
    using MM = MdlMap;
    static_assert(O::Kx == ORD_SEQ && O::Ax == ORD_SEQ && O::Bu == ORD_SEQ, "the forward sweep spells out sequential sums");
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31;
    const int warp = tid >> 5;
    const int h = lane & 1;
    const bool odd = h != 0;
    const unsigned pe = lane & ~1u;
    constexpr unsigned FULLM = 0xffffffffu, EVEN = 0x55555555u;
    constexpr int XROW = NX * NH, UROW = NU * (NH - 1);

    uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
    uint32_t s_dyz = sb + SS::DYZ + tid * 16, s_pn = sb + SS::PN + tid * 16, s_mdl = sb + SS::MDL;
    asm volatile("" : "+r"(s_dyz), "+r"(s_pn), "+r"(s_mdl));   // keep the bases in registers (see tmpc_kernel_sysp.cuh)
    // ---- model image: own-row-adjacent majors
    {
        double *md = reinterpret_cast<double *>(smem + SS::MDL);
        for (int e = tid; e < MM::LEN; e += 256) {
            double v;
            if (e < MM::A) v = P.K[e];
            else if (e < MM::B) v = P.A[e - MM::A];
            else if (e < MM::BR) v = P.B[e - MM::B];
            else if (e < MM::QI) { const int q = e - MM::BR; v = P.B[(q >> 2) + NX * (q & 3)]; }        // Brm[k*4 + r] = Bdyn(k, r)
            else if (e < MM::M) v = P.Qi[e - MM::QI];
            else if (e < MM::KR) v = P.M[e - MM::M];
            else { const int q = e - MM::KR; v = P.K[(q / NX) + NU * (q % NX)]; }                       // Krm[k*12 + r] = Kinf(k, r)
            md[e] = v;
        }
    }
    uint32_t tcol;
    {
        uint32_t *slot = reinterpret_cast<uint32_t *>(smem + SS::TMSLOT);
        if (tid < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();   // (also publishes the model image)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tcol = *slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
    }
    TVecD<12, NH> sgv;   // own rows of g (0..5) and v (6..11) per stage: 24 of the thread's 256 tensor-memory columns
    sgv.base = tcol;

    // coefficient pair of the lane's rows (r0, r0 + 1) at element offset `off` of the image (for h = 0) + hs * h
    auto cf = [&](int off, int hs) -> double2 { return ldc2(s_mdl + (uint32_t)((off + hs * h) * 8)); };
    // whole vectors from the pair's halves (even lane: low rows, odd lane: high rows)
    auto px = [&](double v) -> double { return __shfl_xor_sync(FULLM, v, 1); };
    auto gather12 = [&](const double (&own)[OX], double (&full)[NX]) {
#pragma unroll
        for (int t = 0; t < OX; ++t) {
            const double o = px(own[t]);
            full[t] = odd ? o : own[t];
            full[OX + t] = odd ? own[t] : o;
        }
    };
    auto gather4 = [&](const double (&own)[OU], double (&full)[NU]) {
#pragma unroll
        for (int t = 0; t < OU; ++t) {
            const double o = px(own[t]);
            full[t] = odd ? o : own[t];
            full[OU + t] = odd ? own[t] : o;
        }
    };
    auto lds_dyz = [&](int i, double (&d)[OU], double (&y)[OU], double (&z)[OU]) {
        const uint32_t b = s_dyz + (uint32_t)(i * 3) * (256 * 16);
        const double2 t0 = ldsd2(b), t1 = ldsd2(b + 256 * 16), t2 = ldsd2(b + 2 * 256 * 16);
        d[0] = t0.x; d[1] = t0.y; y[0] = t1.x; y[1] = t1.y; z[0] = t2.x; z[1] = t2.y;
    };
    auto sts_chunk = [&](int i, int c, const double (&v)[OU]) { stsd2(s_dyz + (uint32_t)(i * 3 + c) * (256 * 16), make_double2(v[0], v[1])); };

    double x0o[OX];
#pragma unroll
    for (int j = 0; j < OX; ++j) x0o[j] = 0.0;
    long long inst = -1;
    int it = 0;
    int phase = PH_FREE;
    bool exhausted = false;
    int deferred = 0;
    bool spec = false, counted = true;
    double res[4] = {0.0, 0.0, 0.0, 0.0};
    unsigned long long n_iter = 0, n_solved = 0, n_trips = 0, n_inst = 0;
    const bool seed_shared = a.xref_stride == 0;
    if (seed_shared) {   // own rows of the p_N seed -(Xref_{N-1}^T Pinf) (admm.cpp:83): one reference for the whole batch
        double xr[NX];
        gload<double, NX>(a.Xref + (NH - 1) * NX, xr);
#pragma unroll 1
        for (int j = 0; j < OX; j += 2) {
            double pv[2];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int col = OX * h + j + jj;
                pv[jj] = -dot<double, O::XtP, NX, FAST>([&](int k) { return P.Pf[k + col * NX]; }, [&](int k) { return xr[k]; });
            }
            stsd2(s_pn + (uint32_t)(j >> 1) * (256 * 16), make_double2(pv[0], pv[1]));
        }
    }

    for (;;) {
        // ------------------------------------------------------------------ refill (per PAIR)
        const bool need = (phase == PH_FREE) && !exhausted;
        unsigned m = __ballot_sync(FULLM, need) & EVEN;
        {
            const bool others_busy = __ballot_sync(FULLM, phase != PH_FREE) != 0;
            if (m && __popc(m) < 2 && deferred < 1 && others_busy) { ++deferred; m = 0; }
            else deferred = 0;
        }
        if (m) {
            const int leader = __ffs(m) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(a.counter, (unsigned long long)__popc(m));
            base = __shfl_sync(FULLM, base, leader);
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
                    res[0] = res[1] = res[2] = res[3] = 0.0;
                    {
                        const double2 *xp = reinterpret_cast<const double2 *>(a.x0 + inst * NX + OX * h);
#pragma unroll
                        for (int j = 0; j < OX / 2; ++j) { const double2 t = __ldg(xp + j); x0o[2 * j] = t.x; x0o[2 * j + 1] = t.y; }
                    }
                    if (!seed_shared) {
                        double xr[NX];
                        gload<double, NX>(a.Xref + inst * a.xref_stride + (NH - 1) * NX, xr);
#pragma unroll 1
                        for (int j = 0; j < OX; j += 2) {
                            double pv[2];
#pragma unroll
                            for (int jj = 0; jj < 2; ++jj) {
                                const int col = OX * h + j + jj;
                                pv[jj] = -dot<double, O::XtP, NX, FAST>([&](int k) { return P.Pf[k + col * NX]; }, [&](int k) { return xr[k]; });
                            }
                            stsd2(s_pn + (uint32_t)(j >> 1) * (256 * 16), make_double2(pv[0], pv[1]));
                        }
                    }
                    if (WARM && a.wd) {
#pragma unroll 1
                        for (int i = 0; i < NH - 1; ++i) {
                            const double2 td = __ldg(reinterpret_cast<const double2 *>(a.wd + inst * UROW + i * NU + OU * h));
                            const double2 ty = __ldg(reinterpret_cast<const double2 *>(a.wy + inst * UROW + i * NU + OU * h));
                            const double2 tz = __ldg(reinterpret_cast<const double2 *>(a.wz + inst * UROW + i * NU + OU * h));
                            const uint32_t b = s_dyz + (uint32_t)(i * 3) * (256 * 16);
                            stsd2(b, td); stsd2(b + 256 * 16, ty); stsd2(b + 2 * 256 * 16, tz);
                        }
                    } else {
                        const double2 zz = make_double2(0.0, 0.0);
#pragma unroll 1
                        for (int c = 0; c < 3 * (NH - 1); ++c) stsd2(s_dyz + (uint32_t)c * (256 * 16), zz);
                    }
                } else {
                    exhausted = true;
                }
            }
            // tcgen05 is warp-collective: every lane rewrites its g / v cells, refilled lanes with zeros (cold) or the caller's warm
            // state, the others with what they hold
            const bool wfill = fill && WARM && a.wd;
#pragma unroll 1
            for (int i = 0; i < NH; ++i) {
                double gv[12];
                sgv.load(i, gv);
                if (wfill) {
                    const double2 *gp = reinterpret_cast<const double2 *>(a.wg + inst * XROW + i * NX + OX * h);
                    const double2 *vp = reinterpret_cast<const double2 *>(a.wv + inst * XROW + i * NX + OX * h);
#pragma unroll
                    for (int j = 0; j < OX / 2; ++j) {
                        const double2 tg = __ldg(gp + j), tv = __ldg(vp + j);
                        gv[2 * j] = tg.x; gv[2 * j + 1] = tg.y; gv[OX + 2 * j] = tv.x; gv[OX + 2 * j + 1] = tv.y;
                    }
                } else if (fill) {
#pragma unroll
                    for (int j = 0; j < 12; ++j) gv[j] = 0.0;
                }
                sgv.store(i, gv);
            }
            tm_wait_st();
        }
        if (__all_sync(FULLM, phase == PH_FREE)) break;
        ++n_trips;

        const bool emit = (phase == PH_EMIT);
        if (phase == PH_RUN) ++it;

        // ------------------------------------------------------------------ forward sweep
        // forward_pass (admm.cpp:27-37) + update_slack (:45-61) + update_dual (:67-71) + residual maxima (:95-98)
        double pri_x = 0.0, dua_x = 0.0, pri_u = 0.0, dua_u = 0.0;
        {
            double xo_[OX], xs[NX];
#pragma unroll
            for (int j = 0; j < OX; ++j) xo_[j] = x0o[j];
            gather12(xo_, xs);
            const bool wr = emit || (spec && phase == PH_RUN);
            double *xo = (wr && a.x) ? a.x + inst * XROW + OX * h : nullptr;
            double *uo = (wr && a.u) ? a.u + inst * UROW + OU * h : nullptr;
            double *u0o = (wr && a.u0) ? a.u0 + inst * NU + OU * h : nullptr;
            double *go = (WARM && emit && a.wg) ? a.wg + inst * XROW + OX * h : nullptr;
            double *yo = (WARM && emit && a.wy) ? a.wy + inst * UROW + OU * h : nullptr;
            auto state_part = [&](int i) {   // own rows   (:48, :59, :70, :95, :96)
                double gv[12];
                sgv.load(i, gv);
                if (WARM && go) {
#pragma unroll
                    for (int j = 0; j < OX / 2; ++j) reinterpret_cast<double2 *>(go + i * NX)[j] = make_double2(gv[2 * j], gv[2 * j + 1]);
                }
                const int bi = i * NX + OX * h;
#pragma unroll
                for (int j = 0; j < OX; ++j) {
                    const double xg = N::add(xo_[j], gv[j]);
                    const double vn = clampd(xg, P.xmin[bi + j], P.xmax[bi + j]);
                    pri_x = maxabs(pri_x, N::sub(xo_[j], vn));
                    dua_x = maxabs(dua_x, N::sub(gv[OX + j], vn));
                    gv[j] = N::sub(xg, vn);   // (g + x) - vnew: the sum is the one above (addition commutes bit for bit)
                    gv[OX + j] = vn;
                }
                sgv.store(i, gv);
                if (xo) {
#pragma unroll
                    for (int j = 0; j < OX / 2; ++j) reinterpret_cast<double2 *>(xo + i * NX)[j] = make_double2(xo_[2 * j], xo_[2 * j + 1]);
                }
            };
#pragma unroll 1
            for (int i = 0; i < NH - 1; ++i) {
                state_part(i);
                double d[OU], y[OU], z[OU], uo_[OU], us[NU], kx[OU], ax[OX], bu[OX];
                lds_dyz(i, d, y, z);
                if (WARM && yo) *reinterpret_cast<double2 *>(yo + i * NU) = make_double2(y[0], y[1]);
                // Kinf x (own 2 rows) and Adyn x (own 6 rows): sequential sums over the columns, all chains advancing together   (:31, :35)
#pragma unroll
                for (int k = 0; k < NX; ++k) {
                    const double2 ck = cf(MM::K + NU * k, OU);
                    const double2 c0 = cf(MM::A + NX * k, OX), c1 = cf(MM::A + NX * k + 2, OX), c2 = cf(MM::A + NX * k + 4, OX);
                    const double cc[OX] = {c0.x, c0.y, c1.x, c1.y, c2.x, c2.y};
                    if (k == 0) {
                        kx[0] = N::mul(ck.x, xs[0]); kx[1] = N::mul(ck.y, xs[0]);
#pragma unroll
                        for (int r = 0; r < OX; ++r) ax[r] = N::mul(cc[r], xs[0]);
                    } else {
                        kx[0] = FAST ? N::fma(ck.x, xs[k], kx[0]) : N::add(N::mul(ck.x, xs[k]), kx[0]);
                        kx[1] = FAST ? N::fma(ck.y, xs[k], kx[1]) : N::add(N::mul(ck.y, xs[k]), kx[1]);
#pragma unroll
                        for (int r = 0; r < OX; ++r) ax[r] = FAST ? N::fma(cc[r], xs[k], ax[r]) : N::add(N::mul(cc[r], xs[k]), ax[r]);
                    }
                }
                {
                    const int bi = i * NU + OU * h;
                    double yn[OU], zn[OU];
#pragma unroll
                    for (int r = 0; r < OU; ++r) {
                        uo_[r] = N::sub(-kx[r], d[r]);                                            // :31
                        const double uy = N::add(uo_[r], y[r]);                                   // :47
                        zn[r] = clampd(uy, P.umin[bi + r], P.umax[bi + r]);                       // :53
                        pri_u = maxabs(pri_u, N::sub(uo_[r], zn[r]));                             // :97
                        dua_u = maxabs(dua_u, N::sub(z[r], zn[r]));                               // :98
                        yn[r] = N::sub(uy, zn[r]);                                                // :69  (y + u) - znew
                    }
                    sts_chunk(i, 1, yn);
                    sts_chunk(i, 2, zn);
                    if (uo) *reinterpret_cast<double2 *>(uo + i * NU) = make_double2(uo_[0], uo_[1]);
                    if (u0o && i == 0) *reinterpret_cast<double2 *>(u0o) = make_double2(uo_[0], uo_[1]);
                    gather4(uo_, us);
                }
                // Bdyn u (own 6 rows)   (:35)
#pragma unroll
                for (int k = 0; k < NU; ++k) {
                    const double2 c0 = cf(MM::B + NX * k, OX), c1 = cf(MM::B + NX * k + 2, OX), c2 = cf(MM::B + NX * k + 4, OX);
                    const double cc[OX] = {c0.x, c0.y, c1.x, c1.y, c2.x, c2.y};
#pragma unroll
                    for (int r = 0; r < OX; ++r) {
                        if constexpr (FAST) ax[r] = N::fma(cc[r], us[k], ax[r]);
                        else if (k == 0) bu[r] = N::mul(cc[r], us[0]);
                        else bu[r] = N::add(N::mul(cc[r], us[k]), bu[r]);
                    }
                }
#pragma unroll
                for (int r = 0; r < OX; ++r) xo_[r] = FAST ? ax[r] : N::add(ax[r], bu[r]);         // :35
                gather12(xo_, xs);
            }
            state_part(NH - 1);
            tm_wait_st();   // the backward sweep reads the cells this sweep wrote
        }
        pri_x = maxabs(pri_x, __shfl_xor_sync(FULLM, pri_x, 1));
        dua_x = maxabs(dua_x, __shfl_xor_sync(FULLM, dua_x, 1));
        pri_u = maxabs(pri_u, __shfl_xor_sync(FULLM, pri_u, 1));
        dua_u = maxabs(dua_u, __shfl_xor_sync(FULLM, dua_u, 1));

        // ------------------------------------------------------------------ termination (admm.cpp:91-109, :135-138)
        bool final_bwd = false;
        if (phase == PH_RUN) {
            const bool chk = (it % P.check_term) == 0;
            if (chk) {
                res[0] = pri_x;
                res[1] = N::mul(dua_x, P.rho);
                res[2] = pri_u;
                res[3] = N::mul(dua_u, P.rho);
            }
            const bool conv = chk && res[0] < P.pri_tol && res[2] < P.pri_tol && res[1] < P.dua_tol && res[3] < P.dua_tol;
            if (conv || it >= P.max_iter) {
                if (h == 0) {
                    if (a.iter) a.iter[inst] = it;
                    if (a.status) a.status[inst] = conv ? 1 : 11;
                    if (a.resid) {
                        reinterpret_cast<double2 *>(a.resid + inst * 4)[0] = make_double2(res[0], res[1]);
                        reinterpret_cast<double2 *>(a.resid + inst * 4)[1] = make_double2(res[2], res[3]);
                    }
                    n_iter += (unsigned)it;
                    n_solved += conv ? 1u : 0u;
                    ++n_inst;
                }
                final_bwd = !conv;
                phase = spec ? PH_FREE : PH_EMIT;
            } else if constexpr (!WARM) {
                constexpr double SF = TMPC_SPEC_FACTOR_F64;
                const bool next_chk = ((it + 1) % P.check_term) == 0;
                spec = (it + 1 >= P.max_iter) ||
                       (next_chk && res[0] < SF * P.pri_tol && res[2] < SF * P.pri_tol && res[1] < SF * P.dua_tol && res[3] < SF * P.dua_tol);
            }
        } else if (phase == PH_EMIT) {
            phase = PH_FREE;
        }
        if (a.done) {
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
            double po[OX], ps[NX];
            const double *xr_base = a.Xref + (inst < 0 ? 0 : inst) * a.xref_stride + OX * h;
            double *wdo = wout ? a.wd + inst * UROW + OU * h : nullptr;
            double *wvo = wout ? a.wv + inst * XROW + OX * h : nullptr;
            double *wzo = wout ? a.wz + inst * UROW + OU * h : nullptr;
            const double rho_l = P.rho, nrho_l = P.nrho;
            {
                double gv[12], pn[OX];
                sgv.load(NH - 1, gv);
#pragma unroll
                for (int j = 0; j < OX / 2; ++j) { const double2 t = ldsd2(s_pn + (uint32_t)j * (256 * 16)); pn[2 * j] = t.x; pn[2 * j + 1] = t.y; }
                if (WARM && wvo) {
#pragma unroll
                    for (int j = 0; j < OX / 2; ++j) reinterpret_cast<double2 *>(wvo + (NH - 1) * NX)[j] = make_double2(gv[OX + 2 * j], gv[OX + 2 * j + 1]);
                }
#pragma unroll
                for (int j = 0; j < OX; ++j) {
                    if constexpr (FAST) po[j] = N::fma(nrho_l, N::sub(gv[OX + j], gv[j]), pn[j]);
                    else po[j] = N::sub(pn[j], N::mul(rho_l, N::sub(gv[OX + j], gv[j])));        // :84
                }
                gather12(po, ps);
            }
#pragma unroll 1
            for (int i = NH - 2; i >= 0; --i) {
                double d[OU], y[OU], z[OU], gv[12], r_[OU], rs[NU], q[OX], s_[OU], ss[NU], dn[OU];
                lds_dyz(i, d, y, z);
                sgv.load(i, gv);
                if (WARM && wvo) {
#pragma unroll
                    for (int j = 0; j < OX / 2; ++j) reinterpret_cast<double2 *>(wvo + i * NX)[j] = make_double2(gv[OX + 2 * j], gv[OX + 2 * j + 1]);
                    *reinterpret_cast<double2 *>(wzo + i * NU) = make_double2(z[0], z[1]);
                }
#pragma unroll
                for (int r = 0; r < OU; ++r) r_[r] = N::mul(nrho_l, N::sub(z[r], y[r]));          // :80
                gather4(r_, rs);
#pragma unroll
                for (int j = 0; j < OX / 2; ++j) {
                    const double2 xr = __ldg(reinterpret_cast<const double2 *>(xr_base + i * NX) + j);
                    const double xv[2] = {xr.x, xr.y};
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const int r = 2 * j + jj;
                        const double cq = -N::mul(xv[jj], P.Qd[OX * h + r]);                      // :81
                        if constexpr (FAST) q[r] = N::fma(nrho_l, N::sub(gv[OX + r], gv[r]), cq);
                        else q[r] = N::sub(cq, N::mul(rho_l, N::sub(gv[OX + r], gv[r])));         // :82
                    }
                }
                // Bdyn^T p, own 2 rows, in the reference's order   (:19)
                {
                    double e0[NX], e1[NX];
#pragma unroll
                    for (int k = 0; k < NX; ++k) {
                        const double2 c = cf(MM::BR + NU * k, OU);
                        e0[k] = c.x; e1[k] = c.y;
                    }
                    const double b0 = dot<double, O::Btp, NX, FAST>([&](int k) { return e0[k]; }, [&](int k) { return ps[k]; });
                    const double b1 = dot<double, O::Btp, NX, FAST>([&](int k) { return e1[k]; }, [&](int k) { return ps[k]; });
                    s_[0] = N::add(b0, r_[0]);
                    s_[1] = N::add(b1, r_[1]);
                }
                gather4(s_, ss);
                {   // d_i = Quu_inv s, own 2 rows   (:19)
                    double e0[NU], e1[NU];
#pragma unroll
                    for (int k = 0; k < NU; ++k) {
                        const double2 c = cf(MM::QI + NU * k, OU);
                        e0[k] = c.x; e1[k] = c.y;
                    }
                    dn[0] = dot<double, O::Qs, NU, FAST>([&](int k) { return e0[k]; }, [&](int k) { return ss[k]; });
                    dn[1] = dot<double, O::Qs, NU, FAST>([&](int k) { return e1[k]; }, [&](int k) { return ss[k]; });
                }
                if (cont) sts_chunk(i, 0, dn);
                if (WARM && wdo) *reinterpret_cast<double2 *>(wdo + i * NU) = make_double2(dn[0], dn[1]);
                // p_i = q + AmBKt p - Kinf^T r, own 6 rows, two rows at a time   (:20)
#pragma unroll
                for (int j = 0; j < OX; j += 2) {
                    double m0[NX], m1[NX], k0[NU], k1[NU];
#pragma unroll
                    for (int k = 0; k < NX; ++k) {
                        const double2 c = cf(MM::M + NX * k + j, OX);
                        m0[k] = c.x; m1[k] = c.y;
                    }
#pragma unroll
                    for (int k = 0; k < NU; ++k) {
                        const double2 c = cf(MM::KR + NX * k + j, OX);
                        k0[k] = c.x; k1[k] = c.y;
                    }
                    const double mp0 = dot<double, O::Mp, NX, FAST>([&](int k) { return m0[k]; }, [&](int k) { return ps[k]; });
                    const double mp1 = dot<double, O::Mp, NX, FAST>([&](int k) { return m1[k]; }, [&](int k) { return ps[k]; });
                    const double kr0 = dot<double, O::Ktr, NU, FAST>([&](int k) { return k0[k]; }, [&](int k) { return rs[k]; });
                    const double kr1 = dot<double, O::Ktr, NU, FAST>([&](int k) { return k1[k]; }, [&](int k) { return rs[k]; });
                    po[j] = N::sub(N::add(q[j], mp0), kr0);
                    po[j + 1] = N::sub(N::add(q[j + 1], mp1), kr1);
                }
                gather12(po, ps);
            }
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
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tcol) : "memory");   // (warp 0: tcol is the base)
}

}  // namespace tmpc
