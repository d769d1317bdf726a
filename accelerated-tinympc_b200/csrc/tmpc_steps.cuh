// The six step functions of the reference (admm.hpp:13-18) as stand-alone batched kernels on a full workspace in
// global memory ([instance][stage][dim]).  They exist for unit parity with the reference's exported step functions
// and for the host-side single-instance API; the production path is the fused persistent kernel.
#pragma once
#include "tmpc_kernel.cuh"

namespace tmpc {

template <class T> struct StepArgs {
    long long batch;
    T *x, *u, *q, *r, *p, *d, *v, *vnew, *z, *znew, *g, *y;
    const T *Xref;
    long long xref_stride;
    T *resid;   // [batch][4] pri_x, dua_x, pri_u, dua_u (in/out)
    int *term;  // [batch] termination_condition result
    int iter;
};

enum { STEP_FORWARD = 0, STEP_SLACK = 1, STEP_DUAL = 2, STEP_LINCOST = 3, STEP_TERM = 4, STEP_BACKWARD = 5 };

template <class T, int NX, int NU, int NH, bool FAST>
__global__ void step_kernel(const __grid_constant__ Model<T, NX, NU, NH> P, const __grid_constant__ StepArgs<T> a, int which)
{
    using N = Num<T>;
    using O = Orders<T, NX, NU>;
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= a.batch) return;
    constexpr int XROW = NX * NH, UROW = NU * (NH - 1);
    T *x = a.x + b * XROW, *u = a.u + b * UROW, *q = a.q + b * XROW, *r = a.r + b * UROW, *p = a.p + b * XROW;
    T *d = a.d + b * UROW, *v = a.v + b * XROW, *vnew = a.vnew + b * XROW, *z = a.z + b * UROW, *znew = a.znew + b * UROW;
    T *g = a.g + b * XROW, *y = a.y + b * UROW;
    const T *xr = a.Xref + b * a.xref_stride;
    if (which == STEP_FORWARD) {                                                      // admm.cpp:27-37
        T xi[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) xi[j] = x[j];
#pragma unroll 1
        for (int i = 0; i < NH - 1; ++i) {
            T ui[NU], xn[NX];
#pragma unroll
            for (int rr = 0; rr < NU; ++rr) {
                T kx = dot<T, O::Kx, NX, FAST>([&](int k) { return P.K[rr + k * NU]; }, [&](int k) { return xi[k]; });
                ui[rr] = N::sub(-kx, d[i * NU + rr]);
                u[i * NU + rr] = ui[rr];
            }
#pragma unroll
            for (int rr = 0; rr < NX; ++rr) {
                T ax = dot<T, O::Ax, NX, FAST>([&](int k) { return P.A[rr + k * NX]; }, [&](int k) { return xi[k]; });
                T bu = dot<T, O::Bu, NU, FAST>([&](int k) { return P.B[rr + k * NX]; }, [&](int k) { return ui[k]; });
                xn[rr] = N::add(ax, bu);
            }
#pragma unroll
            for (int j = 0; j < NX; ++j) { xi[j] = xn[j]; x[(i + 1) * NX + j] = xn[j]; }
        }
    } else if (which == STEP_SLACK) {                                                 // :45-61 (bounds are +-inf when disabled)
        for (int k = 0; k < UROW; ++k) znew[k] = N::mn(P.umax[k], N::mx(P.umin[k], N::add(u[k], y[k])));
        for (int k = 0; k < XROW; ++k) vnew[k] = N::mn(P.xmax[k], N::mx(P.xmin[k], N::add(x[k], g[k])));
    } else if (which == STEP_DUAL) {                                                  // :67-71
        for (int k = 0; k < UROW; ++k) y[k] = N::sub(N::add(y[k], u[k]), znew[k]);
        for (int k = 0; k < XROW; ++k) g[k] = N::sub(N::add(g[k], x[k]), vnew[k]);
    } else if (which == STEP_LINCOST) {                                               // :77-85
        for (int k = 0; k < UROW; ++k) r[k] = N::mul(P.nrho, N::sub(znew[k], y[k]));
        for (int i = 0; i < NH; ++i)
            for (int j = 0; j < NX; ++j) {
                const int k = i * NX + j;
                q[k] = N::sub(-N::mul(xr[k], P.Qd[j]), N::mul(P.rho, N::sub(vnew[k], g[k])));
            }
        T xl[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) xl[j] = xr[(NH - 1) * NX + j];
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            T pn = -dot<T, O::XtP, NX, FAST>([&](int k) { return P.Pf[k + j * NX]; }, [&](int k) { return xl[k]; });
            const int k = (NH - 1) * NX + j;
            p[k] = N::sub(pn, N::mul(P.rho, N::sub(vnew[k], g[k])));
        }
    } else if (which == STEP_TERM) {                                                  // :91-109
        int ok = 0;
        if (a.iter % P.check_term == 0) {
            T px = T(0), dx = T(0), pu = T(0), du = T(0);
            for (int k = 0; k < XROW; ++k) {
                px = N::mx(px, N::abs(N::sub(x[k], vnew[k])));
                dx = N::mx(dx, N::abs(N::sub(v[k], vnew[k])));
            }
            for (int k = 0; k < UROW; ++k) {
                pu = N::mx(pu, N::abs(N::sub(u[k], znew[k])));
                du = N::mx(du, N::abs(N::sub(z[k], znew[k])));
            }
            dx = N::mul(dx, P.rho);
            du = N::mul(du, P.rho);
            a.resid[b * 4 + 0] = px; a.resid[b * 4 + 1] = dx; a.resid[b * 4 + 2] = pu; a.resid[b * 4 + 3] = du;
            ok = (px < P.pri_tol && pu < P.pri_tol && dx < P.dua_tol && du < P.dua_tol) ? 1 : 0;
        }
        if (a.term) a.term[b] = ok;
    } else if (which == STEP_BACKWARD) {                                              // :15-22
        T pn[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) pn[j] = p[(NH - 1) * NX + j];
#pragma unroll 1
        for (int i = NH - 2; i >= 0; --i) {
            T ri[NU], s[NU], pi[NX];
#pragma unroll
            for (int j = 0; j < NU; ++j) ri[j] = r[i * NU + j];
#pragma unroll
            for (int rr = 0; rr < NU; ++rr)
                s[rr] = N::add(dot<T, O::Btp, NX, FAST>([&](int k) { return P.B[k + rr * NX]; }, [&](int k) { return pn[k]; }), ri[rr]);
#pragma unroll
            for (int rr = 0; rr < NU; ++rr)
                d[i * NU + rr] = dot<T, O::Qs, NU, FAST>([&](int k) { return P.Qi[rr + k * NU]; }, [&](int k) { return s[k]; });
#pragma unroll
            for (int rr = 0; rr < NX; ++rr) {
                T mp = dot<T, O::Mp, NX, FAST>([&](int k) { return P.M[rr + k * NX]; }, [&](int k) { return pn[k]; });
                T kr = dot<T, O::Ktr, NU, FAST>([&](int k) { return P.K[k + rr * NU]; }, [&](int k) { return ri[k]; });
                pi[rr] = N::sub(N::add(q[i * NX + rr], mp), kr);
            }
#pragma unroll
            for (int j = 0; j < NX; ++j) { pn[j] = pi[j]; p[i * NX + j] = pi[j]; }
        }
    }
}

}  // namespace tmpc
