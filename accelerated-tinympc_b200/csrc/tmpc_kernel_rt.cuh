// Run-time-shape ADMM kernel: any nx, nu <= 64 and any horizon, float or double.
//
// The reference fixes NSTATES / NINPUTS / NHORIZON at compile time (glob_opts.hpp:3-9) and its vendored Eigen picks the
// floating-point EVALUATION ORDER of every product from those sizes at compile time.  The three BASELINE shapes have their
// own kernels (tmpc_kernel_f32 / _small / _warp .cuh) with that order baked into templates; this kernel serves every other
// shape with the order expressed as DATA: for each of the 8 products on the path the host (tmpc_orders_rt.hpp) emits a
// small postfix "reduction program" -- (leaf index, number of adds that follow) per term -- that reproduces the reference
// build's summation tree exactly, including the rows that Eigen's linear-vectorised assignment peels off by the run-time
// address of the destination column.  PARITY results are therefore bit-identical to tiny_solve for any shape
// (oracle/pin_shapes.py pins the dispatch rule against the compiled reference on 60+ shapes).
//
// Structure (same as tmpc_kernel.cuh, /root/reference/src/tinympc/admm.cpp:111-152): persistent grid, one thread owns one
// instance at a time, per-lane early exit + refill from a global counter.  The per-instance state does not fit on chip
// for arbitrary shapes, so it lives in a per-lane scratch area of HBM laid out [element][lane]: every access of a warp is
// one coalesced 128-byte (float) line and the working set of the resident lanes stays in the 126 MB L2 for small shapes.
// v/z are double-buffered (vnew/znew are written to the other buffer and the roles swap when the iteration continues),
// which gives the reference's "v, z one iteration behind on an early exit" for the warm-start write-back without extra
// traffic.  Scratch accesses bypass L1 (ld/st.global.cg): L1 is left to the shared model, bounds and programs, which are
// read through the read-only path at warp-uniform addresses.  The vectors a mat-vec sweeps (x_i, u_i / p_{i+1}, r_i and
// the product under construction) and the operand stack of the reduction programs are indexed at run time, so they live
// in SHARED memory, [element][thread] (conflict-free), sized by the shape at launch: (3 max(nx,nu) + 8) scalars per thread.
#pragma once
#include "tmpc_kernel.cuh"
#include "tmpc_steps.cuh"

namespace tmpc {

constexpr int RT_MAXD = 64;    // nx, nu <= 64
constexpr int RT_STACK = 8;    // operand stack of a reduction program (host checks the depth)
constexpr int RT_BLOCK = 128;
constexpr int RT_MLP = 8;      // scratch elements whose loads are issued together in the element-wise loops
constexpr int RT_MLPU = 4;     // ... in the loops that run a dot product per element

// One evaluation-order variant of one product: its reduction program and a ROW-MAJOR copy of the coefficient matrix with
// the columns of every row already in program order, so a dot product walks one contiguous run of K coefficients.
template <class T> struct VarRT {
    const T *coef;               // [rows][K]
    const unsigned short *prog;  // [K]: leaf | (adds << 8)
    int kind;                    // 0 sequential chain (no program read), 1 terms in natural order, 2 gather by leaf index,
                                 // 3 compile-time-K code for `order` (coefficients in natural order, no program read)
    int order;                   // ORD_* of tmpc_kernel.cuh (kind 3)
};
template <class T> struct ProdRT {
    VarRT<T> a, b;               // a = packet-path rows, b = scalar-path rows
};

template <class T> struct ModelRT {
    int nx, nu, N;
    const T *Qd, *xmin, *xmax, *umin, *umax;   // device image
    ProdRT<T> Kx, Ax, Bu, Btp, Qs, Mp, Ktr, XtP;
    // rows [lo, hi) of a destination column take variant a, the others variant b (tmpc_orders_rt.hpp):
    //   head < 0: all rows a;  else unrolled assignment: [0, head);  rt_*: peeled by the address of the column
    int head_Kx, head_Ax, head_Qs, head_Mp;
    int rt_u, rt_x, rt_p, off_u, off_x, off_p, pk, sb;
    T rho, nrho, pri_tol, dua_tol;
    int max_iter, check_term;
    T *scratch;       // [elements_per_lane][lanes]
    long long lanes;  // gridDim.x * blockDim.x
    int warm;
    // per-instance box bounds (tmpc_set_instance_bounds): [instance][N][nx] / [instance][N-1][nu]; a null pair = the shared rows above
    const T *ixmin, *ixmax, *iumin, *iumax;
};

// dynamic shared memory of one block
__host__ __device__ inline size_t rt_smem_bytes(int nx, int nu, size_t scalar)
{
    const int D = nx > nu ? nx : nu;
    return (size_t)(3 * D + RT_STACK + 3 * RT_MLPU) * RT_BLOCK * scalar;
}

// elements of per-lane scratch
__host__ __device__ inline long long rt_scratch_elems(int nx, int nu, int N)
{
    return 2LL * nx + 5LL * nu * (N - 1) + 4LL * nx * N;
}

// Tree-shaped orders with K known at compile time: the reduction is the fully unrolled register code of tmpc_kernel.cuh
// (dot<T, ORD, K>), one instance per K <= 64, reached through a warp-uniform switch.  Not inlined: one copy per scalar type.
template <class T, int ORD, int K> __device__ __forceinline__ T dot_fixed_k(const T *__restrict__ c, const T *__restrict__ xs)
{
    if constexpr (ORD == ORD_GEMV_ROW && K < Num<T>::PK) {
        return T(0);   // never selected by the host (tmpc_orders_rt.hpp)
    } else {
        return dot<T, ORD, K, false>([&](int k) { return __ldg(c + k); }, [&](int k) { return xs[k * RT_BLOCK]; });
    }
}
#define TMPC_RT_CASES4(b) TMPC_RT_CASE(b) TMPC_RT_CASE(b + 1) TMPC_RT_CASE(b + 2) TMPC_RT_CASE(b + 3)
#define TMPC_RT_CASES16(b) TMPC_RT_CASES4(b) TMPC_RT_CASES4(b + 4) TMPC_RT_CASES4(b + 8) TMPC_RT_CASES4(b + 12)
template <class T, int ORD> __device__ __noinline__ T dot_fixed(int K, const T *__restrict__ c, const T *__restrict__ xs)
{
    switch (K) {
#define TMPC_RT_CASE(k) case (k): return dot_fixed_k<T, ORD, (k)>(c, xs);
        TMPC_RT_CASES16(1) TMPC_RT_CASES16(17) TMPC_RT_CASES16(33) TMPC_RT_CASES16(49)
#undef TMPC_RT_CASE
    }
    return T(0);
}

// sum_k c[k] x(k): PARITY = the variant's tree over individually rounded products; FAST = one FMA chain (`acc0` continues
// a chain when `chain` is set).  c = the row's K coefficients in program order; xs = the thread's vector in shared memory
// (element stride RT_BLOCK); st = the thread's operand stack in shared memory.
template <class T, bool FAST>
__device__ __forceinline__ T dot_rt(const VarRT<T> &v, int row, int K, const T *__restrict__ xs, T *__restrict__ st, T acc0 = T(0),
                                    bool chain = false)
{
    using N = Num<T>;
    const T *__restrict__ c = v.coef + row * K;
    if constexpr (FAST) {
        T acc = chain ? N::fma(__ldg(c), xs[0], acc0) : N::mul(__ldg(c), xs[0]);
#pragma unroll 4
        for (int k = 1; k < K; ++k) acc = N::fma(__ldg(c + k), xs[k * RT_BLOCK], acc);
        return acc;
    } else {
        if (v.kind == 0) {
            T acc = N::mul(__ldg(c), xs[0]);
#pragma unroll 4
            for (int k = 1; k < K; ++k) acc = N::add(N::mul(__ldg(c + k), xs[k * RT_BLOCK]), acc);
            return acc;
        }
        if (v.kind == 3) {
            if (v.order == ORD_TREE) return dot_fixed<T, ORD_TREE>(K, c, xs);
            if (v.order == ORD_VECREDUX) return dot_fixed<T, ORD_VECREDUX>(K, c, xs);
            return dot_fixed<T, ORD_GEMV_ROW>(K, c, xs);
        }
        const unsigned short *__restrict__ pg = v.prog;
        const bool gather = v.kind == 2;
        int sp = 0;
        T acc = T(0);
        for (int k = 0; k < K; ++k) {
            const unsigned w = __ldg(pg + k);
            int nm = w >> 8;
            const T e = N::mul(__ldg(c + k), xs[(gather ? (int)(w & 0xff) : k) * RT_BLOCK]);
            if (nm == 0) {                 // push
                if (k) st[(sp++ & (RT_STACK - 1)) * RT_BLOCK] = acc;
                acc = e;
            } else {                       // first add joins the top of the stack with the new term, the rest pop
                acc = N::add(acc, e);
                while (--nm) acc = N::add(st[(--sp & (RT_STACK - 1)) * RT_BLOCK], acc);
            }
        }
        return acc;
    }
}

template <class T> __device__ __forceinline__ void rt_head(const ModelRT<T> &P, int R, int head, int rt, int off, int col, int &lo, int &hi)
{
    if (head < 0) { lo = 0; hi = R; return; }
    if (!rt) { lo = 0; hi = head; return; }
    const int mask = P.pk - 1;
    int first = (P.pk - (((off + col * R * P.sb) / P.sb) & mask)) & mask;
    if (first > R) first = R;
    lo = first;
    hi = first + ((R - first) / P.pk) * P.pk;
}

template <class T, bool FAST>
__global__ void __launch_bounds__(RT_BLOCK) admm_kernel_rt(const __grid_constant__ ModelRT<T> P, const __grid_constant__ SolveArgs<T> a)
{
    using N = Num<T>;
    const int nx = P.nx, nu = P.nu, NH = P.N;
    const int XROW = nx * NH, UROW = nu * (NH - 1);
    const long long lanes = P.lanes;
    const unsigned lane = threadIdx.x & 31;
    constexpr unsigned FULLM = 0xffffffffu;
    T *S = P.scratch + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    auto at = [&](int e) -> T * { return S + (long long)e * lanes; };   // hot loops walk these pointers by `lanes`
    auto ld = [&](int e) -> T { return __ldcg(at(e)); };
    auto stg = [&](int e, T v) { __stcg(at(e), v); };
    // run-time indexed vectors + operand stack in shared memory, [element][thread]
    extern __shared__ __align__(16) unsigned char rt_smem[];
    const int D = nx > nu ? nx : nu;
    T *va = reinterpret_cast<T *>(rt_smem) + threadIdx.x;
    T *vb = va + D * RT_BLOCK, *vc = vb + D * RT_BLOCK, *stk = vc + D * RT_BLOCK;
    T *pf = stk + RT_STACK * RT_BLOCK;   // prefetched scratch values of RT_MLPU rows (3 per row)
    constexpr int VS = RT_BLOCK;   // stride between consecutive elements of a thread's vector
    // scratch map
    const int oX0 = 0, oPN = nx, oD = 2 * nx, oY = oD + UROW, oZ0 = oY + UROW, oZ1 = oZ0 + UROW, oU = oZ1 + UROW,
              oG = oU + UROW, oV0 = oG + XROW, oV1 = oV0 + XROW, oXo = oV1 + XROW;

    long long inst = -1;
    int it = 0, cur = 0;
    bool active = false, exhausted = false;
    T res[4] = {T(0), T(0), T(0), T(0)};
    unsigned long long n_iter = 0, n_solved = 0, n_trips = 0, n_inst = 0;

    for (;;) {
        // ------------------------------------------------------------------ lane refill
        const bool need = !active && !exhausted;
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
                    inst = ci;
                    active = true;
                    it = 0;
                    cur = 0;
                    res[0] = res[1] = res[2] = res[3] = T(0);
                    for (int j = 0; j < nx; ++j) stg(oX0 + j, __ldg(a.x0 + inst * nx + j));
                    // p_N seed: -(Xref_{N-1}^T * Pinf)   (admm.cpp:83)
                    const T *xr = a.Xref + inst * a.xref_stride + (long long)(NH - 1) * nx;
                    for (int j = 0; j < nx; ++j) va[j * VS] = __ldg(xr + j);
                    for (int j = 0; j < nx; ++j) stg(oPN + j, -dot_rt<T, FAST>(P.XtP.a, j, nx, va, stk));
                    if (P.warm && a.wd) {
                        for (int e = 0; e < UROW; ++e) {
                            stg(oD + e, a.wd[inst * UROW + e]);
                            stg(oY + e, a.wy[inst * UROW + e]);
                            stg(oZ0 + e, a.wz[inst * UROW + e]);
                        }
                        for (int e = 0; e < XROW; ++e) {
                            stg(oG + e, a.wg[inst * XROW + e]);
                            stg(oV0 + e, a.wv[inst * XROW + e]);
                        }
                    } else {
                        for (int e = 0; e < UROW; ++e) { stg(oD + e, T(0)); stg(oY + e, T(0)); stg(oZ0 + e, T(0)); }
                        for (int e = 0; e < XROW; ++e) { stg(oG + e, T(0)); stg(oV0 + e, T(0)); }
                    }
                } else {
                    exhausted = true;
                }
            }
        }
        if (__all_sync(FULLM, !active)) break;
        ++n_trips;
        if (!active) continue;
        ++it;
        const int oV = cur ? oV1 : oV0, oVn = cur ? oV0 : oV1, oZ = cur ? oZ1 : oZ0, oZn = cur ? oZ0 : oZ1;

        // ------------------------------------------------------------------ forward sweep
        // forward_pass (admm.cpp:27-37) + update_slack (:45-61) + update_dual (:67-71) + residual maxima (:95-98)
        T pri_x = T(0), dua_x = T(0), pri_u = T(0), dua_u = T(0);
        for (int j = 0; j < nx; ++j) va[j * VS] = ld(oX0 + j);
        const T *bxl = P.xmin, *bxh = P.xmax, *bul = P.umin, *buh = P.umax;
        if (P.ixmin) { bxl = P.ixmin + inst * XROW; bxh = P.ixmax + inst * XROW; }
        if (P.iumin) { bul = P.iumin + inst * UROW; buh = P.iumax + inst * UROW; }
        for (int i = 0; i < NH; ++i) {
            T *pG = at(oG + i * nx), *pV = at(oV + i * nx), *pVn = at(oVn + i * nx), *pXo = at(oXo + i * nx);
            T *pD = at(oD + i * nu), *pY = at(oY + i * nu), *pZ = at(oZ + i * nu), *pZn = at(oZn + i * nu), *pU = at(oU + i * nu);
            // the scratch reads of RT_MLP elements are issued together (memory-level parallelism: the state streams from L2/HBM)
            for (int j0 = 0; j0 < nx; j0 += RT_MLP) {
                T gg[RT_MLP], vv[RT_MLP];
#pragma unroll
                for (int t = 0; t < RT_MLP; ++t)
                    if (j0 + t < nx) { gg[t] = __ldcg(pG + t * lanes); vv[t] = __ldcg(pV + t * lanes); }
#pragma unroll
                for (int t = 0; t < RT_MLP; ++t)
                    if (j0 + t < nx) {
                        const int j = j0 + t;
                        const T x = va[j * VS];
                        T g = gg[t];
                        T vn = N::add(x, g);                                                    // :48
                        vn = N::mn(__ldg(bxh + j), N::mx(__ldg(bxl + j), vn));                  // :59
                        pri_x = N::mx(pri_x, N::abs(N::sub(x, vn)));                            // :95
                        dua_x = N::mx(dua_x, N::abs(N::sub(vv[t], vn)));                        // :96
                        g = N::sub(N::add(g, x), vn);                                           // :70
                        __stcg(pG + t * lanes, g);
                        __stcg(pVn + t * lanes, vn);
                        __stcg(pXo + t * lanes, x);
                    }
                pG += RT_MLP * lanes; pV += RT_MLP * lanes; pVn += RT_MLP * lanes; pXo += RT_MLP * lanes;
            }
            bxl += nx; bxh += nx;
            if (i == NH - 1) break;
            int lo, hi;
            rt_head(P, nu, P.head_Kx, P.rt_u, P.off_u, i, lo, hi);
            for (int r0 = 0; r0 < nu; r0 += RT_MLPU) {
                {
                    T dd[RT_MLPU], yy[RT_MLPU], zz[RT_MLPU];
#pragma unroll
                    for (int t = 0; t < RT_MLPU; ++t)
                        if (r0 + t < nu) { dd[t] = __ldcg(pD + t * lanes); yy[t] = __ldcg(pY + t * lanes); zz[t] = __ldcg(pZ + t * lanes); }
#pragma unroll
                    for (int t = 0; t < RT_MLPU; ++t) { pf[(3 * t) * VS] = dd[t]; pf[(3 * t + 1) * VS] = yy[t]; pf[(3 * t + 2) * VS] = zz[t]; }
                }
#pragma unroll 1
                for (int t = 0; t < RT_MLPU; ++t)
                    if (r0 + t < nu) {
                        const int r = r0 + t;
                        const T kx = dot_rt<T, FAST>((r >= lo && r < hi) ? P.Kx.a : P.Kx.b, r, nx, va, stk);
                        const T u = N::sub(-kx, pf[(3 * t) * VS]);                              // :31
                        T y = pf[(3 * t + 1) * VS];
                        T zn = N::add(u, y);                                                    // :47
                        zn = N::mn(__ldg(buh + r), N::mx(__ldg(bul + r), zn));                  // :53
                        pri_u = N::mx(pri_u, N::abs(N::sub(u, zn)));                            // :97
                        dua_u = N::mx(dua_u, N::abs(N::sub(pf[(3 * t + 2) * VS], zn)));         // :98
                        y = N::sub(N::add(y, u), zn);                                           // :69
                        __stcg(pY + t * lanes, y);
                        __stcg(pZn + t * lanes, zn);
                        __stcg(pU + t * lanes, u);
                        vb[r * VS] = u;
                    }
                pD += RT_MLPU * lanes; pY += RT_MLPU * lanes; pZ += RT_MLPU * lanes; pZn += RT_MLPU * lanes; pU += RT_MLPU * lanes;
            }
            bul += nu; buh += nu;
            rt_head(P, nx, P.head_Ax, P.rt_x, P.off_x, i + 1, lo, hi);
            for (int r = 0; r < nx; ++r) {
                const bool pa = r >= lo && r < hi;
                const T ax = dot_rt<T, FAST>(pa ? P.Ax.a : P.Ax.b, r, nx, va, stk);
                if constexpr (FAST) {
                    vc[r * VS] = dot_rt<T, FAST>(P.Bu.a, r, nu, vb, stk, ax, true);
                } else {
                    const T bu = dot_rt<T, FAST>(pa ? P.Bu.a : P.Bu.b, r, nu, vb, stk);
                    vc[r * VS] = N::add(ax, bu);                                                // :35
                }
            }
            for (int j = 0; j < nx; ++j) va[j * VS] = vc[j * VS];
        }

        // ------------------------------------------------------------------ termination (admm.cpp:91-109, :135-138)
        const bool chk = (it % P.check_term) == 0;
        if (chk) {
            res[0] = pri_x;
            res[1] = N::mul(dua_x, P.rho);
            res[2] = pri_u;
            res[3] = N::mul(dua_u, P.rho);
        }
        const bool conv = chk && res[0] < P.pri_tol && res[2] < P.pri_tol && res[1] < P.dua_tol && res[3] < P.dua_tol;
        const bool fin = conv || it >= P.max_iter;
        const bool wback = P.warm && a.wd;
        // a max_iter exit still does v = vnew, z = znew and the backward pass of its last iteration (admm.cpp:141-144)
        if (!fin || (!conv && wback)) {
            cur ^= 1;
            const int oVc = cur ? oV1 : oV0, oZc = cur ? oZ1 : oZ0;
            // -------------------------------------------------------------- backward sweep
            // update_linear_cost (admm.cpp:77-85) recomputed per stage + backward_pass_grad (:15-22)
            const T *xr_base = a.Xref + inst * a.xref_stride;
            for (int j = 0; j < nx; ++j) {
                const int e = (NH - 1) * nx + j;
                const T dv = N::sub(ld(oVc + e), ld(oG + e));
                if constexpr (FAST) va[j * VS] = N::fma(P.nrho, dv, ld(oPN + j));
                else va[j * VS] = N::sub(ld(oPN + j), N::mul(P.rho, dv));                        // :84
            }
            for (int i = NH - 2; i >= 0; --i) {
                {
                    const T *pZ = at(oZc + i * nu), *pY = at(oY + i * nu);
                    for (int j0 = 0; j0 < nu; j0 += RT_MLP) {
                        T zz[RT_MLP], yy[RT_MLP];
#pragma unroll
                        for (int t = 0; t < RT_MLP; ++t)
                            if (j0 + t < nu) { zz[t] = __ldcg(pZ + t * lanes); yy[t] = __ldcg(pY + t * lanes); }
#pragma unroll
                        for (int t = 0; t < RT_MLP; ++t)
                            if (j0 + t < nu) vb[(j0 + t) * VS] = N::mul(P.nrho, N::sub(zz[t], yy[t]));   // :80
                        pZ += RT_MLP * lanes; pY += RT_MLP * lanes;
                    }
                }
                for (int r = 0; r < nu; ++r) {
                    const T bp = dot_rt<T, FAST>(P.Btp.a, r, nx, va, stk);
                    vc[r * VS] = N::add(bp, vb[r * VS]);
                }
                int lo, hi;
                rt_head(P, nu, P.head_Qs, 0, 0, i, lo, hi);
                T *pD = at(oD + i * nu);
                for (int r = 0; r < nu; ++r) {
                    __stcg(pD, dot_rt<T, FAST>((r >= lo && r < hi) ? P.Qs.a : P.Qs.b, r, nu, vc, stk));   // :19
                    pD += lanes;
                }
                rt_head(P, nx, P.head_Mp, P.rt_p, P.off_p, i, lo, hi);
                const T *pV = at(oVc + i * nx), *pG = at(oG + i * nx);
                for (int r0 = 0; r0 < nx; r0 += RT_MLPU) {   // p_i is built in vc (s is dead)
                    {
                        T vv[RT_MLPU], gg[RT_MLPU], xr[RT_MLPU];
#pragma unroll
                        for (int t = 0; t < RT_MLPU; ++t)
                            if (r0 + t < nx) {
                                vv[t] = __ldcg(pV + t * lanes); gg[t] = __ldcg(pG + t * lanes);
                                xr[t] = __ldg(xr_base + i * nx + r0 + t);
                            }
#pragma unroll
                        for (int t = 0; t < RT_MLPU; ++t) { pf[(3 * t) * VS] = vv[t]; pf[(3 * t + 1) * VS] = gg[t]; pf[(3 * t + 2) * VS] = xr[t]; }
                    }
#pragma unroll 1
                    for (int t = 0; t < RT_MLPU; ++t)
                        if (r0 + t < nx) {
                            const int r = r0 + t;
                            const T cq = -N::mul(pf[(3 * t + 2) * VS], __ldg(P.Qd + r));        // :81
                            const T dv = N::sub(pf[(3 * t) * VS], pf[(3 * t + 1) * VS]);
                            T q;
                            if constexpr (FAST) q = N::fma(P.nrho, dv, cq);
                            else q = N::sub(cq, N::mul(P.rho, dv));                              // :82
                            const T mp = dot_rt<T, FAST>((r >= lo && r < hi) ? P.Mp.a : P.Mp.b, r, nx, va, stk);
                            const T kr = dot_rt<T, FAST>(P.Ktr.a, r, nu, vb, stk);
                            vc[r * VS] = N::sub(N::add(q, mp), kr);                              // :20
                        }
                    pV += RT_MLPU * lanes; pG += RT_MLPU * lanes;
                }
                for (int j = 0; j < nx; ++j) va[j * VS] = vc[j * VS];
            }
        }
        if (fin) {
            if (a.iter) a.iter[inst] = it;
            if (a.status) a.status[inst] = conv ? 1 : 11;
            if (a.resid) {
                a.resid[inst * 4 + 0] = res[0];
                a.resid[inst * 4 + 1] = res[1];
                a.resid[inst * 4 + 2] = res[2];
                a.resid[inst * 4 + 3] = res[3];
            }
            if (a.x) for (int e = 0; e < XROW; ++e) a.x[inst * XROW + e] = ld(oXo + e);
            if (a.u) for (int e = 0; e < UROW; ++e) a.u[inst * UROW + e] = ld(oU + e);
            if (a.u0) for (int e = 0; e < nu; ++e) a.u0[inst * nu + e] = ld(oU + e);
            if (wback) {
                // converged: cur was not advanced -> v, z are those of the previous iteration (SURVEY 8a note W)
                const int oVc = cur ? oV1 : oV0, oZc = cur ? oZ1 : oZ0;
                for (int e = 0; e < UROW; ++e) {
                    a.wd[inst * UROW + e] = ld(oD + e);
                    a.wy[inst * UROW + e] = ld(oY + e);
                    a.wz[inst * UROW + e] = ld(oZc + e);
                }
                for (int e = 0; e < XROW; ++e) {
                    a.wg[inst * XROW + e] = ld(oG + e);
                    a.wv[inst * XROW + e] = ld(oVc + e);
                }
            }
            n_iter += (unsigned)it;
            n_solved += conv ? 1u : 0u;
            ++n_inst;
            active = false;
            if (a.done) { __threadfence(); atomicAdd(a.done + (inst >> a.done_shift), 1u); }
        }
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

// The six step functions of the reference (admm.hpp:13-18) for any shape, on full workspaces in global memory
// ([instance][stage][dim]); same evaluation orders as the fused kernel.  Unit-test surface (see tmpc_steps.cuh).
template <class T, bool FAST>
__global__ void __launch_bounds__(RT_BLOCK) step_kernel_rt(const __grid_constant__ ModelRT<T> P, const __grid_constant__ StepArgs<T> a, int which)
{
    using N = Num<T>;
    const int nx = P.nx, nu = P.nu, NH = P.N;
    const int XROW = nx * NH, UROW = nu * (NH - 1);
    extern __shared__ __align__(16) unsigned char rt_smem[];
    const int D = nx > nu ? nx : nu;
    T *va = reinterpret_cast<T *>(rt_smem) + threadIdx.x;
    T *vb = va + D * RT_BLOCK, *vc = vb + D * RT_BLOCK, *stk = vc + D * RT_BLOCK;
    constexpr int VS = RT_BLOCK;
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= a.batch) return;
    T *x = a.x + b * XROW, *u = a.u + b * UROW, *q = a.q + b * XROW, *r = a.r + b * UROW, *p = a.p + b * XROW;
    T *d = a.d + b * UROW, *v = a.v + b * XROW, *vnew = a.vnew + b * XROW, *z = a.z + b * UROW, *znew = a.znew + b * UROW;
    T *g = a.g + b * XROW, *y = a.y + b * UROW;
    const T *xr = a.Xref + b * a.xref_stride;
    int lo, hi;
    if (which == STEP_FORWARD) {                                                      // admm.cpp:27-37
        for (int j = 0; j < nx; ++j) va[j * VS] = x[j];
        for (int i = 0; i < NH - 1; ++i) {
            rt_head(P, nu, P.head_Kx, P.rt_u, P.off_u, i, lo, hi);
            for (int rr = 0; rr < nu; ++rr) {
                const T kx = dot_rt<T, FAST>((rr >= lo && rr < hi) ? P.Kx.a : P.Kx.b, rr, nx, va, stk);
                const T ui = N::sub(-kx, d[i * nu + rr]);
                u[i * nu + rr] = ui;
                vb[rr * VS] = ui;
            }
            rt_head(P, nx, P.head_Ax, P.rt_x, P.off_x, i + 1, lo, hi);
            for (int rr = 0; rr < nx; ++rr) {
                const bool pa = rr >= lo && rr < hi;
                const T ax = dot_rt<T, FAST>(pa ? P.Ax.a : P.Ax.b, rr, nx, va, stk);
                const T bu = dot_rt<T, FAST>(pa ? P.Bu.a : P.Bu.b, rr, nu, vb, stk);
                vc[rr * VS] = N::add(ax, bu);
            }
            for (int j = 0; j < nx; ++j) { va[j * VS] = vc[j * VS]; x[(i + 1) * nx + j] = vc[j * VS]; }
        }
    } else if (which == STEP_SLACK) {                                                 // :45-61 (bounds are +-inf when disabled)
        for (int k = 0; k < UROW; ++k) znew[k] = N::mn(__ldg(P.umax + k), N::mx(__ldg(P.umin + k), N::add(u[k], y[k])));
        for (int k = 0; k < XROW; ++k) vnew[k] = N::mn(__ldg(P.xmax + k), N::mx(__ldg(P.xmin + k), N::add(x[k], g[k])));
    } else if (which == STEP_DUAL) {                                                  // :67-71
        for (int k = 0; k < UROW; ++k) y[k] = N::sub(N::add(y[k], u[k]), znew[k]);
        for (int k = 0; k < XROW; ++k) g[k] = N::sub(N::add(g[k], x[k]), vnew[k]);
    } else if (which == STEP_LINCOST) {                                               // :77-85
        for (int k = 0; k < UROW; ++k) r[k] = N::mul(P.nrho, N::sub(znew[k], y[k]));
        for (int i = 0; i < NH; ++i)
            for (int j = 0; j < nx; ++j) {
                const int k = i * nx + j;
                q[k] = N::sub(-N::mul(xr[k], __ldg(P.Qd + j)), N::mul(P.rho, N::sub(vnew[k], g[k])));
            }
        for (int j = 0; j < nx; ++j) va[j * VS] = xr[(NH - 1) * nx + j];
        for (int j = 0; j < nx; ++j) {
            const T pn = -dot_rt<T, FAST>(P.XtP.a, j, nx, va, stk);
            const int k = (NH - 1) * nx + j;
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
        for (int j = 0; j < nx; ++j) va[j * VS] = p[(NH - 1) * nx + j];
        for (int i = NH - 2; i >= 0; --i) {
            for (int j = 0; j < nu; ++j) vb[j * VS] = r[i * nu + j];
            for (int rr = 0; rr < nu; ++rr) vc[rr * VS] = N::add(dot_rt<T, FAST>(P.Btp.a, rr, nx, va, stk), vb[rr * VS]);
            rt_head(P, nu, P.head_Qs, 0, 0, i, lo, hi);
            for (int rr = 0; rr < nu; ++rr) d[i * nu + rr] = dot_rt<T, FAST>((rr >= lo && rr < hi) ? P.Qs.a : P.Qs.b, rr, nu, vc, stk);
            rt_head(P, nx, P.head_Mp, P.rt_p, P.off_p, i, lo, hi);
            for (int rr = 0; rr < nx; ++rr) {
                const T mp = dot_rt<T, FAST>((rr >= lo && rr < hi) ? P.Mp.a : P.Mp.b, rr, nx, va, stk);
                const T kr = dot_rt<T, FAST>(P.Ktr.a, rr, nu, vb, stk);
                vc[rr * VS] = N::sub(N::add(q[i * nx + rr], mp), kr);
            }
            for (int j = 0; j < nx; ++j) { va[j * VS] = vc[j * VS]; p[i * nx + j] = vc[j * VS]; }
        }
    }
}

// The examples' plant step x1 = Adyn * x0 + Bdyn * u.col(0) (quadrotor_hovering.cpp:108) for any shape (see plant_kernel in
// tmpc_batch.cuh).  Both products are REGULAR Eigen products here: rows >= 8 && depth >= 8 go through the column-major GEMV
// (sequential for every row); smaller ones stay coefficient-based inside the sum and follow the packet / scalar rows of
// the assignment to a 16-byte aligned x1 (rows [0, nx/pk*pk) sequential, the rest the scalar tree) -- pinned against the
// compiled reference by oracle/pin_shapes.py.
template <class T, bool FAST>
__global__ void __launch_bounds__(RT_BLOCK) plant_kernel_rt(const __grid_constant__ ModelRT<T> P, long long batch, T *x0, const T *u, long long u_stride,
                                                            T *x_next_hist, T *u0_hist, const int *iter, const int *status, int *iter_hist, int *status_hist)
{
    using N = Num<T>;
    const int nx = P.nx, nu = P.nu;
    extern __shared__ __align__(16) unsigned char rt_smem[];
    const int D = nx > nu ? nx : nu;
    T *va = reinterpret_cast<T *>(rt_smem) + threadIdx.x;
    T *vb = va + D * RT_BLOCK, *vc = vb + D * RT_BLOCK, *stk = vc + D * RT_BLOCK;
    constexpr int VS = RT_BLOCK;
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= batch) return;
    for (int j = 0; j < nx; ++j) va[j * VS] = x0[b * nx + j];
    for (int j = 0; j < nu; ++j) vb[j * VS] = u[b * u_stride + j];
    const int head = P.head_Ax < 0 ? nx : (nx / P.pk) * P.pk;
    const int head_a = nx >= 8 ? nx : head, head_b = (nx >= 8 && nu >= 8) ? nx : head;
    for (int r = 0; r < nx; ++r) {
        const T ax = dot_rt<T, FAST>(r < head_a ? P.Ax.a : P.Ax.b, r, nx, va, stk);
        if constexpr (FAST) vc[r * VS] = dot_rt<T, FAST>(P.Bu.a, r, nu, vb, stk, ax, true);
        else vc[r * VS] = N::add(ax, dot_rt<T, FAST>(r < head_b ? P.Bu.a : P.Bu.b, r, nu, vb, stk));
    }
    for (int r = 0; r < nx; ++r) {
        const T v = vc[r * VS];
        x0[b * nx + r] = v;
        if (x_next_hist) x_next_hist[b * nx + r] = v;
    }
    if (u0_hist)
        for (int j = 0; j < nu; ++j) u0_hist[b * nu + j] = vb[j * VS];
    if (iter_hist) iter_hist[b] = iter[b];
    if (status_hist) status_hist[b] = status[b];
}

}  // namespace tmpc
