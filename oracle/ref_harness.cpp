// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// POD (ctypes-friendly) shim around the UNMODIFIED reference solver.  It is compiled once per
// (nx, nu, N, scalar) configuration by oracle/Makefile against the reference sources *where they
// lie* under /root/reference (the build stages symlinks to src/tinympc/{admm.cpp,admm.hpp,types.hpp}
// next to a generated glob_opts.hpp, which is what tiny_codegen itself does, codegen.cpp:131-160,
// 626-641).  Output: oracle/_ref/libref_<cfg>.so.  Nothing here re-implements solver math: every
// numerical result comes from the reference's own tiny_solve / step functions (admm.hpp:10-18).
//
// Wire layout of all arrays = the reference's own (Eigen column-major): a "nx x N" trajectory is
// [stage][state] contiguous (tiny_wrapper.cpp:27), matrices are column-major (codegen.cpp:245-252).
#include <tinympc/admm.hpp>

#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include <memory>

typedef tinytype T;

extern "C" {

struct ref_problem {
    const void *Kinf, *Pinf, *Quu_inv, *AmBKt, *Adyn, *Bdyn, *Q;  // col-major, scalar = tinytype
    const void *x_min, *x_max, *u_min, *u_max;                    // [N][nx], [N-1][nu]
    double rho, abs_pri_tol, abs_dua_tol;
    int32_t max_iter, check_termination, en_state_bound, en_input_bound;
};

// State that is live across tiny_solve calls (SURVEY 8a "state liveness") + the rest of the workspace.
struct ref_state {          // every pointer nullable; [B][...] in the wire layout
    void *d, *y, *g, *v, *z;                 // in/out (warm start); null => zeros in, nothing out
    void *vnew, *znew, *q, *r, *p;           // out only (workspace as left by tiny_solve)
};

void ref_info(int32_t *nx, int32_t *nu, int32_t *N, int32_t *scalar_bytes)
{
    *nx = NSTATES; *nu = NINPUTS; *N = NHORIZON; *scalar_bytes = (int32_t)sizeof(T);
}

}  // extern "C"

namespace {

struct Inst {
    TinyCache cache;
    // 16-byte aligned like a global/static TinyWorkspace (how the reference's examples and generated code hold it:
    // quadrotor_hovering.cpp:25-28, codegen.cpp:350-470).  Matters only for shapes whose nx or nu is not a multiple of
    // the SSE packet: Eigen then peels rows by the run-time address of each column (see tinympc_oracle.c).
    alignas(16) TinyWorkspace work;
    TinySettings settings;
    TinySolver solver;
};

void load_problem(Inst &I, const ref_problem *P)
{
    std::memset((void *)&I.work, 0, sizeof(I.work));
    std::memset((void *)&I.cache, 0, sizeof(I.cache));
    I.cache.rho = (T)P->rho;
    I.cache.Kinf = Eigen::Map<const tiny_MatrixNuNx>((const T *)P->Kinf);
    I.cache.Pinf = Eigen::Map<const tiny_MatrixNxNx>((const T *)P->Pinf);
    I.cache.Quu_inv = Eigen::Map<const tiny_MatrixNuNu>((const T *)P->Quu_inv);
    I.cache.AmBKt = Eigen::Map<const tiny_MatrixNxNx>((const T *)P->AmBKt);
    I.cache.coeff_d2p.setZero();
    I.work.Adyn = Eigen::Map<const tiny_MatrixNxNx>((const T *)P->Adyn);
    I.work.Bdyn = Eigen::Map<const tiny_MatrixNxNu>((const T *)P->Bdyn);
    I.work.Q = Eigen::Map<const tiny_VectorNx>((const T *)P->Q);
    I.work.R.setZero();
    if (P->u_min) I.work.u_min = Eigen::Map<const tiny_MatrixNuNhm1>((const T *)P->u_min);
    if (P->u_max) I.work.u_max = Eigen::Map<const tiny_MatrixNuNhm1>((const T *)P->u_max);
    if (P->x_min) I.work.x_min = Eigen::Map<const tiny_MatrixNxNh>((const T *)P->x_min);
    if (P->x_max) I.work.x_max = Eigen::Map<const tiny_MatrixNxNh>((const T *)P->x_max);
    I.settings.abs_pri_tol = (T)P->abs_pri_tol;
    I.settings.abs_dua_tol = (T)P->abs_dua_tol;
    I.settings.max_iter = P->max_iter;
    I.settings.check_termination = P->check_termination;
    I.settings.en_state_bound = P->en_state_bound;
    I.settings.en_input_bound = P->en_input_bound;
    I.solver.settings = &I.settings;
    I.solver.cache = &I.cache;
    I.solver.work = &I.work;
}

constexpr int NXN = NSTATES * NHORIZON;
constexpr int NUN = NINPUTS * (NHORIZON - 1);

template <class M> void get(M &m, const void *base, int64_t b, int n)
{
    if (base) std::memcpy(m.data(), (const T *)base + b * n, sizeof(T) * n);
    else m.setZero();
}
template <class M> void put(const M &m, void *base, int64_t b, int n)
{
    if (base) std::memcpy((T *)base + b * n, m.data(), sizeof(T) * n);
}

void run_range(const ref_problem *P, int64_t b0, int64_t b1, const T *x0, const T *Xref,
               int64_t xref_stride, const ref_state *S, T *x_out, T *u_out, int32_t *iter_out,
               int32_t *status_out, T *resid_out, int32_t *ret_out)
{
    std::unique_ptr<Inst> I(new Inst);
    load_problem(*I, P);
    TinyWorkspace &w = I->work;
    for (int64_t b = b0; b < b1; ++b) {
        // everything that tiny_solve does not overwrite before reading is (re)initialised here,
        // exactly like the zeroing block of examples/quadrotor_hovering.cpp:49-71
        w.x.setZero(); w.u.setZero(); w.q.setZero(); w.r.setZero(); w.p.setZero();
        w.vnew.setZero(); w.znew.setZero();
        get(w.d, S ? S->d : nullptr, b, NUN);
        get(w.y, S ? S->y : nullptr, b, NUN);
        get(w.g, S ? S->g : nullptr, b, NXN);
        get(w.v, S ? S->v : nullptr, b, NXN);
        get(w.z, S ? S->z : nullptr, b, NUN);
        std::memcpy(w.Xref.data(), Xref + b * xref_stride, sizeof(T) * NXN);
        for (int j = 0; j < NSTATES; ++j) w.x(j, 0) = x0[b * NSTATES + j];
        int rc = tiny_solve(&I->solver);
        if (ret_out) ret_out[b] = rc;
        put(w.x, x_out, b, NXN);
        put(w.u, u_out, b, NUN);
        if (iter_out) iter_out[b] = w.iter;
        if (status_out) status_out[b] = w.status;
        if (resid_out) {
            resid_out[4 * b + 0] = w.primal_residual_state;
            resid_out[4 * b + 1] = w.dual_residual_state;
            resid_out[4 * b + 2] = w.primal_residual_input;
            resid_out[4 * b + 3] = w.dual_residual_input;
        }
        if (S) {
            put(w.d, S->d, b, NUN); put(w.y, S->y, b, NUN); put(w.g, S->g, b, NXN);
            put(w.v, S->v, b, NXN); put(w.z, S->z, b, NUN);
            put(w.vnew, S->vnew, b, NXN); put(w.znew, S->znew, b, NUN);
            put(w.q, S->q, b, NXN); put(w.r, S->r, b, NUN); put(w.p, S->p, b, NXN);
        }
    }
}

}  // namespace

extern "C" {

// Loop the reference's tiny_solve (admm.cpp:111) over a batch, one private TinySolver per thread
// (tiny_solve is re-entrant on distinct solvers).  xref_stride = 0 => one shared Xref.
int ref_solve_batch(const ref_problem *P, int64_t B, const void *x0, const void *Xref,
                    int64_t xref_stride, const ref_state *S, void *x_out, void *u_out,
                    int32_t *iter_out, int32_t *status_out, void *resid_out, int32_t *ret_out,
                    int32_t nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if ((int64_t)nthreads > B) nthreads = (int32_t)(B > 0 ? B : 1);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        int64_t b0 = B * t / nthreads, b1 = B * (t + 1) / nthreads;
        th.emplace_back(run_range, P, b0, b1, (const T *)x0, (const T *)Xref, xref_stride, S,
                        (T *)x_out, (T *)u_out, iter_out, status_out, (T *)resid_out, ret_out);
    }
    for (auto &t : th) t.join();
    return 0;
}

// One reference step function (admm.hpp:13-18) on a caller-provided workspace image.
//   ws layout (scalars): x[NXN] u[NUN] q[NXN] r[NUN] p[NXN] d[NUN] v[NXN] vnew[NXN] z[NUN] znew[NUN]
//                        g[NXN] y[NUN] Xref[NXN] resid[4]   (resid order: pri_x, dua_x, pri_u, dua_u)
//   which: 0 forward_pass, 1 update_slack, 2 update_dual, 3 update_linear_cost,
//          4 termination_condition (iter given), 5 backward_pass_grad
int ref_step(const ref_problem *P, int32_t which, void *ws_, int32_t iter)
{
    std::unique_ptr<Inst> I(new Inst);
    load_problem(*I, P);
    TinyWorkspace &w = I->work;
    T *ws = (T *)ws_;
    T *px = ws, *pu = px + NXN, *pq = pu + NUN, *pr = pq + NXN, *pp = pr + NUN, *pd = pp + NXN,
      *pv = pd + NUN, *pvn = pv + NXN, *pz = pvn + NXN, *pzn = pz + NUN, *pg = pzn + NUN,
      *py = pg + NXN, *pxr = py + NUN, *pres = pxr + NXN;
    get(w.x, px, 0, NXN); get(w.u, pu, 0, NUN); get(w.q, pq, 0, NXN); get(w.r, pr, 0, NUN);
    get(w.p, pp, 0, NXN); get(w.d, pd, 0, NUN); get(w.v, pv, 0, NXN); get(w.vnew, pvn, 0, NXN);
    get(w.z, pz, 0, NUN); get(w.znew, pzn, 0, NUN); get(w.g, pg, 0, NXN); get(w.y, py, 0, NUN);
    get(w.Xref, pxr, 0, NXN);
    w.iter = iter;
    w.primal_residual_state = pres[0]; w.dual_residual_state = pres[1];
    w.primal_residual_input = pres[2]; w.dual_residual_input = pres[3];
    int rc = 0;
    switch (which) {
    case 0: forward_pass(&I->solver); break;
    case 1: update_slack(&I->solver); break;
    case 2: update_dual(&I->solver); break;
    case 3: update_linear_cost(&I->solver); break;
    case 4: rc = termination_condition(&I->solver) ? 1 : 0; break;
    case 5: backward_pass_grad(&I->solver); break;
    default: return -1;
    }
    put(w.x, px, 0, NXN); put(w.u, pu, 0, NUN); put(w.q, pq, 0, NXN); put(w.r, pr, 0, NUN);
    put(w.p, pp, 0, NXN); put(w.d, pd, 0, NUN); put(w.v, pv, 0, NXN); put(w.vnew, pvn, 0, NXN);
    put(w.z, pz, 0, NUN); put(w.znew, pzn, 0, NUN); put(w.g, pg, 0, NXN); put(w.y, py, 0, NUN);
    pres[0] = w.primal_residual_state; pres[1] = w.dual_residual_state;
    pres[2] = w.primal_residual_input; pres[3] = w.dual_residual_input;
    return rc;
}

// The plant step of the reference's closed-loop examples, the verbatim expression of
// examples/quadrotor_hovering.cpp:108 and quadrotor_tracking.cpp:112 on the same Eigen types:
//     x1 = work.Adyn * x0 + work.Bdyn * work.u.col(0);
// u = the [N-1][nu] input trajectory of the workspace (col(0) is used).  Also returns the tracking error the
// examples print before each solve, (x0 - work.Xref.col(1)).norm() (hovering.cpp:92), when xref/err are given.
int ref_plant_step(const ref_problem *P, const void *x0_, const void *u_, void *x1_, const void *xref_, void *err_)
{
    std::unique_ptr<Inst> I(new Inst);
    load_problem(*I, P);
    TinyWorkspace &work = I->work;
    // 16-byte aligned like the examples' locals when NSTATES * sizeof(tinytype) is a multiple of 16; for other sizes
    // Eigen peels rows by the run-time address of x1, so the alignment is fixed here (tinympc_oracle.c, plant step)
    alignas(16) tiny_VectorNx x0;
    alignas(16) tiny_VectorNx x1;
    std::memcpy(x0.data(), x0_, sizeof(T) * NSTATES);
    get(work.u, u_, 0, NUN);
    x1 = work.Adyn * x0 + work.Bdyn * work.u.col(0);
    std::memcpy(x1_, x1.data(), sizeof(T) * NSTATES);
    if (xref_ && err_) {
        get(work.Xref, xref_, 0, NXN);
        *(T *)err_ = (x0 - work.Xref.col(1)).norm();
    }
    return 0;
}

}  // extern "C"
