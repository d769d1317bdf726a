// Host C++ layer: the reference's entry points (include/tinympc/*.hpp) on top of the C ABI (include/tmpc.h).
// Nothing numerical from the solver loop runs here; tiny_precompute is the only host-side math (the cold,
// once-per-model Riccati recursion the reference performs inside tiny_codegen, codegen.cpp:254-292).
#include "tinympc/tiny_api.hpp"
#include "tmpc.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(const std::string &m) { g_err = m; return -1; }

constexpr int kDtype = sizeof(tinytype) == 4 ? TMPC_F32 : TMPC_F64;

struct Backend {
    tmpc_ctx *ctx = nullptr;        // device 0: single instances, step functions, device-memory batches
    tmpc_multi *multi = nullptr;    // every selected device: host-memory batches (created on first use)
    int policy = TMPC_ORDER_PARITY;
    int devices = 0;                // tiny_set_devices: 0 = every visible device, n = the first n
    // what was last pushed to the device contexts: tmpc_set_model / tmpc_set_settings run only when something changed
    uint64_t pushed_ctx = 0, pushed_multi = 0;
    // per-instance bounds (tiny_set_instance_bounds): host copies so that they survive a context re-creation and can be
    // installed on whichever backend the next tiny_solve_batch uses
    int64_t ib_batch = 0;
    bool ib_on_device = false;
    std::vector<tinytype> ib[4];
    const tinytype *ib_dev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool ib_ctx = false, ib_multi = false;   // installed on ctx / on multi
};

Backend *backend(TinySolver *s)
{
    if (!s->backend) s->backend = new Backend;
    return static_cast<Backend *>(s->backend);
}

// FNV-1a over everything tmpc_set_model / tmpc_set_settings take: the reference lets callers edit cache / work / settings
// fields freely between solves, so every call looks at them -- but the device copies are rebuilt (a stream sync, a blocking
// copy and an image rebuild inside tmpc_set_model) only when the fingerprint moved.
uint64_t fingerprint(const TinySolver *s)
{
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void *p, size_t n) {
        const unsigned char *b = static_cast<const unsigned char *>(p);
        for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    };
    const TinyWorkspace &w = *s->work;
    const TinyCache &c = *s->cache;
    const TinySettings &t = *s->settings;
    for (const tiny_Matrix *m : {&c.Kinf, &c.Pinf, &c.Quu_inv, &c.AmBKt, &w.Adyn, &w.Bdyn, &w.Q, &w.x_min, &w.x_max, &w.u_min, &w.u_max})
        mix(m->data(), sizeof(tinytype) * (size_t)m->rows() * m->cols());
    mix(&c.rho, sizeof c.rho);
    mix(&t.abs_pri_tol, sizeof t.abs_pri_tol); mix(&t.abs_dua_tol, sizeof t.abs_dua_tol);
    const int iv[4] = {t.max_iter, t.check_termination, t.en_state_bound, t.en_input_bound};
    mix(iv, sizeof iv);
    return h | 1ull;   // never 0 (= nothing pushed yet)
}

// (re)create the device context if needed and push model + settings when they changed since the last push
int sync_model(TinySolver *s)
{
    Backend *b = backend(s);
    if (!b->ctx) {
        int rc = tmpc_create(&b->ctx, 0, s->nx, s->nu, s->N, kDtype, b->policy);
        if (rc != TMPC_OK) return fail(std::string("tmpc_create: ") + tmpc_last_error(nullptr));
        b->pushed_ctx = 0;
        b->ib_ctx = false;
    }
    const uint64_t fp = fingerprint(s);
    if (fp == b->pushed_ctx) return 0;
    const TinyWorkspace &w = *s->work;
    const TinyCache &c = *s->cache;
    int rc = tmpc_set_model(b->ctx, c.Kinf.data(), c.Pinf.data(), c.Quu_inv.data(), c.AmBKt.data(), w.Adyn.data(),
                            w.Bdyn.data(), w.Q.data(), (double)c.rho, w.x_min.data(), w.x_max.data(), w.u_min.data(),
                            w.u_max.data());
    if (rc != TMPC_OK) return fail(std::string("tmpc_set_model: ") + tmpc_last_error(b->ctx));
    const TinySettings &t = *s->settings;
    rc = tmpc_set_settings(b->ctx, (double)t.abs_pri_tol, (double)t.abs_dua_tol, t.max_iter, t.check_termination,
                           t.en_state_bound, t.en_input_bound);
    if (rc != TMPC_OK) return fail(std::string("tmpc_set_settings: ") + tmpc_last_error(b->ctx));
    b->pushed_ctx = fp;
    return 0;
}

// the same for the multi-device backend (host-memory batches); returns 1 when there is only one device to use
int sync_multi(TinySolver *s)
{
    Backend *b = backend(s);
    if (!b->multi) {
        int rc = tmpc_multi_create(&b->multi, b->devices, nullptr, s->nx, s->nu, s->N, kDtype, b->policy);
        if (rc != TMPC_OK) return fail(std::string("tmpc_multi_create: ") + tmpc_multi_last_error(nullptr));
        b->pushed_multi = 0;
        b->ib_multi = false;
    }
    const uint64_t fp = fingerprint(s);
    if (fp == b->pushed_multi) return 0;
    const TinyWorkspace &w = *s->work;
    const TinyCache &c = *s->cache;
    int rc = tmpc_multi_set_model(b->multi, c.Kinf.data(), c.Pinf.data(), c.Quu_inv.data(), c.AmBKt.data(), w.Adyn.data(),
                                  w.Bdyn.data(), w.Q.data(), (double)c.rho, w.x_min.data(), w.x_max.data(), w.u_min.data(),
                                  w.u_max.data());
    if (rc != TMPC_OK) return fail(std::string("tmpc_multi_set_model: ") + tmpc_multi_last_error(b->multi));
    const TinySettings &t = *s->settings;
    rc = tmpc_multi_set_settings(b->multi, (double)t.abs_pri_tol, (double)t.abs_dua_tol, t.max_iter, t.check_termination,
                                 t.en_state_bound, t.en_input_bound);
    if (rc != TMPC_OK) return fail(std::string("tmpc_multi_set_settings: ") + tmpc_multi_last_error(b->multi));
    b->pushed_multi = fp;
    return 0;
}

void drop_contexts(Backend *b)
{
    if (b->ctx) { tmpc_destroy(b->ctx); b->ctx = nullptr; }
    if (b->multi) { tmpc_multi_destroy(b->multi); b->multi = nullptr; }
    b->pushed_ctx = b->pushed_multi = 0;
    b->ib_ctx = b->ib_multi = false;
}

// ---- small dense helpers in double, row-major std::vector (cold path only)
typedef std::vector<double> Mat;
Mat mul(const Mat &A, int ar, int ac, const Mat &B, int bc)
{
    Mat C((size_t)ar * bc, 0.0);
    for (int i = 0; i < ar; ++i)
        for (int k = 0; k < ac; ++k) {
            const double a = A[(size_t)i * ac + k];
            for (int j = 0; j < bc; ++j) C[(size_t)i * bc + j] += a * B[(size_t)k * bc + j];
        }
    return C;
}
Mat transpose(const Mat &A, int r, int c)
{
    Mat T((size_t)r * c);
    for (int i = 0; i < r; ++i)
        for (int j = 0; j < c; ++j) T[(size_t)j * r + i] = A[(size_t)i * c + j];
    return T;
}
bool inverse(Mat A, int n, Mat &inv)  // Gauss-Jordan with partial pivoting
{
    inv.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r)
            if (std::fabs(A[(size_t)r * n + c]) > std::fabs(A[(size_t)piv * n + c])) piv = r;
        if (A[(size_t)piv * n + c] == 0.0) return false;
        if (piv != c)
            for (int j = 0; j < n; ++j) {
                std::swap(A[(size_t)piv * n + j], A[(size_t)c * n + j]);
                std::swap(inv[(size_t)piv * n + j], inv[(size_t)c * n + j]);
            }
        const double d = 1.0 / A[(size_t)c * n + c];
        for (int j = 0; j < n; ++j) { A[(size_t)c * n + j] *= d; inv[(size_t)c * n + j] *= d; }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const double f = A[(size_t)r * n + c];
            if (f == 0.0) continue;
            for (int j = 0; j < n; ++j) { A[(size_t)r * n + j] -= f * A[(size_t)c * n + j]; inv[(size_t)r * n + j] -= f * inv[(size_t)c * n + j]; }
        }
    }
    return true;
}
Mat to_rowmajor(const tiny_Matrix &m)
{
    Mat A((size_t)m.rows() * m.cols());
    for (int i = 0; i < m.rows(); ++i)
        for (int j = 0; j < m.cols(); ++j) A[(size_t)i * m.cols() + j] = (double)m(i, j);
    return A;
}
void from_rowmajor(tiny_Matrix &m, const Mat &A, int r, int c)
{
    m.resize(r, c);
    for (int i = 0; i < r; ++i)
        for (int j = 0; j < c; ++j) m(i, j) = (tinytype)A[(size_t)i * c + j];
}

// single-instance step through tmpc_step on the host workspace
int run_step(TinySolver *s, int which, int *term)
{
    if (sync_model(s) != 0) return -1;
    TinyWorkspace &w = *s->work;
    tinytype res[4] = {w.primal_residual_state, w.dual_residual_state, w.primal_residual_input, w.dual_residual_input};
    tmpc_workspace ws;
    ws.x = w.x.data(); ws.u = w.u.data(); ws.q = w.q.data(); ws.r = w.r.data(); ws.p = w.p.data(); ws.d = w.d.data();
    ws.v = w.v.data(); ws.vnew = w.vnew.data(); ws.z = w.z.data(); ws.znew = w.znew.data(); ws.g = w.g.data(); ws.y = w.y.data();
    ws.Xref = w.Xref.data(); ws.xref_shared = 1; ws.resid = res;
    int32_t t = 0;
    ws.term = &t;
    int rc = tmpc_step(backend(s)->ctx, which, 1, &ws, w.iter, TMPC_MEM_HOST, nullptr);
    if (rc != TMPC_OK) return fail(std::string("tmpc_step: ") + tmpc_last_error(backend(s)->ctx));
    w.primal_residual_state = res[0]; w.dual_residual_state = res[1];
    w.primal_residual_input = res[2]; w.dual_residual_input = res[3];
    if (term) *term = t;
    return 0;
}

}  // namespace

namespace tinyhost { int set_error(const std::string &m) { return fail(m); } }   // for the other host translation units

extern "C" {

const char *tiny_last_error(void) { return g_err.c_str(); }

int tiny_setup(TinySolver **out, int nx, int nu, int N, const tinytype *Adyn, const tinytype *Bdyn, const tinytype *Q,
               const tinytype *R, tinytype rho, const tinytype *x_min, const tinytype *x_max, const tinytype *u_min,
               const tinytype *u_max, int verbose)
{
    if (!out || !Adyn || !Bdyn || !Q || !R || nx < 1 || nu < 1 || N < 2) return fail("tiny_setup: bad argument");
    TinySolver *s = new TinySolver;
    s->settings = new TinySettings;
    s->cache = new TinyCache;
    s->work = new TinyWorkspace;
    s->nx = nx; s->nu = nu; s->N = N; s->backend = nullptr;
    TinyWorkspace &w = *s->work;
    TinyCache &c = *s->cache;
    // zero every work array, as examples/quadrotor_hovering.cpp:49-71
    for (tiny_Matrix *m : {&w.x, &w.q, &w.p, &w.v, &w.vnew, &w.g, &w.x_min, &w.x_max, &w.Xref}) m->resize(nx, N);
    for (tiny_Matrix *m : {&w.u, &w.r, &w.d, &w.z, &w.znew, &w.y, &w.u_min, &w.u_max, &w.Uref}) m->resize(nu, N - 1);
    w.primal_residual_state = w.primal_residual_input = w.dual_residual_state = w.dual_residual_input = 0;
    w.status = 0; w.iter = 0;
    w.Q.resize(nx, 1); w.R.resize(nu, 1); w.Qu.resize(nu, 1);
    w.Adyn.resize(nx, nx); w.Bdyn.resize(nx, nu);
    std::memcpy(w.Adyn.data(), Adyn, sizeof(tinytype) * nx * nx);   // column-major in, column-major stored
    std::memcpy(w.Bdyn.data(), Bdyn, sizeof(tinytype) * nx * nu);
    std::memcpy(w.Q.data(), Q, sizeof(tinytype) * nx);
    std::memcpy(w.R.data(), R, sizeof(tinytype) * nu);
    c.rho = rho;
    c.Kinf.resize(nu, nx); c.Pinf.resize(nx, nx); c.Quu_inv.resize(nu, nu); c.AmBKt.resize(nx, nx); c.coeff_d2p.resize(nx, nu);
    TinySettings &t = *s->settings;
    t.abs_pri_tol = (tinytype)1e-3; t.abs_dua_tol = (tinytype)1e-3; t.max_iter = 100; t.check_termination = 1;
    t.en_state_bound = (x_min && x_max) ? 1 : 0;   // codegen.cpp:227-243
    t.en_input_bound = (u_min && u_max) ? 1 : 0;
    if (t.en_state_bound) {
        std::memcpy(w.x_min.data(), x_min, sizeof(tinytype) * nx * N);
        std::memcpy(w.x_max.data(), x_max, sizeof(tinytype) * nx * N);
    }
    if (t.en_input_bound) {
        std::memcpy(w.u_min.data(), u_min, sizeof(tinytype) * nu * (N - 1));
        std::memcpy(w.u_max.data(), u_max, sizeof(tinytype) * nu * (N - 1));
    }
    if (verbose) printf("tiny_setup: nx=%d nu=%d N=%d rho=%g state bounds %s, input bounds %s\n", nx, nu, N, (double)rho,
                        t.en_state_bound ? "on" : "off", t.en_input_bound ? "on" : "off");
    *out = s;
    return 0;
}

int tiny_precompute(TinySolver *s)
{
    if (!s) return fail("tiny_precompute: NULL solver");
    const int n = s->nx, m = s->nu;
    const double rho = (double)s->cache->rho;
    const Mat A = to_rowmajor(s->work->Adyn), B = to_rowmajor(s->work->Bdyn);
    const Mat At = transpose(A, n, n), Bt = transpose(B, n, m);
    Mat Q1((size_t)n * n, 0.0), R1((size_t)m * m, 0.0);
    for (int i = 0; i < n; ++i) Q1[(size_t)i * n + i] = (double)s->work->Q(i) + rho;      // codegen.cpp:255-258
    for (int i = 0; i < m; ++i) R1[(size_t)i * m + i] = (double)s->work->R(i) + rho;
    Mat Ktp1((size_t)m * n, 0.0), Ptp1((size_t)n * n, 0.0), Kinf((size_t)m * n, 0.0), Pinf((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) Ptp1[(size_t)i * n + i] = rho;
    int sweeps = 1000;
    for (int it = 0; it < 1000; ++it) {                                                   // codegen.cpp:273-285
        Mat BtP = mul(Bt, m, n, Ptp1, n);
        Mat S = mul(BtP, m, n, B, m);
        for (size_t k = 0; k < S.size(); ++k) S[k] += R1[k];
        Mat Sinv;
        if (!inverse(S, m, Sinv)) return fail("tiny_precompute: R + B'PB is singular");
        Kinf = mul(mul(Sinv, m, m, BtP, n), m, n, A, n);
        Mat BK = mul(B, n, m, Kinf, n);
        Mat AmBK(A);
        for (size_t k = 0; k < AmBK.size(); ++k) AmBK[k] -= BK[k];
        Pinf = mul(mul(At, n, n, Ptp1, n), n, n, AmBK, n);
        for (size_t k = 0; k < Pinf.size(); ++k) Pinf[k] += Q1[k];
        double dmax = 0.0;
        for (size_t k = 0; k < Kinf.size(); ++k) dmax = std::fmax(dmax, std::fabs(Kinf[k] - Ktp1[k]));
        if (dmax < 1e-5) { sweeps = it + 1; break; }
        Ktp1 = Kinf;
        Ptp1 = Pinf;
    }
    Mat S = mul(mul(Bt, m, n, Pinf, n), m, n, B, m);                                       // codegen.cpp:290-292
    for (size_t k = 0; k < S.size(); ++k) S[k] += R1[k];
    Mat Quu_inv;
    if (!inverse(S, m, Quu_inv)) return fail("tiny_precompute: R + B'PB is singular");
    Mat BK = mul(B, n, m, Kinf, n);
    Mat AmBK(A);
    for (size_t k = 0; k < AmBK.size(); ++k) AmBK[k] -= BK[k];
    Mat AmBKt = transpose(AmBK, n, n);
    Mat d2p = mul(transpose(Kinf, m, n), n, m, R1, m);
    Mat t2 = mul(mul(AmBKt, n, n, Pinf, n), n, n, B, m);
    for (size_t k = 0; k < d2p.size(); ++k) d2p[k] -= t2[k];
    from_rowmajor(s->cache->Kinf, Kinf, m, n);
    from_rowmajor(s->cache->Pinf, Pinf, n, n);
    from_rowmajor(s->cache->Quu_inv, Quu_inv, m, m);
    from_rowmajor(s->cache->AmBKt, AmBKt, n, n);
    from_rowmajor(s->cache->coeff_d2p, d2p, n, m);
    return sweeps;
}

int tiny_precompute_raw(int nx, int nu, const tinytype *Adyn, const tinytype *Bdyn, const tinytype *Q, const tinytype *R, tinytype rho,
                        tinytype *Kinf, tinytype *Pinf, tinytype *Quu_inv, tinytype *AmBKt)
{
    TinySolver *s = nullptr;
    if (tiny_setup(&s, nx, nu, 2, Adyn, Bdyn, Q, R, rho, nullptr, nullptr, nullptr, nullptr, 0) != 0) return -1;
    const int sweeps = tiny_precompute(s);
    if (sweeps >= 0) {
        if (Kinf) std::memcpy(Kinf, s->cache->Kinf.data(), sizeof(tinytype) * nu * nx);
        if (Pinf) std::memcpy(Pinf, s->cache->Pinf.data(), sizeof(tinytype) * nx * nx);
        if (Quu_inv) std::memcpy(Quu_inv, s->cache->Quu_inv.data(), sizeof(tinytype) * nu * nu);
        if (AmBKt) std::memcpy(AmBKt, s->cache->AmBKt.data(), sizeof(tinytype) * nx * nx);
    }
    tiny_free(s);
    return sweeps;
}

int tiny_rollout_batch(TinySolver *s, const TinyRolloutIn *in, TinyRolloutOut *out)
{
    if (!s || !in || !out) return fail("tiny_rollout_batch: null argument");
    if (in->batch < 1 || in->steps < 1 || !in->x0) return fail("tiny_rollout_batch: batch, steps >= 1 and x0 are required");
    if (!in->table && !in->Xref) return fail("tiny_rollout_batch: a fixed Xref or a reference table is required");
    if (in->table && in->table_rows < s->N) return fail("tiny_rollout_batch: the reference table needs at least N rows");
    Backend *b = backend(s);
    if (b->ib_batch) return fail("tiny_rollout_batch: per-instance bounds are set (tiny_set_instance_bounds): clear them first");
    if (in->batch >= 32768 && (b->devices ? b->devices : tmpc_device_count()) > 1) {
        // every selected device runs the loop of its contiguous instance range (tmpc_multi_rollout)
        if (sync_multi(s) != 0) return -1;
        tmpc_rollout_args r;
        std::memset(&r, 0, sizeof r);
        r.batch = in->batch; r.steps = in->steps; r.reset_duals = in->reset_duals ? 1 : 0;
        r.x0 = in->x0; r.Xref = in->Xref; r.xref_shared = in->xref_shared ? 1 : 0;
        r.table = in->table; r.table_rows = in->table_rows; r.start = in->start;
        r.x0_hist = out->x_hist; r.u0_hist = out->u0_hist; r.iter_hist = out->iter_hist; r.status_hist = out->status_hist;
        r.x = out->x; r.u = out->u;
        if (tmpc_multi_rollout(b->multi, &r) != TMPC_OK) return fail(std::string("tmpc_multi_rollout: ") + tmpc_multi_last_error(b->multi));
        return 0;
    }
    if (sync_model(s) != 0) return -1;
    tmpc_batch *bt = nullptr;
    if (tmpc_batch_create(b->ctx, in->batch, &bt) != TMPC_OK) return fail(std::string("tmpc_batch_create: ") + tmpc_last_error(b->ctx));
    auto bail = [&](const char *what) {
        g_err = std::string(what) + ": " + tmpc_batch_last_error(bt);
        tmpc_batch_destroy(bt);
        return -1;
    };
    if (tmpc_batch_set_x0(bt, in->x0, TMPC_MEM_HOST) != TMPC_OK) return bail("tmpc_batch_set_x0");
    if (in->table) {
        if (tmpc_batch_set_xref_table(bt, in->table, in->table_rows, in->start, TMPC_MEM_HOST) != TMPC_OK) return bail("tmpc_batch_set_xref_table");
    } else if (tmpc_batch_set_xref(bt, in->Xref, in->xref_shared ? 1 : 0, TMPC_MEM_HOST) != TMPC_OK) return bail("tmpc_batch_set_xref");
    if (tmpc_batch_rollout(bt, in->steps, in->reset_duals ? 1 : 0, out->x_hist, out->u0_hist, out->iter_hist, out->status_hist, TMPC_MEM_HOST) != TMPC_OK)
        return bail("tmpc_batch_rollout");
    if (out->x && tmpc_batch_get(bt, TMPC_GET_X, out->x, TMPC_MEM_HOST) != TMPC_OK) return bail("tmpc_batch_get(x)");
    if (out->u && tmpc_batch_get(bt, TMPC_GET_U, out->u, TMPC_MEM_HOST) != TMPC_OK) return bail("tmpc_batch_get(u)");
    tmpc_batch_destroy(bt);
    return 0;
}

int tiny_set_order_policy(TinySolver *s, int policy)
{
    if (!s || (policy != TMPC_ORDER_PARITY && policy != TMPC_ORDER_FAST)) return fail("tiny_set_order_policy: bad argument");
    Backend *b = backend(s);
    // the policy is a property of the device contexts: they are re-created on the next call; per-instance bounds are kept
    // in the Backend and re-installed then (install_instance_bounds)
    if (b->policy != policy) drop_contexts(b);
    b->policy = policy;
    return 0;
}

int tiny_set_devices(TinySolver *s, int n)
{
    if (!s || n < 0) return fail("tiny_set_devices: bad argument");
    const int visible = tmpc_device_count();
    if (n > visible) return fail("tiny_set_devices: more devices than are visible");
    Backend *b = backend(s);
    if (b->devices != n && b->multi) { tmpc_multi_destroy(b->multi); b->multi = nullptr; b->pushed_multi = 0; b->ib_multi = false; }
    b->devices = n;
    return 0;
}

int tiny_solve(TinySolver *s)
{
    if (!s) return fail("tiny_solve: NULL solver");
    if (sync_model(s) != 0) return -1;
    TinyWorkspace &w = *s->work;
    tmpc_warm warm = {w.d.data(), w.y.data(), w.g.data(), w.v.data(), w.z.data()};
    tinytype res[4];
    int32_t it = 0, st = 0;
    tmpc_solve_args a;
    std::memset(&a, 0, sizeof a);
    a.batch = 1;
    a.x0 = w.x.col(0);          // work.x.col(0) = x0 (quadrotor_hovering.cpp:95)
    std::vector<tinytype> x0(w.x.col(0), w.x.col(0) + s->nx);   // x is also an output: keep the input separate
    a.x0 = x0.data();
    a.Xref = w.Xref.data();
    a.xref_shared = 1;
    a.mem = TMPC_MEM_HOST;
    a.warm = &warm;
    a.x = w.x.data(); a.u = w.u.data(); a.iter = &it; a.status = &st; a.resid = res;
    int rc = tmpc_solve(backend(s)->ctx, &a);
    if (rc != TMPC_OK) return fail(std::string("tmpc_solve: ") + tmpc_last_error(backend(s)->ctx));
    w.iter = it; w.status = st;
    w.primal_residual_state = res[0]; w.dual_residual_state = res[1];
    w.primal_residual_input = res[2]; w.dual_residual_input = res[3];
    return st == TMPC_STATUS_SOLVED ? 0 : 1;    // admm.cpp:137,151
}

// Host batches of at least 32,768 instances go over every selected device (tmpc_multi: one ctx + one worker thread per
// device, contiguous instance ranges); everything else -- small batches, device-memory batches -- runs on device 0.
static bool use_multi(TinySolver *s, const TinyBatchIn *in)
{
    Backend *b = backend(s);
    if (in->on_device || in->batch < 32768 || b->ib_on_device) return false;
    const int want = b->devices ? b->devices : tmpc_device_count();
    return want > 1;
}

// per-instance bounds live in the Backend; put them on the context family that is about to solve (and take them off the other
// one so that a later call with the shared box is not surprised)
static int install_instance_bounds(TinySolver *s, bool multi)
{
    Backend *b = backend(s);
    const int mem = b->ib_on_device ? TMPC_MEM_DEVICE : TMPC_MEM_HOST;
    const tinytype *p[4];
    for (int k = 0; k < 4; ++k) p[k] = b->ib_on_device ? b->ib_dev[k] : b->ib[k].data();
    if (b->ib_batch == 0) return 0;   // nothing set: tiny_set_instance_bounds(0) cleared both families
    if (multi) {
        if (!b->ib_multi) {
            int rc = tmpc_multi_set_instance_bounds(b->multi, b->ib_batch, p[0], p[1], p[2], p[3]);
            if (rc != TMPC_OK) return fail(std::string("tmpc_multi_set_instance_bounds: ") + tmpc_multi_last_error(b->multi));
            b->ib_multi = true;
        }
        return 0;
    }
    if (!b->ib_ctx) {
        int rc = tmpc_set_instance_bounds(b->ctx, b->ib_batch, p[0], p[1], p[2], p[3], mem);
        if (rc != TMPC_OK) return fail(std::string("tmpc_set_instance_bounds: ") + tmpc_last_error(b->ctx));
        b->ib_ctx = true;
    }
    return 0;
}

int tiny_solve_batch(TinySolver *s, const TinyBatchIn *in, TinyBatchOut *out)
{
    if (!s || !in || !out) return fail("tiny_solve_batch: NULL argument");
    Backend *b = backend(s);
    tmpc_warm warm = {in->d, in->y, in->g, in->v, in->z};
    const bool any = in->d || in->y || in->g || in->v || in->z;
    tmpc_solve_args a;
    std::memset(&a, 0, sizeof a);
    a.batch = in->batch; a.x0 = in->x0; a.Xref = in->Xref; a.xref_shared = in->xref_shared;
    a.mem = in->on_device ? TMPC_MEM_DEVICE : TMPC_MEM_HOST;
    a.warm = any ? &warm : nullptr;
    a.x = out->x; a.u = out->u; a.u0 = out->u0; a.iter = out->iter; a.status = out->status; a.resid = out->resid;
    a.stream = in->stream;
    if (use_multi(s, in)) {
        if (sync_multi(s) != 0) return -1;
        if (install_instance_bounds(s, true) != 0) return -1;
        int rc = tmpc_multi_solve(b->multi, &a);
        if (rc != TMPC_OK) return fail(std::string("tmpc_multi_solve: ") + tmpc_multi_last_error(b->multi));
        return 0;
    }
    if (sync_model(s) != 0) return -1;
    if (install_instance_bounds(s, false) != 0) return -1;
    int rc = tmpc_solve(b->ctx, &a);
    if (rc != TMPC_OK) return fail(std::string("tmpc_solve: ") + tmpc_last_error(b->ctx));
    return 0;
}

int tiny_set_instance_bounds(TinySolver *s, int64_t batch, const tinytype *x_min, const tinytype *x_max, const tinytype *u_min,
                             const tinytype *u_max, int on_device)
{
    if (!s) return fail("tiny_set_instance_bounds: NULL argument");
    if (batch < 0) return fail("tiny_set_instance_bounds: negative batch");
    if (batch > 0 && (!x_min || !x_max || !u_min || !u_max)) return fail("tiny_set_instance_bounds: all four bound arrays must be given");
    Backend *b = backend(s);
    // forget what is installed anywhere, keep the new boxes here (host arrays are copied; device arrays stay the caller's until
    // the context has taken its copy, which happens below)
    if (b->ib_ctx && b->ctx) tmpc_set_instance_bounds(b->ctx, 0, nullptr, nullptr, nullptr, nullptr, TMPC_MEM_HOST);
    if (b->ib_multi && b->multi) tmpc_multi_set_instance_bounds(b->multi, 0, nullptr, nullptr, nullptr, nullptr);
    b->ib_ctx = b->ib_multi = false;
    b->ib_batch = batch;
    b->ib_on_device = batch > 0 && on_device != 0;
    const tinytype *src[4] = {x_min, x_max, u_min, u_max};
    const size_t per[4] = {(size_t)s->nx * s->N, (size_t)s->nx * s->N, (size_t)s->nu * (s->N - 1), (size_t)s->nu * (s->N - 1)};
    for (int k = 0; k < 4; ++k) {
        b->ib[k].clear();
        b->ib_dev[k] = nullptr;
        if (batch > 0 && !on_device) b->ib[k].assign(src[k], src[k] + (size_t)batch * per[k]);
        if (batch > 0 && on_device) b->ib_dev[k] = src[k];
    }
    if (batch == 0) return 0;
    if (sync_model(s) != 0) return -1;   // creates the ctx on first use: argument errors surface here, as before
    return install_instance_bounds(s, false);
}

void forward_pass(TinySolver *s) { run_step(s, 0, nullptr); }
void update_slack(TinySolver *s) { run_step(s, 1, nullptr); }
void update_dual(TinySolver *s) { run_step(s, 2, nullptr); }
void update_linear_cost(TinySolver *s) { run_step(s, 3, nullptr); }
bool termination_condition(TinySolver *s)
{
    int t = 0;
    run_step(s, 4, &t);
    return t != 0;
}
void backward_pass_grad(TinySolver *s) { run_step(s, 5, nullptr); }

void tiny_free(TinySolver *s)
{
    if (!s) return;
    if (s->backend) {
        Backend *b = static_cast<Backend *>(s->backend);
        drop_contexts(b);
        delete b;
    }
    delete s->settings;
    delete s->cache;
    delete s->work;
    delete s;
}

}  // extern "C"
