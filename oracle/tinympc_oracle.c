/* TEST INFRASTRUCTURE ONLY.
 *
 * CPU oracle for the batched TinyMPC ADMM path: a plain-C restatement of
 * /root/reference/src/tinympc/admm.cpp:15-152 (tiny_solve and its six step functions), bit-exact
 * against the reference compiled "-O3" without -m flags (SSE2, no FMA) for every shape listed in
 * select_orders() below.  PARITY IS PINNED: tests/test_oracle_vs_ref.py compares it bit for bit with
 * oracle/_ref (the unmodified reference built from /root/reference by oracle/Makefile) whenever that
 * library is present, and tests/test_golden.py compares it with the committed fixtures under
 * tests/golden/ that oracle/make_golden.py generated from oracle/_ref.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product (include/tmpc.h, accelerated-tinympc_b200/) never does.
 *
 * Build: make -C oracle oracle   (gcc -O2 -ffp-contract=off: no a*b+c fusion, IEEE semantics)
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { ORD_SEQ = 0, ORD_VECREDUX = 1, ORD_TREE = 2, ORD_VECLOOP = 3, ORD_GEMV_COL = 4, ORD_GEMV_ROW = 5 };

/* evaluation order of the 8 products on the path (names: lhs then rhs) */
typedef struct {
    int Kx;  /* Kinf * x_i                 admm.cpp:31 */
    int Ax;  /* Adyn * x_i                 admm.cpp:35 */
    int Bu;  /* Bdyn * u_i                 admm.cpp:35 */
    int Btp; /* Bdyn^T * p_{i+1}           admm.cpp:19 */
    int Qs;  /* Quu_inv * (B^T p + r)      admm.cpp:19 */
    int Mp;  /* AmBKt * p_{i+1}            admm.cpp:20 */
    int Ktr; /* Kinf^T * r_i               admm.cpp:20 */
    int XtP; /* Xref_{N-1}^T * Pinf        admm.cpp:83 */
    /* Mixed rows (output length not a multiple of the packet width): rows [0, head) take the order above, rows
     * [head, R) take ORD_TREE / ORD_SEQ (the scalar coefficient path of the linear-vectorised assignment,
     * AssignEvaluator.h dense_assignment_loop<LinearVectorizedTraversal, CompleteUnrolling>).  head < 0 = all rows. */
    int head_Kx, head_Ax, head_Qs, head_Mp;
    int tail_x, tail_u; /* order of the scalar tail rows for K = nx and K = nu products */
    /* Assignments too costly for complete unrolling (AssignEvaluator.h copy_using_evaluator_traits:
     * size * (1 + source coefficient cost) > EIGEN_UNROLLING_LIMIT * pk) run dense_assignment_loop<
     * LinearVectorizedTraversal, NoUnrolling>, which peels rows up to the first 16-byte aligned ADDRESS of the
     * destination column: rows [first, first + ((R-first)/pk)*pk) take the packet path, the others the scalar path.
     * rt_* = that applies to destination u.col(i) / x.col(i+1) / p.col(i); off_* = byte offset of the member in a
     * 16-byte aligned TinyWorkspace (types.hpp:52-97 member order; a member is 16-aligned iff its byte size is a
     * multiple of 16, else scalar-aligned). */
    int rt_u, rt_x, rt_p, off_u, off_x, off_p, pk, sb;
} orders_t;

/* rows [lo, hi) of an R-row destination column `col` take the packet order */
static void head_range(const orders_t *o, int R, int unrolled_head, int rt, int off, int col, int *lo, int *hi)
{
    if (unrolled_head < 0) { *lo = 0; *hi = R; return; }
    if (!rt) { *lo = 0; *hi = unrolled_head; return; }
    int mask = o->pk - 1;
    int first = (o->pk - (((off + col * R * o->sb) / o->sb) & mask)) & mask;
    if (first > R) first = R;
    *lo = first;
    *hi = first + ((R - first) / o->pk) * o->pk;
}

typedef struct {
    int32_t nx, nu, N, scalar_bytes;
    const void *Kinf, *Pinf, *Quu_inv, *AmBKt, *Adyn, *Bdyn, *Q;
    const void *x_min, *x_max, *u_min, *u_max;
    double rho, abs_pri_tol, abs_dua_tol;
    int32_t max_iter, check_termination, en_state_bound, en_input_bound;
} oracle_problem;

typedef struct {
    void *d, *y, *g, *v, *z;       /* in/out, nullable */
    void *vnew, *znew, *q, *r, *p; /* out, nullable */
} oracle_state;

/* Eigen's compile-time dispatch restated (GeneralProduct.h product_type_selector,
 * ProductEvaluators.h:574-575 CanVectorizeLhs, Redux.h traversal/unrolling selection).
 * pk = SSE packet width for the scalar. */
static int g_override[8] = {-1, -1, -1, -1, -1, -1, -1, -1};

/* experiment hook used while deriving the orders for a new shape (tests/tools only):
 * idx in the orders_t field order; order < 0 restores the built-in dispatch */
void oracle_set_order_override(int idx, int order)
{
    if (idx >= 0 && idx < 8) g_override[idx] = order;
}

static orders_t select_orders_builtin(int nx, int nu, int N, int scalar_bytes);

static orders_t select_orders(int nx, int nu, int N, int scalar_bytes)
{
    orders_t o = select_orders_builtin(nx, nu, N, scalar_bytes);
    int *f = &o.Kx;
    for (int i = 0; i < 8; ++i)
        if (g_override[i] >= 0) f[i] = g_override[i];
    return o;
}

static orders_t select_orders_builtin(int nx, int nu, int N, int scalar_bytes)
{
    orders_t o;
    const int pk = 16 / scalar_bytes;
    const int large = 8; /* EIGEN_CACHEFRIENDLY_PRODUCT_THRESHOLD, arch/Default/Settings.h:30-31 */
    /* lazyProduct(col-major lhs, column rhs) (admm.cpp:31,35): packet path over the rows when
     * rows % pk == 0 (ProductEvaluators.h:574-575) -> sequential over k.  A 1-row Kinf is stored row-major,
     * so its single coefficient is a vectorised redux over the contiguous row. */
    /* Completely unrolled scalar redux = half-split tree; beyond Redux.h's unrolling limit (cost 4K-1 against
     * EIGEN_UNROLLING_LIMIT * pk for the linear traversal, i.e. K <= 110 float / 55 double) it is a plain loop. */
    const int unroll_k = (110 * pk + 1) / 4;
    o.tail_x = nx <= unroll_k ? ORD_TREE : ORD_SEQ;
    o.tail_u = nu <= unroll_k ? ORD_TREE : ORD_SEQ;
    o.Kx = (nu == 1) ? (nx <= unroll_k ? ORD_VECREDUX : ORD_VECLOOP) : ORD_SEQ;
    o.head_Kx = (nu == 1) ? -1 : (nu / pk) * pk;
    o.Ax = ORD_SEQ;
    o.Bu = ORD_SEQ;
    o.head_Ax = (nx / pk) * pk;
    if (nx == 1) {
        /* a 1-row Bdyn is stored row-major (like a 1-row Kinf): its single coefficient is a vectorised redux over the
         * contiguous row */
        o.Bu = nu <= unroll_k ? ORD_VECREDUX : ORD_VECLOOP;
        o.head_Ax = -1;
    }
    /* Bdyn^T * p (admm.cpp:19) is a regular product evaluated into a temporary: GeneralProduct.h
     * product_type_selector<rows=nu, 1, depth=nx>: both >= 8 -> row-major GEMV; otherwise coefficient /
     * inner product = completely unrolled vectorised redux. */
    o.Btp = (nu >= large && nx >= large) ? ORD_GEMV_ROW : ORD_VECREDUX;
    /* Quu_inv * s: nu >= 8 -> col-major GEMV (sequential); else packet path (sequential). */
    o.Qs = ORD_SEQ;
    o.head_Qs = (nu >= large) ? -1 : (nu / pk) * pk;
    /* q + AmBKt.lazyProduct(p) - Kinf^T.lazyProduct(r) (admm.cpp:20): Kinf^T is a row-major nx x nu lhs, which
     * has no packet path against a single column, so the whole expression is evaluated per coefficient and
     * AmBKt's strided row gives a scalar tree; when nu == 1, Kinf^T is a contiguous column vector, the
     * expression is packet-evaluable and AmBKt*p takes the sequential packet path. */
    o.Mp = (nu == 1) ? ORD_SEQ : o.tail_x;
    o.head_Mp = (nu == 1) ? (nx / pk) * pk : -1;
    o.Ktr = nu <= unroll_k ? ORD_VECREDUX : ORD_VECLOOP;
    o.XtP = nx <= unroll_k ? ORD_VECREDUX : ORD_VECLOOP;
    if (o.Btp == ORD_VECREDUX && nx > unroll_k) o.Btp = ORD_VECLOOP;
    o.pk = pk; o.sb = scalar_bytes;
    {
        const int lim = 110 * pk;
        /* u.col(i) = -Kinf.lazyProduct(x) - d: source cost (4nx-1) + 1 + 1 + 1 */
        o.rt_u = (nu != 1 && nu % pk != 0 && nu * (1 + 4 * nx + 2) > lim);
        /* x.col(i+1) = A.lazyProduct(x) + B.lazyProduct(u): (4nx-1) + (4nu-1) + 1 */
        o.rt_x = (nx % pk != 0 && nx * (1 + 4 * nx + 4 * nu - 1) > lim);
        /* p.col(i) = q + AmBKt.lazyProduct(p) - Kinf^T.lazyProduct(r), packet-evaluable only when nu == 1 */
        o.rt_p = (nu == 1 && nx % pk != 0 && nx * (1 + 1 + (4 * nx - 1) + 1 + 3 + 1) > lim);
        int off = 0, bx = nx * N * scalar_bytes, bu = nu * (N - 1) * scalar_bytes;
        int sizes[6] = {bx, bu, bx, bu, bx, bu}; /* x u q r p d */
        int offs[6];
        for (int k = 0; k < 6; ++k) {
            int al = (sizes[k] % 16 == 0) ? 16 : scalar_bytes;
            off = (off + al - 1) / al * al;
            offs[k] = off;
            off += sizes[k];
        }
        o.off_x = offs[0]; o.off_u = offs[1]; o.off_p = offs[4];
    } /* admm.cpp:83: row vector * matrix, coefficient = redux over a contiguous column */
    return o;
}

/* 1 iff select_orders() has been verified bit-for-bit against the compiled reference for this shape
 * (tests/test_oracle_vs_ref.py, f32 and f64).  Other shapes run, but their parity is UNPINNED. */
int oracle_shape_pinned(int nx, int nu, int N)
{
    (void)N;
    return (nx == 12 && nu == 4) || (nx == 4 && nu == 1) || (nx == 32 && nu == 8);
}

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define T float
#define SFX(n) CAT(n, _f32)
#define PK 4
#define FABS fabsf
#include "tinympc_oracle_impl.h"
#undef T
#undef SFX
#undef PK
#undef FABS

#define T double
#define SFX(n) CAT(n, _f64)
#define PK 2
#define FABS fabs
#include "tinympc_oracle_impl.h"
#undef T
#undef SFX
#undef PK
#undef FABS

int oracle_solve_batch(const oracle_problem *in, int64_t B, const void *x0, const void *Xref,
                       int64_t xref_stride, const oracle_state *S, void *x_out, void *u_out,
                       int32_t *iter_out, int32_t *status_out, void *resid_out, int32_t nthreads)
{
    if (in->scalar_bytes == 4)
        return oracle_solve_batch_f32(in, B, x0, Xref, xref_stride, S, x_out, u_out, iter_out,
                                      status_out, resid_out, nthreads);
    if (in->scalar_bytes == 8)
        return oracle_solve_batch_f64(in, B, x0, Xref, xref_stride, S, x_out, u_out, iter_out,
                                      status_out, resid_out, nthreads);
    return -1;
}

int oracle_step(const oracle_problem *in, int32_t which, void *ws, int32_t iter)
{
    if (in->scalar_bytes == 4) return oracle_step_f32(in, which, ws, iter);
    if (in->scalar_bytes == 8) return oracle_step_f64(in, which, ws, iter);
    return -1;
}

/* batched plant step x1 = Adyn x0 + Bdyn u0; u0 of instance b = u0[b * u_stride .. + nu) */
int oracle_plant_step(const oracle_problem *in, int64_t B, const void *x0, const void *u0, int64_t u_stride, void *x1)
{
    if (in->scalar_bytes == 4) return oracle_plant_step_f32(in, B, x0, u0, u_stride, x1);
    if (in->scalar_bytes == 8) return oracle_plant_step_f64(in, B, x0, u0, u_stride, x1);
    return -1;
}
