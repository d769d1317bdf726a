/* TEST INFRASTRUCTURE ONLY -- see tinympc_oracle.c.  Included once per scalar type with
 *   T      scalar type            SFX(name)  name##_f32 / name##_f64         PK  SSE packet width 16/sizeof(T)
 *
 * Plain-C restatement of the reference's cached-Riccati ADMM iteration, written against
 * /root/reference/src/tinympc/admm.cpp (line numbers cited per function) with the floating-point
 * EVALUATION ORDER that the reference's vendored Eigen 3.4.90 produces in an "-O3, no -m flags"
 * (SSE2, no FMA) build -- SURVEY.md appendix A.2.  All products a*b are individually rounded, then
 * summed in one of these orders:
 *
 *   ORD_SEQ      acc = e0; acc = e_k + acc, k = 1..K-1          (Eigen lazy-product packet path over the
 *                rows of a column-major lhs: ProductEvaluators.h etor_product_packet_impl, pmadd = mul+add)
 *   ORD_VECREDUX completely unrolled *vectorised* redux of cwiseProduct: lanes l=0..PK-1 each reduce the
 *                packets j of e_{j*PK+l} with the recursive half split tree(s,len) = tree(s,len/2) +
 *                tree(s+len/2, len-len/2); lanes combine as (l0+l2)+(l1+l3) [float] / l0+l1 [double];
 *                a remainder K%PK is reduced by the scalar tree and added last (Redux.h redux_vec_unroller,
 *                redux_impl<LinearVectorizedTraversal,CompleteUnrolling>; SSE predux)
 *   ORD_TREE     completely unrolled *scalar* redux: the same half split over the K products
 *                (Redux.h redux_novec_unroller)
 *   ORD_VECLOOP  (K too large for complete unrolling, Redux.h redux_impl<LinearVectorizedTraversal,
 *                NoUnrolling>): two packet accumulators over alternating packets, then a third packet if
 *                one is left, predux, then scalar tail
 *   ORD_GEMV_COL / ORD_GEMV_ROW   Eigen's general_matrix_vector_product kernels (dimension >= 8), see below
 *
 * Which order each of the 8 products uses depends on the SHAPE (Eigen dispatches at compile time);
 * select_orders() encodes that dispatch and tests/test_oracle_vs_ref.py pins every shape this repo
 * ships against the compiled reference (oracle/_ref) bit for bit.
 */

typedef struct {
    int nx, nu, N;
    const T *Kinf, *Pinf, *Quu_inv, *AmBKt, *Adyn, *Bdyn, *Q; /* column-major */
    const T *x_min, *x_max, *u_min, *u_max;                   /* [N][nx], [N-1][nu] */
    T rho, abs_pri_tol, abs_dua_tol;
    int max_iter, check_termination, en_state_bound, en_input_bound;
    orders_t ord;
} SFX(prob);

typedef struct {
    T *x, *u, *q, *r, *p, *d, *v, *vnew, *z, *znew, *g, *y, *Xref;
    T pri_x, dua_x, pri_u, dua_u;
    int iter, status;
    T *e; /* scratch for K products */
} SFX(work);

static T SFX(tree)(const T *e, int len)
{
    if (len == 1) return e[0];
    int h = len / 2;
    T a = SFX(tree)(e, h);
    T b = SFX(tree)(e + h, len - h);
    return a + b;
}

/* tree over packets [s, s+len) for lane l */
static T SFX(ptree)(const T *e, int s, int len, int l)
{
    if (len == 1) return e[s * PK + l];
    int h = len / 2;
    T a = SFX(ptree)(e, s, h, l);
    T b = SFX(ptree)(e, s + h, len - h, l);
    return a + b;
}

static T SFX(predux)(const T *lane)
{
#if PK == 4
    return (lane[0] + lane[2]) + (lane[1] + lane[3]);
#else
    return lane[0] + lane[1];
#endif
}

static T SFX(reduce)(const T *e, int K, int order)
{
    T lane[PK];
    switch (order) {
    case ORD_SEQ: {
        T acc = e[0];
        for (int k = 1; k < K; ++k) acc = e[k] + acc;
        return acc;
    }
    case ORD_TREE:
        return SFX(tree)(e, K);
    case ORD_VECREDUX: {
        int np = K / PK;
        if (np == 0) return SFX(tree)(e, K);
        for (int l = 0; l < PK; ++l) lane[l] = SFX(ptree)(e, 0, np, l);
        T res = SFX(predux)(lane);
        if (np * PK != K) res = res + SFX(tree)(e + np * PK, K - np * PK);
        return res;
    }
    case ORD_VECLOOP: {
        /* Redux.h redux_impl<Func,Evaluator,LinearVectorizedTraversal,NoUnrolling>::run with
         * alignedStart = 0 (unaligned packet loads are allowed on SSE): */
        int alignedSize2 = (K / (2 * PK)) * (2 * PK);
        int alignedSize = (K / PK) * PK;
        T res;
        if (alignedSize) {
            T p0[PK], p1[PK];
            for (int l = 0; l < PK; ++l) p0[l] = e[l];
            if (alignedSize > PK) {
                for (int l = 0; l < PK; ++l) p1[l] = e[PK + l];
                for (int i = 2 * PK; i < alignedSize2; i += 2 * PK)
                    for (int l = 0; l < PK; ++l) {
                        p0[l] = p0[l] + e[i + l];
                        p1[l] = p1[l] + e[i + PK + l];
                    }
                for (int l = 0; l < PK; ++l) p0[l] = p0[l] + p1[l];
                if (alignedSize > alignedSize2)
                    for (int l = 0; l < PK; ++l) p0[l] = p0[l] + e[alignedSize2 + l];
            }
            res = SFX(predux)(p0);
            for (int i = alignedSize; i < K; ++i) res = res + e[i];
        } else {
            res = e[0];
            for (int i = 1; i < K; ++i) res = res + e[i];
        }
        return res;
    }
    case ORD_GEMV_ROW: {
        /* general_matrix_vector_product<RowMajor> (GeneralMatrixVector.h:329-517) into a zeroed temporary
         * with alpha = 1: per row, lane l accumulates packets sequentially from 0 (0 + e = e), predux,
         * then the scalar tail is added one by one; res = 0 + 1*cc. */
        int full = (K / PK) * PK;
        T res;
        if (full) {
            for (int l = 0; l < PK; ++l) lane[l] = e[l];
            for (int j = PK; j < full; j += PK)
                for (int l = 0; l < PK; ++l) lane[l] = e[j + l] + lane[l];
            res = SFX(predux)(lane);
        } else {
            res = (T)0;
        }
        for (int j = full; j < K; ++j) res = res + e[j];
        return res;
    }
    default:
        return (T)NAN;
    }
}

/* out[r] = sum_k M[r + k*ld_k ...] : generic strided mat-vec with a named order.
 * M(r,k) = m[r*rs + k*ks]; vec x[k]. */
static void SFX(matvec2)(T *out, const T *m, int rs, int ks, const T *x, int R, int K, int order, int lo, int hi,
                         int tail_order, T *e)
{
    if (order == ORD_GEMV_COL) order = ORD_SEQ;
    for (int r = 0; r < R; ++r) {
        for (int k = 0; k < K; ++k) e[k] = m[r * rs + k * ks] * x[k];
        out[r] = SFX(reduce)(e, K, (r >= lo && r < hi) ? order : tail_order);
    }
}

static void SFX(matvec)(T *out, const T *m, int rs, int ks, const T *x, int R, int K, int order, T *e)
{
    if (order == ORD_GEMV_COL) {
        /* general_matrix_vector_product<ColMajor> (GeneralMatrixVector.h): dest is a zeroed temporary,
         * all columns in one block (cols < 128); per row: c = 0; c = e_k + c ...; res = c*alpha + res with
         * alpha = 1, res = 0  ==> identical to a sequential sum seeded with (0 + e_0) = e_0. */
        order = ORD_SEQ;
    }
    for (int r = 0; r < R; ++r) {
        for (int k = 0; k < K; ++k) e[k] = m[r * rs + k * ks] * x[k];
        out[r] = SFX(reduce)(e, K, order);
    }
}

/* admm.cpp:27-37 */
static void SFX(forward_pass)(const SFX(prob) *P, SFX(work) *w)
{
    const int n = P->nx, m = P->nu, N = P->N;
    T Kx[64], Ax[64], Bu[64];
    for (int i = 0; i < N - 1; ++i) {
        const T *xi = w->x + i * n;
        T *ui = w->u + i * m;
        int lo, hi;
        head_range(&P->ord, m, P->ord.head_Kx, P->ord.rt_u, P->ord.off_u, i, &lo, &hi);
        SFX(matvec2)(Kx, P->Kinf, 1, m, xi, m, n, P->ord.Kx, lo, hi, P->ord.tail_x, w->e); /* :31 */
        for (int r = 0; r < m; ++r) ui[r] = (-Kx[r]) - w->d[i * m + r];
        head_range(&P->ord, n, P->ord.head_Ax, P->ord.rt_x, P->ord.off_x, i + 1, &lo, &hi);
        SFX(matvec2)(Ax, P->Adyn, 1, n, xi, n, n, P->ord.Ax, lo, hi, P->ord.tail_x, w->e); /* :35 */
        SFX(matvec2)(Bu, P->Bdyn, 1, n, ui, n, m, P->ord.Bu, lo, hi, P->ord.tail_u, w->e);
        for (int r = 0; r < n; ++r) w->x[(i + 1) * n + r] = Ax[r] + Bu[r];
    }
}

static T SFX(clampv)(T lo, T hi, T v)
{
    /* x_max.cwiseMin(x_min.cwiseMax(v)) with SSE min/max semantics on finite data */
    T a = (lo > v) ? lo : v; /* pmax(lo, v) = _mm_max_ps(v?..) -- value-identical for finite inputs */
    return (hi < a) ? hi : a;
}

/* admm.cpp:45-61 */
static void SFX(update_slack)(const SFX(prob) *P, SFX(work) *w)
{
    const int nxn = P->nx * P->N, nun = P->nu * (P->N - 1);
    for (int k = 0; k < nun; ++k) w->znew[k] = w->u[k] + w->y[k];
    for (int k = 0; k < nxn; ++k) w->vnew[k] = w->x[k] + w->g[k];
    if (P->en_input_bound)
        for (int k = 0; k < nun; ++k) w->znew[k] = SFX(clampv)(P->u_min[k], P->u_max[k], w->znew[k]);
    if (P->en_state_bound)
        for (int k = 0; k < nxn; ++k) w->vnew[k] = SFX(clampv)(P->x_min[k], P->x_max[k], w->vnew[k]);
}

/* admm.cpp:67-71 */
static void SFX(update_dual)(const SFX(prob) *P, SFX(work) *w)
{
    const int nxn = P->nx * P->N, nun = P->nu * (P->N - 1);
    for (int k = 0; k < nun; ++k) w->y[k] = (w->y[k] + w->u[k]) - w->znew[k];
    for (int k = 0; k < nxn; ++k) w->g[k] = (w->g[k] + w->x[k]) - w->vnew[k];
}

/* admm.cpp:77-85 */
static void SFX(update_linear_cost)(const SFX(prob) *P, SFX(work) *w)
{
    const int n = P->nx, N = P->N, nxn = n * N, nun = P->nu * (N - 1);
    const T nrho = -P->rho;
    T xp[64];
    for (int k = 0; k < nun; ++k) w->r[k] = nrho * (w->znew[k] - w->y[k]);              /* :80 */
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < n; ++j) w->q[i * n + j] = -(w->Xref[i * n + j] * P->Q[j]);  /* :81 */
    for (int k = 0; k < nxn; ++k) w->q[k] = w->q[k] - P->rho * (w->vnew[k] - w->g[k]);  /* :82 */
    /* :83  p_N = -(Xref_N^T * Pinf): row vector times matrix; output j = sum_k Xref[k]*Pinf(k,j) */
    SFX(matvec)(xp, P->Pinf, n, 1, w->Xref + (N - 1) * n, n, n, P->ord.XtP, w->e);
    for (int j = 0; j < n; ++j) w->p[(N - 1) * n + j] = -xp[j];
    for (int j = 0; j < n; ++j)                                                           /* :84 */
        w->p[(N - 1) * n + j] =
            w->p[(N - 1) * n + j] - P->rho * (w->vnew[(N - 1) * n + j] - w->g[(N - 1) * n + j]);
}

static T SFX(maxabsdiff)(const T *a, const T *b, int len)
{
    T m = FABS(a[0] - b[0]);
    for (int k = 1; k < len; ++k) {
        T t = FABS(a[k] - b[k]);
        if (t > m) m = t;
    }
    return m;
}

/* admm.cpp:91-109 */
static int SFX(termination_condition)(const SFX(prob) *P, SFX(work) *w)
{
    const int nxn = P->nx * P->N, nun = P->nu * (P->N - 1);
    if (w->iter % P->check_termination == 0) {
        w->pri_x = SFX(maxabsdiff)(w->x, w->vnew, nxn);
        w->dua_x = SFX(maxabsdiff)(w->v, w->vnew, nxn) * P->rho;
        w->pri_u = SFX(maxabsdiff)(w->u, w->znew, nun);
        w->dua_u = SFX(maxabsdiff)(w->z, w->znew, nun) * P->rho;
        if (w->pri_x < P->abs_pri_tol && w->pri_u < P->abs_pri_tol && w->dua_x < P->abs_dua_tol &&
            w->dua_u < P->abs_dua_tol)
            return 1;
    }
    return 0;
}

/* admm.cpp:15-22 */
static void SFX(backward_pass_grad)(const SFX(prob) *P, SFX(work) *w)
{
    const int n = P->nx, m = P->nu, N = P->N;
    T s[64], Mp[64], Ktr[64];
    for (int i = N - 2; i >= 0; --i) {
        const T *pn = w->p + (i + 1) * n;
        const T *ri = w->r + i * m;
        /* :19  d_i = Quu_inv * (B^T p_{i+1} + r_i);  (B^T)(r,k) = B[k + r*n] */
        SFX(matvec)(s, P->Bdyn, n, 1, pn, m, n, P->ord.Btp, w->e);
        for (int r = 0; r < m; ++r) s[r] = s[r] + ri[r];
        int lo, hi;
        head_range(&P->ord, m, P->ord.head_Qs, 0, 0, i, &lo, &hi);
        SFX(matvec2)(w->d + i * m, P->Quu_inv, 1, m, s, m, m, P->ord.Qs, lo, hi, P->ord.tail_u, w->e);
        /* :20  p_i = q_i + AmBKt p_{i+1} - Kinf^T r_i;  (K^T)(r,k) = K[k + r*m] */
        head_range(&P->ord, n, P->ord.head_Mp, P->ord.rt_p, P->ord.off_p, i, &lo, &hi);
        SFX(matvec2)(Mp, P->AmBKt, 1, n, pn, n, n, P->ord.Mp, lo, hi, P->ord.tail_x, w->e);
        SFX(matvec)(Ktr, P->Kinf, m, 1, ri, n, m, P->ord.Ktr, w->e);
        for (int r = 0; r < n; ++r) w->p[i * n + r] = (w->q[i * n + r] + Mp[r]) - Ktr[r];
    }
}

/* admm.cpp:111-152 */
static int SFX(tiny_solve)(const SFX(prob) *P, SFX(work) *w)
{
    const int nxn = P->nx * P->N, nun = P->nu * (P->N - 1);
    w->status = 11;
    w->iter = 1;
    for (int i = 0; i < P->max_iter; ++i) {
        w->iter = i + 1;
        SFX(forward_pass)(P, w);
        SFX(update_slack)(P, w);
        SFX(update_dual)(P, w);
        SFX(update_linear_cost)(P, w);
        if (SFX(termination_condition)(P, w)) {
            w->status = 1;
            return 0;
        }
        memcpy(w->v, w->vnew, sizeof(T) * nxn);
        memcpy(w->z, w->znew, sizeof(T) * nun);
        SFX(backward_pass_grad)(P, w);
    }
    return 1;
}

static void SFX(load_prob)(SFX(prob) *P, const oracle_problem *in)
{
    P->nx = in->nx; P->nu = in->nu; P->N = in->N;
    P->Kinf = (const T *)in->Kinf; P->Pinf = (const T *)in->Pinf; P->Quu_inv = (const T *)in->Quu_inv;
    P->AmBKt = (const T *)in->AmBKt; P->Adyn = (const T *)in->Adyn; P->Bdyn = (const T *)in->Bdyn;
    P->Q = (const T *)in->Q;
    P->x_min = (const T *)in->x_min; P->x_max = (const T *)in->x_max;
    P->u_min = (const T *)in->u_min; P->u_max = (const T *)in->u_max;
    P->rho = (T)in->rho; P->abs_pri_tol = (T)in->abs_pri_tol; P->abs_dua_tol = (T)in->abs_dua_tol;
    P->max_iter = in->max_iter; P->check_termination = in->check_termination;
    P->en_state_bound = in->en_state_bound; P->en_input_bound = in->en_input_bound;
    P->ord = select_orders(in->nx, in->nu, in->N, (int)sizeof(T));
}

static void SFX(alloc_work)(SFX(work) *w, int nxn, int nun, int kmax)
{
    T *buf = (T *)calloc((size_t)(7 * nxn + 6 * nun + kmax), sizeof(T));
    w->x = buf; w->q = w->x + nxn; w->p = w->q + nxn; w->v = w->p + nxn; w->vnew = w->v + nxn;
    w->g = w->vnew + nxn; w->Xref = w->g + nxn;
    w->u = w->Xref + nxn; w->r = w->u + nun; w->d = w->r + nun; w->z = w->d + nun;
    w->znew = w->z + nun; w->y = w->znew + nun; w->e = w->y + nun;
}

typedef struct {
    const oracle_problem *in;
    int64_t b0, b1;
    const T *x0, *Xref;
    int64_t xref_stride;
    const oracle_state *S;
    T *x_out, *u_out, *resid_out;
    int32_t *iter_out, *status_out;
} SFX(job);

#define GETV(dst, base, b, len) do { if (base) memcpy(dst, (const T *)(base) + (b) * (len), sizeof(T) * (len)); else memset(dst, 0, sizeof(T) * (len)); } while (0)
#define PUTV(src, base, b, len) do { if (base) memcpy((T *)(base) + (b) * (len), src, sizeof(T) * (len)); } while (0)

static void *SFX(run_range)(void *arg)
{
    SFX(job) *J = (SFX(job) *)arg;
    SFX(prob) P;
    SFX(load_prob)(&P, J->in);
    const int n = P.nx, nxn = P.nx * P.N, nun = P.nu * (P.N - 1);
    SFX(work) w;
    SFX(alloc_work)(&w, nxn, nun, 64);
    const oracle_state *S = J->S;
    for (int64_t b = J->b0; b < J->b1; ++b) {
        memset(w.x, 0, sizeof(T) * nxn); memset(w.u, 0, sizeof(T) * nun);
        memset(w.q, 0, sizeof(T) * nxn); memset(w.r, 0, sizeof(T) * nun);
        memset(w.p, 0, sizeof(T) * nxn); memset(w.vnew, 0, sizeof(T) * nxn);
        memset(w.znew, 0, sizeof(T) * nun);
        GETV(w.d, S ? S->d : NULL, b, nun); GETV(w.y, S ? S->y : NULL, b, nun);
        GETV(w.g, S ? S->g : NULL, b, nxn); GETV(w.v, S ? S->v : NULL, b, nxn);
        GETV(w.z, S ? S->z : NULL, b, nun);
        memcpy(w.Xref, J->Xref + b * J->xref_stride, sizeof(T) * nxn);
        memcpy(w.x, J->x0 + b * n, sizeof(T) * n);
        SFX(tiny_solve)(&P, &w);
        PUTV(w.x, J->x_out, b, nxn); PUTV(w.u, J->u_out, b, nun);
        if (J->iter_out) J->iter_out[b] = w.iter;
        if (J->status_out) J->status_out[b] = w.status;
        if (J->resid_out) {
            J->resid_out[4 * b + 0] = w.pri_x; J->resid_out[4 * b + 1] = w.dua_x;
            J->resid_out[4 * b + 2] = w.pri_u; J->resid_out[4 * b + 3] = w.dua_u;
        }
        if (S) {
            PUTV(w.d, S->d, b, nun); PUTV(w.y, S->y, b, nun); PUTV(w.g, S->g, b, nxn);
            PUTV(w.v, S->v, b, nxn); PUTV(w.z, S->z, b, nun);
            PUTV(w.vnew, S->vnew, b, nxn); PUTV(w.znew, S->znew, b, nun);
            PUTV(w.q, S->q, b, nxn); PUTV(w.r, S->r, b, nun); PUTV(w.p, S->p, b, nxn);
        }
    }
    free(w.x);
    return NULL;
}

int SFX(oracle_solve_batch)(const oracle_problem *in, int64_t B, const void *x0, const void *Xref,
                            int64_t xref_stride, const oracle_state *S, void *x_out, void *u_out,
                            int32_t *iter_out, int32_t *status_out, void *resid_out, int32_t nthreads)
{
    if (in->nx > 64 || in->nu > 64 || in->check_termination < 1) return -1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > B) nthreads = (int32_t)(B > 0 ? B : 1);
    pthread_t th[256];
    SFX(job) jobs[256];
    for (int t = 0; t < nthreads; ++t) {
        SFX(job) j = {in, B * t / nthreads, B * (t + 1) / nthreads, (const T *)x0, (const T *)Xref,
                      xref_stride, S, (T *)x_out, (T *)u_out, (T *)resid_out, iter_out, status_out};
        jobs[t] = j;
        if (nthreads == 1) SFX(run_range)(&jobs[t]);
        else pthread_create(&th[t], NULL, SFX(run_range), &jobs[t]);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    return 0;
}

/* same workspace image and `which` codes as ref_step in ref_harness.cpp */
int SFX(oracle_step)(const oracle_problem *in, int32_t which, void *ws_, int32_t iter)
{
    SFX(prob) P;
    SFX(load_prob)(&P, in);
    const int nxn = P.nx * P.N, nun = P.nu * (P.N - 1);
    T *ws = (T *)ws_;
    T escratch[64];
    SFX(work) w;
    w.x = ws; w.u = w.x + nxn; w.q = w.u + nun; w.r = w.q + nxn; w.p = w.r + nun; w.d = w.p + nxn;
    w.v = w.d + nun; w.vnew = w.v + nxn; w.z = w.vnew + nxn; w.znew = w.z + nun; w.g = w.znew + nun;
    w.y = w.g + nxn; w.Xref = w.y + nun;
    T *res = w.Xref + nxn;
    w.e = escratch;
    w.iter = iter;
    w.pri_x = res[0]; w.dua_x = res[1]; w.pri_u = res[2]; w.dua_u = res[3];
    int rc = 0;
    switch (which) {
    case 0: SFX(forward_pass)(&P, &w); break;
    case 1: SFX(update_slack)(&P, &w); break;
    case 2: SFX(update_dual)(&P, &w); break;
    case 3: SFX(update_linear_cost)(&P, &w); break;
    case 4: rc = SFX(termination_condition)(&P, &w); break;
    case 5: SFX(backward_pass_grad)(&P, &w); break;
    default: return -1;
    }
    res[0] = w.pri_x; res[1] = w.dua_x; res[2] = w.pri_u; res[3] = w.dua_u;
    return rc;
}

/* The plant step of the reference's closed-loop examples (quadrotor_hovering.cpp:108, quadrotor_tracking.cpp:112):
 *     x1 = work.Adyn * x0 + work.Bdyn * work.u.col(0)
 * evaluates exactly like one stage of forward_pass (admm.cpp:35): (Adyn x0) and (Bdyn u0) each in their product
 * order, then one add per row.  Pinned against oracle/_ref's ref_plant_step by tests/test_oracle_vs_ref.py. */
int SFX(oracle_plant_step)(const oracle_problem *in, int64_t B, const void *x0_, const void *u0_, int64_t u_stride,
                           void *x1_)
{
    SFX(prob) P;
    SFX(load_prob)(&P, in);
    const int n = P.nx, m = P.nu;
    if (n > 64 || m > 64) return -1;
    T Ax[64], Bu[64], e[64];
    for (int64_t b = 0; b < B; ++b) {
        const T *x0 = (const T *)x0_ + b * n, *u0 = (const T *)u0_ + b * u_stride;
        T *x1 = (T *)x1_ + b * n;
        /* work.Adyn * x0 and work.Bdyn * u are REGULAR products here (not lazyProduct): GeneralProduct.h
         * product_type_selector sends rows >= 8 && depth >= 8 to the column-major GEMV (sequential, every row, into a
         * temporary); smaller ones stay coefficient-based inside the sum and follow the packet / scalar rows of the
         * assignment to a 16-byte aligned x1: rows [0, n/pk*pk) sequential, the rest the scalar tree. */
        const int ax_all = n >= 8, bu_all = n >= 8 && m >= 8;
        const int head = P.ord.head_Ax < 0 ? n : P.ord.head_Ax;
        SFX(matvec2)(Ax, P.Adyn, 1, n, x0, n, n, P.ord.Ax, 0, ax_all ? n : head, P.ord.tail_x, e);
        SFX(matvec2)(Bu, P.Bdyn, 1, n, u0, n, m, P.ord.Bu, 0, bu_all ? n : head, P.ord.tail_u, e);
        for (int r = 0; r < n; ++r) x1[r] = Ax[r] + Bu[r];
    }
    return 0;
}

#undef GETV
#undef PUTV
