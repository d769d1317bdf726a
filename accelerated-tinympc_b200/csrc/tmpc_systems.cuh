// Per-instance SYSTEMS (include/tmpc.h "tmpc_systems_*"): every instance of a batch brings its own
// model (Adyn, Bdyn, Q, R, rho), so its own cache.  Two pieces:
//
//   * precompute_kernel: the reference's cache recursion (/root/reference/src/tinympc/codegen.cpp:254-292 -- Riccati
//     fixed point on Q+rho, R+rho from P = rho*I, at most 1000 sweeps, stop when max|dKinf| < 1e-5, then Quu_inv and
//     AmBKt) for one instance per thread, in double, with the SAME operation order as the host tiny_precompute
//     (host/tiny_api.cpp: i-k-j products accumulated from zero, Gauss-Jordan with partial pivoting) and FMA
//     contraction off, so device and host caches are bit-identical.  It writes the instance's SysBlock.
//   * the PERSYS instances of the generic ADMM kernel (tmpc_kernel.cuh) read their coefficients from that block.
//
// Bounds, tolerances and iteration limits stay those of the ctx (shared).  Included by tmpc_api.cu.
#pragma once

namespace tmpc {

// C[R x C] = A[R x K] * B[K x C], row-major, accumulated from zero in i-k-j order (host mul())
template <int R, int K, int C> __device__ __forceinline__ void rmul(const double *A, const double *B, double *Cm)
{
    for (int i = 0; i < R * C; ++i) Cm[i] = 0.0;
    for (int i = 0; i < R; ++i)
        for (int k = 0; k < K; ++k) {
            const double a = A[i * K + k];
            for (int j = 0; j < C; ++j) Cm[i * C + j] = __dadd_rn(Cm[i * C + j], __dmul_rn(a, B[k * C + j]));
        }
}

// Gauss-Jordan with partial pivoting (host inverse()); A is destroyed.  false = singular
template <int N> __device__ bool rinverse(double *A, double *inv)
{
    for (int i = 0; i < N * N; ++i) inv[i] = 0.0;
    for (int i = 0; i < N; ++i) inv[i * N + i] = 1.0;
    for (int c = 0; c < N; ++c) {
        int piv = c;
        for (int r = c + 1; r < N; ++r)
            if (fabs(A[r * N + c]) > fabs(A[piv * N + c])) piv = r;
        if (A[piv * N + c] == 0.0) return false;
        if (piv != c)
            for (int j = 0; j < N; ++j) {
                double t = A[piv * N + j]; A[piv * N + j] = A[c * N + j]; A[c * N + j] = t;
                t = inv[piv * N + j]; inv[piv * N + j] = inv[c * N + j]; inv[c * N + j] = t;
            }
        const double d = __ddiv_rn(1.0, A[c * N + c]);
        for (int j = 0; j < N; ++j) { A[c * N + j] = __dmul_rn(A[c * N + j], d); inv[c * N + j] = __dmul_rn(inv[c * N + j], d); }
        for (int r = 0; r < N; ++r) {
            if (r == c) continue;
            const double f = A[r * N + c];
            if (f == 0.0) continue;
            for (int j = 0; j < N; ++j) {
                A[r * N + j] = __dsub_rn(A[r * N + j], __dmul_rn(f, A[c * N + j]));
                inv[r * N + j] = __dsub_rn(inv[r * N + j], __dmul_rn(f, inv[c * N + j]));
            }
        }
    }
    return true;
}

template <class T> struct PrecomputeArgs {
    long long batch;
    const T *Adyn, *Bdyn;   // [batch][nx*nx], [batch][nx*nu] column-major
    const T *Q, *R;         // [batch][nx], [batch][nu]
    const T *rho;           // [batch]
    int q_plus_rho;         // work.Q of the block: Q + rho (what tiny_codegen stores) or Q as given (the examples)
    T *blocks;              // [batch][SysBlock::STRIDE]
    int *sweeps;            // [batch] Riccati sweeps used (1000 = never converged; -1 = singular R + B'PB)
};

template <class T, int NX, int NU>
__global__ void __launch_bounds__(64) precompute_kernel(const PrecomputeArgs<T> a)
{
    using SB = SysBlock<NX, NU>;
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= a.batch) return;
    constexpr int n = NX, m = NU;
    double A[n * n], B[n * m], At[n * n], Bt[m * n];                 // row-major, like the host's to_rowmajor()
    double Ktp1[m * n], Ptp1[n * n], Kinf[m * n], Pinf[n * n];
    double BtP[m * n], S[m * m], Sinv[m * m], T1[m * n], BK[n * n], AmBK[n * n], T2[n * n];
    const T *Ac = a.Adyn + b * (n * n), *Bc = a.Bdyn + b * (n * m);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) { A[i * n + j] = (double)Ac[i + j * n]; At[j * n + i] = A[i * n + j]; }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) { B[i * m + j] = (double)Bc[i + j * n]; Bt[j * n + i] = B[i * m + j]; }
    const double rho = (double)a.rho[b];
    double q1[n], r1[m];
    for (int i = 0; i < n; ++i) q1[i] = __dadd_rn((double)a.Q[b * n + i], rho);          // codegen.cpp:255-258
    for (int i = 0; i < m; ++i) r1[i] = __dadd_rn((double)a.R[b * m + i], rho);
    for (int i = 0; i < m * n; ++i) { Ktp1[i] = 0.0; Kinf[i] = 0.0; }
    for (int i = 0; i < n * n; ++i) { Ptp1[i] = 0.0; Pinf[i] = 0.0; }
    for (int i = 0; i < n; ++i) Ptp1[i * n + i] = rho;
    int sweeps = 1000;
    bool singular = false;
    for (int it = 0; it < 1000; ++it) {                                                   // codegen.cpp:273-285
        rmul<m, n, n>(Bt, Ptp1, BtP);
        rmul<m, n, m>(BtP, B, S);
        for (int i = 0; i < m; ++i) S[i * m + i] = __dadd_rn(S[i * m + i], r1[i]);
        if (!rinverse<m>(S, Sinv)) { singular = true; break; }
        rmul<m, m, n>(Sinv, BtP, T1);
        rmul<m, n, n>(T1, A, Kinf);
        rmul<n, m, n>(B, Kinf, BK);
        for (int i = 0; i < n * n; ++i) AmBK[i] = __dsub_rn(A[i], BK[i]);
        rmul<n, n, n>(At, Ptp1, T2);
        rmul<n, n, n>(T2, AmBK, Pinf);
        for (int i = 0; i < n; ++i) Pinf[i * n + i] = __dadd_rn(Pinf[i * n + i], q1[i]);
        double dmax = 0.0;
        for (int i = 0; i < m * n; ++i) dmax = fmax(dmax, fabs(__dsub_rn(Kinf[i], Ktp1[i])));
        if (dmax < 1e-5) { sweeps = it + 1; break; }
        for (int i = 0; i < m * n; ++i) Ktp1[i] = Kinf[i];
        for (int i = 0; i < n * n; ++i) Ptp1[i] = Pinf[i];
    }
    if (!singular) {                                                                       // codegen.cpp:290-292
        rmul<m, n, n>(Bt, Pinf, BtP);
        rmul<m, n, m>(BtP, B, S);
        for (int i = 0; i < m; ++i) S[i * m + i] = __dadd_rn(S[i * m + i], r1[i]);
        if (!rinverse<m>(S, Sinv)) singular = true;
        rmul<n, m, n>(B, Kinf, BK);
        for (int i = 0; i < n * n; ++i) AmBK[i] = __dsub_rn(A[i], BK[i]);
    }
    T *blk = a.blocks + b * SB::STRIDE;
    for (int i = 0; i < SB::STRIDE; ++i) blk[i] = T(0);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) { const T v = (T)Kinf[i * n + j]; blk[SB::K + i + j * m] = v; blk[SB::Krm + i * n + j] = v; }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const T av = Ac[i + j * n], mv = (T)AmBK[j * n + i];                                 // AmBKt = (A - B K)^T
            blk[SB::A + i + j * n] = av; blk[SB::Arm + i * n + j] = av;
            blk[SB::M + i + j * n] = mv; blk[SB::Mrm + i * n + j] = mv;
            blk[SB::Pf + i + j * n] = (T)Pinf[i * n + j];
        }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) { const T v = Bc[i + j * n]; blk[SB::B + i + j * n] = v; blk[SB::Brm + i * m + j] = v; }
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) { const T v = (T)Sinv[i * m + j]; blk[SB::Qi + i + j * m] = v; blk[SB::Qirm + i * m + j] = v; }
    for (int i = 0; i < n; ++i) blk[SB::Qd + i] = a.q_plus_rho ? (T)q1[i] : a.Q[b * n + i];
    blk[SB::RHO] = a.rho[b];
    if (a.sweeps) a.sweeps[b] = singular ? -1 : sweeps;
}

}  // namespace tmpc

struct tmpc_systems_impl {
    tmpc_ctx_impl *c = nullptr;
    int64_t B = 0;
    void *blocks = nullptr;   // device [B][stride] in ctx dtype
    int *sweeps = nullptr;    // device [B]
    int stride = 0;
};
#define SYS(s) reinterpret_cast<tmpc_systems_impl *>(s)

namespace {

template <class T, int NX, int NU>
cudaError_t launch_precompute(tmpc_systems_impl *sy, const void *A, const void *Bm, const void *Q, const void *R, const void *rho,
                              int q_plus_rho, cudaStream_t s)
{
    tmpc::PrecomputeArgs<T> pa;
    pa.batch = sy->B; pa.Adyn = (const T *)A; pa.Bdyn = (const T *)Bm; pa.Q = (const T *)Q; pa.R = (const T *)R; pa.rho = (const T *)rho;
    pa.q_plus_rho = q_plus_rho; pa.blocks = (T *)sy->blocks; pa.sweeps = sy->sweeps;
    const unsigned blocks = (unsigned)((sy->B + 63) / 64);
    tmpc::precompute_kernel<T, NX, NU><<<blocks, 64, 0, s>>>(pa);
    return cudaGetLastError();
}

int sys_stride(int nx, int nu)
{
    if (nx == 12 && nu == 4) return tmpc::SysBlock<12, 4>::STRIDE;
    if (nx == 4 && nu == 1) return tmpc::SysBlock<4, 1>::STRIDE;
    return 0;
}

bool lookup_kernel_sys(int nx, int nu, int N, int dtype, int policy, bool warm, bool const_bounds, KernelInfo &out)
{
    // fp32 12/4/10 default: two lanes per instance, row pairs streamed from tensor memory (tmpc_kernel_sysp.cuh).  TMPC_KERNEL =
    // sys_thread: the same with one thread per instance (tmpc_kernel_sys.cuh); sys_rows: the first TMEM-resident kernel (row by
    // row); sys_global: coefficients re-read from the global block
    const char *e = getenv("TMPC_KERNEL");
    const bool global_coeffs = e && !strcmp(e, "sys_global"), rows = e && !strcmp(e, "sys_rows"), thr = e && !strcmp(e, "sys_thread");
    if (!global_coeffs && !rows && tmpc_dispatch::lookup_sys_pairs(nx, nu, N, dtype, policy, warm, const_bounds, thr ? 1 : 0, out)) return true;
    return tmpc_dispatch::lookup_sys(nx, nu, N, dtype, policy, warm, global_coeffs, out);
}

}  // namespace

extern "C" {

int tmpc_systems_precompute(tmpc_ctx *ctx, int64_t batch, const void *Adyn, const void *Bdyn, const void *Q, const void *R,
                            const void *rho, int32_t q_plus_rho, int32_t mem, tmpc_systems **out)
{
    if (!ctx || !out) return TMPC_ERR_INVALID;
    *out = nullptr;
    tmpc_ctx_impl *c = CTX(ctx);
    if (batch < 1 || !Adyn || !Bdyn || !Q || !R || !rho) return fail(c, TMPC_ERR_INVALID, "tmpc_systems_precompute: bad argument");
    const int stride = sys_stride(c->nx, c->nu);
    if (!stride) return fail(c, TMPC_ERR_UNSUPPORTED, "per-instance systems are compiled for nx/nu = 12/4 and 4/1");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const size_t es = esize(c);
    const int nx = c->nx, nu = c->nu;
    tmpc_systems_impl *sy = new tmpc_systems_impl;
    sy->c = c; sy->B = batch; sy->stride = stride;
    auto bail = [&](const std::string &m) { if (sy->blocks) cudaFree(sy->blocks); if (sy->sweeps) cudaFree(sy->sweeps); delete sy; return fail(c, TMPC_ERR_CUDA, m); };
    if (cudaMalloc(&sy->blocks, (size_t)batch * stride * es) != cudaSuccess) return bail("tmpc_systems_precompute: cudaMalloc failed");
    if (cudaMalloc((void **)&sy->sweeps, (size_t)batch * sizeof(int)) != cudaSuccess) return bail("tmpc_systems_precompute: cudaMalloc failed");
    const void *src[5] = {Adyn, Bdyn, Q, R, rho};
    const size_t per[5] = {(size_t)nx * nx * es, (size_t)nx * nu * es, (size_t)nx * es, (size_t)nu * es, es};
    void *dev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaStream_t s = c->stream;
    bool ok = true;
    if (mem != TMPC_MEM_DEVICE) {
        for (int k = 0; k < 5 && ok; ++k) {
            ok = cudaMalloc(&dev[k], per[k] * batch) == cudaSuccess &&
                 cudaMemcpyAsync(dev[k], src[k], per[k] * batch, cudaMemcpyHostToDevice, s) == cudaSuccess;
        }
    } else {
        for (int k = 0; k < 5; ++k) dev[k] = const_cast<void *>(src[k]);
    }
    cudaError_t e = cudaErrorUnknown;
    if (ok) {
        const bool f32 = c->dtype == TMPC_F32;
        if (nx == 12) e = f32 ? launch_precompute<float, 12, 4>(sy, dev[0], dev[1], dev[2], dev[3], dev[4], q_plus_rho, s)
                              : launch_precompute<double, 12, 4>(sy, dev[0], dev[1], dev[2], dev[3], dev[4], q_plus_rho, s);
        else e = f32 ? launch_precompute<float, 4, 1>(sy, dev[0], dev[1], dev[2], dev[3], dev[4], q_plus_rho, s)
                     : launch_precompute<double, 4, 1>(sy, dev[0], dev[1], dev[2], dev[3], dev[4], q_plus_rho, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    }
    if (mem != TMPC_MEM_DEVICE)
        for (int k = 0; k < 5; ++k) if (dev[k]) cudaFree(dev[k]);
    if (!ok || e != cudaSuccess) return bail(std::string("tmpc_systems_precompute: ") + (ok ? cudaGetErrorString(e) : "staging failed"));
    *out = reinterpret_cast<tmpc_systems *>(sy);
    return TMPC_OK;
}

int tmpc_systems_destroy(tmpc_systems *sp)
{
    if (!sp) return TMPC_OK;
    tmpc_systems_impl *sy = SYS(sp);
    cudaSetDevice(sy->c->device);
    cudaStreamSynchronize(sy->c->stream);
    if (sy->blocks) cudaFree(sy->blocks);
    if (sy->sweeps) cudaFree(sy->sweeps);
    delete sy;
    return TMPC_OK;
}

int tmpc_systems_get(tmpc_systems *sp, int32_t what, void *dst)
{
    if (!sp || !dst) return TMPC_ERR_INVALID;
    tmpc_systems_impl *sy = SYS(sp);
    tmpc_ctx_impl *c = sy->c;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const size_t es = esize(c);
    const int nx = c->nx, nu = c->nu;
    if (what == TMPC_SYS_SWEEPS) {
        CUDA_TRY(c, cudaMemcpy(dst, sy->sweeps, (size_t)sy->B * sizeof(int), cudaMemcpyDeviceToHost));
        return TMPC_OK;
    }
    int off = 0, len = 0;
    int oK, oA, oB, oQi, oM, oPf, oQd, oRho;
    if (nx == 12) { using SB = tmpc::SysBlock<12, 4>; oK = SB::K; oA = SB::A; oB = SB::B; oQi = SB::Qi; oM = SB::M; oPf = SB::Pf; oQd = SB::Qd; oRho = SB::RHO; }
    else { using SB = tmpc::SysBlock<4, 1>; oK = SB::K; oA = SB::A; oB = SB::B; oQi = SB::Qi; oM = SB::M; oPf = SB::Pf; oQd = SB::Qd; oRho = SB::RHO; }
    switch (what) {
    case TMPC_SYS_KINF: off = oK; len = nu * nx; break;
    case TMPC_SYS_PINF: off = oPf; len = nx * nx; break;
    case TMPC_SYS_QUU_INV: off = oQi; len = nu * nu; break;
    case TMPC_SYS_AMBKT: off = oM; len = nx * nx; break;
    case TMPC_SYS_ADYN: off = oA; len = nx * nx; break;
    case TMPC_SYS_BDYN: off = oB; len = nx * nu; break;
    case TMPC_SYS_Q: off = oQd; len = nx; break;
    case TMPC_SYS_RHO: off = oRho; len = 1; break;
    default: return fail(c, TMPC_ERR_INVALID, "bad tmpc_systems_get selector");
    }
    CUDA_TRY(c, cudaMemcpy2D(dst, (size_t)len * es, (const char *)sy->blocks + (size_t)off * es, (size_t)sy->stride * es, (size_t)len * es,
                             (size_t)sy->B, cudaMemcpyDeviceToHost));
    return TMPC_OK;
}

int tmpc_solve_systems(tmpc_ctx *ctx, const tmpc_solve_args *a, const tmpc_systems *sp)
{
    if (!ctx || !a || !sp) return TMPC_ERR_INVALID;
    tmpc_ctx_impl *c = CTX(ctx);
    const tmpc_systems_impl *sy = reinterpret_cast<const tmpc_systems_impl *>(sp);
    if (sy->c != c) return fail(c, TMPC_ERR_INVALID, "systems object belongs to another ctx");
    if (!c->has_model) return fail(c, TMPC_ERR_STATE, "tmpc_set_model has not been called (bounds and settings come from the ctx)");
    if (a->batch != sy->B) return fail(c, TMPC_ERR_INVALID, "batch differs from the systems object");
    if (a->mem != TMPC_MEM_DEVICE) return fail(c, TMPC_ERR_INVALID, "tmpc_solve_systems takes device buffers (TMPC_MEM_DEVICE)");
    if (!a->x0 || !a->Xref) return fail(c, TMPC_ERR_INVALID, "x0 / Xref must not be NULL");
    const bool warm = a->warm != nullptr;
    if (warm && (!a->warm->d || !a->warm->y || !a->warm->g || !a->warm->v || !a->warm->z))
        return fail(c, TMPC_ERR_INVALID, "warm state needs all of d, y, g, v, z");
    CUDA_TRY(c, cudaSetDevice(c->device));
    KernelInfo ki;
    if (!lookup_kernel_sys(c->nx, c->nu, c->N, c->dtype, c->policy, warm, c->const_bounds, ki)) return fail(c, TMPC_ERR_UNSUPPORTED, "no per-instance-systems kernel for this shape");
    DevArgs da{};
    da.batch = a->batch; da.x0 = a->x0; da.Xref = a->Xref;
    da.xref_stride = a->xref_shared ? 0 : (long long)c->nx * c->N;
    if (warm) { da.wd = a->warm->d; da.wy = a->warm->y; da.wg = a->warm->g; da.wv = a->warm->v; da.wz = a->warm->z; }
    da.x = a->x; da.u = a->u; da.iter = a->iter; da.status = a->status; da.resid = a->resid; da.u0 = a->u0;
    da.sys = sy->blocks;
    cudaStream_t s = a->stream ? (cudaStream_t)a->stream : c->stream;
    c->stats.instances = a->batch;
    c->stats.iterations = c->stats.solved = c->stats.trips = 0;
    c->stats.launches = 0;
    // longest-expected-first schedule with each instance's own Kinf (tmpc_api.cu lpt_prepare)
    bool ev0_done = false;
    c->lpt_used = 0;
    {
        const int rc = order_after_previous(c, s);   // the scheduling buffers may still be read by the previous launch
        if (rc != TMPC_OK) return rc;
    }
    if (c->nx == 12 && c->nu == 4 && lpt_wanted(c, ki, da)) {
        using SB = tmpc::SysBlock<12, 4>;
        CUDA_TRY(c, cudaEventRecord(c->ev0, s));
        ev0_done = true;
        const char *kb = (const char *)sy->blocks + (size_t)SB::K * esize(c);
        int rc0 = lpt_prepare(c, da, s, kb, (long long)sy->stride);
        if (rc0 != TMPC_OK) return rc0;
        c->lpt_used = 1;
    }
    int rc = launch_kernel_info(c, ki, da, s, true, ev0_done);
    if (rc == TMPC_OK) { c->stats_pending = true; c->stats.pattern = 0; }
    return rc;
}

}  // extern "C"
