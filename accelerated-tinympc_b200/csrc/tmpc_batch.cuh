// Device-resident batch of workspaces (include/tmpc.h "tmpc_batch_*"): the reference's wrapper API
// (/root/reference/src/tinympc/tiny_wrapper.cpp:5-176: set_x0 / set_xref / reset_dual_variables / call_tiny_solve /
// get_x / get_u on ONE global workspace) with a leading batch dimension, and the closed loop of its examples
// (examples/quadrotor_hovering.cpp:90-114, quadrotor_tracking.cpp:93-118) run entirely on the device:
//
//     per MPC step:  Xref window from a table  ->  y = g = 0  ->  tiny_solve (warm: d, v, z carried)  ->  x0 <- A x0 + B u_0
//
// as three kernels per step on one stream (window gather, the persistent ADMM kernel, plant step) with no host
// round trip, no H2D/D2H and the warm-start state {d,y,g,v,z} never leaving HBM.  Included by tmpc_api.cu.
#pragma once

namespace tmpc {

// Xref[b][i][:] = table[w0 + i][:], w0 = min(start[b] + step, rows - N)   (tracking.cpp:101: Xref_total.block(0, k); the
// window stops moving at the end of the table, which the reference's loop bound k < NTOTAL-NHORIZON-1 never reaches)
template <class T>
__global__ void xref_window_kernel(long long batch, int nx, int N, const T *table, long long rows, const int *start, int step, T *Xref)
{
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long per = (long long)N * nx;
    if (idx >= batch * per) return;
    const long long b = idx / per;
    const int i = (int)((idx - b * per) / nx), j = (int)(idx % nx);
    long long w0 = (start ? start[b] : 0) + step;
    if (w0 > rows - N) w0 = rows - N;
    Xref[idx] = table[(w0 + i) * nx + j];
}

// The examples' plant step x1 = Adyn * x0 + Bdyn * u.col(0) (quadrotor_hovering.cpp:108).  It evaluates like one stage of
// forward_pass (admm.cpp:35): both products in their own order, one add per row -- pinned against the compiled
// reference by tests/test_oracle_vs_ref.py::test_plant_step.  Also records the step's u0 / iter / status histories.
template <class T, int NX, int NU, int NH, bool FAST>
__global__ void plant_kernel(const __grid_constant__ Model<T, NX, NU, NH> P, long long batch, T *x0, const T *u, long long u_stride, T *x_next_hist,
                             T *u0_hist, const int *iter, const int *status, int *iter_hist, int *status_hist)
{
    using N = Num<T>;
    using O = Orders<T, NX, NU>;
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= batch) return;
    T x[NX], u0[NU];
#pragma unroll
    for (int j = 0; j < NX; ++j) x[j] = x0[b * NX + j];
#pragma unroll
    for (int j = 0; j < NU; ++j) u0[j] = u[b * u_stride + j];   // u_stride: nu (N-1) = row 0 of the trajectory, nu = the u0-only buffer
#pragma unroll
    for (int r = 0; r < NX; ++r) {
        T v;
        if constexpr (FAST) {
            v = dot<T, O::Ax, NX, true>([&](int k) { return P.A[r + k * NX]; }, [&](int k) { return x[k]; });
#pragma unroll
            for (int k = 0; k < NU; ++k) v = N::fma(P.B[r + k * NX], u0[k], v);
        } else {
            const T ax = dot<T, O::Ax, NX, false>([&](int k) { return P.A[r + k * NX]; }, [&](int k) { return x[k]; });
            const T bu = dot<T, O::Bu, NU, false>([&](int k) { return P.B[r + k * NX]; }, [&](int k) { return u0[k]; });
            v = N::add(ax, bu);
        }
        x0[b * NX + r] = v;
        if (x_next_hist) x_next_hist[b * NX + r] = v;
    }
    if (u0_hist) {
#pragma unroll
        for (int j = 0; j < NU; ++j) u0_hist[b * NU + j] = u0[j];
    }
    if (iter_hist) iter_hist[b] = iter[b];
    if (status_hist) status_hist[b] = status[b];
}

}  // namespace tmpc

struct tmpc_batch_impl {
    tmpc_ctx_impl *c = nullptr;
    int64_t B = 0;
    char *base = nullptr;          // one device allocation
    void *x0 = nullptr, *xref = nullptr, *d = nullptr, *y = nullptr, *g = nullptr, *v = nullptr, *z = nullptr;
    void *x = nullptr, *u = nullptr, *resid = nullptr;
    int *iter = nullptr, *status = nullptr;
    bool xref_shared = true;
    void *table = nullptr;         // [rows][nx]
    int64_t table_rows = 0;
    int *start = nullptr;          // [B] or null
    int64_t steps_done = 0;        // window offset of the next rollout step
    bool iter_valid = false;       // `iter` holds the iteration counts of a previous solve of these instances (schedule key)
    void *u0 = nullptr;            // [B][nu]: u(:,0) of a controls-only rollout step (allocated by the first rollout)
    const void *plant_u = nullptr; // what the plant step reads: `u` (row stride nu (N-1)) or `u0` (row stride nu)
    long long plant_u_stride = 0;
    std::string err;
};
#define BAT(b) reinterpret_cast<tmpc_batch_impl *>(b)

namespace {

int bfail(tmpc_batch_impl *b, int code, const std::string &msg)
{
    if (b) { b->err = msg; if (b->c) b->c->err = msg; }
    return code;
}
#define BCUDA_TRY(b, call)                                                                         \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return bfail(b, TMPC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

template <class T, int NX, int NU, int NH>
cudaError_t launch_plant(tmpc_ctx_impl *c, tmpc_batch_impl *b, void *x_next_hist, void *u0_hist, int *iter_hist, int *status_hist, cudaStream_t s)
{
    const int threads = 128;
    const unsigned blocks = (unsigned)((b->B + threads - 1) / threads);
    const tmpc::Model<T, NX, NU, NH> *m = reinterpret_cast<const tmpc::Model<T, NX, NU, NH> *>(c->model.data());
    if (c->policy == TMPC_ORDER_PARITY)
        tmpc::plant_kernel<T, NX, NU, NH, false><<<blocks, threads, 0, s>>>(*m, b->B, (T *)b->x0, (const T *)b->plant_u, b->plant_u_stride, (T *)x_next_hist,
                                                                               (T *)u0_hist, b->iter, b->status, iter_hist, status_hist);
    else
        tmpc::plant_kernel<T, NX, NU, NH, true><<<blocks, threads, 0, s>>>(*m, b->B, (T *)b->x0, (const T *)b->plant_u, b->plant_u_stride, (T *)x_next_hist,
                                                                              (T *)u0_hist, b->iter, b->status, iter_hist, status_hist);
    return cudaGetLastError();
}

template <class T>
cudaError_t launch_plant_rt(tmpc_ctx_impl *c, tmpc_batch_impl *b, void *x_next_hist, void *u0_hist, int *iter_hist, int *status_hist, cudaStream_t s)
{
    const unsigned blocks = (unsigned)((b->B + tmpc::RT_BLOCK - 1) / tmpc::RT_BLOCK);
    const tmpc::ModelRT<T> &m = *reinterpret_cast<const tmpc::ModelRT<T> *>(c->model_rt.data());
    const size_t smem = tmpc::rt_smem_bytes(c->nx, c->nu, sizeof(T));
    if (c->policy == TMPC_ORDER_PARITY) {
        cudaFuncSetAttribute(tmpc::plant_kernel_rt<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tmpc::plant_kernel_rt<T, false><<<blocks, tmpc::RT_BLOCK, smem, s>>>(m, b->B, (T *)b->x0, (const T *)b->plant_u, b->plant_u_stride, (T *)x_next_hist, (T *)u0_hist,
                                                                              b->iter, b->status, iter_hist, status_hist);
    } else {
        cudaFuncSetAttribute(tmpc::plant_kernel_rt<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tmpc::plant_kernel_rt<T, true><<<blocks, tmpc::RT_BLOCK, smem, s>>>(m, b->B, (T *)b->x0, (const T *)b->plant_u, b->plant_u_stride, (T *)x_next_hist, (T *)u0_hist,
                                                                             b->iter, b->status, iter_hist, status_hist);
    }
    return cudaGetLastError();
}

cudaError_t dispatch_plant(tmpc_ctx_impl *c, tmpc_batch_impl *b, void *xh, void *uh, int *ih, int *sh, cudaStream_t s)
{
    const bool f32 = c->dtype == TMPC_F32;
    if (c->nx == 12 && c->nu == 4 && c->N == 10)
        return f32 ? launch_plant<float, 12, 4, 10>(c, b, xh, uh, ih, sh, s) : launch_plant<double, 12, 4, 10>(c, b, xh, uh, ih, sh, s);
    if (c->nx == 4 && c->nu == 1 && c->N == 10)
        return f32 ? launch_plant<float, 4, 1, 10>(c, b, xh, uh, ih, sh, s) : launch_plant<double, 4, 1, 10>(c, b, xh, uh, ih, sh, s);
    if (c->nx == 32 && c->nu == 8 && c->N == 50 && f32) return launch_plant<float, 32, 8, 50>(c, b, xh, uh, ih, sh, s);
    if (c->rt_ready) return f32 ? launch_plant_rt<float>(c, b, xh, uh, ih, sh, s) : launch_plant_rt<double>(c, b, xh, uh, ih, sh, s);
    return cudaErrorInvalidValue;
}

int batch_copy_in(tmpc_batch_impl *b, void *dst, const void *src, size_t bytes, int mem)
{
    if (!src) return bfail(b, TMPC_ERR_INVALID, "NULL source");
    BCUDA_TRY(b, cudaSetDevice(b->c->device));
    BCUDA_TRY(b, cudaMemcpyAsync(dst, src, bytes, mem == TMPC_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, b->c->stream));
    if (mem != TMPC_MEM_DEVICE) BCUDA_TRY(b, cudaStreamSynchronize(b->c->stream));   // the caller may reuse its buffer
    return TMPC_OK;
}
int batch_copy_out(tmpc_batch_impl *b, void *dst, const void *src, size_t bytes, int mem)
{
    if (!dst) return bfail(b, TMPC_ERR_INVALID, "NULL destination");
    BCUDA_TRY(b, cudaSetDevice(b->c->device));
    BCUDA_TRY(b, cudaMemcpyAsync(dst, src, bytes, mem == TMPC_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, b->c->stream));
    if (mem != TMPC_MEM_DEVICE) BCUDA_TRY(b, cudaStreamSynchronize(b->c->stream));
    return TMPC_OK;
}

// Longest-first schedule of a closed loop: consecutive MPC steps of one instance need almost the same number of iterations, so the
// previous solve's `iter` ranks this one's load (sorted on the stream ahead of the kernel: 1M instances in ~40 us).  Only
// for batches that keep the GPU busy for several rounds of lanes; TMPC_LPT=0 turns it off.
// controls_only: neither x nor u is written (nothing can read them before the next solve overwrites them: the steps of a rollout
// but the last); u(:,0) goes to b->u0, which is all the plant step needs -- and all the kernels then skip their emission work
// roll_steps > 1: ONE launch runs that many MPC steps per instance (fused closed loop, tmpc_kernel_f32.cuh ROLL); rx / ru / ri / rs
// receive the histories of the steps before the last (SolveArgs::roll_*), the last step's outputs land where a plain solve's do
int batch_solve_async(tmpc_batch_impl *b, bool controls_only = false, int roll_steps = 0, void *rx = nullptr, void *ru = nullptr, int *ri = nullptr,
                      int *rs = nullptr)
{
    tmpc_ctx_impl *c = b->c;
    DevArgs da{};
    da.batch = b->B; da.x0 = b->x0; da.Xref = b->xref;
    da.xref_stride = b->xref_shared ? 0 : (long long)c->nx * c->N;
    da.wd = b->d; da.wy = b->y; da.wg = b->g; da.wv = b->v; da.wz = b->z;
    da.x = b->x; da.u = b->u; da.iter = b->iter; da.status = b->status; da.resid = b->resid;
    b->plant_u = b->u; b->plant_u_stride = (long long)c->nu * (c->N - 1);
    if (controls_only && b->u0) {
        da.x = nullptr; da.u = nullptr; da.u0 = b->u0;
        b->plant_u = b->u0; b->plant_u_stride = c->nu;
    }
    if (roll_steps > 1) {
        da.roll_steps = roll_steps; da.roll_x = rx; da.roll_u0 = ru; da.roll_iter = ri; da.roll_status = rs;
        if (b->table) { da.roll_table = b->table; da.roll_rows = b->table_rows; da.roll_start = b->start; da.roll_step0 = (int)b->steps_done; }
    }
    c->stats.instances = b->B;
    c->stats.launches = 0;
    const char *e = getenv("TMPC_LPT");
    const bool uniform_claims = roll_steps > 1 && c->nx == 4;   // fused 4/1/10 loops: every claim is a few iterations per step, index order keeps the rows coalesced
    if (b->iter_valid && !uniform_claims && !(e && !strcmp(e, "0")) && !c->ib_batch && b->B >= 4LL * 256 * c->sm_count && b->B < (1LL << 31)) {
        int rc = order_after_previous(c, c->stream);
        int bits = 1;
        while ((1 << bits) <= c->max_iter && bits < 31) ++bits;
        if (rc == TMPC_OK) rc = lpt_from_keys(c, da, c->stream, reinterpret_cast<const unsigned *>(b->iter), bits);
        if (rc != TMPC_OK) return rc;
    }
    int rc = launch_device(c, da, true, c->stream, true);
    if (rc == TMPC_OK) { c->stats_pending = true; b->iter_valid = true; }
    return rc;
}

}  // namespace

extern "C" {

int tmpc_batch_create(tmpc_ctx *ctx, int64_t batch, tmpc_batch **out)
{
    if (!ctx || !out) return TMPC_ERR_INVALID;
    *out = nullptr;
    tmpc_ctx_impl *c = CTX(ctx);
    if (batch < 1) return fail(c, TMPC_ERR_INVALID, "batch must be >= 1");
    if (!c->has_model) return fail(c, TMPC_ERR_STATE, "tmpc_set_model has not been called");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const size_t es = esize(c), xrow = (size_t)c->nx * c->N, urow = (size_t)c->nu * (c->N - 1);
    auto a256 = [](size_t v) { return (v + 255) & ~size_t(255); };
    const size_t sz_x0 = a256(batch * c->nx * es), sz_xr = a256(batch * xrow * es), sz_u = a256(batch * urow * es),
                 sz_i = a256(batch * 4), sz_r = a256(batch * 4 * es);
    const size_t total = sz_x0 + 4 * sz_xr /*xref g v x*/ + 4 * sz_u /*d y z u*/ + 2 * sz_i + sz_r;
    tmpc_batch_impl *b = new tmpc_batch_impl;
    b->c = c; b->B = batch;
    cudaError_t e = cudaMalloc((void **)&b->base, total);
    if (e != cudaSuccess) { delete b; return fail(c, TMPC_ERR_CUDA, std::string("tmpc_batch_create: cudaMalloc: ") + cudaGetErrorString(e)); }
    char *p = b->base;
    auto take = [&](size_t n) { char *q = p; p += n; return (void *)q; };
    b->x0 = take(sz_x0); b->xref = take(sz_xr); b->g = take(sz_xr); b->v = take(sz_xr); b->x = take(sz_xr);
    b->d = take(sz_u); b->y = take(sz_u); b->z = take(sz_u); b->u = take(sz_u);
    b->iter = (int *)take(sz_i); b->status = (int *)take(sz_i); b->resid = take(sz_r);
    e = cudaMemsetAsync(b->base, 0, total, c->stream);     // the zeroing block of the examples (hovering.cpp:49-71)
    if (e != cudaSuccess) { cudaFree(b->base); delete b; return fail(c, TMPC_ERR_CUDA, "tmpc_batch_create: memset failed"); }
    *out = reinterpret_cast<tmpc_batch *>(b);
    return TMPC_OK;
}

int tmpc_batch_destroy(tmpc_batch *bt)
{
    if (!bt) return TMPC_OK;
    tmpc_batch_impl *b = BAT(bt);
    cudaSetDevice(b->c->device);
    cudaStreamSynchronize(b->c->stream);
    if (b->base) cudaFree(b->base);
    if (b->u0) cudaFree(b->u0);
    if (b->table) cudaFree(b->table);
    if (b->start) cudaFree(b->start);
    delete b;
    return TMPC_OK;
}

int tmpc_batch_set_x0(tmpc_batch *bt, const void *x0, int32_t mem)
{
    if (!bt) return TMPC_ERR_INVALID;
    tmpc_batch_impl *b = BAT(bt);
    return batch_copy_in(b, b->x0, x0, (size_t)b->B * b->c->nx * esize(b->c), mem);
}

int tmpc_batch_set_xref(tmpc_batch *bt, const void *xref, int32_t shared, int32_t mem)
{
    if (!bt) return TMPC_ERR_INVALID;
    tmpc_batch_impl *b = BAT(bt);
    const size_t xrow = (size_t)b->c->nx * b->c->N * esize(b->c);
    int rc = batch_copy_in(b, b->xref, xref, shared ? xrow : (size_t)b->B * xrow, mem);
    if (rc == TMPC_OK) b->xref_shared = shared != 0;
    return rc;
}

int tmpc_batch_set_xref_table(tmpc_batch *bt, const void *table, int64_t rows, const int32_t *start, int32_t mem)
{
    if (!bt) return TMPC_ERR_INVALID;
    tmpc_batch_impl *b = BAT(bt);
    tmpc_ctx_impl *c = b->c;
    if (!table || rows < c->N) return bfail(b, TMPC_ERR_INVALID, "the reference table needs at least N rows");
    BCUDA_TRY(b, cudaSetDevice(c->device));
    BCUDA_TRY(b, cudaStreamSynchronize(c->stream));
    if (b->table) { cudaFree(b->table); b->table = nullptr; }
    if (b->start) { cudaFree(b->start); b->start = nullptr; }
    const size_t bytes = (size_t)rows * c->nx * esize(c);
    BCUDA_TRY(b, cudaMalloc(&b->table, bytes));
    int rc = batch_copy_in(b, b->table, table, bytes, mem);
    if (rc != TMPC_OK) return rc;
    if (start) {
        BCUDA_TRY(b, cudaMalloc((void **)&b->start, (size_t)b->B * sizeof(int)));
        rc = batch_copy_in(b, b->start, start, (size_t)b->B * sizeof(int), mem);
        if (rc != TMPC_OK) return rc;
    }
    b->table_rows = rows;
    b->steps_done = 0;
    b->iter_valid = false;
    b->xref_shared = false;
    return TMPC_OK;
}

int tmpc_batch_reset_dual_variables(tmpc_batch *bt)
{
    if (!bt) return TMPC_ERR_INVALID;
    tmpc_batch_impl *b = BAT(bt);
    tmpc_ctx_impl *c = b->c;
    const size_t es = esize(c);
    BCUDA_TRY(b, cudaSetDevice(c->device));
    BCUDA_TRY(b, cudaMemsetAsync(b->y, 0, (size_t)b->B * c->nu * (c->N - 1) * es, c->stream));
    BCUDA_TRY(b, cudaMemsetAsync(b->g, 0, (size_t)b->B * c->nx * c->N * es, c->stream));
    return TMPC_OK;
}

int tmpc_batch_reset(tmpc_batch *bt)
{
    if (!bt) return TMPC_ERR_INVALID;
    tmpc_batch_impl *b = BAT(bt);
    tmpc_ctx_impl *c = b->c;
    const size_t es = esize(c), xb = (size_t)b->B * c->nx * c->N * es, ub = (size_t)b->B * c->nu * (c->N - 1) * es;
    BCUDA_TRY(b, cudaSetDevice(c->device));
    BCUDA_TRY(b, cudaMemsetAsync(b->d, 0, ub, c->stream));
    BCUDA_TRY(b, cudaMemsetAsync(b->y, 0, ub, c->stream));
    BCUDA_TRY(b, cudaMemsetAsync(b->z, 0, ub, c->stream));
    BCUDA_TRY(b, cudaMemsetAsync(b->g, 0, xb, c->stream));
    BCUDA_TRY(b, cudaMemsetAsync(b->v, 0, xb, c->stream));
    b->steps_done = 0;
    b->iter_valid = false;
    return TMPC_OK;
}

int tmpc_batch_solve(tmpc_batch *bt)
{
    if (!bt) return TMPC_ERR_INVALID;
    tmpc_batch_impl *b = BAT(bt);
    BCUDA_TRY(b, cudaSetDevice(b->c->device));
    return batch_solve_async(b);
}

int tmpc_batch_get(tmpc_batch *bt, int32_t what, void *dst, int32_t mem)
{
    if (!bt) return TMPC_ERR_INVALID;
    tmpc_batch_impl *b = BAT(bt);
    tmpc_ctx_impl *c = b->c;
    const size_t es = esize(c), xrow = (size_t)c->nx * c->N, urow = (size_t)c->nu * (c->N - 1);
    const void *src = nullptr;
    size_t per = 0;
    switch (what) {
    case TMPC_GET_X: src = b->x; per = xrow * es; break;
    case TMPC_GET_U: src = b->u; per = urow * es; break;
    case TMPC_GET_ITER: src = b->iter; per = 4; break;
    case TMPC_GET_STATUS: src = b->status; per = 4; break;
    case TMPC_GET_RESID: src = b->resid; per = 4 * es; break;
    case TMPC_GET_X0: src = b->x0; per = c->nx * es; break;
    case TMPC_GET_D: src = b->d; per = urow * es; break;
    case TMPC_GET_Y: src = b->y; per = urow * es; break;
    case TMPC_GET_Z: src = b->z; per = urow * es; break;
    case TMPC_GET_G: src = b->g; per = xrow * es; break;
    case TMPC_GET_V: src = b->v; per = xrow * es; break;
    default: return bfail(b, TMPC_ERR_INVALID, "bad tmpc_batch_get selector");
    }
    return batch_copy_out(b, dst, src, (size_t)b->B * per, mem);
}

}  // extern "C"

// hist_stride: instances per step row of the caller's HOST history arrays (>= batch; tmpc_multi_rollout: the whole job's batch, this
// device's range starting at the pointers passed) -- device histories are always dense
int batch_rollout(tmpc_batch *bt, int32_t steps, int32_t reset_duals, void *x0_hist, void *u0_hist, int32_t *iter_hist,
                  int32_t *status_hist, int32_t mem, int64_t hist_stride)
{
    if (!bt) return TMPC_ERR_INVALID;
    tmpc_batch_impl *b = BAT(bt);
    tmpc_ctx_impl *c = b->c;
    if (steps < 0) return bfail(b, TMPC_ERR_INVALID, "negative step count");
    if (steps == 0) return TMPC_OK;
    BCUDA_TRY(b, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const size_t es = esize(c);
    const int64_t B = b->B;
    const size_t nxb = (size_t)B * c->nx * es, nub = (size_t)B * c->nu * es, ib = (size_t)B * 4;
    // histories: device buffers when the caller's are on the host
    const bool host = mem != TMPC_MEM_DEVICE;
    char *dx = nullptr, *du = nullptr;
    int *di = nullptr, *ds = nullptr;
    auto cleanup = [&]() {
        if (host) { if (dx) cudaFree(dx); if (du) cudaFree(du); if (di) cudaFree(di); if (ds) cudaFree(ds); }
    };
    auto dev_hist = [&](void *user, size_t bytes, void **dptr) -> cudaError_t {
        *dptr = nullptr;
        if (!user) return cudaSuccess;
        if (!host) { *dptr = user; return cudaSuccess; }
        return cudaMalloc(dptr, bytes);
    };
    cudaError_t e;
    if ((e = dev_hist(x0_hist, nxb * (steps + 1), (void **)&dx)) != cudaSuccess || (e = dev_hist(u0_hist, nub * steps, (void **)&du)) != cudaSuccess ||
        (e = dev_hist(iter_hist, ib * steps, (void **)&di)) != cudaSuccess || (e = dev_hist(status_hist, ib * steps, (void **)&ds)) != cudaSuccess) {
        cleanup();
        return bfail(b, TMPC_ERR_CUDA, std::string("tmpc_batch_rollout: history allocation: ") + cudaGetErrorString(e));
    }
    if (dx) BCUDA_TRY(b, cudaMemcpyAsync(dx, b->x0, nxb, cudaMemcpyDeviceToDevice, s));
    if (!b->u0 && steps > 1) BCUDA_TRY(b, cudaMalloc(&b->u0, (size_t)B * c->nu * es));
    cudaEvent_t r0, r1;
    BCUDA_TRY(b, cudaEventCreate(&r0));
    BCUDA_TRY(b, cudaEventCreate(&r1));
    float kernel_ms = 0.f;
    BCUDA_TRY(b, cudaEventRecord(r0, s));
    int rc = TMPC_OK;
    int k0 = 0;
    if (steps >= 2 && reset_duals && roll_supported(c) && !getenv("TMPC_ROLLOUT_FULL_OUTPUTS")) {
        // Fused closed loop: every lane of ONE persistent launch takes an instance through all `steps` solves (plant step, dual reset and
        // the warm d / v / z hand-over stay on chip); only the last step's plant update is left to the loop below
        c->duals_zero_next = true;
        rc = batch_solve_async(b, false, steps, dx ? dx + nxb : nullptr, du, di, ds);
        if (rc == TMPC_OK && b->table) {
            // the lanes read their windows from the table themselves; the batch's Xref buffer is left as the per-step path leaves it
            // (the last step's window), for a wrapper-style solve that may follow
            const long long n = (long long)B * c->nx * c->N;
            tmpc::xref_window_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(B, c->nx, c->N, (const float *)b->table, b->table_rows, b->start,
                                                                                     (int)(b->steps_done + steps - 1), (float *)b->xref);
        }
        if (rc == TMPC_OK) {
            k0 = steps - 1;
            e = dispatch_plant(c, b, dx ? dx + nxb * steps : nullptr, du ? du + nub * k0 : nullptr, di ? di + (size_t)B * k0 : nullptr,
                               ds ? ds + (size_t)B * k0 : nullptr, s);
            if (e != cudaSuccess) rc = bfail(b, TMPC_ERR_CUDA, std::string("plant step: ") + cudaGetErrorString(e));
            b->steps_done += steps;
            k0 = steps;
        }
    }
    for (int k = k0; k < steps && rc == TMPC_OK; ++k) {
        if (b->table) {   // 2. reference window (tracking.cpp:101)
            const long long n = (long long)B * c->nx * c->N;
            const unsigned blocks = (unsigned)((n + 255) / 256);
            if (c->dtype == TMPC_F32)
                tmpc::xref_window_kernel<float><<<blocks, 256, 0, s>>>(B, c->nx, c->N, (const float *)b->table, b->table_rows, b->start,
                                                                         (int)(b->steps_done), (float *)b->xref);
            else
                tmpc::xref_window_kernel<double><<<blocks, 256, 0, s>>>(B, c->nx, c->N, (const double *)b->table, b->table_rows, b->start,
                                                                          (int)(b->steps_done), (double *)b->xref);
        }
        if (reset_duals) {   // 3. y = 0, g = 0 (hovering.cpp:100-101)
            KernelInfo ki;
            const bool in_kernel = !c->ib_batch && lookup_kernel(c->nx, c->nu, c->N, c->dtype, c->policy, true, ki, c->pattern) && ki.model_kind == 1;
            if (in_kernel) c->duals_zero_next = true;   // the fp32 12/4/10 kernel zero-fills them on chip: 624 B per instance neither written nor read
            else {
                cudaMemsetAsync(b->y, 0, (size_t)B * c->nu * (c->N - 1) * es, s);
                cudaMemsetAsync(b->g, 0, (size_t)B * c->nx * c->N * es, s);
            }
        }
        // 4. tiny_solve.  Only the LAST step's trajectories can ever be read (tmpc_batch_get): the others are controls-only solves
        rc = batch_solve_async(b, k + 1 < steps && !getenv("TMPC_ROLLOUT_FULL_OUTPUTS"));
        if (rc != TMPC_OK) break;
        // 5. plant step + histories
        e = dispatch_plant(c, b, dx ? dx + nxb * (k + 1) : nullptr, du ? du + nub * k : nullptr, di ? di + (size_t)B * k : nullptr,
                           ds ? ds + (size_t)B * k : nullptr, s);
        if (e != cudaSuccess) { rc = bfail(b, TMPC_ERR_CUDA, std::string("plant step: ") + cudaGetErrorString(e)); break; }
        b->steps_done += 1;
    }
    if (rc == TMPC_OK) {
        cudaEventRecord(r1, s);
        if (host) {
            // one row of `batch` instances per step; the caller's rows are hist_stride instances apart
            const size_t hs = (size_t)(hist_stride > 0 ? hist_stride : B);
            if (dx) cudaMemcpy2DAsync(x0_hist, hs * c->nx * es, dx, nxb, nxb, (size_t)steps + 1, cudaMemcpyDeviceToHost, s);
            if (du) cudaMemcpy2DAsync(u0_hist, hs * c->nu * es, du, nub, nub, (size_t)steps, cudaMemcpyDeviceToHost, s);
            if (di) cudaMemcpy2DAsync(iter_hist, hs * 4, di, ib, ib, (size_t)steps, cudaMemcpyDeviceToHost, s);
            if (ds) cudaMemcpy2DAsync(status_hist, hs * 4, ds, ib, ib, (size_t)steps, cudaMemcpyDeviceToHost, s);
        }
        e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = bfail(b, TMPC_ERR_CUDA, std::string("tmpc_batch_rollout: ") + cudaGetErrorString(e));
        else cudaEventElapsedTime(&kernel_ms, r0, r1);
    }
    cudaEventDestroy(r0);
    cudaEventDestroy(r1);
    cleanup();
    if (rc == TMPC_OK) {
        c->stats.launches = k0 ? 2 : steps * (2 + (b->table ? 1 : 0));
        c->rollout_ms = kernel_ms;
    }
    return rc;
}

extern "C" {

int tmpc_batch_rollout(tmpc_batch *bt, int32_t steps, int32_t reset_duals, void *x0_hist, void *u0_hist, int32_t *iter_hist,
                       int32_t *status_hist, int32_t mem)
{
    return batch_rollout(bt, steps, reset_duals, x0_hist, u0_hist, iter_hist, status_hist, mem, 0);
}

float tmpc_batch_last_rollout_ms(const tmpc_batch *bt) { return bt ? reinterpret_cast<const tmpc_batch_impl *>(bt)->c->rollout_ms : 0.f; }

const char *tmpc_batch_last_error(const tmpc_batch *bt) { return bt ? reinterpret_cast<const tmpc_batch_impl *>(bt)->err.c_str() : ""; }

}  // extern "C"
