// C ABI (include/tmpc.h) over the sm_100a kernels in tmpc_kernel.cuh.
// Owns: device/model state per context, kernel dispatch by (shape, dtype, policy), the chunked
// host<->device pipeline for TMPC_MEM_HOST callers, statistics.  No solver arithmetic runs on the host.
#include "tmpc.h"
#include "tmpc_dispatch.hpp"
#include "tmpc_kernel.cuh"
#include "tmpc_kernel_f32.cuh"
#include "tmpc_kernel_warp.cuh"
#include "tmpc_kernel_small.cuh"
#include "tmpc_steps.cuh"
#include "tmpc_kernel_rt.cuh"
#include "tmpc_orders_rt.hpp"

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <string>
#include <type_traits>
#include <vector>

namespace {

thread_local std::string g_create_error;

using tmpc_dispatch::KernelInfo;

struct tmpc_ctx_impl {
    int device = 0;
    int nx = 0, nu = 0, N = 0, dtype = 0, policy = 0;
    bool has_model = false;
    float rollout_ms = 0.f;  // device time of the last tmpc_batch_rollout
    bool const_bounds = false;  // bounds identical at every horizon stage (fp32 12/4/10 image)
    int pattern = 0;  // structural-sparsity specialisation the current model conforms to (0 = dense)
    bool warm_variant_ready = false;
    std::string err;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // Launches of one ctx share the work counter, the statistics slot, the scheduling buffers and the run-time-shape scratch:
    // they are serialised on the device through this event, whatever stream each one is queued on (order_after_previous).
    cudaEvent_t last_launch = nullptr;
    bool launched = false;
    int sm_count = 0;
    // model image exactly as the kernel's Model<T,...> struct (built on the host, passed by value)
    std::vector<unsigned char> model;      // tmpc::Model<T,...> image
    std::vector<unsigned char> model_f32;  // tmpc::ModelF32<...> image (packed fp32 kernel)
    tmpc::ModelWarp model_w{};             // warp-per-instance kernel: scalars + pointers into d_model_w
    float *d_model_w = nullptr;            // device image: fwd4 | bwd4 | Pt | Qd | xmin | xmax | umin | umax
    size_t d_model_w_floats = 0;
    // run-time-shape kernel (tmpc_kernel_rt.cuh): ModelRT<T> image, its device arrays, per-lane scratch
    std::vector<unsigned char> model_rt;
    void *d_model_rt = nullptr;
    size_t d_model_rt_bytes = 0;
    void *d_rt_scratch = nullptr;
    size_t d_rt_scratch_bytes = 0;
    // per-lane coalesced scratch of the fp32 12/4/10 kernel (SolveArgs::scratch)
    void *d_lane_scratch = nullptr;
    size_t d_lane_scratch_bytes = 0;
    void *d_pn_seed = nullptr;     // p_N seeds of a batch with per-instance Xref (pre-pass of the fp32 12/4/10 kernel)
    size_t d_pn_seed_bytes = 0;
    int reserve_sms_next = 0;      // host pipeline: the next launch leaves this many SMs to the ranking kernels that run beside it
    bool duals_zero_next = false;  // tmpc_batch rollout with reset duals: the next launch of the fp32 12/4/10 kernel zero-fills y, g itself
    void *d_model_f32 = nullptr;   // device copy of model_f32 (TMPC_KERNEL=f32_tma_cache: the kernel stages it into shared memory by TMA)
    bool rt_ready = false;
    // per-instance box bounds (tmpc_set_instance_bounds): device copies xmin | xmax | umin | umax, 0 = not set
    void *d_ib[4] = {nullptr, nullptr, nullptr, nullptr};
    long long ib_batch = 0;
    // longest-expected-first schedule (lpt_prepare): device copy of Kinf, sort buffers
    void *d_kinf = nullptr;
    unsigned *lpt_buf = nullptr;     // keys_in | keys_out | vals_in | vals_out, `lpt_cap` entries each
    size_t lpt_cap = 0;
    void *lpt_temp = nullptr;
    size_t lpt_temp_bytes = 0;
    int lpt_used = 0;                // the last solve ran with the schedule
    // settings
    double pri = 1e-3, dua = 1e-3;
    int max_iter = 100, check_term = 1, en_state = 1, en_input = 1;
    // raw model copies (dtype of ctx) kept to rebuild the image when settings change
    std::vector<unsigned char> Kinf, Pinf, Quu_inv, AmBKt, Adyn, Bdyn, Q, xmin, xmax, umin, umax;
    double rho = 0;
    bool has_xb = false, has_ub = false;
    // device scratch
    unsigned long long *d_counter = nullptr;  // [0] counter, [1..4] stats
    // host staging (TMPC_MEM_HOST)
    struct Stage {
        void *h_in = nullptr, *h_out = nullptr;  // pinned
        void *d_in = nullptr, *d_out = nullptr;
        size_t in_bytes = 0, out_bytes = 0;
        cudaStream_t s = nullptr;
        cudaEvent_t done = nullptr;
    } stage[3];
    int64_t stage_chunk = 0;
    // gated host pipeline (cold solves): full-batch device image + per-chunk completion counters
    void *g_in = nullptr, *g_out = nullptr;
    size_t g_in_bytes = 0, g_out_bytes = 0;
    unsigned *g_done = nullptr;
    size_t g_done_n = 0;
    cudaStream_t g_copy = nullptr, g_in_stream = nullptr;
    cudaEvent_t g_h2d = nullptr, g_first = nullptr;
    unsigned *g_gate = nullptr;   // [0] instances whose inputs have arrived, [1] gate timeout flag
    // stats of the last solve
    tmpc_stats stats{};
    bool stats_pending = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending_events;
};

#define CTX(c) reinterpret_cast<tmpc_ctx_impl *>(c)

int fail(tmpc_ctx_impl *c, int code, const std::string &msg)
{
    if (c) c->err = msg;
    else g_create_error = msg;
    return code;
}

#define CUDA_TRY(c, call)                                                                            \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(c, TMPC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
    } while (0)

// ---------------------------------------------------------------------------------------------
// kernel table (the instantiations live in k_*.cu, one translation unit per kernel family: tmpc_dispatch.hpp)
// ---------------------------------------------------------------------------------------------
// development switch: TMPC_KERNEL=generic | f32_smem | f32_tmem (default) selects the fp32 12/4/10 kernel
int kernel_variant()
{
    const char *e = getenv("TMPC_KERNEL");
    if (!e) return 2;
    if (!strcmp(e, "generic")) return 0;
    if (!strcmp(e, "f32_smem")) return 1;
    if (!strcmp(e, "f32_tma_cache")) return 3;
    return 2;
}

// Compiled shapes.  Thread-per-instance needs the per-instance state to fit shared memory:
//   quadrotor 12/4/10: 360 scalars -> 128 threads (f32) / 64 threads (f64) per SM
//   cartpole   4/1/10: 111 scalars -> 256 threads (f32) / 128 (f64)
// pattern: bit 0..7 = structural-sparsity pattern id (0 dense), bit 8 = bounds constant over the horizon
bool force_rt();
bool lookup_kernel_rt(int nx, int nu, int N, int dtype, int policy, KernelInfo &out);
bool lookup_kernel(int nx, int nu, int N, int dtype, int policy, bool warm, KernelInfo &out, int pattern_bits = 0)
{
    const int pattern = pattern_bits & 0xff;
    const bool cb = (pattern_bits & 0x100) != 0;
    if (force_rt()) return lookup_kernel_rt(nx, nu, N, dtype, policy, out);
    const char *e = getenv("TMPC_KERNEL");
    const bool generic = e && !strcmp(e, "generic");
    if (nx == 12 && nu == 4 && N == 10) {
        if (dtype == TMPC_F32) {
            const int v = kernel_variant();
            if (v == 0) return tmpc_dispatch::lookup_generic(nx, nu, N, dtype, policy, warm, 0, out);
            return tmpc_dispatch::lookup_f32(policy, warm, pattern, cb, v, out);   // (v == 3 falls back to 2 where it has no instance)
        }
        // double: two lanes per instance, 128 instances / SM (tmpc_kernel_f64p.cuh).  TMPC_KERNEL=f64_thread: one thread per instance,
        // g, v in tensor memory, 128 / SM; generic: all state in shared memory, 64 / SM
        if (!generic && !(e && !strcmp(e, "f64_thread")) && tmpc_dispatch::lookup_f64p(nx, nu, N, dtype, policy, warm, out)) return true;
        return tmpc_dispatch::lookup_generic(nx, nu, N, dtype, policy, warm, generic ? 1 : 0, out);
    }
    if (nx == 4 && nu == 1 && N == 10) {
        if (dtype == TMPC_F32) {
            // register-resident single-input kernel (tmpc_kernel_small.cuh); TMPC_KERNEL=generic | generic_unroll:
            // the shared-memory kernel (444 B of state: 512 instances / SM)
            if (generic) return tmpc_dispatch::lookup_generic(nx, nu, N, dtype, policy, warm, 0, out);
            if (e && !strcmp(e, "generic_unroll")) return tmpc_dispatch::lookup_generic(nx, nu, N, dtype, policy, warm, 2, out);
            if (e && !strcmp(e, "small256")) return tmpc_dispatch::lookup_small(256, policy, warm, out);
            if (e && !strcmp(e, "small512")) return tmpc_dispatch::lookup_small(512, policy, warm, out);
            return tmpc_dispatch::lookup_small(384, policy, warm, out);
        }
        return tmpc_dispatch::lookup_generic(nx, nu, N, dtype, policy, warm, 0, out);
    }
    // large shape: one warp per instance.  g,v in tensor memory -> 16 instances / SM (TMPC_KERNEL=warp_smem: all state in
    // shared memory, 17.8 KB each -> 12 instances / SM)
    if (nx == 32 && nu == 8 && N == 50 && dtype == TMPC_F32)
        return tmpc_dispatch::lookup_warp(e && !strcmp(e, "warp_smem") ? 2 : e && !strcmp(e, "warp4") ? 0 : e && !strcmp(e, "warp4x2") ? 3 : 1, policy, warm, out);
    return lookup_kernel_rt(nx, nu, N, dtype, policy, out);
}

bool shape_parity_pinned(int nx, int nu)
{
    // evaluation orders verified bit-for-bit against the compiled reference (oracle/, tests/): the three BASELINE shapes
    // by the fixtures under tests/golden, every other shape up to 64 by the dispatch rule that oracle/pin_shapes.py
    // checks against the compiled reference on 60+ shapes (tmpc_orders_rt.hpp)
    return nx >= 1 && nu >= 1 && nx <= tmpc::RT_MAXD && nu <= tmpc::RT_MAXD;
}

bool rt_shape_ok(int nx, int nu, int N)
{
    return nx >= 1 && nu >= 1 && N >= 2 && nx <= tmpc::RT_MAXD && nu <= tmpc::RT_MAXD &&
           (long long)nx * N < (1LL << 24) && (long long)nu * N < (1LL << 24);
}

bool force_rt()
{
    const char *e = getenv("TMPC_KERNEL");
    return e && !strcmp(e, "rt");
}

// run-time-shape kernel: any other shape (TMPC_KERNEL=rt forces it for the compiled shapes too)
bool lookup_kernel_rt(int nx, int nu, int N, int dtype, int policy, KernelInfo &out)
{
    if (!rt_shape_ok(nx, nu, N)) return false;
    const bool fast = policy == TMPC_ORDER_FAST;
    if (dtype == TMPC_F32) out.fn = fast ? (const void *)&tmpc::admm_kernel_rt<float, true> : (const void *)&tmpc::admm_kernel_rt<float, false>;
    else out.fn = fast ? (const void *)&tmpc::admm_kernel_rt<double, true> : (const void *)&tmpc::admm_kernel_rt<double, false>;
    out.smem = tmpc::rt_smem_bytes(nx, nu, dtype == TMPC_F32 ? 4 : 8);
    out.block = tmpc::RT_BLOCK;
    out.model_bytes = dtype == TMPC_F32 ? sizeof(tmpc::ModelRT<float>) : sizeof(tmpc::ModelRT<double>);
    out.model_kind = 3;
    out.per_block = tmpc::RT_BLOCK;
    return true;
}

// ---------------------------------------------------------------------------------------------
// model image
// ---------------------------------------------------------------------------------------------
template <class T, int NX, int NU, int NH> void build_model_t(tmpc_ctx_impl *c)
{
    using M = tmpc::Model<T, NX, NU, NH>;
    c->model.assign(sizeof(M), 0);
    M *m = reinterpret_cast<M *>(c->model.data());
    auto cp = [](T *dst, const std::vector<unsigned char> &src, size_t n) { std::memcpy(dst, src.data(), n * sizeof(T)); };
    cp(m->K, c->Kinf, NU * NX);
    cp(m->A, c->Adyn, NX * NX);
    cp(m->B, c->Bdyn, NX * NU);
    cp(m->Qi, c->Quu_inv, NU * NU);
    cp(m->M, c->AmBKt, NX * NX);
    cp(m->Pf, c->Pinf, NX * NX);
    cp(m->Qd, c->Q, NX);
    const T inf = std::numeric_limits<T>::infinity();
    // a disabled (or absent) bound becomes +-inf: min(+inf, max(-inf, v)) == v exactly, so the kernel
    // needs no branch for en_state_bound / en_input_bound (admm.cpp:51,57)
    const bool xs = c->en_state && c->has_xb, us = c->en_input && c->has_ub;
    for (int k = 0; k < NH * NX; ++k) {
        m->xmin[k] = xs ? reinterpret_cast<const T *>(c->xmin.data())[k] : -inf;
        m->xmax[k] = xs ? reinterpret_cast<const T *>(c->xmax.data())[k] : inf;
    }
    for (int k = 0; k < (NH - 1) * NU; ++k) {
        m->umin[k] = us ? reinterpret_cast<const T *>(c->umin.data())[k] : -inf;
        m->umax[k] = us ? reinterpret_cast<const T *>(c->umax.data())[k] : inf;
    }
    m->rho = (T)c->rho;
    m->nrho = -(T)c->rho;
    m->pri_tol = (T)c->pri;
    m->dua_tol = (T)c->dua;
    m->max_iter = c->max_iter;
    m->check_term = c->check_term;
}

template <int NX, int NU, int NH> void build_model_f32(tmpc_ctx_impl *c, std::vector<unsigned char> &img)
{
    using M = tmpc::ModelF32<NX, NU, NH>;
    constexpr int RS = NU + NX;
    img.assign(sizeof(M), 0);
    M *m = reinterpret_cast<M *>(img.data());
    const float *K = reinterpret_cast<const float *>(c->Kinf.data());     // [r + k*NU]
    const float *A = reinterpret_cast<const float *>(c->Adyn.data());     // [r + k*NX]
    const float *B = reinterpret_cast<const float *>(c->Bdyn.data());     // [r + k*NX]
    const float *Qi = reinterpret_cast<const float *>(c->Quu_inv.data());
    const float *Mm = reinterpret_cast<const float *>(c->AmBKt.data());
    const float *Pf = reinterpret_cast<const float *>(c->Pinf.data());
    // does the model conform to a compiled structural pattern?  Every coefficient the pattern drops must be an
    // exact zero, every multiply it skips an exact one.  TMPC_DENSE=1 forces the dense instance.
    c->pattern = 0;
    int mperm[NX];
    for (int r = 0; r < NX; ++r) mperm[r] = r;
    if constexpr (NX == 12 && NU == 4) {
        using PQ = tmpc::PatQuadrotor;
        bool ok = !getenv("TMPC_DENSE");
        for (int r = 0; r < NX && ok; ++r)
            for (int k = 0; k < NX && ok; ++k) {
                const float a = A[r + k * NX], mm = Mm[r + k * NX];
                if (!((PQ::a_nz[r] >> k) & 1u) && a != 0.f) ok = false;
                if (((PQ::a_one[r] >> k) & 1u) && a != 1.f) ok = false;
                if (!((PQ::m_nz[r] >> k) & 1u) && mm != 0.f) ok = false;
            }
        for (int j = 0; j < NX / 2 && ok; ++j)   // the pair masks must cover both rows of each pair
            for (int h = 0; h < 2; ++h)
                if (PQ::m_nz[PQ::m_perm[2 * j + h]] & ~PQ::m_pair_nz[j]) ok = false;
        if (ok) {
            c->pattern = PQ::id;
            for (int r = 0; r < NX; ++r) mperm[r] = PQ::m_perm[r];   // AmBKt rows stored in pair order
        }
    }
    m->nz2 = make_float2(-0.0f, -0.0f);
    for (int k = 0; k < NX; ++k) {
        for (int r = 0; r < NU; ++r) m->KA[k * RS + r] = K[r + k * NU];
        for (int r = 0; r < NX; ++r) m->KA[k * RS + NU + r] = A[r + k * NX];
        for (int r = 0; r < NU; ++r) m->BM[k * RS + r] = B[k + r * NX];       // (B^T)(r,k)
        for (int r = 0; r < NX; ++r) m->BM[k * RS + NU + r] = Mm[mperm[r] + k * NX];
        for (int j = 0; j < NX; ++j) m->Pt[k * NX + j] = Pf[k + j * NX];      // Pinf(k,j)
    }
    for (int k = 0; k < NU; ++k) {
        for (int r = 0; r < NX; ++r) m->Bc[k * NX + r] = B[r + k * NX];
        for (int r = 0; r < NU; ++r) m->Qi[k * NU + r] = Qi[r + k * NU];
        for (int j = 0; j < NX; ++j) m->Kr[k * NX + j] = K[k + j * NU];       // (K^T)(j,k)
    }
    std::memcpy(m->Qd, c->Q.data(), NX * sizeof(float));
    const float inf = std::numeric_limits<float>::infinity();
    const bool xs = c->en_state && c->has_xb, us = c->en_input && c->has_ub;
    for (int k = 0; k < NH * NX; ++k) {
        m->xmin[k] = xs ? reinterpret_cast<const float *>(c->xmin.data())[k] : -inf;
        m->xmax[k] = xs ? reinterpret_cast<const float *>(c->xmax.data())[k] : inf;
    }
    for (int k = 0; k < (NH - 1) * NU; ++k) {
        m->umin[k] = us ? reinterpret_cast<const float *>(c->umin.data())[k] : -inf;
        m->umax[k] = us ? reinterpret_cast<const float *>(c->umax.data())[k] : inf;
    }
    m->rho = (float)c->rho;
    m->nrho = -(float)c->rho;
    m->pri_tol = (float)c->pri;
    m->dua_tol = (float)c->dua;
    m->max_iter = c->max_iter;
    m->check_term = c->check_term;
    // bounds identical at every stage?  (then the CB kernel instance reads stage 0's rows with fixed addresses)
    bool cb = !getenv("TMPC_NO_CONST_BOUNDS");
    for (int i = 1; i < NH && cb; ++i)
        for (int j = 0; j < NX && cb; ++j)
            if (m->xmin[i * NX + j] != m->xmin[j] || m->xmax[i * NX + j] != m->xmax[j]) cb = false;
    for (int i = 1; i < NH - 1 && cb; ++i)
        for (int j = 0; j < NU && cb; ++j)
            if (m->umin[i * NU + j] != m->umin[j] || m->umax[i * NU + j] != m->umax[j]) cb = false;
    c->const_bounds = cb;
}

// lane-major coefficient image of the warp-per-instance kernel (layout: tmpc_kernel_warp.cuh ModelWarp)
bool build_model_warp(tmpc_ctx_impl *c)
{
    constexpr int NX = tmpc::WNX, NU = tmpc::WNU;
    const int NH = c->N;
    const float *K = reinterpret_cast<const float *>(c->Kinf.data());     // [r + k*NU]
    const float *A = reinterpret_cast<const float *>(c->Adyn.data());     // [r + k*NX]
    const float *B = reinterpret_cast<const float *>(c->Bdyn.data());     // [r + k*NX]
    const float *Qi = reinterpret_cast<const float *>(c->Quu_inv.data()); // [r + k*NU]
    const float *Mm = reinterpret_cast<const float *>(c->AmBKt.data());
    const float *Pf = reinterpret_cast<const float *>(c->Pinf.data());
    const size_t n_fwd = (size_t)tmpc::WARP_FWD4 * 32 * 4, n_bwd = (size_t)tmpc::WARP_BWD4 * 32 * 4;
    const size_t n_x = (size_t)NH * NX, n_u = (size_t)(NH - 1) * NU;
    const size_t total = n_fwd + n_bwd + NX * NX + NX + 2 * n_x + 2 * n_u;
    std::vector<float> h(total, 0.f);
    float *fwd = h.data(), *bwd = fwd + n_fwd, *Pt = bwd + n_bwd, *Qd = Pt + NX * NX;
    float *xmin = Qd + NX, *xmax = xmin + n_x, *umin = xmax + n_x, *umax = umin + n_u;
    auto at4 = [](float *base, int group, int lane, int t) -> float & { return base[((size_t)group * 32 + lane) * 4 + t]; };
    for (int lane = 0; lane < 32; ++lane) {
        const int ur = lane >> 2, sl = lane & 3;
        for (int k = 0; k < NX; ++k) {
            at4(fwd, k / 4, lane, k % 4) = A[lane + k * NX];
            at4(fwd, 8 + k / 4, lane, k % 4) = K[ur + k * NU];
            at4(bwd, k / 4, lane, k % 4) = Mm[lane + k * NX];
        }
        for (int k = 0; k < NU; ++k) {
            at4(fwd, 16 + k / 4, lane, k % 4) = B[lane + k * NX];
            at4(bwd, 8 + k / 4, lane, k % 4) = B[(4 * k + sl) + ur * NX];   // (B^T)(ur, 4k + sl): packet k of SIMD lane sl
            at4(bwd, 10 + k / 4, lane, k % 4) = Qi[ur + k * NU];
            at4(bwd, 12 + k / 4, lane, k % 4) = K[k + lane * NU];           // (K^T)(lane, k)
        }
    }
    for (int k = 0; k < NX; ++k)
        for (int j = 0; j < NX; ++j) Pt[k * NX + j] = Pf[k + j * NX];
    std::memcpy(Qd, c->Q.data(), NX * sizeof(float));
    const float inf = std::numeric_limits<float>::infinity();
    const bool xs = c->en_state && c->has_xb, us = c->en_input && c->has_ub;
    for (size_t k = 0; k < n_x; ++k) {
        xmin[k] = xs ? reinterpret_cast<const float *>(c->xmin.data())[k] : -inf;
        xmax[k] = xs ? reinterpret_cast<const float *>(c->xmax.data())[k] : inf;
    }
    for (size_t k = 0; k < n_u; ++k) {
        umin[k] = us ? reinterpret_cast<const float *>(c->umin.data())[k] : -inf;
        umax[k] = us ? reinterpret_cast<const float *>(c->umax.data())[k] : inf;
    }
    if (c->d_model_w_floats < total) {
        if (c->d_model_w) cudaFree(c->d_model_w);
        c->d_model_w = nullptr; c->d_model_w_floats = 0;
        if (cudaMalloc((void **)&c->d_model_w, total * sizeof(float)) != cudaSuccess) return false;
        c->d_model_w_floats = total;
    }
    // a solve still in flight on the ctx stream may be reading the old image
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return false;
    if (cudaMemcpy(c->d_model_w, h.data(), total * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    float *d = c->d_model_w;
    tmpc::ModelWarp &m = c->model_w;
    m.fwd4 = reinterpret_cast<const float4 *>(d);
    m.bwd4 = reinterpret_cast<const float4 *>(d + n_fwd);
    m.Pt = d + n_fwd + n_bwd;
    m.Qd = m.Pt + NX * NX;
    m.xmin = m.Qd + NX; m.xmax = m.xmin + n_x; m.umin = m.xmax + n_x; m.umax = m.umin + n_u;
    m.rho = (float)c->rho; m.nrho = -(float)c->rho;
    m.pri_tol = (float)c->pri; m.dua_tol = (float)c->dua;
    m.max_iter = c->max_iter; m.check_term = c->check_term;
    m.nz2 = make_float2(-0.0f, -0.0f);
    return true;
}

// Run-time-shape kernel: device image = Qd | xmin | xmax | umin | umax | for each product and order variant a row-major
// copy of its coefficient matrix with the columns in program order | the reduction programs (uint16).
// ModelRT<T> holds pointers into it.
template <class T> bool build_model_rt_t(tmpc_ctx_impl *c)
{
    const int nx = c->nx, nu = c->nu, N = c->N;
    const tmpc_rt::Orders o = tmpc_rt::build_orders(nx, nu, N, (int)sizeof(T), c->policy == TMPC_ORDER_FAST);
    if (o.max_depth > tmpc::RT_STACK) return false;
    const T *Kinf = reinterpret_cast<const T *>(c->Kinf.data()), *A = reinterpret_cast<const T *>(c->Adyn.data()),
            *Bm = reinterpret_cast<const T *>(c->Bdyn.data()), *Qi = reinterpret_cast<const T *>(c->Quu_inv.data()),
            *M = reinterpret_cast<const T *>(c->AmBKt.data()), *Pf = reinterpret_cast<const T *>(c->Pinf.data());
    std::vector<T> img;
    auto al4 = [&]() { while (img.size() % 4) img.push_back(T(0)); };
    const size_t n_x = (size_t)N * nx, n_u = (size_t)(N - 1) * nu;
    const size_t o_qd = img.size();
    img.insert(img.end(), reinterpret_cast<const T *>(c->Q.data()), reinterpret_cast<const T *>(c->Q.data()) + nx);
    al4();
    const T inf = std::numeric_limits<T>::infinity();
    const bool xs = c->en_state && c->has_xb, us = c->en_input && c->has_ub;
    size_t o_b[4];
    for (int w = 0; w < 4; ++w) {
        o_b[w] = img.size();
        const bool isx = w < 2, en = isx ? xs : us;
        const std::vector<unsigned char> &src = w == 0 ? c->xmin : w == 1 ? c->xmax : w == 2 ? c->umin : c->umax;
        const size_t n = isx ? n_x : n_u;
        for (size_t k = 0; k < n; ++k) img.push_back(en ? reinterpret_cast<const T *>(src.data())[k] : ((w & 1) ? inf : -inf));
        al4();
    }
    // coefficient copies: element (r, k) of each product as the ADMM sweeps use it
    struct PD { const tmpc_rt::Prod *p; int R, K; std::function<T(int, int)> at; };
    const PD pd[8] = {
        {&o.Kx, nu, nx, [&](int r, int k) { return Kinf[r + k * nu]; }},   // Kinf x
        {&o.Ax, nx, nx, [&](int r, int k) { return A[r + k * nx]; }},      // Adyn x
        {&o.Bu, nx, nu, [&](int r, int k) { return Bm[r + k * nx]; }},     // Bdyn u
        {&o.Btp, nu, nx, [&](int r, int k) { return Bm[k + r * nx]; }},    // Bdyn^T p
        {&o.Qs, nu, nu, [&](int r, int k) { return Qi[r + k * nu]; }},     // Quu_inv s
        {&o.Mp, nx, nx, [&](int r, int k) { return M[r + k * nx]; }},      // AmBKt p
        {&o.Ktr, nx, nu, [&](int r, int k) { return Kinf[k + r * nu]; }},  // Kinf^T r
        {&o.XtP, nx, nx, [&](int r, int k) { return Pf[k + r * nx]; }},    // Xref^T Pinf
    };
    size_t o_cf[8][2];
    for (int q = 0; q < 8; ++q)
        for (int v = 0; v < 2; ++v) {
            const tmpc_rt::Variant &var = v ? pd[q].p->b : pd[q].p->a;
            if (v == 1 && var.prog == pd[q].p->a.prog) { o_cf[q][1] = o_cf[q][0]; continue; }
            o_cf[q][v] = img.size();
            for (int r = 0; r < pd[q].R; ++r)
                for (int k = 0; k < pd[q].K; ++k)
                    img.push_back(pd[q].at(r, var.kind == tmpc_rt::K_FIXED ? k : (o.prog[var.prog + k] & 0xff)));
            al4();
        }
    const size_t prog_off = img.size() * sizeof(T);
    const size_t bytes = prog_off + ((o.prog.size() * 2 + 15) & ~size_t(15));
    std::vector<unsigned char> raw(bytes, 0);
    std::memcpy(raw.data(), img.data(), img.size() * sizeof(T));
    std::memcpy(raw.data() + prog_off, o.prog.data(), o.prog.size() * 2);
    if (cudaSetDevice(c->device) != cudaSuccess) return false;
    // the previous image may still be in use by a kernel
    if (cudaDeviceSynchronize() != cudaSuccess) return false;
    if (c->d_model_rt_bytes < bytes) {
        if (c->d_model_rt) cudaFree(c->d_model_rt);
        c->d_model_rt = nullptr; c->d_model_rt_bytes = 0;
        if (cudaMalloc(&c->d_model_rt, bytes) != cudaSuccess) return false;
        c->d_model_rt_bytes = bytes;
    }
    if (cudaMemcpy(c->d_model_rt, raw.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) return false;
    c->model_rt.assign(sizeof(tmpc::ModelRT<T>), 0);
    tmpc::ModelRT<T> &m = *reinterpret_cast<tmpc::ModelRT<T> *>(c->model_rt.data());
    const T *d = reinterpret_cast<const T *>(c->d_model_rt);
    const unsigned short *dprog = reinterpret_cast<const unsigned short *>(reinterpret_cast<const unsigned char *>(c->d_model_rt) + prog_off);
    m.nx = nx; m.nu = nu; m.N = N;
    m.Qd = d + o_qd; m.xmin = d + o_b[0]; m.xmax = d + o_b[1]; m.umin = d + o_b[2]; m.umax = d + o_b[3];
    tmpc::ProdRT<T> *dst[8] = {&m.Kx, &m.Ax, &m.Bu, &m.Btp, &m.Qs, &m.Mp, &m.Ktr, &m.XtP};
    for (int q = 0; q < 8; ++q)
        for (int v = 0; v < 2; ++v) {
            const tmpc_rt::Variant &var = v ? pd[q].p->b : pd[q].p->a;
            tmpc::VarRT<T> &dv = v ? dst[q]->b : dst[q]->a;
            dv.coef = d + o_cf[q][v];
            dv.prog = dprog + var.prog;
            dv.kind = var.kind;
            dv.order = var.order;
        }
    m.head_Kx = o.head_Kx; m.head_Ax = o.head_Ax; m.head_Qs = o.head_Qs; m.head_Mp = o.head_Mp;
    m.rt_u = o.rt_u; m.rt_x = o.rt_x; m.rt_p = o.rt_p; m.off_u = o.off_u; m.off_x = o.off_x; m.off_p = o.off_p; m.pk = o.pk; m.sb = o.sb;
    m.rho = (T)c->rho; m.nrho = -(T)c->rho; m.pri_tol = (T)c->pri; m.dua_tol = (T)c->dua;
    m.max_iter = c->max_iter; m.check_term = c->check_term;
    c->rt_ready = true;
    return true;
}

bool build_model_rt(tmpc_ctx_impl *c)
{
    if (!rt_shape_ok(c->nx, c->nu, c->N)) return false;
    return c->dtype == TMPC_F32 ? build_model_rt_t<float>(c) : build_model_rt_t<double>(c);
}

bool build_model_shape(tmpc_ctx_impl *c);
bool ib_on_f32_kernel(const tmpc_ctx_impl *c);
bool build_model(tmpc_ctx_impl *c)
{
    if (!build_model_shape(c)) return false;
    // per-instance bounds run on the run-time-shape kernel whatever the shape: keep its image current
    KernelInfo probe;
    if (c->ib_batch && !ib_on_f32_kernel(c) && !(lookup_kernel(c->nx, c->nu, c->N, c->dtype, c->policy, false, probe) && probe.model_kind == 3))
        return build_model_rt(c);
    return true;
}

bool build_model_shape(tmpc_ctx_impl *c)
{
    KernelInfo probe;
    if (lookup_kernel(c->nx, c->nu, c->N, c->dtype, c->policy, false, probe) && probe.model_kind == 3) {
        if (!build_model_rt(c)) return false;
        // the stand-alone step kernels and the plant step exist for the compiled shapes only
        const bool f32 = c->dtype == TMPC_F32;
        if (c->nx == 12 && c->nu == 4 && c->N == 10) f32 ? build_model_t<float, 12, 4, 10>(c) : build_model_t<double, 12, 4, 10>(c);
        if (c->nx == 4 && c->nu == 1 && c->N == 10) f32 ? build_model_t<float, 4, 1, 10>(c) : build_model_t<double, 4, 1, 10>(c);
        if (c->nx == 32 && c->nu == 8 && c->N == 50 && f32) build_model_t<float, 32, 8, 50>(c);
        return true;
    }
    if (c->nx == 32 && c->nu == 8 && c->dtype == TMPC_F32) {
        if (cudaSetDevice(c->device) != cudaSuccess) return false;
        if (!build_model_warp(c)) return false;
        build_model_t<float, 32, 8, 50>(c);   // step kernels
        return true;
    }
    if (c->dtype == TMPC_F32 && c->nx == 12 && c->nu == 4 && c->N == 10) build_model_f32<12, 4, 10>(c, c->model_f32);
    const bool f32 = c->dtype == TMPC_F32;
    if (c->nx == 12 && c->nu == 4 && c->N == 10) {
        f32 ? build_model_t<float, 12, 4, 10>(c) : build_model_t<double, 12, 4, 10>(c);
        return true;
    }
    if (c->nx == 4 && c->nu == 1 && c->N == 10) {
        f32 ? build_model_t<float, 4, 1, 10>(c) : build_model_t<double, 4, 1, 10>(c);
        return true;
    }
    return false;
}

size_t esize(const tmpc_ctx_impl *c) { return c->dtype == TMPC_F32 ? 4 : 8; }

struct DevArgs {  // type-erased tmpc::SolveArgs<T> (identical layout for float/double: pointers + int64)
    long long batch;
    const void *x0;
    const void *Xref;
    long long xref_stride;
    void *wd, *wy, *wg, *wv, *wz;
    void *x, *u;
    int *iter, *status;
    void *resid;
    unsigned long long *counter;
    unsigned long long *stats;
    unsigned *done;
    int done_shift;
    const void *sys;
    unsigned *gate;
    const unsigned *order;
    void *u0;
    void *scratch;
    int sc_ib, sc_wm, sc_chunks;
    int test_flags;
    const void *ixmin, *ixmax, *iumin, *iumax;
    const void *model_g;
    long long gate_split;
    int roll_steps, roll_step0;
    void *roll_x, *roll_u0;
    int *roll_iter, *roll_status;
    const void *roll_table;
    long long roll_rows;
    const int *roll_start;
    const void *pn_seed;
};
static_assert(sizeof(DevArgs) == sizeof(tmpc::SolveArgs<float>), "arg layout");
static_assert(sizeof(DevArgs) == sizeof(tmpc::SolveArgs<double>), "arg layout");

void *model_param(tmpc_ctx_impl *c, const KernelInfo &ki)
{
    if (ki.model_kind == 2) return (void *)&c->model_w;
    if (ki.model_kind == 3) return (void *)c->model_rt.data();
    return ki.model_kind == 1 ? (void *)c->model_f32.data() : (void *)c->model.data();
}

// Everything a launch of this ctx does on `s` from here on runs after the previous launch of the ctx has finished.
int order_after_previous(tmpc_ctx_impl *c, cudaStream_t s)
{
    if (c->launched) CUDA_TRY(c, cudaStreamWaitEvent(s, c->last_launch, 0));
    return TMPC_OK;
}
int mark_launch(tmpc_ctx_impl *c, cudaStream_t s)
{
    CUDA_TRY(c, cudaEventRecord(c->last_launch, s));
    c->launched = true;
    return TMPC_OK;
}

// Grid of the persistent kernel for `da.batch` instances; for the run-time-shape kernel also its per-lane scratch and
// the launch-dependent fields of its model image.
int plan_launch(tmpc_ctx_impl *c, const KernelInfo &ki, const DevArgs &da, cudaStream_t s, long long &blocks)
{
    blocks = (da.batch + ki.per_block - 1) / ki.per_block;
    long long max_blocks = std::max(1, c->sm_count - c->reserve_sms_next);
    c->reserve_sms_next = 0;
    if (ki.model_kind == 3) {
        // persistent grid = every block the SMs can hold; the per-lane scratch (state of the resident instances) is
        // capped at 16 GB of HBM
        if (!c->rt_ready) return fail(c, TMPC_ERR_STATE, "run-time-shape model image not built");
        int per_sm = 1;
        CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ki.fn, ki.block, ki.smem));
        if (per_sm < 1) per_sm = 1;
        if (const char *e = getenv("TMPC_RT_BLOCKS_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(e)));
        max_blocks = (long long)c->sm_count * per_sm;
        const long long per_block_bytes = tmpc::rt_scratch_elems(c->nx, c->nu, c->N) * ki.block * (long long)esize(c);
        const long long cap = (16LL << 30) / per_block_bytes;
        if (cap < 1) return fail(c, TMPC_ERR_UNSUPPORTED, "shape needs more than 16 GB of scratch for one block");
        if (max_blocks > cap) max_blocks = cap;
    }
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    if (ki.model_kind == 3) {
        const size_t need = (size_t)tmpc::rt_scratch_elems(c->nx, c->nu, c->N) * (size_t)blocks * ki.block * esize(c);
        if (c->d_rt_scratch_bytes < need) {
            CUDA_TRY(c, cudaDeviceSynchronize());   // kernels of earlier calls / chunks may still use the old area
            if (c->d_rt_scratch) cudaFree(c->d_rt_scratch);
            c->d_rt_scratch = nullptr; c->d_rt_scratch_bytes = 0;
            CUDA_TRY(c, cudaMalloc(&c->d_rt_scratch, need));
            c->d_rt_scratch_bytes = need;
        }
        if (c->dtype == TMPC_F32) {
            auto &m = *reinterpret_cast<tmpc::ModelRT<float> *>(c->model_rt.data());
            m.scratch = (float *)c->d_rt_scratch; m.lanes = blocks * ki.block; m.warm = da.wd ? 1 : 0;
            const bool xs = c->ib_batch && c->en_state, us = c->ib_batch && c->en_input;
            m.ixmin = xs ? (const float *)c->d_ib[0] : nullptr; m.ixmax = xs ? (const float *)c->d_ib[1] : nullptr;
            m.iumin = us ? (const float *)c->d_ib[2] : nullptr; m.iumax = us ? (const float *)c->d_ib[3] : nullptr;
        } else {
            auto &m = *reinterpret_cast<tmpc::ModelRT<double> *>(c->model_rt.data());
            m.scratch = (double *)c->d_rt_scratch; m.lanes = blocks * ki.block; m.warm = da.wd ? 1 : 0;
            const bool xs = c->ib_batch && c->en_state, us = c->ib_batch && c->en_input;
            m.ixmin = xs ? (const double *)c->d_ib[0] : nullptr; m.ixmax = xs ? (const double *)c->d_ib[1] : nullptr;
            m.iumin = us ? (const double *)c->d_ib[2] : nullptr; m.iumax = us ? (const double *)c->d_ib[3] : nullptr;
        }
    }
    (void)s;
    return TMPC_OK;
}

// fp32 12/4/10 kernel: which regions of the per-lane coalesced scratch this launch needs (tmpc_kernel_f32.cuh ScratchMap)
int plan_lane_scratch(tmpc_ctx_impl *c, const KernelInfo &ki, DevArgs &da, long long blocks, bool ib, cudaStream_t s)
{
    da.scratch = nullptr; da.sc_ib = da.sc_wm = -1; da.sc_chunks = 0; da.test_flags = 0;
    da.ixmin = da.ixmax = da.iumin = da.iumax = nullptr;
    da.model_g = nullptr;
    if (ki.model_kind != 1) {
        c->duals_zero_next = false;   // (the other kernels take their duals from the caller's buffers)
        if (ki.scratch_per_block) {   // eight-warp 32/8/50 kernel: g, v rows of two slots per warp
            const size_t need = (size_t)blocks * ki.scratch_per_block;
            if (c->d_lane_scratch_bytes < need) {
                CUDA_TRY(c, cudaDeviceSynchronize());
                if (c->d_lane_scratch) cudaFree(c->d_lane_scratch);
                c->d_lane_scratch = nullptr; c->d_lane_scratch_bytes = 0;
                CUDA_TRY(c, cudaMalloc(&c->d_lane_scratch, need));
                c->d_lane_scratch_bytes = need;
            }
            da.scratch = c->d_lane_scratch;
        }
        return TMPC_OK;
    }
    if (kernel_variant() == 3) {
        // (re-uploaded per launch: 3.9 KB on the launch stream, ordered before the kernel; the A/B variant is not tuned for launch cost)
        if (!c->d_model_f32) CUDA_TRY(c, cudaMalloc(&c->d_model_f32, c->model_f32.size()));
        CUDA_TRY(c, cudaMemcpyAsync(c->d_model_f32, c->model_f32.data(), c->model_f32.size(), cudaMemcpyHostToDevice, s));
        da.model_g = c->d_model_f32;
    }
    if (const char *e = getenv("TMPC_TEST_MIRROR")) da.test_flags = atoi(e) & 11;   // (8: fused closed loop, exact hand-over at every step)
    if (c->duals_zero_next && da.wd) da.test_flags |= 4;
    c->duals_zero_next = false;
    using SM = tmpc::ScratchMap<12, 4, 10>;
    int chunks = 0;
    if (ib) {
        da.sc_ib = chunks; chunks += SM::IB_CHUNKS;
        if (c->en_state) { da.ixmin = c->d_ib[0]; da.ixmax = c->d_ib[1]; }
        if (c->en_input) { da.iumin = c->d_ib[2]; da.iumax = c->d_ib[3]; }
    }
    if (da.wd) { da.sc_wm = chunks; chunks += SM::WM_CHUNKS * (da.roll_steps > 1 ? 2 : 1); }   // (a fused closed loop keeps two mirror areas)
    da.sc_chunks = chunks;
    if (!chunks) return TMPC_OK;
    const size_t need = (size_t)blocks * ki.block * chunks * 16;
    if (c->d_lane_scratch_bytes < need) {
        CUDA_TRY(c, cudaDeviceSynchronize());   // an earlier launch may still use the old area
        if (c->d_lane_scratch) cudaFree(c->d_lane_scratch);
        c->d_lane_scratch = nullptr; c->d_lane_scratch_bytes = 0;
        CUDA_TRY(c, cudaMalloc(&c->d_lane_scratch, need));
        c->d_lane_scratch_bytes = need;
    }
    da.scratch = c->d_lane_scratch;
    return TMPC_OK;
}

// Launch one persistent kernel (already chosen) for one device-resident batch on `s`.
int launch_kernel_info(tmpc_ctx_impl *c, const KernelInfo &ki, DevArgs &da, cudaStream_t s, bool time_it, bool ev0_done = false)
{
    {
        const int rc = order_after_previous(c, s);   // (a no-op wait when the caller already ordered this stream)
        if (rc != TMPC_OK) return rc;
    }
    CUDA_TRY(c, cudaFuncSetAttribute(ki.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ki.smem));
    CUDA_TRY(c, cudaMemsetAsync(c->d_counter, 0, 5 * sizeof(unsigned long long), s));
    da.counter = c->d_counter;
    da.stats = c->d_counter + 1;
    long long blocks = 1;
    {
        int rc = plan_launch(c, ki, da, s, blocks);
        if (rc == TMPC_OK) rc = plan_lane_scratch(c, ki, da, blocks, c->ib_batch != 0 && !da.sys, s);
        if (rc != TMPC_OK) return rc;
        if (da.roll_steps > 1 && ki.model_kind == 0 && getenv("TMPC_ROLL_ASYNC_REFILL")) da.test_flags |= 16;   // A/B of admm_kernel_small_roll
        if (da.sys) {   // refill policy of the row-pair systems kernel (tmpc_kernel_sys.cuh), tuning only
            if (const char *e = getenv("TMPC_SYS_REFILL_MIN")) da.test_flags |= (atoi(e) & 31) << 8;
            if (const char *e = getenv("TMPC_SYS_DEFER_MAX")) da.test_flags |= (atoi(e) & 15) << 16;
        }
    }
    void *params[2] = {model_param(c, ki), &da};
    if (time_it && !ev0_done) CUDA_TRY(c, cudaEventRecord(c->ev0, s));
    da.pn_seed = nullptr;
    if (ki.model_kind == 1 && da.xref_stride != 0 && !da.gate && !da.roll_table && da.batch >= 8192 && !getenv("TMPC_NO_PN_PREPASS")) {
        // per-instance reference trajectories, inputs resident: the p_N seeds in one coalesced pre-pass (part of the solve: inside the events)
        const size_t need = (size_t)da.batch * c->nx * sizeof(float);
        if (c->d_pn_seed_bytes < need) {
            CUDA_TRY(c, cudaDeviceSynchronize());
            if (c->d_pn_seed) cudaFree(c->d_pn_seed);
            c->d_pn_seed = nullptr; c->d_pn_seed_bytes = 0;
            CUDA_TRY(c, cudaMalloc(&c->d_pn_seed, need));
            c->d_pn_seed_bytes = need;
        }
        const void *xr = da.Xref;
        long long stride = da.xref_stride, nb = da.batch;
        void *out = c->d_pn_seed;
        void *pp[5] = {c->model_f32.data(), (void *)&xr, &stride, &nb, &out};
        CUDA_TRY(c, cudaLaunchKernel(tmpc_dispatch::lookup_f32_pn_seed(c->policy), dim3((unsigned)((nb + 127) / 128)), dim3(128), pp, 0, s));
        da.pn_seed = c->d_pn_seed;
        c->stats.launches += 1;
    }
    CUDA_TRY(c, cudaLaunchKernel(ki.fn, dim3((unsigned)blocks), dim3(ki.block), params, ki.smem, s));
    if (time_it) CUDA_TRY(c, cudaEventRecord(c->ev1, s));
    {
        const int rc = mark_launch(c, s);
        if (rc != TMPC_OK) return rc;
    }
    c->stats.launches += 1;
    c->stats.lanes = (int32_t)(blocks * ki.per_block);
    c->stats.pattern = c->pattern;
    return TMPC_OK;
}

// Launch the persistent kernel for one device-resident batch on `s`.
// ---------------------------------------------------------------------------------------------
// Longest-expected-first schedule.  The persistent kernels hand instances to lanes in claim order; iteration counts
// range from a handful to max_iter, so with instances claimed in index order the launch ends with a tail in which
// a few lanes finish 100-iteration instances while the rest of the GPU idles (simulated on the hover workload: makespan
// 7.9 % above the mean lane load in index order, 1.1 % with this schedule).  How long an instance runs is governed by how
// hard the box constraints bite, and the saturation of the unconstrained LQR input, max_r |Kinf (x0 - Xref_0)|_r, ranks
// it well (Spearman 0.83 against the true iteration count).  Pre-pass on the solve's stream: one key per instance,
// CUB radix sort of (key, index) on the upper 16 key bits (sign, exponent, 7 mantissa bits), descending; the kernels then claim order[k].  Results are
// unaffected (instances are independent); only used for device-resident batches without completion counters, where
// completion in index order does not matter.  Measured on B200 (profiles/r01_lpt_schedule.log): pre-pass 66 us per 1M
// instances (key kernel 23 us + sort 43 us); 1M hover instances 10.93 -> 10.40 ms, 4M cartpole 12.39 -> 11.49 ms, 32/8/50 warm
// re-solve 19.7 -> 18.0 ms; a workload without a tail pays the pre-pass (hover at mult 0.1: 5.29 -> 5.41 ms).
template <class T>
__global__ void lpt_key_kernel(long long batch, int nx, int nu, const T *__restrict__ K0, long long k_stride, const T *__restrict__ x0,
                               const T *__restrict__ Xref, long long xref_stride, unsigned *__restrict__ keys, unsigned *__restrict__ vals)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= batch) return;
    const T *x = x0 + i * nx, *xr = Xref + i * xref_stride;
    const T *K = K0 + i * k_stride;   // k_stride = 0: the shared Kinf; else the instance's own (per-instance systems)
    float m = 0.f;
    // 8 rows of Kinf at a time in registers, every x0 / Xref element read once per row block (16-byte loads when the rows
    // are 16-byte aligned, i.e. nx a multiple of the vector width: the API requires 16-byte aligned device buffers)
    constexpr int VEC = 16 / (int)sizeof(T);
    const bool vec = (nx % VEC) == 0 && (xref_stride % VEC) == 0;
    for (int r0 = 0; r0 < nu; r0 += 8) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        auto col = [&](int k, float d) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r0 + r < nu) acc[r] = fmaf((float)__ldg(K + r0 + r + k * nu), d, acc[r]);
        };
        if (vec) {
            using V = typename tmpc::Num<T>::vec_t;
            for (int k = 0; k < nx; k += VEC) {
                const V a = __ldg(reinterpret_cast<const V *>(x + k)), b = __ldg(reinterpret_cast<const V *>(xr + k));
#pragma unroll
                for (int e = 0; e < VEC; ++e) col(k + e, (float)(((const T *)&a)[e] - ((const T *)&b)[e]));
            }
        } else {
            for (int k = 0; k < nx; ++k) col(k, (float)(__ldg(x + k) - __ldg(xr + k)));
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (r0 + r < nu) m = fmaxf(m, fabsf(acc[r]));
    }
    keys[i] = __float_as_uint(m);   // m >= 0: the bit pattern is monotone in m
    vals[i] = (unsigned)i;
}

bool lpt_wanted(const tmpc_ctx_impl *c, const KernelInfo &ki, const DevArgs &da)
{
    const char *e = getenv("TMPC_LPT");
    if (e && !strcmp(e, "0")) return false;
    if (da.done || da.gate || !c->d_kinf || da.batch >= (1LL << 31)) return false;
    if (e && !strcmp(e, "1")) return true;
    // Per-instance reference trajectories (tracking): measured 3 % SLOWER with the schedule (7.02 -> 7.35 ms per 1M launch):
    // x0 = Xref_0 + noise there, so the key ranks nothing, and the permuted order scatters the per-iteration Xref reads.
    if (da.xref_stride != 0) return false;
    // Warm starts stream 2.9 KB of state per instance in and out of [instance]-major buffers; with thread-per-instance
    // kernels the permuted order turns that into scattered 144..480-byte rows (measured: device rollout of 1M hover instances
    // 7.85 -> 9.65 ms per MPC step).  The warp-per-instance kernel moves 17 KB contiguous per instance and gains (19.7 -> 18.0 ms).
    // (a fused closed loop streams the state once per `roll_steps` solves and its claims are that much longer: there the order pays)
    if (da.wd && ki.model_kind != 2 && da.roll_steps <= 1) return false;
    if (da.roll_steps > 1 && ki.model_kind != 1) return false;   // (4/1/10 closed loops: a handful of iterations per step, uniform claims)
    const long long lanes = (long long)ki.per_block * c->sm_count;
    return da.batch >= 2 * lanes && da.batch >= 16384;
}

__global__ void iota_kernel(unsigned *__restrict__ out, long long n, unsigned offset)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = offset + (unsigned)i;
}

// K / k_stride: Kinf (column-major) of instance i at K + i * k_stride; default = the ctx's shared Kinf.
// tail > 0 (host pipeline): only the LEADING `tail` instances are ranked; the claim order becomes
// [tail, tail+1, ..., batch-1, the leading `tail` instances longest-expected-first].
// part: 0 = everything; 1 = buffers + the index-ordered head of the order only (before the kernel starts); 2 = rank the tail only
int lpt_prepare(tmpc_ctx_impl *c, DevArgs &da, cudaStream_t s, const void *K = nullptr, long long k_stride = 0, long long tail = 0, int part = 0)
{
    if (!K) { K = c->d_kinf; k_stride = 0; }
    const size_t B = (size_t)da.batch;
    if (c->lpt_cap < B) {
        CUDA_TRY(c, cudaDeviceSynchronize());
        if (c->lpt_buf) cudaFree(c->lpt_buf);
        if (c->lpt_temp) cudaFree(c->lpt_temp);
        c->lpt_buf = nullptr; c->lpt_temp = nullptr; c->lpt_cap = 0; c->lpt_temp_bytes = 0;
        CUDA_TRY(c, cudaMalloc((void **)&c->lpt_buf, 4 * B * sizeof(unsigned)));
        size_t tb = 0;
        CUDA_TRY(c, cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, (const unsigned *)nullptr, (unsigned *)nullptr,
                                                               (const unsigned *)nullptr, (unsigned *)nullptr, (int)B, 0, 32, s));
        CUDA_TRY(c, cudaMalloc(&c->lpt_temp, tb));
        c->lpt_temp_bytes = tb;
        c->lpt_cap = B;
    }
    unsigned *keys_in = c->lpt_buf, *keys_out = keys_in + c->lpt_cap, *vals_in = keys_out + c->lpt_cap, *vals_out = vals_in + c->lpt_cap;
    const size_t n_rank = tail > 0 ? (size_t)tail : B;
    unsigned *sorted_out = vals_out + (B - n_rank);
    if (n_rank < B && part != 2) {
        iota_kernel<<<(unsigned)((B - n_rank + 255) / 256), 256, 0, s>>>(vals_out, (long long)(B - n_rank), (unsigned)n_rank);
        CUDA_TRY(c, cudaGetLastError());
    }
    da.order = vals_out;
    if (part == 1) return TMPC_OK;
    const unsigned blocks = (unsigned)((n_rank + 255) / 256);
    if (c->dtype == TMPC_F32)
        lpt_key_kernel<float><<<blocks, 256, 0, s>>>((long long)n_rank, c->nx, c->nu, (const float *)K, k_stride, (const float *)da.x0, (const float *)da.Xref,
                                                     da.xref_stride, keys_in, vals_in);
    else
        lpt_key_kernel<double><<<blocks, 256, 0, s>>>((long long)n_rank, c->nx, c->nu, (const double *)K, k_stride, (const double *)da.x0,
                                                      (const double *)da.Xref, da.xref_stride, keys_in, vals_in);
    CUDA_TRY(c, cudaGetLastError());
    size_t tb = c->lpt_temp_bytes;
    CUDA_TRY(c, cub::DeviceRadixSort::SortPairsDescending(c->lpt_temp, tb, (const unsigned *)keys_in, keys_out, (const unsigned *)vals_in, sorted_out,
                                                           (int)n_rank, 16, 32, s));
    c->stats.launches += 1;   // the key kernel (the sort's kernels are CUB's)
    return TMPC_OK;
}

// per-instance bounds run on the specialised kernel where one exists with the IB option (fp32 12/4/10, TMEM variant)
bool ib_on_f32_kernel(const tmpc_ctx_impl *c)
{
    return c->nx == 12 && c->nu == 4 && c->N == 10 && c->dtype == TMPC_F32 && kernel_variant() == 2 && !force_rt() && !getenv("TMPC_IB_RT");
}

// Claim order from a per-instance integer key the caller already has on the device (tmpc_batch: the iteration counts of the
// previous closed-loop step predict this step's almost perfectly): instances sorted by key, largest first.
int lpt_from_keys(tmpc_ctx_impl *c, DevArgs &da, cudaStream_t s, const unsigned *keys, int key_bits)
{
    const size_t B = (size_t)da.batch;
    if (c->lpt_cap < B) {
        CUDA_TRY(c, cudaDeviceSynchronize());
        if (c->lpt_buf) cudaFree(c->lpt_buf);
        if (c->lpt_temp) cudaFree(c->lpt_temp);
        c->lpt_buf = nullptr; c->lpt_temp = nullptr; c->lpt_cap = 0; c->lpt_temp_bytes = 0;
        CUDA_TRY(c, cudaMalloc((void **)&c->lpt_buf, 4 * B * sizeof(unsigned)));
        size_t tb = 0;
        CUDA_TRY(c, cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, (const unsigned *)nullptr, (unsigned *)nullptr,
                                                               (const unsigned *)nullptr, (unsigned *)nullptr, (int)B, 0, 32, s));
        CUDA_TRY(c, cudaMalloc(&c->lpt_temp, tb));
        c->lpt_temp_bytes = tb;
        c->lpt_cap = B;
    }
    unsigned *keys_out = c->lpt_buf + c->lpt_cap, *vals_in = keys_out + c->lpt_cap, *vals_out = vals_in + c->lpt_cap;
    iota_kernel<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(vals_in, (long long)B, 0u);
    CUDA_TRY(c, cudaGetLastError());
    size_t tb = c->lpt_temp_bytes;
    CUDA_TRY(c, cub::DeviceRadixSort::SortPairsDescending(c->lpt_temp, tb, keys, keys_out, (const unsigned *)vals_in, vals_out, (int)B, 0, key_bits, s));
    da.order = vals_out;
    c->stats.launches += 1;
    return TMPC_OK;
}

// Fused closed loop (ROLL instances of the fp32 12/4/10 kernel): PARITY order, tensor-memory variant, one box for the batch.
// TMPC_ROLL=0 keeps one launch per MPC step.
bool roll_supported(const tmpc_ctx_impl *c)
{
    const char *e = getenv("TMPC_ROLL");
    if (e && !strcmp(e, "0")) return false;
    if (c->dtype != TMPC_F32 || c->policy != TMPC_ORDER_PARITY || force_rt() || c->ib_batch) return false;
    if (c->nx == 12 && c->nu == 4 && c->N == 10) return kernel_variant() == 2;
    if (c->nx == 4 && c->nu == 1 && c->N == 10) {   // register-resident kernel (tmpc_kernel_small.cuh), not with the generic development variants
        const char *k = getenv("TMPC_KERNEL");
        return !k || !strncmp(k, "small", 5);
    }
    return false;
}

int launch_device(tmpc_ctx_impl *c, DevArgs &da, bool warm, cudaStream_t s, bool time_it)
{
    KernelInfo ki;
    const bool ib_f32 = c->ib_batch && ib_on_f32_kernel(c);
    if (da.roll_steps > 1) {
        // fused closed loop (tmpc_batch_rollout): the caller has checked roll_supported()
        bool ok = warm && da.wd && !c->ib_batch && !da.sys;
        if (ok && c->nx == 4) {
            const char *k = getenv("TMPC_KERNEL");
            // 256 lanes per SM: the only size at which the unrolled sweeps keep every register (measured: 16.7 ms per 8 steps x 16.7M
            // instances; 384 lanes spill 450 B: 19.6 ms)
            ok = tmpc_dispatch::lookup_small_roll(k && !strcmp(k, "small384") ? 384 : k && !strcmp(k, "small512") ? 512 : 256, ki);
        } else if (ok) ok = tmpc_dispatch::lookup_f32_roll(c->pattern, c->const_bounds, ki);
        if (!ok) return fail(c, TMPC_ERR_UNSUPPORTED, "fused closed loop: no kernel");
    } else if (ib_f32) {
        // per-instance bounds on the specialised fp32 12/4/10 kernel: each lane copies its instance's box into its coalesced
        // scratch rows at refill and projects onto it (IB instances of the kernel)
        if (da.sys) return fail(c, TMPC_ERR_UNSUPPORTED, "per-instance bounds with per-instance systems");
        if (da.batch != c->ib_batch) return fail(c, TMPC_ERR_INVALID, "batch differs from the batch of tmpc_set_instance_bounds");
        if (!tmpc_dispatch::lookup_f32(c->policy, warm, c->pattern, false, 2, ki, true)) return fail(c, TMPC_ERR_UNSUPPORTED, "no kernel for this shape");
    } else if (c->ib_batch) {
        // other shapes: the run-time-shape kernel reads each instance's own rows (index order, no scheduling pre-pass)
        if (da.sys) return fail(c, TMPC_ERR_UNSUPPORTED, "per-instance bounds with per-instance systems");
        if (da.batch != c->ib_batch) return fail(c, TMPC_ERR_INVALID, "batch differs from the batch of tmpc_set_instance_bounds");
        if (!c->rt_ready || !lookup_kernel_rt(c->nx, c->nu, c->N, c->dtype, c->policy, ki))
            return fail(c, TMPC_ERR_UNSUPPORTED, "no kernel for this shape");
    } else if (!lookup_kernel(c->nx, c->nu, c->N, c->dtype, c->policy, warm, ki, c->pattern | (c->const_bounds ? 0x100 : 0)))
        return fail(c, TMPC_ERR_UNSUPPORTED, "no kernel for this shape");
    bool ev0_done = false;
    c->lpt_used = da.order ? 2 : 0;   // 2: the caller (host pipeline, closed-loop batch) brought its own claim order
    {
        const int rc = order_after_previous(c, s);   // the scheduling buffers may still be read by the previous launch
        if (rc != TMPC_OK) return rc;
    }
    if (!da.sys && (!c->ib_batch || ib_f32) && !da.order && lpt_wanted(c, ki, da)) {
        // the pre-pass is part of the solve: it runs on the same stream inside the timed region
        if (time_it) { CUDA_TRY(c, cudaEventRecord(c->ev0, s)); ev0_done = true; }
        const int rc = lpt_prepare(c, da, s);
        if (rc != TMPC_OK) return rc;
        c->lpt_used = 1;
    }
    return launch_kernel_info(c, ki, da, s, time_it, ev0_done);
}

int ensure_stage(tmpc_ctx_impl *c, int k, size_t in_bytes, size_t out_bytes)
{
    auto &st = c->stage[k];
    if (!st.s) CUDA_TRY(c, cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
    if (!st.done) CUDA_TRY(c, cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming));
    if (st.in_bytes < in_bytes) {
        if (st.h_in) cudaFreeHost(st.h_in);
        if (st.d_in) cudaFree(st.d_in);
        st.h_in = st.d_in = nullptr;
        CUDA_TRY(c, cudaMallocHost(&st.h_in, in_bytes));
        CUDA_TRY(c, cudaMalloc(&st.d_in, in_bytes));
        st.in_bytes = in_bytes;
    }
    if (st.out_bytes < out_bytes) {
        if (st.h_out) cudaFreeHost(st.h_out);
        if (st.d_out) cudaFree(st.d_out);
        st.h_out = st.d_out = nullptr;
        CUDA_TRY(c, cudaMallocHost(&st.h_out, out_bytes));
        CUDA_TRY(c, cudaMalloc(&st.d_out, out_bytes));
        st.out_bytes = out_bytes;
    }
    return TMPC_OK;
}


typedef int (*wait_value32_fn)(cudaStream_t, unsigned long long /*CUdeviceptr*/, unsigned, unsigned);
typedef int (*write_value32_fn)(cudaStream_t, unsigned long long /*CUdeviceptr*/, unsigned, unsigned);

// Cold solve from/to HOST memory: ONE persistent-kernel launch over the whole batch, outputs copied back while the
// kernel is still running.  The kernel bumps done[inst >> shift] after the last store of each instance; the copy
// stream waits on those counters with stream memory operations (cuStreamWaitValue32 GEQ) and then DMAs that
// chunk's slices straight into the caller's buffers.  No host thread is involved between launch and the final sync.
int launch_device(tmpc_ctx_impl *c, DevArgs &da, bool warm, cudaStream_t s, bool time_it);

int solve_host_gated(tmpc_ctx_impl *c, const tmpc_solve_args *a)
{
    const size_t es = esize(c);
    const int nx = c->nx, nu = c->nu, N = c->N;
    const size_t xrow = (size_t)nx * N, urow = (size_t)nu * (N - 1);
    const int64_t B = a->batch;
    static wait_value32_fn wait32 = nullptr;
    static write_value32_fn write32 = nullptr;
    static bool wait32_probed = false;
    if (!wait32_probed) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn &&
            qr == cudaDriverEntryPointSuccess)
            wait32 = reinterpret_cast<wait_value32_fn>(fn);
        else
            cudaGetLastError();
        fn = nullptr;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn &&
            qr == cudaDriverEntryPointSuccess)
            write32 = reinterpret_cast<write_value32_fn>(fn);
        else
            cudaGetLastError();
        wait32_probed = true;
    }
    // TMPC_NO_STREAM_MEMOPS=1: behave as on a driver without stream memory operations (plain copy after the kernel, inputs
    // complete before the launch) -- the fallback is exercised by the tests through this switch
    const bool memops = !getenv("TMPC_NO_STREAM_MEMOPS");
    const wait_value32_fn waitf = memops ? wait32 : nullptr;
    const write_value32_fn writef = memops ? write32 : nullptr;
    // 65,536 instances per D2H chunk.  The read-back is PCIe-bound from its first byte (the kernel produces 57 GB/s of outputs,
    // PCIe carries 56): smaller chunks (measured 16K: +5 %, 4K: +35 % time) or small leading chunks (+1 %) only add per-copy cost.
    int shift = 16;
    if (const char *e = getenv("TMPC_D2H_SHIFT")) shift = std::max(10, std::min(20, atoi(e)));
    while (shift > 10 && (B >> shift) < 8) --shift;   // small batches: at least ~8 chunks, >= 1024 instances each
    const int64_t CH = (int64_t)1 << shift;
    const int nch = (int)((B + CH - 1) / CH);
    // device image: in = x0 | [Xref];  out = whichever of x | u | u0 | iter | status | resid the caller asked for
    const size_t in_bytes = (size_t)B * nx * es + (a->xref_shared ? xrow * es : (size_t)B * xrow * es);
    auto a16 = [](size_t v) { return (v + 15) & ~size_t(15); };
    size_t o = 0;
    auto region = [&](const void *want, size_t per) -> size_t { const size_t at = o; if (want) o = a16(o + (size_t)B * per); return at; };
    const size_t o_x = region(a->x, xrow * es), o_u = region(a->u, urow * es), o_u0 = region(a->u0, nu * es);
    const size_t o_it = region(a->iter, 4), o_st = region(a->status, 4), o_rs = region(a->resid, 4 * es);
    const size_t out_bytes = o + 64;
    if (c->g_in_bytes < in_bytes) {
        if (c->g_in) cudaFree(c->g_in);
        c->g_in = nullptr; c->g_in_bytes = 0;
        CUDA_TRY(c, cudaMalloc(&c->g_in, in_bytes));
        c->g_in_bytes = in_bytes;
    }
    if (c->g_out_bytes < out_bytes) {
        if (c->g_out) cudaFree(c->g_out);
        c->g_out = nullptr; c->g_out_bytes = 0;
        CUDA_TRY(c, cudaMalloc(&c->g_out, out_bytes));
        c->g_out_bytes = out_bytes;
    }
    if (c->g_done_n < (size_t)nch) {
        if (c->g_done) cudaFree(c->g_done);
        c->g_done = nullptr; c->g_done_n = 0;
        CUDA_TRY(c, cudaMalloc((void **)&c->g_done, sizeof(unsigned) * nch));
        c->g_done_n = nch;
    }
    if (!c->g_copy) CUDA_TRY(c, cudaStreamCreateWithFlags(&c->g_copy, cudaStreamNonBlocking));
    if (!c->g_h2d) CUDA_TRY(c, cudaEventCreateWithFlags(&c->g_h2d, cudaEventDisableTiming));
    cudaStream_t s = c->stream;
    char *din = (char *)c->g_in;
    char *d_x0 = din, *d_xref = din + (size_t)B * nx * es;
    DevArgs da{};
    da.batch = B; da.x0 = d_x0; da.Xref = d_xref; da.xref_stride = a->xref_shared ? 0 : (long long)xrow;
    // Inputs: the H2D of x0 (and of per-instance Xref) is OVERLAPPED with the kernel.  It runs in 131,072-instance chunks on
    // its own stream; after each chunk a stream memory operation advances an arrival counter; the kernel starts as soon
    // as the first chunk is in and its lanes check the counter before touching an instance (gate_wait).  Everything is
    // enqueued before the kernel launch, so even a staged (pageable) copy cannot wait on the kernel.
    const bool overlap = writef != nullptr && B < (int64_t)0xffffffffu && !getenv("TMPC_NO_H2D_OVERLAP");
    if (!c->g_in_stream) CUDA_TRY(c, cudaStreamCreateWithFlags(&c->g_in_stream, cudaStreamNonBlocking));
    if (!c->g_first) CUDA_TRY(c, cudaEventCreateWithFlags(&c->g_first, cudaEventDisableTiming));
    if (!c->g_gate) CUDA_TRY(c, cudaMalloc((void **)&c->g_gate, 4 * sizeof(unsigned)));
    cudaStream_t hs = c->g_in_stream;
    {
        // the previous launch of this ctx may still read the input image, the claim order or the gate
        const int rc = order_after_previous(c, hs);
        if (rc != TMPC_OK) return rc;
    }
    CUDA_TRY(c, cudaMemsetAsync(c->g_gate, 0, 4 * sizeof(unsigned), hs));
    if (a->xref_shared) CUDA_TRY(c, cudaMemcpyAsync(d_xref, a->Xref, xrow * es, cudaMemcpyHostToDevice, hs));
    // Tail-sorted schedule.  Claimed in index order, a launch ends with a tail in which a few lanes finish max_iter-long
    // instances while the rest of the GPU idles (+7.5 % on the hover workload); the full longest-expected-first order of the
    // device path needs every x0 before the first claim, i.e. no H2D overlap.  Middle way: a LEADING segment of the batch is
    // ranked by the same key (lpt_prepare) and claimed LAST, longest first; the rest streams in behind the running kernel
    // and is claimed in index order.  Simulated on the hover workload's real iteration counts: makespan 1.018 x the mean
    // lane load with a quarter of the batch ranked, 1.013 with half (index order 1.075, fully sorted 1.013).
    //   late (default): the persistent kernel is launched on all SMs BUT TWO as soon as a first small chunk of the
    //     index-ordered part is in; the ranked segment (a quarter of the batch) goes over LAST and its key kernel + radix sort run
    //     beside the solver on the two free SMs (a full grid leaves them no register file); the lanes reach those claims
    //     milliseconds after the ranking is done (gate[2]).  Nothing but 32,768 instances of H2D precedes the launch.
    //     Used when the caller wants neither x nor u back (controls-only: u0 / iter / status): measured 10.77 ms per 1M-instance
    //     step against 10.84 ms tail-first on one GPU, and the gap grows when eight ranks share the host's H2D path.
    //   first: the ranked segment (a quarter) goes over FIRST and is ranked before the launch, full grid.  Used for
    //     trajectory outputs: those calls are bound by the D2H of 648 B per instance, and the ranked segment's outputs only
    //     complete at the very end of the kernel -- with half the batch ranked that read-back (340 MB) would trail the
    //     kernel instead of hiding behind it (measured 16.9 vs 14.3 ms per step).
    //   TMPC_TAIL_FIRST=1 / TMPC_TAIL_LATE=1 force one or the other.
    int64_t T0 = 0;
    const bool late = getenv("TMPC_TAIL_LATE") ? true : getenv("TMPC_TAIL_FIRST") ? false : (!a->x && !a->u);
    {
        KernelInfo ki;
        const char *e = getenv("TMPC_LPT");
        const bool off = e && !strcmp(e, "0");
        if (!off && overlap && a->xref_shared && !c->ib_batch && c->d_kinf && nch >= 8 &&
            lookup_kernel(c->nx, c->nu, c->N, c->dtype, c->policy, false, ki, c->pattern | (c->const_bounds ? 0x100 : 0)) &&
            B >= 2 * (int64_t)ki.per_block * c->sm_count && B >= 16384) {
            int div = 4;   // share of the batch that is ranked: 1 / div  (TMPC_TAIL_DIV: tuning; measured in late mode, one GPU, 1M hover
                           // instances, controls-only: 2 -> 10.80 ms, 3 -> 10.78, 4 -> 10.71, 8 -> 10.95; the ranked segment's
                           // read-back trails the kernel, which costs eight ranks sharing one host D2H path more than one)
            if (const char *dv = getenv("TMPC_TAIL_DIV")) div = std::max(2, std::min(64, atoi(dv)));
            T0 = ((B / div + CH - 1) / CH) * CH;
        }
    }
    const int64_t ICH = overlap ? 131072 : B;
    const bool gate_stall_test = getenv("TMPC_TEST_GATE_STALL") != nullptr;
    auto h2d = [&](int64_t b0, int64_t n) -> int {
        CUDA_TRY(c, cudaMemcpyAsync(d_x0 + b0 * nx * es, (const char *)a->x0 + b0 * nx * es, (size_t)n * nx * es, cudaMemcpyHostToDevice, hs));
        if (!a->xref_shared)
            CUDA_TRY(c, cudaMemcpyAsync(d_xref + b0 * xrow * es, (const char *)a->Xref + b0 * xrow * es, (size_t)n * xrow * es,
                                        cudaMemcpyHostToDevice, hs));
        return TMPC_OK;
    };
    int64_t fed = 0;
    if (T0 > 0 && !late) {
        int rc = h2d(0, T0);
        if (rc != TMPC_OK) return rc;
        if ((rc = lpt_prepare(c, da, hs, nullptr, 0, T0)) != TMPC_OK) return rc;
        if (writef(hs, (unsigned long long)(uintptr_t)c->g_gate, (unsigned)T0, 0u) != 0) return fail(c, TMPC_ERR_CUDA, "cuStreamWriteValue32 failed");
        CUDA_TRY(c, cudaEventRecord(c->g_first, hs));
        fed = T0;
    }
    if (T0 > 0 && late) {
        const int rc = lpt_prepare(c, da, hs, nullptr, 0, T0, 1);   // the index-ordered head of the claim order, before the launch
        if (rc != TMPC_OK) return rc;
        da.gate_split = T0;
        fed = T0;
    }
    bool first_chunk = true;
    for (int64_t b0 = fed; b0 < B;) {
        // (late: a small first chunk, so that the kernel starts at once; PCIe then delivers ~6e8 instances/s to a kernel that
        //  consumes 1e8)
        const int64_t n = std::min<int64_t>((late && T0 > 0 && first_chunk) ? 32768 : ICH, B - b0);
        const int rc = h2d(b0, n);
        if (rc != TMPC_OK) return rc;
        // TMPC_TEST_GATE_STALL=1 (tests): the last arrival is never announced, so the lanes that claim those instances give up
        if (overlap && !(gate_stall_test && b0 + n >= B) &&
            writef(hs, (unsigned long long)(uintptr_t)c->g_gate, (unsigned)(b0 + n), 0u) != 0)
            return fail(c, TMPC_ERR_CUDA, "cuStreamWriteValue32 failed");
        if (b0 == 0 || (late && T0 > 0 && first_chunk)) CUDA_TRY(c, cudaEventRecord(c->g_first, hs));
        first_chunk = false;
        b0 += n;
    }
    if (T0 > 0 && late) {
        // the ranked segment: over the bus last, ranked on the SMs the solver leaves free, announced through gate[2]
        int rc = h2d(0, T0);
        if (rc == TMPC_OK) rc = lpt_prepare(c, da, hs, nullptr, 0, T0, 2);
        if (rc != TMPC_OK) return rc;
        if (!gate_stall_test && writef(hs, (unsigned long long)(uintptr_t)(c->g_gate + 2), 1u, 0u) != 0)
            return fail(c, TMPC_ERR_CUDA, "cuStreamWriteValue32 failed");
        c->reserve_sms_next = 2;
        if (const char *e = getenv("TMPC_SORT_SMS")) c->reserve_sms_next = std::max(1, std::min(16, atoi(e)));
    }
    if (!overlap) CUDA_TRY(c, cudaEventRecord(c->g_first, hs));   // everything must be in before the kernel starts
    {
        const int rc = order_after_previous(c, s);   // ... and the completion counters / output image
        if (rc != TMPC_OK) return rc;
    }
    CUDA_TRY(c, cudaMemsetAsync(c->g_done, 0, sizeof(unsigned) * nch, s));
    CUDA_TRY(c, cudaStreamWaitEvent(s, c->g_first, 0));
    // the copy stream must not evaluate its waits against counters left by a previous solve
    CUDA_TRY(c, cudaEventRecord(c->g_h2d, s));
    CUDA_TRY(c, cudaStreamWaitEvent(c->g_copy, c->g_h2d, 0));
    char *dout = (char *)c->g_out;
    char *d_x = dout + o_x, *d_u = dout + o_u, *d_u0 = dout + o_u0, *d_it = dout + o_it, *d_st = dout + o_st, *d_rs = dout + o_rs;
    da.x = a->x ? d_x : nullptr; da.u = a->u ? d_u : nullptr; da.u0 = a->u0 ? d_u0 : nullptr;
    da.iter = a->iter ? (int *)d_it : nullptr; da.status = a->status ? (int *)d_st : nullptr; da.resid = a->resid ? d_rs : nullptr;
    da.done = waitf ? c->g_done : nullptr;
    da.done_shift = shift;
    da.gate = overlap ? c->g_gate : nullptr;
    int rc = launch_device(c, da, false, s, true);
    if (rc != TMPC_OK) return rc;
    cudaStream_t cs = c->g_copy;
    if (waitf) {
        // Whatever happened inside the kernel (a lane that gave up at the input gate never reports its instance), the copy
        // stream must drain: once the kernel is over every completion counter is released.
        // (0x7f7f7f7f: cuStreamWaitValue32 GEQ is the cyclic comparison (int32)(*addr - value) >= 0, so "all ones" would read as -1)
        CUDA_TRY(c, cudaMemsetAsync(c->g_done, 0x7f, sizeof(unsigned) * nch, s));
    } else {                                            // no stream mem-ops available: plain copy after the kernel
        CUDA_TRY(c, cudaEventRecord(c->g_h2d, s));
        CUDA_TRY(c, cudaStreamWaitEvent(cs, c->g_h2d, 0));
    }
    // chunks in the order they complete: the index-ordered part first, the tail-sorted leading part last
    const int k_first = (int)(T0 / CH);
    for (int kk = 0; kk < nch; ++kk) {
        const int k = (kk + k_first) % nch;
        const int64_t b0 = (int64_t)k * CH, n = std::min<int64_t>(CH, B - b0);
        if (waitf) {
            int e = waitf(cs, (unsigned long long)(uintptr_t)(c->g_done + k), (unsigned)n, 0u /*CU_STREAM_WAIT_VALUE_GEQ*/);
            if (e != 0) return fail(c, TMPC_ERR_CUDA, "cuStreamWaitValue32 failed");
        }
        auto back = [&](void *dst, const char *src, size_t per) {
            if (dst) cudaMemcpyAsync((char *)dst + b0 * per, src + b0 * per, n * per, cudaMemcpyDeviceToHost, cs);
        };
        back(a->x, d_x, xrow * es); back(a->u, d_u, urow * es); back(a->u0, d_u0, nu * es); back(a->iter, d_it, 4);
        back(a->status, d_st, 4); back(a->resid, d_rs, 4 * es);
    }
    CUDA_TRY(c, cudaStreamSynchronize(cs));
    CUDA_TRY(c, cudaStreamSynchronize(s));
    CUDA_TRY(c, cudaStreamSynchronize(hs));
    if (overlap) {
        unsigned g[2] = {0, 0};
        CUDA_TRY(c, cudaMemcpy(g, c->g_gate, sizeof g, cudaMemcpyDeviceToHost));
        if (g[1]) return fail(c, TMPC_ERR_CUDA, "input gate timed out: the overlapped H2D did not deliver an instance's inputs within 2 s; outputs are incomplete");
    }
    c->stats_pending = true;
    return TMPC_OK;
}

bool is_pinned(const void *p)
{
    if (!p) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

template <class T, int NX, int NU, int NH>
cudaError_t launch_step(tmpc_ctx_impl *c, const tmpc::StepArgs<T> &sa, int which, cudaStream_t s)
{
    const int threads = 64;
    const unsigned blocks = (unsigned)((sa.batch + threads - 1) / threads);
    const tmpc::Model<T, NX, NU, NH> *m = reinterpret_cast<const tmpc::Model<T, NX, NU, NH> *>(c->model.data());
    if (c->policy == TMPC_ORDER_PARITY) tmpc::step_kernel<T, NX, NU, NH, false><<<blocks, threads, 0, s>>>(*m, sa, which);
    else tmpc::step_kernel<T, NX, NU, NH, true><<<blocks, threads, 0, s>>>(*m, sa, which);
    return cudaGetLastError();
}

}  // namespace

extern "C" {

const char *tmpc_version(void) { return "tmpc 0.2 (sm_100a)"; }

int tmpc_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *tmpc_last_error(const tmpc_ctx *ctx)
{
    return ctx ? reinterpret_cast<const tmpc_ctx_impl *>(ctx)->err.c_str() : g_create_error.c_str();
}

int tmpc_create(tmpc_ctx **out, int device, int nx, int nu, int N, int dtype, int order_policy)
{
    if (!out) return fail(nullptr, TMPC_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (dtype != TMPC_F32 && dtype != TMPC_F64) return fail(nullptr, TMPC_ERR_INVALID, "bad dtype");
    if (order_policy != TMPC_ORDER_PARITY && order_policy != TMPC_ORDER_FAST)
        return fail(nullptr, TMPC_ERR_INVALID, "bad order policy");
    KernelInfo ki;
    if (!lookup_kernel(nx, nu, N, dtype, order_policy, false, ki)) {
        char b[160];
        snprintf(b, sizeof b, "shape nx=%d nu=%d N=%d is outside what the kernels cover (1 <= nx, nu <= 64, N >= 2)", nx, nu, N);
        return fail(nullptr, TMPC_ERR_UNSUPPORTED, b);
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, TMPC_ERR_CUDA,
                    std::string("no CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, TMPC_ERR_INVALID, "bad device index");
    tmpc_ctx_impl *c = new tmpc_ctx_impl;
    c->device = device; c->nx = nx; c->nu = nu; c->N = N; c->dtype = dtype; c->policy = order_policy;
    auto bail = [&](const char *what, cudaError_t er) {
        std::string msg = std::string(what) + ": " + cudaGetErrorString(er);
        if (c->d_counter) cudaFree(c->d_counter);
        if (c->last_launch) cudaEventDestroy(c->last_launch);
        if (c->ev1) cudaEventDestroy(c->ev1);
        if (c->ev0) cudaEventDestroy(c->ev0);
        if (c->stream) cudaStreamDestroy(c->stream);
        delete c;
        return fail(nullptr, TMPC_ERR_CUDA, msg);
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
    c->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
        delete c;
        return fail(nullptr, TMPC_ERR_UNSUPPORTED, "device is not sm_100-class; kernels are built for sm_100a only");
    }
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
    if ((e = cudaEventCreate(&c->ev0)) != cudaSuccess) return bail("event", e);
    if ((e = cudaEventCreate(&c->ev1)) != cudaSuccess) return bail("event", e);
    if ((e = cudaEventCreateWithFlags(&c->last_launch, cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
    if ((e = cudaMalloc(&c->d_counter, 5 * sizeof(unsigned long long))) != cudaSuccess) return bail("cudaMalloc", e);
    c->stats.parity_pinned = shape_parity_pinned(nx, nu) && order_policy == TMPC_ORDER_PARITY;
    *out = reinterpret_cast<tmpc_ctx *>(c);
    return TMPC_OK;
}

int tmpc_destroy(tmpc_ctx *ctx)
{
    if (!ctx) return TMPC_OK;
    tmpc_ctx_impl *c = CTX(ctx);
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto &st : c->stage) {
        if (st.s) { cudaStreamSynchronize(st.s); cudaStreamDestroy(st.s); }
        if (st.done) cudaEventDestroy(st.done);
        if (st.h_in) cudaFreeHost(st.h_in);
        if (st.h_out) cudaFreeHost(st.h_out);
        if (st.d_in) cudaFree(st.d_in);
        if (st.d_out) cudaFree(st.d_out);
    }
    if (c->g_in) cudaFree(c->g_in);
    if (c->g_out) cudaFree(c->g_out);
    if (c->g_done) cudaFree(c->g_done);
    if (c->g_copy) cudaStreamDestroy(c->g_copy);
    if (c->g_in_stream) cudaStreamDestroy(c->g_in_stream);
    if (c->g_h2d) cudaEventDestroy(c->g_h2d);
    if (c->g_first) cudaEventDestroy(c->g_first);
    if (c->g_gate) cudaFree(c->g_gate);
    if (c->d_counter) cudaFree(c->d_counter);
    if (c->d_model_w) cudaFree(c->d_model_w);
    if (c->d_model_rt) cudaFree(c->d_model_rt);
    if (c->d_rt_scratch) cudaFree(c->d_rt_scratch);
    if (c->d_pn_seed) cudaFree(c->d_pn_seed);
    if (c->d_lane_scratch) cudaFree(c->d_lane_scratch);
    if (c->d_model_f32) cudaFree(c->d_model_f32);
    for (void *p : c->d_ib) if (p) cudaFree(p);
    if (c->d_kinf) cudaFree(c->d_kinf);
    if (c->lpt_buf) cudaFree(c->lpt_buf);
    if (c->lpt_temp) cudaFree(c->lpt_temp);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->last_launch) cudaEventDestroy(c->last_launch);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return TMPC_OK;
}

int tmpc_set_model(tmpc_ctx *ctx, const void *Kinf, const void *Pinf, const void *Quu_inv, const void *AmBKt,
                   const void *Adyn, const void *Bdyn, const void *Q, double rho, const void *x_min,
                   const void *x_max, const void *u_min, const void *u_max)
{
    if (!ctx) return TMPC_ERR_INVALID;
    tmpc_ctx_impl *c = CTX(ctx);
    if (!Kinf || !Pinf || !Quu_inv || !AmBKt || !Adyn || !Bdyn || !Q)
        return fail(c, TMPC_ERR_INVALID, "model pointers must not be NULL");
    if ((x_min == nullptr) != (x_max == nullptr) || (u_min == nullptr) != (u_max == nullptr))
        return fail(c, TMPC_ERR_INVALID, "bounds must be given as min/max pairs");
    const size_t es = esize(c);
    const int nx = c->nx, nu = c->nu, N = c->N;
    auto cp = [&](std::vector<unsigned char> &dst, const void *src, size_t n) {
        dst.resize(n * es);
        if (src) std::memcpy(dst.data(), src, n * es);
    };
    cp(c->Kinf, Kinf, nu * nx); cp(c->Pinf, Pinf, nx * nx); cp(c->Quu_inv, Quu_inv, nu * nu);
    cp(c->AmBKt, AmBKt, nx * nx); cp(c->Adyn, Adyn, nx * nx); cp(c->Bdyn, Bdyn, nx * nu); cp(c->Q, Q, nx);
    c->has_xb = x_min != nullptr;
    c->has_ub = u_min != nullptr;
    cp(c->xmin, x_min, nx * N); cp(c->xmax, x_max, nx * N);
    cp(c->umin, u_min, nu * (N - 1)); cp(c->umax, u_max, nu * (N - 1));
    c->rho = rho;
    c->has_model = true;
    if (!build_model(c)) return fail(c, TMPC_ERR_UNSUPPORTED, "shape not compiled");
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaDeviceSynchronize());   // an earlier solve may still read the previous copy
    if (!c->d_kinf) CUDA_TRY(c, cudaMalloc(&c->d_kinf, (size_t)nu * nx * es));
    CUDA_TRY(c, cudaMemcpy(c->d_kinf, c->Kinf.data(), (size_t)nu * nx * es, cudaMemcpyHostToDevice));
    return TMPC_OK;
}

int tmpc_set_settings(tmpc_ctx *ctx, double abs_pri_tol, double abs_dua_tol, int max_iter, int check_termination,
                      int en_state_bound, int en_input_bound)
{
    if (!ctx) return TMPC_ERR_INVALID;
    tmpc_ctx_impl *c = CTX(ctx);
    if (check_termination < 1)
        return fail(c, TMPC_ERR_INVALID, "check_termination must be >= 1 (0 is undefined behaviour in the reference)");
    if (max_iter < 1) return fail(c, TMPC_ERR_INVALID, "max_iter must be >= 1");
    c->pri = abs_pri_tol; c->dua = abs_dua_tol; c->max_iter = max_iter; c->check_term = check_termination;
    c->en_state = en_state_bound; c->en_input = en_input_bound;
    if (c->has_model && !build_model(c)) return fail(c, TMPC_ERR_UNSUPPORTED, "shape not compiled");
    return TMPC_OK;
}

int tmpc_set_instance_bounds(tmpc_ctx *ctx, int64_t batch, const void *x_min, const void *x_max, const void *u_min,
                             const void *u_max, int32_t mem)
{
    if (!ctx) return TMPC_ERR_INVALID;
    tmpc_ctx_impl *c = CTX(ctx);
    if (batch < 0) return fail(c, TMPC_ERR_INVALID, "negative batch");
    if (mem != TMPC_MEM_HOST && mem != TMPC_MEM_DEVICE) return fail(c, TMPC_ERR_INVALID, "bad mem kind");
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaDeviceSynchronize());   // an earlier solve may still read the previous copies
    for (void *&p : c->d_ib) { if (p) cudaFree(p); p = nullptr; }
    c->ib_batch = 0;
    if (batch == 0) return TMPC_OK;
    if (!x_min || !x_max || !u_min || !u_max) return fail(c, TMPC_ERR_INVALID, "all four bound arrays must be given");
    if (!rt_shape_ok(c->nx, c->nu, c->N)) return fail(c, TMPC_ERR_UNSUPPORTED, "shape outside the run-time-shape kernel's range");
    const size_t es = esize(c);
    const size_t nb[4] = {(size_t)c->nx * c->N, (size_t)c->nx * c->N, (size_t)c->nu * (c->N - 1), (size_t)c->nu * (c->N - 1)};
    const void *src[4] = {x_min, x_max, u_min, u_max};
    for (int w = 0; w < 4; ++w) {
        const size_t bytes = (size_t)batch * nb[w] * es;
        CUDA_TRY(c, cudaMalloc(&c->d_ib[w], bytes));
        CUDA_TRY(c, cudaMemcpy(c->d_ib[w], src[w], bytes, mem == TMPC_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice));
    }
    c->ib_batch = batch;
    if (c->has_model && !build_model(c)) {
        for (void *&p : c->d_ib) { if (p) cudaFree(p); p = nullptr; }
        c->ib_batch = 0;
        return fail(c, TMPC_ERR_UNSUPPORTED, "shape not compiled");
    }
    return TMPC_OK;
}

int tmpc_solve(tmpc_ctx *ctx, const tmpc_solve_args *a)
{
    if (!ctx || !a) return TMPC_ERR_INVALID;
    tmpc_ctx_impl *c = CTX(ctx);
    if (!c->has_model) return fail(c, TMPC_ERR_STATE, "tmpc_set_model has not been called");
    if (a->batch < 0) return fail(c, TMPC_ERR_INVALID, "negative batch");
    if (a->batch > 0 && (!a->x0 || !a->Xref)) return fail(c, TMPC_ERR_INVALID, "x0 / Xref must not be NULL");
    if (a->warm && (!a->warm->d || !a->warm->y || !a->warm->g || !a->warm->v || !a->warm->z))
        return fail(c, TMPC_ERR_INVALID, "warm state needs all of d, y, g, v, z");
    CUDA_TRY(c, cudaSetDevice(c->device));
    c->stats.instances = a->batch;
    c->stats.iterations = c->stats.solved = c->stats.trips = 0;
    c->stats.launches = 0;
    c->stats.lanes = 0;
    c->stats.kernel_ms = 0.f;
    c->stats_pending = false;
    if (a->batch == 0) return TMPC_OK;

    const size_t es = esize(c);
    const int nx = c->nx, nu = c->nu, N = c->N;
    const size_t xrow = (size_t)nx * N, urow = (size_t)nu * (N - 1);
    const bool warm = a->warm != nullptr;

    if (a->mem == TMPC_MEM_DEVICE) {
        auto mis = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; };
        if (mis(a->x0) || mis(a->Xref) || mis(a->x) || mis(a->u) || mis(a->u0) ||
            (warm && (mis(a->warm->d) || mis(a->warm->y) || mis(a->warm->g) || mis(a->warm->v) || mis(a->warm->z))))
            return fail(c, TMPC_ERR_INVALID, "device buffers must be 16-byte aligned");
        DevArgs da{};
        da.batch = a->batch; da.x0 = a->x0; da.Xref = a->Xref;
        da.xref_stride = a->xref_shared ? 0 : (long long)xrow;
        if (warm) { da.wd = a->warm->d; da.wy = a->warm->y; da.wg = a->warm->g; da.wv = a->warm->v; da.wz = a->warm->z; }
        da.x = a->x; da.u = a->u; da.iter = a->iter; da.status = a->status; da.resid = a->resid; da.u0 = a->u0;
        cudaStream_t s = a->stream ? (cudaStream_t)a->stream : c->stream;
        int rc = launch_device(c, da, warm, s, true);
        if (rc != TMPC_OK) return rc;
        c->stats_pending = true;
        return TMPC_OK;
    }
    if (a->mem != TMPC_MEM_HOST) return fail(c, TMPC_ERR_INVALID, "bad mem kind");
    if (c->ib_batch && a->batch != c->ib_batch)
        return fail(c, TMPC_ERR_INVALID, "batch differs from the batch of tmpc_set_instance_bounds");
    if (!warm && (c->ib_batch || !getenv("TMPC_HOST_CHUNKED"))) return solve_host_gated(c, a);
    if (c->ib_batch)   // the chunked pipeline launches per chunk with chunk-relative instance numbers
        return fail(c, TMPC_ERR_UNSUPPORTED, "per-instance bounds with a warm start from host memory: use device buffers or tmpc_batch");
    // (the chunk loop below launches through plan_launch directly; it never sees per-instance bounds)

    // ---- host buffers: chunked 3-deep pipeline  H2D(k+1) | solve(k) | D2H(k-1) on three streams
    // input chunk image : x0 | [Xref per instance] | [warm d y z g v]
    // output chunk image: x | u | iter | status | resid | [warm d y z g v]
    const size_t in_per = nx * es + (a->xref_shared ? 0 : xrow * es) + (warm ? (3 * urow + 2 * xrow) * es : 0);
    const size_t out_per = (xrow + urow + nu) * es + 8 + 4 * es;
    int64_t chunk = std::min<int64_t>(a->batch, 131072);
    if (a->batch > chunk) chunk = (int64_t)((a->batch + ((a->batch + chunk - 1) / chunk) - 1) / ((a->batch + chunk - 1) / chunk));
    chunk = (chunk + 3) & ~int64_t(3);  // keep every sub-array 16-byte aligned
    const size_t xref_sh_bytes = xrow * es;
    void *d_xref_shared = nullptr;
    if (a->xref_shared) {
        CUDA_TRY(c, cudaMalloc(&d_xref_shared, xref_sh_bytes));
        CUDA_TRY(c, cudaMemcpyAsync(d_xref_shared, a->Xref, xref_sh_bytes, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
    const int nchunks = (int)((a->batch + chunk - 1) / chunk);
    const int depth = nchunks < 3 ? nchunks : 3;
    for (int k = 0; k < depth; ++k) {
        int rc = ensure_stage(c, k, in_per * chunk, out_per * chunk + (warm ? (3 * urow + 2 * xrow) * es * chunk : 0));
        if (rc != TMPC_OK) { if (d_xref_shared) cudaFree(d_xref_shared); return rc; }
    }
    // direct DMA from/to user memory when it is pinned, else bounce through the pinned staging buffers
    const bool direct = is_pinned(a->x0) && is_pinned(a->x) && is_pinned(a->u) && is_pinned(a->iter) &&
                        is_pinned(a->status) && is_pinned(a->resid) && (a->xref_shared || is_pinned(a->Xref)) &&
                        (!warm || (is_pinned(a->warm->d) && is_pinned(a->warm->y) && is_pinned(a->warm->g) &&
                                   is_pinned(a->warm->v) && is_pinned(a->warm->z)));
    float total_ms = 0.f;
    std::vector<cudaEvent_t> kev(2 * nchunks, nullptr);
    int rc_all = TMPC_OK;
    auto cleanup = [&]() {
        for (auto e : kev) if (e) cudaEventDestroy(e);
        if (d_xref_shared) cudaFree(d_xref_shared);
    };
    struct OutPtrs { char *x, *u, *iter, *status, *resid, *wd, *wy, *wz, *wg, *wv; };
    std::vector<int64_t> cb(nchunks), cn(nchunks);
    auto drain = [&](int k) -> int {  // copy chunk k's outputs from staging to user memory (bounce mode)
        auto &st = c->stage[k % 3];
        CUDA_TRY(c, cudaStreamSynchronize(st.s));
        if (!direct) {
            const int64_t b0 = cb[k], n = cn[k];
            char *h = (char *)st.h_out;
            size_t off = 0;
            auto take = [&](void *dst, size_t per) {
                if (dst) std::memcpy((char *)dst + b0 * per, h + off, n * per);
                off += chunk * per;
            };
            take(a->x, xrow * es); take(a->u, urow * es); take(a->iter, 4); take(a->status, 4); take(a->resid, 4 * es);
            take(a->u0, nu * es);
            if (warm) {
                take(a->warm->d, urow * es); take(a->warm->y, urow * es); take(a->warm->z, urow * es);
                take(a->warm->g, xrow * es); take(a->warm->v, xrow * es);
            }
        }
        return TMPC_OK;
    };
    for (int k = 0; k < nchunks && rc_all == TMPC_OK; ++k) {
        auto &st = c->stage[k % 3];
        if (k >= 3) { rc_all = drain(k - 3); if (rc_all != TMPC_OK) break; }
        const int64_t b0 = (int64_t)k * chunk, n = std::min<int64_t>(chunk, a->batch - b0);
        cb[k] = b0; cn[k] = n;
        // ---- H2D
        char *din = (char *)st.d_in;
        size_t off = 0;
        char *hin = (char *)st.h_in;
        auto put = [&](const void *src, size_t per) -> char * {
            char *d = din + off;
            if (direct) cudaMemcpyAsync(d, (const char *)src + b0 * per, n * per, cudaMemcpyHostToDevice, st.s);
            else std::memcpy(hin + off, (const char *)src + b0 * per, n * per);
            off += chunk * per;
            return d;
        };
        DevArgs da{};
        da.batch = n;
        da.x0 = put(a->x0, nx * es);
        if (a->xref_shared) { da.Xref = d_xref_shared; da.xref_stride = 0; }
        else { da.Xref = put(a->Xref, xrow * es); da.xref_stride = (long long)xrow; }
        char *dout = (char *)st.d_out;
        size_t oo = 0;
        auto outp = [&](size_t per) { char *p = dout + oo; oo += chunk * per; return p; };
        da.x = outp(xrow * es); da.u = outp(urow * es);
        da.iter = (int *)outp(4); da.status = (int *)outp(4); da.resid = outp(4 * es);
        da.u0 = outp(nu * es);
        if (!a->u0) da.u0 = nullptr;
        char *w_in[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        if (warm) {
            // warm state is in place on the device: stage it straight into the OUTPUT image, solve there
            da.wd = outp(urow * es); da.wy = outp(urow * es); da.wz = outp(urow * es);
            da.wg = outp(xrow * es); da.wv = outp(xrow * es);
            const void *src[5] = {a->warm->d, a->warm->y, a->warm->z, a->warm->g, a->warm->v};
            void *dst[5] = {da.wd, da.wy, da.wz, da.wg, da.wv};
            const size_t per[5] = {urow * es, urow * es, urow * es, xrow * es, xrow * es};
            for (int q = 0; q < 5; ++q) {
                if (direct) cudaMemcpyAsync(dst[q], (const char *)src[q] + b0 * per[q], n * per[q], cudaMemcpyHostToDevice, st.s);
                else {
                    w_in[q] = hin + off;
                    std::memcpy(hin + off, (const char *)src[q] + b0 * per[q], n * per[q]);
                    off += chunk * per[q];
                }
            }
        }
        if (!direct) {
            const size_t plain = chunk * (nx * es + (a->xref_shared ? 0 : xrow * es));
            cudaMemcpyAsync(din, hin, plain, cudaMemcpyHostToDevice, st.s);
            if (warm) {
                void *dst[5] = {da.wd, da.wy, da.wz, da.wg, da.wv};
                const size_t per[5] = {urow * es, urow * es, urow * es, xrow * es, xrow * es};
                for (int q = 0; q < 5; ++q) cudaMemcpyAsync(dst[q], w_in[q], n * per[q], cudaMemcpyHostToDevice, st.s);
            }
        }
        // ---- solve (kernels of different chunks serialise on the device through the shared work counter,
        //      so each chunk gets its own counter slot: reuse the ctx counter but order launches on st.s
        //      after the previous chunk's kernel)
        if (k > 0) cudaStreamWaitEvent(st.s, kev[2 * (k - 1) + 1], 0);
        else if ((rc_all = order_after_previous(c, st.s)) != TMPC_OK) break;
        cudaEventCreate(&kev[2 * k]);
        cudaEventCreate(&kev[2 * k + 1]);
        {
            KernelInfo ki;
            lookup_kernel(c->nx, c->nu, c->N, c->dtype, c->policy, warm, ki, c->pattern | (c->const_bounds ? 0x100 : 0));
            cudaFuncSetAttribute(ki.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ki.smem);
            // stats accumulate over chunks; only the work counter is reset per chunk
            cudaMemsetAsync(c->d_counter, 0, (k == 0 ? 5 : 1) * sizeof(unsigned long long), st.s);
            da.counter = c->d_counter;
            da.stats = c->d_counter + 1;
            long long blocks = 1;
            if ((rc_all = plan_launch(c, ki, da, st.s, blocks)) != TMPC_OK) break;
            if ((rc_all = plan_lane_scratch(c, ki, da, blocks, false, st.s)) != TMPC_OK) break;
            void *params[2] = {model_param(c, ki), &da};
            cudaEventRecord(kev[2 * k], st.s);
            cudaError_t e = cudaLaunchKernel(ki.fn, dim3((unsigned)blocks), dim3(ki.block), params, ki.smem, st.s);
            cudaEventRecord(kev[2 * k + 1], st.s);
            if (e != cudaSuccess) { rc_all = fail(c, TMPC_ERR_CUDA, std::string("launch: ") + cudaGetErrorString(e)); break; }
            if ((rc_all = mark_launch(c, st.s)) != TMPC_OK) break;
            c->stats.launches += 1;
            c->stats.lanes = (int32_t)(blocks * ki.per_block);
            c->stats.pattern = c->pattern;
        }
        // ---- D2H
        {
            size_t o2 = 0;
            char *hout = (char *)st.h_out;
            auto get = [&](void *dst, size_t per) {
                if (direct) { if (dst) cudaMemcpyAsync((char *)dst + b0 * per, dout + o2, n * per, cudaMemcpyDeviceToHost, st.s); }
                else if (dst) cudaMemcpyAsync(hout + o2, dout + o2, n * per, cudaMemcpyDeviceToHost, st.s);
                o2 += chunk * per;
            };
            get(a->x, xrow * es); get(a->u, urow * es); get(a->iter, 4); get(a->status, 4); get(a->resid, 4 * es);
            get(a->u0, nu * es);
            if (warm) {
                get(a->warm->d, urow * es); get(a->warm->y, urow * es); get(a->warm->z, urow * es);
                get(a->warm->g, xrow * es); get(a->warm->v, xrow * es);
            }
        }
    }
    if (rc_all == TMPC_OK)
        for (int k = std::max(0, nchunks - 3); k < nchunks && rc_all == TMPC_OK; ++k) rc_all = drain(k);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc_all == TMPC_OK && e != cudaSuccess) rc_all = fail(c, TMPC_ERR_CUDA, std::string("solve: ") + cudaGetErrorString(e));
    if (rc_all == TMPC_OK) {
        for (int k = 0; k < nchunks; ++k) {
            float ms = 0.f;
            if (kev[2 * k] && cudaEventElapsedTime(&ms, kev[2 * k], kev[2 * k + 1]) == cudaSuccess) total_ms += ms;
        }
        c->stats.kernel_ms = total_ms;
        unsigned long long h[5];
        if (cudaMemcpy(h, c->d_counter, sizeof h, cudaMemcpyDeviceToHost) == cudaSuccess) {
            c->stats.iterations = (int64_t)h[1]; c->stats.solved = (int64_t)h[2]; c->stats.trips = (int64_t)h[3];
        }
    }
    cleanup();
    return rc_all;
}

int tmpc_step(tmpc_ctx *ctx, int which, int64_t batch, const tmpc_workspace *ws, int32_t iter, int32_t mem, void *stream)
{
    if (!ctx || !ws) return TMPC_ERR_INVALID;
    tmpc_ctx_impl *c = CTX(ctx);
    if (!c->has_model) return fail(c, TMPC_ERR_STATE, "tmpc_set_model has not been called");
    if (which < 0 || which > 5) return fail(c, TMPC_ERR_INVALID, "bad step index");
    if (batch <= 0) return batch == 0 ? TMPC_OK : fail(c, TMPC_ERR_INVALID, "negative batch");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const size_t es = esize(c);
    const size_t xrow = (size_t)c->nx * c->N, urow = (size_t)c->nu * (c->N - 1);
    void *arr[12] = {ws->x, ws->u, ws->q, ws->r, ws->p, ws->d, ws->v, ws->vnew, ws->z, ws->znew, ws->g, ws->y};
    const size_t per[12] = {xrow, urow, xrow, urow, xrow, urow, xrow, xrow, urow, urow, xrow, urow};
    for (void *p : arr) if (!p) return fail(c, TMPC_ERR_INVALID, "workspace arrays must not be NULL");
    if (!ws->Xref) return fail(c, TMPC_ERR_INVALID, "Xref must not be NULL");
    cudaStream_t s = (mem == TMPC_MEM_DEVICE && stream) ? (cudaStream_t)stream : c->stream;
    void *dev[12];
    void *d_xref = nullptr, *d_res = nullptr;
    int *d_term = nullptr;
    const size_t xref_n = ws->xref_shared ? xrow : xrow * batch;
    if (mem == TMPC_MEM_HOST) {
        for (int k = 0; k < 12; ++k) {
            CUDA_TRY(c, cudaMalloc(&dev[k], per[k] * batch * es));
            CUDA_TRY(c, cudaMemcpyAsync(dev[k], arr[k], per[k] * batch * es, cudaMemcpyHostToDevice, s));
        }
        CUDA_TRY(c, cudaMalloc(&d_xref, xref_n * es));
        CUDA_TRY(c, cudaMemcpyAsync(d_xref, ws->Xref, xref_n * es, cudaMemcpyHostToDevice, s));
        CUDA_TRY(c, cudaMalloc(&d_res, 4 * batch * es));
        if (ws->resid) CUDA_TRY(c, cudaMemcpyAsync(d_res, ws->resid, 4 * batch * es, cudaMemcpyHostToDevice, s));
        else CUDA_TRY(c, cudaMemsetAsync(d_res, 0, 4 * batch * es, s));
        CUDA_TRY(c, cudaMalloc((void **)&d_term, sizeof(int) * batch));
    } else {
        for (int k = 0; k < 12; ++k) dev[k] = arr[k];
        d_xref = const_cast<void *>(ws->Xref);
        d_res = ws->resid;
        d_term = ws->term;
        if (which == 4 && !d_res) return fail(c, TMPC_ERR_INVALID, "step 4 needs resid");
    }
    cudaError_t e = cudaErrorInvalidValue;
    auto go = [&](auto tag, auto nx_, auto nu_, auto nh_) {
        using T = decltype(tag);
        tmpc::StepArgs<T> sa;
        sa.batch = batch;
        T **dst[12] = {&sa.x, &sa.u, &sa.q, &sa.r, &sa.p, &sa.d, &sa.v, &sa.vnew, &sa.z, &sa.znew, &sa.g, &sa.y};
        for (int k = 0; k < 12; ++k) *dst[k] = (T *)dev[k];
        sa.Xref = (const T *)d_xref;
        sa.xref_stride = ws->xref_shared ? 0 : (long long)xrow;
        sa.resid = (T *)d_res;
        sa.term = d_term;
        sa.iter = iter;
        e = launch_step<T, decltype(nx_)::value, decltype(nu_)::value, decltype(nh_)::value>(c, sa, which, s);
    };
    using I = std::integral_constant<int, 0>;
    (void)sizeof(I);
    const bool f32 = c->dtype == TMPC_F32;
    if (c->nx == 12 && c->nu == 4 && c->N == 10) {
        if (f32) go(float(), std::integral_constant<int, 12>(), std::integral_constant<int, 4>(), std::integral_constant<int, 10>());
        else go(double(), std::integral_constant<int, 12>(), std::integral_constant<int, 4>(), std::integral_constant<int, 10>());
    } else if (c->nx == 4 && c->nu == 1 && c->N == 10) {
        if (f32) go(float(), std::integral_constant<int, 4>(), std::integral_constant<int, 1>(), std::integral_constant<int, 10>());
        else go(double(), std::integral_constant<int, 4>(), std::integral_constant<int, 1>(), std::integral_constant<int, 10>());
    } else if (c->nx == 32 && c->nu == 8 && c->N == 50 && f32) {
        go(float(), std::integral_constant<int, 32>(), std::integral_constant<int, 8>(), std::integral_constant<int, 50>());
    } else if (c->rt_ready) {
        // any other shape: the run-time-shape step kernel (tmpc_kernel_rt.cuh)
        auto go_rt = [&](auto tag) {
            using T = decltype(tag);
            tmpc::StepArgs<T> sa;
            sa.batch = batch;
            T **dst[12] = {&sa.x, &sa.u, &sa.q, &sa.r, &sa.p, &sa.d, &sa.v, &sa.vnew, &sa.z, &sa.znew, &sa.g, &sa.y};
            for (int k = 0; k < 12; ++k) *dst[k] = (T *)dev[k];
            sa.Xref = (const T *)d_xref;
            sa.xref_stride = ws->xref_shared ? 0 : (long long)xrow;
            sa.resid = (T *)d_res;
            sa.term = d_term;
            sa.iter = iter;
            const tmpc::ModelRT<T> &m = *reinterpret_cast<const tmpc::ModelRT<T> *>(c->model_rt.data());
            const size_t smem = tmpc::rt_smem_bytes(c->nx, c->nu, sizeof(T));
            const unsigned blocks = (unsigned)((batch + tmpc::RT_BLOCK - 1) / tmpc::RT_BLOCK);
            if (c->policy == TMPC_ORDER_PARITY) {
                cudaFuncSetAttribute(tmpc::step_kernel_rt<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                tmpc::step_kernel_rt<T, false><<<blocks, tmpc::RT_BLOCK, smem, s>>>(m, sa, which);
            } else {
                cudaFuncSetAttribute(tmpc::step_kernel_rt<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                tmpc::step_kernel_rt<T, true><<<blocks, tmpc::RT_BLOCK, smem, s>>>(m, sa, which);
            }
            e = cudaGetLastError();
        };
        if (f32) go_rt(float()); else go_rt(double());
    } else {
        return fail(c, TMPC_ERR_UNSUPPORTED, "no step kernels for this shape");
    }
    if (e != cudaSuccess) return fail(c, TMPC_ERR_CUDA, std::string("step launch: ") + cudaGetErrorString(e));
    if (mem == TMPC_MEM_HOST) {
        for (int k = 0; k < 12; ++k) CUDA_TRY(c, cudaMemcpyAsync(arr[k], dev[k], per[k] * batch * es, cudaMemcpyDeviceToHost, s));
        if (ws->resid) CUDA_TRY(c, cudaMemcpyAsync(ws->resid, d_res, 4 * batch * es, cudaMemcpyDeviceToHost, s));
        if (ws->term) CUDA_TRY(c, cudaMemcpyAsync(ws->term, d_term, sizeof(int) * batch, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(c, cudaStreamSynchronize(s));
        for (int k = 0; k < 12; ++k) cudaFree(dev[k]);
        cudaFree(d_xref); cudaFree(d_res); cudaFree(d_term);
    }
    return TMPC_OK;
}

int tmpc_get_stats(tmpc_ctx *ctx, tmpc_stats *out)
{
    if (!ctx || !out) return TMPC_ERR_INVALID;
    tmpc_ctx_impl *c = CTX(ctx);
    if (c->stats_pending) {
        CUDA_TRY(c, cudaSetDevice(c->device));
        CUDA_TRY(c, cudaEventSynchronize(c->ev1));
        float ms = 0.f;
        CUDA_TRY(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        c->stats.kernel_ms = ms;
        unsigned long long h[5];
        CUDA_TRY(c, cudaMemcpy(h, c->d_counter, sizeof h, cudaMemcpyDeviceToHost));
        c->stats.iterations = (int64_t)h[1]; c->stats.solved = (int64_t)h[2]; c->stats.trips = (int64_t)h[3];
        c->stats_pending = false;
    }
    c->stats.scheduled = c->lpt_used;
    *out = c->stats;
    return TMPC_OK;
}

int tmpc_host_alloc(void **ptr, uint64_t bytes)
{
    if (!ptr) return TMPC_ERR_INVALID;
    return cudaHostAlloc(ptr, bytes, cudaHostAllocPortable) == cudaSuccess ? TMPC_OK : TMPC_ERR_CUDA;
}
int tmpc_host_free(void *ptr) { return cudaFreeHost(ptr) == cudaSuccess ? TMPC_OK : TMPC_ERR_CUDA; }

}  // extern "C"

#include "tmpc_batch.cuh"
#include "tmpc_systems.cuh"
#include "tmpc_multi.cuh"
