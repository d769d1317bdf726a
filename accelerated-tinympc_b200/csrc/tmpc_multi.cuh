// tmpc_multi: one host batch over several devices from ONE process (include/tmpc.h).  Included at the end of tmpc_api.cu.
//
// The reference has nothing to shard (one TinySolver, one thread: quadrotor_hovering.cpp:104); its callers are plain C++
// loops, so the multi-GPU form has to live behind the same C entry points rather than behind a launcher.  Instances never
// interact (admm.cpp:111-152 touches one workspace), so the batch splits into contiguous index ranges [B r / G, B (r+1) / G)
// -- the same rule as sharding.py for the one-process-per-GPU bench -- with no inter-device traffic at all.  One persistent
// host worker thread per device drives that device's ctx (a ctx is single-caller); tmpc_multi_solve hands each worker its
// range of the caller's arrays and waits for all of them.
#pragma once
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>

namespace {

struct MultiWorker {
    tmpc_ctx *ctx = nullptr;
    int device = 0;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;   // pending job (empty = none)
    bool has_job = false, busy = false, quit = false;
    int rc = TMPC_OK;
    tmpc_stats stats{};
    int64_t b0 = 0, n = 0;      // range of the last solve
};

struct tmpc_multi_impl {
    int nx = 0, nu = 0, N = 0, dtype = 0, policy = 0;
    std::vector<std::unique_ptr<MultiWorker>> w;
    int used = 0;               // devices that took part in the last solve
    int64_t ib_batch = 0;
    std::string err;
};

#define MULTI(m) reinterpret_cast<tmpc_multi_impl *>(m)

void worker_main(MultiWorker *w)
{
    cudaSetDevice(w->device);
    std::unique_lock<std::mutex> lk(w->mu);
    for (;;) {
        w->cv.wait(lk, [&] { return w->has_job || w->quit; });
        if (w->quit) return;
        std::function<int()> job = std::move(w->job);
        w->has_job = false;
        lk.unlock();
        const int rc = job();
        lk.lock();
        w->rc = rc;
        w->busy = false;
        w->cv.notify_all();
    }
}

void submit(MultiWorker *w, std::function<int()> job)
{
    std::lock_guard<std::mutex> lk(w->mu);
    w->job = std::move(job);
    w->has_job = true;
    w->busy = true;
    w->cv.notify_all();
}

int wait_done(MultiWorker *w)
{
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv.wait(lk, [&] { return !w->busy; });
    return w->rc;
}

// run `f(worker index)` on the first `n` workers concurrently; first failure wins
int run_all(tmpc_multi_impl *m, int n, const std::function<int(int)> &f)
{
    for (int i = 0; i < n; ++i) submit(m->w[i].get(), [=] { return f(i); });
    int rc_all = TMPC_OK;
    for (int i = 0; i < n; ++i) {
        const int rc = wait_done(m->w[i].get());
        if (rc != TMPC_OK && rc_all == TMPC_OK) {
            rc_all = rc;
            m->err = "device " + std::to_string(m->w[i]->device) + ": " + tmpc_last_error(m->w[i]->ctx);
        }
    }
    return rc_all;
}

int mfail(tmpc_multi_impl *m, int code, const std::string &msg)
{
    if (m) m->err = msg;
    else g_create_error = msg;
    return code;
}

}  // namespace

extern "C" {

int tmpc_multi_create(tmpc_multi **out, int ndev, const int *devices, int nx, int nu, int N, int dtype, int order_policy)
{
    if (!out) return mfail(nullptr, TMPC_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int visible = 0;
    cudaError_t e = cudaGetDeviceCount(&visible);
    if (e != cudaSuccess || visible == 0)
        return mfail(nullptr, TMPC_ERR_CUDA, std::string("no CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e));
    if (ndev < 0 || ndev > visible) return mfail(nullptr, TMPC_ERR_INVALID, "ndev exceeds the visible devices");
    if (ndev == 0) { ndev = visible; devices = nullptr; }
    auto *m = new tmpc_multi_impl;
    m->nx = nx; m->nu = nu; m->N = N; m->dtype = dtype; m->policy = order_policy;
    for (int i = 0; i < ndev; ++i) {
        const int dev = devices ? devices[i] : i;
        tmpc_ctx *ctx = nullptr;
        const int rc = tmpc_create(&ctx, dev, nx, nu, N, dtype, order_policy);
        if (rc != TMPC_OK) {
            const std::string why = g_create_error;
            for (auto &w : m->w) tmpc_destroy(w->ctx);
            delete m;
            return mfail(nullptr, rc, "device " + std::to_string(dev) + ": " + why);
        }
        m->w.emplace_back(new MultiWorker);
        m->w.back()->ctx = ctx;
        m->w.back()->device = dev;
    }
    for (auto &w : m->w) w->th = std::thread(worker_main, w.get());
    *out = reinterpret_cast<tmpc_multi *>(m);
    return TMPC_OK;
}

int tmpc_multi_destroy(tmpc_multi *mm)
{
    if (!mm) return TMPC_OK;
    tmpc_multi_impl *m = MULTI(mm);
    for (auto &w : m->w) {
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->quit = true;
            w->cv.notify_all();
        }
        if (w->th.joinable()) w->th.join();
        tmpc_destroy(w->ctx);
    }
    delete m;
    return TMPC_OK;
}

int tmpc_multi_device_count(const tmpc_multi *mm) { return mm ? (int)reinterpret_cast<const tmpc_multi_impl *>(mm)->w.size() : 0; }

tmpc_ctx *tmpc_multi_ctx(tmpc_multi *mm, int index)
{
    if (!mm) return nullptr;
    tmpc_multi_impl *m = MULTI(mm);
    return (index >= 0 && index < (int)m->w.size()) ? m->w[index]->ctx : nullptr;
}

const char *tmpc_multi_last_error(const tmpc_multi *mm)
{
    return mm ? reinterpret_cast<const tmpc_multi_impl *>(mm)->err.c_str() : g_create_error.c_str();
}

int tmpc_multi_set_model(tmpc_multi *mm, const void *Kinf, const void *Pinf, const void *Quu_inv, const void *AmBKt, const void *Adyn,
                         const void *Bdyn, const void *Q, double rho, const void *x_min, const void *x_max, const void *u_min,
                         const void *u_max)
{
    if (!mm) return TMPC_ERR_INVALID;
    tmpc_multi_impl *m = MULTI(mm);
    return run_all(m, (int)m->w.size(), [=](int i) {
        return tmpc_set_model(m->w[i]->ctx, Kinf, Pinf, Quu_inv, AmBKt, Adyn, Bdyn, Q, rho, x_min, x_max, u_min, u_max);
    });
}

int tmpc_multi_set_settings(tmpc_multi *mm, double abs_pri_tol, double abs_dua_tol, int max_iter, int check_termination,
                            int en_state_bound, int en_input_bound)
{
    if (!mm) return TMPC_ERR_INVALID;
    tmpc_multi_impl *m = MULTI(mm);
    return run_all(m, (int)m->w.size(), [=](int i) {
        return tmpc_set_settings(m->w[i]->ctx, abs_pri_tol, abs_dua_tol, max_iter, check_termination, en_state_bound, en_input_bound);
    });
}

namespace {
// devices a batch of B instances is spread over: at least 16,384 instances per device (below that a second device costs
// more in launch + copy set-up than it saves), and with per-instance bounds exactly the split they were installed with
int devices_for(const tmpc_multi_impl *m, int64_t B)
{
    const int64_t by_size = std::max<int64_t>(1, B / 16384);
    return (int)std::min<int64_t>((int64_t)m->w.size(), by_size);
}
}  // namespace

int tmpc_multi_set_instance_bounds(tmpc_multi *mm, int64_t batch, const void *x_min, const void *x_max, const void *u_min,
                                   const void *u_max)
{
    if (!mm) return TMPC_ERR_INVALID;
    tmpc_multi_impl *m = MULTI(mm);
    if (batch < 0) return mfail(m, TMPC_ERR_INVALID, "negative batch");
    if (batch > 0 && (!x_min || !x_max || !u_min || !u_max)) return mfail(m, TMPC_ERR_INVALID, "all four bound arrays must be given");
    const size_t es = m->dtype == TMPC_F32 ? 4 : 8;
    const size_t xb = (size_t)m->nx * m->N * es, ub = (size_t)m->nu * (m->N - 1) * es;
    const int G = batch > 0 ? devices_for(m, batch) : (int)m->w.size();
    const int rc = run_all(m, (int)m->w.size(), [=](int i) {
        const int64_t b0 = i < G ? batch * i / G : 0, b1 = i < G ? batch * (i + 1) / G : 0;
        if (b1 == b0) return tmpc_set_instance_bounds(m->w[i]->ctx, 0, nullptr, nullptr, nullptr, nullptr, TMPC_MEM_HOST);
        return tmpc_set_instance_bounds(m->w[i]->ctx, b1 - b0, (const char *)x_min + b0 * xb, (const char *)x_max + b0 * xb,
                                        (const char *)u_min + b0 * ub, (const char *)u_max + b0 * ub, TMPC_MEM_HOST);
    });
    m->ib_batch = rc == TMPC_OK ? batch : 0;
    return rc;
}

int tmpc_multi_solve(tmpc_multi *mm, const tmpc_solve_args *a)
{
    if (!mm || !a) return TMPC_ERR_INVALID;
    tmpc_multi_impl *m = MULTI(mm);
    if (a->mem != TMPC_MEM_HOST) return mfail(m, TMPC_ERR_INVALID, "tmpc_multi_solve takes host arrays (device arrays belong to one device: use tmpc_multi_ctx + tmpc_solve)");
    if (a->batch < 0) return mfail(m, TMPC_ERR_INVALID, "negative batch");
    if (m->ib_batch && a->batch != m->ib_batch) return mfail(m, TMPC_ERR_INVALID, "batch differs from the batch of tmpc_multi_set_instance_bounds");
    const int64_t B = a->batch;
    const int G = devices_for(m, B);
    m->used = G;
    const size_t es = m->dtype == TMPC_F32 ? 4 : 8;
    const size_t xrow = (size_t)m->nx * m->N * es, urow = (size_t)m->nu * (m->N - 1) * es;
    const size_t x0b = (size_t)m->nx * es, u0b = (size_t)m->nu * es;
    const tmpc_solve_args src = *a;
    const tmpc_warm wsrc = a->warm ? *a->warm : tmpc_warm{};
    const bool has_warm = a->warm != nullptr;
    return run_all(m, G, [=](int i) {
        MultiWorker *w = m->w[i].get();
        const int64_t b0 = B * i / G, b1 = B * (i + 1) / G;
        w->b0 = b0; w->n = b1 - b0;
        auto at = [&](const void *p, size_t per) -> void * { return p ? (void *)((const char *)p + (size_t)b0 * per) : nullptr; };
        tmpc_solve_args s = src;
        tmpc_warm wm;
        s.batch = b1 - b0;
        s.x0 = at(src.x0, x0b);
        s.Xref = src.xref_shared ? src.Xref : at(src.Xref, xrow);
        if (has_warm) {
            wm.d = at(wsrc.d, urow); wm.y = at(wsrc.y, urow); wm.z = at(wsrc.z, urow); wm.g = at(wsrc.g, xrow); wm.v = at(wsrc.v, xrow);
            s.warm = &wm;
        }
        s.x = at(src.x, xrow); s.u = at(src.u, urow); s.u0 = at(src.u0, u0b);
        s.iter = (int32_t *)at(src.iter, 4); s.status = (int32_t *)at(src.status, 4); s.resid = at(src.resid, 4 * es);
        s.stream = nullptr;
        int rc = tmpc_solve(w->ctx, &s);
        if (rc == TMPC_OK) rc = tmpc_get_stats(w->ctx, &w->stats);
        return rc;
    });
}

int tmpc_multi_rollout(tmpc_multi *mm, const tmpc_rollout_args *a)
{
    if (!mm || !a) return TMPC_ERR_INVALID;
    tmpc_multi_impl *m = MULTI(mm);
    if (a->batch < 1 || a->steps < 1 || !a->x0 || (!a->table && !a->Xref)) return mfail(m, TMPC_ERR_INVALID, "tmpc_multi_rollout: batch, steps >= 1, x0 and a reference are required");
    if (m->ib_batch) return mfail(m, TMPC_ERR_UNSUPPORTED, "tmpc_multi_rollout with per-instance bounds");
    const int64_t B = a->batch;
    const int G = devices_for(m, B);
    m->used = G;
    const size_t es = m->dtype == TMPC_F32 ? 4 : 8;
    const size_t xrow = (size_t)m->nx * m->N * es, urow = (size_t)m->nu * (m->N - 1) * es, x0b = (size_t)m->nx * es, u0b = (size_t)m->nu * es;
    const tmpc_rollout_args r = *a;
    return run_all(m, G, [=](int i) {
        MultiWorker *w = m->w[i].get();
        const int64_t b0 = B * i / G, b1 = B * (i + 1) / G;
        w->b0 = b0; w->n = b1 - b0;
        auto at = [&](const void *p, size_t per) -> void * { return p ? (void *)((const char *)p + (size_t)b0 * per) : nullptr; };
        tmpc_batch *bt = nullptr;
        int rc = tmpc_batch_create(w->ctx, b1 - b0, &bt);
        if (rc != TMPC_OK) return rc;
        rc = tmpc_batch_set_x0(bt, at(r.x0, x0b), TMPC_MEM_HOST);
        if (rc == TMPC_OK) {
            if (r.table) rc = tmpc_batch_set_xref_table(bt, r.table, r.table_rows, r.start ? r.start + b0 : nullptr, TMPC_MEM_HOST);
            else rc = tmpc_batch_set_xref(bt, r.xref_shared ? r.Xref : at(r.Xref, xrow), r.xref_shared ? 1 : 0, TMPC_MEM_HOST);
        }
        // the histories are [step][batch][...]: this device's columns start at instance b0 of every step row
        if (rc == TMPC_OK)
            rc = batch_rollout(bt, r.steps, r.reset_duals, at(r.x0_hist, x0b), at(r.u0_hist, u0b), (int32_t *)at(r.iter_hist, 4),
                               (int32_t *)at(r.status_hist, 4), TMPC_MEM_HOST, B);
        if (rc == TMPC_OK && r.x) rc = tmpc_batch_get(bt, TMPC_GET_X, at(r.x, xrow), TMPC_MEM_HOST);
        if (rc == TMPC_OK && r.u) rc = tmpc_batch_get(bt, TMPC_GET_U, at(r.u, urow), TMPC_MEM_HOST);
        if (rc == TMPC_OK) rc = tmpc_get_stats(w->ctx, &w->stats);
        tmpc_batch_destroy(bt);   // (keeps the ctx's last error text when something failed above)
        return rc;
    });
}

int tmpc_multi_get_stats(tmpc_multi *mm, tmpc_stats *total, tmpc_stats *per_device)
{
    if (!mm || !total) return TMPC_ERR_INVALID;
    tmpc_multi_impl *m = MULTI(mm);
    tmpc_stats t{};
    for (int i = 0; i < (int)m->w.size(); ++i) {
        tmpc_stats s{};
        if (i < m->used) s = m->w[i]->stats;
        if (per_device) per_device[i] = s;
        if (i >= m->used) continue;
        t.instances += s.instances; t.iterations += s.iterations; t.solved += s.solved; t.trips += s.trips;
        t.launches += s.launches; t.lanes += s.lanes;
        t.kernel_ms = std::max(t.kernel_ms, s.kernel_ms);
        t.parity_pinned = s.parity_pinned; t.pattern = s.pattern; t.scheduled = s.scheduled;
    }
    *total = t;
    return TMPC_OK;
}

}  // extern "C"
