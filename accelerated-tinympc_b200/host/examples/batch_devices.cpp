// tiny_solve_batch over every visible GPU from one process, and the controls-only output mask.
// The reference's callers are plain C++ loops around tiny_solve (quadrotor_hovering.cpp:104) that consume u(:,0)
// (:110); this is the batched form of that call: one host batch of quadrotor instances, solved
//   (1) on one device, full outputs                      (tiny_set_devices(s, 1))
//   (2) on every visible device, full outputs            (tiny_set_devices(s, 0): contiguous instance ranges, one per device)
//   (3) on every visible device, controls only           (out.u0 + iter + status: 24 instead of 648 bytes per solve back)
// Self-checking: (2) must equal (1) bit for bit, instance by instance; (3) must equal u(:,0), iter, status of (1).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mpcdata.hpp"
#include "tinympc/tiny_api.hpp"
#include "tmpc.h"

static void load_cache(TinySolver *s, const MpcData &d)
{
    const auto K = d.cast<tinytype>("Kinf"), P = d.cast<tinytype>("Pinf"), Qi = d.cast<tinytype>("Quu_inv"), M = d.cast<tinytype>("AmBKt");
    std::memcpy(s->cache->Kinf.data(), K.data(), sizeof(tinytype) * K.size());
    std::memcpy(s->cache->Pinf.data(), P.data(), sizeof(tinytype) * P.size());
    std::memcpy(s->cache->Quu_inv.data(), Qi.data(), sizeof(tinytype) * Qi.size());
    std::memcpy(s->cache->AmBKt.data(), M.data(), sizeof(tinytype) * M.size());
}

// SURVEY 8d generator: u01(seed, idx) = (splitmix64(seed ^ splitmix64(idx)) >> 40) * 2^-24
static uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static double u01(uint64_t seed, uint64_t idx) { return (double)(splitmix64(seed ^ splitmix64(idx)) >> 40) * (1.0 / 16777216.0); }

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s <problem_data dir> [instances]\n", argv[0]); return 2; }
    const std::string dir = argv[1];
    const int64_t Bn = argc > 2 ? atoll(argv[2]) : 200000;
    const int nx = 12, nu = 4, N = 10;
    MpcData d(dir + "/quadrotor_20hz.mpcdata");
    const auto A = d.cast<tinytype>("Adyn"), B = d.cast<tinytype>("Bdyn"), Q = d.cast<tinytype>("Q"), R = d.cast<tinytype>("R");
    std::vector<tinytype> xlo(nx * N, -5), xhi(nx * N, 5), ulo(nu * (N - 1), (tinytype)-0.5), uhi(nu * (N - 1), (tinytype)0.5);
    TinySolver *s = nullptr;
    if (tiny_setup(&s, nx, nu, N, A.data(), B.data(), Q.data(), R.data(), (tinytype)d.scalars.at("rho"), xlo.data(), xhi.data(),
                   ulo.data(), uhi.data(), 0) != 0) { fprintf(stderr, "%s\n", tiny_last_error()); return 1; }
    load_cache(s, d);
    const double scale[12] = {2, 2, 2, .2, .2, .2, .5, .5, .5, .5, .5, .5}, hover[12] = {0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<tinytype> x0((size_t)Bn * nx), xref((size_t)N * nx, 0);
    for (int i = 0; i < N; ++i) xref[(size_t)i * nx + 2] = 2;
    for (int64_t b = 0; b < Bn; ++b)
        for (int j = 0; j < nx; ++j) x0[(size_t)b * nx + j] = (tinytype)(hover[j] + 0.25 * scale[j] * (2.0 * u01(1234, (uint64_t)(12 * b + j)) - 1.0));

    struct Out { std::vector<tinytype> x, u, u0; std::vector<int32_t> it, st; };
    auto run = [&](int devices, bool controls_only, Out &o, double &ms) -> int {
        if (tiny_set_devices(s, devices) != 0) { fprintf(stderr, "%s\n", tiny_last_error()); return 1; }
        o.it.assign(Bn, 0); o.st.assign(Bn, 0);
        TinyBatchIn in;
        std::memset(&in, 0, sizeof in);
        in.batch = Bn; in.x0 = x0.data(); in.Xref = xref.data(); in.xref_shared = 1;
        TinyBatchOut out;
        std::memset(&out, 0, sizeof out);
        out.iter = o.it.data(); out.status = o.st.data();
        if (controls_only) { o.u0.assign((size_t)Bn * nu, 0); out.u0 = o.u0.data(); }
        else { o.x.assign((size_t)Bn * N * nx, 0); o.u.assign((size_t)Bn * (N - 1) * nu, 0); out.x = o.x.data(); out.u = o.u.data(); }
        for (int rep = 0; rep < 2; ++rep) {   // first call creates the contexts
            const auto t0 = std::chrono::steady_clock::now();
            if (tiny_solve_batch(s, &in, &out) != 0) { fprintf(stderr, "tiny_solve_batch: %s\n", tiny_last_error()); return 1; }
            ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
        return 0;
    };
    const int ndev = tmpc_device_count();
    Out one, all, ctl;
    double ms1 = 0, msa = 0, msc = 0;
    if (run(1, false, one, ms1) || run(0, false, all, msa) || run(0, true, ctl, msc)) return 1;
    long long iters = 0;
    for (int64_t b = 0; b < Bn; ++b) iters += one.it[b];
    if (one.it != all.it || one.st != all.st || one.x != all.x || one.u != all.u) { printf("FAIL: %d devices differ from one device\n", ndev); return 2; }
    if (one.it != ctl.it || one.st != ctl.st) { printf("FAIL: controls-only iteration counts differ\n"); return 2; }
    for (int64_t b = 0; b < Bn; ++b)
        if (std::memcmp(&ctl.u0[(size_t)b * nu], &one.u[(size_t)b * (N - 1) * nu], sizeof(tinytype) * nu) != 0) {
            printf("FAIL: u0 of instance %lld differs from u(:,0)\n", (long long)b);
            return 2;
        }
    printf("batch_devices ok: %lld instances, %lld iterations, %d device(s); pageable host arrays: one device %.1f ms, all devices %.1f ms, "
           "all devices controls-only %.1f ms\n", (long long)Bn, iters, ndev, ms1, msa, msc);
    tiny_free(s);
    return 0;
}
