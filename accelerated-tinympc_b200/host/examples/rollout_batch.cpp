// tiny_rollout_batch: the examples' closed loop (quadrotor_hovering.cpp:90-114) for a batch of quadrotors in ONE call, against the
// same loop written the way the reference's examples write it -- per MPC step: reset duals, tiny_solve_batch with the warm state
// carried in host arrays.  Self-checking: every applied control, iteration count and status of the one-call loop, and the last
// step's trajectories, must equal the step-by-step loop bit for bit (the step-by-step loop takes its measurements from the
// one-call loop's plant states).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mpcdata.hpp"
#include "tinympc/tiny_api.hpp"

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
    if (argc < 2) { fprintf(stderr, "usage: %s <problem_data dir> [instances] [steps]\n", argv[0]); return 2; }
    const std::string dir = argv[1];
    const int64_t Bn = argc > 2 ? atoll(argv[2]) : 50000;
    const int steps = argc > 3 ? atoi(argv[3]) : 6;
    const int nx = 12, nu = 4, N = 10;
    MpcData d(dir + "/quadrotor_20hz.mpcdata");
    const auto A = d.cast<tinytype>("Adyn"), B = d.cast<tinytype>("Bdyn"), Q = d.cast<tinytype>("Q"), R = d.cast<tinytype>("R");
    std::vector<tinytype> xlo(nx * N, -5), xhi(nx * N, 5), ulo(nu * (N - 1), (tinytype)-0.5), uhi(nu * (N - 1), (tinytype)0.5);
    TinySolver *s = nullptr;
    if (tiny_setup(&s, nx, nu, N, A.data(), B.data(), Q.data(), R.data(), (tinytype)d.scalars.at("rho"), xlo.data(), xhi.data(),
                   ulo.data(), uhi.data(), 0) != 0) { fprintf(stderr, "%s\n", tiny_last_error()); return 1; }
    {
        const auto K = d.cast<tinytype>("Kinf"), P = d.cast<tinytype>("Pinf"), Qi = d.cast<tinytype>("Quu_inv"), M = d.cast<tinytype>("AmBKt");
        std::memcpy(s->cache->Kinf.data(), K.data(), sizeof(tinytype) * K.size());
        std::memcpy(s->cache->Pinf.data(), P.data(), sizeof(tinytype) * P.size());
        std::memcpy(s->cache->Quu_inv.data(), Qi.data(), sizeof(tinytype) * Qi.size());
        std::memcpy(s->cache->AmBKt.data(), M.data(), sizeof(tinytype) * M.size());
    }
    // (every visible device takes part: tiny_rollout_batch and tiny_solve_batch split host batches of >= 32,768 instances)
    const double scale[12] = {2, 2, 2, .2, .2, .2, .5, .5, .5, .5, .5, .5}, hover[12] = {0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<tinytype> x0((size_t)Bn * nx), xref((size_t)N * nx, 0);
    for (int i = 0; i < N; ++i) xref[(size_t)i * nx + 2] = 2;
    for (int64_t b = 0; b < Bn; ++b)
        for (int j = 0; j < nx; ++j) x0[(size_t)b * nx + j] = (tinytype)(hover[j] + 0.25 * scale[j] * (2.0 * u01(1234, (uint64_t)(12 * b + j)) - 1.0));

    // ---- (1) one call
    std::vector<tinytype> xh((size_t)(steps + 1) * Bn * nx), uh((size_t)steps * Bn * nu), xl((size_t)Bn * N * nx), ul((size_t)Bn * (N - 1) * nu);
    std::vector<int32_t> ih((size_t)steps * Bn), sh((size_t)steps * Bn);
    TinyRolloutIn rin;
    std::memset(&rin, 0, sizeof rin);
    rin.batch = Bn; rin.steps = steps; rin.reset_duals = 1; rin.x0 = x0.data(); rin.Xref = xref.data(); rin.xref_shared = 1;
    TinyRolloutOut rout;
    std::memset(&rout, 0, sizeof rout);
    rout.x_hist = xh.data(); rout.u0_hist = uh.data(); rout.iter_hist = ih.data(); rout.status_hist = sh.data(); rout.x = xl.data(); rout.u = ul.data();
    double ms_one = 0;
    for (int rep = 0; rep < 2; ++rep) {   // the first call creates the context
        const auto t0 = std::chrono::steady_clock::now();
        if (tiny_rollout_batch(s, &rin, &rout) != 0) { fprintf(stderr, "tiny_rollout_batch: %s\n", tiny_last_error()); return 1; }
        ms_one = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }

    // ---- (2) the loop of the examples, one tiny_solve_batch per step
    std::vector<tinytype> xc = x0, x((size_t)Bn * N * nx), u((size_t)Bn * (N - 1) * nu);
    std::vector<tinytype> wd((size_t)Bn * (N - 1) * nu, 0), wy(wd), wz(wd), wg((size_t)Bn * N * nx, 0), wv(wg);
    std::vector<int32_t> it(Bn), st(Bn);
    long long bad = 0, iters = 0;
    const auto t1 = std::chrono::steady_clock::now();
    for (int k = 0; k < steps; ++k) {
        std::fill(wy.begin(), wy.end(), (tinytype)0);      // quadrotor_hovering.cpp:100-101
        std::fill(wg.begin(), wg.end(), (tinytype)0);
        TinyBatchIn in;
        std::memset(&in, 0, sizeof in);
        in.batch = Bn; in.x0 = xc.data(); in.Xref = xref.data(); in.xref_shared = 1;
        in.d = wd.data(); in.y = wy.data(); in.g = wg.data(); in.v = wv.data(); in.z = wz.data();
        TinyBatchOut out;
        std::memset(&out, 0, sizeof out);
        out.x = x.data(); out.u = u.data(); out.iter = it.data(); out.status = st.data();
        if (tiny_solve_batch(s, &in, &out) != 0) { fprintf(stderr, "tiny_solve_batch: %s\n", tiny_last_error()); return 1; }
        for (int64_t b = 0; b < Bn; ++b) {
            const tinytype *ub = &u[(size_t)b * (N - 1) * nu];
            bad += it[b] != ih[(size_t)k * Bn + b];
            bad += st[b] != sh[(size_t)k * Bn + b];
            bad += std::memcmp(ub, &uh[((size_t)k * Bn + b) * nu], sizeof(tinytype) * nu) != 0;
            iters += it[b];
        }
        // next measurement: the plant state the one-call loop recorded (its plant step is checked against the reference's own in
        // tests/test_gpu_batch.py; this example checks the SOLVES of the fused loop against the step-by-step path)
        std::memcpy(xc.data(), &xh[(size_t)(k + 1) * Bn * nx], sizeof(tinytype) * xc.size());
    }
    const double ms_loop = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
    bad += std::memcmp(x.data(), xl.data(), sizeof(tinytype) * x.size()) != 0;
    bad += std::memcmp(u.data(), ul.data(), sizeof(tinytype) * u.size()) != 0;
    printf("instances %lld steps %d iterations %lld  one call %.1f ms  loop of tiny_solve_batch %.1f ms  mismatches %lld\n", (long long)Bn, steps, iters,
           ms_one, ms_loop, bad);
    tiny_free(s);
    if (bad) { printf("rollout batch MISMATCH\n"); return 1; }
    printf("rollout batch ok\n");
    return 0;
}
