// Closed-loop MPC drivers on the host C++ API: the same experiments as the reference's quadrotor examples
// (examples/quadrotor_hovering.cpp:83-114 -- 70 steps towards a hover set-point; examples/quadrotor_tracking.cpp:
// 84-118 -- 290 steps along the y-axis line), every tiny_solve executed on the GPU.
//   usage: quadrotor_loops <problem_data dir> hover|track [-v]
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "mpcdata.hpp"
#include "tinympc/tiny_api.hpp"

static void load_cache(TinySolver *s, const MpcData &d)
{
    auto put = [&](tiny_Matrix &dst, const char *name) {
        const auto v = d.cast<tinytype>(name);
        std::memcpy(dst.data(), v.data(), sizeof(tinytype) * v.size());
    };
    put(s->cache->Kinf, "Kinf"); put(s->cache->Pinf, "Pinf"); put(s->cache->Quu_inv, "Quu_inv"); put(s->cache->AmBKt, "AmBKt");
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s <problem_data dir> hover|track [-v]\n", argv[0]); return 2; }
    const std::string dir = argv[1], mode = argv[2];
    const bool verbose = argc > 3 && !strcmp(argv[3], "-v");
    const int nx = 12, nu = 4, N = 10;
    MpcData d(dir + "/quadrotor_20hz.mpcdata");
    const auto A = d.cast<tinytype>("Adyn"), B = d.cast<tinytype>("Bdyn"), Q = d.cast<tinytype>("Q"), R = d.cast<tinytype>("R");
    std::vector<tinytype> xlo(nx * N, -5), xhi(nx * N, 5), ulo(nu * (N - 1), (tinytype)-0.5), uhi(nu * (N - 1), (tinytype)0.5);
    TinySolver *s = nullptr;
    if (tiny_setup(&s, nx, nu, N, A.data(), B.data(), Q.data(), R.data(), (tinytype)d.scalars.at("rho"), xlo.data(), xhi.data(),
                   ulo.data(), uhi.data(), 0) != 0) { fprintf(stderr, "%s\n", tiny_last_error()); return 1; }
    load_cache(s, d);   // the examples use the shipped cache, not a fresh precompute
    TinyWorkspace &w = *s->work;

    std::vector<tinytype> table;   // 12 x 301 column-major
    int steps = 70;
    std::vector<tinytype> x0(nx, 0);
    if (mode == "hover") {
        for (int i = 0; i < N; ++i) w.Xref(2, i) = 2;                        // hover at z = 2
        const tinytype init[12] = {0, 1, 0, (tinytype)0.2, 0, 0, (tinytype)0.1, 0, 0, 0, 0, 0};
        x0.assign(init, init + 12);
    } else {
        MpcData t(dir + "/quadrotor_20hz_y_axis_line.mpcdata");
        table = t.cast<tinytype>("Xref_total");
        steps = 301 - N - 1;
        for (int i = 0; i < N; ++i) w.Xref.setCol(i, &table[(size_t)i * nx]);
        x0.assign(table.begin(), table.begin() + nx);
    }
    for (int k = 0; k < steps; ++k) {
        double e2 = 0;
        for (int j = 0; j < nx; ++j) { const double e = (double)x0[j] - (double)w.Xref(j, 1); e2 += e * e; }
        if (mode == "hover") printf("tracking error at step %2d: %.4f\n", k, std::sqrt(e2));
        else printf("tracking error: %g\n", std::sqrt(e2));
        w.x.setCol(0, x0.data());                                            // 1. measurement
        if (mode != "hover")                                                 // 2. reference window
            for (int i = 0; i < N; ++i) w.Xref.setCol(i, &table[(size_t)(k + i) * nx]);
        w.y.setZero();                                                       // 3. reset duals
        w.g.setZero();
        const int rc = tiny_solve(s);                                        // 4. solve on the GPU
        if (rc < 0) { fprintf(stderr, "tiny_solve failed: %s\n", tiny_last_error()); return 1; }
        if (verbose) printf("  iter %d status %d u0 % .9g % .9g % .9g % .9g\n", w.iter, w.status, (double)w.u(0, 0),
                            (double)w.u(1, 0), (double)w.u(2, 0), (double)w.u(3, 0));
        std::vector<tinytype> x1(nx);                                        // 5. plant: x+ = A x + B u0
        for (int r = 0; r < nx; ++r) {
            tinytype acc = 0;
            for (int c = 0; c < nx; ++c) acc += w.Adyn(r, c) * x0[c];
            for (int c = 0; c < nu; ++c) acc += w.Bdyn(r, c) * w.u(c, 0);
            x1[r] = acc;
        }
        x0 = x1;
    }
    tiny_free(s);
    return 0;
}
