// tiny_setup + tiny_precompute on the reference's cartpole model (examples/codegen_cartpole.cpp:22-28), then the
// 300-step closed loop that file carries for generated code (:75-122, max_iter 150), solved on the GPU.
//   usage: cartpole_precompute <problem_data dir>
#include <cstdio>
#include <map>
#include <vector>

#include "mpcdata.hpp"
#include "tinympc/tiny_api.hpp"

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s <problem_data dir>\n", argv[0]); return 2; }
    const int n = 4, m = 1, N = 10;
    MpcData d(std::string(argv[1]) + "/cartpole.mpcdata");
    const auto A = d.cast<tinytype>("Adyn"), B = d.cast<tinytype>("Bdyn"), Q = d.cast<tinytype>("Q"), R = d.cast<tinytype>("R");
    const tinytype rho = (tinytype)d.scalars.at("rho");
    std::vector<tinytype> xlo(n * N, -5), xhi(n * N, 5), ulo(m * (N - 1), -5), uhi(m * (N - 1), 5);
    TinySolver *s = nullptr;
    if (tiny_setup(&s, n, m, N, A.data(), B.data(), Q.data(), R.data(), rho, xlo.data(), xhi.data(), ulo.data(), uhi.data(), 0) != 0) {
        fprintf(stderr, "%s\n", tiny_last_error());
        return 1;
    }
    const int sweeps = tiny_precompute(s);
    printf("Kinf converged after %d iterations\n", sweeps);
    printf("Kinf = %.10f %.10f %.10f %.10f\n", (double)s->cache->Kinf(0, 0), (double)s->cache->Kinf(0, 1),
           (double)s->cache->Kinf(0, 2), (double)s->cache->Kinf(0, 3));
    printf("Quu_inv = %.10f\n", (double)s->cache->Quu_inv(0, 0));
    // generated code stores Q + rho in work.Q (codegen.cpp:255,433)
    for (int i = 0; i < n; ++i) s->work->Q(i) += rho;
    s->settings->max_iter = 150;
    std::vector<tinytype> x0 = {0, 0, (tinytype)0.1, 0};
    std::map<int, int> hist;
    for (int k = 0; k < 300; ++k) {
        s->work->x.setCol(0, x0.data());
        s->work->y.setZero();
        s->work->g.setZero();
        if (tiny_solve(s) < 0) { fprintf(stderr, "tiny_solve failed: %s\n", tiny_last_error()); return 1; }
        hist[s->work->iter]++;
        std::vector<tinytype> x1(n);
        for (int r = 0; r < n; ++r) {
            tinytype acc = 0;
            for (int c = 0; c < n; ++c) acc += s->work->Adyn(r, c) * x0[c];
            acc += s->work->Bdyn(r, 0) * s->work->u(0, 0);
            x1[r] = acc;
        }
        x0 = x1;
    }
    printf("iteration histogram:");
    for (auto &kv : hist) printf(" %d:%d", kv.first, kv.second);
    printf("\n");
    tiny_free(s);
    return 0;
}
