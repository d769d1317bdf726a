// The reference's examples/codegen_random.cpp problem (nx = 2, nu = 2, N = 3, its data :20-39 incl. the inverted box
// bounds) through the host API on the GPU: tiny_setup + tiny_precompute (what tiny_codegen computes, codegen.cpp:254-292)
// and one tiny_solve (what the generated tiny_main does).  The shape has no specialised kernel: it runs the
// run-time-shape kernel with the evaluation order Eigen picks for 2/2/3.
//   usage: codegen_random [x0_0 x0_1]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tinympc/tiny_api.hpp"

int main(int argc, char **argv)
{
    const int n = 2, m = 2, N = 3;
    const tinytype A[n * n] = {1, 5, 1, 2}, B[n * m] = {3, 3, 4, 1}, Q[n] = {1, 1}, R[m] = {2, 2};   // column-major
    const tinytype rho = (tinytype)0.1;
    const tinytype xlo[n * N] = {1, 2, 1, 2, 1, 2}, xhi[n * N] = {-1, -2, -1, -2, -1, -2};
    const tinytype ulo[m * (N - 1)] = {2, 3, 2, 3}, uhi[m * (N - 1)] = {-2, -3, -2, -3};
    TinySolver *s = nullptr;
    if (tiny_setup(&s, n, m, N, A, B, Q, R, rho, xlo, xhi, ulo, uhi, 0) != 0) {
        fprintf(stderr, "%s\n", tiny_last_error());
        return 1;
    }
    const int sweeps = tiny_precompute(s);
    printf("riccati sweeps %d\n", sweeps);
    for (int i = 0; i < n; ++i) s->work->Q(i) += rho;   // generated code stores Q + rho (codegen.cpp:255,433)
    tinytype x0[n] = {(tinytype)(argc > 2 ? atof(argv[1]) : 0.5), (tinytype)(argc > 2 ? atof(argv[2]) : -0.3)};
    s->work->x.setCol(0, x0);
    const int rc = tiny_solve(s);
    if (rc < 0) { fprintf(stderr, "tiny_solve failed: %s\n", tiny_last_error()); return 1; }
    printf("rc %d iter %d status %d\n", rc, s->work->iter, s->work->status);
    printf("u0 %.17g %.17g\n", (double)s->work->u(0, 0), (double)s->work->u(1, 0));
    printf("xN %.17g %.17g\n", (double)s->work->x(0, N - 1), (double)s->work->x(1, N - 1));
    printf("Kinf %.12g %.12g %.12g %.12g\n", (double)s->cache->Kinf(0, 0), (double)s->cache->Kinf(1, 0), (double)s->cache->Kinf(0, 1),
           (double)s->cache->Kinf(1, 1));
    tiny_free(s);
    return 0;
}
