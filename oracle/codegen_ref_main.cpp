// TEST INFRASTRUCTURE ONLY.  Drives the UNMODIFIED reference code generator (src/tinympc/codegen.cpp:218
// tiny_codegen) on a problem read from a text file, so that oracle/make_golden.py can commit what the reference
// emits (src/tiny_data_workspace.cpp, tinympc/glob_opts.hpp) as golden files for the codegen-compatible
// import/export of the product (include/tinympc/tiny_api.hpp tiny_export_data_workspace / tiny_import_data_workspace).
//   usage: codegen_ref <problem.txt> <tinympc_dir> <output_dir>
//   problem.txt: nx nu N rho abs_pri_tol abs_dua_tol max_iter check_termination, then column-major
//                Adyn[nx*nx] Bdyn[nx*nu] Q[nx] R[nu] x_min[nx*N] x_max[nx*N] u_min[nu*(N-1)] u_max[nu*(N-1)]
#include <tinympc/codegen.hpp>

#include <cstdio>
#include <vector>

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    FILE *f = fopen(argv[1], "r");
    if (!f) return 3;
    int nx, nu, N, max_iter, check;
    double rho, pri, dua;
    if (fscanf(f, "%d %d %d %lf %lf %lf %d %d", &nx, &nu, &N, &rho, &pri, &dua, &max_iter, &check) != 8) return 4;
    auto rd = [&](int n) {
        std::vector<tinytype> v(n);
        for (int i = 0; i < n; ++i) { double t; if (fscanf(f, "%lf", &t) != 1) exit(5); v[i] = (tinytype)t; }
        return v;
    };
    auto A = rd(nx * nx), B = rd(nx * nu), Q = rd(nx), R = rd(nu), xmin = rd(nx * N), xmax = rd(nx * N), umin = rd(nu * (N - 1)),
         umax = rd(nu * (N - 1));
    fclose(f);
    return tiny_codegen(nx, nu, N, A.data(), B.data(), Q.data(), R.data(), xmin.data(), xmax.data(), umin.data(), umax.data(),
                        (tinytype)rho, (tinytype)pri, (tinytype)dua, max_iter, check, 0, argv[2], argv[3]) == 1 ? 0 : 1;
}
