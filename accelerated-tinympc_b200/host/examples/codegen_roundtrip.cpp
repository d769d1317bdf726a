// Codegen-compatible data files, both directions (no GPU needed):
//   codegen_roundtrip import <tiny_data_workspace.cpp> <out.cpp> [<glob_opts out>]   parse a generated data file and write it back
//   codegen_roundtrip fresh  <model.mpcdata> <N> <xbound> <ubound> <out.cpp>         tiny_setup + tiny_precompute, then export what
//                                                                                    tiny_codegen would emit (work.Q/R = Q+rho, R+rho)
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "mpcdata.hpp"
#include "tinympc/tiny_api.hpp"

int main(int argc, char **argv)
{
    if (argc >= 4 && !strcmp(argv[1], "import")) {
        TinySolver *s = nullptr;
        if (tiny_import_data_workspace(&s, argv[2]) != 0) { fprintf(stderr, "%s\n", tiny_last_error()); return 1; }
        printf("imported nx=%d nu=%d N=%d rho=%.17g max_iter=%d Kinf(0,0)=%.17g\n", s->nx, s->nu, s->N, (double)s->cache->rho,
               s->settings->max_iter, (double)s->cache->Kinf(0, 0));
        if (tiny_export_data_workspace(s, argv[3]) != 0) { fprintf(stderr, "%s\n", tiny_last_error()); return 1; }
        if (argc > 4 && tiny_export_glob_opts(s, argv[4]) != 0) { fprintf(stderr, "%s\n", tiny_last_error()); return 1; }
        tiny_free(s);
        return 0;
    }
    if (argc >= 7 && !strcmp(argv[1], "fresh")) {
        MpcData d(argv[2]);
        const int nx = (int)d.scalars.at("nx"), nu = (int)d.scalars.at("nu"), N = atoi(argv[3]);
        const tinytype xb = (tinytype)atof(argv[4]), ub = (tinytype)atof(argv[5]);
        const auto A = d.cast<tinytype>("Adyn"), B = d.cast<tinytype>("Bdyn"), Q = d.cast<tinytype>("Q"), R = d.cast<tinytype>("R");
        std::vector<tinytype> xlo(nx * N, -xb), xhi(nx * N, xb), ulo(nu * (N - 1), -ub), uhi(nu * (N - 1), ub);
        TinySolver *s = nullptr;
        const tinytype rho = (tinytype)d.scalars.at("rho");
        if (tiny_setup(&s, nx, nu, N, A.data(), B.data(), Q.data(), R.data(), rho, xlo.data(), xhi.data(), ulo.data(), uhi.data(), 0) != 0) {
            fprintf(stderr, "%s\n", tiny_last_error());
            return 1;
        }
        const int sweeps = tiny_precompute(s);
        printf("precompute: %d sweeps\n", sweeps);
        for (int i = 0; i < nx; ++i) s->work->Q(i) += rho;   // what tiny_codegen stores (codegen.cpp:255-256)
        for (int i = 0; i < nu; ++i) s->work->R(i) += rho;
        if (tiny_export_data_workspace(s, argv[6]) != 0) { fprintf(stderr, "%s\n", tiny_last_error()); return 1; }
        tiny_free(s);
        return 0;
    }
    fprintf(stderr, "usage: %s import <data.cpp> <out.cpp> [glob_opts out] | fresh <model.mpcdata> <N> <xbound> <ubound> <out.cpp>\n", argv[0]);
    return 2;
}
