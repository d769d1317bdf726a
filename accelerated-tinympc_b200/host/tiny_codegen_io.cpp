// Codegen-compatible data import / export: read and write the generated-code data files of the
// reference's tiny_codegen (/root/reference/src/tinympc/codegen.cpp:322-477 "tiny_data_workspace.cpp": TinySettings,
// TinyCache and TinyWorkspace as Eigen comma initialisers in ROW-MAJOR order, "(tinytype)%.16f", codegen.cpp:118-129;
// and codegen.cpp:131-160 "glob_opts.hpp").  A cache computed here can be dropped into a generated embedded project,
// and a generated project's cache can be solved on the GPU, without Eigen on either side.  Host-only, cold path.
#include "tinympc/tiny_api.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

extern "C" const char *tiny_last_error(void);
namespace tinyhost { int set_error(const std::string &m); }

namespace {

void print_matrix(FILE *f, const tiny_Matrix &m)   // codegen.cpp:118-129: mat.reshaped<RowMajor>()
{
    const int n = m.rows() * m.cols();
    int k = 0;
    for (int i = 0; i < m.rows(); ++i)
        for (int j = 0; j < m.cols(); ++j, ++k) fprintf(f, "(tinytype)%.16f%s", (double)m(i, j), k < n - 1 ? "," : "");
}
void entry(FILE *f, const char *type, const tiny_Matrix &m, const char *name, bool last = false)
{
    fprintf(f, "\t(%s() << ", type);
    print_matrix(f, m);
    fprintf(f, ").finished()%s\t// %s\n", last ? "" : ",", name);
}

struct Parsed {
    std::vector<double> scalars;                       // plain numeric lines in order of appearance
    std::map<std::string, std::vector<double>> mats;   // by trailing comment label
    std::vector<std::string> order;
};

// one line of the data file: "\t(tiny_MatrixNuNx() << (tinytype)a,(tinytype)b).finished(),\t// Kinf" or "\t(tinytype)0.1,\t// rho"
bool parse_file(const char *path, Parsed &P, std::string &err)
{
    FILE *f = fopen(path, "r");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    std::string line;
    int ch;
    std::vector<std::string> lines;
    while ((ch = fgetc(f)) != EOF) {
        if (ch == '\n') { lines.push_back(line); line.clear(); }
        else line.push_back((char)ch);
    }
    if (!line.empty()) lines.push_back(line);
    fclose(f);
    for (const std::string &ln : lines) {
        const size_t c = ln.find("//");
        if (c == std::string::npos) continue;
        std::string label = ln.substr(c + 2);
        while (!label.empty() && label.front() == ' ') label.erase(label.begin());
        while (!label.empty() && (label.back() == ' ' || label.back() == '\r')) label.pop_back();
        const std::string body = ln.substr(0, c);
        std::vector<double> vals;
        const size_t lsh = body.find("<<");
        size_t pos = lsh == std::string::npos ? 0 : lsh + 2;
        while (pos < body.size()) {
            const size_t t = body.find("(tinytype)", pos);
            const char *start;
            if (t != std::string::npos) { start = body.c_str() + t + 10; pos = t + 10; }
            else if (lsh == std::string::npos && vals.empty()) {   // "\t100,\t\t" integer entry
                start = body.c_str() + pos;
                while (*start == '\t' || *start == ' ') ++start;
            } else break;
            char *end = nullptr;
            const double v = strtod(start, &end);
            if (end == start) break;
            vals.push_back(v);
            pos = (size_t)(end - body.c_str());
            if (t == std::string::npos) break;
        }
        if (vals.empty()) continue;
        P.mats[label] = vals;
        P.order.push_back(label);
    }
    return true;
}

bool fill(tiny_Matrix &m, int r, int c, const Parsed &P, const char *label, std::string &err)   // row-major in the file
{
    auto it = P.mats.find(label);
    if (it == P.mats.end() || (int)it->second.size() != r * c) {
        err = std::string("entry '") + label + "' missing or of the wrong size";
        return false;
    }
    m.resize(r, c);
    int k = 0;
    for (int i = 0; i < r; ++i)
        for (int j = 0; j < c; ++j, ++k) m(i, j) = (tinytype)it->second[k];
    return true;
}

}  // namespace

extern "C" {

int tiny_export_data_workspace(const TinySolver *s, const char *path)
{
    if (!s || !path) return tinyhost::set_error("tiny_export_data_workspace: NULL argument");
    FILE *f = fopen(path, "w");
    if (!f) return tinyhost::set_error(std::string("tiny_export_data_workspace: cannot open ") + path);
    const TinySettings &t = *s->settings;
    const TinyCache &c = *s->cache;
    const TinyWorkspace &w = *s->work;
    const int nx = s->nx, nu = s->nu, N = s->N;
    fprintf(f, "/*\n */\n\n#include <tinympc/tiny_data_workspace.hpp>\n\n#ifdef __cplusplus\nextern \"C\" {\n#endif\n\n");
    fprintf(f, "/* User settings */\nTinySettings settings = {\n");
    fprintf(f, "\t(tinytype)%.16f,\t// primal tolerance\n", (double)t.abs_pri_tol);
    fprintf(f, "\t(tinytype)%.16f,\t// dual tolerance\n", (double)t.abs_dua_tol);
    fprintf(f, "\t%d,\t\t// max iterations\n", t.max_iter);
    fprintf(f, "\t%d,\t\t// iterations per termination check\n", t.check_termination);
    fprintf(f, "\t%d,\t\t// enable state constraints\n", t.en_state_bound);
    fprintf(f, "\t%d\t\t// enable input constraints\n", t.en_input_bound);
    fprintf(f, "};\n\n");
    fprintf(f, "/* Matrices that must be recomputed with changes in time step, rho */\nTinyCache cache = {\n");
    fprintf(f, "\t(tinytype)%.16f,\t// rho (step size/penalty)\n", (double)c.rho);
    entry(f, "tiny_MatrixNuNx", c.Kinf, "Kinf");
    entry(f, "tiny_MatrixNxNx", c.Pinf, "Pinf");
    entry(f, "tiny_MatrixNuNu", c.Quu_inv, "Quu_inv");
    entry(f, "tiny_MatrixNxNx", c.AmBKt, "AmBKt");
    entry(f, "tiny_MatrixNxNu", c.coeff_d2p, "coeff_d2p");
    fprintf(f, "};\n\n");
    fprintf(f, "/* Problem variables */\nTinyWorkspace work = {\n");
    const tiny_Matrix zx(nx, N), zu(nu, N - 1), zq(nu, 1);
    // the generator always writes zero work arrays (codegen.cpp:386-429); so do we: a data file is a cold start
    for (const char *nm : {"x", "u", "q", "r", "p", "d", "v", "vnew", "z", "znew", "g", "y"}) {
        const bool isx = !strcmp(nm, "x") || !strcmp(nm, "q") || !strcmp(nm, "p") || !strcmp(nm, "v") || !strcmp(nm, "vnew") || !strcmp(nm, "g");
        entry(f, isx ? "tiny_MatrixNxNh" : "tiny_MatrixNuNhm1", isx ? zx : zu, nm);
    }
    fprintf(f, "\t(tinytype)%.16f,\t// state primal residual\n", 0.0);
    fprintf(f, "\t(tinytype)%.16f,\t// input primal residual\n", 0.0);
    fprintf(f, "\t(tinytype)%.16f,\t// state dual residual\n", 0.0);
    fprintf(f, "\t(tinytype)%.16f,\t// input dual residual\n", 0.0);
    fprintf(f, "\t%d,\t// solve status\n", 0);
    fprintf(f, "\t%d,\t// solve iteration\n", 0);
    entry(f, "tiny_VectorNx", w.Q, "Q");
    entry(f, "tiny_VectorNu", w.R, "R");
    entry(f, "tiny_MatrixNxNx", w.Adyn, "Adyn");
    entry(f, "tiny_MatrixNxNu", w.Bdyn, "Bdyn");
    entry(f, "tiny_MatrixNuNhm1", w.u_min, "u_min");
    entry(f, "tiny_MatrixNuNhm1", w.u_max, "u_max");
    entry(f, "tiny_MatrixNxNh", w.x_min, "x_min");
    entry(f, "tiny_MatrixNxNh", w.x_max, "x_max");
    entry(f, "tiny_MatrixNxNh", zx, "Xref");
    entry(f, "tiny_MatrixNuNhm1", zu, "Uref");
    entry(f, "tiny_VectorNu", zq, "Qu", true);
    fprintf(f, "};\n\n");
    fprintf(f, "TinySolver tiny_data_solver = {&settings, &cache, &work};\n\n#ifdef __cplusplus\n}\n#endif\n\n");
    fclose(f);
    return 0;
}

int tiny_export_glob_opts(const TinySolver *s, const char *path)
{
    if (!s || !path) return tinyhost::set_error("tiny_export_glob_opts: NULL argument");
    FILE *f = fopen(path, "w");
    if (!f) return tinyhost::set_error(std::string("tiny_export_glob_opts: cannot open ") + path);
    // codegen.cpp:146-155; the generator hard-codes float there, we write the scalar type this library was built for
    fprintf(f, "/*\n */\n\n#pragma once\n\ntypedef %s tinytype;\n\n#define NSTATES %d\n#define NINPUTS %d\n#define NHORIZON %d",
            sizeof(tinytype) == 4 ? "float" : "double", s->nx, s->nu, s->N);
    fclose(f);
    return 0;
}

int tiny_import_data_workspace(TinySolver **out, const char *path)
{
    if (!out || !path) return tinyhost::set_error("tiny_import_data_workspace: NULL argument");
    Parsed P;
    std::string err;
    if (!parse_file(path, P, err)) return tinyhost::set_error("tiny_import_data_workspace: " + err);
    auto need = [&](const char *label) -> const std::vector<double> * {
        auto it = P.mats.find(label);
        return it == P.mats.end() ? nullptr : &it->second;
    };
    const std::vector<double> *pinf = need("Pinf"), *quu = need("Quu_inv"), *x = need("x");
    if (!pinf || !quu || !x) return tinyhost::set_error("tiny_import_data_workspace: not a tiny_data_workspace file (Pinf / Quu_inv / x missing)");
    const int nx = (int)std::lround(std::sqrt((double)pinf->size())), nu = (int)std::lround(std::sqrt((double)quu->size()));
    if (nx < 1 || nu < 1 || (size_t)nx * nx != pinf->size() || (size_t)nu * nu != quu->size() || x->size() % nx)
        return tinyhost::set_error("tiny_import_data_workspace: inconsistent matrix sizes");
    const int N = (int)(x->size() / nx);
    std::vector<tinytype> z((size_t)nx * nx + (size_t)nx * nu + nx + nu, 0);
    TinySolver *s = nullptr;
    // sizes first (tiny_setup zero-fills every array), then every entry of the file
    if (tiny_setup(&s, nx, nu, N, z.data(), z.data(), z.data(), z.data(), 0, nullptr, nullptr, nullptr, nullptr, 0) != 0) return -1;
    TinyCache &c = *s->cache;
    TinyWorkspace &w = *s->work;
    TinySettings &t = *s->settings;
    bool ok = fill(c.Kinf, nu, nx, P, "Kinf", err) && fill(c.Pinf, nx, nx, P, "Pinf", err) && fill(c.Quu_inv, nu, nu, P, "Quu_inv", err) &&
              fill(c.AmBKt, nx, nx, P, "AmBKt", err) && fill(c.coeff_d2p, nx, nu, P, "coeff_d2p", err) && fill(w.Q, nx, 1, P, "Q", err) &&
              fill(w.R, nu, 1, P, "R", err) && fill(w.Adyn, nx, nx, P, "Adyn", err) && fill(w.Bdyn, nx, nu, P, "Bdyn", err) &&
              fill(w.u_min, nu, N - 1, P, "u_min", err) && fill(w.u_max, nu, N - 1, P, "u_max", err) &&
              fill(w.x_min, nx, N, P, "x_min", err) && fill(w.x_max, nx, N, P, "x_max", err) && fill(w.Xref, nx, N, P, "Xref", err);
    auto scalar = [&](const char *label, double &v) {
        auto it = P.mats.find(label);
        if (it == P.mats.end() || it->second.size() != 1) { ok = false; err = std::string("entry '") + label + "' missing"; return; }
        v = it->second[0];
    };
    double rho = 0, pri = 0, dua = 0, mi = 0, ct = 0, es = 0, ei = 0;
    scalar("rho (step size/penalty)", rho); scalar("primal tolerance", pri); scalar("dual tolerance", dua);
    scalar("max iterations", mi); scalar("iterations per termination check", ct);
    scalar("enable state constraints", es); scalar("enable input constraints", ei);
    if (!ok) { tiny_free(s); return tinyhost::set_error("tiny_import_data_workspace: " + err); }
    c.rho = (tinytype)rho;
    t.abs_pri_tol = (tinytype)pri; t.abs_dua_tol = (tinytype)dua; t.max_iter = (int)mi; t.check_termination = (int)ct;
    t.en_state_bound = (int)es; t.en_input_bound = (int)ei;
    *out = s;
    return 0;
}

}  // extern "C"
