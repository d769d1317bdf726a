// tiny_set_instance_bounds through the host API: the wrapper's set_xmin / set_xmax / set_umin / set_umax
// (tiny_wrapper.cpp:43-129) with a leading batch dimension in front of tiny_solve_batch.  Self-checking on the 2/2/3
// system of the reference's examples/codegen_random.cpp (run-time-shape kernel):
//   (1) every instance given the shared box         -> bit-identical to the shared-bounds solve
//   (2) odd instances given a tighter input box     -> even instances unchanged, odd ones obey their own box in z-space
//   (3) bounds cleared (batch = 0)                   -> bit-identical to the shared-bounds solve again
#include <cstdio>
#include <cstring>
#include <vector>

#include "tinympc/tiny_api.hpp"

static int solve(TinySolver *s, int64_t B, const std::vector<tinytype> &x0, const std::vector<tinytype> &xref, std::vector<tinytype> &x,
                 std::vector<tinytype> &u, std::vector<int32_t> &it)
{
    std::vector<int32_t> st(B);
    TinyBatchIn in;
    std::memset(&in, 0, sizeof in);
    in.batch = B; in.x0 = x0.data(); in.Xref = xref.data(); in.xref_shared = 1;
    TinyBatchOut out;
    std::memset(&out, 0, sizeof out);
    out.x = x.data(); out.u = u.data(); out.iter = it.data(); out.status = st.data();
    return tiny_solve_batch(s, &in, &out);
}

int main()
{
    const int n = 2, m = 2, N = 3;
    const int64_t Bn = 64;
    const tinytype A[n * n] = {1, 5, 1, 2}, B[n * m] = {3, 3, 4, 1}, Q[n] = {1, 1}, R[m] = {2, 2};   // column-major
    const tinytype rho = (tinytype)0.1;
    tinytype xlo[n * N], xhi[n * N], ulo[m * (N - 1)], uhi[m * (N - 1)];
    for (int k = 0; k < n * N; ++k) { xlo[k] = -2; xhi[k] = 2; }
    for (int k = 0; k < m * (N - 1); ++k) { ulo[k] = (tinytype)-0.3; uhi[k] = (tinytype)0.3; }
    TinySolver *s = nullptr;
    if (tiny_setup(&s, n, m, N, A, B, Q, R, rho, xlo, xhi, ulo, uhi, 0) != 0) { fprintf(stderr, "%s\n", tiny_last_error()); return 1; }
    tiny_precompute(s);
    std::vector<tinytype> x0(Bn * n), xref(N * n, 0), xa(Bn * N * n), ua(Bn * (N - 1) * m), xb(xa.size()), ub(ua.size());
    std::vector<int32_t> ia(Bn), ib(Bn);
    for (int64_t i = 0; i < Bn; ++i) { x0[i * n] = (tinytype)(0.02 * (i - 30)); x0[i * n + 1] = (tinytype)(-0.01 * (i % 7)); }
    if (solve(s, Bn, x0, xref, xa, ua, ia) != 0) { fprintf(stderr, "shared: %s\n", tiny_last_error()); return 1; }
    auto same = [&](const char *what) {
        if (xa != xb || ua != ub || ia != ib) { printf("FAIL %s\n", what); return false; }
        return true;
    };
    // (1)
    std::vector<tinytype> bxl(Bn * N * n, -2), bxh(Bn * N * n, 2), bul(Bn * (N - 1) * m, (tinytype)-0.3), buh(Bn * (N - 1) * m, (tinytype)0.3);
    if (tiny_set_instance_bounds(s, Bn, bxl.data(), bxh.data(), bul.data(), buh.data(), 0) != 0) { fprintf(stderr, "%s\n", tiny_last_error()); return 1; }
    if (solve(s, Bn, x0, xref, xb, ub, ib) != 0) { fprintf(stderr, "per-instance: %s\n", tiny_last_error()); return 1; }
    if (!same("same boxes")) return 2;
    // (2)
    for (int64_t i = 1; i < Bn; i += 2)
        for (int k = 0; k < (N - 1) * m; ++k) { bul[i * (N - 1) * m + k] = (tinytype)-0.05; buh[i * (N - 1) * m + k] = (tinytype)0.05; }
    if (tiny_set_instance_bounds(s, Bn, bxl.data(), bxh.data(), bul.data(), buh.data(), 0) != 0) { fprintf(stderr, "%s\n", tiny_last_error()); return 1; }
    if (solve(s, Bn, x0, xref, xb, ub, ib) != 0) { fprintf(stderr, "tight: %s\n", tiny_last_error()); return 1; }
    int changed = 0;
    for (int64_t i = 0; i < Bn; ++i) {
        const bool eq = std::memcmp(&ua[i * (N - 1) * m], &ub[i * (N - 1) * m], sizeof(tinytype) * (N - 1) * m) == 0 && ia[i] == ib[i];
        if (i % 2 == 0 && !eq) { printf("FAIL even instance %lld changed\n", (long long)i); return 2; }
        if (i % 2 == 1 && !eq) ++changed;
    }
    if (changed == 0) { printf("FAIL tighter boxes changed nothing\n"); return 2; }
    // a batch of another size is refused while the boxes are set
    {
        std::vector<tinytype> x1(x0.begin(), x0.begin() + n);
        std::vector<tinytype> xo(N * n), uo((N - 1) * m);
        std::vector<int32_t> io(1);
        if (solve(s, 1, x1, xref, xo, uo, io) == 0) { printf("FAIL batch mismatch accepted\n"); return 2; }
    }
    // (3)
    if (tiny_set_instance_bounds(s, 0, nullptr, nullptr, nullptr, nullptr, 0) != 0) { fprintf(stderr, "%s\n", tiny_last_error()); return 1; }
    if (solve(s, Bn, x0, xref, xb, ub, ib) != 0) { fprintf(stderr, "cleared: %s\n", tiny_last_error()); return 1; }
    if (!same("cleared")) return 2;
    printf("instance bounds ok: %d of %lld odd instances changed by their own box\n", changed, (long long)(Bn / 2));
    tiny_free(s);
    return 0;
}
