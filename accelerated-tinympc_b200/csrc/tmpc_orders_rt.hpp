// Host side of the run-time-shape kernel (tmpc_kernel_rt.cuh): which floating-point evaluation order the reference's
// build gives each of the 8 products on the ADMM path for an arbitrary (nx, nu, N, scalar), emitted as postfix
// reduction programs.
//
// The reference computes with fixed-size Eigen 3.4.90 expressions (/root/reference/src/tinympc/admm.cpp:19,20,31,35,83),
// so the order is decided at compile time by Eigen's dispatch, for an "-O3" SSE2 build (packet = 16 bytes):
//   * lazyProduct of a column-major lhs with a column: the destination column is assigned packet-wise where it can be
//     (ProductEvaluators.h etor_product_packet_impl: acc = e_k + acc, sequential over k) and coefficient-wise elsewhere
//     (coeff() = (lhs.row(r).transpose().cwiseProduct(rhs)).sum(): Redux.h, completely unrolled half-split tree up to the
//     unrolling limit, a plain loop beyond it);
//   * which rows are "packet rows": all of them when rows % packet == 0 (inner-vectorised traversal); otherwise the
//     linear-vectorised traversal (AssignEvaluator.h): rows [0, rows/packet*packet) when the assignment is completely
//     unrolled, and rows [first, first + (rows-first)/packet*packet) with `first` = distance of the destination COLUMN'S
//     ADDRESS to the next 16-byte boundary when it is not (dense_assignment_loop<LinearVectorizedTraversal, NoUnrolling>);
//     the column address follows from the member layout of TinyWorkspace (types.hpp:52-97) held 16-byte aligned, as the
//     reference's examples and generated code hold it (globals);
//   * inner products over a contiguous vector (Kinf^T r, Xref^T Pinf, Bdyn^T p below the GEMV threshold, a 1-row Kinf):
//     vectorised redux (Redux.h redux_vec_unroller / the NoUnrolling two-accumulator loop beyond the limit);
//   * dimensions >= 8 (EIGEN_CACHEFRIENDLY_PRODUCT_THRESHOLD) of the regular products of admm.cpp:19: GEMV kernels
//     (GeneralMatrixVector.h): column-major = sequential, row-major = per-lane sequential + predux + scalar tail.
// oracle/tinympc_oracle.c restates the same rule independently for the tests, and oracle/pin_shapes.py pins that
// restatement to the compiled reference on 60+ shapes x {float, double}; the GPU tests then pin this file to the oracle.
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <vector>

namespace tmpc_rt {

enum Order { SEQ = 0, VECREDUX = 1, TREE = 2, VECLOOP = 3, GEMV_ROW = 5 };

// expression tree of one reduction over K individually rounded products
struct Expr {
    struct Node { int leaf, l, r, depth; };
    std::vector<Node> n;
    int leaf(int k) { n.push_back({k, -1, -1, 1}); return (int)n.size() - 1; }
    int add(int a, int b)
    {
        // Sethi-Ullman number: operand-stack need when the deeper side is evaluated first
        const int da = n[a].depth, db = n[b].depth;
        n.push_back({-1, a, b, da == db ? da + 1 : std::max(da, db)});
        return (int)n.size() - 1;
    }
    int tree(int s, int len) { return len == 1 ? leaf(s) : add(tree(s, len / 2), tree(s + len / 2, len - len / 2)); }
    int ptree(int s, int len, int l, int pk)
    {
        return len == 1 ? leaf(s * pk + l) : add(ptree(s, len / 2, l, pk), ptree(s + len / 2, len - len / 2, l, pk));
    }
    int predux(const int *lane, int pk) { return pk == 4 ? add(add(lane[0], lane[2]), add(lane[1], lane[3])) : add(lane[0], lane[1]); }
};

inline int build(Expr &x, int order, int K, int pk)
{
    int lane[4];
    switch (order) {
    case SEQ: {
        int acc = x.leaf(0);
        for (int k = 1; k < K; ++k) acc = x.add(acc, x.leaf(k));
        return acc;
    }
    case TREE:
        return x.tree(0, K);
    case VECREDUX: {
        const int np = K / pk;
        if (np == 0) return x.tree(0, K);
        for (int l = 0; l < pk; ++l) lane[l] = x.ptree(0, np, l, pk);
        int res = x.predux(lane, pk);
        if (np * pk != K) res = x.add(res, x.tree(np * pk, K - np * pk));
        return res;
    }
    case VECLOOP: {
        const int a2 = (K / (2 * pk)) * (2 * pk), a1 = (K / pk) * pk;
        if (!a1) {
            int acc = x.leaf(0);
            for (int k = 1; k < K; ++k) acc = x.add(acc, x.leaf(k));
            return acc;
        }
        int p0[4], p1[4];
        for (int l = 0; l < pk; ++l) p0[l] = x.leaf(l);
        if (a1 > pk) {
            for (int l = 0; l < pk; ++l) p1[l] = x.leaf(pk + l);
            for (int i = 2 * pk; i < a2; i += 2 * pk)
                for (int l = 0; l < pk; ++l) {
                    p0[l] = x.add(p0[l], x.leaf(i + l));
                    p1[l] = x.add(p1[l], x.leaf(i + pk + l));
                }
            for (int l = 0; l < pk; ++l) p0[l] = x.add(p0[l], p1[l]);
            if (a1 > a2)
                for (int l = 0; l < pk; ++l) p0[l] = x.add(p0[l], x.leaf(a2 + l));
        }
        int res = x.predux(p0, pk);
        for (int i = a1; i < K; ++i) res = x.add(res, x.leaf(i));
        return res;
    }
    case GEMV_ROW: {
        const int full = (K / pk) * pk;
        if (!full) {   // 0 + e_0 + e_1 ...: value-identical to the sequential sum
            int acc = x.leaf(0);
            for (int k = 1; k < K; ++k) acc = x.add(acc, x.leaf(k));
            return acc;
        }
        for (int l = 0; l < pk; ++l) {
            lane[l] = x.leaf(l);
            for (int j = pk; j < full; j += pk) lane[l] = x.add(lane[l], x.leaf(j + l));
        }
        int res = x.predux(lane, pk);
        for (int j = full; j < K; ++j) res = x.add(res, x.leaf(j));
        return res;
    }
    }
    return -1;
}

// postfix program: one entry per term, leaf | (adds that follow << 8).  SEQ / TREE are emitted left operand first, which
// keeps their terms in natural order (the kernel then walks the vector instead of gathering); the vectorised orders visit
// the terms out of order anyway and are emitted deeper operand first (a + b == b + a bit for bit on non-NaN data).
// Returns the operand-stack depth the program needs.
inline int emit(const Expr &x, int root, bool inorder, std::vector<uint16_t> &out)
{
    const size_t start = out.size();
    std::function<void(int)> go = [&](int i) {
        const Expr::Node &nd = x.n[i];
        if (nd.leaf >= 0) { out.push_back((uint16_t)nd.leaf); return; }
        const bool lfirst = inorder || x.n[nd.l].depth >= x.n[nd.r].depth;
        go(lfirst ? nd.l : nd.r);
        go(lfirst ? nd.r : nd.l);
        out.back() = (uint16_t)(out.back() + 0x100);
    };
    go(root);
    int sp = 0, mx = 0;
    for (size_t k = start; k < out.size(); ++k) {
        ++sp;
        mx = std::max(mx, sp);
        sp -= out[k] >> 8;
    }
    return mx;
}

// K_FIXED: TREE / VECREDUX / GEMV_ROW run through the kernel's compile-time-K code (one unrolled instance per K <= 64);
// the program is then unused and the coefficient copy stays in natural order.  The interpreter (K_INORDER / K_GATHER)
// remains for VECLOOP (K beyond Eigen's unrolling limit: double, K > 55).
enum Kind { K_SEQ = 0, K_INORDER = 1, K_GATHER = 2, K_FIXED = 3 };

struct Variant {
    int prog;   // offset into Orders::prog (K entries)
    int kind;   // K_SEQ: acc = e_k + acc, no program needed; K_INORDER: terms 0..K-1 in order; K_GATHER: by leaf index
    int order;
};
struct Prod { Variant a, b; };

struct Orders {
    Prod Kx, Ax, Bu, Btp, Qs, Mp, Ktr, XtP;
    int head_Kx, head_Ax, head_Qs, head_Mp;
    int rt_u, rt_x, rt_p, off_u, off_x, off_p, pk, sb;
    std::vector<uint16_t> prog;
    int max_depth = 0;
};

// fast = FAST policy: every product is one sequential FMA chain, no order to reproduce
inline Orders build_orders(int nx, int nu, int N, int sb, bool fast = false)
{
    Orders o{};
    const int pk = 16 / sb, large = 8;
    const int unroll_k = (110 * pk + 1) / 4;   // Redux.h: cost 4K-1 <= EIGEN_UNROLLING_LIMIT * pk
    const int tail_x = nx <= unroll_k ? TREE : SEQ, tail_u = nu <= unroll_k ? TREE : SEQ;
    const int vred_x = nx <= unroll_k ? VECREDUX : VECLOOP, vred_u = nu <= unroll_k ? VECREDUX : VECLOOP;
    auto add_prog = [&](int order, int K) {
        if (fast) order = SEQ;
        Expr x;
        const int root = build(x, order, K, pk);
        Variant v;
        v.prog = (int)o.prog.size();
        v.order = order;
        const bool inorder = order == SEQ || order == TREE;
        o.max_depth = std::max(o.max_depth, emit(x, root, inorder, o.prog));
        bool ident = true, seq = true;
        for (int k = 0; k < K; ++k) {
            const uint16_t w = o.prog[v.prog + k];
            ident = ident && (w & 0xff) == k;
            seq = seq && (w >> 8) == (k ? 1 : 0);
        }
        v.kind = (ident && seq) ? K_SEQ : ident ? K_INORDER : K_GATHER;
        if (v.kind != K_SEQ && (order == TREE || order == VECREDUX || (order == GEMV_ROW && K >= pk))) v.kind = K_FIXED;
        return v;
    };
    auto two = [&](int oa, int ob, int K) {
        Prod p;
        p.a = add_prog(oa, K);
        p.b = (ob == oa || fast) ? p.a : add_prog(ob, K);
        return p;
    };
    o.Kx = two(nu == 1 ? vred_x : SEQ, tail_x, nx);
    o.head_Kx = nu == 1 ? -1 : (nu / pk) * pk;
    o.Ax = two(SEQ, tail_x, nx);
    // a 1-row Bdyn is stored row-major (like a 1-row Kinf): its single coefficient is a vectorised redux over the row
    o.Bu = nx == 1 ? two(vred_u, vred_u, nu) : two(SEQ, tail_u, nu);
    o.head_Ax = nx == 1 ? -1 : (nx / pk) * pk;
    o.Btp = two((nu >= large && nx >= large) ? GEMV_ROW : vred_x, vred_x, nx);
    o.Qs = two(SEQ, tail_u, nu);
    o.head_Qs = nu >= large ? -1 : (nu / pk) * pk;
    o.Mp = two(nu == 1 ? SEQ : tail_x, tail_x, nx);
    o.head_Mp = nu == 1 ? (nx / pk) * pk : -1;
    o.Ktr = two(vred_u, vred_u, nu);
    o.XtP = two(vred_x, vred_x, nx);
    o.pk = pk;
    o.sb = sb;
    const int lim = 110 * pk;   // AssignEvaluator.h: size * (dst + src coefficient cost) <= EIGEN_UNROLLING_LIMIT * pk
    o.rt_u = (nu != 1 && nu % pk != 0 && nu * (1 + 4 * nx + 2) > lim);
    o.rt_x = (nx % pk != 0 && nx * (1 + 4 * nx + 4 * nu - 1) > lim);
    o.rt_p = (nu == 1 && nx % pk != 0 && nx * (1 + 1 + (4 * nx - 1) + 1 + 3 + 1) > lim);
    // TinyWorkspace members x u q r p d in declaration order; a member is 16-byte aligned iff its size is a multiple of 16
    const int bx = nx * N * sb, bu = nu * (N - 1) * sb;
    const int sizes[6] = {bx, bu, bx, bu, bx, bu};
    int offs[6], off = 0;
    for (int k = 0; k < 6; ++k) {
        const int al = (sizes[k] % 16 == 0) ? 16 : sb;
        off = (off + al - 1) / al * al;
        offs[k] = off;
        off += sizes[k];
    }
    o.off_x = offs[0];
    o.off_u = offs[1];
    o.off_p = offs[4];
    return o;
}

}  // namespace tmpc_rt
