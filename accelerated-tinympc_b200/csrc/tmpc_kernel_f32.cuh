// fp32 production kernel for sm_100a: same algorithm and bit-exact results as tmpc_kernel.cuh, re-engineered
// around what the B200 SM actually offers (numbers: profiles/r01_*):
//
//  * 2 warps per scheduler instead of 1.  The per-instance state {d,y,z,g,v,p_N} is 1440 B, so shared memory
//    alone holds 128-160 instances per SM = one warp per SMSP, and every LDS / constant-load latency is exposed
//    (v1: 50 % issue utilisation, stall = short scoreboard).  Here g and v (2/3 of the state) live in TENSOR
//    MEMORY, used as a per-thread scratchpad through tcgen05.st / tcgen05.ld (.32x32b: thread t of a warp owns
//    TMEM lane 32*(warp%4)+t; 240 of the 256 columns of its half), d/y/z/p_N stay in shared memory: 256
//    instances per SM.  Measured TMEM scratch bandwidth: 351 B/clk/SM load, 256 B/clk/SM store (tools/ubench_tmem.cu).
//  * packed fp32: accumulations and element-wise work run as FADD2 / FFMA2 on pairs of output rows, which
//    halves the issue slots of the FMA pipe's work (the pipe itself still does 128 lane-ops/clk/SM).  In PARITY
//    mode the PRODUCTS stay scalar FMUL: ptxas contracts mul.f32x2 + add.f32x2 into FFMA2 even with -fmad=false.
//  * stacked coefficient matrices [K;A] and [B^T;AmBKt], stored so that the coefficients of consecutive output rows
//    for one input column are adjacent in the constant bank: one LDCU.128 feeds 4 rows (8 FP instructions).
#pragma once
#include "tmpc_kernel.cuh"

#ifndef TMPC_REFILL_PAIRS
#define TMPC_REFILL_PAIRS 1
#endif

namespace tmpc {

template <int NX, int NU, int NH> struct alignas(16) ModelF32 {
    static constexpr int RS = NU + NX;  // stacked rows
    float KA[NX * RS];   // column k: [Kinf(:,k) ; Adyn(:,k)]
    float Bc[NU * NX];   // column k: Bdyn(:,k)
    float BM[NX * RS];   // column k: [Bdyn(k,:)^T ; AmBKt(:,k)]     (B^T p and AmBKt p share p_k)
    float Qi[NU * NU];   // column k: Quu_inv(:,k)
    float Kr[NU * NX];   // column k: Kinf(k,:)^T                     (Kinf^T r)
    float Pt[NX * NX];   // column k: Pinf(k,:)^T                     (Xref^T Pinf)
    float Qd[NX];
    float xmin[NH * NX], xmax[NH * NX];
    float umin[(NH - 1) * NU], umax[(NH - 1) * NU];
    float rho, nrho, pri_tol, dua_tol;
    int max_iter, check_term;
    float2 nz2;          // (-0.f, -0.f), opaque to the compiler: addend of the exactly-rounded packed product (prod2)
};

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, neg2(b)); }
// Exactly rounded PAIR of products in one FMA-pipe instruction.  sm_100 has no packed multiply that survives ptxas:
// mul.rn.f32x2 followed by add.rn.f32x2 is contracted into FFMA2 (tools/ubench_fp32.cu), which would break the
// reference's "round the product, then round the sum".  fma(a, b, -0) rounds the exact product once and adding -0
// changes neither value nor sign, so with Z = (-0, -0) read from the kernel parameters (unknown to the compiler, hence
// neither folded nor contracted into the following add) FFMA2 R, R.F32, UR.F32x2, Z IS the packed multiply.
__device__ __forceinline__ float2 prod2(float2 a, float2 b, float2 Z) { return __ffma2_rn(a, b, Z); }
// Where the packed product is used (tuning switches).  Measured on B200, 1,048,576 hover instances, PARITY, quadrotor
// pattern (gpurun_out/variants_v6.log): DENSE/M/ELT = 1/1/1 12.69 ms, 0/0/0 11.40, 0/0/1 11.36, 0/1/0 11.28, 0/1/1 11.37.
// FFMA2 occupies the FMA pipe for two cycles like the two FMULs it replaces, and its coefficient PAIR has to come
// through LDC.64 into registers where FMUL takes the coefficient straight from the constant bank (FMUL R,R,c[][] /
// UR), so the packed product does not pay for the dense mat-vecs; it does for the paired AmBKt rows.
#ifndef TMPC_SPEC_FACTOR
#define TMPC_SPEC_FACTOR 2.0f    // speculative emission when every residual is within this factor of its tolerance (measured 1.25: 11.278 ms, 2: 11.216, 4: 11.227, always: 11.309; the remaining 6.6 % of extra lane-trips at 1M instances are the tail: exhausted lanes waiting for their warp)
#endif
#ifndef TMPC_PROD2_DENSE
#define TMPC_PROD2_DENSE 0   // dense row-pair mat-vecs
#endif
#ifndef TMPC_PROD2_M
#define TMPC_PROD2_M 1       // AmBKt rows of a structural pattern evaluated as pairs (else row by row, scalar)
#endif
#ifndef TMPC_PROD2_ELT
#define TMPC_PROD2_ELT 0     // element-wise products of update_linear_cost
#endif
__device__ __forceinline__ float2 prodd(float2 c, float x, float2 Z)
{
    if constexpr (TMPC_PROD2_DENSE) return prod2(c, f2(x, x), Z);
    else return f2(__fmul_rn(c.x, x), __fmul_rn(c.y, x));
}
__device__ __forceinline__ float2 prode(float2 a, float2 b, float2 Z)
{
    if constexpr (TMPC_PROD2_ELT) return prod2(a, b, Z);
    else return f2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
}

// ---- TMEM scratch (per-thread columns): tm_ld8 / tm_ld4 / tm_st8 / tm_st4 / tm_wait_st are in tmpc_kernel.cuh ----
// The loaded registers are only valid after tcgen05.wait::ld; passing them through the wait as in/out operands
// gives the compiler the data dependence (it must not schedule a use above the wait).
template <int N> __device__ __forceinline__ void tm_wait_ld(float (&r)[N])
{
    static_assert(N == 12 || N == 24, "wait helper sized for one or two 12-vectors");
    if constexpr (N == 12) {
        asm volatile("tcgen05.wait::ld.sync.aligned;"
                     : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]),
                       "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]) :: "memory");
    } else {
        asm volatile("tcgen05.wait::ld.sync.aligned;"
                     : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]),
                       "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]), "+f"(r[14]),
                       "+f"(r[15]), "+f"(r[16]), "+f"(r[17]), "+f"(r[18]), "+f"(r[19]), "+f"(r[20]), "+f"(r[21]),
                       "+f"(r[22]), "+f"(r[23]) :: "memory");
    }
}

// g and v for one thread: NX*NH columns each.  TM = tensor memory, else shared memory (same interface).
template <int NX, int NH, int BLOCK, bool TM> struct XStore;
template <int NX, int NH, int BLOCK> struct XStore<NX, NH, BLOCK, true> {
    static_assert(NX == 12, "TMEM path is laid out for 12-vectors: stage i = 24 adjacent columns [g_i | v_i] (x16 + x8)");
    uint32_t base;
    __device__ __forceinline__ XStore(uint32_t tmem_base, int warp)
    {
        base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
    }
    // issue loads of g_i and v_i into gv[0..11], gv[12..23]; valid after wait()
    __device__ __forceinline__ void load_issue(int i, float (&gv)[24]) const
    {
        tm_ld16(base + i * 24, gv); tm_ld8(base + i * 24 + 16, gv + 16);
    }
    __device__ __forceinline__ void wait(float (&gv)[24]) const { tm_wait_ld<24>(gv); }
    __device__ __forceinline__ void store(int i, const float (&g)[12], const float (&v)[12]) const
    {
        float t[24];
#pragma unroll
        for (int j = 0; j < 12; ++j) { t[j] = g[j]; t[12 + j] = v[j]; }
        tm_st16(base + i * 24, t); tm_st8(base + i * 24 + 16, t + 16);
    }
    __device__ __forceinline__ void fence_st() const { tm_wait_st(); }
};
template <int NX, int NH, int BLOCK> struct XStore<NX, NH, BLOCK, false> {
    SVec<float, NX, NH, BLOCK> sg, sv;
    static constexpr size_t BYTES = 2 * SVec<float, NX, NH, BLOCK>::BYTES;
    __device__ __forceinline__ XStore(unsigned char *p, int tid) : sg(p, tid), sv(p + SVec<float, NX, NH, BLOCK>::BYTES, tid) {}
    __device__ __forceinline__ void load_issue(int i, float (&gv)[2 * NX]) const
    {
        float g[NX], v[NX];
        sg.load(i, g); sv.load(i, v);
#pragma unroll
        for (int j = 0; j < NX; ++j) { gv[j] = g[j]; gv[NX + j] = v[j]; }
    }
    __device__ __forceinline__ void wait(float (&)[2 * NX]) const {}
    __device__ __forceinline__ void store(int i, const float (&g)[NX], const float (&v)[NX]) const { sg.store(i, g); sv.store(i, v); }
    __device__ __forceinline__ void fence_st() const {}
};

// ---- packed row-pair reductions -----------------------------------------------------------------------
// e(k) returns the float2 of products for one row pair at input index k.
template <int S, int LEN, class E> __device__ __forceinline__ float2 tree2(const E &e)
{
    if constexpr (LEN == 1) return e(S);
    else {
        constexpr int H = LEN / 2;
        float2 a = tree2<S, H>(e);
        float2 b = tree2<S + H, LEN - H>(e);
        return add2(a, b);
    }
}
template <int S, int LEN, int L, class E> __device__ __forceinline__ float2 ptree2(const E &e)
{
    if constexpr (LEN == 1) return e(S * 4 + L);
    else {
        constexpr int H = LEN / 2;
        float2 a = ptree2<S, H, L>(e);
        float2 b = ptree2<S + H, LEN - H, L>(e);
        return add2(a, b);
    }
}
template <int ORD, int K, class E> __device__ __forceinline__ float2 reduce2(const E &e)
{
    if constexpr (ORD == ORD_SEQ) {
        float2 acc = e(0);
#pragma unroll
        for (int k = 1; k < K; ++k) acc = add2(e(k), acc);
        return acc;
    } else if constexpr (ORD == ORD_TREE) {
        return tree2<0, K>(e);
    } else {
        static_assert(ORD == ORD_VECREDUX && K % 4 == 0, "packed path supports K % 4 == 0 vector reductions");
        float2 l0 = ptree2<0, K / 4, 0>(e), l1 = ptree2<0, K / 4, 1>(e), l2 = ptree2<0, K / 4, 2>(e), l3 = ptree2<0, K / 4, 3>(e);
        return add2(add2(l0, l2), add2(l1, l3));
    }
}

// out2[j] (row pair j of R rows) = sum_k c[k*RS + R0 + 2j .. +1] * x[k], rows R0..R0+R-1 of a stacked matrix
template <int ORD, int R, int K, int RS, int R0, bool FAST>
__device__ __forceinline__ void matvec2(const float *c, const float (&x)[K], float2 (&out)[R / 2], const float2 Z)
{
    static_assert(R % 2 == 0, "row pairs");
    if constexpr (FAST) {
#pragma unroll
        for (int j = 0; j < R / 2; ++j) out[j] = __fmul2_rn(f2(c[R0 + 2 * j], c[R0 + 2 * j + 1]), f2(x[0], x[0]));
#pragma unroll
        for (int k = 1; k < K; ++k)
#pragma unroll
            for (int j = 0; j < R / 2; ++j)
                out[j] = __ffma2_rn(f2(c[k * RS + R0 + 2 * j], c[k * RS + R0 + 2 * j + 1]), f2(x[k], x[k]), out[j]);
    } else if constexpr (ORD == ORD_SEQ) {
        // column sweep: all rows advance together, one coefficient column (R adjacent values) per step
#pragma unroll
        for (int j = 0; j < R / 2; ++j) out[j] = prodd(f2(c[R0 + 2 * j], c[R0 + 2 * j + 1]), x[0], Z);
#pragma unroll
        for (int k = 1; k < K; ++k)
#pragma unroll
            for (int j = 0; j < R / 2; ++j)
                out[j] = add2(prodd(f2(c[k * RS + R0 + 2 * j], c[k * RS + R0 + 2 * j + 1]), x[k], Z), out[j]);
    } else {
        // tree orders need all K products of a row before the first add: go 2 row pairs (4 rows = one LDCU.128 per k) at a time
#pragma unroll
        for (int j0 = 0; j0 < R / 2; j0 += 2) {
#pragma unroll
            for (int jj = 0; jj < 2 && j0 + jj < R / 2; ++jj) {
                const int j = j0 + jj;
                auto e = [&](int k) { return prodd(f2(c[k * RS + R0 + 2 * j], c[k * RS + R0 + 2 * j + 1]), x[k], Z); };
                out[j] = reduce2<ORD, K>(e);
            }
        }
    }
}


// ---- structural sparsity of the model, known at compile time -------------------------------------------------
// A term whose coefficient is exactly zero contributes +-0 to its sum, and adding +-0 is the identity on every
// non-zero partial sum, so dropping it leaves every VALUE of the reference computation unchanged (only the sign of
// an exactly-zero result can differ; finite data, as everywhere on this path).  A coefficient that is exactly 1
// needs no multiply.  The pattern is a compile-time property of the kernel instance; tmpc_set_model() checks the
// actual matrices against it and falls back to the dense instance when they do not conform.
//   a_nz[r] bit k : Adyn(r,k) may be non-zero      a_one[r] bit k : Adyn(r,k) == 1 exactly
//   m_nz[r] bit k : AmBKt(r,k) may be non-zero
template <int NX> struct PatDense {
    static constexpr bool sparse = false;
    static constexpr int id = 0;
};
// Linearised quadrotor of the reference's examples (problem_data/quadrotor_*hz_params.hpp): Adyn = I + 14 couplings;
// the shipped 7-decimal AmBKt has the z / vz rows and columns decoupled.
struct PatQuadrotor {
    static constexpr bool sparse = true;
    static constexpr int id = 1;
    static constexpr uint32_t a_nz[12] = {0x451, 0x28a, 0x104, 0x208, 0x410, 0x820, 0x450, 0x288, 0x100, 0x200, 0x400, 0x800};
    static constexpr uint32_t a_one[12] = {0x001, 0x002, 0x004, 0x008, 0x010, 0x020, 0x040, 0x080, 0x100, 0x200, 0x400, 0x800};
    static constexpr uint32_t m_nz[12] = {0xefb, 0xefb, 0x104, 0xefb, 0xefb, 0xefb, 0xefb, 0xefb, 0x104, 0xefb, 0xefb, 0xefb};
    // AmBKt rows are evaluated as PAIRS (one FFMA2 + one FADD2 per term for two rows): m_perm lists the rows in pair
    // order (the host stores the AmBKt part of the stacked image in this row order), rows of a pair share a mask
    static constexpr int m_perm[12] = {0, 1, 3, 9, 4, 5, 6, 7, 10, 11, 2, 8};
    static constexpr uint32_t m_pair_nz[6] = {0xefb, 0xefb, 0xefb, 0xefb, 0xefb, 0x104};
};

__host__ __device__ constexpr bool bits_any(uint32_t m, int s, int len) { return ((m >> s) & ((len >= 32) ? 0xffffffffu : ((1u << len) - 1u))) != 0; }

// sequential order (acc = e_0; acc = e_k + acc) over the non-zero terms of one row
template <uint32_t NZ, uint32_t ONE, int K, int KEND, bool STARTED, bool FAST, class C, class X>
__device__ __forceinline__ float sp_seq(const C &c, const X &x, float acc)
{
    if constexpr (K == KEND) {
        return STARTED ? acc : 0.f;
    } else if constexpr (((NZ >> K) & 1u) == 0u) {
        return sp_seq<NZ, ONE, K + 1, KEND, STARTED, FAST>(c, x, acc);
    } else {
        constexpr bool one = ((ONE >> K) & 1u) != 0u;
        if constexpr (!STARTED) acc = one ? x(K) : __fmul_rn(c(K), x(K));
        else if constexpr (one) acc = __fadd_rn(x(K), acc);
        else if constexpr (FAST) acc = __fmaf_rn(c(K), x(K), acc);
        else acc = __fadd_rn(__fmul_rn(c(K), x(K)), acc);
        return sp_seq<NZ, ONE, K + 1, KEND, true, FAST>(c, x, acc);
    }
}
// scalar tree order (recursive half split) over the non-zero terms of one row; precondition: some bit of NZ in [S, S+LEN)
template <uint32_t NZ, int S, int LEN, class E> __device__ __forceinline__ float sp_tree(const E &e)
{
    if constexpr (LEN == 1) {
        return e(S);
    } else {
        constexpr int H = LEN / 2;
        constexpr bool la = bits_any(NZ, S, H), lb = bits_any(NZ, S + H, LEN - H);
        if constexpr (la && lb) return __fadd_rn(sp_tree<NZ, S, H>(e), sp_tree<NZ, S + H, LEN - H>(e));
        else if constexpr (la) return sp_tree<NZ, S, H>(e);
        else return sp_tree<NZ, S + H, LEN - H>(e);
    }
}
// row R of a sparse mat-vec: c(k) coefficient, x(k) input
template <int ORD, uint32_t NZ, uint32_t ONE, int K, bool FAST, class C, class X>
__device__ __forceinline__ float sp_row(const C &c, const X &x)
{
    if constexpr (!bits_any(NZ, 0, K)) return 0.f;
    else if constexpr (FAST || ORD == ORD_SEQ) return sp_seq<NZ, ONE, 0, K, false, FAST>(c, x, 0.f);
    else {
        static_assert(ORD == ORD_TREE, "sparse rows: sequential or scalar-tree order");
        return sp_tree<NZ, 0, K>([&](int k) { return __fmul_rn(c(k), x(k)); });
    }
}
// all rows of a sparse mat-vec, resolved at compile time
template <int ORD, class MASKS, int R, int NR, int K, bool FAST, class CF, class X>
__device__ __forceinline__ void sp_rows(const CF &cf, const X &x, float (&out)[NR])
{
    if constexpr (R < NR) {
        out[R] = sp_row<ORD, MASKS::nz(R), MASKS::one(R), K, FAST>([&](int k) { return cf(R, k); }, x);
        sp_rows<ORD, MASKS, R + 1, NR, K, FAST>(cf, x, out);
    }
}
// the same for a PAIR of rows that share a mask: e(k) = float2 of exactly rounded products
template <uint32_t NZ, int S, int LEN, class E> __device__ __forceinline__ float2 sp_tree2(const E &e)
{
    if constexpr (LEN == 1) {
        return e(S);
    } else {
        constexpr int H = LEN / 2;
        constexpr bool la = bits_any(NZ, S, H), lb = bits_any(NZ, S + H, LEN - H);
        if constexpr (la && lb) return add2(sp_tree2<NZ, S, H>(e), sp_tree2<NZ, S + H, LEN - H>(e));
        else if constexpr (la) return sp_tree2<NZ, S, H>(e);
        else return sp_tree2<NZ, S + H, LEN - H>(e);
    }
}
template <uint32_t NZ, int K, int KEND, bool STARTED, bool FAST, class C2, class X>
__device__ __forceinline__ float2 sp_seq2(const C2 &c2, const X &x, float2 acc, const float2 Z)
{
    if constexpr (K == KEND) {
        return STARTED ? acc : f2(0.f, 0.f);
    } else if constexpr (((NZ >> K) & 1u) == 0u) {
        return sp_seq2<NZ, K + 1, KEND, STARTED, FAST>(c2, x, acc, Z);
    } else {
        if constexpr (!STARTED) acc = FAST ? __fmul2_rn(c2(K), f2(x(K), x(K))) : prod2(c2(K), f2(x(K), x(K)), Z);
        else if constexpr (FAST) acc = __ffma2_rn(c2(K), f2(x(K), x(K)), acc);
        else acc = add2(prod2(c2(K), f2(x(K), x(K)), Z), acc);
        return sp_seq2<NZ, K + 1, KEND, true, FAST>(c2, x, acc, Z);
    }
}
// AmBKt p over row pairs in PAT::m_perm order; out[] in NATURAL row order.  c2(j, k) = coefficient pair of pair j
template <int ORD, class PAT, int J, int NX, bool FAST, class CF2, class X>
__device__ __forceinline__ void sp_pairs_m(const CF2 &cf2, const X &x, float (&out)[NX], const float2 Z)
{
    if constexpr (J < NX / 2) {
        constexpr uint32_t NZ = PAT::m_pair_nz[J];
        float2 r;
        if constexpr (!bits_any(NZ, 0, NX)) r = f2(0.f, 0.f);
        else if constexpr (FAST || ORD == ORD_SEQ)
            r = sp_seq2<NZ, 0, NX, false, FAST>([&](int k) { return cf2(J, k); }, x, f2(0.f, 0.f), Z);
        else {
            static_assert(ORD == ORD_TREE, "sparse row pairs: sequential or scalar-tree order");
            r = sp_tree2<NZ, 0, NX>([&](int k) { return prod2(cf2(J, k), f2(x(k), x(k)), Z); });
        }
        out[PAT::m_perm[2 * J]] = r.x;
        out[PAT::m_perm[2 * J + 1]] = r.y;
        sp_pairs_m<ORD, PAT, J + 1, NX, FAST>(cf2, x, out, Z);
    }
}
template <class PAT> struct MaskA {
    __host__ __device__ static constexpr uint32_t nz(int r) { return PAT::a_nz[r]; }
    __host__ __device__ static constexpr uint32_t one(int r) { return PAT::a_one[r]; }
};
template <class PAT, int R, int NX> __device__ __forceinline__ void unpermute_m(const float (&in)[NX], float (&out)[NX])
{
    if constexpr (R < NX) {
        out[PAT::m_perm[R]] = in[R];
        unpermute_m<PAT, R + 1, NX>(in, out);
    }
}
template <class PAT> struct MaskMP {   // AmBKt rows in the stored (pair) order
    __host__ __device__ static constexpr uint32_t nz(int r) { return PAT::m_nz[PAT::m_perm[r]]; }
    __host__ __device__ static constexpr uint32_t one(int) { return 0u; }
};
template <class PAT> struct MaskM {
    __host__ __device__ static constexpr uint32_t nz(int r) { return PAT::m_nz[r]; }
    __host__ __device__ static constexpr uint32_t one(int) { return 0u; }
};

// This lane's rows of the per-lane coalesced scratch (SolveArgs::scratch): chunk c of the lane is one float4; the 32 lanes of
// a warp hold chunk c in 512 contiguous bytes, so every access below is a fully coalesced LDG.128 / STG.128 that stays in
// L2 (ld/st.global.cg: the lines are private to the lane, L1 would only be thrashed).
struct LaneScratch {
    float4 *base;
    __device__ __forceinline__ LaneScratch(void *scratch, int chunks, int block_threads)
    {
        const long long gwarp = (long long)blockIdx.x * (block_threads >> 5) + (threadIdx.x >> 5);
        base = reinterpret_cast<float4 *>(scratch) + (gwarp * chunks) * 32 + (threadIdx.x & 31);
    }
    __device__ __forceinline__ float4 ld(int c) const { return __ldcg(base + c * 32); }
    __device__ __forceinline__ void st(int c, float4 v) const { __stcg(base + c * 32, v); }
    template <int D> __device__ __forceinline__ void ldv(int c, float (&o)[D]) const
    {
#pragma unroll
        for (int k = 0; k < D / 4; ++k) {
            const float4 t = ld(c + k);
            o[4 * k] = t.x; o[4 * k + 1] = t.y; o[4 * k + 2] = t.z; o[4 * k + 3] = t.w;
        }
    }
    template <int D> __device__ __forceinline__ void stv(int c, const float (&o)[D]) const
    {
#pragma unroll
        for (int k = 0; k < D / 4; ++k) st(c + k, make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]));
    }
    // the same through a row pointer the caller steps itself (chunk offsets become immediates of the stores)
    __device__ __forceinline__ float4 *row(int c) const { return base + (long long)c * 32; }
    template <int D> __device__ static __forceinline__ void stp(float4 *p, const float (&o)[D])
    {
#pragma unroll
        for (int k = 0; k < D / 4; ++k) __stcg(p + k * 32, make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]));
    }
};
// chunk maps (NX = 12, NU = 4): per-instance box of stage i = 8 chunks [xmin(3) xmax(3) umin(1) umax(1)]; warm mirror of
// stage i = 5 chunks [v(3) z(1) d(1)] (last stage: v only)
template <int NX, int NU, int NH> struct ScratchMap {
    static constexpr int CX = NX / 4, CU = NU / 4;
    static constexpr int IB_STAGE = 2 * CX + 2 * CU, IB_CHUNKS = IB_STAGE * NH;
    static constexpr int WM_STAGE = CX + 2 * CU, WM_CHUNKS = WM_STAGE * (NH - 1) + CX;
};
#ifndef TMPC_MIRROR_FACTOR
#define TMPC_MIRROR_FACTOR 4.0f   // warm start: a backward sweep mirrors d / v / z when the next iteration may converge, i.e. every
                                  // residual is within this factor of its tolerance (see the kernel's backward section)
#endif

// Pre-pass for batches with per-instance reference trajectories: the p_N seed -(Xref_{N-1}^T Pinf) (admm.cpp:83) of every instance,
// one thread each, same evaluation order as the kernel's own seed_pn.
template <int NX, int NU, int NH, bool FAST>
__global__ void pn_seed_kernel(const __grid_constant__ ModelF32<NX, NU, NH> P, const float *__restrict__ Xref, long long stride, long long batch,
                               float *__restrict__ out)
{
    using O = Orders<float, NX, NU>;
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= batch) return;
    float xr[NX], pn[NX];
    gload<float, NX>(Xref + b * stride + (NH - 1) * NX, xr);
#pragma unroll
    for (int j = 0; j < NX; ++j)
        pn[j] = -dot<float, O::XtP, NX, FAST>([&](int k) { return P.Pt[k * NX + j]; }, [&](int k) { return xr[k]; });
    gstore<float, NX>(out + b * NX, pn);
}

template <int NX, int NU, int NH, int BLOCK, bool TM> struct SmemLayoutF32 {
    using SU = SVec<float, NU, NH - 1, BLOCK>;
    using SP = SVec<float, NX, 1, BLOCK>;
    static constexpr size_t XBYTES = TM ? 0 : 2 * SVec<float, NX, NH, BLOCK>::BYTES;
    static constexpr size_t BYTES = 3 * SU::BYTES + SP::BYTES + XBYTES + 16;
    // CSM instances: + the model image (128-byte aligned) and its mbarrier
    static constexpr size_t BYTES_CSM = (BYTES + 127) / 128 * 128 + sizeof(ModelF32<NX, NU, NH>) + 16;
    // ROLL instances (TM): + two x rows per lane, the step's measurement and the plant's next state, and the control applied
    static constexpr size_t BYTES_ROLL = BYTES + 2 * SP::BYTES + SVec<float, NU, 1, BLOCK>::BYTES;
};

// CB: the bounds are the same at every horizon stage (the usual box constraints, e.g. every example of the reference):
// the kernel reads stage 0's row with compile-time addresses (operands straight from the constant bank) instead of
// indexing the per-stage table with the loop counter (LDC.64 with a register index: 2.7 % of the instructions).
// IB: every instance has its own box (tmpc_set_instance_bounds): the bound operands of the projection come from the lane's
// scratch rows (filled at refill from the ctx's [instance][stage][dim] copy) instead of the constant bank.
// CSM: the A/B the project brief asks for -- "the shared cache is staged into shared memory once per CTA via TMA".  The model
// image (3.9 KB: stacked Kinf/Adyn/Bdyn/AmBKt/Quu_inv/Pinf, Q, bounds, tolerances) is fetched from a device copy by ONE
// cp.async.bulk per CTA, completing on an mbarrier, and every coefficient is then an LDS (uniform address: a broadcast)
// instead of a constant-bank operand folded into the FMUL.  Measured against the default (profiles/r02_cache_smem_ab.md):
// kept as TMPC_KERNEL=f32_tma_cache, not the default.
template <int NX, int NU, int NH, int BLOCK, bool FAST, bool WARM, bool TM, class PAT = PatDense<NX>, bool CB = false, bool IB = false,
          bool CSM = false, bool ROLL = false>
__global__ void __launch_bounds__(BLOCK, 1)
admm_kernel_f32(const __grid_constant__ ModelF32<NX, NU, NH> Pc, const __grid_constant__ SolveArgs<float> a)
{
    static_assert(!(IB && CB), "per-instance bounds are never constant over the batch");
    extern __shared__ __align__(16) unsigned char smem[];
    const ModelF32<NX, NU, NH> *Pp = &Pc;
    if constexpr (CSM) {
        using M = ModelF32<NX, NU, NH>;
        static_assert(sizeof(M) % 16 == 0, "bulk copies move multiples of 16 bytes");
        constexpr size_t OFF = (SmemLayoutF32<NX, NU, NH, BLOCK, TM>::BYTES + 127) / 128 * 128;
        M *sm = reinterpret_cast<M *>(smem + OFF);
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(smem + OFF + sizeof(M));
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((uint32_t)sizeof(M)) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dst), "l"(a.model_g), "r"((uint32_t)sizeof(M)), "r"(bar) : "memory");
        }
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar) : "memory");
        }
        Pp = sm;
    }
    const ModelF32<NX, NU, NH> &P = *Pp;
    using SM = ScratchMap<NX, NU, NH>;
    const LaneScratch sc(a.scratch, a.sc_chunks, BLOCK);

    static_assert(NX % 4 == 0 && NU % 4 == 0, "packed kernel: 16-byte vectors of x and u");
    using O = Orders<float, NX, NU>;
    using L = SmemLayoutF32<NX, NU, NH, BLOCK, TM>;
    constexpr int RS = NU + NX;
    const float2 Z = P.nz2;   // (-0, -0): see prod2
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const unsigned lane = tid & 31;
    constexpr unsigned FULLM = 0xffffffffu;
    constexpr int XROW = NX * NH, UROW = NU * (NH - 1);

    unsigned char *sp = smem;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sp); sp += 16;
    typename L::SU sd(sp, tid); sp += L::SU::BYTES;
    typename L::SU sy(sp, tid); sp += L::SU::BYTES;
    typename L::SU sz(sp, tid); sp += L::SU::BYTES;
    typename L::SP spn(sp, tid); sp += L::SP::BYTES;
    // ROLL (TM instances: nothing else lives past spn): x_1 of the latest forward sweep, and the step's x0 -- registers are what keeps
    // the coefficient pairs of the backward sweep out of the loop (profiles/r02_ncu_fused_loop.md)
    typename L::SP sx1(sp, tid);
    typename L::SP sx0(sp + L::SP::BYTES, tid);
    SVec<float, NU, 1, BLOCK> su0(sp + 2 * L::SP::BYTES, tid);   // u(:,0) of the latest forward sweep

    uint32_t tmem_base = 0;
    if constexpr (TM) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;"
                         :: "r"((uint32_t)__cvta_generic_to_shared(tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_base = *tmem_slot;
    }
    auto make_xs = [&]() {
        if constexpr (TM) return XStore<NX, NH, BLOCK, true>(tmem_base, warp);
        else return XStore<NX, NH, BLOCK, false>(sp, tid);
    };
    auto xs = make_xs();

    long long inst = -1;
    int it = 0;
    int phase = PH_FREE;
    bool exhausted = false;
    bool deferred = false;   // warp-uniform: the previous trip postponed a single-lane refill
    long long redo = -1;       // WARM: instance to solve AGAIN from its untouched warm input, mirror forced on (see the backward section)
    bool force_mirror = false; // WARM: this run mirrors d / v / z in every backward sweep
    bool mirrored = false;     // WARM: the previous iteration's backward sweep mirrored d / v / z into the lane's scratch rows
    int flush = 0;             // WARM: the instance `finst` converged in the previous trip and its warm state is still on chip / in
    long long finst = -1;      // the scratch rows: 1 = write g, y back; 2 = also d, v, z of the iteration before (from the scratch)
    bool spec = false;   // speculative emission: this trip's x,u go straight to the output because the lane is
                         // expected to terminate in it (residuals within SPEC_FACTOR of tolerance, or last iteration)
    float x0[NX];
    float res[4] = {0.f, 0.f, 0.f, 0.f};
    using cnt_t = typename std::conditional<ROLL, unsigned, unsigned long long>::type;   // (per lane: far below 2^32 either way)
    cnt_t n_iter = 0, n_solved = 0, n_trips = 0, n_inst = 0;
#pragma unroll
    for (int j = 0; j < NX; ++j) x0[j] = 0.f;
    // ROLL -- the examples' closed loop (quadrotor_hovering.cpp:90-114) for this lane's instance WITHOUT leaving the lane: when step
    // rs < S-1 ends, the next step starts from the state on chip (x0 <- x_1 of the solve = the plant step, y = g = 0, d / v / z as the
    // reference carries them) instead of a write-back, a re-launch and a refill; the last step is a plain warm solve.  The warm
    // state the reference leaves after an early exit is one iteration behind (see the backward section): it comes from the mirror
    // rows, of which a ROLL launch has two sets: `area` receives this step's mirrors, area ^ 1 holds the state the step started from
    // (the previous step's last mirror), which is also what a step is solved again from when its mirror was not predicted.
    // LAZY first iteration.  After an early exit the chip holds v, z of the terminating iteration; the reference's (one behind) differ
    // from them in ONE place of the next step: the dual residuals of its first iteration (admm.cpp:96,98), i.e. only in whether that
    // iteration may terminate.  So the next step starts on the on-chip v, z -- no reload.  Its first dual residuals are then measured
    // against v, z that are off by exactly what the previous step's LAST dual residuals measured (dprev, each below the tolerance):
    // when one of them exceeds (dprev + tolerance) by a margin, or a primal residual fails, the exact test fails as well (triangle
    // inequality; the 0.1 % margin covers the roundings) and the step simply goes on.  Otherwise -- the first iteration MIGHT end the
    // step -- the step is taken again from the exact state in the mirror rows (pend 3).  Every result stays bit-exact; in the hover
    // closed loop the first dual residual of a step is 10-100 tolerances, so the second path is for nearly stationary instances.
    static_assert(!ROLL || (WARM && TM && !IB && !CSM), "fused closed loop: warm, tensor-memory instance, shared box");
    const int S = ROLL ? a.roll_steps : 1;
    int rs = 0;           // MPC step of `inst` being solved
    int pend = 0;         // what the top of the next trip does: 1 next step after an early exit, 2 after a max_iter exit, 3 same step again
    int area = 0, varea = 0, farea = 0;
    bool gzero = false;   // this trip's forward sweep reads g as zero (reset duals; g lives in tensor memory, written collectively)
    bool vprev = false;   // exact hand-over: before this trip's forward sweep v is replaced by the mirror rows `varea` (one iteration behind)
    bool lazy = false;    // this step's first iteration ran on the on-chip v, z (see above)
    float dprev_x = 0.f, dprev_u = 0.f;   // rho-scaled dual residuals of the iteration that ended the previous step
    auto wm = [&](int ar) -> int { return a.sc_wm + (ROLL ? ar * SM::WM_CHUNKS : 0); };

    // p_N seed: -(Xref_{N-1}^T Pinf)  (admm.cpp:83).  It depends on the reference trajectory only, so with one Xref shared by the
    // batch it is computed once per lane here instead of at every refill: the refill section runs at warp level nearly every
    // trip (32 lanes x 1/34 terminations per trip at the headline workload) and the seed was 60 % of its instructions.
    auto seed_pn = [&](const float *xl) {
        float xr[NX], pn[NX];
        gload<float, NX>(xl, xr);
#pragma unroll
        for (int j = 0; j < NX; ++j)
            pn[j] = -dot<float, O::XtP, NX, FAST>([&](int k) { return P.Pt[k * NX + j]; }, [&](int k) { return xr[k]; });
        spn.store(0, pn);
    };
    const bool shared_xref = (a.xref_stride == 0);
    // ROLL with a reference table: the window of MPC step `step` of this lane's instance (tmpc.h tmpc_batch_set_xref_table)
    const bool tab = ROLL && a.roll_table != nullptr;
    auto window = [&](int step) -> const float * {
        long long w0 = (long long)(a.roll_start ? __ldg(a.roll_start + (inst < 0 ? 0 : inst)) : 0) + a.roll_step0 + step;
        if (w0 > a.roll_rows - NH) w0 = a.roll_rows - NH;
        return a.roll_table + w0 * NX;
    };
    // Controls-only callers (x = u = NULL, u0 given: what an MPC loop applies, quadrotor_hovering.cpp:110): u(:,0) is stored by
    // every trip's stage 0, so the trip in which the lane terminates has already delivered it and neither an emission trip
    // nor a speculative one is needed -- warm starts included (their state is written back by the flush at the next trip).
    const bool u0only = !a.x && !a.u;
    const bool duals_zero = WARM && (a.test_flags & 4);   // the caller reset y and g: zero-fill instead of reading them
    if (shared_xref) seed_pn(a.Xref + (NH - 1) * NX);

    for (;;) {
        if constexpr (ROLL) {
            // ------------------------------------------------------------------ next MPC step (or the same one again) on this lane
            if (pend) {
                // where the state the coming run starts from (d, v, z) is: this step's last mirror after an early exit past the first
                // iteration or a max_iter exit (whose backward sweep mirrored what it left on chip); otherwise the step's own start state
                const int src = (pend == 3 || (pend == 1 && it == 1)) ? (area ^ 1) : area;
                // exact hand-over (reload z, v read from the mirror rows in the sweep): a step taken again, and after a step that ended
                // at its first iteration (the next one probably does too); otherwise lazy
                const bool exact = pend == 3 || (pend == 1 && (it == 1 || (a.test_flags & 8)));
                lazy = (pend == 1) && !exact;
                dprev_x = res[1]; dprev_u = res[3];
                if (exact) {
                    // z (and for a repeated step d) of the reference's state: the forward sweeps have overwritten them on chip
                    float4 tz[NH - 1], td[NH - 1];
#pragma unroll
                    for (int i = 0; i < NH - 1; ++i) {
                        tz[i] = sc.ld(wm(src) + SM::WM_STAGE * i + SM::CX);
                        if (pend == 3) td[i] = sc.ld(wm(src) + SM::WM_STAGE * i + SM::CX + SM::CU);
                    }
#pragma unroll
                    for (int i = 0; i < NH - 1; ++i) {
                        const float t[NU] = {tz[i].x, tz[i].y, tz[i].z, tz[i].w};
                        sz.store(i, t);
                        if (pend == 3) { const float t2[NU] = {td[i].x, td[i].y, td[i].z, td[i].w}; sd.store(i, t2); }
                    }
                    vprev = true; varea = src;
                }
                {   // y = 0 (hovering.cpp:100); g = 0 is applied by the next forward sweep
                    float zu[NU];
#pragma unroll
                    for (int j = 0; j < NU; ++j) zu[j] = 0.f;
#pragma unroll 1
                    for (int i = 0; i < NH - 1; ++i) sy.store(i, zu);
                }
                gzero = true;
                if (pend != 3) {
                    // plant step (hovering.cpp:108): x1 = Adyn x0 + Bdyn u0 is stage 1 of the solve's last forward sweep, which left it in
                    // the lane's shared-memory row; the instance's x0 row follows (the measurement the LAST step starts from is what the
                    // caller's plant step needs after the launch)
                    sx1.load(0, x0);
                    sx0.store(0, x0);
                    gstore<float, NX>(const_cast<float *>(a.x0) + inst * NX, x0);
                    if (a.roll_x) gstore<float, NX>(a.roll_x + ((long long)rs * a.batch + inst) * NX, x0);
                    if (a.roll_u0) { float t[NU]; su0.load(0, t); gstore<float, NU>(a.roll_u0 + ((long long)rs * a.batch + inst) * NU, t); }
                    rs += 1;
                    if (tab) seed_pn(window(rs) + (NH - 1) * NX);   // the window moves on: p_N seed of the new step
                    area = src ^ 1;
                    force_mirror = (a.test_flags & 2) != 0;
                } else {
                    force_mirror = true;
                }
                it = 0; mirrored = false; pend = 0;
                res[0] = res[1] = res[2] = res[3] = 0.f;
                spec = (rs >= S - 1) && (P.max_iter <= 1);
            }
            // exact hand-over: v of the reference's state into tensor memory (collective: the other lanes write back what they read).
            // Off the common path -- a lazy step never gets here
            if (__any_sync(FULLM, vprev)) {
#pragma unroll 1
                for (int i = 0; i < NH; ++i) {
                    float gv[2 * NX], vsub[NX];
                    xs.load_issue(i, gv);
                    if (vprev) sc.ldv<NX>(wm(varea) + SM::WM_STAGE * i, vsub);
                    xs.wait(gv);
#pragma unroll
                    for (int j = 0; j < NX; ++j)
                        if (vprev) gv[NX + j] = vsub[j];
                    xs.store(i, *reinterpret_cast<float(*)[NX]>(gv), *reinterpret_cast<float(*)[NX]>(gv + NX));
                }
                xs.fence_st();
                vprev = false;
            }
        }
        // ------------------------------------------------------------------ lane refill (warp-uniform branch)
        const bool need = (phase == PH_FREE) && !exhausted;
        unsigned m = __ballot_sync(FULLM, need);
        if constexpr (TMPC_REFILL_PAIRS) {
            // The refill section runs for the whole warp; with one free lane it is deferred by one trip so that it usually
            // serves two (another lane frees up with probability ~0.9 per trip at the headline workload)
            const bool others_busy = __ballot_sync(FULLM, phase != PH_FREE) != 0;
            if (m && __popc(m) < 2 && !deferred && others_busy) { deferred = true; m = 0; }
            else deferred = false;
        }
        // WARM: a lane that converged in the previous trip writes its warm state back HERE, before this trip's emission sweep or a
        // refill touches g / y (the section is warp-level anyway: tcgen05.ld/st are collective)
        unsigned mf = 0;
        if constexpr (WARM) mf = __ballot_sync(FULLM, flush != 0);
        if (m | mf) {
            const bool refill_now = need && ((m >> lane) & 1u);   // (a deferred single lane waits for the next trip)
            const unsigned mc = __ballot_sync(FULLM, refill_now && redo < 0);   // lanes that claim a NEW instance (the others run theirs again)
            const int leader = mc ? __ffs(mc) - 1 : 0;
            unsigned long long base = 0;
            if (mc && (int)lane == leader) base = atomicAdd(a.counter, (unsigned long long)__popc(mc));
            base = __shfl_sync(FULLM, base, leader);
            bool fill = false;
            if constexpr (WARM) {
                if (flush) {
                    // y of the terminating iteration (still in shared memory); d, v, z of the iteration before it from the scratch
                    // rows, fetched in two batches so that their L2 latency is paid twice, not once per row
                    const long long fi = ROLL ? inst : finst;   // (a ROLL lane flushes before its refill below replaces inst)
                    float *wyo = a.wy + fi * UROW;
#pragma unroll
                    for (int i = 0; i < NH - 1; ++i) { float t[NU]; sy.load(i, t); gstore<float, NU>(wyo + i * NU, t); }
                    if (flush == 2) {
                        float *wvo = a.wv + fi * XROW, *wzo = a.wz + fi * UROW, *wdo = a.wd + fi * UROW;
                        float4 tv[NH * SM::CX];
#pragma unroll
                        for (int i = 0; i < NH; ++i)
#pragma unroll
                            for (int c = 0; c < SM::CX; ++c) tv[i * SM::CX + c] = sc.ld(wm(farea) + SM::WM_STAGE * i + c);
#pragma unroll
                        for (int i = 0; i < NH; ++i)
#pragma unroll
                            for (int c = 0; c < SM::CX; ++c) reinterpret_cast<float4 *>(wvo + i * NX)[c] = tv[i * SM::CX + c];
                        float4 tz[NH - 1], td[NH - 1];
#pragma unroll
                        for (int i = 0; i < NH - 1; ++i) { tz[i] = sc.ld(wm(farea) + SM::WM_STAGE * i + SM::CX); td[i] = sc.ld(wm(farea) + SM::WM_STAGE * i + SM::CX + SM::CU); }
#pragma unroll
                        for (int i = 0; i < NH - 1; ++i) { reinterpret_cast<float4 *>(wzo)[i] = tz[i]; reinterpret_cast<float4 *>(wdo)[i] = td[i]; }
                    }
                }
            }
            if (refill_now) {
                long long ni = redo;
                if (redo < 0) {
                    const long long idx = (long long)base + __popc(mc & ((1u << lane) - 1u));
                    if (idx < a.batch) ni = claim_instance(a, idx);
                }
                if (ni >= 0) {
                    inst = ni; phase = PH_RUN; it = 0; fill = true;
                    force_mirror = (redo >= 0) || (a.test_flags & 2); redo = -1; mirrored = false;
                    spec = (P.max_iter <= 1) && !u0only && !(ROLL && S > 1);
                    if constexpr (ROLL) { rs = 0; pend = 0; area = 0; gzero = false; vprev = false; lazy = false; }
                    res[0] = res[1] = res[2] = res[3] = 0.f;
                    gload<float, NX>(a.x0 + inst * NX, x0);
                    if constexpr (ROLL) sx0.store(0, x0);
                    if (tab) seed_pn(window(0) + (NH - 1) * NX);
                    else if (!shared_xref) {
                        // (the refill section is warp-level code: with the seeds from the pre-pass it issues three loads instead of a 12 x 12 product)
                        if (a.pn_seed) { float pn[NX]; gload<float, NX>(a.pn_seed + inst * NX, pn); spn.store(0, pn); }
                        else seed_pn(a.Xref + inst * a.xref_stride + (NH - 1) * NX);
                    }
                    if constexpr (IB) {
                        // the instance's own box -> the lane's scratch rows (a missing / disabled family = +-inf).  Loads are issued
                        // half a horizon at a time, so the copy costs a handful of memory round trips instead of one per stage
                        const float inf = __int_as_float(0x7f800000);
                        constexpr int HB = (NH + 1) / 2;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float4 lo[HB * SM::CX], hi[HB * SM::CX];
#pragma unroll
                            for (int q = 0; q < HB; ++q) {
                                const int i = h * HB + q;
#pragma unroll
                                for (int c = 0; c < SM::CX; ++c) {
                                    if (i < NH && a.ixmin) {
                                        lo[q * SM::CX + c] = __ldg(reinterpret_cast<const float4 *>(a.ixmin + inst * XROW + i * NX) + c);
                                        hi[q * SM::CX + c] = __ldg(reinterpret_cast<const float4 *>(a.ixmax + inst * XROW + i * NX) + c);
                                    } else {
                                        lo[q * SM::CX + c] = make_float4(-inf, -inf, -inf, -inf);
                                        hi[q * SM::CX + c] = make_float4(inf, inf, inf, inf);
                                    }
                                }
                            }
#pragma unroll
                            for (int q = 0; q < HB; ++q) {
                                const int i = h * HB + q;
                                if (i < NH) {
#pragma unroll
                                    for (int c = 0; c < SM::CX; ++c) {
                                        sc.st(a.sc_ib + SM::IB_STAGE * i + c, lo[q * SM::CX + c]);
                                        sc.st(a.sc_ib + SM::IB_STAGE * i + SM::CX + c, hi[q * SM::CX + c]);
                                    }
                                }
                            }
                        }
                        {
                            static_assert(SM::CU == 1, "one 16-byte chunk per input-bound row");
                            float4 ul[NH - 1], uh[NH - 1];
#pragma unroll
                            for (int i = 0; i < NH - 1; ++i) {
                                if (a.iumin) {
                                    ul[i] = __ldg(reinterpret_cast<const float4 *>(a.iumin + inst * UROW + i * NU));
                                    uh[i] = __ldg(reinterpret_cast<const float4 *>(a.iumax + inst * UROW + i * NU));
                                } else { ul[i] = make_float4(-inf, -inf, -inf, -inf); uh[i] = make_float4(inf, inf, inf, inf); }
                            }
#pragma unroll
                            for (int i = 0; i < NH - 1; ++i) {
                                sc.st(a.sc_ib + SM::IB_STAGE * i + 2 * SM::CX, ul[i]);
                                sc.st(a.sc_ib + SM::IB_STAGE * i + 2 * SM::CX + SM::CU, uh[i]);
                            }
                        }
                    }
                    if (WARM && a.wd) {
                        // warm d, y, z: every load in flight before the first store (one memory round trip, not one per stage)
                        float td[NH - 1][NU], ty[NH - 1][NU], tz[NH - 1][NU];
#pragma unroll
                        for (int i = 0; i < NH - 1; ++i) {
                            gload<float, NU>(a.wd + inst * UROW + i * NU, td[i]);
                            if (duals_zero) {
#pragma unroll
                                for (int j = 0; j < NU; ++j) ty[i][j] = 0.f;
                            } else gload<float, NU>(a.wy + inst * UROW + i * NU, ty[i]);
                            gload<float, NU>(a.wz + inst * UROW + i * NU, tz[i]);
                        }
#pragma unroll
                        for (int i = 0; i < NH - 1; ++i) { sd.store(i, td[i]); sy.store(i, ty[i]); sz.store(i, tz[i]); }
                        if constexpr (ROLL) {   // the state step 0 starts from -> mirror rows 1
#pragma unroll
                            for (int i = 0; i < NH - 1; ++i) {
                                sc.stv<NU>(wm(1) + SM::WM_STAGE * i + SM::CX + SM::CU, td[i]);
                                sc.stv<NU>(wm(1) + SM::WM_STAGE * i + SM::CX, tz[i]);
                            }
                        }
                    } else {
                        float zu[NU];
#pragma unroll
                        for (int j = 0; j < NU; ++j) zu[j] = 0.f;
#pragma unroll 1
                        for (int i = 0; i < NH - 1; ++i) { sd.store(i, zu); sy.store(i, zu); sz.store(i, zu); }
                    }
                } else {
                    exhausted = true;
                }
            }
            // g, v: every lane of the warp takes part (tcgen05 is warp-collective); lanes that are not being refilled write back
            // what they hold.  RB stages per round trip: the section is bound by the latency of the tensor-memory reads and -- with a
            // warm start -- of the global loads of the new instance's g / v rows, not by instructions.  A lane that is flushing a
            // converged instance stores the g it reads before anything overwrites it.
            const bool wfill = WARM && a.wd;
            constexpr int RB = WARM ? 5 : 2;
            float *wgo = (WARM && flush) ? a.wg + (ROLL ? inst : finst) * XROW : nullptr;
            auto refill_batch = [&](int i0, auto nb_tag) {
                constexpr int NB = decltype(nb_tag)::value;
                float gv[NB][2 * NX];
#pragma unroll
                for (int q = 0; q < NB; ++q) xs.load_issue(i0 + q, gv[q]);
#pragma unroll
                for (int q = 0; q < NB; ++q) xs.wait(gv[q]);
                if constexpr (WARM) {
                    if (wgo) {
#pragma unroll
                        for (int q = 0; q < NB; ++q) gstore<float, NX>(wgo + (i0 + q) * NX, *reinterpret_cast<float(*)[NX]>(gv[q]));
                    }
                }
                if (fill) {
                    if (wfill) {
#pragma unroll
                        for (int q = 0; q < NB; ++q) {
                            if (duals_zero) {
#pragma unroll
                                for (int j = 0; j < NX; ++j) gv[q][j] = 0.f;
                            } else gload<float, NX>(a.wg + inst * XROW + (i0 + q) * NX, *reinterpret_cast<float(*)[NX]>(gv[q]));
                            gload<float, NX>(a.wv + inst * XROW + (i0 + q) * NX, *reinterpret_cast<float(*)[NX]>(gv[q] + NX));
                            if constexpr (ROLL) sc.stv<NX>(wm(1) + SM::WM_STAGE * (i0 + q), *reinterpret_cast<float(*)[NX]>(gv[q] + NX));
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < NB; ++q)
#pragma unroll
                            for (int j = 0; j < 2 * NX; ++j) gv[q][j] = 0.f;
                    }
                }
#pragma unroll
                for (int q = 0; q < NB; ++q) xs.store(i0 + q, *reinterpret_cast<float(*)[NX]>(gv[q]), *reinterpret_cast<float(*)[NX]>(gv[q] + NX));
            };
            if (m) {   // (a flush alone rewrites nothing)
#pragma unroll 1
                for (int i = 0; i + RB <= NH; i += RB) refill_batch(i, std::integral_constant<int, RB>());
#pragma unroll 1
                for (int i = NH / RB * RB; i < NH; ++i) refill_batch(i, std::integral_constant<int, 1>());
                xs.fence_st();
            } else if (WARM && mf) {
                // flush without a refill in this warp: g still has to be read out of tensor memory (collectively) and stored
#pragma unroll 1
                for (int i = 0; i + RB <= NH; i += RB) {
                    float gv[RB][2 * NX];
#pragma unroll
                    for (int q = 0; q < RB; ++q) xs.load_issue(i + q, gv[q]);
#pragma unroll
                    for (int q = 0; q < RB; ++q) xs.wait(gv[q]);
                    if (wgo) {
#pragma unroll
                        for (int q = 0; q < RB; ++q) gstore<float, NX>(wgo + (i + q) * NX, *reinterpret_cast<float(*)[NX]>(gv[q]));
                    }
                }
#pragma unroll 1
                for (int i = NH / RB * RB; i < NH; ++i) {
                    float gv[2 * NX];
                    xs.load_issue(i, gv);
                    xs.wait(gv);
                    if (wgo) gstore<float, NX>(wgo + i * NX, *reinterpret_cast<float(*)[NX]>(gv));
                }
            }
            if constexpr (WARM) flush = 0;
        }
        if (__all_sync(FULLM, phase == PH_FREE)) break;
        ++n_trips;

        const bool emit = (phase == PH_EMIT);
        if (phase == PH_RUN) ++it;

        // ------------------------------------------------------------------ forward sweep
        // forward_pass (admm.cpp:27-37) + update_slack (:45-61) + update_dual (:67-71) + residual maxima (:95-98)
        float pri_x = 0.f, dua_x = 0.f, pri_u = 0.f, dua_u = 0.f;
        {
            float x[NX];
            if constexpr (ROLL) sx0.load(0, x);
            else {
#pragma unroll
                for (int j = 0; j < NX; ++j) x[j] = x0[j];
            }
            const bool wr = emit || (spec && phase == PH_RUN);
            float *xo = (wr && a.x) ? a.x + inst * XROW : nullptr;
            float *uo = (wr && a.u) ? a.u + inst * UROW : nullptr;
            const bool stash = ROLL && rs < S - 1 && phase == PH_RUN;   // a step the lane continues from: u(:,0) and x_1 kept on chip

            auto xpart = [&](int i, float (&gv)[2 * NX], const float (&bxl)[NX], const float (&bxh)[NX]) {
                // state slack / dual / residuals for stage i (uses x_i)
                xs.wait(gv);
                if constexpr (ROLL) {   // first sweep of a step on this lane: g = 0
#pragma unroll
                    for (int j = 0; j < NX; ++j)
                        if (gzero) gv[j] = 0.f;
                }
                float g[NX], vn[NX];
#pragma unroll
                for (int j = 0; j < NX; j += 2) {
                    const float2 x2 = f2(x[j], x[j + 1]), g2 = f2(gv[j], gv[j + 1]), v2 = f2(gv[NX + j], gv[NX + j + 1]);
                    float2 t = add2(x2, g2);                                                         // :48
                    if constexpr (IB) {
                        t.x = fminf(bxh[j], fmaxf(bxl[j], t.x));                                   // :59, the instance's own box
                        t.y = fminf(bxh[j + 1], fmaxf(bxl[j + 1], t.y));
                    } else {
                        const int bi = CB ? 0 : i * NX;
                        t.x = fminf(P.xmax[bi + j], fmaxf(P.xmin[bi + j], t.x));                   // :59
                        t.y = fminf(P.xmax[bi + j + 1], fmaxf(P.xmin[bi + j + 1], t.y));
                    }
                    const float2 rp = sub2(x2, t), rd = sub2(v2, t);
                    pri_x = fmaxf(pri_x, fmaxf(fabsf(rp.x), fabsf(rp.y)));                           // :95
                    dua_x = fmaxf(dua_x, fmaxf(fabsf(rd.x), fabsf(rd.y)));                           // :96
                    const float2 gn = sub2(add2(g2, x2), t);                                         // :70
                    g[j] = gn.x; g[j + 1] = gn.y; vn[j] = t.x; vn[j + 1] = t.y;
                }
                xs.store(i, g, vn);
                if (xo) gstore<float, NX>(xo + i * NX, x);
            };

#pragma unroll 1
            for (int i = 0; i < NH - 1; ++i) {
                float gv[2 * NX];
                xs.load_issue(i, gv);
                // the instance's own box for this stage: 8 coalesced 16-byte loads from the lane's scratch rows, in flight
                // behind the mat-vecs below
                float bxl[NX], bxh[NX], bul[NU], buh[NU];
                if constexpr (IB) {
                    sc.ldv<NU>(a.sc_ib + SM::IB_STAGE * i + 2 * SM::CX, bul);
                    sc.ldv<NU>(a.sc_ib + SM::IB_STAGE * i + 2 * SM::CX + SM::CU, buh);
                    sc.ldv<NX>(a.sc_ib + SM::IB_STAGE * i, bxl);
                    sc.ldv<NX>(a.sc_ib + SM::IB_STAGE * i + SM::CX, bxh);
                }
                float d[NU], y[NU], z[NU];
                sd.load(i, d); sy.load(i, y); sz.load(i, z);
                // [K;A] x_i in one column sweep
                float2 ka[RS / 2];
                if constexpr (PAT::sparse) {
                    // Kinf x_i dense (packed row pairs), Adyn x_i over its non-zero terms only
                    float2 kk[NU / 2];
                    float aa[NX];
                    matvec2<O::Kx, NU, NX, RS, 0, FAST>(P.KA, x, kk, Z);
                    sp_rows<O::Ax, MaskA<PAT>, 0, NX, NX, FAST>([&](int r, int k) { return P.KA[k * RS + NU + r]; },
                                                                  [&](int k) { return x[k]; }, aa);
#pragma unroll
                    for (int j = 0; j < NU / 2; ++j) ka[j] = kk[j];
#pragma unroll
                    for (int j = 0; j < NX / 2; ++j) ka[NU / 2 + j] = f2(aa[2 * j], aa[2 * j + 1]);
                } else if constexpr (O::Kx == ORD_SEQ && O::Ax == ORD_SEQ) {
                    matvec2<ORD_SEQ, RS, NX, RS, 0, FAST>(P.KA, x, ka, Z);
                } else {
                    float2 kk[NU / 2], aa[NX / 2];
                    matvec2<O::Kx, NU, NX, RS, 0, FAST>(P.KA, x, kk, Z);
                    matvec2<O::Ax, NX, NX, RS, NU, FAST>(P.KA, x, aa, Z);
#pragma unroll
                    for (int j = 0; j < NU / 2; ++j) ka[j] = kk[j];
#pragma unroll
                    for (int j = 0; j < NX / 2; ++j) ka[NU / 2 + j] = aa[j];
                }
                float u[NU], zn[NU];
#pragma unroll
                for (int r = 0; r < NU; r += 2) {
                    const float2 d2 = f2(d[r], d[r + 1]), y2 = f2(y[r], y[r + 1]), z2 = f2(z[r], z[r + 1]);
                    const float2 u2 = sub2(neg2(ka[r / 2]), d2);                                     // :31
                    float2 t = add2(u2, y2);                                                         // :47
                    if constexpr (IB) {
                        t.x = fminf(buh[r], fmaxf(bul[r], t.x));                                   // :53, the instance's own box
                        t.y = fminf(buh[r + 1], fmaxf(bul[r + 1], t.y));
                    } else {
                        const int bu = CB ? 0 : i * NU;
                        t.x = fminf(P.umax[bu + r], fmaxf(P.umin[bu + r], t.x));                   // :53
                        t.y = fminf(P.umax[bu + r + 1], fmaxf(P.umin[bu + r + 1], t.y));
                    }
                    const float2 rp = sub2(u2, t), rd = sub2(z2, t);
                    pri_u = fmaxf(pri_u, fmaxf(fabsf(rp.x), fabsf(rp.y)));                           // :97
                    dua_u = fmaxf(dua_u, fmaxf(fabsf(rd.x), fabsf(rd.y)));                           // :98
                    const float2 yn = sub2(add2(y2, u2), t);                                         // :69
                    u[r] = u2.x; u[r + 1] = u2.y; zn[r] = t.x; zn[r + 1] = t.y; y[r] = yn.x; y[r + 1] = yn.y;
                }
                sy.store(i, y);
                sz.store(i, zn);
                if (uo) gstore<float, NU>(uo + i * NU, u);
                // u(:,0): with the trajectory outputs it is written by the emitting trip; a controls-only solve stores it in EVERY
                // trip (16 bytes per lane, the terminating trip's value is the last one written) -- cheaper than keeping it in
                // four registers across the sweep (measured: 10.21 -> 10.10 ms per 1M-instance launch)
                if constexpr (!ROLL) {
                    if (i == 0 && a.u0 && (u0only ? phase == PH_RUN : wr)) gstore<float, NU>(a.u0 + inst * NU, u);
                } else {
                    if (i == 0 && !stash && a.u0 && (u0only ? phase == PH_RUN : wr)) gstore<float, NU>(a.u0 + inst * NU, u);
                    if (i == 0 && stash) su0.store(0, u);   // every trip: the last one written counts
                }
                // x_{i+1} = A x_i + B u_i                                                            :35
                float2 xn[NX / 2];
                if constexpr (FAST) {
#pragma unroll
                    for (int j = 0; j < NX / 2; ++j) xn[j] = ka[NU / 2 + j];
#pragma unroll
                    for (int k = 0; k < NU; ++k)
#pragma unroll
                        for (int j = 0; j < NX / 2; ++j)
                            xn[j] = __ffma2_rn(f2(P.Bc[k * NX + 2 * j], P.Bc[k * NX + 2 * j + 1]), f2(u[k], u[k]), xn[j]);
                } else {
                    float2 bu[NX / 2];
                    matvec2<O::Bu, NX, NU, NX, 0, false>(P.Bc, u, bu, Z);
#pragma unroll
                    for (int j = 0; j < NX / 2; ++j) xn[j] = add2(ka[NU / 2 + j], bu[j]);
                }
                xpart(i, gv, bxl, bxh);
#pragma unroll
                for (int j = 0; j < NX / 2; ++j) { x[2 * j] = xn[j].x; x[2 * j + 1] = xn[j].y; }
                if constexpr (ROLL) {   // x_1 = Adyn x0 + Bdyn u0: the plant's next state if this trip ends the step
                    if (i == 0 && stash) sx1.store(0, x);
                }
            }
            {
                float gv[2 * NX];
                xs.load_issue(NH - 1, gv);
                float bxl[NX], bxh[NX];
                if constexpr (IB) {
                    sc.ldv<NX>(a.sc_ib + SM::IB_STAGE * (NH - 1), bxl);
                    sc.ldv<NX>(a.sc_ib + SM::IB_STAGE * (NH - 1) + SM::CX, bxh);
                }
                xpart(NH - 1, gv, bxl, bxh);
            }
            xs.fence_st();
            if constexpr (ROLL) gzero = false;
        }

        // ------------------------------------------------------------------ termination (admm.cpp:91-109, :135-138)
        bool final_bwd = false;
        bool chk_now = false;
        bool finished = false;   // every output of this lane's instance is (or will be, after this trip's backward) written
        if (phase == PH_RUN) {
            const bool chk = (it % P.check_term) == 0;
            chk_now = chk;
            if (chk) {
                res[0] = pri_x; res[1] = __fmul_rn(dua_x, P.rho); res[2] = pri_u; res[3] = __fmul_rn(dua_u, P.rho);
            }
            const bool conv = chk && res[0] < P.pri_tol && res[2] < P.pri_tol && res[1] < P.dua_tol && res[3] < P.dua_tol;
            bool rolled = false;
            if constexpr (ROLL) {
                if (lazy && it == 1 && chk && res[0] < P.pri_tol && res[2] < P.pri_tol &&
                    !(res[1] > __fmul_rn(__fadd_rn(dprev_x, P.dua_tol), 1.001f) || res[3] > __fmul_rn(__fadd_rn(dprev_u, P.dua_tol), 1.001f))) {
                    pend = 3; rolled = true;   // the lazy first iteration might end the step: again, from the exact state
                } else if (conv && it > 1 && !mirrored) {
                    pend = 3; rolled = true;   // early exit without the mirror: the step again, from its start state, mirror forced
                } else if (rs < S - 1 && (conv || it >= P.max_iter)) {
                    if (a.roll_iter) a.roll_iter[(long long)rs * a.batch + inst] = it;
                    if (a.roll_status) a.roll_status[(long long)rs * a.batch + inst] = conv ? 1 : 11;
                    n_iter += (unsigned)it; n_solved += conv ? 1u : 0u; ++n_inst;
                    pend = conv ? 1 : 2; rolled = true;   // (2: this trip's backward sweep still runs, admm.cpp:141-144, and mirrors)
                }
            }
            if (rolled) {
            } else if (!ROLL && WARM && a.wd && conv && it > 1 && !mirrored) {
                // Converged, but the previous backward sweep did not mirror d / v / z (the predicate below did not see it coming:
                // not observed in 150,000 hover / closed-loop solves at factor 4, but nothing guarantees it).  Nothing of this
                // instance has touched the caller's warm buffers yet: solve it again from there with the mirror forced on.
                redo = inst; phase = PH_FREE; spec = false;
            } else if (conv || it >= P.max_iter) {
                if (a.iter) a.iter[inst] = it;
                if (a.status) a.status[inst] = conv ? 1 : 11;
                if (a.resid) *reinterpret_cast<float4 *>(a.resid + inst * 4) = make_float4(res[0], res[1], res[2], res[3]);
                n_iter += (unsigned)it; n_solved += conv ? 1u : 0u; ++n_inst;
                final_bwd = !conv;
                if constexpr (!ROLL) {
                    if (WARM && conv && a.wd) { flush = it > 1 ? 2 : 1; finst = inst; }   // written back at the top of the next trip
                } else {
                    // the caller's d, v, z are those of step 0: always written, from the mirror rows that hold the reference's state -- after
                    // a max_iter exit the ones this trip's backward sweep is about to fill (no direct write-back path in these instances)
                    flush = 2; farea = (!conv || it > 1) ? area : (area ^ 1);
                }
                if (u0only) { phase = PH_FREE; finished = true; }   // u(:,0) of this very trip is already in the output
                else if (spec) { phase = PH_FREE; finished = true; }   // x,u of this very trip are already in the output
                else phase = PH_EMIT;
                spec = false;
            } else if (!u0only && !(ROLL && rs < S - 1)) {
                // predict termination in the next trip: last allowed iteration, or every residual within 25 % of its
                // tolerance at a check (ADMM crawls across the threshold); a wrong guess only costs stores
                constexpr float SF = TMPC_SPEC_FACTOR;
                const bool next_chk = ((it + 1) % P.check_term) == 0;
                spec = (it + 1 >= P.max_iter) ||
                       (next_chk && res[0] < SF * P.pri_tol && res[2] < SF * P.pri_tol && res[1] < SF * P.dua_tol && res[3] < SF * P.dua_tol);
            }
        } else if (phase == PH_EMIT) {
            phase = PH_FREE;
            finished = true;
        }

        // ------------------------------------------------------------------ backward sweep
        // update_linear_cost (admm.cpp:77-85) recomputed per stage + backward_pass_grad (:15-22)
        const bool cont = (phase == PH_RUN) && !(ROLL && (pend == 1 || pend == 3));
        // Warm start: the caller's buffers must end up as the reference leaves its workspace (SURVEY 8a note W):
        //   max_iter exit (wfbw): this trip's backward still runs (admm.cpp:141-144): d, v = vnew, z = znew, g, y of this iteration;
        //   early exit  (flush): g, y of this iteration, but d, v, z of the iteration BEFORE (admm.cpp:135-138) -- which the forward
        //                        sweep has already overwritten on chip.  So a continuing lane MIRRORS d, v, z into its coalesced
        //                        scratch rows (wmir) in the backward sweep of every iteration after which the next one may converge:
        //                        all residuals within TMPC_MIRROR_FACTOR of their tolerance (at the reference's own tolerances the
        //                        largest ratio seen one iteration before convergence is 3.0; a miss is caught above and the
        //                        instance is solved again).  That is 3-5 mirrored sweeps per solve instead of 14-34, and the
        //                        caller's buffers are written exactly once, at the top of the trip after the instance ends.
        const bool wfbw = !ROLL && WARM && final_bwd && a.wd;
        const bool rfbw = ROLL && final_bwd;   // fused loop: the last step's max_iter exit mirrors instead
        bool wmir = false;
        if constexpr (WARM) {
            constexpr float MF = TMPC_MIRROR_FACTOR;
            const bool next_chk = ((it + 1) % P.check_term) == 0;
            const bool near = !chk_now || (res[0] < MF * P.pri_tol && res[2] < MF * P.pri_tol && res[1] < MF * P.dua_tol && res[3] < MF * P.dua_tol);
            wmir = (cont || rfbw) && a.wd && (force_mirror || (ROLL && (pend == 2 || rfbw)) || (next_chk && near && !(a.test_flags & 1)));
            if (cont) mirrored = wmir;
        }
        if (__any_sync(FULLM, cont || wfbw || rfbw)) {
            float p[NX];
            const float *xr_base = tab ? window(rs) : a.Xref + (inst < 0 ? 0 : inst) * a.xref_stride;
            float *wdo = wfbw ? a.wd + inst * UROW : nullptr;
            float *wvo = wfbw ? a.wv + inst * XROW : nullptr;
            float *wzo = wfbw ? a.wz + inst * UROW : nullptr;
            float *wgo = wfbw ? a.wg + inst * XROW : nullptr;
            float *wyo = wfbw ? a.wy + inst * UROW : nullptr;
            float4 *mrow = nullptr;   // this sweep's mirror rows, stage by stage (only dereferenced under wmir)
            if constexpr (WARM) mrow = sc.row(wm(area) + SM::WM_STAGE * (NH - 1));
            {
                float gv[2 * NX], pn[NX];
                xs.load_issue(NH - 1, gv);
                spn.load(0, pn);
                xs.wait(gv);
                if constexpr (WARM) {
                    if (wfbw) {
                        gstore<float, NX>(wgo + (NH - 1) * NX, *reinterpret_cast<float(*)[NX]>(gv));
                        gstore<float, NX>(wvo + (NH - 1) * NX, *reinterpret_cast<float(*)[NX]>(gv + NX));
                    } else if (wmir) LaneScratch::stp<NX>(mrow, *reinterpret_cast<float(*)[NX]>(gv + NX));
                }
#pragma unroll
                for (int j = 0; j < NX; ++j) {
                    const float dvg = __fsub_rn(gv[NX + j], gv[j]);
                    if constexpr (FAST) p[j] = __fmaf_rn(P.nrho, dvg, pn[j]);
                    else p[j] = __fsub_rn(pn[j], __fmul_rn(P.rho, dvg));                             // :84
                }
            }
#pragma unroll 1
            for (int i = NH - 2; i >= 0; --i) {
                float gv[2 * NX];
                xs.load_issue(i, gv);
                if constexpr (WARM) mrow -= SM::WM_STAGE * 32;
                float z[NU], y[NU], r[NU], xr[NX];
                sz.load(i, z);
                sy.load(i, y);
                gload<float, NX>(xr_base + i * NX, xr);
#pragma unroll
                for (int j = 0; j < NU; j += 2) {                                                    // :80
                    float2 t = sub2(f2(z[j], z[j + 1]), f2(y[j], y[j + 1]));
                    if constexpr (FAST) t = __fmul2_rn(t, f2(P.nrho, P.nrho));
                    else t = prode(t, f2(P.nrho, P.nrho), Z);
                    r[j] = t.x; r[j + 1] = t.y;
                }
                // [B^T ; AmBKt] p_{i+1}
                float2 bm[RS / 2];
                if constexpr (PAT::sparse) {
                    float2 bb[NU / 2];
                    float mm[NX];
                    matvec2<O::Btp, NU, NX, RS, 0, FAST>(P.BM, p, bb, Z);
                    if constexpr (TMPC_PROD2_M || FAST) {
                        sp_pairs_m<O::Mp, PAT, 0, NX, FAST>(
                            [&](int j, int k) { return f2(P.BM[k * RS + NU + 2 * j], P.BM[k * RS + NU + 2 * j + 1]); },
                            [&](int k) { return p[k]; }, mm, Z);
                    } else {   // row by row (scalar FMUL with the coefficient as a direct constant-bank operand)
                        float mt[NX];
                        sp_rows<O::Mp, MaskMP<PAT>, 0, NX, NX, FAST>([&](int r, int k) { return P.BM[k * RS + NU + r]; },
                                                                       [&](int k) { return p[k]; }, mt);
                        unpermute_m<PAT, 0, NX>(mt, mm);
                    }
#pragma unroll
                    for (int j = 0; j < NU / 2; ++j) bm[j] = bb[j];
#pragma unroll
                    for (int j = 0; j < NX / 2; ++j) bm[NU / 2 + j] = f2(mm[2 * j], mm[2 * j + 1]);
                } else if constexpr (FAST) {
                    matvec2<ORD_SEQ, RS, NX, RS, 0, true>(P.BM, p, bm, Z);
                } else {
                    float2 bb[NU / 2], mm[NX / 2];
                    matvec2<O::Btp, NU, NX, RS, 0, false>(P.BM, p, bb, Z);
                    matvec2<O::Mp, NX, NX, RS, NU, false>(P.BM, p, mm, Z);
#pragma unroll
                    for (int j = 0; j < NU / 2; ++j) bm[j] = bb[j];
#pragma unroll
                    for (int j = 0; j < NX / 2; ++j) bm[NU / 2 + j] = mm[j];
                }
                float s[NU];
#pragma unroll
                for (int j = 0; j < NU; j += 2) {
                    const float2 t = add2(bm[j / 2], f2(r[j], r[j + 1]));
                    s[j] = t.x; s[j + 1] = t.y;
                }
                float2 d2[NU / 2];
                matvec2<O::Qs, NU, NU, NU, 0, FAST>(P.Qi, s, d2, Z);                                    // :19
                float d[NU];
#pragma unroll
                for (int j = 0; j < NU / 2; ++j) { d[2 * j] = d2[j].x; d[2 * j + 1] = d2[j].y; }
                sd.store(i, d, cont);
                if constexpr (WARM) {
                    if (wfbw) gstore<float, NU>(wdo + i * NU, d);
                    else if (wmir) { LaneScratch::stp<NU>(mrow + (SM::CX + SM::CU) * 32, d); LaneScratch::stp<NU>(mrow + SM::CX * 32, z); }
                }
                float2 kr[NX / 2];
                matvec2<O::Ktr, NX, NU, NX, 0, FAST>(P.Kr, r, kr, Z);
                xs.wait(gv);
                if constexpr (WARM) {
                    if (wfbw) {
                        gstore<float, NX>(wgo + i * NX, *reinterpret_cast<float(*)[NX]>(gv));
                        gstore<float, NU>(wyo + i * NU, y);
                        gstore<float, NX>(wvo + i * NX, *reinterpret_cast<float(*)[NX]>(gv + NX));
                        gstore<float, NU>(wzo + i * NU, z);
                    } else if (wmir) {
                        LaneScratch::stp<NX>(mrow, *reinterpret_cast<float(*)[NX]>(gv + NX));
                    }
                }
#pragma unroll
                for (int j = 0; j < NX; j += 2) {
                    const float2 cq = neg2(prode(f2(xr[j], xr[j + 1]), f2(P.Qd[j], P.Qd[j + 1]), Z));              // :81
                    const float2 dvg = sub2(f2(gv[NX + j], gv[NX + j + 1]), f2(gv[j], gv[j + 1]));
                    float2 q;
                    if constexpr (FAST) q = __ffma2_rn(f2(P.nrho, P.nrho), dvg, cq);
                    else q = sub2(cq, prode(dvg, f2(P.rho, P.rho), Z));                                            // :82
                    const float2 pn = sub2(add2(q, bm[NU / 2 + j / 2]), kr[j / 2]);                                // :20
                    p[j] = pn.x; p[j + 1] = pn.y;
                }
            }
        }
        // signal the host-side D2H pipeline: all stores of this instance (incl. warm state of a final backward) are issued
        if (finished && a.done) { __threadfence(); atomicAdd(a.done + (inst >> a.done_shift), 1u); }
    }

    if (a.stats) {
        unsigned long long t_iter = n_iter, t_solved = n_solved, t_trips = n_trips, t_inst = n_inst;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            t_iter += __shfl_down_sync(FULLM, t_iter, o);
            t_solved += __shfl_down_sync(FULLM, t_solved, o);
            t_trips += __shfl_down_sync(FULLM, t_trips, o);
            t_inst += __shfl_down_sync(FULLM, t_inst, o);
        }
        if (lane == 0) {
            atomicAdd(a.stats + 0, t_iter);
            atomicAdd(a.stats + 1, t_solved);
            atomicAdd(a.stats + 2, t_trips);
            atomicAdd(a.stats + 3, t_inst);
        }
    }
    if constexpr (TM) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base) : "memory");
    }
}

}  // namespace tmpc
