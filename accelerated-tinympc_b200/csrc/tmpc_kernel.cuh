// Persistent thread-per-instance ADMM kernel for sm_100a.
//
// One CUDA thread owns one MPC instance at a time and runs the reference's whole tiny_solve loop
// (/root/reference/src/tinympc/admm.cpp:111-152) on it; when the instance terminates the lane emits its
// trajectory and claims the next instance from a global work counter (per-lane early exit + refill), so
// lanes of a warp sit at different ADMM iterations of different instances while executing the same code.
//
//   * shared model/cache (Kinf, Adyn, Bdyn, Quu_inv, AmBKt, Pinf, Q, bounds, rho, tolerances) lives in the
//     kernel-parameter constant bank (__grid_constant__): every mat-vec coefficient is a warp-uniform
//     c[0][imm] operand of FMUL/FFMA and costs no load instruction and no register;
//   * the state that survives an iteration {d, y, z, g, v} (+ p_N seed) sits in shared memory, laid out
//     [16-byte chunk][thread] so a warp's LDS.128/STS.128 is conflict-free; x, u, q, r, p, vnew, znew are
//     transient registers (q, r are recomputed bit-identically in the backward sweep);
//   * forward_pass + update_slack + update_dual + the four residual maxima are fused into one sweep over
//     the horizon; update_linear_cost + backward_pass_grad into the reverse sweep.
//
// Arithmetic policy (template FAST): false = products rounded individually and summed in the reference's
// exact order (orders.h) with FMA contraction disabled -> bit-identical to the "-O3" SSE2 reference build;
// true = FMA chains.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace tmpc {

// ------------------------------------------------------------------------------------------------
// scalar traits: explicitly rounded ops (never contracted by the compiler)
// ------------------------------------------------------------------------------------------------
template <class T> struct Num;
template <> struct Num<float> {
    static constexpr int PK = 4;  // SSE packet width of the reference build for this scalar
    using vec_t = float4;
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float mn(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float mx(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
};
template <> struct Num<double> {
    static constexpr int PK = 2;
    using vec_t = double2;
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ double mn(double a, double b) { return fmin(a, b); }
    static __device__ __forceinline__ double mx(double a, double b) { return fmax(a, b); }
    static __device__ __forceinline__ double abs(double a) { return fabs(a); }
};

// ------------------------------------------------------------------------------------------------
// evaluation orders of the reference build (restated and pinned in oracle/tinympc_oracle.c)
// ------------------------------------------------------------------------------------------------
enum { ORD_SEQ = 0, ORD_VECREDUX = 1, ORD_TREE = 2, ORD_GEMV_ROW = 5 };

template <class T, int NX, int NU> struct Orders {
    static constexpr int PK = Num<T>::PK;
    static constexpr int LARGE = 8;
    static constexpr int Kx = (NU == 1) ? ORD_VECREDUX : (NU % PK == 0 ? ORD_SEQ : ORD_TREE);
    static constexpr int Ax = (NX % PK == 0) ? ORD_SEQ : ORD_TREE;
    static constexpr int Bu = (NX % PK == 0) ? ORD_SEQ : ORD_TREE;
    static constexpr int Btp = (NU >= LARGE && NX >= LARGE) ? ORD_GEMV_ROW : ORD_VECREDUX;
    static constexpr int Qs = (NU >= LARGE || NU % PK == 0 || NU == 1) ? ORD_SEQ : ORD_TREE;
    static constexpr int Mp = (NU == 1 && NX % PK == 0) ? ORD_SEQ : ORD_TREE;
    static constexpr int Ktr = ORD_VECREDUX;
    static constexpr int XtP = ORD_VECREDUX;
};

// e(k) = k-th individually rounded product.  All recursion is resolved at compile time.
template <class T, int S, int LEN, class E> __device__ __forceinline__ T red_tree(const E &e)
{
    if constexpr (LEN == 1) {
        return e(S);
    } else {
        constexpr int H = LEN / 2;
        T a = red_tree<T, S, H>(e);
        T b = red_tree<T, S + H, LEN - H>(e);
        return Num<T>::add(a, b);
    }
}
// tree over packets [S, S+LEN) for SIMD lane L of the reference's packet
template <class T, int S, int LEN, int L, class E> __device__ __forceinline__ T red_ptree(const E &e)
{
    constexpr int PK = Num<T>::PK;
    if constexpr (LEN == 1) {
        return e(S * PK + L);
    } else {
        constexpr int H = LEN / 2;
        T a = red_ptree<T, S, H, L>(e);
        T b = red_ptree<T, S + H, LEN - H, L>(e);
        return Num<T>::add(a, b);
    }
}
template <class T> __device__ __forceinline__ T predux(const T (&l)[Num<T>::PK])
{
    if constexpr (Num<T>::PK == 4) return Num<T>::add(Num<T>::add(l[0], l[2]), Num<T>::add(l[1], l[3]));
    else return Num<T>::add(l[0], l[1]);
}
template <class T, int K, int L, class E> __device__ __forceinline__ T red_lane_seq(const E &e)
{
    constexpr int PK = Num<T>::PK;
    T acc = e(L);
#pragma unroll
    for (int j = PK; j + PK <= K; j += PK) acc = Num<T>::add(e(j + L), acc);
    return acc;
}

// sum of the K individually rounded products e(0..K-1) in the named order of the reference build
template <class T, int ORD, int K, class E>
__device__ __forceinline__ T reduce_products(const E &e)
{
    using N = Num<T>;
    constexpr int PK = N::PK;
    {
        if constexpr (ORD == ORD_SEQ) {
            T acc = e(0);
#pragma unroll
            for (int k = 1; k < K; ++k) acc = N::add(e(k), acc);
            return acc;
        } else if constexpr (ORD == ORD_TREE) {
            return red_tree<T, 0, K>(e);
        } else if constexpr (ORD == ORD_VECREDUX) {
            constexpr int NP = K / PK;
            if constexpr (NP == 0) {
                return red_tree<T, 0, K>(e);
            } else {
                T l[PK];
                l[0] = red_ptree<T, 0, NP, 0>(e);
                l[1] = red_ptree<T, 0, NP, 1>(e);
                if constexpr (PK == 4) {
                    l[2] = red_ptree<T, 0, NP, 2>(e);
                    l[3] = red_ptree<T, 0, NP, 3>(e);
                }
                T r = predux<T>(l);
                if constexpr (NP * PK != K) r = N::add(r, red_tree<T, NP * PK, K - NP * PK>(e));
                return r;
            }
        } else {  // ORD_GEMV_ROW
            constexpr int FULL = (K / PK) * PK;
            static_assert(FULL > 0, "row-major GEMV order needs K >= packet");
            T l[PK];
            l[0] = red_lane_seq<T, K, 0>(e);
            l[1] = red_lane_seq<T, K, 1>(e);
            if constexpr (PK == 4) {
                l[2] = red_lane_seq<T, K, 2>(e);
                l[3] = red_lane_seq<T, K, 3>(e);
            }
            T r = predux<T>(l);
#pragma unroll
            for (int j = FULL; j < K; ++j) r = N::add(r, e(j));
            return r;
        }
    }
}

// sum_k c(k)*x(k) in the named order (PARITY) or as one FMA chain (FAST).  (Spelled out rather than forwarded to reduce_products: with the
// forwarding form ptxas allocates and schedules the warp-per-instance kernel differently and it runs 8 % slower -- measured, round 2.)
template <class T, int ORD, int K, bool FAST, class C, class X>
__device__ __forceinline__ T dot(const C &c, const X &x)
{
    using N = Num<T>;
    constexpr int PK = N::PK;
    if constexpr (FAST) {
        T acc = N::mul(c(0), x(0));
#pragma unroll
        for (int k = 1; k < K; ++k) acc = N::fma(c(k), x(k), acc);
        return acc;
    } else {
        auto e = [&](int k) -> T { return N::mul(c(k), x(k)); };
        if constexpr (ORD == ORD_SEQ) {
            T acc = e(0);
#pragma unroll
            for (int k = 1; k < K; ++k) acc = N::add(e(k), acc);
            return acc;
        } else if constexpr (ORD == ORD_TREE) {
            return red_tree<T, 0, K>(e);
        } else if constexpr (ORD == ORD_VECREDUX) {
            constexpr int NP = K / PK;
            if constexpr (NP == 0) {
                return red_tree<T, 0, K>(e);
            } else {
                T l[PK];
                l[0] = red_ptree<T, 0, NP, 0>(e);
                l[1] = red_ptree<T, 0, NP, 1>(e);
                if constexpr (PK == 4) {
                    l[2] = red_ptree<T, 0, NP, 2>(e);
                    l[3] = red_ptree<T, 0, NP, 3>(e);
                }
                T r = predux<T>(l);
                if constexpr (NP * PK != K) r = N::add(r, red_tree<T, NP * PK, K - NP * PK>(e));
                return r;
            }
        } else {  // ORD_GEMV_ROW
            constexpr int FULL = (K / PK) * PK;
            static_assert(FULL > 0, "row-major GEMV order needs K >= packet");
            T l[PK];
            l[0] = red_lane_seq<T, K, 0>(e);
            l[1] = red_lane_seq<T, K, 1>(e);
            if constexpr (PK == 4) {
                l[2] = red_lane_seq<T, K, 2>(e);
                l[3] = red_lane_seq<T, K, 3>(e);
            }
            T r = predux<T>(l);
#pragma unroll
            for (int j = FULL; j < K; ++j) r = N::add(r, e(j));
            return r;
        }
    }
}

// out[r] = sum_k c(r, k) x(k), r = 0..R-1, every row in the named order -- written so that the R independent chains ADVANCE
// TOGETHER: sequential orders as a column sweep (term k of every row before term k+1 of any), tree orders four rows at a
// time with their products formed k-major.  The one-warp-per-scheduler instances of the generic kernel (fp64: 128 threads per
// SM) have nothing but this instruction-level parallelism to cover the FP64 pipe's latency; row-after-row dot products left
// the scheduler waiting on each row's own chain (ncu, profiles/r02_ncu_fp64.md: issue slots 25 % busy).
template <class T, int ORD, int R, int K, bool FAST, class C, class X>
__device__ __forceinline__ void matvec_rows(const C &c, const X &x, T (&out)[R])
{
    using N = Num<T>;
    if constexpr (FAST || ORD == ORD_SEQ) {
#pragma unroll
        for (int r = 0; r < R; ++r) out[r] = N::mul(c(r, 0), x(0));
#pragma unroll
        for (int k = 1; k < K; ++k)
#pragma unroll
            for (int r = 0; r < R; ++r) out[r] = FAST ? N::fma(c(r, k), x(k), out[r]) : N::add(N::mul(c(r, k), x(k)), out[r]);
    } else {
        constexpr int G = sizeof(T) == 8 ? 2 : 4;   // (a double product is two registers: larger groups spill)
#pragma unroll
        for (int r0 = 0; r0 < R; r0 += G) {
            T e[G][K];
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int j = 0; j < G; ++j)
                    if (r0 + j < R) e[j][k] = N::mul(c(r0 + j, k), x(k));
#pragma unroll
            for (int j = 0; j < G; ++j)
                if (r0 + j < R) out[r0 + j] = reduce_products<T, ORD, K>([&](int k) -> T { return e[j][k]; });
        }
    }
}

// ------------------------------------------------------------------------------------------------
// kernel parameters (constant bank)
// ------------------------------------------------------------------------------------------------
template <class T, int NX, int NU, int NH> struct alignas(16) Model {
    T K[NU * NX];   // Kinf     (nu x nx) col-major: K[r + k*NU]
    T A[NX * NX];   // Adyn                          A[r + k*NX]
    T B[NX * NU];   // Bdyn     (nx x nu)            B[r + k*NX]
    T Qi[NU * NU];  // Quu_inv
    T M[NX * NX];   // AmBKt
    T Pf[NX * NX];  // Pinf
    T Qd[NX];       // work.Q
    T xmin[NH * NX], xmax[NH * NX];
    T umin[(NH - 1) * NU], umax[(NH - 1) * NU];
    T rho, nrho, pri_tol, dua_tol;
    int max_iter, check_term;
};

template <class T> struct SolveArgs {
    long long batch;
    const T *x0;
    const T *Xref;
    long long xref_stride;  // elements between instances (0 = shared)
    T *wd, *wy, *wg, *wv, *wz;  // warm state, in place (all or none)
    T *x, *u;
    int *iter, *status;
    T *resid;
    unsigned long long *counter;  // next unclaimed instance
    unsigned long long *stats;    // [0] iterations [1] solved [2] lane-trips [3] instances
    unsigned *done;               // nullable: done[inst >> done_shift] += 1 when every output of inst is written
    int done_shift;
    const T *sys;                 // per-instance systems (PERSYS kernels): [instance][SysBlock::STRIDE], else null
    unsigned *gate;               // nullable: gate[0] = number of leading instances whose inputs have arrived in device memory
                                  // (advanced by stream memory operations between the chunks of an overlapped H2D), gate[1] = timeout flag
    const unsigned *order;        // nullable: the k-th claim of the work counter solves instance order[k] (longest-expected-first
                                  // schedule built by the host pre-pass, tmpc_api.cu lpt_prepare); null = instance k
    T *u0;                        // nullable: out [batch][nu] = u(:,0), the control an MPC caller applies (quadrotor_hovering.cpp:110)
    // Per-lane coalesced scratch of the fp32 12/4/10 kernel (tmpc_kernel_f32.cuh LaneScratch): 16-byte chunks laid out
    // [warp of the grid][chunk][lane of the warp], so that a warp's access to one chunk is one contiguous 512-byte run.  It
    // holds what is private to the lane's CURRENT instance, lives for many iterations and does not fit on chip:
    //   sc_ib: the instance's own box (tmpc_set_instance_bounds), copied in at refill, read by every forward sweep;
    //   sc_wm: mirror of d / v / z of a warm-started solve, written by the backward sweeps after which the next iteration may
    //          converge (the reference leaves d / v / z one iteration behind on an early exit), copied to the caller's
    //          buffers once, at termination.
    // Each sc_* is the first chunk of its region or -1.  L2-resident: <= 128 chunks x 37,888 lanes = 78 MB.
    // (A third candidate, -(Xref o Q) of per-instance trajectories precomputed at refill, LOST to re-reading the 48-byte Xref
    // rows in place through L1: 7.95 vs 7.25 ms per 1M tracking instances, profiles/r02_scratch_ab.log.)
    void *scratch;
    int sc_ib, sc_wm, sc_chunks;
    int test_flags;                           // 1, 2: tests only (1 = never predict the mirror: every early exit takes the re-solve
                                              // fall-back; 2 = mirror in every backward sweep); 4 = the warm duals y, g are ZERO
                                              // (closed loop with reset duals, quadrotor_hovering.cpp:100-101): not read at all;
                                              // 8: tests only, fused closed loop: exact hand-over at every step (no lazy first iteration)
    const T *ixmin, *ixmax, *iumin, *iumax;   // per-instance boxes [instance][N][nx] / [instance][N-1][nu]; a null pair = unbounded
    const void *model_g;                      // device copy of the kernel's model image (CSM instances of the fp32 kernel: staged
                                              // into shared memory by one TMA bulk copy per CTA), else null
    long long gate_split;                     // host pipeline, 0 or T: instances [0, T) are the tail segment -- transferred LAST, ranked on
                                              // the SMs the launch leaves free, claimed last (claims batch-T .. batch-1); gate[2] != 0 once
                                              // they have arrived and order[] holds their ranking.  gate[0] then covers [T, batch) only.
    // Fused closed loop (ROLL instances of the fp32 12/4/10 kernel, tmpc_kernel_f32.cuh): roll_steps > 1 MPC steps per instance in ONE
    // launch -- solve, plant step x0 <- x_1 of the solve, y = g = 0, warm d / v / z carried on chip -- for steps 0 .. roll_steps-2,
    // then the last step as a plain warm solve (outputs, state written back).  Histories of the steps the kernel completes
    // itself (each nullable): roll_x[k] = the state after step k, roll_u0[k] = the control applied, roll_iter / roll_status[k].
    int roll_steps, roll_step0;
    T *roll_x, *roll_u0;          // [roll_steps-1][batch][nx] / [..][nu]
    int *roll_iter, *roll_status; // [roll_steps-1][batch]
    // ... tracking a reference TABLE (quadrotor_tracking.cpp:101): step k of instance b follows rows w0 .. w0+N-1 of roll_table
    // [roll_rows][nx], w0 = min(roll_start[b] (null = 0) + roll_step0 + k, roll_rows - N); null table = the fixed Xref above
    const T *roll_table;
    long long roll_rows;
    const int *roll_start;
    // fp32 12/4/10 kernel, per-instance Xref: p_N seeds -(Xref_{N-1}^T Pinf) [batch][nx] computed by a pre-pass (null = by the lanes)
    const T *pn_seed;
};

// instance solved by the idx-th claim of the work counter
template <class T> __device__ __forceinline__ long long claimed_instance(const SolveArgs<T> &a, long long idx)
{
    return a.order ? (long long)__ldg(a.order + idx) : idx;
}

// Overlapped H2D: an instance may be claimed before its x0 / Xref chunk has landed.  Wait until gate[slot] exceeds `v` (rarely:
// the kernel consumes ~5 GB/s of inputs, PCIe delivers 30-55); give up after ~2 s instead of hanging the device (the host then
// releases the completion counters and reports the timeout).
template <class T> __device__ __forceinline__ bool gate_spin(const SolveArgs<T> &a, int slot, long long v)
{
    volatile unsigned *g = a.gate;
    if ((long long)g[slot] <= v) {
        const long long t0 = clock64();
        while ((long long)g[slot] <= v) {
            __nanosleep(200);
            if (clock64() - t0 > 4000000000LL) { g[1] = 1u; return false; }
        }
    }
    __threadfence();   // the inputs were written before the counter advanced: order the reads after the observation
    return true;
}
// The instance solved by the idx-th claim, once its inputs are in device memory; -1 if they never arrived.
// gate[0] = end of the index-ordered part that has arrived, gate[2] = the tail segment (see gate_split) is in and ranked.
template <class T> __device__ __forceinline__ long long claim_instance(const SolveArgs<T> &a, long long idx)
{
    if (!a.gate) return claimed_instance(a, idx);
    if (a.gate_split > 0 && idx >= a.batch - a.gate_split && !gate_spin(a, 2, 0)) return -1;   // before order[idx] is read
    const long long inst = claimed_instance(a, idx);
    if (inst >= a.gate_split && !gate_spin(a, 0, inst)) return -1;
    return inst;
}

// Per-instance system block (PERSYS: every instance brings its own model + cache; the "systems" batching axis),
// written by the batched precompute kernel.  Every matrix a mat-vec sweeps ROW-wise is stored row-major next to its
// column-major copy, so that the coefficient vector of each dot product is one contiguous, 16-byte aligned run that the
// kernel fetches with vector loads (4x fewer L1 requests than element-wise access to a lane-private 2-4 KB block).
__host__ __device__ constexpr int sys_al(int v) { return (v + 3) / 4 * 4; }
template <int NX, int NU> struct SysBlock {
    // the coefficients the ADMM loop sweeps, each dot product's vector contiguous; this prefix [0, TMLEN) is also the
    // image a lane keeps in tensor memory (SYS == 2: TMEM column = block offset)
    static constexpr int Krm = 0,                       // Kinf     row-major
                         Arm = sys_al(Krm + NU * NX),   // Adyn     row-major
                         Brm = sys_al(Arm + NX * NX),   // Bdyn     row-major
                         B = sys_al(Brm + NX * NU),     // Bdyn     col-major  (rows of Bdyn^T)
                         Qirm = sys_al(B + NX * NU),    // Quu_inv  row-major
                         Mrm = sys_al(Qirm + NU * NU),  // AmBKt    row-major
                         K = sys_al(Mrm + NX * NX),     // Kinf     col-major  (rows of Kinf^T)
                         TMLEN = sys_al(K + NU * NX),
                         // the rest: seed / getters
                         A = TMLEN,                     // Adyn     col-major
                         Qi = sys_al(A + NX * NX),      // Quu_inv  col-major
                         M = sys_al(Qi + NU * NU),      // AmBKt    col-major
                         Pf = sys_al(M + NX * NX),      // Pinf     col-major  (columns = what Xref^T Pinf sweeps)
                         Qd = sys_al(Pf + NX * NX),     // work.Q
                         RHO = sys_al(Qd + NX), LEN = RHO + 1, STRIDE = sys_al(LEN);
};

// ---- tensor memory as a per-thread scratchpad (tcgen05.ld/st.32x32b: thread t of a warp owns TMEM lane 32*(warp%4)+t) ----
__device__ __forceinline__ void tm_ld8(uint32_t a, float *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
                 : "r"(a) : "memory");
}
__device__ __forceinline__ void tm_ld4(uint32_t a, float *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]) : "r"(a) : "memory");
}
__device__ __forceinline__ void tm_st8(uint32_t a, const float *r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(a), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]) : "memory");
}
__device__ __forceinline__ void tm_st4(uint32_t a, const float *r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                 :: "r"(a), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]) : "memory");
}
__device__ __forceinline__ void tm_ld16(uint32_t a, float *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]), "=f"(r[9]),
                   "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
                 : "r"(a) : "memory");
}
__device__ __forceinline__ void tm_st16(uint32_t a, const float *r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(a), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]), "f"(r[8]), "f"(r[9]),
                    "f"(r[10]), "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15]) : "memory");
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// The loaded registers are only valid after tcgen05.wait::ld; passing them through the wait as in/out operands
// gives the compiler the data dependence (it must not schedule a use above the wait).
__device__ __forceinline__ void tm_wait4(float *r)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]) :: "memory");
}
__device__ __forceinline__ void tm_wait8(float *r)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]) :: "memory");
}
__device__ __forceinline__ void tm_wait12(float *r)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]),
                   "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]) :: "memory");
}
__device__ __forceinline__ void tm_wait16(float *r, float *q)   // a 12-vector and a 4-vector in flight together
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]),
                   "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]), "+f"(q[0]), "+f"(q[1]), "+f"(q[2]), "+f"(q[3]) :: "memory");
}

// per-thread array of STAGES vectors of D scalars in shared memory.
//  D % VEC == 0: [chunk][thread] with 16-byte chunks (conflict-free 128-bit access);
//  otherwise   : [element][thread] (conflict-free 32/64-bit access).
template <class T, int D, int STAGES, int BLOCK> struct SVec {
    using vec_t = typename Num<T>::vec_t;
    static constexpr int VEC = 16 / sizeof(T);
    static constexpr bool CHUNK = (D % VEC == 0);
    static constexpr int LEN = D * STAGES;
    static constexpr size_t BYTES =
        CHUNK ? size_t(LEN / VEC) * BLOCK * 16 : ((size_t(LEN) * BLOCK * sizeof(T) + 15) / 16) * 16;
    unsigned char *base;
    __device__ __forceinline__ SVec(unsigned char *b, int tid) : base(b + (CHUNK ? tid * 16 : tid * int(sizeof(T)))) {}
    __device__ __forceinline__ void load(int i, T (&o)[D]) const
    {
        if constexpr (CHUNK) {
#pragma unroll
            for (int c = 0; c < D / VEC; ++c) {
                vec_t t = *reinterpret_cast<const vec_t *>(base + size_t(i * (D / VEC) + c) * BLOCK * 16);
                o[c * VEC + 0] = t.x;
                o[c * VEC + 1] = t.y;
                if constexpr (VEC == 4) {
                    o[c * VEC + 2] = ((const T *)&t)[2];
                    o[c * VEC + 3] = ((const T *)&t)[3];
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < D; ++j)
                o[j] = *reinterpret_cast<const T *>(base + size_t(i * D + j) * BLOCK * sizeof(T));
        }
    }
    __device__ __forceinline__ void store(int i, const T (&o)[D], bool pred = true) const
    {
        if (!pred) return;
        if constexpr (CHUNK) {
#pragma unroll
            for (int c = 0; c < D / VEC; ++c) {
                vec_t t;
                t.x = o[c * VEC + 0];
                t.y = o[c * VEC + 1];
                if constexpr (VEC == 4) {
                    ((T *)&t)[2] = o[c * VEC + 2];
                    ((T *)&t)[3] = o[c * VEC + 3];
                }
                *reinterpret_cast<vec_t *>(base + size_t(i * (D / VEC) + c) * BLOCK * 16) = t;
            }
        } else {
#pragma unroll
            for (int j = 0; j < D; ++j)
                *reinterpret_cast<T *>(base + size_t(i * D + j) * BLOCK * sizeof(T)) = o[j];
        }
    }
};

// ---- g / v of a DOUBLE-precision instance in tensor memory (SYS == 3): a double is two 32-bit TMEM cells, stage i of the
// thread's vector sits at columns [base + 2 D i, + 2 D) of its TMEM lane.  tcgen05.ld/st are warp-collective: every lane of
// the warp must call load/store together (the sweeps do; the refill path is written accordingly).
__device__ __forceinline__ void tm_ld8u(uint32_t a, uint32_t *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a) : "memory");
}
__device__ __forceinline__ void tm_ld16u(uint32_t a, uint32_t *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(a) : "memory");
}
__device__ __forceinline__ void tm_st8u(uint32_t a, const uint32_t *r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tm_st16u(uint32_t a, const uint32_t *r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                    "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tm_wait8u(uint32_t *r)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]) :: "memory");
}
__device__ __forceinline__ void tm_wait24u(uint32_t *r)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                   "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]) :: "memory");
}
template <int D, int STAGES> struct TVecD {
    static_assert(D == 12 || D == 4, "TMEM rows of 24 or 8 cells");
    static constexpr int COLS = 2 * D * STAGES;
    uint32_t base = 0;
    __device__ __forceinline__ void load(int i, double (&o)[D]) const
    {
        uint32_t r[2 * D];
        const uint32_t a = base + (uint32_t)(i * 2 * D);
        if constexpr (D == 12) { tm_ld16u(a, r); tm_ld8u(a + 16, r + 16); tm_wait24u(r); }
        else { tm_ld8u(a, r); tm_wait8u(r); }
#pragma unroll
        for (int j = 0; j < D; ++j) o[j] = __hiloint2double((int)r[2 * j + 1], (int)r[2 * j]);
    }
    __device__ __forceinline__ void store(int i, const double (&o)[D]) const
    {
        uint32_t r[2 * D];
#pragma unroll
        for (int j = 0; j < D; ++j) { r[2 * j] = (uint32_t)__double2loint(o[j]); r[2 * j + 1] = (uint32_t)__double2hiint(o[j]); }
        const uint32_t a = base + (uint32_t)(i * 2 * D);
        if constexpr (D == 12) { tm_st16u(a, r); tm_st8u(a + 16, r + 16); }
        else tm_st8u(a, r);
    }
};

// global [instance][stage][dim] rows; 16-byte vector access when D is a multiple of the vector width
template <class T, int D> __device__ __forceinline__ void gload(const T *p, T (&o)[D])
{
    using vec_t = typename Num<T>::vec_t;
    constexpr int VEC = 16 / sizeof(T);
    if constexpr (D % VEC == 0) {
#pragma unroll
        for (int c = 0; c < D / VEC; ++c) {
            vec_t t = __ldg(reinterpret_cast<const vec_t *>(p) + c);
#pragma unroll
            for (int e = 0; e < VEC; ++e) o[c * VEC + e] = ((const T *)&t)[e];
        }
    } else {
#pragma unroll
        for (int j = 0; j < D; ++j) o[j] = __ldg(p + j);
    }
}
template <class T, int D> __device__ __forceinline__ void gstore(T *p, const T (&o)[D])
{
    using vec_t = typename Num<T>::vec_t;
    constexpr int VEC = 16 / sizeof(T);
    if constexpr (D % VEC == 0) {
#pragma unroll
        for (int c = 0; c < D / VEC; ++c) {
            vec_t t;
#pragma unroll
            for (int e = 0; e < VEC; ++e) ((T *)&t)[e] = o[c * VEC + e];
            reinterpret_cast<vec_t *>(p)[c] = t;
        }
    } else {
#pragma unroll
        for (int j = 0; j < D; ++j) p[j] = o[j];
    }
}

template <class T, int NX, int NU, int NH, int BLOCK> struct SmemLayout {
    using SU = SVec<T, NU, NH - 1, BLOCK>;
    using SX = SVec<T, NX, NH, BLOCK>;
    using SP = SVec<T, NX, 1, BLOCK>;
    static constexpr size_t BYTES = 3 * SU::BYTES + 2 * SX::BYTES + SP::BYTES;
    static constexpr size_t BYTES_NOGV = 3 * SU::BYTES + SP::BYTES;   // g, v in tensor memory (SYS == 3)
};

enum { PH_FREE = 0, PH_RUN = 1, PH_EMIT = 2 };

// coefficient vector of one dot product for the per-instance-systems kernels: `off` elements into the lane's block
// (16-byte vector loads through the read-only path), or (TM) the same offset in the lane's tensor-memory columns.
// issue() starts the fetch, ready() completes it, so the next row is in flight while the current one is consumed.
template <class T, int K, bool TM> __device__ __forceinline__ void crow_issue(const T *blk, uint32_t tcol, int off, T (&c)[K])
{
    if constexpr (TM) {
        static_assert(sizeof(T) == 4 && (K == 12 || K == 4), "TMEM rows: 12 or 4 floats");
        if constexpr (K == 12) { tm_ld8(tcol + off, c); tm_ld4(tcol + off + 8, c + 8); }
        else tm_ld4(tcol + off, c);
    } else {
        gload<T, K>(blk + off, c);
    }
}
template <class T, int K, bool TM> __device__ __forceinline__ void crow_ready(T (&c)[K])
{
    if constexpr (TM) { if constexpr (K == 12) tm_wait12(c); else tm_wait4(c); }
}
template <class T, int K1, int K2, bool TM> __device__ __forceinline__ void crow_ready2(T (&c1)[K1], T (&c2)[K2])
{
    if constexpr (TM) { static_assert(K1 == 12 && K2 == 4, "12 + 4"); tm_wait16(c1, c2); }
}

// SYS: 3 = one shared model, DOUBLE precision, g and v in tensor memory (480 of the lane's 512 cells at 12/4/10): 128 instead of
//          64 instances per SM -- at 64 threads two of the SM's four schedulers have no warp at all;
//      0 = one shared model (constant bank); 1 = per-instance systems, coefficients fetched from the lane's global block;
//      2 = per-instance systems, the lane's loop coefficients (SysBlock prefix, 496 floats at 12/4) resident in TENSOR
//          MEMORY: 128 threads per SM, one TMEM lane (512 columns) each, rewritten warp-collectively when a lane refills
template <class T, int NX, int NU, int NH, int BLOCK, bool FAST, bool WARM, bool UNROLL, int SYS = 0>
__global__ void __launch_bounds__(BLOCK, 1)
admm_kernel(const __grid_constant__ Model<T, NX, NU, NH> P, const __grid_constant__ SolveArgs<T> a)
{
    constexpr bool PERSYS = SYS == 1 || SYS == 2, SYSTM = SYS == 2, TMGV = SYS == 3;
    static_assert(!TMGV || (BLOCK == 128 && sizeof(T) == 8 && 2 * 2 * NX * NH <= 512), "g/v in TMEM: 4 warps, doubles, 4 NX NH cells per lane");
    static_assert(!SYSTM || (BLOCK == 128 && sizeof(T) == 4 && SysBlock<NX, NU>::TMLEN <= 512 && SysBlock<NX, NU>::TMLEN % 8 == 0),
                  "TMEM-resident systems: 4 warps, one 512-column TMEM lane per thread, float");
    // model source: the shared constant-bank image, or (PERSYS) this lane's own SysBlock: global memory read through the
    // read-only path, or its prefix kept in tensor memory (the ~2 KB of coefficients per instance do not fit registers)
    using SB = SysBlock<NX, NU>;
    const T *blk = PERSYS ? a.sys : nullptr;   // idle lanes keep a valid block (instance 0)
    T rho_l = P.rho, nrho_l = P.nrho;
    auto mQd = [&](int i) -> T { if constexpr (PERSYS) return __ldg(blk + SB::Qd + i); else return P.Qd[i]; };
    using N = Num<T>;
    using O = Orders<T, NX, NU>;
    using L = SmemLayout<T, NX, NU, NH, BLOCK>;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31;
    constexpr unsigned FULLM = 0xffffffffu;
    constexpr int XROW = NX * NH, UROW = NU * (NH - 1);

    unsigned char *sp = smem;
    typename L::SU sd(sp, tid); sp += L::SU::BYTES;
    typename L::SU sy(sp, tid); sp += L::SU::BYTES;
    typename L::SU sz(sp, tid); sp += L::SU::BYTES;
    using GV = std::conditional_t<TMGV, TVecD<(TMGV ? NX : 12), NH>, typename L::SX>;
    GV sg = [&] { if constexpr (TMGV) return GV{}; else { GV t(sp, tid); sp += L::SX::BYTES; return t; } }();
    GV sv = [&] { if constexpr (TMGV) return GV{}; else { GV t(sp, tid); sp += L::SX::BYTES; return t; } }();
    typename L::SP spn(sp, tid);
    uint32_t tcol = 0;   // SYSTM / TMGV: TMEM address of this thread's lane, column 0
    if constexpr (SYSTM || TMGV) {
        uint32_t *slot = reinterpret_cast<uint32_t *>(smem + (TMGV ? L::BYTES_NOGV : L::BYTES));
        if (tid < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tcol = *slot + ((uint32_t)(((tid >> 5) & 3) * 32) << 16);
        if constexpr (TMGV) { sg.base = tcol; sv.base = tcol + (uint32_t)(2 * NX * NH); }
    }

    // the p_N seed depends on Xref_{N-1} and Pinf only: with one model and one Xref for the whole batch every lane computes
    // it once here instead of inside the refill section, which runs at warp level nearly every trip
    const bool seed_shared = !PERSYS && a.xref_stride == 0;
    if (seed_shared) {
        T xr[NX], pn[NX];
        gload<T, NX>(a.Xref + (NH - 1) * NX, xr);
#pragma unroll
        for (int j = 0; j < NX; ++j)
            pn[j] = -dot<T, O::XtP, NX, FAST>([&](int k) { return P.Pf[k + j * NX]; }, [&](int k) { return xr[k]; });
        spn.store(0, pn);
    }
    long long inst = -1;
    int it = 0;
    int phase = PH_FREE;
    bool exhausted = false;
    bool hit_max = false;  // terminated by max_iter without converging (backward of that iteration still runs)
    bool deferred = false;  // warp-uniform: the previous trip postponed a single-lane refill
    T x0[NX];
    T res[4] = {T(0), T(0), T(0), T(0)};
    unsigned long long n_iter = 0, n_solved = 0, n_trips = 0, n_inst = 0;
#pragma unroll
    for (int j = 0; j < NX; ++j) x0[j] = T(0);

    for (;;) {
        // ------------------------------------------------------------------ lane refill
        const bool need = (phase == PH_FREE) && !exhausted;
        unsigned m = __ballot_sync(FULLM, need);
        {   // a single free lane waits one trip for a second one: the refill section runs for the whole warp (tmpc_kernel_f32.cuh)
            const bool others_busy = __ballot_sync(FULLM, phase != PH_FREE) != 0;
            if (m && __popc(m) < 2 && !deferred && others_busy) { deferred = true; m = 0; }
            else deferred = false;
        }
        if (m) {
            const int leader = __ffs(m) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(a.counter, (unsigned long long)__popc(m));
            base = __shfl_sync(FULLM, base, leader);
            if (need) {
                const long long idx = (long long)base + __popc(m & ((1u << lane) - 1u));
                const long long ci = idx < a.batch ? claim_instance(a, idx) : -1;
                if (ci >= 0) {
                    inst = ci;
                    phase = PH_RUN;
                    it = 0;
                    hit_max = false;
                    res[0] = res[1] = res[2] = res[3] = T(0);
                    gload<T, NX>(a.x0 + inst * NX, x0);
                    if constexpr (PERSYS) {
                        blk = a.sys + inst * SB::STRIDE;
                        rho_l = __ldg(blk + SB::RHO);
                        nrho_l = -rho_l;
                    }
                    // p_N seed: -(Xref_{N-1}^T * Pinf)   (admm.cpp:83); with one shared model and Xref it was computed once, above
                    if (!seed_shared) {
                        T xr[NX], pn[NX];
                        gload<T, NX>(a.Xref + inst * a.xref_stride + (NH - 1) * NX, xr);
#pragma unroll
                        for (int j = 0; j < NX; ++j) {
                            T c[NX];
                            if constexpr (PERSYS) gload<T, NX>(blk + SB::Pf + j * NX, c);
                            pn[j] = -dot<T, O::XtP, NX, FAST>([&](int k) { if constexpr (PERSYS) return c[k]; else return P.Pf[k + j * NX]; },
                                                              [&](int k) { return xr[k]; });
                        }
                        spn.store(0, pn);
                    }
                    if (WARM && a.wd) {
#pragma unroll 1
                        for (int i = 0; i < NH - 1; ++i) {
                            T t[NU];
                            gload<T, NU>(a.wd + inst * UROW + i * NU, t); sd.store(i, t);
                            gload<T, NU>(a.wy + inst * UROW + i * NU, t); sy.store(i, t);
                            gload<T, NU>(a.wz + inst * UROW + i * NU, t); sz.store(i, t);
                        }
                        if constexpr (!TMGV) {
#pragma unroll 1
                            for (int i = 0; i < NH; ++i) {
                                T t[NX];
                                gload<T, NX>(a.wg + inst * XROW + i * NX, t); sg.store(i, t);
                                gload<T, NX>(a.wv + inst * XROW + i * NX, t); sv.store(i, t);
                            }
                        }
                    } else {
                        T zu[NU], zx[NX];
#pragma unroll
                        for (int j = 0; j < NU; ++j) zu[j] = T(0);
#pragma unroll
                        for (int j = 0; j < NX; ++j) zx[j] = T(0);
#pragma unroll 1
                        for (int i = 0; i < NH - 1; ++i) { sd.store(i, zu); sy.store(i, zu); sz.store(i, zu); }
                        if constexpr (!TMGV) {
#pragma unroll 1
                            for (int i = 0; i < NH; ++i) { sg.store(i, zx); sv.store(i, zx); }
                        }
                    }
                } else {
                    exhausted = true;
                }
            }
            if constexpr (TMGV) {
                // tcgen05 is warp-collective: every lane of the warp rewrites its g / v cells, refilled lanes with zeros (cold) or
                // the caller's warm state, the others with what they hold
                const bool fill = need && !exhausted;
                const bool wfill = fill && WARM && a.wd;
#pragma unroll 1
                for (int i = 0; i < NH; ++i) {
                    T g[NX], v[NX];
                    sg.load(i, g);
                    sv.load(i, v);
                    if (wfill) {
                        gload<T, NX>(a.wg + inst * XROW + i * NX, g);
                        gload<T, NX>(a.wv + inst * XROW + i * NX, v);
                    } else if (fill) {
#pragma unroll
                        for (int j = 0; j < NX; ++j) g[j] = v[j] = T(0);
                    }
                    sg.store(i, g);
                    sv.store(i, v);
                }
                tm_wait_st();
            }
            if constexpr (SYSTM) {
                // tcgen05 is warp-collective: every lane rewrites its columns, lanes that were not refilled with what they hold
                const bool fill = need && !exhausted;
                // eight 8-column groups per tensor-memory round trip: the section is bound by the latency of the reads
                auto put8 = [&](int c8, float (&t)[8]) {
                    if (fill) {
                        const float4 lo = __ldg(reinterpret_cast<const float4 *>(blk + c8)), hi = __ldg(reinterpret_cast<const float4 *>(blk + c8 + 4));
                        t[0] = lo.x; t[1] = lo.y; t[2] = lo.z; t[3] = lo.w; t[4] = hi.x; t[5] = hi.y; t[6] = hi.z; t[7] = hi.w;
                    }
                    tm_st8(tcol + c8, t);
                };
                constexpr int G32 = (SB::TMLEN / 64) * 64;
#pragma unroll 1
                for (int c = 0; c < G32; c += 64) {
                    float t[8][8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) tm_ld8(tcol + c + 8 * q, t[q]);
#pragma unroll
                    for (int q = 0; q < 8; ++q) tm_wait8(t[q]);
#pragma unroll
                    for (int q = 0; q < 8; ++q) put8(c + 8 * q, t[q]);
                }
#pragma unroll 1
                for (int c8 = G32; c8 < SB::TMLEN; c8 += 8) {
                    float t[8];
                    tm_ld8(tcol + c8, t);
                    tm_wait8(t);
                    put8(c8, t);
                }
                tm_wait_st();
            }
        }
        if (__all_sync(FULLM, phase == PH_FREE)) break;
        ++n_trips;

        const bool emit = (phase == PH_EMIT);
        if (phase == PH_RUN) ++it;

        // ------------------------------------------------------------------ forward sweep
        // forward_pass (admm.cpp:27-37) + update_slack (:45-61) + update_dual (:67-71) + residual maxima (:95-98)
        T pri_x = T(0), dua_x = T(0), pri_u = T(0), dua_u = T(0);
        {
            T x[NX];
#pragma unroll
            for (int j = 0; j < NX; ++j) x[j] = x0[j];
            T *xo = (emit && a.x) ? a.x + inst * XROW : nullptr;
            T *uo = (emit && a.u) ? a.u + inst * UROW : nullptr;
            T *go = (WARM && emit && a.wg) ? a.wg + inst * XROW : nullptr;
            T *yo = (WARM && emit && a.wy) ? a.wy + inst * UROW : nullptr;

            auto stage = [&](int i, bool last) {
                T g[NX], v[NX], vn[NX];
                sg.load(i, g);
                sv.load(i, v);
                if (WARM && go) gstore<T, NX>(go + i * NX, g);
#pragma unroll
                for (int j = 0; j < NX; ++j) {
                    vn[j] = N::add(x[j], g[j]);                                                  // :48
                    vn[j] = N::mn(P.xmax[i * NX + j], N::mx(P.xmin[i * NX + j], vn[j]));         // :59
                    pri_x = N::mx(pri_x, N::abs(N::sub(x[j], vn[j])));                           // :95
                    dua_x = N::mx(dua_x, N::abs(N::sub(v[j], vn[j])));                           // :96
                    g[j] = N::sub(N::add(g[j], x[j]), vn[j]);                                    // :70
                }
                sg.store(i, g);
                sv.store(i, vn);
                if (xo) gstore<T, NX>(xo + i * NX, x);
                if (!last) {
                    T d[NU], y[NU], z[NU], u[NU], zn[NU];
                    sd.load(i, d);
                    sy.load(i, y);
                    sz.load(i, z);
                    if (WARM && yo) gstore<T, NU>(yo + i * NU, y);
                    T ck[2][NX];
                    T kxv[NU];
                    if constexpr (PERSYS) crow_issue<T, NX, SYSTM>(blk, tcol, SB::Krm, ck[0]);
                    else matvec_rows<T, O::Kx, NU, NX, FAST>([&](int r, int k) { return P.K[r + k * NU]; }, [&](int k) { return x[k]; }, kxv);
#pragma unroll
                    for (int r = 0; r < NU; ++r) {
                        T kx;
                        if constexpr (PERSYS) {
                            crow_ready<T, NX, SYSTM>(ck[r & 1]);
                            if (r + 1 < NU) crow_issue<T, NX, SYSTM>(blk, tcol, SB::Krm + (r + 1) * NX, ck[(r + 1) & 1]);
                            kx = dot<T, O::Kx, NX, FAST>([&](int k) { return ck[r & 1][k]; }, [&](int k) { return x[k]; });
                        } else {
                            kx = kxv[r];
                        }
                        u[r] = N::sub(-kx, d[r]);                                                // :31
                        zn[r] = N::add(u[r], y[r]);                                              // :47
                        zn[r] = N::mn(P.umax[i * NU + r], N::mx(P.umin[i * NU + r], zn[r]));     // :53
                        pri_u = N::mx(pri_u, N::abs(N::sub(u[r], zn[r])));                       // :97
                        dua_u = N::mx(dua_u, N::abs(N::sub(z[r], zn[r])));                       // :98
                        y[r] = N::sub(N::add(y[r], u[r]), zn[r]);                                // :69
                    }
                    sy.store(i, y);
                    sz.store(i, zn);
                    if (uo) gstore<T, NU>(uo + i * NU, u);
                    if (emit && a.u0 && i == 0) gstore<T, NU>(a.u0 + inst * NU, u);
                    T xn[NX];
                    T ca[2][NX], cb[2][NU];
                    if constexpr (PERSYS) { crow_issue<T, NX, SYSTM>(blk, tcol, SB::Arm, ca[0]); crow_issue<T, NU, SYSTM>(blk, tcol, SB::Brm, cb[0]); }
                    if constexpr (!PERSYS) {
                        // shared model: all rows of Adyn x (and of Bdyn u) advance together (matvec_rows)
                        matvec_rows<T, O::Ax, NX, NX, FAST>([&](int r, int k) { return P.A[r + k * NX]; }, [&](int k) { return x[k]; }, xn);
                        if constexpr (FAST) {
#pragma unroll
                            for (int k = 0; k < NU; ++k)
#pragma unroll
                                for (int r = 0; r < NX; ++r) xn[r] = N::fma(P.B[r + k * NX], u[k], xn[r]);
                        } else {
                            T bu[NX];
                            matvec_rows<T, O::Bu, NX, NU, false>([&](int r, int k) { return P.B[r + k * NX]; }, [&](int k) { return u[k]; }, bu);
#pragma unroll
                            for (int r = 0; r < NX; ++r) xn[r] = N::add(xn[r], bu[r]);                  // :35
                        }
                    }
#pragma unroll
                    for (int r = 0; r < (PERSYS ? NX : 0); ++r) {
                        if constexpr (PERSYS) {
                            if constexpr (SYSTM) crow_ready2<T, NX, NU, true>(ca[r & 1], cb[r & 1]);
                            if (r + 1 < NX) {
                                crow_issue<T, NX, SYSTM>(blk, tcol, SB::Arm + (r + 1) * NX, ca[(r + 1) & 1]);
                                crow_issue<T, NU, SYSTM>(blk, tcol, SB::Brm + (r + 1) * NU, cb[(r + 1) & 1]);
                            }
                        }
                        T ax = dot<T, O::Ax, NX, FAST>([&](int k) { if constexpr (PERSYS) return ca[r & 1][k]; else return P.A[r + k * NX]; },
                                                       [&](int k) { return x[k]; });
                        if constexpr (FAST) {
                            T acc = ax;
#pragma unroll
                            for (int k = 0; k < NU; ++k) acc = N::fma(PERSYS ? cb[r & 1][k] : P.B[r + k * NX], u[k], acc);
                            xn[r] = acc;
                        } else {
                            T bu = dot<T, O::Bu, NU, FAST>([&](int k) { if constexpr (PERSYS) return cb[r & 1][k]; else return P.B[r + k * NX]; },
                                                           [&](int k) { return u[k]; });
                            xn[r] = N::add(ax, bu);                                              // :35
                        }
                    }
#pragma unroll
                    for (int j = 0; j < NX; ++j) x[j] = xn[j];
                }
            };
            if constexpr (UNROLL) {
#pragma unroll
                for (int i = 0; i < NH - 1; ++i) stage(i, false);
            } else {
#pragma unroll 1
                for (int i = 0; i < NH - 1; ++i) stage(i, false);
            }
            stage(NH - 1, true);
            if constexpr (TMGV) tm_wait_st();   // the backward sweep reads the cells this sweep wrote
        }

        // ------------------------------------------------------------------ termination (admm.cpp:91-109, :135-138)
        bool final_bwd = false;  // WARM: max_iter exit still runs the backward pass of its last iteration
        if (phase == PH_RUN) {
            const bool chk = (it % P.check_term) == 0;
            if (chk) {
                res[0] = pri_x;
                res[1] = N::mul(dua_x, rho_l);
                res[2] = pri_u;
                res[3] = N::mul(dua_u, rho_l);
            }
            const bool conv = chk && res[0] < P.pri_tol && res[2] < P.pri_tol && res[1] < P.dua_tol &&
                              res[3] < P.dua_tol;
            if (conv || it >= P.max_iter) {
                if (a.iter) a.iter[inst] = it;
                if (a.status) a.status[inst] = conv ? 1 : 11;
                if (a.resid) {
                    a.resid[inst * 4 + 0] = res[0];
                    a.resid[inst * 4 + 1] = res[1];
                    a.resid[inst * 4 + 2] = res[2];
                    a.resid[inst * 4 + 3] = res[3];
                }
                n_iter += (unsigned)it;
                n_solved += conv ? 1u : 0u;
                ++n_inst;
                hit_max = !conv;
                final_bwd = !conv;
                phase = PH_EMIT;
            }
        } else if (phase == PH_EMIT) {
            phase = PH_FREE;  // trajectory was written by this trip's forward sweep
            if (a.done) { __threadfence(); atomicAdd(a.done + (inst >> a.done_shift), 1u); }
        }

        // ------------------------------------------------------------------ backward sweep
        // update_linear_cost (admm.cpp:77-85) recomputed per stage + backward_pass_grad (:15-22)
        const bool cont = (phase == PH_RUN);
        const bool wout = WARM && (cont || final_bwd) && a.wd;
        if (__any_sync(FULLM, cont || wout)) {
            T p[NX];
            const T *xr_base = a.Xref + (inst < 0 ? 0 : inst) * a.xref_stride;
            T *wdo = wout ? a.wd + inst * UROW : nullptr;
            T *wvo = wout ? a.wv + inst * XROW : nullptr;
            T *wzo = wout ? a.wz + inst * UROW : nullptr;
            {
                T v[NX], g[NX], pn[NX];
                sv.load(NH - 1, v);
                sg.load(NH - 1, g);
                spn.load(0, pn);
                if (WARM && wvo) gstore<T, NX>(wvo + (NH - 1) * NX, v);
#pragma unroll
                for (int j = 0; j < NX; ++j) {
                    if constexpr (FAST) p[j] = N::fma(nrho_l, N::sub(v[j], g[j]), pn[j]);
                    else p[j] = N::sub(pn[j], N::mul(rho_l, N::sub(v[j], g[j])));                // :84
                }
            }
            auto bstage = [&](int i) {
                T z[NU], y[NU], r[NU], v[NX], g[NX], xr[NX], q[NX];
                sz.load(i, z);
                sy.load(i, y);
                sv.load(i, v);
                sg.load(i, g);
                gload<T, NX>(xr_base + i * NX, xr);
                if (WARM && wvo) { gstore<T, NX>(wvo + i * NX, v); gstore<T, NU>(wzo + i * NU, z); }
#pragma unroll
                for (int j = 0; j < NU; ++j) r[j] = N::mul(nrho_l, N::sub(z[j], y[j]));          // :80
#pragma unroll
                for (int j = 0; j < NX; ++j) {
                    T cq = -N::mul(xr[j], mQd(j));                                              // :81
                    if constexpr (FAST) q[j] = N::fma(nrho_l, N::sub(v[j], g[j]), cq);
                    else q[j] = N::sub(cq, N::mul(rho_l, N::sub(v[j], g[j])));                   // :82
                }
                T s[NU], d[NU];
                T cbt[2][NX];
                if constexpr (PERSYS) crow_issue<T, NX, SYSTM>(blk, tcol, SB::B, cbt[0]);
                if constexpr (!PERSYS) {
                    T bp[NU];
                    matvec_rows<T, O::Btp, NU, NX, FAST>([&](int rr, int k) { return P.B[k + rr * NX]; }, [&](int k) { return p[k]; }, bp);
#pragma unroll
                    for (int j = 0; j < NU; ++j) s[j] = N::add(bp[j], r[j]);
                    matvec_rows<T, O::Qs, NU, NU, FAST>([&](int rr, int k) { return P.Qi[rr + k * NU]; }, [&](int k) { return s[k]; }, d);   // :19
                }
#pragma unroll
                for (int r_ = 0; r_ < (PERSYS ? NU : 0); ++r_) {
                    if constexpr (PERSYS) {
                        crow_ready<T, NX, SYSTM>(cbt[r_ & 1]);
                        if (r_ + 1 < NU) crow_issue<T, NX, SYSTM>(blk, tcol, SB::B + (r_ + 1) * NX, cbt[(r_ + 1) & 1]);
                    }
                    T bp = dot<T, O::Btp, NX, FAST>([&](int k) { if constexpr (PERSYS) return cbt[r_ & 1][k]; else return P.B[k + r_ * NX]; },
                                                    [&](int k) { return p[k]; });
                    s[r_] = N::add(bp, r[r_]);
                }
                T cqi[2][NU];
                if constexpr (PERSYS) crow_issue<T, NU, SYSTM>(blk, tcol, SB::Qirm, cqi[0]);
#pragma unroll
                for (int r_ = 0; r_ < (PERSYS ? NU : 0); ++r_) {
                    if constexpr (PERSYS) {
                        crow_ready<T, NU, SYSTM>(cqi[r_ & 1]);
                        if (r_ + 1 < NU) crow_issue<T, NU, SYSTM>(blk, tcol, SB::Qirm + (r_ + 1) * NU, cqi[(r_ + 1) & 1]);
                    }
                    d[r_] = dot<T, O::Qs, NU, FAST>([&](int k) { if constexpr (PERSYS) return cqi[r_ & 1][k]; else return P.Qi[r_ + k * NU]; },
                                                    [&](int k) { return s[k]; });                // :19
                }
                sd.store(i, d, cont);
                if (WARM && wdo) gstore<T, NU>(wdo + i * NU, d);
                T pn[NX];
                T cm[2][NX], ckt[2][NU];
                if constexpr (PERSYS) { crow_issue<T, NX, SYSTM>(blk, tcol, SB::Mrm, cm[0]); crow_issue<T, NU, SYSTM>(blk, tcol, SB::K, ckt[0]); }
                if constexpr (!PERSYS) {
                    T mp[NX], kr[NX];
                    matvec_rows<T, O::Mp, NX, NX, FAST>([&](int rr, int k) { return P.M[rr + k * NX]; }, [&](int k) { return p[k]; }, mp);
                    matvec_rows<T, O::Ktr, NX, NU, FAST>([&](int rr, int k) { return P.K[k + rr * NU]; }, [&](int k) { return r[k]; }, kr);
#pragma unroll
                    for (int j = 0; j < NX; ++j) pn[j] = N::sub(N::add(q[j], mp[j]), kr[j]);            // :20
                }
#pragma unroll
                for (int r_ = 0; r_ < (PERSYS ? NX : 0); ++r_) {
                    if constexpr (PERSYS) {
                        if constexpr (SYSTM) crow_ready2<T, NX, NU, true>(cm[r_ & 1], ckt[r_ & 1]);
                        if (r_ + 1 < NX) {
                            crow_issue<T, NX, SYSTM>(blk, tcol, SB::Mrm + (r_ + 1) * NX, cm[(r_ + 1) & 1]);
                            crow_issue<T, NU, SYSTM>(blk, tcol, SB::K + (r_ + 1) * NU, ckt[(r_ + 1) & 1]);
                        }
                    }
                    T mp = dot<T, O::Mp, NX, FAST>([&](int k) { if constexpr (PERSYS) return cm[r_ & 1][k]; else return P.M[r_ + k * NX]; },
                                                   [&](int k) { return p[k]; });
                    T kr = dot<T, O::Ktr, NU, FAST>([&](int k) { if constexpr (PERSYS) return ckt[r_ & 1][k]; else return P.K[k + r_ * NU]; },
                                                    [&](int k) { return r[k]; });
                    pn[r_] = N::sub(N::add(q[r_], mp), kr);                                      // :20
                }
#pragma unroll
                for (int j = 0; j < NX; ++j) p[j] = pn[j];
            };
            if constexpr (UNROLL) {
#pragma unroll
                for (int i = NH - 2; i >= 0; --i) bstage(i);
            } else {
#pragma unroll 1
                for (int i = NH - 2; i >= 0; --i) bstage(i);
            }
        }
        (void)hit_max;
    }

    // ---------------------------------------------------------------------- statistics
    if (a.stats) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_iter += __shfl_down_sync(FULLM, n_iter, o);
            n_solved += __shfl_down_sync(FULLM, n_solved, o);
            n_trips += __shfl_down_sync(FULLM, n_trips, o);
            n_inst += __shfl_down_sync(FULLM, n_inst, o);
        }
        if (lane == 0) {
            atomicAdd(a.stats + 0, n_iter);
            atomicAdd(a.stats + 1, n_solved);
            atomicAdd(a.stats + 2, n_trips);
            atomicAdd(a.stats + 3, n_inst);
        }
    }
    if constexpr (SYSTM || TMGV) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tcol) : "memory");
    }
}

}  // namespace tmpc
