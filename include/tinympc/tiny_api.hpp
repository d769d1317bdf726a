// Set-up / precompute / batched solve entry points named by the project brief.  The reference has no
// tiny_setup or tiny_precompute: "setup" is the hand-written initialisation block of each example
// (examples/quadrotor_hovering.cpp:33-78) and the cache math lives inside tiny_codegen (codegen.cpp:254-292);
// these functions package exactly those two pieces.
#pragma once
#include <stdint.h>
#include "admm.hpp"

#ifdef __cplusplus
extern "C" {
#endif

/* Allocate a solver for an nx/nu/N problem and fill it like the examples do: model (column-major inputs, as
 * tiny_codegen takes them, codegen.cpp:245-252), raw Q and R, rho, bounds (a NULL min/max pair disables that
 * bound: en_*_bound = 0, codegen.cpp:227-243), all work arrays zero, settings tol 1e-3 / max_iter 100 /
 * check_termination 1.  The cache is NOT computed here: call tiny_precompute or fill solver->cache yourself. */
int tiny_setup(TinySolver **out, int nx, int nu, int N, const tinytype *Adyn, const tinytype *Bdyn,
               const tinytype *Q, const tinytype *R, tinytype rho, const tinytype *x_min, const tinytype *x_max,
               const tinytype *u_min, const tinytype *u_max, int verbose);

/* Kinf, Pinf, Quu_inv, AmBKt, coeff_d2p by the reference's recursion (codegen.cpp:254-292): Riccati fixed point
 * on Q+rho, R+rho from P = rho*I, at most 1000 sweeps, stop when max|dKinf| < 1e-5; computed in double, stored
 * as tinytype.  Returns the number of sweeps (the reference never reports non-convergence; neither do we).
 * work.Q keeps the value the caller gave (the examples pass raw Q; generated code passes Q+rho). */
int tiny_precompute(TinySolver *solver);

/* The same recursion on plain column-major arrays (no solver object): what a foreign caller binds, and the host twin of
 * the batched device precompute (tmpc_systems_precompute), which it matches bit for bit.  Returns the sweep count. */
int tiny_precompute_raw(int nx, int nu, const tinytype *Adyn, const tinytype *Bdyn, const tinytype *Q, const tinytype *R, tinytype rho,
                        tinytype *Kinf, tinytype *Pinf, tinytype *Quu_inv, tinytype *AmBKt);

typedef struct {
    int64_t batch;
    const tinytype *x0;    /* [batch][nx] */
    const tinytype *Xref;  /* [N][nx] if xref_shared, else [batch][N][nx] */
    int xref_shared;
    int on_device;         /* 0: host pointers, 1: device pointers (then `stream` is a cudaStream_t) */
    void *stream;
    tinytype *d, *y, *g, *v, *z; /* warm state in place, all or none (NULL = cold start) */
} TinyBatchIn;

typedef struct {
    tinytype *x;      /* [batch][N][nx]   */
    tinytype *u;      /* [batch][N-1][nu] */
    int32_t *iter;    /* [batch] */
    int32_t *status;  /* [batch] 1 solved / 11 max_iter */
    tinytype *resid;  /* [batch][4] primal_state, dual_state, primal_input, dual_input; nullable */
    tinytype *u0;     /* [batch][nu] u(:,0) alone: what an MPC loop applies (quadrotor_hovering.cpp:110); nullable.  Every
                         pointer of this struct may be NULL (= not wanted): a controls-only caller sets u0, iter, status */
} TinyBatchOut;

/* tiny_solve for `batch` instances sharing solver's model, cache, bounds and settings.  0 = the call worked
 * (instances that stop at max_iter have status 11), negative = error. */
int tiny_solve_batch(TinySolver *solver, const TinyBatchIn *in, TinyBatchOut *out);

/* The closed loop every example of the reference runs around tiny_solve (quadrotor_hovering.cpp:90-114, quadrotor_tracking.cpp:93-118,
 * codegen_cartpole.cpp:75-122), for `batch` instances and `steps` MPC steps in ONE call, entirely on the device:
 *     [Xref = rows w0 .. w0+N-1 of `table`, w0 = min(start[b] + k, rows - N)]  ->  [y = g = 0 if reset_duals]  ->  tiny_solve
 *     (warm: d, v, z carried from step to step)  ->  x0 <- Adyn x0 + Bdyn u(:,0)
 * starting from the cold workspace the examples zero.  Host pointers.  Forwards to tmpc_batch_* (include/tmpc.h); with reset_duals
 * on a float build of the 12/4/10 or 4/1/10 shape the whole loop is one persistent kernel launch.  Batches of at least 32,768
 * instances are spread over the devices tiny_set_devices selects (contiguous instance ranges, tmpc_multi_rollout). */
typedef struct {
    int64_t batch;
    int32_t steps;
    int32_t reset_duals;
    const tinytype *x0;      /* [batch][nx] initial measurements */
    const tinytype *Xref;    /* fixed reference: [N][nx] if xref_shared, else [batch][N][nx]; ignored when table != NULL */
    int32_t xref_shared;
    const tinytype *table;   /* reference table [table_rows][nx] (Xref_total transposed), or NULL */
    int64_t table_rows;
    const int32_t *start;    /* [batch] first row of each instance's window at step 0; NULL = 0 */
} TinyRolloutIn;

typedef struct {             /* every pointer may be NULL (= not wanted) */
    tinytype *x_hist;        /* [steps+1][batch][nx]: the measurement each step started from; entry `steps` = the final plant state */
    tinytype *u0_hist;       /* [steps][batch][nu]: the control applied */
    int32_t *iter_hist;      /* [steps][batch] */
    int32_t *status_hist;    /* [steps][batch] */
    tinytype *x, *u;         /* trajectories of the LAST solve: [batch][N][nx], [batch][N-1][nu] */
} TinyRolloutOut;

int tiny_rollout_batch(TinySolver *solver, const TinyRolloutIn *in, TinyRolloutOut *out);

/* One box per instance for tiny_solve_batch: the wrapper's set_xmin / set_xmax / set_umin / set_umax
 * (tiny_wrapper.cpp:43-129) with a leading batch dimension.  x_min, x_max [batch][N][nx], u_min, u_max [batch][N-1][nu]
 * (host arrays, or device arrays when on_device != 0); copied.  While set, tiny_solve_batch must be called with exactly
 * this batch (tiny_solve itself, one instance, is refused); batch = 0 returns to the bounds of solver->work.
 * Forwards to tmpc_set_instance_bounds (include/tmpc.h). */
int tiny_set_instance_bounds(TinySolver *solver, int64_t batch, const tinytype *x_min, const tinytype *x_max,
                             const tinytype *u_min, const tinytype *u_max, int on_device);

/* policy 0 = bit-exact order of the reference's -O3 SSE2 build (default), 1 = FMA-contracted.  Per-instance bounds survive
 * the change. */
int tiny_set_order_policy(TinySolver *solver, int policy);

/* Devices tiny_solve_batch spreads a HOST batch over: 0 = every visible device (default), n = the first n.  Batches of at
 * least 32,768 instances are split into contiguous instance ranges, one per device, solved concurrently (tmpc_multi_solve);
 * results are identical to a one-device solve.  Device-memory batches and tiny_solve run on device 0. */
int tiny_set_devices(TinySolver *solver, int n);

/* Codegen-compatible data files (the reference's tiny_codegen output, codegen.cpp:322-477 and :131-160).
 * export: write solver's settings, cache, model and bounds as a `tiny_data_workspace.cpp` / `glob_opts.hpp` that a
 *         generated TinyMPC project compiles as is (work arrays zero, as the generator writes them; Q and R exactly as
 *         stored in solver->work -- the generator stores Q+rho, R+rho, codegen.cpp:255-256,431-436).
 * import: parse a generated `tiny_data_workspace.cpp` (sizes are inferred from the matrices) into a new solver. */
int tiny_export_data_workspace(const TinySolver *solver, const char *path);
int tiny_export_glob_opts(const TinySolver *solver, const char *path);
int tiny_import_data_workspace(TinySolver **out, const char *path);

void tiny_free(TinySolver *solver);
const char *tiny_last_error(void);

#ifdef __cplusplus
}
#endif
