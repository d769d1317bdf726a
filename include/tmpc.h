/* tmpc.h -- C ABI of the B200-native batched TinyMPC solver (libtmpc_cuda.so).
 *
 * This is the drop-in boundary for the cached-Riccati ADMM hot path of ucb-bar/Accelerated-TinyMPC:
 * one tmpc_solve() call runs the reference's whole `tiny_solve` loop (src/tinympc/admm.cpp:111-152:
 * forward_pass :27-37, update_slack :45-61, update_dual :67-71, update_linear_cost :77-85,
 * termination_condition :91-109, backward_pass_grad :15-22) for `batch` independent instances that share
 * one model/cache, inside one persistent sm_100a kernel.  Plain pointers and sizes only: no C++/Eigen/torch
 * types cross this boundary.  The reference has no FFI for a batch; the per-call pieces it does expose are
 * cited on each entry point.  Host-side C++ mirrors of TinySolver/TinyCache/TinyWorkspace/TinySettings and
 * tiny_setup/tiny_precompute/tiny_solve live in include/tinympc/ and call only this header.
 *
 * Conventions (identical to the reference's wire layout):
 *   - matrices are COLUMN-MAJOR (Eigen default, types.hpp:13-21; tiny_codegen inputs, codegen.cpp:245-252);
 *   - trajectories are [instance][stage][dim] (= column-major nx x N per instance, tiny_wrapper.cpp:27,154);
 *   - scalar type is float (TMPC_F32, what tiny_codegen emits, codegen.cpp:152) or double (TMPC_F64, the
 *     shipped glob_opts.hpp:3); every array of a ctx uses that one scalar type.
 * Every function returns TMPC_OK (0) or a negative tmpc_status; tmpc_last_error() gives the text.
 * There is NO CPU fallback: without a CUDA device every call that needs one fails with TMPC_ERR_CUDA.
 */
#ifndef TMPC_H
#define TMPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMPC_VERSION_MAJOR 0
#define TMPC_VERSION_MINOR 2

typedef enum {
    TMPC_OK = 0,
    TMPC_ERR_INVALID = -1,     /* bad argument (null pointer, check_termination < 1, ...) */
    TMPC_ERR_UNSUPPORTED = -2, /* shape / dtype / policy has no compiled kernel */
    TMPC_ERR_CUDA = -3,        /* CUDA runtime error (see tmpc_last_error) */
    TMPC_ERR_STATE = -4        /* call order (solve before set_model, ...) */
} tmpc_status;

typedef enum { TMPC_F32 = 0, TMPC_F64 = 1 } tmpc_dtype;

/* Arithmetic policy.
 *  PARITY: individually rounded products summed in the exact order of the reference's "-O3" SSE2 build
 *           -> per-instance iteration counts, status and x/u are bit-identical to tiny_solve.
 *  FAST:   the same recurrences with FMA contraction and sequential accumulation; agrees with the reference
 *          as well as the reference agrees with itself across compiler flags. */
typedef enum { TMPC_ORDER_PARITY = 0, TMPC_ORDER_FAST = 1 } tmpc_order;

typedef enum { TMPC_MEM_HOST = 0, TMPC_MEM_DEVICE = 1 } tmpc_mem;

/* status values written per instance: the reference's work->status (admm.cpp:114,136) */
#define TMPC_STATUS_SOLVED 1
#define TMPC_STATUS_UNSOLVED 11

typedef struct tmpc_ctx tmpc_ctx;

/* Create a solver context bound to one CUDA device for one problem shape.
 * Replaces the compile-time NSTATES/NINPUTS/NHORIZON/tinytype macros (glob_opts.hpp:3-9). */
int tmpc_create(tmpc_ctx **out, int device, int nx, int nu, int N, int dtype, int order_policy);
int tmpc_destroy(tmpc_ctx *ctx);

/* Shared model + cache + bounds: what the examples put into TinyCache and the parameter half of
 * TinyWorkspace (types.hpp:26-34,83-91; quadrotor_hovering.cpp:33-47).  Host pointers, ctx dtype.
 *   Kinf [nu x nx], Pinf [nx x nx], Quu_inv [nu x nu], AmBKt [nx x nx], Adyn [nx x nx], Bdyn [nx x nu]: col-major
 *   Q [nx] (used as given by update_linear_cost, admm.cpp:81);
 *   x_min,x_max [N][nx], u_min,u_max [N-1][nu]; a NULL pair leaves that bound unconstrained. */
int tmpc_set_model(tmpc_ctx *ctx, const void *Kinf, const void *Pinf, const void *Quu_inv, const void *AmBKt,
                   const void *Adyn, const void *Bdyn, const void *Q, double rho, const void *x_min,
                   const void *x_max, const void *u_min, const void *u_max);

/* TinySettings (types.hpp:39-47).  check_termination < 1 is rejected (it is `% 0` UB in admm.cpp:93);
 * max_iter < 1 is rejected. */
int tmpc_set_settings(tmpc_ctx *ctx, double abs_pri_tol, double abs_dua_tol, int max_iter,
                      int check_termination, int en_state_bound, int en_input_bound);

/* Per-instance box bounds: the wrapper's set_xmin / set_xmax / set_umin / set_umax (tiny_wrapper.cpp:43-129 fill
 * work.x_min ... work.u_max of its ONE workspace) with a leading batch dimension.  Arrays of the ctx dtype,
 *   x_min, x_max [batch][N][nx], u_min, u_max [batch][N-1][nu]  (host or device, `mem` = tmpc_mem); all four given.
 * The library keeps its own device copy.  While set, tmpc_solve / tmpc_batch_solve must be called with exactly
 * this batch; instance i is projected onto its own box (admm.cpp:53,59) -- en_state_bound / en_input_bound still
 * switch each family off.  fp32 12/4/10 keeps its specialised kernel (each lane copies its instance's box into its
 * coalesced, L2-resident scratch rows when it claims the instance and reads the stage's 128 bytes from there in every
 * forward sweep); every other shape runs on the run-time-shape kernel in index order.
 * batch = 0 (pointers ignored) returns to the shared bounds of tmpc_set_model.  Not applied by tmpc_step and
 * tmpc_solve_systems; a warm start from HOST memory is refused (use device buffers or a tmpc_batch). */
int tmpc_set_instance_bounds(tmpc_ctx *ctx, int64_t batch, const void *x_min, const void *x_max, const void *u_min,
                             const void *u_max, int32_t mem);

/* State that is live across tiny_solve calls (read before written: admm.cpp:31,47-48,69-70,96,98), per
 * instance, IN PLACE: read at the start of the solve, left exactly as the reference leaves its workspace
 * (d from the last executed backward pass, v/z one iteration behind vnew/znew on an early exit).
 * Cold start = these arrays zero-filled (quadrotor_hovering.cpp:49-71).  All five must be given. */
typedef struct {
    void *d; /* [batch][N-1][nu] */
    void *y; /* [batch][N-1][nu] */
    void *g; /* [batch][N][nx]   */
    void *v; /* [batch][N][nx]   */
    void *z; /* [batch][N-1][nu] */
} tmpc_warm;

typedef struct {
    int64_t batch;
    const void *x0;      /* [batch][nx]  -> work.x.col(0) (quadrotor_hovering.cpp:95, set_x0 tiny_wrapper.cpp:5) */
    const void *Xref;    /* [N][nx] shared, or [batch][N][nx] (set_xref, tiny_wrapper.cpp:21-41) */
    int32_t xref_shared; /* 1: one Xref for every instance */
    int32_t mem;         /* tmpc_mem: where x0/Xref/warm/outputs live */
    tmpc_warm *warm;     /* NULL = cold start, nothing written back */
    void *x;             /* out [batch][N][nx]    (get_x, tiny_wrapper.cpp:152-163); NULL = skip */
    void *u;             /* out [batch][N-1][nu]  (get_u, tiny_wrapper.cpp:165-176); NULL = skip */
    int32_t *iter;       /* out [batch] work->iter   (admm.cpp:120); NULL = skip */
    int32_t *status;     /* out [batch] work->status (1 solved / 11 not); NULL = skip */
    void *resid;         /* out [batch][4]: primal_state, dual_state, primal_input, dual_input (admm.cpp:95-98) */
    void *stream;        /* cudaStream_t (TMPC_MEM_DEVICE); NULL = the ctx's own stream */
    void *u0;            /* out [batch][nu]: u(:,0) alone -- the control an MPC loop applies (quadrotor_hovering.cpp:110
                            `x1 = Adyn x0 + Bdyn u.col(0)`); NULL = skip.  Every output pointer is its own mask bit: a
                            controls-only caller passes x = u = NULL and u0 (+ iter / status) and gets 24 instead of 648
                            bytes per quadrotor solve back, and the kernel then needs no trajectory-emission pass */
} tmpc_solve_args;

/* Run tiny_solve for every instance of the batch.  With TMPC_MEM_DEVICE the call is asynchronous on
 * `stream`; with TMPC_MEM_HOST it stages through pinned buffers in chunks (H2D / solve / D2H overlapped)
 * and returns when the outputs are in host memory.  Returns TMPC_OK even if some instances stop at
 * max_iter (their status is 11, as in the reference).
 * Ordering: a ctx owns one work counter, one statistics slot and its scratch areas, so the launches of one ctx
 * are SERIALISED ON THE DEVICE whatever streams they are queued on (each launch waits for the previous launch
 * of the same ctx through an event); independent solves that should overlap need one ctx each.  One host thread
 * at a time may call into a ctx. */
int tmpc_solve(tmpc_ctx *ctx, const tmpc_solve_args *args);

typedef struct {
    int64_t instances;        /* instances solved by the last tmpc_solve */
    int64_t iterations;       /* sum of per-instance iteration counts */
    int64_t solved;           /* instances with status 1 */
    int64_t trips;            /* lane-trips executed by the persistent kernel (iterations + emission + idle) */
    int32_t launches;         /* kernels launched by the last tmpc_solve */
    int32_t lanes;            /* resident instance slots (threads) the kernel ran with */
    float kernel_ms;          /* device time of those kernels (CUDA events on the launching stream) */
    int32_t parity_pinned;    /* 1 if this shape's evaluation order is verified against the reference */
    int32_t pattern;          /* model-structure specialisation the kernel ran with: 0 dense, 1 quadrotor (exact zeros
                                 and ones of Adyn / AmBKt dropped; value-identical results) */
    int32_t scheduled;        /* 1 if the last solve used the longest-expected-first schedule (device-resident batches of at
                                 least twice the resident lanes: a key per instance + a radix sort ahead of the kernel, on the
                                 same stream and inside kernel_ms; TMPC_LPT=0 disables it) */
} tmpc_stats;

/* Statistics of the last tmpc_solve on this ctx (synchronises the ctx's stream). */
int tmpc_get_stats(tmpc_ctx *ctx, tmpc_stats *out);

/* Full per-instance workspace (every array of TinyWorkspace that the step functions touch, types.hpp:55-76),
 * [batch][stage][dim]; all pointers required except resid/term. */
typedef struct {
    void *x, *u, *q, *r, *p, *d, *v, *vnew, *z, *znew, *g, *y;
    const void *Xref;    /* [N][nx] shared or [batch][N][nx] */
    int32_t xref_shared;
    void *resid;         /* [batch][4], read and written by step 4 */
    int32_t *term;       /* [batch] out: termination_condition() result (step 4) */
} tmpc_workspace;

/* One of the reference's step functions (admm.hpp:13-18) on every instance of a batch, as its own kernel:
 * which = 0 forward_pass, 1 update_slack, 2 update_dual, 3 update_linear_cost, 4 termination_condition (uses
 * `iter` for the check_termination test), 5 backward_pass_grad.  Unit-test surface; synchronous for host memory. */
int tmpc_step(tmpc_ctx *ctx, int which, int64_t batch, const tmpc_workspace *ws, int32_t iter, int32_t mem, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Device-resident batch of workspaces: the reference's wrapper API (tiny_wrapper.hpp:14-23 -- set_x0, set_xref,
 * reset_dual_variables, call_tiny_solve, get_x, get_u on ONE global workspace) with a leading batch dimension,
 * and the closed loop of the reference's examples kept on the device.  A tmpc_batch owns, in HBM, what
 * TinyWorkspace holds per instance: x(:,0), Xref, the warm state d y g v z, and the results x u iter status
 * residuals.  Every solve is a warm start from whatever the workspace holds (exactly like tiny_solve); a new
 * batch is zero-filled like the examples' init block (quadrotor_hovering.cpp:49-71).
 * `mem` says where the caller's array lives (tmpc_mem).  Bounds and model are the ctx's (tmpc_set_model /
 * tmpc_set_settings: set_umin/set_umax/set_xmin/set_xmax of the wrapper apply to every instance of the ctx;
 * tmpc_set_instance_bounds gives every instance its own).
 * All work is queued on the ctx's stream; calls with host arrays return when the copy is complete. */
typedef struct tmpc_batch tmpc_batch;
int tmpc_batch_create(tmpc_ctx *ctx, int64_t batch, tmpc_batch **out);
int tmpc_batch_destroy(tmpc_batch *b);
int tmpc_batch_set_x0(tmpc_batch *b, const void *x0, int32_t mem);                    /* [batch][nx]  set_x0, tiny_wrapper.cpp:5-19 */
int tmpc_batch_set_xref(tmpc_batch *b, const void *xref, int32_t shared, int32_t mem); /* [N][nx] or [batch][N][nx]  set_xref :21-41 */
/* Reference trajectory table [rows][nx] (examples/quadrotor_tracking.cpp:84) + per-instance first row start[batch]
 * (NULL = 0): at rollout step k instance b tracks rows w0..w0+N-1, w0 = min(start[b] + k, rows - N) (tracking.cpp:101). */
int tmpc_batch_set_xref_table(tmpc_batch *b, const void *table, int64_t rows, const int32_t *start, int32_t mem);
int tmpc_batch_reset_dual_variables(tmpc_batch *b);                                    /* y = 0, g = 0  reset_dual_variables :131-140 */
int tmpc_batch_reset(tmpc_batch *b);                                                   /* d y g v z = 0: cold start */
int tmpc_batch_solve(tmpc_batch *b);                                                   /* call_tiny_solve :142-150 for every instance (async) */
typedef enum {
    TMPC_GET_X = 0,      /* [batch][N][nx]    get_x, tiny_wrapper.cpp:152-163 */
    TMPC_GET_U = 1,      /* [batch][N-1][nu]  get_u :165-176 */
    TMPC_GET_ITER = 2,   /* int32 [batch] */
    TMPC_GET_STATUS = 3, /* int32 [batch] */
    TMPC_GET_RESID = 4,  /* [batch][4] */
    TMPC_GET_X0 = 5,     /* [batch][nx]: the current measurement (after a rollout: the plant state) */
    TMPC_GET_D = 6, TMPC_GET_Y = 7, TMPC_GET_Z = 8, /* [batch][N-1][nu] */
    TMPC_GET_G = 9, TMPC_GET_V = 10                 /* [batch][N][nx] */
} tmpc_batch_field;
int tmpc_batch_get(tmpc_batch *b, int32_t what, void *dst, int32_t mem);
/* The examples' closed loop (quadrotor_hovering.cpp:90-114, quadrotor_tracking.cpp:93-118) for `steps` MPC steps,
 * entirely on the device: [reference window from the table] -> [y = g = 0 if reset_duals] -> tiny_solve ->
 * x0 <- Adyn x0 + Bdyn u(:,0) (the examples' plant step, same evaluation order).  Histories are optional (NULL):
 * x0_hist [steps+1][batch][nx] (entry 0 = the initial state), u0_hist [steps][batch][nu], iter_hist / status_hist
 * [steps][batch].  Synchronous.
 * With reset_duals != 0 on an fp32 PARITY context of the 12/4/10 or 4/1/10 shape (shared bounds) the whole rollout is ONE
 * persistent launch: every lane takes its instance through all the steps with the state on chip (same results bit for bit;
 * the environment switch TMPC_ROLL=0 keeps one launch per step). */
int tmpc_batch_rollout(tmpc_batch *b, int32_t steps, int32_t reset_duals, void *x0_hist, void *u0_hist, int32_t *iter_hist,
                       int32_t *status_hist, int32_t mem);
float tmpc_batch_last_rollout_ms(const tmpc_batch *b);   /* device time of the last rollout (CUDA events on the ctx stream) */
const char *tmpc_batch_last_error(const tmpc_batch *b);

/* ---------------------------------------------------------------------------------------------------------
 * Per-instance SYSTEMS: every instance of the batch has its own model (Adyn, Bdyn, Q, R, rho) and therefore its own
 * cache.  tmpc_systems_precompute runs the reference's cache recursion (tiny_codegen, codegen.cpp:254-292: Riccati
 * fixed point on Q+rho, R+rho from P = rho*I, <= 1000 sweeps, stop at max|dKinf| < 1e-5; Quu_inv, AmBKt) for all
 * instances on the device, one instance per thread, in double, bit-identical to the host tiny_precompute, and keeps
 * the result device-resident; tmpc_solve_systems is tmpc_solve with those per-instance models.  Bounds, tolerances,
 * max_iter and check_termination remain the ctx's (tmpc_set_model must have been called; its matrices are unused here).
 *   Adyn [batch][nx*nx], Bdyn [batch][nx*nu] column-major; Q [batch][nx]; R [batch][nu]; rho [batch]; ctx dtype.
 *   q_plus_rho: the Q used by update_linear_cost (admm.cpp:81) is Q+rho (as tiny_codegen stores it, codegen.cpp:255,433)
 *               or Q as given (as the examples do, quadrotor_20hz_params.hpp:89). */
typedef struct tmpc_systems tmpc_systems;
int tmpc_systems_precompute(tmpc_ctx *ctx, int64_t batch, const void *Adyn, const void *Bdyn, const void *Q, const void *R,
                            const void *rho, int32_t q_plus_rho, int32_t mem, tmpc_systems **out);
int tmpc_systems_destroy(tmpc_systems *s);
typedef enum {
    TMPC_SYS_KINF = 0,    /* [batch][nu*nx] column-major */
    TMPC_SYS_PINF = 1,    /* [batch][nx*nx] */
    TMPC_SYS_QUU_INV = 2, /* [batch][nu*nu] */
    TMPC_SYS_AMBKT = 3,   /* [batch][nx*nx] */
    TMPC_SYS_ADYN = 4, TMPC_SYS_BDYN = 5, TMPC_SYS_Q = 6, TMPC_SYS_RHO = 7,
    TMPC_SYS_SWEEPS = 8   /* int32 [batch]: Riccati sweeps (1000 = not converged, as silent as the reference; -1 = singular) */
} tmpc_systems_field;
int tmpc_systems_get(tmpc_systems *s, int32_t what, void *dst_host);
/* tmpc_solve for a batch whose instance i uses system i.  Device buffers only (args->mem = TMPC_MEM_DEVICE). */
int tmpc_solve_systems(tmpc_ctx *ctx, const tmpc_solve_args *args, const tmpc_systems *systems);

/* ---------------------------------------------------------------------------------------------------------
 * One batch over SEVERAL devices from one process: the multi-GPU form of tmpc_solve for HOST-memory callers (the reference's
 * callers are plain C++ loops, quadrotor_hovering.cpp:104; there is nothing to shard in the reference).  A tmpc_multi owns one
 * tmpc_ctx and one host worker thread per device; tmpc_multi_solve splits the batch into contiguous instance ranges
 * [B*r/G, B*(r+1)/G) (instances never interact, admm.cpp:111-152: no inter-device traffic), every worker runs tmpc_solve
 * (TMPC_MEM_HOST) on its range of the caller's arrays concurrently, and the call returns when all outputs are in host memory.
 * Results are identical, instance by instance, to a single-device tmpc_solve.  Small batches use fewer devices (at least
 * 16,384 instances per device).  Pinned caller memory (tmpc_host_alloc) is DMA'd directly by every device.
 *   ndev = 0: every visible device; else devices[0..ndev) (NULL = 0..ndev-1). */
typedef struct tmpc_multi tmpc_multi;
int tmpc_multi_create(tmpc_multi **out, int ndev, const int *devices, int nx, int nu, int N, int dtype, int order_policy);
int tmpc_multi_destroy(tmpc_multi *m);
int tmpc_multi_device_count(const tmpc_multi *m);
tmpc_ctx *tmpc_multi_ctx(tmpc_multi *m, int index);   /* the index-th device's ctx (device-memory calls, batches, steps) */
int tmpc_multi_set_model(tmpc_multi *m, const void *Kinf, const void *Pinf, const void *Quu_inv, const void *AmBKt,
                         const void *Adyn, const void *Bdyn, const void *Q, double rho, const void *x_min,
                         const void *x_max, const void *u_min, const void *u_max);
int tmpc_multi_set_settings(tmpc_multi *m, double abs_pri_tol, double abs_dua_tol, int max_iter, int check_termination,
                            int en_state_bound, int en_input_bound);
/* per-instance boxes for the whole batch (HOST arrays; each device keeps its own range); batch = 0 clears them */
int tmpc_multi_set_instance_bounds(tmpc_multi *m, int64_t batch, const void *x_min, const void *x_max, const void *u_min,
                                   const void *u_max);
int tmpc_multi_solve(tmpc_multi *m, const tmpc_solve_args *args);   /* args->mem must be TMPC_MEM_HOST; args->stream ignored */
/* tmpc_batch_rollout for one HOST batch spread over the devices: every device runs the closed loop of its contiguous instance range
 * (from the cold workspace) and writes its slices of the histories.  Pointers as in tmpc_batch_set_x0 / set_xref / set_xref_table /
 * rollout / get; x0 and one of (Xref, table) are required, every output may be NULL. */
typedef struct {
    int64_t batch;
    int32_t steps, reset_duals;
    const void *x0;          /* [batch][nx] */
    const void *Xref;        /* [N][nx] if xref_shared, else [batch][N][nx]; ignored when table != NULL */
    int32_t xref_shared;
    const void *table;       /* [table_rows][nx] or NULL */
    int64_t table_rows;
    const int32_t *start;    /* [batch] or NULL */
    void *x0_hist;           /* [steps+1][batch][nx] */
    void *u0_hist;           /* [steps][batch][nu] */
    int32_t *iter_hist, *status_hist;   /* [steps][batch] */
    void *x, *u;             /* last step's trajectories [batch][N][nx] / [batch][N-1][nu] */
} tmpc_rollout_args;
int tmpc_multi_rollout(tmpc_multi *m, const tmpc_rollout_args *args);
/* totals over the devices of the last tmpc_multi_solve (kernel_ms = the slowest device's); per_device: NULL or an array of
 * tmpc_multi_device_count() entries */
int tmpc_multi_get_stats(tmpc_multi *m, tmpc_stats *total, tmpc_stats *per_device);
const char *tmpc_multi_last_error(const tmpc_multi *m);

/* Pinned host allocation helpers for TMPC_MEM_HOST callers that want full PCIe speed (portable: every device of the
 * process can DMA from / to it). */
int tmpc_host_alloc(void **ptr, uint64_t bytes);
int tmpc_host_free(void *ptr);

const char *tmpc_last_error(const tmpc_ctx *ctx); /* ctx may be NULL: last error of tmpc_create */
int tmpc_device_count(void);                      /* visible sm_100-class CUDA devices (0 when there is none) */
const char *tmpc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TMPC_H */
