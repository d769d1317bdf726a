// The reference's solver entry points (/root/reference/src/tinympc/admm.hpp:10-18), same names and meaning,
// executed on the B200 through the C ABI of include/tmpc.h.  There is no CPU implementation behind them.
#pragma once
#include "types.hpp"

#ifdef __cplusplus
extern "C" {
#endif

/* Run ADMM on solver->work (one instance).  Returns 0 when converged (work->status = 1), 1 when max_iter was
 * reached (status = 11) as admm.cpp:111-152; a negative value is a device/argument error (tiny_last_error()).
 * Reads x(:,0), Xref, d, y, g, v, z; leaves x, u, iter, status, the four residuals and d, y, g, v, z exactly as
 * the reference does (warm start).  vnew, znew, q, r, p are scratch that the next solve rewrites before reading
 * and are not refreshed on the host. */
int tiny_solve(TinySolver *solver);

/* The six step functions, one instance, each a device launch on the solver's workspace (unit-test surface). */
void forward_pass(TinySolver *solver);
void update_slack(TinySolver *solver);
void update_dual(TinySolver *solver);
void update_linear_cost(TinySolver *solver);
bool termination_condition(TinySolver *solver);
void backward_pass_grad(TinySolver *solver);

#ifdef __cplusplus
}
#endif
