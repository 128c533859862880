/*
 * pygent_b200 -- C ABI of one compiled "problem" (system dynamics + cost), sm_100a.
 *
 * This is the drop-in boundary of the simulation-and-planning hot path.  In the reference
 * (mpritzkoleit/pygent) the only native seam on this path is
 *
 *     sympy_to_c.convert_to_c(args, expr, cfilepath=..., use_exisiting_so=False) -> callable
 *
 * (call sites: pygent/modeling_scripts/cart_pole.py:159-170, acrobot.py:147-157,
 *  cart_pole_double_parallel.py:157-168, cart_pole_double_serial.py:170-181,
 *  cart_pole_triple.py:189-200, pygent/algorithms/ilqr.py:558-571): sympy expressions are
 * turned into C, compiled into a shared object and called through ctypes, one scalar
 * evaluation per call.  The B200 replacement compiles the SAME expressions into one shared
 * library per problem whose entry points below evaluate whole batched loops of the
 * reference on the GPU:
 *
 *   pg_rollout        <- the `for k: environment.step(u_k)` loops (pygent/environments.py:242-276,
 *                        fast_step :291-322; pygent/algorithms/ilqr.py:501-516 init_trajectory)
 *   pg_ilqr_forward   <- iLQR.forward_pass for one or all line-search alphas (ilqr.py:187-231, :429-437)
 *   pg_ilqr_backward  <- iLQR.linearization + cost_lin + backward_pass (ilqr.py:233-346, :518-533, :597-610)
 *   pg_ilqr_select    <- the line-search bookkeeping, accept/reject, mu schedule and stopping tests of
 *                        iLQR.run_optim (ilqr.py:418-485, :612-622)
 *   pg_ilqr_commit    <- "self.xx, self.uu, self.KK, self.kk = best candidate" (ilqr.py:458-464)
 *   pg_observe        <- helpers.observation (pygent/helpers.py:20-29)
 *   pg_eval           <- one batched evaluation of f, A, B / the cost expansion (what the reference's
 *                        generated callables return; used by the parity tests)
 *
 * Conventions
 *   - All array arguments are DEVICE pointers owned by the caller, except where a parameter is
 *     documented as "host".  Nothing is allocated or retained by the library.
 *   - Layout is trajectory-minor ("SoA"): the batch index b is the fastest dimension of every array,
 *     e.g. X[(k*n + i)*B + b].  `dtype` selects the scalar type of every `void*` array.
 *   - Every function enqueues its work on `stream` (a cudaStream_t passed as void*) and returns
 *     immediately; return value 0 = success, >0 = cudaError_t of the launch, <0 = PG_E_* below.
 *   - No CPU fallback exists: without a CUDA device the functions return the CUDA error.
 */
#ifndef PYGENT_B200_H
#define PYGENT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_ABI_VERSION 7

#define PG_F64 0
#define PG_F32 1

#define PG_INTEG_RK4 0   /* classical RK4, S substeps per control interval (parity schedule S=4) */
#define PG_INTEG_EULER 1 /* forward Euler, StateSpaceModel.fast_step (environments.py:291-322)  */

#define PG_LIN_COST_FD 4 /* OR-ed into lin_mode: iLQR(finite_diff=True), the cost expansion by central differences of c*dt with
                            eps = 1e-3 (helpers.py:41-128 via ilqr.py:240-242, :598-603); fp64 only */

#define PG_ROLLOUT_FINAL_COST 1 /* add fcost(x_T)*dt to the cost (ilqr.py:513) */
#define PG_ROLLOUT_ACCUMULATE 2 /* cost[b] holds the sum of earlier time segments of the same rollout */

#define PG_E_BADARG (-1)
#define PG_E_UNSUPPORTED (-2)

#define PG_MAX_ALPHAS 16

/* per-trajectory solver status written by pg_ilqr_select */
#define PG_ST_RUNNING 0
#define PG_ST_CONV_FUN 1   /* "Converged: small improvement" (ilqr.py:468-471) */
#define PG_ST_CONV_GRAD 2  /* "Converged: small gradient"    (ilqr.py:472-475) */
#define PG_ST_MU_MAX 3     /* "Diverged: no improvement"     (ilqr.py:482-485) */
#define PG_ST_MAX_ITERS 4

typedef struct {
    int32_t abi_version;
    int32_t n;               /* state dimension   */
    int32_t m;               /* control dimension */
    int32_t n_obs;           /* observation dimension: angles expand to [cos, sin] */
    int32_t has_ab;          /* analytic A, B compiled in (else central differences, helpers.py:130-148) */
    int32_t has_cost_derivs; /* analytic cost expansion compiled in */
    int32_t cost_uses_xnext; /* running cost depends on x_{k+1} */
    int32_t reserved;
    char name[64];
    char key[32];
} pg_problem_info_t;

int pg_get_problem_info(pg_problem_info_t* out);

/* Open-loop rollout of B independent environments over T control intervals.
 *   x0 [n,B]; U [T,m,B] or NULL (u == 0, ilqr.py:505); X [T+1,n,B] or NULL (history, X[0] = x0);
 *   xN [n,B] or NULL; cost [B] or NULL: sum_k (c(x_k,u_k,x_{k+1},t_{k+1}) + terminal_cost*terminated_{k+1})*dt
 *   (+ fcost(x_T)*dt with PG_ROLLOUT_FINAL_COST; added to the incoming cost[b] with PG_ROLLOUT_ACCUMULATE,
 *   which lets a long rollout be fed in time segments while the next segment's controls are still in
 *   flight from the host); terminated [B] or NULL: terminate(x_T).
 *   t0: simulation time of x0 (the reference's env.tt[-1]); time advances by repeated t += dt. */
int pg_rollout(int dtype, int integrator, int64_t B, int32_t T, int32_t S, double dt, double t0,
               const void* x0, const void* U, void* X, void* xN, void* cost, uint8_t* terminated,
               double terminal_cost, int32_t flags, void* stream);

/* The same rollout, load-balanced for large batches: persistent worker warps (3 per warp scheduler) pull
 * (32-trajectory group, 20-step time chunk) units from a ticket queue, so that all schedulers carry the same
 * load even when B/32 is not a multiple of the scheduler count (65,536 envs: 3.46 warps per scheduler).
 * Results are bit-identical to pg_rollout.  xN and cost are required (they carry the state between chunks);
 * workspace: device scratch of at least pg_rollout_workspace_bytes(B, T) bytes; it must not be shared by calls that may run
 * concurrently.  Bytes [0, 4) are a STICKY status word (i32): zero it once after allocation; a worker that gives up waiting for
 * its unit (~2 s; never expected) sets it to 1 and no launch ever clears it, so the caller can test it whenever it next
 * synchronises -- non-zero means the results of some earlier launch on this workspace are incomplete. */
int64_t pg_rollout_workspace_bytes(int64_t B, int32_t T);

/* pg_rollout / pg_rollout_balanced with the controls drawn INSIDE the kernel instead of read from U (SURVEY.md 8(d), config 2:
 * "u[k,b] ~ U(-uMax, uMax) ... or in-kernel Philox with the same counter scheme, validated against the host stream"):
 *     u[k, j, b] = (2 U - 1) * u_max[j],   U = ((r1 << 32 | r0) >> 11) * 2^-53,
 *     (r0, r1) = Philox2x32-10(counter = (b, 4 (k_offset + k) + j), key = Philox2x32-10((seed mod 2^32, seed div 2^32), 0x5047454E).r0)
 *     (one 64-bit block per action; B <= 2^32, k_offset + T <= 2^30, else PG_E_BADARG).
 * u_max: host array [m].  workspace == NULL: the plain kernel; else the ticket-queue kernel (pg_rollout_balanced's rules).
 * pg_random_actions writes the same stream to U [T, m, B] (for validation and for kernels that read U); the host-side
 * restatement is pygent_b200/random_actions.py.  A rollout over pg_random_actions' U equals pg_rollout_random bit for bit. */
int pg_rollout_random(int dtype, int integrator, int64_t B, int32_t T, int32_t S, double dt, double t0, const void* x0, uint64_t seed,
                      int32_t k_offset, const double* u_max, void* X, void* xN, void* cost, uint8_t* terminated, double terminal_cost,
                      int32_t flags, void* workspace, int64_t workspace_bytes, void* stream);
int pg_random_actions(int dtype, int64_t B, int32_t T, uint64_t seed, int32_t k_offset, const double* u_max, void* U, void* stream);
int pg_rollout_balanced(int dtype, int integrator, int64_t B, int32_t T, int32_t S, double dt, double t0,
                        const void* x0, const void* U, void* X, void* xN, void* cost, uint8_t* terminated,
                        double terminal_cost, int32_t flags, void* workspace, int64_t workspace_bytes, void* stream);

/* Alpha window of a SERIAL line search split over two launches (the reference stops at the first alpha that passes its
 * z-test, ilqr.py:453-454 -- about 4.8 of its 11 alphas on average -- so the later alphas are evaluated only for the
 * trajectories pg_ilqr_linesearch_filter keeps).  Both launches get the full alphas / output arrays:
 *   phase 1 evaluates alphas [0, n_first), phase 2 alphas [n_first, n_alpha) --
 *   unless *live_count <= small_count (device int32, read at kernel start): then phase 1 evaluates ALL alphas and phase 2
 *   none, because a second dependent chain of T control intervals costs more than it saves on a GPU that is not full.
 * NULL or phase 0: the whole search in one launch. */
typedef struct pg_alpha_window {
    int32_t phase;
    int32_t n_first;
    int32_t small_count;
    int32_t reserved;
    const int32_t* live_count;
} pg_alpha_window;

/* Closed-loop line-search rollouts u_k = K_j (x_k - xbar_j) + alpha k_j + ubar_j, j = min(Tk-1, k).
 *   alphas: host array of n_alpha values (<= PG_MAX_ALPHAS), all evaluated in one launch;
 *   alpha_dev [B] or NULL: when given, n_alpha must be 1 and each trajectory uses its own alpha;
 *   KK [Tk,m,n,B], kk [Tk,m,B]; ulim: host array [m] or NULL -- clip u to +-ulim (constrained=True);
 *   terminal_cost: the environment's terminal_cost (environments.py:275), 0 by default;
 *   X_out [n_alpha,T+1,n,B] / U_out [n_alpha,T,m,B] or NULL; cost_out [n_alpha,B];
 *   diverged_out [n_alpha,B] or NULL: any(x > 1e8) over the rollout (ilqr.py:435);
 *   terminated_out [n_alpha,B] or NULL; xN_out [n_alpha,n,B] or NULL: terminal state of every candidate (the state
 *   the reference's planner environment is left in after a forward pass; used by pg_mpc_shift);
 *   xS [n_alpha, pg_ilqr_num_snapshots(T), n, B] or NULL: state of every candidate at the boundaries of the 20-step time
 *   chunks (what pg_ilqr_commit_chunked restarts from);
 *   active [B] or NULL: trajectories with active==0 are skipped;
 *   index [<=B] (i32) + count [1] (i32, device) or both NULL: compacted work list -- only trajectories
 *   index[0..*count) are processed (written by pg_ilqr_select; late iLQR iterations touch few trajectories). */
int pg_ilqr_forward(int dtype, int integrator, int64_t B, int32_t T, int32_t S, double dt,
                    int32_t n_alpha, const double* alphas, const void* alpha_dev,
                    const void* x0, const void* xbar, const void* ubar,
                    const void* KK, const void* kk, int32_t Tk, const double* ulim, double terminal_cost,
                    void* X_out, void* U_out, void* cost_out, uint8_t* diverged_out,
                    uint8_t* terminated_out, void* xN_out, void* xS, const uint8_t* active,
                    const int32_t* index, const int32_t* count, const pg_alpha_window* window, void* stream);

/* pg_ilqr_forward for large line searches, load-balanced like pg_rollout_balanced: persistent worker warps pull
 * (32-entry group of the work list, alpha, 20-step time chunk) units from a ticket queue.  Costs and flags only (no
 * X_out / U_out); cost_out, diverged_out and xN_out are required (they carry a candidate between chunks).  Bit-identical
 * to pg_ilqr_forward.  workspace: device scratch of at least pg_ilqr_forward_workspace_bytes(B, T, n_alpha) bytes, with the
 * same sticky status word in bytes [0, 4) as pg_rollout_balanced. */
int64_t pg_ilqr_forward_workspace_bytes(int64_t B, int32_t T, int32_t n_alpha);
int pg_ilqr_forward_balanced(int dtype, int integrator, int64_t B, int32_t T, int32_t S, double dt,
                             int32_t n_alpha, const double* alphas, const void* alpha_dev,
                             const void* x0, const void* xbar, const void* ubar,
                             const void* KK, const void* kk, int32_t Tk, const double* ulim, double terminal_cost,
                             void* cost_out, uint8_t* diverged_out, uint8_t* terminated_out, void* xN_out, void* xS,
                             const int32_t* index, const int32_t* count,
                             void* workspace, int64_t workspace_bytes, const pg_alpha_window* window, void* stream);

/* Work lists of a SERIAL line search evaluated in two launches of pg_ilqr_forward(_balanced) (see pg_alpha_window):
 *   phase 0: index_out / count_out = the entries of index_in whose backward pass succeeded (bw_success [B]);
 *   phase 1: ... = the entries that, after the first n_first alphas (host array alphas; cost_cand, diverged [>= n_first, B];
 *            cost [B] the current costs, dV [2,B]), have neither passed the z-test of pg_ilqr_select nor diverged.
 *            Empty when *count_in <= small_count (launch 1 evaluated every alpha then).
 * count_out must be zero on entry (device int32); order within a warp is kept.  pg_ilqr_select reads only the alphas the
 * serial search tries, so evaluating the remaining alphas for the phase-1 list alone changes none of its decisions. */
int pg_ilqr_linesearch_filter(int dtype, int64_t B, int32_t phase, int32_t n_first, const double* alphas, double zmin,
                              const void* cost, const void* dV, const uint8_t* bw_success, const void* cost_cand,
                              const uint8_t* diverged, const int32_t* index_in, const int32_t* count_in,
                              int32_t* index_out, int32_t* count_out, int32_t small_count, void* stream);

/* Backward Riccati pass along (xbar, ubar).
 *   lin_mode 0: analytic A, B (requires has_ab); 1: central differences eps=1e-3 of f; | PG_LIN_COST_FD: the cost
 *   expansion (running and final) by central differences as well (finite_diff=True; PG_E_UNSUPPORTED in fp32).
 *   mu [B] per-trajectory regularisation; reg_type 1 or 2 (ilqr.py:291-295);
 *   ulim host [m] or NULL: box-constrained gains (ilqr.py:304-327, exact-clip restatement);
 *   Vxx_T [n,n,B] or NULL: the FINAL backward pass (ilqr.py:246-260, :487-494) -- terminal value Hessian P from the
 *   discrete Riccati equation at the final-cost equilibrium (solved on the host like the reference does, scipy), terminal
 *   gradient zero; the caller passes mu = 0 and ulim = NULL for it;
 *   outputs: KK [T,m,n,B], kk [T,m,B], dV [2,B] = (sum V1, sum V2), gnorm [B] (ilqr.py:418-420),
 *   success [B] (0 = Quu not positive definite at some step; KK/kk then incomplete). */
int pg_ilqr_backward(int dtype, int64_t B, int32_t T, double dt, int32_t reg_type, int32_t lin_mode,
                     const void* mu, const void* xbar, const void* ubar, const double* ulim, const void* Vxx_T,
                     void* KK, void* kk, void* dV, void* gnorm, uint8_t* success,
                     const uint8_t* active, const int32_t* index, const int32_t* count, void* stream);

typedef struct {
    int32_t n_alpha;
    int32_t parallel;     /* iLQR(parallel=...) (ilqr.py:453) */
    int32_t iteration;    /* 1-based iteration number of this call */
    int32_t max_iters;
    double alphas[PG_MAX_ALPHAS];
    double zmin, tol_fun, tol_grad;
    double mu_min, mu_max, mu_d0;
} pg_select_params_t;

/* One line-search decision per trajectory; updates the per-trajectory solver state in place.
 *   in : cost_cand [n_alpha,B], diverged [n_alpha,B], dV [2,B], gnorm [B], bw_success [B]
 *   io : cost [B], mu [B], mu_d [B], dcost [B], success_gradient [B] (u8), status [B] (i32), iters [B] (i32)
 *   out: alpha_best [B], accept [B] (u8), active [B] (u8), n_active (device int32 counter: += trajectories still
 *        running), index_out [B] or NULL: their indices, compacted (the next iteration's work list);
 *   index_in / count_in: this iteration's work list or NULL (all B);
 *   last_tried [B] (i32) or NULL: index of the last alpha the (serial) search evaluated, -1 when the backward pass
 *   failed and no forward pass ran. */
int pg_ilqr_select(int dtype, int64_t B, const pg_select_params_t* params,
                   const void* cost_cand, const uint8_t* diverged, const void* dV, const void* gnorm,
                   const uint8_t* bw_success, void* cost, void* mu, void* mu_d, void* dcost,
                   uint8_t* success_gradient, int32_t* status, int32_t* iters,
                   void* alpha_best, uint8_t* accept, uint8_t* active, int32_t* n_active,
                   const int32_t* index_in, const int32_t* count_in, int32_t* index_out, int32_t* last_tried,
                   void* stream);

/* For accepted trajectories: re-run the winning alpha and store it as the new nominal trajectory
 * (xbar, ubar updated in place) and copy the candidate gains into the accepted gains. */
int pg_ilqr_commit(int dtype, int integrator, int64_t B, int32_t T, int32_t S, double dt,
                   const void* alpha_best, const uint8_t* accept, const void* x0,
                   void* xbar, void* ubar, const void* KK, const void* kk, const double* ulim,
                   void* KK_acc, void* kk_acc, const int32_t* index, const int32_t* count, void* stream);

/* --- iLQR over dynamics integrated and linearised OUTSIDE this library (the learned-MLP model of MBRL,
 * include/pygent_b200_mlp.h; reference: pygent/algorithms/mbrl.py:76-93 builds its NMPC/iLQR planner on
 * StateSpaceModel(self.ode, ...), whose iLQR.linearization falls back to finite differences, ilqr.py:526).  The
 * problem library still owns the cost (sympy -> CUDA), the Riccati recursion and the line-search logic. */

/* pg_ilqr_backward with caller-supplied CONTINUOUS-time Jacobians A [T,n,n,B], Bm [T,n,m,B] of the dynamics along
 * (xbar, ubar); discretised like the reference, Ad = I + A dt, Bd = B dt (ilqr.py:530-531).
 * lin_flags: 0, or PG_LIN_COST_FD for the central-difference cost expansion (finite_diff=True; fp64 only). */
int pg_ilqr_backward_ext(int dtype, int64_t B, int32_t T, double dt, int32_t reg_type, int32_t lin_flags, const void* mu,
                         const void* xbar, const void* ubar, const void* A, const void* Bm, const double* ulim,
                         const void* Vxx_T, void* KK, void* kk, void* dV, void* gnorm, uint8_t* success,
                         const int32_t* index, const int32_t* count, void* stream);

/* Cost of candidate trajectories produced elsewhere: X [n_alpha,T+1,n,B], U [n_alpha,T,m,B] ->
 * cost_out [n_alpha,B] = sum_k (c(x_k,u_k,x_{k+1},t_{k+1}) + terminal_cost*terminated_{k+1}) dt + fcost(x_T) dt
 * (ilqr.py:213-224; t_0 = t0, time advances by repeated t += dt), diverged_out / terminated_out [n_alpha,B] as in
 * pg_ilqr_forward. */
int pg_traj_cost(int dtype, int64_t B, int32_t T, double dt, double t0, int32_t n_alpha, const void* X, const void* U,
                 double terminal_cost, void* cost_out, uint8_t* diverged_out, uint8_t* terminated_out,
                 const int32_t* index, const int32_t* count, void* stream);

/* pg_ilqr_commit without re-integration: for accepted trajectories copy candidate alpha_best[b] (looked up in the
 * host array alphas [n_alpha]) from X_cand / U_cand into xbar / ubar and the candidate gains into the accepted ones. */
int pg_ilqr_commit_copy(int dtype, int64_t B, int32_t T, int32_t n_alpha, const double* alphas,
                        const void* alpha_best, const uint8_t* accept, const void* X_cand, const void* U_cand,
                        void* xbar, void* ubar, const void* KK, const void* kk, void* KK_acc, void* kk_acc,
                        const int32_t* index, const int32_t* count, void* stream);

/* --- receding-horizon control: MPCAgent (pygent/algorithms/nmpc.py:126-299) --------------------------------------
 * pg_mpc_action = take_action (:187-211) after the planner iteration: u = K_0 (x - xbar_0) + ubar_0 + alpha k_0
 * (+ noise [m,B] or NULL), clipped to +-ulim (host [m]).  KK, kk, xbar, ubar: the planner's arrays (first interval is
 * read), alpha [B] = alpha_best, x [n,B] the measured state, u_out [m,B]. */
int pg_mpc_action(int dtype, int64_t B, const void* KK, const void* kk, const void* xbar, const void* ubar,
                  const void* alpha, const void* x, const void* noise, const double* ulim, void* u_out, void* stream);

/* pg_mpc_shift = shift_planner (:257-277), in place: cost -= c(x_0,u_0,x_1) dt + fcost(x_T) dt; xbar, ubar move one
 * interval forward, ubar_{T-1} = 0; the LAST entries of the gains KK, kk are zeroed (they are not shifted -- reference
 * behaviour); one interval with u = 0 is integrated from the planner-environment state and appended as xbar_T, and
 * cost += its transition cost (at time t_env + dt) + fcost(xbar_T) dt.  The planner-environment state is
 * xN_cand[last_tried[b]] (outputs of pg_ilqr_forward / pg_ilqr_select: end of the last line-search candidate
 * evaluated) or, when last_tried[b] < 0 or the pointers are NULL, xe [n,B]; xe receives the appended state.
 * x_new_ext [n,B] or NULL: the appended state when the caller integrates it (learned dynamics). */
int pg_mpc_shift(int dtype, int integrator, int64_t B, int32_t T, int32_t S, double dt, double t_env,
                 void* xbar, void* ubar, void* KK, void* kk, void* cost, void* xe, const void* xN_cand,
                 int32_t n_alpha, const int32_t* last_tried, const void* x_new_ext, double terminal_cost, void* stream);

/* pg_ilqr_commit with the winner's time chunks integrated in parallel: every (accepted trajectory, chunk) pair restarts
 * from the snapshot xS of candidate alpha_best[b] (looked up in the host array alphas) and overwrites its piece of
 * xbar / ubar in place (a chunk writes its own start state and leaves its end state to the next chunk's thread, so
 * nothing is read after it was overwritten), copying the candidate gains of its intervals as it goes.  Same results as
 * pg_ilqr_commit; n_chunks times shorter dependent chains. */
int32_t pg_ilqr_num_snapshots(int32_t T);
int pg_ilqr_commit_chunked(int dtype, int integrator, int64_t B, int32_t T, int32_t S, double dt, int32_t n_alpha,
                           const double* alphas, const void* alpha_best, const uint8_t* accept, const void* x0,
                           void* xbar, void* ubar, const void* KK, const void* kk, const double* ulim, const void* xS,
                           void* KK_acc, void* kk_acc, const int32_t* index, const int32_t* count, void* stream);

/* Terminal value of the FINAL backward pass, per trajectory, on the device (iLQR.backward_pass(final=True), ilqr.py:246-260):
 *   x* = argmin fcost(x) started at the trajectory's terminal state (the reference: scipy.optimize.minimize; here Newton's method
 *        on the generated gradient and Hessian of the final cost -- one step for the quadratic final costs of the examples),
 *   Ad = I + A(x*, 0) dt, Bd = B(x*, 0) dt, (Q, R, N) = the expansion of c dt at (x*, 0, t_final)   (ilqr.py:250-256),
 *   P  = scipy.linalg.solve_discrete_are(Ad, Bd, Q, R, s=N), by the structure-preserving doubling algorithm.
 * xbar [T+1, n, B] -> P [n, n, B] (symmetric; hand it to pg_ilqr_backward as Vxx_T), x_eq [n, B] or NULL, ok [B] (u8: both
 * iterations converged).  fp64 only; lin_mode 0 (analytic Jacobians) or 1 (central differences). */
int pg_ilqr_terminal_value(int dtype, int64_t B, int32_t T, double dt, double t_final, int32_t lin_mode, const void* xbar, void* P,
                           void* x_eq, uint8_t* ok, const int32_t* index, const int32_t* count, void* stream);

/* Fused tail solver (the "persistent fused loop" of iLQR.run_optim, ilqr.py:399-485): ONE launch that finishes every trajectory
 * of a work list.  A warp owns a trajectory and iterates backward sweep (laid over the lanes: one column of the value Hessian each) -> all n_alpha line-search candidates side by
 * side (one lane each) -> the reference's decision logic -> winner copied into the nominal trajectory, until that
 * trajectory stops (status != 0 or iters == select.max_iters), then takes the next one.  The arithmetic and every decision
 * are those of pg_ilqr_backward / pg_ilqr_forward / pg_ilqr_select / pg_ilqr_commit (bit-identical results), without the
 * launch boundaries, the host and the slowest trajectory pacing everybody; meant for the tail of a solve, when the running
 * trajectories no longer fill the GPU.  All arrays as in those functions; the struct holds plain pointers and sizes only.
 * Supported: fp64, lin_mode 0 / 1 (no PG_LIN_COST_FD), no externally supplied Jacobians or terminal value; else
 * PG_E_UNSUPPORTED (run the launch-per-step functions instead). */
typedef struct {
    int32_t dtype, integrator, T, S, n_alpha, reg_type, lin_mode, constrained; /* constrained: clip with ulim (both passes) */
    int64_t B;
    double dt, terminal_cost;
    double ulim[2];
    pg_select_params_t select;
    const void* x0;                                   /* [n, B] */
    void *xbar, *ubar, *KK, *kk;                      /* nominal trajectory and accepted gains: in / out */
    void *KKc, *kkc, *dV, *gnorm;                     /* candidate gains and backward-pass outputs (scratch, as pg_ilqr_backward) */
    uint8_t* bw_success;
    void* cost_cand;                                  /* [n_alpha, B] line-search outputs (scratch, as pg_ilqr_forward) */
    uint8_t *diverged, *terminated;
    void* xN;                                         /* [n_alpha, n, B] or NULL */
    void *cost, *mu, *mu_d, *dcost;                   /* solver state, as pg_ilqr_select */
    uint8_t* success_gradient;
    int32_t *status, *iters;
    void* alpha_best;
    uint8_t *accept, *active;
    int32_t* last_tried;                              /* or NULL */
    const int32_t *index, *count;                     /* work list (device) and its length (device, <= max_count) */
    int32_t max_count;
    void* workspace;                                  /* pg_ilqr_solve_workspace_bytes(T, max_count) bytes */
    int64_t workspace_bytes;
    /* Rounds: with round_iters > 0 a launch gives every trajectory of the list at most that many iterations and writes those
     * still running to index_out / count_out (device; count_out is zeroed by the call; order unspecified), the work list of
     * the next call.  Each call deals the trajectories to the warps anew -- the first 4 x SMs of the list get a warp scheduler
     * of their own -- so a few long-running trajectories never share a scheduler for the rest of a solve; a list of at most
     * 4 x SMs trajectories is finished by the call that finds it, whatever round_iters says.  0: run to the end. */
    int32_t round_iters;
    int32_t* index_out;
    int32_t* count_out;
} pg_ilqr_solve_t;
int64_t pg_ilqr_solve_workspace_bytes(int32_t T, int32_t max_count);
/* Trajectories the fused solver works on simultaneously (one resident warp each); a longer work list is handed out through a
 * ticket counter as warps become free.  Switching to pg_ilqr_solve pays once the running trajectories fit. */
int32_t pg_ilqr_solve_capacity(void);
int pg_ilqr_solve(const pg_ilqr_solve_t* args, void* stream);

/* o [n_obs,B] = observation(x [n,B]) : angle -> [cos, sin] (helpers.py:20-29). */
int pg_observe(int dtype, int64_t B, const void* x, void* o, void* stream);

/* Batched point evaluation (parity tests): any output may be NULL.
 *   x [n,B], u [m,B], t [B] -> f [n,B], A [n,n,B], Bm [n,m,B] (lin_mode as above),
 *   c [B] running cost (x' = x), cx [n,B], cu [m,B], Cxx [n,n,B], Cuu [m,m,B], Cxu [n,m,B],
 *   cf [B], cfx [n,B], Cfxx [n,n,B]  (costs NOT multiplied by dt). */
int pg_eval(int dtype, int64_t B, int32_t lin_mode, const void* x, const void* u, const void* t,
            void* f, void* A, void* Bm, void* c, void* cx, void* cu, void* Cxx, void* Cuu, void* Cxu,
            void* cf, void* cfx, void* Cfxx, void* stream);

#ifdef __cplusplus
}
#endif
#endif
