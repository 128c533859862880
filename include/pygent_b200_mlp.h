/*
 * pygent_b200 -- C ABI of the learned-MLP dynamics path (BASELINE.json config 5), sm_100a.
 *
 * Reference: pygent/nn_models.py:152-202 (NNDynamics: (oDim+uDim) -> 500 -> ReLU -> 500 -> ReLU -> dxDim, fp32,
 * input/output standardisation) and pygent/algorithms/mbrl.py:125-130 (MBRL.ode with sparse_dyn:
 * rhs = [x[n/2:], nn(obs(x), u) / dt]), integrated with forward Euler (StateSpaceModel.fast_step,
 * pygent/environments.py:291-322; MBRL builds its planner with fastForward=True, mbrl.py:81).
 *
 * One control interval for a batch of rows (trajectories, line-search candidates, finite-difference points):
 *     o  = (observation(x) - oMean) / oVar                     fp64 -> fp32   (helpers.py:20-29, nn_models.py:193-197)
 *     h1 = relu(W1 [o; (u - uMean)/uVar] + b1)                 CUDA cores (K <= 16)
 *     h2 = relu(W2 h1 + b2)                                    500 x 500: the dense GEMM of this path
 *     dv = (W3 h2 + b3) * dxVar + dxMean                       CUDA cores (N <= 4)
 *     x' = x + dt * [x[n/2:], dv / dt]                         fp64
 * precision PG_MLP_FP32: layer 2 in fp32 FFMA (tracks the reference's fp32 torch path to ~1e-6);
 * precision PG_MLP_BF16: layer 2 on the 5th-generation tensor cores (tcgen05.mma, bf16 operands, fp32
 *                        accumulation in TMEM), weights streamed from L2 by bulk async copies.  LOWER precision than the
 *                        reference's fp32 network (2e-2 per evaluation): an explicitly approximate fast path.
 * precision PG_MLP_F16X3: layer 2 on the tensor cores at the REFERENCE's precision: H1 and W2 are split x = hi + lo into two
 *                        fp16 numbers (22 mantissa bits) and the product is formed as Ah Wh + Al Wh + Ah Wl with fp32
 *                        accumulation; layer 1 in fp32 on the CUDA cores.  Tracks the fp32 network to ~1e-6 (stated: 1e-5).
 *
 * All pointers are device pointers owned by the caller; state arrays are trajectory-minor ([n, R], R fastest),
 * fp64; every function enqueues on `stream` and returns 0 / a cudaError_t / a negative PG_E_* code.
 */
#ifndef PYGENT_B200_MLP_H
#define PYGENT_B200_MLP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_MLP_FP32 0
#define PG_MLP_BF16 1
#define PG_MLP_F16X3 2

#define PG_MLP_HIDDEN 500     /* the reference's layer width (nn_models.py:179-181) */
#define PG_MLP_HPAD 512       /* padded width used by the kernels */
#define PG_MLP_MAX_IN 16      /* n_obs + m */
#define PG_MLP_MAX_DX 4       /* outputs (n/2 with sparse_dyn) */

typedef struct {
    int32_t n, m, n_obs, dx_dim; /* state, control, observation, network output dimensions; dx_dim == n/2 */
    uint8_t is_angle[8];         /* xIsAngle: angle states enter the network as [cos, sin] */
} pg_mlp_dims_t;

/* Size of the packed parameter block, and packing from the torch layout (host pointers, fp32):
 *   W1 [500, n_obs+m], b1 [500], W2 [500, 500], b2 [500], W3 [dx_dim, 500], b3 [dx_dim],
 *   o_mean/o_var [n_obs], u_mean/u_var [m], dx_mean/dx_var [dx_dim].
 * `packed_host` receives the block; copy it to the device (any 1024-byte aligned address). */
int64_t pg_mlp_packed_bytes(void);
int pg_mlp_pack(const pg_mlp_dims_t* dims, const float* W1, const float* b1, const float* W2, const float* b2,
                const float* W3, const float* b3, const float* o_mean, const float* o_var, const float* u_mean,
                const float* u_var, const float* dx_mean, const float* dx_var, void* packed_host);

/* rhs [n, R] = MBRL.ode(x [n, R], u [m, R]) : batched evaluation of the learned dynamics (used for the
 * finite-difference linearisation, helpers.py:130-148, and by the parity tests). */
int pg_mlp_ode(int precision, const pg_mlp_dims_t* dims, const void* packed, int64_t R, double dt,
               const double* x, const double* u, double* rhs, void* stream);

/* Euler rollout of B trajectories through the learned dynamics, T control intervals.
 *   open loop   : U [T, m, B] (or NULL: u = 0);
 *   closed loop : KK [T, m, n, B], kk [T, m, B], xbar [T+1, n, B], ubar [T, m, B], alpha -- iLQR.forward_pass
 *                 (ilqr.py:206): u_k = K_k (x_k - xbar_k) + alpha k_k + ubar_k; ulim [m] host or NULL clips u.
 *   X [T+1, n, B] receives the trajectory (X[0] = x0), U_out [T, m, B] or NULL the applied controls. */
int pg_mlp_rollout(int precision, const pg_mlp_dims_t* dims, const void* packed, int64_t B, int32_t T, double dt,
                   const double* x0, const double* U, const double* KK, const double* kk, const double* xbar,
                   const double* ubar, double alpha, const double* ulim, double* X, double* U_out, void* stream);

/* iLQR.forward_pass for n_alpha line-search candidates at once (ilqr.py:187-231, :429-437) -- and pg_mlp_rollout is its
 * n_alpha = 1 case: rows of the launch are (candidate, trajectory) pairs.
 *   alphas: host array [n_alpha <= 16];  X [n_alpha, T+1, n, B], U_out [n_alpha, T, m, B] or NULL;
 *   index [<= B] (i32) + count [1] (i32, device) or both NULL: compacted work list of trajectory numbers (the one
 *   pg_ilqr_select writes) -- only those trajectories are rolled out, other slots of X / U_out are left untouched.
 *   The costs of the candidates are evaluated from X / U_out by the problem library (pg_traj_cost). */
int pg_mlp_forward(int precision, const pg_mlp_dims_t* dims, const void* packed, int64_t B, int32_t T, double dt,
                   int32_t n_alpha, const double* alphas, const double* x0, const double* U, const double* KK,
                   const double* kk, const double* xbar, const double* ubar, const double* ulim, double* X, double* U_out,
                   const int32_t* index, const int32_t* count, void* stream);

/* pg_mlp_forward on the tensor-core kernel (PG_MLP_BF16 only) with a load-balanced schedule: one cluster per pair of SMs,
 * each taking an equal share of the (128-row tile, control interval) sequence instead of whole tiles, so that a launch
 * whose tiles do not fill a whole number of waves (45,056 rows = 2.38 waves) loses nothing to the partial wave.  A tile
 * that is split between two clusters is handed over through X itself plus one progress word per tile in `workspace`
 * (device memory, pg_mlp_forward_workspace_bytes(B, n_alpha) bytes, cleared by the call; it must not be shared by calls
 * that may run concurrently).  Results are bit-identical to pg_mlp_forward. */
int64_t pg_mlp_forward_workspace_bytes(int64_t B, int32_t n_alpha);
int pg_mlp_forward_balanced(int precision, const pg_mlp_dims_t* dims, const void* packed, int64_t B, int32_t T, double dt,
                            int32_t n_alpha, const double* alphas, const double* x0, const double* U, const double* KK,
                            const double* kk, const double* xbar, const double* ubar, const double* ulim, double* X,
                            double* U_out, const int32_t* index, const int32_t* count, void* workspace,
                            int64_t workspace_bytes, void* stream);

/* Continuous-time Jacobians of MBRL.ode along a trajectory, for iLQR.linearization (ilqr.py:518-533):
 *   A [T, n, n, B], Bm [T, n, m, B] at (xbar_k, ubar_k), k < T;  xbar [T+1, n, B], ubar [T, m, B].
 *   PG_MLP_LIN_FD       central differences with eps = 1e-3 through 2(n+m) network evaluations -- what the reference does
 *                       (helpers.system_linearization, pygent/helpers.py:130-148); any precision, meant for PG_MLP_FP32;
 *   PG_MLP_LIN_ANALYTIC the exact derivative of the piecewise-linear network by the chain rule, on the tensor cores
 *                       (PG_MLP_BF16 or PG_MLP_F16X3): one forward GEMM for the ReLU masks + one transposed GEMM per output. */
#define PG_MLP_LIN_ANALYTIC 0
#define PG_MLP_LIN_FD 1
int pg_mlp_linearize(int precision, int lin_mode, const pg_mlp_dims_t* dims, const void* packed, int64_t B, int32_t T,
                     double dt, const double* xbar, const double* ubar, double* A, double* Bm, const int32_t* index,
                     const int32_t* count, void* stream);

/* ---- dynamics training (K6): one optimiser step of `MBRL.training` on the device ------------------------------------------
 * Reference: pygent/algorithms/mbrl.py:450-469 (loop over shuffled batches of 512: `training_data`, nn.MSELoss, zero_grad,
 * backward, step), :471-485 (`training_data`: o_, u and dx = x - x_ standardised with the data-set moments; with sparse_dyn
 * only the velocity half of dx), :71-73 (torch.optim.Adagrad(lr = dyn_lr, weight_decay)).
 *
 * Flat fp32 buffers in torch's parameter order of NNDynamics -- W1 [500, n_obs+m], b1 [500], W2 [500, 500], b2 [500],
 * W3 [dx_dim, 500], b3 [dx_dim]: pg_mlp_train_param_count() floats (255,002 for the cart-pole):
 *     params, state_sum (Adagrad's accumulator), bucket [count + 1] (the gradient and, in the last slot, the loss).
 * Data set on the device, row-major fp32: x_ [N, n], x [N, n], o_ [N, n_obs], u [N, m]; idx [R] (i32) selects the rows of the
 * batch (NULL: rows 0..R-1).  moments [48] fp32: o_mean at 0, o_var at 16, u_mean at 32, u_var at 34, dx_mean at 36, dx_var at 40.
 * workspace: pg_mlp_train_workspace_bytes(max_rows) bytes, 1024-byte aligned, set up by pg_mlp_train_init (which also splits
 * W2 into the fp16 hi/lo operand images of the tensor-core GEMMs; pg_mlp_train_update keeps them current) -- call it again
 * whenever params are changed from outside.
 *
 * pg_mlp_train_grad   : forward + backward of R <= max_rows rows.  bucket <- sum over these rows of d(SSE)/d(theta) * 2 / (n_total *
 *                       dx_dim), bucket[count] <- SSE / (n_total * dx_dim): with n_total = R the gradient and value of nn.MSELoss;
 *                       data parallel, every rank passes its slice of the batch and the global n_total, and ONE sum-all-reduce of
 *                       the bucket (count + 1 floats) yields the full-batch gradient and loss on every rank.
 *                       Layer 2 forward, its input gradient and its weight gradient are 128 x 128-tile tcgen05.mma GEMMs with
 *                       fp16 hi/lo split operands (22 mantissa bits, fp32 accumulation); everything else is fp32 on CUDA cores.
 * pg_mlp_train_update : Adagrad  g += wd p;  s += g^2;  p -= lr g / (sqrt(s) + eps)  on the flat buffers.
 * pg_mlp_train_step   : grad (n_total = R) + update, the single-GPU optimiser step. */
#define PG_MLP_TRAIN_MOMENTS 48
int64_t pg_mlp_train_param_count(const pg_mlp_dims_t* dims);
int64_t pg_mlp_train_workspace_bytes(int32_t max_rows);
int pg_mlp_train_init(const pg_mlp_dims_t* dims, const float* params, void* workspace, int64_t workspace_bytes, int32_t max_rows,
                      void* stream);
int pg_mlp_train_grad(const pg_mlp_dims_t* dims, const float* params, const float* moments, const float* x_, const float* x,
                      const float* o_, const float* u, const int32_t* idx, int32_t R, int64_t n_total, void* workspace,
                      int64_t workspace_bytes, int32_t max_rows, float* bucket, void* stream);
int pg_mlp_train_update(const pg_mlp_dims_t* dims, float* params, float* state_sum, const float* bucket, float lr, float weight_decay,
                        float eps, void* workspace, int64_t workspace_bytes, int32_t max_rows, void* stream);
/* Data-parallel exchange over NVLink peer memory, fused with the optimiser (instead of an NCCL all-reduce + pg_mlp_train_update):
 * every rank allocates an exchange block with pg_mlp_train_p2p_alloc (cudaMalloc + CUDA IPC handle, 64 bytes), the handles are
 * exchanged by the host (e.g. torch.distributed.all_gather), every rank maps its peers' blocks with pg_mlp_train_p2p_open, and
 * passes pg_mlp_train_grad the bucket  (float*)own_block + (step & 1) * npad,  npad = (count + 1) rounded up to 256.
 * pg_mlp_train_update_p2p(step) then publishes "my gradient of `step` is complete" into every peer's flag array and runs ONE
 * kernel that waits for all ranks, sums element i of the world's buckets by peer loads in rank order (bit-identical on every
 * rank) and applies Adagrad to it.  blocks: host array [world] of the mapped block pointers (own block at index rank);
 * steps must be numbered 0, 1, 2, ... identically on all ranks.  loss_out (device, 1 float) receives the summed loss.  A rank
 * that waits more than ~4 s gives up and sets the status word (i32 at byte 2 * npad * 4 + 256 of its own block). */
int64_t pg_mlp_train_p2p_bytes(const pg_mlp_dims_t* dims);
int pg_mlp_train_p2p_alloc(int64_t bytes, void** ptr, void* ipc_handle_out);
int pg_mlp_train_p2p_open(const void* ipc_handle, void** ptr);
int pg_mlp_train_p2p_close(void* ptr, int32_t own);
int pg_mlp_train_update_p2p(const pg_mlp_dims_t* dims, float* params, float* state_sum, void* const* blocks, int32_t world, int32_t rank,
                            int64_t step, float lr, float weight_decay, float eps, float* loss_out, void* workspace,
                            int64_t workspace_bytes, int32_t max_rows, void* stream);
int pg_mlp_train_step(const pg_mlp_dims_t* dims, float* params, float* state_sum, const float* moments, const float* x_,
                      const float* x, const float* o_, const float* u, const int32_t* idx, int32_t R, float lr, float weight_decay,
                      float eps, void* workspace, int64_t workspace_bytes, int32_t max_rows, float* bucket, void* stream);

#ifdef __cplusplus
}
#endif
#endif
