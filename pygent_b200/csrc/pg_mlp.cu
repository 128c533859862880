// Learned-MLP dynamics (NNDynamics / MBRL.ode) on the B200: see include/pygent_b200_mlp.h.
//
//   mlp_fp32_kernel : 64 rows per CTA, layer 2 as a register-tiled fp32 FFMA GEMM (reference precision).
//   mlp_bf16_kernel : 128 rows per CTA = one UMMA M tile.  Per control interval and CTA:
//        layer 1 (K <= 16) on CUDA cores -> H1 as bf16 straight into the 128B-swizzled K-major smem
//        image tcgen05.mma reads (A operand, 128 x 512);
//        layer 2: 64 x tcgen05.mma.cta_group::1.kind::f16 (M128 N256 K16), accumulators in TMEM
//        (128 lanes x 512 fp32 columns = all of it), W2 streamed from L2 in pre-swizzled 32 KB chunks by
//        cp.async.bulk through a 3-stage mbarrier ring (no tensor map needed: the chunks are stored in
//        global memory already in their smem image);
//        epilogue: tcgen05.ld (row = TMEM lane = thread) -> +b2, ReLU, layer 3 (N <= 4) -> Euler update (fp64).
//   Rows are trajectories / line-search candidates / finite-difference points; a CTA carries its rows
//   through all T control intervals, the state never leaves registers.
//   mlp_jac_kernel  : analytic Jacobian of the learned dynamics for iLQR's backward pass, 128 rows per CTA, three
//        GEMM passes on the same tcgen05 pipeline: forward (ReLU masks of both hidden layers), then per network
//        output j one pass G1_j = (W3[j,:] * mask2) x W2 against the TRANSPOSED weight images, epilogue
//        (* mask1) x W1 and the chain rule through standardisation and the [cos, sin] observation.  Replaces
//        the 2(n+m) = 10 network evaluations of the reference's central differences (helpers.py:130-148).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pygent_b200_mlp.h"
#include "pg_math.cuh"
#include "pg_tc.cuh"

#define PG_E_BADARG (-1)
#define PG_E_UNSUPPORTED (-2)

namespace {

using namespace pgtc;

constexpr int H = PG_MLP_HPAD;      // 512
constexpr int HID = PG_MLP_HIDDEN;  // 500
constexpr int MAXIN = PG_MLP_MAX_IN;
constexpr int MAXDX = PG_MLP_MAX_DX;
constexpr int CHUNK_N = 128;                   // n-rows (output columns) of one W2 chunk = N of one MMA
constexpr int CHUNK_BYTES = CHUNK_N * 64 * 2;  // one W2 chunk image: 128 n-rows x 64 k bf16 = 16 KB
constexpr int N_CHUNKS = 32;                   // 8 k-blocks x 4 n-quarters

struct alignas(1024) MlpPacked {
    uint16_t W2b[N_CHUNKS][CHUNK_BYTES / 2];  // bf16, 128B-swizzled K-major images, chunk = quarter * 8 + kb
    uint16_t W2bT[N_CHUNKS][CHUNK_BYTES / 2];  // the same for W2^T (contraction over layer 2's OUTPUT index)
    uint16_t W1img[4][CHUNK_BYTES / 2];        // layer 1 as ONE k-block for the tensor core: per output quarter [128 n x 64 k],
                                               // k = 16 seg + j: seg 0 hi(W1b[n][j]), 1 lo(...), 2 hi(...), 3 zero (see layer1_operand)
    uint16_t W2h[N_CHUNKS][CHUNK_BYTES / 2];   // PG_MLP_F16X3: W2 = hi + lo in fp16, same chunk images as W2b
    uint16_t W2l[N_CHUNKS][CHUNK_BYTES / 2];
    uint16_t W2hT[N_CHUNKS][CHUNK_BYTES / 2];  // ... and of W2^T
    uint16_t W2lT[N_CHUNKS][CHUNK_BYTES / 2];
    float W2t[H][H];                          // [k][n] fp32 (FFMA path)
    float W1[H][MAXIN];                       // [k][input]
    float b1[H];
    float b2[H];
    float W3[H][MAXDX];                       // [k][output]
    float b3[MAXDX];
    float W1b[H][MAXIN];                      // W1 row with b1 in slot n_in (the input vector carries a 1 there)
    float4 ep[H];                             // epilogue constants of hidden unit k: {b2, W3[k][0], W3[k][1], W3[k][2]}
    float W3T[MAXDX][H];                      // [output][k]
    float o_mean[MAXIN], o_var[MAXIN];
    float u_mean[2], u_var[2];
    float dx_mean[MAXDX], dx_var[MAXDX];
};

inline uint16_t f32_to_bf16(float f) {  // round to nearest even
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

inline void f32_to_f16_pair(float f, uint16_t& hi, uint16_t& lo) {  // f ~ hi + lo, both fp16 (round to nearest even)
    const float c = f > 65000.0f ? 65000.0f : (f < -65000.0f ? -65000.0f : f);
    const __half h = __float2half_rn(c);
    const __half l = __float2half_rn(c - __half2float(h));
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(l);
}

inline float bf16_to_f32(uint16_t h) {
    const uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

constexpr int MODE_ODE = 0, MODE_ROLLOUT = 1, MODE_LIN_FD = 2, MODE_LIN_AN = 3;
constexpr int MAX_ALPHAS = 16;

struct MlpArgs {
    pg_mlp_dims_t d;
    const MlpPacked* w;
    int mode;
    int64_t R;   // work items without a work list: points (ODE) or trajectories (ROLLOUT, LIN_*)
    int64_t Bs;  // batch stride of every [..., B] array
    int T;
    double dt;
    int n_alpha;
    double alphas[MAX_ALPHAS];
    const int32_t* index;  // optional compacted work list (trajectory numbers) and its device-side length
    const int32_t* count;
    const double* x0;
    const double* U;
    const double* KK;
    const double* kk;
    const double* xbar;
    const double* ubar;
    int clip;
    double ulim[2];
    double* X;
    double* U_out;
    double* rhs;
    double* A;
    double* Bm;
    int32_t* progress;  // balanced rollouts: per 128-row tile, the number of control intervals whose states are in X
};

// Row r of the launch -> what it computes.  ROLLOUT: r = ia * cnt + w (line-search candidate ia of work item w);
// LIN_*: r = k * cnt + w (control interval k of work item w); cnt = *count or R; trajectory b = index[w] or w.
struct Row {
    bool valid;
    int64_t b;
    int ia, k;
    double alpha;
};

__device__ __forceinline__ int64_t work_count(const MlpArgs& a) { return a.count ? (int64_t)*a.count : a.R; }

__device__ __forceinline__ int64_t total_rows(const MlpArgs& a) {
    const int64_t cnt = work_count(a);
    if (a.mode == MODE_ROLLOUT) return cnt * a.n_alpha;
    if (a.mode == MODE_LIN_FD || a.mode == MODE_LIN_AN) return cnt * a.T;
    return cnt;
}

__device__ __forceinline__ Row decode_row(const MlpArgs& a, int64_t r) {
    Row row;
    const int64_t cnt = work_count(a);
    row.valid = r < total_rows(a);
    row.ia = row.k = 0;
    row.alpha = 1.0;
    int64_t w = r;
    if (row.valid && cnt > 0) {
        if (a.mode == MODE_ROLLOUT) {
            row.ia = (int)(r / cnt);
            w = r - (int64_t)row.ia * cnt;
            row.alpha = a.alphas[row.ia];
        } else if (a.mode != MODE_ODE) {
            row.k = (int)(r / cnt);
            w = r - (int64_t)row.k * cnt;
        }
    }
    row.b = (row.valid && a.index) ? (int64_t)a.index[w] : w;
    return row;
}

__device__ __forceinline__ int steps_of(const MlpArgs& a) {
    if (a.mode == MODE_ROLLOUT) return a.T;
    if (a.mode == MODE_LIN_FD) return 2 * (a.d.n + a.d.m);  // x_i +- eps, u_j +- eps (helpers.py:139-146)
    return 1;
}

// ---- per-row pieces shared by the kernels -----------------------------------------------------------------
__device__ __forceinline__ void control_of_row(const MlpArgs& a, const Row& row, int k, const double (&x)[8], double (&u)[2]) {
    const int n = a.d.n, m = a.d.m;
    const int64_t B = a.Bs, b = row.b;
    for (int j = 0; j < m; j++) {
        double v = 0.0;
        if (a.KK) {  // u = K (x - xbar) + alpha k + ubar (ilqr.py:206)
            double acc = 0.0;
            for (int i = 0; i < n; i++) acc += a.KK[(((int64_t)k * m + j) * n + i) * B + b] * (x[i] - a.xbar[((int64_t)k * n + i) * B + b]);
            v = acc + row.alpha * a.kk[((int64_t)k * m + j) * B + b];
            v = v + a.ubar[((int64_t)k * m + j) * B + b];
        } else if (a.U) {
            v = a.U[((int64_t)k * m + j) * B + b];
        }
        if (a.clip) v = fmin(fmax(v, -a.ulim[j]), a.ulim[j]);
        u[j] = v;
    }
}

// network input: observation (angles -> [cos, sin], fp64) cast to fp32, then standardised (nn_models.py:193-198)
__device__ __forceinline__ void network_input(const MlpArgs& a, const pg::MathCtx<double>& mc, const double (&x)[8],
                                              const double (&u)[2], float (&in)[MAXIN]) {
    const MlpPacked* w = a.w;
    int c = 0;
    for (int i = 0; i < a.d.n; i++) {
        if (a.d.is_angle[i]) {
            double s, co;
            mc.sincos(x[i], s, co);
            in[c] = ((float)co - w->o_mean[c]) / w->o_var[c];
            c++;
            in[c] = ((float)s - w->o_mean[c]) / w->o_var[c];
            c++;
        } else {
            in[c] = ((float)x[i] - w->o_mean[c]) / w->o_var[c];
            c++;
        }
    }
    for (int j = 0; j < a.d.m; j++) {
        in[c] = ((float)u[j] - w->u_mean[j]) / w->u_var[j];
        c++;
    }
    in[c++] = 1.0f;  // multiplies the bias column of W1b (zero in W1)
    for (; c < MAXIN; c++) in[c] = 0.0f;
}

__device__ __forceinline__ void load_row(const MlpArgs& a, const Row& row, double (&x)[8], double (&u)[2]) {
    for (int i = 0; i < 8; i++) x[i] = 0.0;
    u[0] = u[1] = 0.0;
    if (!row.valid) return;
    const int n = a.d.n, m = a.d.m;
    const int64_t B = a.Bs, b = row.b;
    if (a.mode == MODE_ROLLOUT) {
        for (int i = 0; i < n; i++) {
            x[i] = a.x0[(int64_t)i * B + b];
            a.X[(((int64_t)row.ia * (a.T + 1)) * n + i) * B + b] = x[i];
        }
    } else if (a.mode == MODE_ODE) {
        for (int i = 0; i < n; i++) x[i] = a.x0[(int64_t)i * B + b];
        for (int j = 0; j < m; j++) u[j] = a.U[(int64_t)j * B + b];
    } else {  // linearisation point (xbar_k, ubar_k)
        for (int i = 0; i < n; i++) x[i] = a.xbar[((int64_t)row.k * n + i) * B + b];
        for (int j = 0; j < m; j++) u[j] = a.ubar[((int64_t)row.k * m + j) * B + b];
    }
}

constexpr double FD_EPS = 1e-3;  // helpers.system_linearization (pygent/helpers.py:130)

// Before the network evaluation of step s: the point (xe, ue) it is evaluated at.
__device__ __forceinline__ void pre_step(const MlpArgs& a, const Row& row, int s, const double (&x)[8], double (&u)[2],
                                         double (&xe)[8], double (&ue)[2]) {
    if (a.mode == MODE_ROLLOUT && row.valid) control_of_row(a, row, s, x, u);
    for (int i = 0; i < 8; i++) xe[i] = x[i];
    ue[0] = u[0];
    ue[1] = u[1];
    if (a.mode == MODE_LIN_FD) {
        const int j = s >> 1;
        const double e = (s & 1) ? -FD_EPS : FD_EPS;
        if (j < a.d.n) xe[j] = x[j] + e;
        else ue[j - a.d.n] = u[j - a.d.n] + e;
    }
}

// After it: rhs = [x[n/2:], float32(1/dt) * (y * dxVar + dxMean)] (mbrl.py:126-128, nn_models.py:200-201), then the
// Euler update (fast_step, environments.py:313-315), the plain right-hand side, or one central-difference column.
__device__ __forceinline__ void post_step(const MlpArgs& a, const Row& row, int s, const float (&y)[MAXDX], double (&x)[8],
                                          const double (&u)[2], const double (&xe)[8], double (&fplus)[8]) {
    const int n = a.d.n, nh = a.d.n / 2, m = a.d.m;
    const int64_t B = a.Bs, b = row.b;
    const MlpPacked* w = a.w;
    const float inv_dt = (float)(1.0 / a.dt);
    double rhs[8];
    for (int i = 0; i < nh; i++) rhs[i] = xe[nh + i];
    for (int j = 0; j < nh; j++) {
        const float dv = (y[j] + w->b3[j]) * w->dx_var[j] + w->dx_mean[j];
        rhs[nh + j] = (double)(inv_dt * dv);
    }
    if (a.mode == MODE_ODE) {
        if (row.valid)
            for (int i = 0; i < n; i++) a.rhs[(int64_t)i * B + b] = rhs[i];
    } else if (a.mode == MODE_ROLLOUT) {
        for (int i = 0; i < n; i++) x[i] = x[i] + a.dt * rhs[i];
        if (row.valid) {
            for (int i = 0; i < n; i++) a.X[(((int64_t)row.ia * (a.T + 1) + s + 1) * n + i) * B + b] = x[i];
            if (a.U_out)
                for (int j = 0; j < m; j++) a.U_out[(((int64_t)row.ia * a.T + s) * m + j) * B + b] = u[j];
        }
    } else {  // MODE_LIN_FD: (f(p + e) - f(p - e)) / (2 eps)
        if ((s & 1) == 0) {
            for (int i = 0; i < n; i++) fplus[i] = rhs[i];
        } else if (row.valid) {
            const int j = s >> 1;
            for (int i = 0; i < n; i++) {
                const double col = (fplus[i] - rhs[i]) / (2 * FD_EPS);
                if (j < n) a.A[(((int64_t)row.k * n + i) * n + j) * B + b] = col;
                else a.Bm[(((int64_t)row.k * n + i) * m + (j - n)) * B + b] = col;
            }
        }
    }
}

// ---- fp32 FFMA kernel --------------------------------------------------------------------------------------
constexpr int F_ROWS = 64, F_HSTR = 513;

__global__ void __launch_bounds__(256) mlp_fp32_kernel(const MlpArgs a) {
    extern __shared__ float fsm[];
    float* H1s = fsm;                        // [64][513]
    float* ins = H1s + F_ROWS * F_HSTR;      // [64][16]
    float* ys = ins + F_ROWS * MAXIN;        // [64][4]
    const MlpPacked* w = a.w;
    const int t = threadIdx.x;
    if ((int64_t)blockIdx.x * F_ROWS >= total_rows(a)) return;  // compacted work list: surplus CTAs
    const bool owner = t < F_ROWS;
    Row row = decode_row(a, (int64_t)blockIdx.x * F_ROWS + (owner ? t : 0));
    row.valid = row.valid && owner;
    const pg::MathCtx<double> mc;
    double x[8], u[2], xe[8], ue[2], fplus[8];
    if (owner) load_row(a, row, x, u);
    const int steps = steps_of(a);
    for (int k = 0; k < steps; k++) {
        if (owner) {
            pre_step(a, row, k, x, u, xe, ue);
            float in[MAXIN];
            network_input(a, mc, xe, ue, in);
#pragma unroll
            for (int j = 0; j < MAXIN; j++) ins[t * MAXIN + j] = in[j];
        }
        __syncthreads();
        {  // layer 1: thread -> (row, quarter of the hidden units)
            const int row = t & 63, q = t >> 6;
            float in[MAXIN];
#pragma unroll
            for (int j = 0; j < MAXIN; j++) in[j] = ins[row * MAXIN + j];
            for (int kk = q * 128; kk < q * 128 + 128; kk++) {
                float acc = w->b1[kk];
#pragma unroll
                for (int j = 0; j < MAXIN; j++) acc = fmaf(w->W1[kk][j], in[j], acc);
                H1s[row * F_HSTR + kk] = kk < HID ? fmaxf(acc, 0.0f) : 0.0f;
            }
        }
        __syncthreads();
        {  // layers 2 + 3: thread -> 4 rows x 4 columns of each 64-column panel
            const int ty = t >> 4, tx = t & 15;
            float yp[4][MAXDX];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < MAXDX; j++) yp[i][j] = 0.0f;
            for (int p = 0; p < H / 64; p++) {
                const int n0 = p * 64 + tx * 4;
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int c = 0; c < 4; c++) acc[i][c] = 0.0f;
                const float* h = H1s + (ty * 4) * F_HSTR;
#pragma unroll 4
                for (int kk = 0; kk < HID; kk++) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(&w->W2t[kk][n0]));
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const float av = h[i * F_HSTR + kk];
                        acc[i][0] = fmaf(av, b.x, acc[i][0]);
                        acc[i][1] = fmaf(av, b.y, acc[i][1]);
                        acc[i][2] = fmaf(av, b.z, acc[i][2]);
                        acc[i][3] = fmaf(av, b.w, acc[i][3]);
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const float bb = w->b2[n0 + c];
                    const float4 w3 = *reinterpret_cast<const float4*>(&w->W3[n0 + c][0]);
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const float h2 = fmaxf(acc[i][c] + bb, 0.0f);
                        yp[i][0] = fmaf(h2, w3.x, yp[i][0]);
                        yp[i][1] = fmaf(h2, w3.y, yp[i][1]);
                        yp[i][2] = fmaf(h2, w3.z, yp[i][2]);
                        yp[i][3] = fmaf(h2, w3.w, yp[i][3]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < MAXDX; j++) {
                    float v = yp[i][j];
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    if (tx == 0) ys[(ty * 4 + i) * MAXDX + j] = v;
                }
        }
        __syncthreads();
        if (owner) {
            float y[MAXDX];
#pragma unroll
            for (int j = 0; j < MAXDX; j++) y[j] = ys[t * MAXDX + j];
            post_step(a, row, k, y, x, u, xe, fplus);
        }
        __syncthreads();
    }
}

// ---- tcgen05 helpers: the generic ones (mbarriers, bulk copies, descriptors, tcgen05.ld) live in pg_tc.cuh ----
// ---- thread-block cluster of CL CTAs sharing one weight stream ------------------------------------------------------
constexpr int CL = 2;
constexpr uint16_t CL_MASK = (1u << CL) - 1;
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// bulk copy global -> the same smem offset of every CTA in the mask; each destination's mbarrier (same offset) gets the bytes
__device__ __forceinline__ void bulk_g2s_multicast(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// D (fp32) = A (bf16, K-major) x B (bf16, K-major), M = 128, N = CHUNK_N
constexpr uint32_t UMMA_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)CHUNK_N >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(UMMA_IDESC), "r"(accumulate)
        : "memory");
}
// The CTA pair as ONE MMA unit (cta_group::2): D [256 x 256] (fp32; rows 0-127 in the leader's TMEM, 128-255 in the
// peer's) += A [256 x 16] (each CTA's own 128 rows, same smem offset) x B [256 x 16] (output columns 0-127 from the
// leader's smem, 128-255 from the peer's).  Issued by the leader only.
constexpr uint32_t UMMA_IDESC_2SM = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(UMMA_IDESC_2SM), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// completion of this thread's earlier MMAs arrives on the barrier at the same offset in every CTA of the mask
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
constexpr int T_ROWS = 128, T_THREADS = 512, T_SPLIT = T_THREADS / T_ROWS;
constexpr int T_STAGES = 4;  // weight ring of the rollout kernel (231,680 B of smem in total)
constexpr int J_STAGES = 3;  // ... of the Jacobian kernel, which also holds the ReLU masks
constexpr int T_A_BYTES = T_ROWS * H * 2;  // 128 KB: 8 k-blocks of [128 rows x 64 k] bf16
constexpr int INS_STRIDE = MAXIN + 4;      // floats per row of the input tile (padded: conflict-free 128-bit reads)
constexpr int T_INS_BYTES = T_ROWS * INS_STRIDE * 4;  // network inputs of the 128 rows; reused for the partial layer-3 sums
constexpr int T_EP_BYTES = H * 16;                    // epilogue constants {b2, W3[.][0..2]} per hidden unit
constexpr int W1S_FLOATS = 8;                         // layer-1 weight rows kept in smem when inputs + 1 <= 8 (else read from L2)
constexpr int T_W1_BYTES = H * W1S_FLOATS * 4;
constexpr int T_CONST_BYTES = T_EP_BYTES + T_W1_BYTES;
constexpr int T_SMEM = T_A_BYTES + T_STAGES * CHUNK_BYTES + T_INS_BYTES + T_CONST_BYTES + 256;
constexpr int PRODUCER_THREAD = 384, ISSUER_THREAD = 416;  // in the warps of accumulator quarter 3, which finishes last

// One 128 x 512 x 512 GEMM of the CTA: D (TMEM, fp32) = A (smem image, bf16) x B (32 chunk images of 16 KB streamed
// from L2 through a 3-stage ring).  The two CTAs of a cluster share the stream: CTA r fetches the chunks with
// i % 2 == r and multicasts each into BOTH rings (with every SM streaming all 512 KB per control interval the kernel was
// L2-bandwidth bound: ~12 us per GEMM against 7.2 us of MMA time); a stage is released to the producers only when the
// MMAs of both CTAs have read it (multicast tcgen05.commit on barriers initialised to CL arrivals).  Chunks come quarter-major -- all 8 k-blocks of output columns 0..127, then 128..255,
// ... -- and each accumulator quarter has its own completion barrier, so the threads that own quarter q run their
// epilogue while the tensor core works on q+1..3.  One thread produces the ring, one issues the 128 MMAs (M128 N128
// K16); both sit in the warps of quarter 3.  `img` selects the weight images (W2b: forward, W2bT: transposed).
#define PG_TMEM_ALLOC_1SM "tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;"
#define PG_TMEM_ALLOC_2SM "tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;"
constexpr bool TWO_SM = true;  // pair the two CTAs of a cluster into one M = 256 MMA unit (false: one M = 128 unit per CTA)

struct GemmPipe {
    uint8_t* As;
    uint8_t* Bs;
    uint64_t* full;
    uint64_t* empty;
    uint64_t* accum;   // [4]
    uint64_t* pfull;   // leader: the peer's copy of stage s has landed (relayed by the peer)
    uint64_t* a_ready; // leader: the peer's A image is written
    uint32_t tmem_base;
    uint32_t chunk_no;  // producer and issuer count chunks the same way
    uint32_t passes;
    uint32_t prefetched;  // producer thread: chunks of THIS pass already requested at the end of the previous one
};

// `next_img`: the weight images of the following pass (nullptr after the last one).  When its own chunks are out the
// producer requests the first STAGES chunks of the next pass right away: the stages free up as this pass's MMAs retire,
// so the next GEMM starts with a full ring instead of paying the L2 latency again while the tensor core idles.
// KB: k-blocks of 64 this pass contracts over (8: a hidden layer of 512; 1: layer 1, whose operand is `a_img`);
// NEXT_KB: the same for the pass that follows.
template <int STAGES, int KB = 8, int NEXT_KB = 8>
__device__ __forceinline__ void gemm_issue(GemmPipe& p, const uint16_t (*img)[CHUNK_BYTES / 2], int t,
                                           const uint16_t (*next_img)[CHUNK_BYTES / 2] = nullptr, const uint8_t* a_img = nullptr) {
    // the control lanes' 31 warp siblings park at __syncwarp() instead of spinning on the completion barrier: as
    // divergent paths of the same warp they would take every other issue slot from the lane that feeds the tensor core
    constexpr int PER_PASS = TWO_SM ? 4 * KB / 2 : 4 * KB;  // chunks this CTA loads per GEMM
    constexpr int NEXT_PER_PASS = TWO_SM ? 4 * NEXT_KB / 2 : 4 * NEXT_KB;
    constexpr int N_PREFETCH = NEXT_PER_PASS < STAGES ? NEXT_PER_PASS : STAGES;
    const uint8_t* A = a_img ? a_img : p.As;
    const int cw = t >> 5;
    if (cw != (PRODUCER_THREAD >> 5) && cw != (ISSUER_THREAD >> 5)) {
        p.chunk_no += PER_PASS;
        return;
    }
    const uint32_t rank = cluster_ctarank();
    if constexpr (TWO_SM) {
        // Chunk i of a pass: output-column half q2 = i / 8 (256 columns), k-block kb = i % 8.  CTA r holds the 128
        // columns [256 q2 + 128 r, +128) of it: weight image quarter 2 q2 + r.  Nothing is multicast: each CTA streams only
        // the half of W2 its own tensor core multiplies.
        if (t == PRODUCER_THREAD) {
            uint32_t c = p.chunk_no + p.prefetched;
            for (int i = (int)p.prefetched; i < PER_PASS; i++, c++) {
                const uint32_t s = c % STAGES, use = c / STAGES;
                mbar_wait(&p.empty[s], (use & 1) ^ 1);
                mbar_expect_tx(&p.full[s], CHUNK_BYTES);
                bulk_g2s(p.Bs + s * CHUNK_BYTES, &img[(2 * (i / KB) + rank) * KB + (i % KB)][0], CHUNK_BYTES, &p.full[s]);
            }
            p.prefetched = 0;
            if (next_img) {
                for (int i = 0; i < N_PREFETCH; i++, c++) {
                    const uint32_t s = c % STAGES, use = c / STAGES;
                    mbar_wait(&p.empty[s], (use & 1) ^ 1);
                    mbar_expect_tx(&p.full[s], CHUNK_BYTES);
                    bulk_g2s(p.Bs + s * CHUNK_BYTES, &next_img[(2 * (i / NEXT_KB) + rank) * NEXT_KB + (i % NEXT_KB)][0], CHUNK_BYTES,
                             &p.full[s]);
                }
                p.prefetched = N_PREFETCH;
            }
        } else if (t == ISSUER_THREAD && rank != 0) {  // peer: tell the leader when a stage of OUR ring is full
            uint32_t c = p.chunk_no;
            for (int i = 0; i < PER_PASS; i++, c++) {
                const uint32_t s = c % STAGES, use = c / STAGES;
                mbar_wait(&p.full[s], use & 1);
                mbar_arrive_remote(&p.pfull[s], 0);
            }
        } else if (t == ISSUER_THREAD) {  // leader: one thread drives both tensor cores
            mbar_wait(p.a_ready, p.passes & 1);  // the peer's A image (ours is ordered by the CTA barrier before this call)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t c = p.chunk_no;
            for (int i = 0; i < PER_PASS; i++, c++) {
                const uint32_t s = c % STAGES, use = c / STAGES;
                mbar_wait(&p.full[s], use & 1);
                mbar_wait(&p.pfull[s], use & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int q2 = i / KB, kb = i % KB;
                const uint32_t a_addr = smem_u32(A + kb * (T_ROWS * 128));
                const uint32_t b_addr = smem_u32(p.Bs + s * CHUNK_BYTES);
#pragma unroll
                for (int ks = 0; ks < 4; ks++)
                    umma_bf16_2sm(p.tmem_base + q2 * 256, umma_desc_sw128(a_addr + ks * 32), umma_desc_sw128(b_addr + ks * 32),
                                  (kb > 0 || ks > 0) ? 1u : 0u);
                umma_commit_2sm(&p.empty[s], CL_MASK);  // frees the stage in both CTAs
                if (kb == KB - 1) {                       // 256 accumulator columns final in both CTAs
                    umma_commit_2sm(&p.accum[2 * q2], CL_MASK);
                    umma_commit_2sm(&p.accum[2 * q2 + 1], CL_MASK);
                }
            }
        }
    } else {
        if (t == PRODUCER_THREAD) {  // weight chunks, L2 -> smem rings of the cluster
            uint32_t c = p.chunk_no;
            for (int i = 0; i < PER_PASS; i++, c++) {
                const uint32_t s = c % STAGES, use = c / STAGES;
                mbar_wait(&p.empty[s], (use & 1) ^ 1);        // both CTAs are done with the stage's previous chunk
                mbar_expect_tx(&p.full[s], CHUNK_BYTES);      // whoever fetches it, this CTA's ring receives the bytes
                if ((uint32_t)(i % CL) == rank) bulk_g2s_multicast(p.Bs + s * CHUNK_BYTES, &img[i][0], CHUNK_BYTES, &p.full[s], CL_MASK);
            }
        } else if (t == ISSUER_THREAD) {  // one thread drives the tensor core for the whole CTA
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t c = p.chunk_no;
            for (int i = 0; i < PER_PASS; i++, c++) {
                const uint32_t s = c % STAGES, use = c / STAGES;
                mbar_wait(&p.full[s], use & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int quarter = i / KB, kb = i % KB;
                const uint32_t a_addr = smem_u32(A + kb * (T_ROWS * 128));
                const uint32_t b_addr = smem_u32(p.Bs + s * CHUNK_BYTES);
#pragma unroll
                for (int ks = 0; ks < 4; ks++)  // K = 16 per instruction = 32 bytes along the swizzled row
                    umma_bf16(p.tmem_base + quarter * CHUNK_N, umma_desc_sw128(a_addr + ks * 32), umma_desc_sw128(b_addr + ks * 32),
                              (kb > 0 || ks > 0) ? 1u : 0u);
                umma_commit_multicast(&p.empty[s], CL_MASK);  // one of the CL arrivals that free the stage in every CTA
                if (kb == KB - 1) umma_commit(&p.accum[quarter]);  // this quarter's accumulators are final
            }
        }
    }
    __syncwarp();
    p.chunk_no += PER_PASS;
}

// The A image of this CTA is complete (call after fence.proxy.async + the CTA barrier): the peer tells the leader.
__device__ __forceinline__ void gemm_a_ready(GemmPipe& p, int t) {
    if constexpr (TWO_SM) {
        if (t == 0 && cluster_ctarank() != 0) mbar_arrive_remote(p.a_ready, 0);
    }
}

// wait until accumulator columns [128 q, 128 q + 128) of the current pass are readable
__device__ __forceinline__ void gemm_wait(const GemmPipe& p, int q) {
    mbar_wait(&p.accum[q], p.passes & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// per-CTA constants -> shared memory (after the smem carve-out only ~40 KB of L1 remain: left in global memory these
// warp-uniform loads missed half of the time)
__device__ __forceinline__ void stage_constants(const MlpPacked* w, float4* eps, float* w1s, int t) {
    for (int i = t; i < H; i += T_THREADS) eps[i] = w->ep[i];
    if (!w1s) return;
    for (int i = t; i < H * (W1S_FLOATS / 4); i += T_THREADS) {
        const int k = i / (W1S_FLOATS / 4), q = i % (W1S_FLOATS / 4);
        reinterpret_cast<float4*>(w1s)[i] = reinterpret_cast<const float4*>(&w->W1b[k][0])[q];
    }
}

// Layer 1 of the CTA's 128 rows, straight into the bf16 A image tcgen05.mma reads (and, with MASK, the ReLU mask words).
// Element (row, kk) of k-block kb lives at
//   kb * 16 KB + (row / 8) * 1024 + (row % 8) * 128 + (((kk / 8) ^ (row % 8)) * 16) + (kk % 8) * 2
// Thread mapping: a warp owns 128 / R hidden units and 32 R rows, lane l the rows l + 32 i: the weight rows (NQ float4
// each, bias folded in) are warp-uniform 128-bit loads reused for R rows -- 64 loads per thread instead of the ~900
// scalar ones of a one-row-per-thread mapping, which kept this stage L1-wavefront bound.
template <int NQ, int R, bool MASK>
__device__ __forceinline__ void layer1_image(const float* w1, int w1_stride, const float* ins, uint8_t* As, uint32_t* m1, int t) {
    constexpr int SETS = 4 / R;             // row sets of 32 R rows
    constexpr int CHUNKS = (H / 8) / (4 * R);  // 16-byte chunks (8 hidden units) per thread
    const int lane = t & 31, wi = t >> 5;
    const int rs = wi % SETS, up = wi / SETS;
    float in[R][NQ * 4];
#pragma unroll
    for (int i = 0; i < R; i++) {
        const float4* src = reinterpret_cast<const float4*>(ins + (rs * 32 * R + lane + 32 * i) * INS_STRIDE);
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const float4 v = src[q];
            in[i][q * 4 + 0] = v.x;
            in[i][q * 4 + 1] = v.y;
            in[i][q * 4 + 2] = v.z;
            in[i][q * 4 + 3] = v.w;
        }
    }
    uint32_t mw[R];
#pragma unroll
    for (int i = 0; i < R; i++) mw[i] = 0;
    for (int c = 0; c < CHUNKS; c++) {
        const int g = up * CHUNKS + c;  // 8 hidden units g*8 .. g*8+7
        float acc[R][8];
#pragma unroll
        for (int i = 0; i < R; i++)
#pragma unroll
            for (int e = 0; e < 8; e++) acc[i][e] = 0.0f;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const float4* wr = reinterpret_cast<const float4*>(w1 + (g * 8 + e) * w1_stride);
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                const float4 ww = wr[q];
#pragma unroll
                for (int i = 0; i < R; i++) {
                    acc[i][e] = fmaf(ww.x, in[i][q * 4 + 0], acc[i][e]);
                    acc[i][e] = fmaf(ww.y, in[i][q * 4 + 1], acc[i][e]);
                    acc[i][e] = fmaf(ww.z, in[i][q * 4 + 2], acc[i][e]);
                    acc[i][e] = fmaf(ww.w, in[i][q * 4 + 3], acc[i][e]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < R; i++) {
            const int row = rs * 32 * R + lane + 32 * i;
            uint32_t packed[4];
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
                const bool plo = (g * 8 + e) < HID && acc[i][e] > 0.0f, phi = (g * 8 + e + 1) < HID && acc[i][e + 1] > 0.0f;
                const float lo = plo ? acc[i][e] : 0.0f, hi = phi ? acc[i][e + 1] : 0.0f;
                if (MASK) mw[i] |= ((plo ? 1u : 0u) << ((g & 3) * 8 + e)) | ((phi ? 1u : 0u) << ((g & 3) * 8 + e + 1));
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(packed[e >> 1]) : "f"(hi), "f"(lo));
            }
            uint8_t* rowp = As + (row >> 3) * 1024 + (row & 7) * 128;
            const int kb = g >> 3, cidx = (g & 7) ^ (row & 7);
            *reinterpret_cast<uint4*>(rowp + kb * (T_ROWS * 128) + cidx * 16) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            if (MASK && (g & 3) == 3) {  // hidden units 32 (g / 4) .. + 31 complete
                m1[row * (H / 32) + (g >> 2)] = mw[i];
                mw[i] = 0;
            }
        }
    }
}

template <bool MASK>
__device__ __forceinline__ void layer1_dispatch(int nin, const MlpPacked* w, const float* w1s, const float* ins, uint8_t* As, uint32_t* m1, int t) {
    const int nq = (nin + 1 + 3) >> 2;  // inputs + the constant 1, in float4 units
    if (nq <= 1) layer1_image<1, 4, MASK>(w1s, W1S_FLOATS, ins, As, m1, t);
    else if (nq == 2) layer1_image<2, 4, MASK>(w1s, W1S_FLOATS, ins, As, m1, t);
    else if (nq == 3) layer1_image<3, 2, MASK>(&w->W1b[0][0], MAXIN, ins, As, m1, t);
    else layer1_image<4, 2, MASK>(&w->W1b[0][0], MAXIN, ins, As, m1, t);
}


// ---- layer 1 on the tensor core -------------------------------------------------------------------------------------
// K = inputs + 1 <= 16 is far too short for an MMA in bf16 ALONE (the network inputs would lose 16 mantissa bits), but as
// a split product it fits ONE 64-wide k-block:  x w ~ xh wh + xh wl + xl wh  with  x = xh + xl, w = wh + wl  in bf16 --
// relative error ~2^-16, below the bf16 rounding of h1 that layer 2 applies anyway.  A row of the operand image is
//   [xh_0..15 | xh_0..15 | xl_0..15 | 0]  against W1img's  [wh | wl | wh | 0]   (bias: the constant-1 input slot),
// in the same 128-byte-swizzled K-major format as a k-block of the A image, so the pass reuses the layer-2 descriptors.
// This replaces 1,024 FFMA per thread and interval (4.5 us of the 14.5 us interval) by 8 MMAs and one TMEM read-out.
__device__ __forceinline__ void layer1_operand(const float (&in)[MAXIN], uint8_t* A1, int row) {
    uint32_t hp[MAXIN / 2], lp[MAXIN / 2];
#pragma unroll
    for (int j = 0; j < MAXIN / 2; j++) {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hp[j]) : "f"(in[2 * j + 1]), "f"(in[2 * j]));
        const float h0 = __uint_as_float(hp[j] << 16), h1 = __uint_as_float(hp[j] & 0xffff0000u);
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lp[j]) : "f"(in[2 * j + 1] - h1), "f"(in[2 * j] - h0));
    }
    uint8_t* rowp = A1 + (row >> 3) * 1024 + (row & 7) * 128;
    const int x = row & 7;
    const uint4 h_lo = make_uint4(hp[0], hp[1], hp[2], hp[3]), h_hi = make_uint4(hp[4], hp[5], hp[6], hp[7]);
    *reinterpret_cast<uint4*>(rowp + ((0 ^ x) << 4)) = h_lo;
    *reinterpret_cast<uint4*>(rowp + ((1 ^ x) << 4)) = h_hi;
    *reinterpret_cast<uint4*>(rowp + ((2 ^ x) << 4)) = h_lo;
    *reinterpret_cast<uint4*>(rowp + ((3 ^ x) << 4)) = h_hi;
    *reinterpret_cast<uint4*>(rowp + ((4 ^ x) << 4)) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
    *reinterpret_cast<uint4*>(rowp + ((5 ^ x) << 4)) = make_uint4(lp[4], lp[5], lp[6], lp[7]);
    *reinterpret_cast<uint4*>(rowp + ((6 ^ x) << 4)) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(rowp + ((7 ^ x) << 4)) = make_uint4(0u, 0u, 0u, 0u);
}

// h1 = relu(layer-1 accumulators) of this thread's row and column quarter: TMEM -> bf16 -> the A image of layer 2
__device__ __forceinline__ void hidden_from_tmem(uint32_t tmem_base, int warp, int part, int row, uint8_t* As) {
    uint8_t* rowp = As + (row >> 3) * 1024 + (row & 7) * 128;
    for (int cb = part * (H / 32 / T_SPLIT); cb < (part + 1) * (H / 32 / T_SPLIT); cb++) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + cb * 32, v);
#pragma unroll
        for (int c8 = 0; c8 < 4; c8++) {
            uint32_t packed[4];
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
                const float lo = fmaxf(__uint_as_float(v[c8 * 8 + e]), 0.0f), hi = fmaxf(__uint_as_float(v[c8 * 8 + e + 1]), 0.0f);
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(packed[e >> 1]) : "f"(hi), "f"(lo));
            }
            const int nn0 = cb * 32 + c8 * 8, kb = nn0 >> 6, cidx = ((nn0 & 63) >> 3) ^ (row & 7);
            *reinterpret_cast<uint4*>(rowp + kb * (T_ROWS * 128) + cidx * 16) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
    }
}

// 512 threads per CTA: thread t works on row (t & 127); the four threads of a row split the hidden units
// (layer 1) and the accumulator columns (epilogue) four ways.  Warp w reads TMEM lanes 32 (w % 4) .. +31, which
// is exactly the row quarter (t & 127) >> 5 of its threads.  Threads 0..127 own the fp64 state of their row.
// ---- hand-over of a tile between two clusters of a balanced rollout --------------------------------------------
__device__ __forceinline__ void publish_progress(int32_t* p, int k) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(k) : "memory");
}
__device__ __forceinline__ void wait_progress(const int32_t* p, int k) {
    unsigned long long t0 = 0;
    for (int spins = 0;; spins++) {
        int v;
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        if (v >= k) return;
        __nanosleep(256);
        if ((spins & 1023) == 1023) {  // the producer runs its piece first, so this never waits long: trap rather than hang
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();
        }
    }
}
// states of interval k of a rollout row, as post_step of interval k - 1 wrote them (by another SM: read through L2)
__device__ __forceinline__ void reload_row(const MlpArgs& a, const Row& row, int k, double (&x)[8], double (&u)[2]) {
    for (int i = 0; i < 8; i++) x[i] = 0.0;
    u[0] = u[1] = 0.0;
    if (!row.valid) return;
    const int n = a.d.n;
    const int64_t B = a.Bs, b = row.b;
    for (int i = 0; i < n; i++) x[i] = __ldcg(&a.X[(((int64_t)row.ia * (a.T + 1) + k) * n + i) * B + b]);
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(T_THREADS, 1) mlp_bf16_kernel(const MlpArgs a) {
    extern __shared__ __align__(1024) uint8_t tsm[];
    uint8_t* As = tsm;
    uint8_t* Bs = tsm + T_A_BYTES;
    float* ins = reinterpret_cast<float*>(tsm + T_A_BYTES + T_STAGES * CHUNK_BYTES);
    float* yp = ins;  // partial layer-3 sums [T_SPLIT][T_ROWS][MAXDX]: the input tile is dead once layer 1 has run
    float4* eps = reinterpret_cast<float4*>(tsm + T_A_BYTES + T_STAGES * CHUNK_BYTES + T_INS_BYTES);
    uint8_t* A1 = reinterpret_cast<uint8_t*>(eps + H);  // layer-1 operand image, one k-block (16 KB, 1024-byte aligned)
    static_assert((T_A_BYTES + T_STAGES * CHUNK_BYTES + T_INS_BYTES + T_EP_BYTES) % 1024 == 0 && T_W1_BYTES >= CHUNK_BYTES, "A1 placement");
    uint64_t* full = reinterpret_cast<uint64_t*>(A1 + T_W1_BYTES);
    uint64_t* empty = full + T_STAGES;
    uint64_t* accum = empty + T_STAGES;
    uint64_t* pfull = accum + 4;
    uint64_t* a_ready = pfull + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);
    const MlpPacked* w = a.w;
    const int t = threadIdx.x, warp = t >> 5, row = t & (T_ROWS - 1), part = t >> 7;
    // The work of a launch is the sequence (tile pair p, control interval k), p-major.  Plain launches give every cluster
    // one whole pair.  Balanced rollouts (a.progress) run one cluster per pair of SMs and cut the sequence into equal
    // ranges, so that no SM idles through a partial last wave: a range covers the tail of one pair, whole pairs, and the
    // head of another.  A cluster walks its range BACKWARDS -- the head [0, k) of its last pair first, the tail [k, T)
    // of its first pair last -- so the cluster that continues a pair finds the states of interval k in X long before it
    // gets there (it has a range of at least T intervals to walk first); `progress` carries the hand-over.
    const int steps = steps_of(a);
    const int64_t pairs = (total_rows(a) + CL * T_ROWS - 1) / (CL * T_ROWS), slot = blockIdx.x / CL;
    int64_t lo, hi;
    if (a.progress) {
        const int64_t slots = min((int64_t)(gridDim.x / CL), pairs);
        if (slot >= slots) return;
        lo = slot * (pairs * steps) / slots;
        hi = (slot + 1) * (pairs * steps) / slots;
    } else {
        if (slot >= pairs) return;  // compacted work list: surplus clusters
        lo = slot * steps;
        hi = lo + steps;
    }
    const bool owner = t < T_ROWS;
    const int nin = a.d.n_obs + a.d.m;
    const pg::MathCtx<double> mc;

    if (t == 0) {
        for (int s = 0; s < T_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], TWO_SM ? 1 : CL);
            mbar_init(&pfull[s], 1);
        }
        for (int q = 0; q < 4; q++) mbar_init(&accum[q], 1);
        mbar_init(a_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // the whole accumulator file of the SM: 128 lanes x 512 fp32 columns
        if constexpr (TWO_SM) {
            asm volatile(PG_TMEM_ALLOC_2SM ::"r"(smem_u32(tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile(PG_TMEM_ALLOC_1SM ::"r"(smem_u32(tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    stage_constants(w, eps, nullptr, t);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();  // the peer's barriers are initialised before anything is multicast to them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    double x[8], u[2], xe[8], ue[2], fplus[8];
    GemmPipe pipe = {As, Bs, full, empty, accum, pfull, a_ready, tmem_base, 0u, 0u, 0u};
    (void)nin;
    if (steps == 0 && owner) {  // T = 0: only the initial states
        Row r0 = decode_row(a, (int64_t)blockIdx.x * T_ROWS + row);
        load_row(a, r0, x, u);
    }
    for (int64_t pos = hi; pos > lo;) {
    const int64_t pair = (pos - 1) / steps, base = pair * steps;
    const int k_begin = (int)(max(lo, base) - base), k_end = (int)(pos - base);
    pos = base + k_begin;
    const bool last_piece = pos <= lo;
    const int64_t tile = pair * CL + (blockIdx.x % CL);
    Row rw = decode_row(a, tile * T_ROWS + row);
    rw.valid = rw.valid && owner;
    if (k_begin > 0) {  // continue where another cluster left this tile
        if (t == 0) wait_progress(&a.progress[tile], k_begin);
        __syncthreads();
    }
    if (owner) {
        if (k_begin == 0) load_row(a, rw, x, u);
        else reload_row(a, rw, k_begin, x, u);
    }
    for (int k = k_begin; k < k_end; k++) {
        if (owner) {
            pre_step(a, rw, k, x, u, xe, ue);
            float in[MAXIN];
            network_input(a, mc, xe, ue, in);
            layer1_operand(in, A1, row);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA unit
        __syncthreads();
        // layer 1: one k-block against W1img, accumulators = pre-activations of all 512 hidden units
        gemm_a_ready(pipe, t);
        gemm_issue<T_STAGES, 1, 8>(pipe, w->W1img, t, w->W2b, A1);
        gemm_wait(pipe, part);
        hidden_from_tmem(tmem_base, warp, part, row, As);
        pipe.passes++;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();  // A image complete, every lane has read its layer-1 accumulators
        gemm_a_ready(pipe, t);
        if (k + 1 < k_end || !last_piece) gemm_issue<T_STAGES, 8, 1>(pipe, w->W2b, t, w->W1img);
        else gemm_issue<T_STAGES, 8, 1>(pipe, w->W2b, t, nullptr);
        // epilogue: TMEM lane = row; this thread's quarter of the 512 columns, as soon as that quarter is final:
        // +b2, ReLU, layer 3
        gemm_wait(pipe, part);
        {
            float y[MAXDX] = {0.0f, 0.0f, 0.0f, 0.0f};
            for (int cb = part * (H / 32 / T_SPLIT); cb < (part + 1) * (H / 32 / T_SPLIT); cb++) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + cb * 32, v);
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    const int nn = cb * 32 + c;
                    const float4 e4 = eps[nn];  // {b2, W3[nn][0..2]}: one 128-bit smem broadcast per column
                    const float h2 = fmaxf(__uint_as_float(v[c]) + e4.x, 0.0f);
                    y[0] = fmaf(h2, e4.y, y[0]);
                    y[1] = fmaf(h2, e4.z, y[1]);
                    y[2] = fmaf(h2, e4.w, y[2]);
                    if (a.d.dx_dim > 3) y[3] = fmaf(h2, w->W3[nn][3], y[3]);
                }
            }
            *reinterpret_cast<float4*>(&yp[(part * T_ROWS + row) * MAXDX]) = make_float4(y[0], y[1], y[2], y[3]);
        }
        pipe.passes++;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();  // every lane has read its accumulators before the next interval's MMAs overwrite them
        if (owner) {
            float y[MAXDX];
#pragma unroll
            for (int j = 0; j < MAXDX; j++) {
                float v = 0.0f;
#pragma unroll
                for (int q = 0; q < T_SPLIT; q++) v += yp[(q * T_ROWS + row) * MAXDX + j];
                y[j] = v;
            }
            post_step(a, rw, k, y, x, u, xe, fplus);
        }
    }
    if (a.progress && k_end < steps) {  // the head of a pair: publish that X holds the states up to interval k_end
        __threadfence();
        __syncthreads();
        if (t == 0) publish_progress(&a.progress[tile], k_end);
    }
    }
    __syncthreads();
    cluster_sync();  // no CTA leaves while its peer may still signal its barriers
    if (warp == 0) {
        if constexpr (TWO_SM) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ---- analytic Jacobian on the tensor cores ---------------------------------------------------------------------
// rhs = [x[n/2:], (1/dt) (dxVar * y(in(x, u)) + dxMean)],  y = W3 relu(W2 relu(W1 in + b1) + b2) + b3, so for output j
//   dy_j/din = ((W3[j,:] * m2) W2 * m1) W1,   m1 = [h1 > 0], m2 = [h2 > 0]
// and d in/d(x, u) is diagonal except for the angle states, which enter as [cos, sin].  Three GEMM passes per 128-row
// tile: forward (masks only), then one pass per output against the transposed weight images.  The accumulators and the
// A image are reused between passes; after a pass has completed the A image doubles as scratch for the partial sums.
constexpr int J_MASK_WORDS = H / 32;                          // 16 mask words per row and layer
constexpr int J_MASK_BYTES = 2 * T_ROWS * J_MASK_WORDS * 4;   // both hidden layers
constexpr int J_SMEM = T_A_BYTES + J_STAGES * CHUNK_BYTES + T_INS_BYTES + T_CONST_BYTES + J_MASK_BYTES + 256;

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(T_THREADS, 1) mlp_jac_kernel(const MlpArgs a) {
    extern __shared__ __align__(1024) uint8_t tsm[];
    uint8_t* As = tsm;
    uint8_t* Bs = tsm + T_A_BYTES;
    float* ins = reinterpret_cast<float*>(tsm + T_A_BYTES + J_STAGES * CHUNK_BYTES);
    float4* eps = reinterpret_cast<float4*>(tsm + T_A_BYTES + J_STAGES * CHUNK_BYTES + T_INS_BYTES);
    float* w1s = reinterpret_cast<float*>(eps + H);
    uint32_t* m1 = reinterpret_cast<uint32_t*>(w1s + H * W1S_FLOATS);
    uint32_t* m2 = m1 + T_ROWS * J_MASK_WORDS;
    uint64_t* full = reinterpret_cast<uint64_t*>(m2 + T_ROWS * J_MASK_WORDS);
    uint64_t* empty = full + J_STAGES;
    uint64_t* accum = empty + J_STAGES;
    uint64_t* pfull = accum + 4;
    uint64_t* a_ready = pfull + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);
    const MlpPacked* w = a.w;
    const int t = threadIdx.x, warp = t >> 5, row = t & (T_ROWS - 1), part = t >> 7;
    if ((int64_t)(blockIdx.x / CL) * CL * T_ROWS >= total_rows(a)) return;  // surplus clusters
    const bool owner = t < T_ROWS;
    Row rw = decode_row(a, (int64_t)blockIdx.x * T_ROWS + row);
    rw.valid = rw.valid && owner;
    const int n = a.d.n, nh = a.d.n / 2, m = a.d.m, nin = a.d.n_obs + a.d.m;
    const pg::MathCtx<double> mc;

    if (t == 0) {
        for (int s = 0; s < J_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], TWO_SM ? 1 : CL);
            mbar_init(&pfull[s], 1);
        }
        for (int q = 0; q < 4; q++) mbar_init(&accum[q], 1);
        mbar_init(a_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if constexpr (TWO_SM) {
            asm volatile(PG_TMEM_ALLOC_2SM ::"r"(smem_u32(tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile(PG_TMEM_ALLOC_1SM ::"r"(smem_u32(tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    stage_constants(w, eps, w1s, t);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();  // the peer's barriers are initialised before anything is multicast to them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    GemmPipe pipe = {As, Bs, full, empty, accum, pfull, a_ready, tmem_base, 0u, 0u, 0u};

    double x[8], u[2];
    if (owner) {
        load_row(a, rw, x, u);
        float in[MAXIN];
        network_input(a, mc, x, u, in);
#pragma unroll
        for (int j = 0; j < MAXIN; j++) ins[row * INS_STRIDE + j] = in[j];
    }
    __syncthreads();
    uint8_t* rowp = As + (row >> 3) * 1024 + (row & 7) * 128;
    constexpr int G0 = H / 8 / T_SPLIT;  // 16 groups of 8 hidden units per thread (passes 1..: one row per thread)
    layer1_dispatch<true>(nin, w, w1s, ins, As, m1, t);  // pass 0, A image: layer 1 + its ReLU mask
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    gemm_a_ready(pipe, t);
    gemm_issue<J_STAGES>(pipe, w->W2b, t, w->W2bT);
    gemm_wait(pipe, part);
    // mask of layer 2: [acc + b2 > 0]
    for (int cb = part * (H / 32 / T_SPLIT); cb < (part + 1) * (H / 32 / T_SPLIT); cb++) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + cb * 32, v);
        uint32_t mw = 0;
#pragma unroll
        for (int c = 0; c < 32; c++) {
            const int nn = cb * 32 + c;
            mw |= ((nn < HID && __uint_as_float(v[c]) + eps[nn].x > 0.0f) ? 1u : 0u) << c;
        }
        m2[row * J_MASK_WORDS + cb] = mw;
    }
    pipe.passes++;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();

    double sx[8], cx[8];  // sin/cos of the angle states for the chain rule
    if (owner) {
        for (int i = 0; i < n; i++) {
            sx[i] = 0.0;
            cx[i] = 1.0;
            if (a.d.is_angle[i]) mc.sincos(x[i], sx[i], cx[i]);
        }
    }
    float* partial = reinterpret_cast<float*>(As);  // [T_SPLIT][T_ROWS][MAXIN], valid between a pass and the next image
    for (int j = 0; j < nh; j++) {
        // A image: g2 = W3[j, :] * m2 (bf16), same swizzled layout as H1
        for (int gi = 0; gi < G0; gi++) {
            const int g = part * G0 + gi;
            const uint32_t bits = (m2[row * J_MASK_WORDS + (g >> 2)] >> ((g & 3) * 8)) & 0xffu;
            const float4 wa = *reinterpret_cast<const float4*>(&w->W3T[j][g * 8]);
            const float4 wb = *reinterpret_cast<const float4*>(&w->W3T[j][g * 8 + 4]);
            const float w3[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
            uint32_t packed[4];
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
                const float lo = ((bits >> e) & 1u) ? w3[e] : 0.0f;
                const float hi = ((bits >> (e + 1)) & 1u) ? w3[e + 1] : 0.0f;
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(packed[e >> 1]) : "f"(hi), "f"(lo));
            }
            const int kb = g >> 3, cidx = (g & 7) ^ (row & 7);
            *reinterpret_cast<uint4*>(rowp + kb * (T_ROWS * 128) + cidx * 16) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        gemm_a_ready(pipe, t);
        gemm_issue<J_STAGES>(pipe, w->W2bT, t, (j + 1 < nh) ? w->W2bT : nullptr);
        gemm_wait(pipe, part);
        // epilogue: g1 = acc * m1, gin = g1 W1 (this thread's quarter of the hidden units)
        {
            float gin[MAXIN];
#pragma unroll
            for (int c = 0; c < MAXIN; c++) gin[c] = 0.0f;
            for (int cb = part * (H / 32 / T_SPLIT); cb < (part + 1) * (H / 32 / T_SPLIT); cb++) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + cb * 32, v);
                const uint32_t mw = m1[row * J_MASK_WORDS + cb];
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    const float g1 = ((mw >> c) & 1u) ? __uint_as_float(v[c]) : 0.0f;
                    const float4* w1 = nin + 1 <= W1S_FLOATS ? reinterpret_cast<const float4*>(w1s + (cb * 32 + c) * W1S_FLOATS)
                                                             : reinterpret_cast<const float4*>(&w->W1[cb * 32 + c][0]);
#pragma unroll
                    for (int q = 0; q < MAXIN / 4; q++) {
                        if (q * 4 < nin) {
                            const float4 ww = w1[q];
                            gin[q * 4 + 0] = fmaf(g1, ww.x, gin[q * 4 + 0]);
                            gin[q * 4 + 1] = fmaf(g1, ww.y, gin[q * 4 + 1]);
                            gin[q * 4 + 2] = fmaf(g1, ww.z, gin[q * 4 + 2]);
                            gin[q * 4 + 3] = fmaf(g1, ww.w, gin[q * 4 + 3]);
                        }
                    }
                }
            }
            gemm_wait(pipe, 3);  // the partial sums alias the A image: every MMA of this pass must have read it
#pragma unroll
            for (int c = 0; c < MAXIN; c++) partial[(part * T_ROWS + row) * MAXIN + c] = gin[c];
        }
        pipe.passes++;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (rw.valid) {
            float gin[MAXIN];
#pragma unroll
            for (int c = 0; c < MAXIN; c++) {
                float v = 0.0f;
#pragma unroll
                for (int q = 0; q < T_SPLIT; q++) v += partial[(q * T_ROWS + row) * MAXIN + c];
                gin[c] = v;
            }
            // d rhs[nh + j] / d(x, u) = (1/dt) dxVar_j * sum_c gin_c * d in_c / d(x, u)
            const double sc = (double)((float)(1.0 / a.dt) * w->dx_var[j]);
            const int64_t B = a.Bs, b = rw.b;
            int c = 0;
            for (int i = 0; i < n; i++) {
                double dcol;
                if (a.d.is_angle[i]) {
                    dcol = (double)gin[c] * (-sx[i]) / (double)w->o_var[c] + (double)gin[c + 1] * cx[i] / (double)w->o_var[c + 1];
                    c += 2;
                } else {
                    dcol = (double)gin[c] / (double)w->o_var[c];
                    c++;
                }
                a.A[(((int64_t)rw.k * n + nh + j) * n + i) * B + b] = sc * dcol;
            }
            for (int jj = 0; jj < m; jj++) {
                a.Bm[(((int64_t)rw.k * n + nh + j) * m + jj) * B + b] = sc * (double)gin[c] / (double)w->u_var[jj];
                c++;
            }
        }
        __syncthreads();  // the partial sums are consumed before the next pass writes its image over them
    }
    if (rw.valid) {  // position half: d x[nh + i] / dx
        const int64_t B = a.Bs, b = rw.b;
        for (int i = 0; i < nh; i++) {
            for (int c = 0; c < n; c++) a.A[(((int64_t)rw.k * n + i) * n + c) * B + b] = (c == nh + i) ? 1.0 : 0.0;
            for (int jj = 0; jj < m; jj++) a.Bm[(((int64_t)rw.k * n + i) * m + jj) * B + b] = 0.0;
        }
    }
    __syncthreads();
    cluster_sync();  // no CTA leaves while its peer may still signal its barriers
    if (warp == 0) {
        if constexpr (TWO_SM) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ---- reference-precision tensor-core path (PG_MLP_F16X3) -----------------------------------------------------------------
// The reference's network is fp32 (nn_models.py:179-202).  Here every operand of the 500 x 500 layer is split x = hi + lo
// into two fp16 numbers (22 mantissa bits) and the product is formed as Ah Wh + Al Wh + Ah Wl with fp32 accumulation in
// TMEM: fp32-class results (the dropped term is 2^-22) at a third of the fp16 tensor rate.  Twice the operand bytes do not
// fit the "whole A image in shared memory" design above (2 x 128 KB), so the GEMM runs K-BLOCK-MAJOR instead:
//   for each block of 64 hidden units kb:  all 512 threads compute layer 1 for (128 rows x 64 units) on the CUDA cores in
//   fp32 -- exactly the reference's arithmetic for that layer -- and write it, split, into one of three 32 KB A buffers;
//   the leader's two issuing threads (one per 256-column half of the accumulators, in control warps of their own) multiply
//   it against the four 16 KB weight chunks of that block (two 256-column halves x {hi, lo}; cta_group::2, M = 256 = both
//   CTAs' rows, N = 256) into ALL 512 accumulator columns,
// and layer 1 of the blocks ahead overlaps the MMAs of block kb.  No layer-1 GEMM pass, no TMEM -> smem conversion of H1, a
// 6-stage weight ring where the bf16 kernel has room for 4.  (X_NST / "2 x" below: the budget the geometry was first written
// for -- two A buffers + eight stages; XR_ABUFS / XR_NST further down redistribute the same 192 KB.)
constexpr int X_NST = 8;
constexpr int X_ABUF_BYTES = 2 * CHUNK_BYTES;  // A_hi | A_lo of one k-block: [128 rows x 64 k] fp16 each
constexpr int X_SMEM = 2 * X_ABUF_BYTES + X_NST * CHUNK_BYTES + T_INS_BYTES + T_CONST_BYTES + 256;
constexpr uint32_t X_IDESC_2SM = umma_idesc_f16(0u, 256u, 256u);
static_assert(X_SMEM <= 232448, "x3 kernels: shared memory");

struct XPipe {
    uint8_t* Ab;       // XR_ABUFS A buffers
    uint8_t* Bs;       // weight ring
    uint64_t *full, *empty, *pfull;  // [XR_NST]
    uint64_t *a_full, *a_empty;      // [XR_ABUFS]
    uint64_t* accum;
    uint32_t tmem_base;
    uint32_t c;       // k-steps issued so far by this CTA (barrier phases derive from it)
    uint32_t total;   // k-steps this CTA will run in total (no weight chunk is requested beyond it)
    uint32_t passes;  // GEMMs completed (phase of `accum`)
};

// 16 values of this thread (its row, hidden units kb * 64 + part * 16 .. + 15) -> split -> the A buffer
__device__ __forceinline__ void x_store16(const float (&v)[16], uint8_t* abuf, int row, int part) {
#pragma unroll
    for (int uu = 0; uu < 2; uu++) {
        float e8[8];
#pragma unroll
        for (int e = 0; e < 8; e++) e8[e] = v[uu * 8 + e];
        uint4 hi, lo;
        f16_split8(e8, hi, lo);
        const int off = chunk_off(row, part * 16 + uu * 8);
        *reinterpret_cast<uint4*>(abuf + off) = hi;
        *reinterpret_cast<uint4*>(abuf + CHUNK_BYTES + off) = lo;
    }
}

// Layer 1 of one block of 64 hidden units in fp32 on the CUDA cores (weights with the bias folded into slot n_in, the input
// vector carries a 1 there; warp-uniform 128-bit reads, broadcast) -> h1 = relu(.) -> A buffer; returns the 16 ReLU mask bits.
template <int NQ>
__device__ __forceinline__ uint32_t x_l1_block(const float* w1, int w1_stride, const float (&in)[16], int kb, int row, int part,
                                              uint8_t* abuf) {
    float v[16];
    uint32_t bits = 0;
#pragma unroll
    for (int e = 0; e < 16; e++) {
        const int k = kb * 64 + part * 16 + e;
        const float4* wr = reinterpret_cast<const float4*>(w1 + (size_t)k * w1_stride);
        float acc = 0.0f;
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const float4 ww = wr[q];
            acc = fmaf(ww.x, in[q * 4 + 0], acc);
            acc = fmaf(ww.y, in[q * 4 + 1], acc);
            acc = fmaf(ww.z, in[q * 4 + 2], acc);
            acc = fmaf(ww.w, in[q * 4 + 3], acc);
        }
        const bool pos = k < HID && acc > 0.0f;
        v[e] = pos ? acc : 0.0f;
        bits |= (pos ? 1u : 0u) << e;
    }
    x_store16(v, abuf, row, part);
    return bits;
}

// The same block with packed arithmetic (fma.rn.f32x2: two hidden units per instruction; per unit the same sequence of
// fused multiply-adds as x_l1_block, so the values are identical), for the rollout kernel, whose layer 1 took as long as the
// tensor core needs for the block (tools/mlp_timing.py).  w1p: the weight rows of a PAIR of units interleaved,
// w1p[(k / 2) * 16 + 2 i + (k & 1)] = W1b[k][i], i < 8 (stage_w1_pairs); in2[i] = {in[i], in[i]}.  Inputs + 1 <= 8.
__device__ __forceinline__ void ffma2(float2& acc, const float2 a, const float2 b) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rc;\n\t}"
        : "+f"(acc.x), "+f"(acc.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
}

__device__ __forceinline__ void stage_w1_pairs(const MlpPacked* w, float* w1p, int t) {
    for (int i = t; i < H * W1S_FLOATS; i += T_THREADS) {
        const int k = i / W1S_FLOATS, c = i % W1S_FLOATS;
        w1p[(k >> 1) * (2 * W1S_FLOATS) + 2 * c + (k & 1)] = w->W1b[k][c];
    }
}

template <int NQ>
__device__ __forceinline__ void x_l1_block_pairs(const float* w1p, const float2 (&in2)[8], int kb, int row, int part, uint8_t* abuf) {
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
        const int k = kb * 64 + part * 16 + e;
        const float4* wr = reinterpret_cast<const float4*>(w1p + (size_t)(k >> 1) * (2 * W1S_FLOATS));
        float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int q = 0; q < 2 * NQ; q++) {  // inputs 2 q, 2 q + 1 of both units
            const float4 ww = wr[q];
            ffma2(acc, make_float2(ww.x, ww.y), in2[2 * q]);
            ffma2(acc, make_float2(ww.z, ww.w), in2[2 * q + 1]);
        }
        // relu; the rows of the padding units (k >= HID) are zero
        v[e] = fmaxf(acc.x, 0.0f);
        v[e + 1] = fmaxf(acc.y, 0.0f);
    }
    x_store16(v, abuf, row, part);
}

__device__ __forceinline__ uint32_t x_l1_dispatch(int nin, const MlpPacked* w, const float* w1s, const float (&in)[16], int kb, int row,
                                                 int part, uint8_t* abuf) {
    const int nq = (nin + 1 + 3) >> 2;  // inputs + the constant 1, in float4 units
    if (nq <= 1) return x_l1_block<1>(w1s, W1S_FLOATS, in, kb, row, part, abuf);
    if (nq == 2) return x_l1_block<2>(w1s, W1S_FLOATS, in, kb, row, part, abuf);
    if (nq == 3) return x_l1_block<3>(&w->W1b[0][0], MAXIN, in, kb, row, part, abuf);
    return x_l1_block<4>(&w->W1b[0][0], MAXIN, in, kb, row, part, abuf);
}

struct XSmem {
    uint8_t *Ab, *Bs;
    uint64_t *full, *empty, *pfull, *a_full, *a_empty, *accum;
    uint32_t* tmem_slot;
};

// The rollout kernel keeps the tensor core's feed out of the compute warps: two more warps hold one control thread each --
// warp 16 issues the MMAs (leader) or relays "my stage is full" (peer), warp 17 requests the weight chunks -- and the 16
// compute warps hand each A block over through an mbarrier (one arrival per warp of BOTH CTAs on the leader's a_full) instead
// of a CTA barrier.  With the control threads inside compute warps every k-block cost layer 1 PLUS the time the issuing
// thread is held by the tensor core's queue: 37 % of all warp samples sat at that barrier
// (profiles/r02_mlp_x3_ncu_summary.txt).  The two control threads must not share a warp either: a thread held at the issue of
// a tcgen05.mma holds its warp, so a producer in a sibling lane only ran while the issuer waited -- the refills then followed
// the MMAs instead of overlapping them (tools/mlp_timing.py: "wait weights" = "issuing").
constexpr int XC_THREADS = T_THREADS + 96;
// -DPG_MLP_TIMING (tools/mlp_timing.py; never in the product build): cycles per phase of mlp_x3_kernel, summed per CTA --
// compute thread 0: [0] owner phase + input hand-over, [1] the eight layer-1 blocks, [5] their waits for a free A buffer,
// [2] waiting for the last MMAs, [3] epilogue, [4] partial sums + state update; the leader's issuing thread: [8] waiting
// for an A block, [9] waiting for weight stages, [10] issuing; [15] intervals.
#ifdef PG_MLP_TIMING
__device__ unsigned long long pg_mlp_timing_buf[1024 * 16];
#define PG_TIME_DECL long long tm_last = clock64()
#define PG_TIME_ADD(slot)                                                                          \
    do {                                                                                           \
        const long long tm_now = clock64();                                                        \
        atomicAdd(&pg_mlp_timing_buf[(blockIdx.x & 1023) * 16 + (slot)], (unsigned long long)(tm_now - tm_last)); \
        tm_last = tm_now;                                                                          \
    } while (0)
#else
#define PG_TIME_DECL
#define PG_TIME_ADD(slot)
#endif
// Ring geometry of the rollout kernel: THREE A buffers and a SIX-stage weight ring in the 192 KB the two A buffers and eight
// stages of the Jacobian kernel take.  With two A buffers layer 1 of block c + 2 could not start before the MMAs of block c
// had retired, and those of block c + 2 not before it was written: ~9 k of the 50 k cycles of an interval were fill and
// drain of that ping-pong.  Weight stage of (k-step c, chunk j): g = 4 c + j -> stage g % 6, use g / 6; the stages were
// never late (tools/mlp_timing.py), and six still cover more than a k-step of MMAs.
constexpr int XR_NST = 6, XR_ABUFS = 3;
static_assert(XR_ABUFS * X_ABUF_BYTES + XR_NST * CHUNK_BYTES == 2 * X_ABUF_BYTES + X_NST * CHUNK_BYTES, "rollout ring: same shared memory");
__device__ __forceinline__ void xr_barriers(uint8_t* base, XSmem& m) {
    m.full = reinterpret_cast<uint64_t*>(base);
    m.empty = m.full + XR_NST;
    m.pfull = m.empty + XR_NST;
    m.a_full = m.pfull + XR_NST;
    m.a_empty = m.a_full + XR_ABUFS;
    m.accum = m.a_empty + XR_ABUFS;
    m.tmem_slot = reinterpret_cast<uint32_t*>(m.accum + 1);
}
__device__ __forceinline__ void xr_setup(const XSmem& m, int t, int warp, uint32_t a_full_count, uint32_t issuers) {
    if (t == 0) {
        for (int s = 0; s < XR_NST; s++) {
            mbar_init(&m.full[s], 1);
            mbar_init(&m.empty[s], 1);
            mbar_init(&m.pfull[s], 1);
        }
        for (int s = 0; s < XR_ABUFS; s++) {
            mbar_init(&m.a_full[s], a_full_count);
            mbar_init(&m.a_empty[s], issuers);
        }
        mbar_init(m.accum, issuers);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // the whole accumulator file of the SM pair: 128 lanes x 512 fp32 columns per CTA
        asm volatile(PG_TMEM_ALLOC_2SM ::"r"(smem_u32(m.tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
// weight chunks of k-step c -> their ring stages.  Producer thread only.
__device__ __forceinline__ void xr_request(const XPipe& p, uint32_t c, const uint16_t (*imgh)[CHUNK_BYTES / 2],
                                           const uint16_t (*imgl)[CHUNK_BYTES / 2], int kb, uint32_t rank) {
#pragma unroll
    for (int j = 0; j < 4; j++) {  // j = 2 * (column half) + (lo part)
        const uint32_t g = 4 * c + j, s = g % XR_NST, n = g / XR_NST;
        mbar_wait(&p.empty[s], (n & 1) ^ 1);
        mbar_expect_tx(&p.full[s], CHUNK_BYTES);
        const uint16_t(*img)[CHUNK_BYTES / 2] = (j & 1) ? imgl : imgh;
        bulk_g2s(p.Bs + s * CHUNK_BYTES, &img[(2 * (j >> 1) + rank) * 8 + kb][0], CHUNK_BYTES, &p.full[s]);
    }
}
__device__ __forceinline__ uint8_t* xr_abuf_parked(const XPipe& p) {
    const uint32_t set = p.c % XR_ABUFS, u = p.c / XR_ABUFS;
    mbar_wait_parked(&p.a_empty[set], (u & 1) ^ 1);
    return p.Ab + set * X_ABUF_BYTES;
}
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(T_THREADS) : "memory"); }
__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// cwarp 0 / 1: the issuer of accumulator half 0 / 1 (weight stages j = 2 * half, 2 * half + 1 of every k-step: the halves are
// disjoint TMEM columns, so two threads can feed the tensor cores side by side -- one thread needed longer to wait for and
// issue a k-step than the MMAs take); cwarp 2: the producer.
// JAC: the Jacobian kernel's sequence -- the first GEMM (8 k-steps) against W2, the others against its transpose.
template <bool JAC>
__device__ __forceinline__ void x_control(const XPipe& p, const MlpPacked* w, int cwarp, int lane, uint32_t rank) {
    const int jh = 2 * cwarp;
    if (lane != 0) {
    } else if (cwarp == 2) {  // producer: a stage is requested again as soon as the MMAs that read it have retired
        for (uint32_t c = 0; c < p.total; c++) {
            if (JAC && c >= 8) xr_request(p, c, w->W2hT, w->W2lT, (int)(c & 7), rank);
            else xr_request(p, c, w->W2h, w->W2l, (int)(c & 7), rank);
        }
    } else if (rank != 0) {  // peer: tell the leader when a stage of OUR ring is full
        for (uint32_t c = 0; c < p.total; c++) {
#pragma unroll
            for (int j = jh; j < jh + 2; j++) {
                const uint32_t g = 4 * c + j, s = g % XR_NST, n = g / XR_NST;
                mbar_wait(&p.full[s], n & 1);
                mbar_arrive_remote(&p.pfull[s], 0);
            }
        }
    } else {  // leader: drives both tensor cores (its half of the accumulator columns)
        PG_TIME_DECL;
        for (uint32_t c = 0; c < p.total; c++) {
            const uint32_t set = c % XR_ABUFS, u = c / XR_ABUFS;
            const int kb = (int)(c & 7);
            PG_TIME_ADD(10);
            mbar_wait(&p.a_full[set], u & 1);
            PG_TIME_ADD(8);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // descriptors: the start address sits in the low 14 bits in 16-byte units, so the k-slice ks (+32 bytes) is + 2 * ks
            const uint64_t da_hi = umma_desc_sw128(smem_u32(p.Ab + set * X_ABUF_BYTES)), da_lo = da_hi + (CHUNK_BYTES >> 4);
#pragma unroll
            for (int j = jh; j < jh + 2; j++) {
                const uint32_t g = 4 * c + j, s = g % XR_NST, n = g / XR_NST;
                PG_TIME_ADD(10);
                mbar_wait(&p.full[s], n & 1);
#ifdef PG_MLP_TIMING
                PG_TIME_ADD(9);
#endif
                mbar_wait(&p.pfull[s], n & 1);
                PG_TIME_ADD(11);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t db = umma_desc_sw128(smem_u32(p.Bs + s * CHUNK_BYTES));
                const uint32_t d = p.tmem_base + (uint32_t)cwarp * 256;
                if ((j & 1) == 0) {  // Wh: Al Wh (small) first, then Ah Wh
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) umma_f16_2sm(d, da_lo + 2 * ks, db + 2 * ks, X_IDESC_2SM, (kb > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) umma_f16_2sm(d, da_hi + 2 * ks, db + 2 * ks, X_IDESC_2SM, 1u);
                } else {  // Wl: Ah Wl
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) umma_f16_2sm(d, da_hi + 2 * ks, db + 2 * ks, X_IDESC_2SM, 1u);
                }
                umma_commit_2sm(&p.empty[s], CL_MASK);  // frees the stage in both CTAs
            }
            umma_commit_2sm(&p.a_empty[set], CL_MASK);  // ... and the A buffer
            if (kb == 7) umma_commit_2sm(p.accum, CL_MASK);  // all 512 accumulator columns final in both CTAs
        }
    }
    __syncwarp();
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(XC_THREADS, 1) mlp_x3_kernel(const MlpArgs a) {
    extern __shared__ __align__(1024) uint8_t tsm[];
    XSmem sm;
    sm.Ab = tsm;
    sm.Bs = tsm + XR_ABUFS * X_ABUF_BYTES;
    float* ins = reinterpret_cast<float*>(sm.Bs + XR_NST * CHUNK_BYTES);
    float* yp = ins;  // partial layer-3 sums [T_SPLIT][T_ROWS][MAXDX]: the input tile is in registers by then
    float4* eps = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(ins) + T_INS_BYTES);
    float* w1s = reinterpret_cast<float*>(eps + H);
    xr_barriers(reinterpret_cast<uint8_t*>(w1s) + T_W1_BYTES, sm);
    const MlpPacked* w = a.w;
    const int t = threadIdx.x, warp = t >> 5, row = t & (T_ROWS - 1), part = t >> 7;
    const int steps = steps_of(a);
    const int64_t pairs = (total_rows(a) + CL * T_ROWS - 1) / (CL * T_ROWS), slot = blockIdx.x / CL;
    int64_t lo, hi;
    if (a.progress) {  // balanced schedule: see mlp_bf16_kernel
        const int64_t slots = min((int64_t)(gridDim.x / CL), pairs);
        if (slot >= slots) return;
        lo = slot * (pairs * steps) / slots;
        hi = (slot + 1) * (pairs * steps) / slots;
    } else {
        if (slot >= pairs) return;
        lo = slot * steps;
        hi = lo + steps;
    }
    const bool owner = t < T_ROWS;
    const int nin = a.d.n_obs + a.d.m;
    const pg::MathCtx<double> mc;
    const uint32_t rank = cluster_ctarank();

    xr_setup(sm, t, warp, CL * (T_THREADS / 32), 2);
    const int nq = (nin + 1 + 3) >> 2;  // inputs + the constant 1, in float4 units
    if (t < T_THREADS) {
        stage_constants(w, eps, nq <= 2 ? nullptr : w1s, t);
        if (nq <= 2) stage_w1_pairs(w, w1s, t);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();  // the peer's barriers are initialised before anything arrives on them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    XPipe pipe = {sm.Ab, sm.Bs, sm.full, sm.empty, sm.pfull, sm.a_full, sm.a_empty, sm.accum, *sm.tmem_slot, 0u, (uint32_t)((hi - lo) * 8), 0u};
    const uint32_t tmem_base = pipe.tmem_base;
    if (warp >= T_THREADS / 32) {  // the control warps
        x_control<false>(pipe, w, warp - T_THREADS / 32, t & 31, rank);
    } else {

    double x[8], u[2], xe[8], ue[2], fplus[8];
    if (steps == 0 && owner) {  // T = 0: only the initial states
        Row r0 = decode_row(a, (int64_t)blockIdx.x * T_ROWS + row);
        load_row(a, r0, x, u);
    }
    bool first = true;
    PG_TIME_DECL;
    for (int64_t pos = hi; pos > lo;) {
    const int64_t pair = (pos - 1) / steps, base = pair * steps;
    const int k_begin = (int)(max(lo, base) - base), k_end = (int)(pos - base);
    pos = base + k_begin;
    const int64_t tile = pair * CL + (blockIdx.x % CL);
    Row rw = decode_row(a, tile * T_ROWS + row);
    rw.valid = rw.valid && owner;
    if (k_begin > 0) {  // continue where another cluster left this tile
        if (t == 0) wait_progress(&a.progress[tile], k_begin);
        compute_sync();
    }
    if (owner) {
        if (k_begin == 0) load_row(a, rw, x, u);
        else reload_row(a, rw, k_begin, x, u);
    }
    for (int k = k_begin; k < k_end; k++) {
        if (!first) compute_sync();  // the previous interval's partial sums (they alias `ins`) have been consumed
        first = false;
        if (owner) {
            pre_step(a, rw, k, x, u, xe, ue);
            float in[MAXIN];
            network_input(a, mc, xe, ue, in);
#pragma unroll
            for (int j = 0; j < MAXIN; j += 4)
                *reinterpret_cast<float4*>(&ins[row * INS_STRIDE + j]) = make_float4(in[j], in[j + 1], in[j + 2], in[j + 3]);
        }
        compute_sync();
        if (t == 0) PG_TIME_ADD(0);
        float in[16];
#pragma unroll
        for (int j = 0; j < MAXIN; j += 4) {
            const float4 v4 = *reinterpret_cast<const float4*>(&ins[row * INS_STRIDE + j]);
            in[j] = v4.x; in[j + 1] = v4.y; in[j + 2] = v4.z; in[j + 3] = v4.w;
        }
        float2 in2[8];
#pragma unroll
        for (int j = 0; j < 8; j++) in2[j] = make_float2(in[j], in[j]);
        for (int kb = 0; kb < 8; kb++) {
            if (t == 0) PG_TIME_ADD(1);
            uint8_t* abuf = xr_abuf_parked(pipe);
            if (t == 0) PG_TIME_ADD(5);
            if (nq <= 1) x_l1_block_pairs<1>(w1s, in2, kb, row, part, abuf);
            else if (nq == 2) x_l1_block_pairs<2>(w1s, in2, kb, row, part, abuf);
            else x_l1_dispatch(nin, w, w1s, in, kb, row, part, abuf);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA unit
            __syncwarp();
            if ((t & 31) == 0) {  // this warp's share of the A block is written: one of the 2 x 16 arrivals the issuer waits for
                if (rank == 0) mbar_arrive_local(&pipe.a_full[pipe.c % XR_ABUFS]);
                else mbar_arrive_remote(&pipe.a_full[pipe.c % XR_ABUFS], 0);
            }
            pipe.c++;
        }
        // epilogue: TMEM lane = row; this thread's quarter of the 512 columns: +b2, ReLU, layer 3
        if (t == 0) PG_TIME_ADD(1);
        mbar_wait_parked(pipe.accum, pipe.passes & 1);
        if (t == 0) PG_TIME_ADD(2);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
            float y[MAXDX] = {0.0f, 0.0f, 0.0f, 0.0f};
            for (int cb = part * (H / 32 / T_SPLIT); cb < (part + 1) * (H / 32 / T_SPLIT); cb++) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + cb * 32, v);
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    const int nn = cb * 32 + c;
                    const float4 e4 = eps[nn];  // {b2, W3[nn][0..2]}: one 128-bit smem broadcast per column
                    const float h2 = fmaxf(__uint_as_float(v[c]) + e4.x, 0.0f);
                    y[0] = fmaf(h2, e4.y, y[0]);
                    y[1] = fmaf(h2, e4.z, y[1]);
                    y[2] = fmaf(h2, e4.w, y[2]);
                    if (a.d.dx_dim > 3) y[3] = fmaf(h2, w->W3[nn][3], y[3]);
                }
            }
            *reinterpret_cast<float4*>(&yp[(part * T_ROWS + row) * MAXDX]) = make_float4(y[0], y[1], y[2], y[3]);
        }
        pipe.passes++;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        compute_sync();  // every lane has read its accumulators before the next interval's MMAs overwrite them
        if (t == 0) PG_TIME_ADD(3);
        if (owner) {
            float y[MAXDX];
#pragma unroll
            for (int j = 0; j < MAXDX; j++) {
                float v = 0.0f;
#pragma unroll
                for (int q = 0; q < T_SPLIT; q++) v += yp[(q * T_ROWS + row) * MAXDX + j];
                y[j] = v;
            }
            post_step(a, rw, k, y, x, u, xe, fplus);
        }
#ifdef PG_MLP_TIMING
        if (t == 0) {
            PG_TIME_ADD(4);
            atomicAdd(&pg_mlp_timing_buf[(blockIdx.x & 1023) * 16 + 15], 1ull);
        }
#endif
    }
    if (a.progress && k_end < steps) {  // the head of a pair: publish that X holds the states up to interval k_end
        __threadfence();
        compute_sync();
        if (t == 0) publish_progress(&a.progress[tile], k_end);
    }
    }
    }  // compute warps
    __syncthreads();
    cluster_sync();  // no CTA leaves while its peer may still signal its barriers
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// Analytic Jacobian at reference precision: the passes of mlp_jac_kernel on the k-block-major split-fp16 GEMM.
constexpr int XJ_SMEM = 2 * X_ABUF_BYTES + X_NST * CHUNK_BYTES + T_W1_BYTES + J_MASK_BYTES + 256;
static_assert(XJ_SMEM <= 232448 && J_MASK_BYTES >= T_INS_BYTES, "jacobian x3 kernel: shared memory");

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(XC_THREADS, 1) mlp_jac_x3_kernel(const MlpArgs a) {
    extern __shared__ __align__(1024) uint8_t tsm[];
    XSmem sm;
    sm.Ab = tsm;
    sm.Bs = tsm + XR_ABUFS * X_ABUF_BYTES;  // ring geometry and control warps of mlp_x3_kernel
    float* w1s = reinterpret_cast<float*>(sm.Bs + XR_NST * CHUNK_BYTES);
    uint32_t* m1 = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(w1s) + T_W1_BYTES);
    uint32_t* m2 = m1 + T_ROWS * J_MASK_WORDS;
    float* ins = reinterpret_cast<float*>(m1);  // the input tile lives in the mask area until every thread has it in registers
    xr_barriers(reinterpret_cast<uint8_t*>(m1) + J_MASK_BYTES, sm);
    const MlpPacked* w = a.w;
    const int t = threadIdx.x, warp = t >> 5, row = t & (T_ROWS - 1), part = t >> 7;
    if ((int64_t)(blockIdx.x / CL) * CL * T_ROWS >= total_rows(a)) return;  // surplus clusters
    const bool owner = t < T_ROWS;
    Row rw = decode_row(a, (int64_t)blockIdx.x * T_ROWS + row);
    rw.valid = rw.valid && owner;
    const int n = a.d.n, nh = a.d.n / 2, m = a.d.m, nin = a.d.n_obs + a.d.m;
    const pg::MathCtx<double> mc;
    const uint32_t rank = cluster_ctarank();

    xr_setup(sm, t, warp, CL * (T_THREADS / 32), 2);
    for (int i = t; i < H * (W1S_FLOATS / 4) && t < T_THREADS; i += T_THREADS) {
        const int k = i / (W1S_FLOATS / 4), q = i % (W1S_FLOATS / 4);
        reinterpret_cast<float4*>(w1s)[i] = reinterpret_cast<const float4*>(&w->W1b[k][0])[q];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    XPipe pipe = {sm.Ab, sm.Bs, sm.full, sm.empty, sm.pfull, sm.a_full, sm.a_empty, sm.accum, *sm.tmem_slot, 0u, (uint32_t)(8 * (1 + nh)), 0u};
    const uint32_t tmem_base = pipe.tmem_base;
    if (warp >= T_THREADS / 32) {  // the control warps
        x_control<true>(pipe, w, warp - T_THREADS / 32, t & 31, rank);
    } else {
    // one of the 2 x 16 warp arrivals the issuers wait for: this warp's share of the A block of k-step pipe.c is written
    auto hand_over = [&]() {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if ((t & 31) == 0) {
            if (rank == 0) mbar_arrive_local(&pipe.a_full[pipe.c % XR_ABUFS]);
            else mbar_arrive_remote(&pipe.a_full[pipe.c % XR_ABUFS], 0);
        }
        pipe.c++;
    };

    double x[8], u[2];
    if (owner) {
        load_row(a, rw, x, u);
        float in[MAXIN];
        network_input(a, mc, x, u, in);
#pragma unroll
        for (int j = 0; j < MAXIN; j++) ins[row * INS_STRIDE + j] = in[j];
    }
    compute_sync();
    float in[16];
#pragma unroll
    for (int j = 0; j < MAXIN; j++) in[j] = ins[row * INS_STRIDE + j];
    compute_sync();  // the mask words alias the input tile
    uint16_t* m1h = reinterpret_cast<uint16_t*>(m1);
    // pass 0: forward, ReLU masks of both layers
    for (int kb = 0; kb < 8; kb++) {
        uint8_t* abuf = xr_abuf_parked(pipe);
        const uint32_t bits = x_l1_dispatch(nin, w, w1s, in, kb, row, part, abuf);
        m1h[row * (2 * J_MASK_WORDS) + kb * 4 + part] = (uint16_t)bits;  // hidden units kb * 64 + part * 16 .. + 15
        hand_over();
    }
    mbar_wait_parked(pipe.accum, pipe.passes & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int cb = part * (H / 32 / T_SPLIT); cb < (part + 1) * (H / 32 / T_SPLIT); cb++) {  // mask of layer 2: [acc + b2 > 0]
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + cb * 32, v);
        uint32_t mw = 0;
#pragma unroll
        for (int c = 0; c < 32; c++) {
            const int nn = cb * 32 + c;
            mw |= ((nn < HID && __uint_as_float(v[c]) + w->b2[nn] > 0.0f) ? 1u : 0u) << c;
        }
        m2[row * J_MASK_WORDS + cb] = mw;
    }
    pipe.passes++;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    compute_sync();

    double sx[8], cx[8];  // sin/cos of the angle states for the chain rule
    if (owner) {
        for (int i = 0; i < n; i++) {
            sx[i] = 0.0;
            cx[i] = 1.0;
            if (a.d.is_angle[i]) mc.sincos(x[i], sx[i], cx[i]);
        }
    }
    float* partial = reinterpret_cast<float*>(sm.Ab);  // [T_SPLIT][T_ROWS][MAXIN]: valid between a pass's last MMA and the next A block
    for (int j = 0; j < nh; j++) {
        for (int kb = 0; kb < 8; kb++) {  // A block: g2 = W3[j, :] * m2
            uint8_t* abuf = xr_abuf_parked(pipe);
            const int k0 = kb * 64 + part * 16;
            const uint32_t bits = (m2[row * J_MASK_WORDS + (k0 >> 5)] >> (k0 & 31)) & 0xffffu;
            float v[16];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float4 w4 = *reinterpret_cast<const float4*>(&w->W3T[j][k0 + q * 4]);
                v[q * 4 + 0] = ((bits >> (q * 4 + 0)) & 1u) ? w4.x : 0.0f;
                v[q * 4 + 1] = ((bits >> (q * 4 + 1)) & 1u) ? w4.y : 0.0f;
                v[q * 4 + 2] = ((bits >> (q * 4 + 2)) & 1u) ? w4.z : 0.0f;
                v[q * 4 + 3] = ((bits >> (q * 4 + 3)) & 1u) ? w4.w : 0.0f;
            }
            x_store16(v, abuf, row, part);
            hand_over();
        }
        mbar_wait_parked(pipe.accum, pipe.passes & 1);  // every MMA of the pass is complete: the A buffers may be reused as scratch
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {  // epilogue: g1 = acc * m1, gin = g1 W1 (this thread's quarter of the hidden units)
            float gin[MAXIN];
#pragma unroll
            for (int c = 0; c < MAXIN; c++) gin[c] = 0.0f;
            for (int cb = part * (H / 32 / T_SPLIT); cb < (part + 1) * (H / 32 / T_SPLIT); cb++) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + cb * 32, v);
                const uint32_t mw = m1[row * J_MASK_WORDS + cb];
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    const float g1 = ((mw >> c) & 1u) ? __uint_as_float(v[c]) : 0.0f;
                    const float4* w1 = nin + 1 <= W1S_FLOATS ? reinterpret_cast<const float4*>(w1s + (cb * 32 + c) * W1S_FLOATS)
                                                             : reinterpret_cast<const float4*>(&w->W1[cb * 32 + c][0]);
#pragma unroll
                    for (int q = 0; q < MAXIN / 4; q++) {
                        if (q * 4 < nin) {
                            const float4 ww = w1[q];
                            gin[q * 4 + 0] = fmaf(g1, ww.x, gin[q * 4 + 0]);
                            gin[q * 4 + 1] = fmaf(g1, ww.y, gin[q * 4 + 1]);
                            gin[q * 4 + 2] = fmaf(g1, ww.z, gin[q * 4 + 2]);
                            gin[q * 4 + 3] = fmaf(g1, ww.w, gin[q * 4 + 3]);
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < MAXIN; c++) partial[(part * T_ROWS + row) * MAXIN + c] = gin[c];
        }
        pipe.passes++;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        compute_sync();
        if (rw.valid) {
            float gin[MAXIN];
#pragma unroll
            for (int c = 0; c < MAXIN; c++) {
                float v = 0.0f;
#pragma unroll
                for (int q = 0; q < T_SPLIT; q++) v += partial[(q * T_ROWS + row) * MAXIN + c];
                gin[c] = v;
            }
            // d rhs[nh + j] / d(x, u) = (1/dt) dxVar_j * sum_c gin_c * d in_c / d(x, u)
            const double sc = (double)((float)(1.0 / a.dt) * w->dx_var[j]);
            const int64_t B = a.Bs, b = rw.b;
            int c = 0;
            for (int i = 0; i < n; i++) {
                double dcol;
                if (a.d.is_angle[i]) {
                    dcol = (double)gin[c] * (-sx[i]) / (double)w->o_var[c] + (double)gin[c + 1] * cx[i] / (double)w->o_var[c + 1];
                    c += 2;
                } else {
                    dcol = (double)gin[c] / (double)w->o_var[c];
                    c++;
                }
                a.A[(((int64_t)rw.k * n + nh + j) * n + i) * B + b] = sc * dcol;
            }
            for (int jj = 0; jj < m; jj++) {
                a.Bm[(((int64_t)rw.k * n + nh + j) * m + jj) * B + b] = sc * (double)gin[c] / (double)w->u_var[jj];
                c++;
            }
        }
        compute_sync();  // the partial sums are consumed before the next pass writes its A blocks over them
    }
    if (rw.valid) {  // position half: d x[nh + i] / dx
        const int64_t B = a.Bs, b = rw.b;
        for (int i = 0; i < nh; i++) {
            for (int c = 0; c < n; c++) a.A[(((int64_t)rw.k * n + i) * n + c) * B + b] = (c == nh + i) ? 1.0 : 0.0;
            for (int jj = 0; jj < m; jj++) a.Bm[(((int64_t)rw.k * n + i) * m + jj) * B + b] = 0.0;
        }
    }
    }  // compute warps
    __syncthreads();
    cluster_sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

int64_t max_rows(const MlpArgs& a) {  // upper bound of total_rows() known on the host (the work list may be shorter)
    if (a.mode == MODE_ROLLOUT) return a.R * a.n_alpha;
    if (a.mode == MODE_LIN_FD || a.mode == MODE_LIN_AN) return a.R * a.T;
    return a.R;
}

// rollouts on a tensor-core kernel (bf16 or split fp16), plain or with the balanced schedule
template <typename K>
int launch_tc(K kernel, int smem, int* resident, bool* attr_set, const MlpArgs& a, int64_t rows, cudaStream_t st, int threads = T_THREADS) {
    if (!*attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        *attr_set = true;
    }
    int64_t clusters = (rows + CL * T_ROWS - 1) / (CL * T_ROWS);
    if (a.progress) {  // balanced rollout: at most one cluster per pair of SMs, all of them resident at once
        if (!*resident) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(CL * 1024);
            cfg.blockDim = dim3(threads);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at;
            at.id = cudaLaunchAttributeClusterDimension;
            at.val.clusterDim.x = CL;
            at.val.clusterDim.y = at.val.clusterDim.z = 1;
            cfg.attrs = &at;
            cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n < 1) {
                (void)cudaGetLastError();
                return PG_E_UNSUPPORTED;
            }
            *resident = n;
        }
        cudaError_t e = cudaMemsetAsync(a.progress, 0, sizeof(int32_t) * (size_t)(clusters * CL), st);
        if (e != cudaSuccess) return (int)e;
        if (clusters > *resident) clusters = *resident;
    }
    kernel<<<(unsigned)(clusters * CL), threads, smem, st>>>(a);
    return 0;
}

int launch(int precision, const MlpArgs& a, cudaStream_t st) {
    const int64_t rows = max_rows(a);
    if (rows == 0) return 0;
    if (a.mode == MODE_LIN_AN) {
        // the analytic chain exists on the tensor-core paths only
        if (precision != PG_MLP_BF16 && precision != PG_MLP_F16X3) return PG_E_BADARG;
        static bool set = false, setx = false;
        const unsigned grid = (unsigned)((rows + CL * T_ROWS - 1) / (CL * T_ROWS)) * CL;
        if (precision == PG_MLP_BF16) {
            if (!set) {
                cudaError_t e = cudaFuncSetAttribute(mlp_jac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, J_SMEM);
                if (e != cudaSuccess) return (int)e;
                set = true;
            }
            mlp_jac_kernel<<<grid, T_THREADS, J_SMEM, st>>>(a);
        } else {
            if (!setx) {
                cudaError_t e = cudaFuncSetAttribute(mlp_jac_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XJ_SMEM);
                if (e != cudaSuccess) return (int)e;
                setx = true;
            }
            mlp_jac_x3_kernel<<<grid, XC_THREADS, XJ_SMEM, st>>>(a);
        }
    } else if (precision == PG_MLP_FP32) {
        const int smem = (F_ROWS * F_HSTR + F_ROWS * MAXIN + F_ROWS * MAXDX) * 4;
        static bool set = false;
        if (!set) {
            cudaError_t e = cudaFuncSetAttribute(mlp_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return (int)e;
            set = true;
        }
        mlp_fp32_kernel<<<(unsigned)((rows + F_ROWS - 1) / F_ROWS), 256, smem, st>>>(a);
    } else if (precision == PG_MLP_BF16) {
        static bool set = false;
        static int resident = 0;
        const int rc = launch_tc(mlp_bf16_kernel, T_SMEM, &resident, &set, a, rows, st);
        if (rc) return rc;
    } else if (precision == PG_MLP_F16X3) {
        static bool set = false;
        static int resident = 0;
        const int rc = launch_tc(mlp_x3_kernel, X_SMEM, &resident, &set, a, rows, st, XC_THREADS);
        if (rc) return rc;
    } else {
        return PG_E_BADARG;
    }
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

bool dims_ok(const pg_mlp_dims_t* d) {
    if (!d || d->n < 2 || d->n > 8 || (d->n & 1) || d->m < 1 || d->m > 2) return false;
    int no = 0;
    for (int i = 0; i < d->n; i++) no += d->is_angle[i] ? 2 : 1;
    return no == d->n_obs && d->n_obs + d->m + 1 <= MAXIN && d->dx_dim == d->n / 2 && d->dx_dim <= MAXDX;
}

}  // namespace

extern "C" {

int64_t pg_mlp_packed_bytes(void) { return (int64_t)sizeof(MlpPacked); }

int pg_mlp_pack(const pg_mlp_dims_t* d, const float* W1, const float* b1, const float* W2, const float* b2, const float* W3,
                const float* b3, const float* o_mean, const float* o_var, const float* u_mean, const float* u_var,
                const float* dx_mean, const float* dx_var, void* packed_host) {
    if (!dims_ok(d) || !W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !packed_host) return PG_E_BADARG;
    MlpPacked* p = (MlpPacked*)packed_host;
    memset(p, 0, sizeof(MlpPacked));
    const int nin = d->n_obs + d->m;
    for (int k = 0; k < HID; k++) {
        for (int j = 0; j < nin; j++) p->W1[k][j] = W1[k * nin + j];
        p->b1[k] = b1[k];
        p->b2[k] = b2[k];
        for (int j = 0; j < d->dx_dim; j++) p->W3[k][j] = W3[j * HID + k];
        for (int n = 0; n < HID; n++) p->W2t[k][n] = W2[n * HID + k];  // W2 is [out n][in k]
    }
    for (int k = 0; k < HID; k++) {
        for (int j = 0; j < nin; j++) p->W1b[k][j] = W1[k * nin + j];
        p->W1b[k][nin] = b1[k];
        p->ep[k] = make_float4(b2[k], W3[k], d->dx_dim > 1 ? W3[HID + k] : 0.0f, d->dx_dim > 2 ? W3[2 * HID + k] : 0.0f);
        for (int j = 0; j < d->dx_dim; j++) p->W3T[j][k] = W3[j * HID + k];
    }
    for (int j = 0; j < d->dx_dim; j++) {
        p->b3[j] = b3[j];
        p->dx_mean[j] = dx_mean ? dx_mean[j] : 0.0f;
        p->dx_var[j] = dx_var ? dx_var[j] : 1.0f;
    }
    for (int j = 0; j < MAXIN; j++) {
        p->o_mean[j] = (o_mean && j < d->n_obs) ? o_mean[j] : 0.0f;
        p->o_var[j] = (o_var && j < d->n_obs) ? o_var[j] : 1.0f;
    }
    for (int j = 0; j < 2; j++) {
        p->u_mean[j] = (u_mean && j < d->m) ? u_mean[j] : 0.0f;
        p->u_var[j] = (u_var && j < d->m) ? u_var[j] : 1.0f;
    }
    // bf16 chunk images: chunk = kb * 4 + quarter holds W2[n = quarter*128 + nl][k = kb*64 + kk] at the swizzled offset,
    // and the transposed set W2[nn = kb*64 + kk][kcol = quarter*128 + nl] (contraction over layer 2's output index)
    for (int kb = 0; kb < 8; kb++)
        for (int qt = 0; qt < 4; qt++) {
            uint16_t* img = p->W2b[qt * 8 + kb];  // load order of the GEMM: quarter-major
            uint16_t* imgT = p->W2bT[qt * 8 + kb];
            for (int nl = 0; nl < CHUNK_N; nl++)
                for (int kk = 0; kk < 64; kk++) {
                    const int n = qt * CHUNK_N + nl, k = kb * 64 + kk;
                    const int off = (nl >> 3) * 1024 + (nl & 7) * 128 + (((kk >> 3) ^ (nl & 7)) << 4) + (kk & 7) * 2;
                    const float wv = (n < HID && k < HID) ? W2[n * HID + k] : 0.0f, wt = (k < HID && n < HID) ? W2[k * HID + n] : 0.0f;
                    img[off >> 1] = f32_to_bf16(wv);
                    imgT[off >> 1] = f32_to_bf16(wt);
                    f32_to_f16_pair(wv, p->W2h[qt * 8 + kb][off >> 1], p->W2l[qt * 8 + kb][off >> 1]);
                    f32_to_f16_pair(wt, p->W2hT[qt * 8 + kb][off >> 1], p->W2lT[qt * 8 + kb][off >> 1]);
                }
        }
    // layer 1 for the tensor core: W1b split into bf16 hi + lo parts (x w ~ xh wh + xh wl + xl wh, error ~2^-16)
    for (int qt = 0; qt < 4; qt++)
        for (int nl = 0; nl < CHUNK_N; nl++)
            for (int kk = 0; kk < 64; kk++) {
                const int n = qt * CHUNK_N + nl, seg = kk >> 4, j = kk & 15;
                const float wv = (n < HID && j <= nin) ? p->W1b[n][j] : 0.0f;
                const uint16_t hi = f32_to_bf16(wv), lo = f32_to_bf16(wv - bf16_to_f32(hi));
                const int off = (nl >> 3) * 1024 + (nl & 7) * 128 + (((kk >> 3) ^ (nl & 7)) << 4) + (kk & 7) * 2;
                p->W1img[qt][off >> 1] = seg == 1 ? lo : (seg == 3 ? (uint16_t)0 : hi);
            }
    return 0;
}

static int forward_impl(int precision, const pg_mlp_dims_t* dims, const void* packed, int64_t B, int32_t T, double dt,
                        int32_t n_alpha, const double* alphas, const double* x0, const double* U, const double* KK,
                        const double* kk, const double* xbar, const double* ubar, const double* ulim, double* X,
                        double* U_out, const int32_t* index, const int32_t* count, int32_t* progress, void* stream);

int pg_mlp_ode(int precision, const pg_mlp_dims_t* dims, const void* packed, int64_t R, double dt, const double* x,
               const double* u, double* rhs, void* stream) {
    if (R == 0) return 0;
    if (!dims_ok(dims) || !packed || R < 0 || !x || !u || !rhs || !(dt > 0)) return PG_E_BADARG;
    MlpArgs a;
    memset(&a, 0, sizeof(a));
    a.d = *dims;
    a.w = (const MlpPacked*)packed;
    a.mode = MODE_ODE;
    a.R = a.Bs = R;
    a.T = 1;
    a.n_alpha = 1;
    a.dt = dt;
    a.x0 = x;
    a.U = u;
    a.rhs = rhs;
    return launch(precision, a, (cudaStream_t)stream);
}

int pg_mlp_rollout(int precision, const pg_mlp_dims_t* dims, const void* packed, int64_t B, int32_t T, double dt,
                   const double* x0, const double* U, const double* KK, const double* kk, const double* xbar,
                   const double* ubar, double alpha, const double* ulim, double* X, double* U_out, void* stream) {
    return pg_mlp_forward(precision, dims, packed, B, T, dt, 1, &alpha, x0, U, KK, kk, xbar, ubar, ulim, X, U_out, nullptr,
                          nullptr, stream);
}

int pg_mlp_forward(int precision, const pg_mlp_dims_t* dims, const void* packed, int64_t B, int32_t T, double dt,
                   int32_t n_alpha, const double* alphas, const double* x0, const double* U, const double* KK,
                   const double* kk, const double* xbar, const double* ubar, const double* ulim, double* X, double* U_out,
                   const int32_t* index, const int32_t* count, void* stream) {
    return forward_impl(precision, dims, packed, B, T, dt, n_alpha, alphas, x0, U, KK, kk, xbar, ubar, ulim, X, U_out, index,
                        count, nullptr, stream);
}

static int forward_impl(int precision, const pg_mlp_dims_t* dims, const void* packed, int64_t B, int32_t T, double dt,
                        int32_t n_alpha, const double* alphas, const double* x0, const double* U, const double* KK,
                        const double* kk, const double* xbar, const double* ubar, const double* ulim, double* X,
                        double* U_out, const int32_t* index, const int32_t* count, int32_t* progress, void* stream) {
    if (B == 0) return 0;
    if (!dims_ok(dims) || !packed || B < 0 || T < 0 || !x0 || !X || !(dt > 0)) return PG_E_BADARG;
    if (n_alpha < 1 || n_alpha > MAX_ALPHAS || !alphas) return PG_E_BADARG;
    if (KK && (!kk || !xbar || !ubar)) return PG_E_BADARG;
    if ((index == nullptr) != (count == nullptr)) return PG_E_BADARG;
    MlpArgs a;
    memset(&a, 0, sizeof(a));
    a.d = *dims;
    a.w = (const MlpPacked*)packed;
    a.mode = MODE_ROLLOUT;
    a.R = a.Bs = B;
    a.T = T;
    a.dt = dt;
    a.n_alpha = n_alpha;
    for (int i = 0; i < n_alpha; i++) a.alphas[i] = alphas[i];
    a.index = index;
    a.count = count;
    a.x0 = x0;
    a.U = U;
    a.KK = KK;
    a.kk = kk;
    a.xbar = xbar;
    a.ubar = ubar;
    a.clip = ulim != nullptr;
    for (int j = 0; j < dims->m && ulim; j++) a.ulim[j] = ulim[j];
    a.X = X;
    a.U_out = U_out;
    a.progress = T > 0 ? progress : nullptr;
    return launch(precision, a, (cudaStream_t)stream);
}

int64_t pg_mlp_forward_workspace_bytes(int64_t B, int32_t n_alpha) {
    if (B < 0 || n_alpha < 1) return 0;
    const int64_t clusters = (B * n_alpha + CL * T_ROWS - 1) / (CL * T_ROWS);
    return (int64_t)sizeof(int32_t) * clusters * CL;
}

int pg_mlp_forward_balanced(int precision, const pg_mlp_dims_t* dims, const void* packed, int64_t B, int32_t T, double dt,
                            int32_t n_alpha, const double* alphas, const double* x0, const double* U, const double* KK,
                            const double* kk, const double* xbar, const double* ubar, const double* ulim, double* X,
                            double* U_out, const int32_t* index, const int32_t* count, void* workspace,
                            int64_t workspace_bytes, void* stream) {
    if (B == 0) return 0;
    if (precision != PG_MLP_BF16 && precision != PG_MLP_F16X3) return PG_E_UNSUPPORTED;  // the schedule exists in the tensor-core kernels only
    if (!workspace || n_alpha < 1 || workspace_bytes < pg_mlp_forward_workspace_bytes(B, n_alpha)) return PG_E_BADARG;
    return forward_impl(precision, dims, packed, B, T, dt, n_alpha, alphas, x0, U, KK, kk, xbar, ubar, ulim, X, U_out, index,
                        count, (int32_t*)workspace, stream);
}

int pg_mlp_linearize(int precision, int lin_mode, const pg_mlp_dims_t* dims, const void* packed, int64_t B, int32_t T,
                     double dt, const double* xbar, const double* ubar, double* A, double* Bm, const int32_t* index,
                     const int32_t* count, void* stream) {
    if (B == 0 || T == 0) return 0;
    if (!dims_ok(dims) || !packed || B < 0 || T < 0 || !xbar || !ubar || !A || !Bm || !(dt > 0)) return PG_E_BADARG;
    if (lin_mode != PG_MLP_LIN_ANALYTIC && lin_mode != PG_MLP_LIN_FD) return PG_E_BADARG;
    if ((index == nullptr) != (count == nullptr)) return PG_E_BADARG;
    MlpArgs a;
    memset(&a, 0, sizeof(a));
    a.d = *dims;
    a.w = (const MlpPacked*)packed;
    a.mode = lin_mode == PG_MLP_LIN_FD ? MODE_LIN_FD : MODE_LIN_AN;
    a.R = a.Bs = B;
    a.T = T;
    a.dt = dt;
    a.n_alpha = 1;
    a.index = index;
    a.count = count;
    a.xbar = xbar;
    a.ubar = ubar;
    a.A = A;
    a.Bm = Bm;
    return launch(precision, a, (cudaStream_t)stream);
}

}  // extern "C"

#ifdef PG_MLP_TIMING
extern "C" int pg_mlp_timing_read(unsigned long long* out, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out, pg_mlp_timing_buf, sizeof(unsigned long long) * 1024 * 16);
    if (e != cudaSuccess) return (int)e;
    if (reset) {
        static unsigned long long zeros[1024 * 16];
        e = cudaMemcpyToSymbol(pg_mlp_timing_buf, zeros, sizeof(zeros));
    }
    return (int)e;
}
#endif
