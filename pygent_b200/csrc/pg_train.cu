// Training step of the learned dynamics model on the B200 (K6): see include/pygent_b200_mlp.h, "dynamics training".
//
// Reference: pygent/algorithms/mbrl.py:450-469 (`training`: MSE on standardised velocity increments, batch 512),
// :471-485 (`training_data`), :71-73 (torch.optim.Adagrad(lr, weight_decay)); network pygent/nn_models.py:179-191
// ((oDim+uDim) -> 500 -> ReLU -> 500 -> ReLU -> dxDim, fp32).
//
// One optimiser step = six launches (+ the gradient all-reduce between the last two when data parallel):
//   train_l1_kernel     gather + standardise the batch rows, layer 1 (K <= 15, CUDA cores, fp32), H1 = relu(.) written as
//                       fp16 hi/lo operand images in BOTH orientations (rows = batch: A of the forward GEMM; rows = hidden
//                       unit: B of the weight-gradient GEMM), ReLU mask bits
//   train_gemm_kernel   Z2 = H1 W2^T on the tensor cores (tcgen05.mma kind::f16, fp32 accumulators in TMEM), epilogue
//                       H2 = relu(Z2 + b2) and the layer-3 partial sums
//   train_loss_kernel   y, e = y - target, per-row squared error, dZ2 = (e W3) * [H2 > 0] as operand images (both orientations),
//                       per-row-block sums for db2 and dW3
//   train_gemm_kernel   dH1 = dZ2 W2 (epilogue: * [H1 > 0] -> dZ1, reduced over the tile's rows for db1 and dW1 without ever
//                       leaving the SM) and dW2 = dZ2^T H1 (epilogue: scaled, straight into the flat gradient bucket) --
//                       64 independent 128 x 64 tiles in ONE launch
//   train_finalize_kernel the small parameters' gradients from the row-block sums, db3 and the loss (fixed order: deterministic)
//   train_adagrad_kernel g += wd p; s += g^2; p -= lr g / (sqrt(s) + eps) on the flat 255,002-float buffers, and the updated
//                       W2 re-split into the fp16 hi/lo images the next step's GEMMs read
//
// Precision.  The reference trains in fp32.  Every GEMM operand is split x = hi + lo into two fp16 numbers (22 mantissa
// bits) and the product is formed as  Ah Bh + Ah Bl + Al Bh  with fp32 accumulation: the dropped term is 2^-22 relative.
// fp16's range (6e-8 ... 65504) fits the operands of this problem: standardised inputs, ReLU activations, weights of
// magnitude ~0.05, and error signals that are kept UNscaled ((y - t) W3, the 2/(N d) of the MSE is applied in fp32 by the
// epilogues).  Values are clamped to +-65000 before the split, so an exploding network shows up as a large loss, not NaN.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/pygent_b200_mlp.h"
#include "pg_tc.cuh"

#define PG_E_BADARG (-1)
#define PG_E_UNSUPPORTED (-2)

namespace {

using namespace pgtc;

constexpr int H = PG_MLP_HPAD;      // 512
constexpr int HID = PG_MLP_HIDDEN;  // 500
constexpr int MAXIN = PG_MLP_MAX_IN;
constexpr int MAXDX = PG_MLP_MAX_DX;
constexpr int CH_BYTES = 128 * 64 * 2;  // one operand chunk image: [128 rows x 64 k] fp16, 16 KB
constexpr int FB = H / 128;             // 4 blocks of 128 hidden units
constexpr int KBH = H / 64;             // 8 k-blocks over the hidden units
constexpr int BN = 64, NB = H / BN;     // GEMM output tile width, column tiles over the hidden units
static_assert(PG_MLP_TRAIN_MOMENTS == 48, "moments: o_mean[16] o_var[16] u_mean[2] u_var[2] dx_mean[4] dx_var[4] (+pad)");

struct Layout {  // offsets into the flat parameter / gradient buffers (torch's parameter order of NNDynamics)
    int nin, dx, oW1, ob1, oW2, ob2, oW3, ob3, NP;
};

Layout make_layout(const pg_mlp_dims_t* d) {
    Layout L;
    L.nin = d->n_obs + d->m;
    L.dx = d->dx_dim;
    L.oW1 = 0;
    L.ob1 = L.oW1 + HID * L.nin;
    L.oW2 = L.ob1 + HID;
    L.ob2 = L.oW2 + HID * HID;
    L.oW3 = L.ob2 + HID;
    L.ob3 = L.oW3 + L.dx * HID;
    L.NP = L.ob3 + L.dx;
    return L;
}

bool dims_ok(const pg_mlp_dims_t* d) {
    if (!d || d->n < 2 || d->n > 8 || (d->n & 1) || d->m < 1 || d->m > 2) return false;
    int no = 0;
    for (int i = 0; i < d->n; i++) no += d->is_angle[i] ? 2 : 1;
    return no == d->n_obs && d->n_obs + d->m + 1 <= MAXIN && d->dx_dim == d->n / 2 && d->dx_dim <= MAXDX;
}

// Workspace: operand images and fp32 intermediates of one step, for up to `rmax` (multiple of 128) batch rows.
struct Ws {
    uint8_t *W2h, *W2l, *W2Th, *W2Tl;    // [FB][KBH] chunks: rows = output unit n, k = input unit (and transposed)
    uint8_t *H1h, *H1l, *dZh, *dZl;      // [rmax/128][KBH] chunks: rows = batch row, k = hidden unit
    uint8_t *H1Th, *H1Tl, *dZTh, *dZTl;  // [FB][rmax/64] chunks: rows = hidden unit, k = batch row
    float *H2;                           // [rmax][H]
    float *ins;                          // [rmax][MAXIN] standardised network inputs
    float *tgt, *err;                    // [rmax][MAXDX] standardised targets, y - target
    float *yp;                           // [rmax][H / 64][MAXDX] layer-3 partial sums per 64-column output tile
    float *partA, *partB;                // [rmax / 128][5][H], [rmax / 128][17][H]: per-row-block sums (db2, dW3 | db1, dW1)
    float *lossrow;                      // [rmax]
    uint8_t *mask1;                      // [rmax][H / 8] bits of [H1 > 0]
    int rmax;
};

int64_t ws_layout(Ws* w, uint8_t* base, int rmax) {
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        uint8_t* p = base ? base + off : nullptr;
        off += (bytes + 1023) / 1024 * 1024;
        return p;
    };
    const int64_t wimg = (int64_t)FB * KBH * CH_BYTES, rimg = (int64_t)(rmax / 128) * KBH * CH_BYTES, timg = (int64_t)FB * (rmax / 64) * CH_BYTES;
    Ws t;
    t.W2h = take(wimg); t.W2l = take(wimg); t.W2Th = take(wimg); t.W2Tl = take(wimg);
    t.H1h = take(rimg); t.H1l = take(rimg); t.dZh = take(rimg); t.dZl = take(rimg);
    t.H1Th = take(timg); t.H1Tl = take(timg); t.dZTh = take(timg); t.dZTl = take(timg);
    t.H2 = (float*)take((int64_t)rmax * H * 4);
    t.ins = (float*)take((int64_t)rmax * MAXIN * 4);
    t.tgt = (float*)take((int64_t)rmax * MAXDX * 4);
    t.err = (float*)take((int64_t)rmax * MAXDX * 4);
    t.yp = (float*)take((int64_t)rmax * (H / 64) * MAXDX * 4);
    t.partA = (float*)take((int64_t)(rmax / 128) * (1 + MAXDX) * H * 4);
    t.partB = (float*)take((int64_t)(rmax / 128) * (1 + MAXIN) * H * 4);
    t.lossrow = (float*)take((int64_t)rmax * 4);
    t.mask1 = take((int64_t)rmax * (H / 8));
    t.rmax = rmax;
    if (w) *w = t;
    return off;
}

struct TrainArgs {
    pg_mlp_dims_t d;
    Layout L;
    Ws w;
    const float* params;
    const float* mom;
    const float *x_, *x, *o_, *u;  // data set, row-major fp32: [N][n], [N][n], [N][n_obs], [N][m]
    const int32_t* idx;            // [R] rows of the batch (NULL: rows 0..R-1)
    int R, Rpad;
    float gscale;      // 2 / (n_total * dx): d MSE / d y = gscale * (y - t)
    float loss_scale;  // 1 / (n_total * dx)
    float* bucket;     // [NP + 1]: gradient (this rank's share of the batch sum) + loss
};

__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
    __half2 h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const float a = fminf(fmaxf(v[2 * e], -65000.0f), 65000.0f), b = fminf(fmaxf(v[2 * e + 1], -65000.0f), 65000.0f);
        h[e] = __floats2half2_rn(a, b);
        const float2 hf = __half22float2(h[e]);
        l[e] = __floats2half2_rn(a - hf.x, b - hf.y);
    }
    hi = make_uint4(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]), *reinterpret_cast<uint32_t*>(&h[2]),
                    *reinterpret_cast<uint32_t*>(&h[3]));
    lo = make_uint4(*reinterpret_cast<uint32_t*>(&l[0]), *reinterpret_cast<uint32_t*>(&l[1]), *reinterpret_cast<uint32_t*>(&l[2]),
                    *reinterpret_cast<uint32_t*>(&l[3]));
}

// An 8 x 8 block (rows r0..r0+7 of the batch, hidden units k0..k0+7) into the operand images of both orientations.
//   row-major set  [rb = r / 128][kb = k / 64]: element (r, k);   transposed set [fb = k / 128][kbr = r / 64]: element (k, r)
__device__ __forceinline__ void store_block(const float (&v)[8][8], int r0, int k0, uint8_t* Mh, uint8_t* Ml, uint8_t* Th, uint8_t* Tl,
                                            int kbr_stride) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint4 hi, lo;
        split8(v[i], hi, lo);
        const int r = r0 + i;
        const size_t off = ((size_t)(r >> 7) * KBH + (k0 >> 6)) * CH_BYTES + chunk_off(r & 127, k0 & 63);
        *reinterpret_cast<uint4*>(Mh + off) = hi;
        *reinterpret_cast<uint4*>(Ml + off) = lo;
    }
#pragma unroll
    for (int c = 0; c < 8; c++) {
        float col[8];
#pragma unroll
        for (int i = 0; i < 8; i++) col[i] = v[i][c];
        uint4 hi, lo;
        split8(col, hi, lo);
        const int k = k0 + c;
        const size_t off = ((size_t)(k >> 7) * kbr_stride + (r0 >> 6)) * CH_BYTES + chunk_off(k & 127, r0 & 63);
        *reinterpret_cast<uint4*>(Th + off) = hi;
        *reinterpret_cast<uint4*>(Tl + off) = lo;
    }
}

constexpr int BLK_THREADS = 128, BLK_ROWS = 16;  // l1 kernel: 2 row groups of 8 x 64 unit groups of 8 per CTA

// ---- gather + standardise + layer 1 -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLK_THREADS) train_l1_kernel(const TrainArgs a) {
    __shared__ float ins_s[BLK_ROWS][MAXIN];
    const int t = threadIdx.x, nin = a.L.nin, n = a.d.n, nh = a.d.n / 2, dx = a.L.dx;
    const int rbase = blockIdx.x * BLK_ROWS;
    if (t < BLK_ROWS) {  // training_data (mbrl.py:471-485)
        const int r = rbase + t;
        float in[MAXIN], tg[MAXDX];
#pragma unroll
        for (int j = 0; j < MAXIN; j++) in[j] = 0.0f;
#pragma unroll
        for (int j = 0; j < MAXDX; j++) tg[j] = 0.0f;
        if (r < a.R) {
            const int64_t s = a.idx ? (int64_t)a.idx[r] : r;
            const float* mom = a.mom;
            for (int j = 0; j < a.d.n_obs; j++) in[j] = (a.o_[s * a.d.n_obs + j] - mom[j]) / mom[16 + j];
            for (int j = 0; j < a.d.m; j++) in[a.d.n_obs + j] = (a.u[s * a.d.m + j] - mom[32 + j]) / mom[34 + j];
            for (int j = 0; j < dx; j++) {
                const float dxv = a.x[s * n + nh + j] - a.x_[s * n + nh + j];
                tg[j] = (dxv - mom[36 + j]) / mom[40 + j];
            }
        }
#pragma unroll
        for (int j = 0; j < MAXIN; j++) {
            ins_s[t][j] = in[j];
            a.w.ins[(size_t)r * MAXIN + j] = in[j];
        }
#pragma unroll
        for (int j = 0; j < MAXDX; j++) a.w.tgt[(size_t)r * MAXDX + j] = tg[j];
    }
    __syncthreads();
    const int kg = t & 63, rg = t >> 6;
    const int r0 = rbase + rg * 8, k0 = kg * 8;
    float v[8][8];
    const float* W1 = a.params + a.L.oW1;
    const float* b1 = a.params + a.L.ob1;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const float b = (k0 + c) < HID ? __ldg(&b1[k0 + c]) : 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) v[i][c] = b;
    }
    for (int j = 0; j < nin; j++) {
        float wj[8], xj[8];
#pragma unroll
        for (int c = 0; c < 8; c++) wj[c] = (k0 + c) < HID ? __ldg(&W1[(k0 + c) * nin + j]) : 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) xj[i] = ins_s[rg * 8 + i][j];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int c = 0; c < 8; c++) v[i][c] = fmaf(xj[i], wj[c], v[i][c]);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t bits = 0;
        const bool live = (r0 + i) < a.R;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const bool pos = live && (k0 + c) < HID && v[i][c] > 0.0f;
            v[i][c] = pos ? v[i][c] : 0.0f;
            bits |= (pos ? 1u : 0u) << c;
        }
        a.w.mask1[(size_t)(r0 + i) * (H / 8) + kg] = (uint8_t)bits;
    }
    store_block(v, r0, k0, a.w.H1h, a.w.H1l, a.w.H1Th, a.w.H1Tl, a.w.rmax / 64);
}

// ---- output, error, loss terms, dZ2 = (e W3) * [H2 > 0], and the batch sums that need H2: db2, dW3 ---------------------------
// CTA = 128 batch rows x 64 hidden units (16 row groups x 8 unit groups of 8 x 8 blocks); grid (Rpad / 128, 8).
constexpr int LS_ACC = 1 + MAXDX;  // per hidden unit: db2, dW3[0..3]

__global__ void __launch_bounds__(128) train_loss_kernel(const TrainArgs a) {
    __shared__ float err_s[128][MAXDX];
    __shared__ float red[16][LS_ACC][64];
    const int t = threadIdx.x, dx = a.L.dx;
    const int mb = blockIdx.x, rbase = mb * 128;
    {
        const int r = rbase + t;
        float e[MAXDX], sq = 0.0f;
#pragma unroll
        for (int j = 0; j < MAXDX; j++) {
            e[j] = 0.0f;
            if (j < dx && r < a.R) {
                float y = 0.0f;
#pragma unroll
                for (int q = 0; q < NB; q++) y += a.w.yp[((size_t)r * NB + q) * MAXDX + j];
                y += __ldg(&a.params[a.L.ob3 + j]);
                e[j] = y - a.w.tgt[(size_t)r * MAXDX + j];
                sq = fmaf(e[j], e[j], sq);
            }
            err_s[t][j] = e[j];
        }
        if (blockIdx.y == 0) {
            *reinterpret_cast<float4*>(&a.w.err[(size_t)r * MAXDX]) = make_float4(e[0], e[1], e[2], e[3]);
            a.w.lossrow[r] = sq;
        }
    }
    __syncthreads();
    const int kgl = t & 7, rg = t >> 3;
    const int r0 = rbase + rg * 8, k0 = (blockIdx.y * 8 + kgl) * 8;
    const float* W3 = a.params + a.L.oW3;
    float w3[MAXDX][8];
#pragma unroll
    for (int j = 0; j < MAXDX; j++)
#pragma unroll
        for (int c = 0; c < 8; c++) w3[j][c] = (j < dx && (k0 + c) < HID) ? __ldg(&W3[j * HID + k0 + c]) : 0.0f;
    float v[8][8], acc[LS_ACC][8];
#pragma unroll
    for (int q = 0; q < LS_ACC; q++)
#pragma unroll
        for (int c = 0; c < 8; c++) acc[q][c] = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float4 ha = *reinterpret_cast<const float4*>(&a.w.H2[(size_t)(r0 + i) * H + k0]);
        const float4 hb = *reinterpret_cast<const float4*>(&a.w.H2[(size_t)(r0 + i) * H + k0 + 4]);
        const float h2[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
        float e[MAXDX];
#pragma unroll
        for (int j = 0; j < MAXDX; j++) e[j] = err_s[rg * 8 + i][j];
#pragma unroll
        for (int c = 0; c < 8; c++) {
            float g = 0.0f;
#pragma unroll
            for (int j = 0; j < MAXDX; j++) g = fmaf(e[j], w3[j][c], g);
            v[i][c] = h2[c] > 0.0f ? g : 0.0f;
            acc[0][c] += v[i][c];
#pragma unroll
            for (int j = 0; j < MAXDX; j++) acc[1 + j][c] = fmaf(e[j], h2[c], acc[1 + j][c]);
        }
    }
    store_block(v, r0, k0, a.w.dZh, a.w.dZl, a.w.dZTh, a.w.dZTl, a.w.rmax / 64);
#pragma unroll
    for (int q = 0; q < LS_ACC; q++)
#pragma unroll
        for (int c = 0; c < 8; c++) red[rg][q][kgl * 8 + c] = acc[q][c];
    __syncthreads();
    for (int o = t; o < LS_ACC * 64; o += 128) {
        const int q = o >> 6, kk = o & 63;
        float s = 0.0f;
#pragma unroll
        for (int g = 0; g < 16; g++) s += red[g][q][kk];
        a.w.partA[((size_t)mb * LS_ACC + q) * H + blockIdx.y * 64 + kk] = s;
    }
}

// ---- 128 x 64 output tiles on the tensor cores -------------------------------------------------------------------------------
// D[128 x 64] (fp32, TMEM) = sum over k-blocks of (Ah + Al)[128 x 64] (Bh + Bl)[64 x 64]^T, the three significant partial
// products as separate tcgen05.mma.kind::f16 instructions into the same accumulator.  Operands arrive as ready-made chunk
// images (A: 16 KB; B: the 8 KB half of a chunk that holds its 64 rows) by cp.async.bulk through a 4-stage mbarrier ring
// (48 KB per stage); warp 4 produces, warp 5 issues, warps 0-3 run the epilogue (thread = TMEM lane = output row).
constexpr int G_STAGES = 4, G_STAGE_BYTES = 2 * CH_BYTES + 2 * (CH_BYTES / 2), G_THREADS = 192;
constexpr int G_EXTRA = 4096;  // barriers, TMEM slot, epilogue constants
constexpr int G_SMEM = G_STAGES * G_STAGE_BYTES + 1024 + G_EXTRA;
constexpr int EP_STRIDE = BN + 1;  // the epilogue's transposition tile [128][65] floats lives in the (drained) ring
static_assert(128 * EP_STRIDE * 4 + 128 * MAXIN * 4 <= G_STAGES * G_STAGE_BYTES, "epilogue tile fits the ring");
constexpr uint32_t G_IDESC = umma_idesc_f16(0u, 128u, (uint32_t)BN);  // fp16 x fp16 -> fp32, M 128, N 64
constexpr int EPI_FWD2 = 0, EPI_DGRAD = 1, EPI_WGRAD = 2;
constexpr int DG_ACC = 1 + MAXIN;  // per hidden unit: db1, dW1[0..15]

struct GemmJob {
    const uint8_t *Ah, *Al, *Bh, *Bl;
    int a_stride, b_stride;  // chunks per 128-row block of the A / B image set
    int kblocks;             // k-blocks of 64 to contract over
    int mbs, nbs;            // tiles of 128 rows / BN columns
    int epi;
};

struct GemmArgs {
    GemmJob job[2];
    int njobs;
    TrainArgs a;
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }  // the 4 epilogue warps only

__global__ void __launch_bounds__(G_THREADS, 1) train_gemm_kernel(const GemmArgs g) {
    extern __shared__ uint8_t gsm_raw[];
    uint8_t* gsm = reinterpret_cast<uint8_t*>(((uintptr_t)gsm_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(gsm + G_STAGES * G_STAGE_BYTES);
    uint64_t* empty = full + G_STAGES;
    uint64_t* accum = empty + G_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
    float* consts = reinterpret_cast<float*>(gsm + G_STAGES * G_STAGE_BYTES + 256);  // [(1 + MAXDX)][BN]: b2, W3 rows of the tile
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    int tile = blockIdx.x, ji = 0;
    if (g.njobs > 1 && tile >= g.job[0].mbs * g.job[0].nbs) {
        tile -= g.job[0].mbs * g.job[0].nbs;
        ji = 1;
    }
    const GemmJob& J = g.job[ji];
    const int mb = tile / J.nbs, nb = tile % J.nbs;

    if (t == 0) {
        for (int s = 0; s < G_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            for (int kb = 0; kb < J.kblocks; kb++) {
                const int s = kb % G_STAGES, use = kb / G_STAGES;
                mbar_wait(&empty[s], (use & 1) ^ 1);
                mbar_expect_tx(&full[s], G_STAGE_BYTES);
                uint8_t* st = gsm + s * G_STAGE_BYTES;
                const size_t ao = ((size_t)mb * J.a_stride + kb) * CH_BYTES;
                const size_t bo = ((size_t)(nb >> 1) * J.b_stride + kb) * CH_BYTES + (nb & 1) * (CH_BYTES / 2);  // rows 64 (nb & 1) .. + 63
                bulk_g2s(st, J.Ah + ao, CH_BYTES, &full[s]);
                bulk_g2s(st + CH_BYTES, J.Al + ao, CH_BYTES, &full[s]);
                bulk_g2s(st + 2 * CH_BYTES, J.Bh + bo, CH_BYTES / 2, &full[s]);
                bulk_g2s(st + 2 * CH_BYTES + CH_BYTES / 2, J.Bl + bo, CH_BYTES / 2, &full[s]);
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        if (lane == 0) {
            for (int kb = 0; kb < J.kblocks; kb++) {
                const int s = kb % G_STAGES, use = kb / G_STAGES;
                mbar_wait(&full[s], use & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = smem_u32(gsm + s * G_STAGE_BYTES);
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {  // K = 16 per instruction = 32 bytes along the swizzled row
                    const uint64_t ah = umma_desc_sw128(st + ks * 32), al = umma_desc_sw128(st + CH_BYTES + ks * 32);
                    const uint64_t bh = umma_desc_sw128(st + 2 * CH_BYTES + ks * 32);
                    const uint64_t bl = umma_desc_sw128(st + 2 * CH_BYTES + CH_BYTES / 2 + ks * 32);
                    umma_f16_1sm(tmem_base, al, bh, G_IDESC, (kb > 0 || ks > 0) ? 1u : 0u);  // small terms first
                    umma_f16_1sm(tmem_base, ah, bl, G_IDESC, 1u);
                    umma_f16_1sm(tmem_base, ah, bh, G_IDESC, 1u);
                }
                umma_commit_1sm(&empty[s]);
            }
            umma_commit_1sm(accum);
        }
        __syncwarp();
    } else {
        const TrainArgs& a = g.a;
        const int rl = warp * 32 + lane;       // TMEM lane = row of the tile
        const int row = mb * 128 + rl;         // global output row of this thread
        const int n0 = nb * BN;                // first output column of the tile
        const uint32_t tbase = tmem_base + ((uint32_t)(warp * 32) << 16);
        float* tile_s = reinterpret_cast<float*>(gsm);                 // [128][EP_STRIDE], valid once the ring has drained
        float* ins_s = tile_s + 128 * EP_STRIDE;                       // [128][MAXIN]
        if (J.epi == EPI_FWD2) {  // stage the tile's b2 and W3 columns while the MMAs run
            if (rl < BN) {
                const int n = n0 + rl;
                consts[rl] = n < HID ? __ldg(&a.params[a.L.ob2 + n]) : 0.0f;
#pragma unroll
                for (int j = 0; j < MAXDX; j++) consts[(1 + j) * BN + rl] = (j < a.L.dx && n < HID) ? __ldg(&a.params[a.L.oW3 + j * HID + n]) : 0.0f;
            }
            epi_bar();
        }
        mbar_wait(accum, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (J.epi == EPI_FWD2) {
            float y[MAXDX] = {0.0f, 0.0f, 0.0f, 0.0f};
            for (int c4 = 0; c4 < BN / 32; c4++) {
                uint32_t v[32];
                tmem_ld32(tbase + c4 * 32, v);
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                    float o[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int cl = c4 * 32 + c + e;
                        const float h2 = (n0 + cl) < HID ? fmaxf(__uint_as_float(v[c + e]) + consts[cl], 0.0f) : 0.0f;
#pragma unroll
                        for (int j = 0; j < MAXDX; j++) y[j] = fmaf(h2, consts[(1 + j) * BN + cl], y[j]);
                        o[e] = h2;
                    }
                    *reinterpret_cast<float4*>(&a.w.H2[(size_t)row * H + n0 + c4 * 32 + c]) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
            *reinterpret_cast<float4*>(&a.w.yp[((size_t)row * NB + nb) * MAXDX]) = make_float4(y[0], y[1], y[2], y[3]);
        } else if (J.epi == EPI_DGRAD) {
            // dZ1 = dH1 * [H1 > 0]; db1[k] = sum_r dZ1[r][k], dW1[k][i] = sum_r dZ1[r][k] in[r][i]: transpose through smem,
            // then thread = hidden unit sums the tile's 128 rows (fixed order: deterministic)
            const uint32_t* m1 = reinterpret_cast<const uint32_t*>(a.w.mask1 + (size_t)row * (H / 8));
#pragma unroll
            for (int q = 0; q < MAXIN / 4; q++)
                *reinterpret_cast<float4*>(&ins_s[rl * MAXIN + q * 4]) = *reinterpret_cast<const float4*>(&a.w.ins[(size_t)row * MAXIN + q * 4]);
            for (int c4 = 0; c4 < BN / 32; c4++) {
                uint32_t v[32];
                tmem_ld32(tbase + c4 * 32, v);
                const uint32_t mw = m1[nb * (BN / 32) + c4];
#pragma unroll
                for (int c = 0; c < 32; c++) tile_s[rl * EP_STRIDE + c4 * 32 + c] = ((mw >> c) & 1u) ? __uint_as_float(v[c]) : 0.0f;
            }
            epi_bar();
            if (rl < BN) {
                float acc[DG_ACC];
#pragma unroll
                for (int i = 0; i < DG_ACC; i++) acc[i] = 0.0f;
                for (int r = 0; r < 128; r++) {
                    const float d = tile_s[r * EP_STRIDE + rl];
                    acc[0] += d;
#pragma unroll
                    for (int q = 0; q < MAXIN / 4; q++) {
                        if (q * 4 < a.L.nin) {
                            const float4 in = *reinterpret_cast<const float4*>(&ins_s[r * MAXIN + q * 4]);
                            acc[1 + q * 4 + 0] = fmaf(d, in.x, acc[1 + q * 4 + 0]);
                            acc[1 + q * 4 + 1] = fmaf(d, in.y, acc[1 + q * 4 + 1]);
                            acc[1 + q * 4 + 2] = fmaf(d, in.z, acc[1 + q * 4 + 2]);
                            acc[1 + q * 4 + 3] = fmaf(d, in.w, acc[1 + q * 4 + 3]);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < DG_ACC; i++) a.w.partB[((size_t)mb * DG_ACC + i) * H + n0 + rl] = acc[i];
            }
        } else {  // EPI_WGRAD: row = output unit n of W2, columns = input units k
            float* G = a.bucket + a.L.oW2;
            for (int c4 = 0; c4 < BN / 32; c4++) {
                uint32_t v[32];
                tmem_ld32(tbase + c4 * 32, v);
                const int k0 = n0 + c4 * 32;
                if (row < HID) {
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        if (k0 + c < HID)  // HID is a multiple of 4: whole float4 groups
                            *reinterpret_cast<float4*>(&G[(size_t)row * HID + k0 + c]) =
                                make_float4(a.gscale * __uint_as_float(v[c]), a.gscale * __uint_as_float(v[c + 1]),
                                            a.gscale * __uint_as_float(v[c + 2]), a.gscale * __uint_as_float(v[c + 3]));
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_base) : "memory");
}

// ---- the small parameters' gradients from the per-row-block partial sums, db3 and the loss (fixed order: deterministic) ------
__global__ void __launch_bounds__(256) train_finalize_kernel(const TrainArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nin = a.L.nin, dx = a.L.dx, rbs = a.Rpad / 128;
    const int slots = LS_ACC + DG_ACC;  // per hidden unit: db2, dW3[4], db1, dW1[16]
    if (i < slots * HID) {
        const int q = i / HID, k = i % HID;
        float s = 0.0f;
        if (q < LS_ACC) {
            for (int mb = 0; mb < rbs; mb++) s += a.w.partA[((size_t)mb * LS_ACC + q) * H + k];
        } else {
            for (int mb = 0; mb < rbs; mb++) s += a.w.partB[((size_t)mb * DG_ACC + (q - LS_ACC)) * H + k];
        }
        s *= a.gscale;
        if (q == 0) a.bucket[a.L.ob2 + k] = s;
        else if (q < LS_ACC) { if (q - 1 < dx) a.bucket[a.L.oW3 + (q - 1) * HID + k] = s; }
        else if (q == LS_ACC) a.bucket[a.L.ob1 + k] = s;
        else if (q - LS_ACC - 1 < nin) a.bucket[a.L.oW1 + k * nin + (q - LS_ACC - 1)] = s;
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {  // db3 and the loss: lane-strided sums, then a fixed shuffle tree
        const int t = threadIdx.x;
        float s[MAXDX + 1];
#pragma unroll
        for (int j = 0; j <= MAXDX; j++) s[j] = 0.0f;
        for (int r = t; r < a.R; r += 32) {
            const float4 e4 = *reinterpret_cast<const float4*>(&a.w.err[(size_t)r * MAXDX]);
            s[0] += e4.x; s[1] += e4.y; s[2] += e4.z; s[3] += e4.w;
            s[MAXDX] += a.w.lossrow[r];
        }
#pragma unroll
        for (int j = 0; j <= MAXDX; j++)
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) s[j] += __shfl_xor_sync(0xffffffffu, s[j], d);
        if (t == 0) {
            for (int j = 0; j < dx; j++) a.bucket[a.L.ob3 + j] = a.gscale * s[j];
            a.bucket[a.L.NP] = a.loss_scale * s[MAXDX];
        }
    }
}

// ---- Adagrad on the flat buffers + re-split of W2 into the operand images ---------------------------------------------------------
__device__ __forceinline__ void pack_w2(const Ws& w, int n, int k, float p) {
    const float c = fminf(fmaxf(p, -65000.0f), 65000.0f);
    const __half hi = __float2half_rn(c), lo = __float2half_rn(c - __half2float(hi));
    const size_t o = ((size_t)(n >> 7) * KBH + (k >> 6)) * CH_BYTES + chunk_off(n & 127, k & 63);   // rows n, contraction k
    const size_t oT = ((size_t)(k >> 7) * KBH + (n >> 6)) * CH_BYTES + chunk_off(k & 127, n & 63);  // rows k, contraction n
    *reinterpret_cast<__half*>(w.W2h + o) = hi;
    *reinterpret_cast<__half*>(w.W2l + o) = lo;
    *reinterpret_cast<__half*>(w.W2Th + oT) = hi;
    *reinterpret_cast<__half*>(w.W2Tl + oT) = lo;
}

__global__ void __launch_bounds__(256) train_pack_kernel(const float* params, Layout L, Ws w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < HID * HID) pack_w2(w, i / HID, i % HID, params[L.oW2 + i]);
}

// torch.optim.Adagrad (lr_decay = 0, initial_accumulator_value = 0): g += wd p; s += g g; p -= lr g / (sqrt(s) + eps)
__global__ void __launch_bounds__(256) train_adagrad_kernel(float* params, float* state_sum, const float* bucket, Layout L, Ws w, float lr,
                                                            float wd, float eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.NP) return;
    float p = params[i];
    const float gr = fmaf(wd, p, bucket[i]);
    const float s = fmaf(gr, gr, state_sum[i]);
    p = p - lr * (gr / (sqrtf(s) + eps));
    state_sum[i] = s;
    params[i] = p;
    if (i >= L.oW2 && i < L.ob2) pack_w2(w, (i - L.oW2) / HID, (i - L.oW2) % HID, p);
}

// ---- data-parallel exchange over NVLink peer memory, fused with the optimiser --------------------------------------------------
// Every rank owns an exchange block in its own HBM, mapped into all ranks (CUDA IPC): two gradient buckets [NPAD] (step parity)
// and a flag word per rank.  After its gradient kernels a rank stores (step + 1) into ITS slot of every peer's flag array; the
// update kernel waits until all of its own slots show the step, then every thread sums element i of the WORLD buckets -- peer
// loads through NVLink / NVSwitch, in rank order, so all ranks form bit-identical sums -- and applies Adagrad at once: one
// kernel is the all-reduce and the optimiser, and the gradients cross the links exactly once (world - 1 reads per element, no
// reduce-scatter / all-gather round trips, no host in the loop).  Reuse of a parity slot is safe: a rank writes the bucket of
// step s + 2 only after its update of step s + 1, which waited for every peer's gradient of step s + 1, which in stream order
// follows that peer's update (the reader) of step s.
constexpr int P2P_MAX_WORLD = 16;

struct P2PTable {
    const float* bucket[P2P_MAX_WORLD];  // peer r's exchange block (bucket of parity p at + p * npad)
    uint32_t* flags[P2P_MAX_WORLD];      // peer r's flag array [P2P_MAX_WORLD]
    int world, rank;
    int64_t npad;
};

__global__ void train_p2p_signal_kernel(P2PTable t, uint32_t step1) {
    const int r = threadIdx.x;
    if (r < t.world) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(t.flags[r] + t.rank), "r"(step1) : "memory");
    }
}

__global__ void __launch_bounds__(256) train_adagrad_p2p_kernel(float* params, float* state_sum, float* loss_out, int32_t* status, Layout L, Ws w,
                                                                P2PTable t, uint32_t step1, float lr, float wd, float eps) {
    __shared__ int ok;
    if (threadIdx.x == 0) {  // wait until every rank has published its gradient of this step
        int good = 1;
        unsigned long long t0 = 0;
        for (int r = 0; r < t.world && good; r++) {
            const uint32_t* f = t.flags[t.rank] + r;
            for (uint32_t spins = 0;; spins++) {
                uint32_t v;
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                if ((int32_t)(v - step1) >= 0) break;
                if ((spins & 1023) == 1023) {  // a rank that never arrives must not hang the GPU: give up after ~4 s
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > 4000000000ull) { good = 0; break; }
                }
            }
        }
        ok = good;
    }
    __syncthreads();
    if (!ok) {
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicExch(status, 1);
        return;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > L.NP) return;
    const size_t off = (size_t)((step1 - 1) & 1) * t.npad + i;
    float g = 0.0f;
    for (int r = 0; r < t.world; r++) g += __ldcv(t.bucket[r] + off);  // fixed rank order: identical on every rank
    if (i == L.NP) {  // the loss rides in the bucket's last slot
        *loss_out = g;
        return;
    }
    float p = params[i];
    const float gr = fmaf(wd, p, g);
    const float s = fmaf(gr, gr, state_sum[i]);
    p = p - lr * (gr / (sqrtf(s) + eps));
    state_sum[i] = s;
    params[i] = p;
    if (i >= L.oW2 && i < L.ob2) pack_w2(w, (i - L.oW2) / HID, (i - L.oW2) % HID, p);
}

int check_last() {
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace

extern "C" {

int64_t pg_mlp_train_param_count(const pg_mlp_dims_t* dims) { return dims_ok(dims) ? make_layout(dims).NP : PG_E_BADARG; }

int64_t pg_mlp_train_workspace_bytes(int32_t max_rows) {
    if (max_rows < 1) return 0;
    return ws_layout(nullptr, nullptr, (max_rows + 127) / 128 * 128);
}

int pg_mlp_train_init(const pg_mlp_dims_t* dims, const float* params, void* workspace, int64_t workspace_bytes, int32_t max_rows,
                      void* stream) {
    if (!dims_ok(dims) || !params || !workspace || max_rows < 1) return PG_E_BADARG;
    if (((uintptr_t)workspace & 1023) != 0) return PG_E_BADARG;
    const int rmax = (max_rows + 127) / 128 * 128;
    Ws w;
    if (ws_layout(&w, (uint8_t*)workspace, rmax) > workspace_bytes) return PG_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(workspace, 0, (size_t)ws_layout(nullptr, nullptr, rmax), st);  // padding of every image is zero
    if (e != cudaSuccess) return (int)e;
    train_pack_kernel<<<(HID * HID + 255) / 256, 256, 0, st>>>(params, make_layout(dims), w);
    return check_last();
}

int pg_mlp_train_grad(const pg_mlp_dims_t* dims, const float* params, const float* moments, const float* x_, const float* x,
                      const float* o_, const float* u, const int32_t* idx, int32_t R, int64_t n_total, void* workspace,
                      int64_t workspace_bytes, int32_t max_rows, float* bucket, void* stream) {
    if (!dims_ok(dims) || !params || !moments || !x_ || !x || !o_ || !u || !workspace || !bucket || R < 0 || n_total < 1 ||
        R > max_rows)
        return PG_E_BADARG;
    const int rmax = (max_rows + 127) / 128 * 128;
    TrainArgs a;
    memset(&a, 0, sizeof(a));
    if (ws_layout(&a.w, (uint8_t*)workspace, rmax) > workspace_bytes) return PG_E_BADARG;
    a.d = *dims;
    a.L = make_layout(dims);
    a.params = params;
    a.mom = moments;
    a.x_ = x_; a.x = x; a.o_ = o_; a.u = u;
    a.idx = idx;
    a.R = R;
    a.Rpad = R > 0 ? (R + 127) / 128 * 128 : 128;
    a.gscale = (float)(2.0 / ((double)n_total * a.L.dx));
    a.loss_scale = (float)(1.0 / ((double)n_total * a.L.dx));
    a.bucket = bucket;
    cudaStream_t st = (cudaStream_t)stream;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(train_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    const int rbs = a.Rpad / 128;
    train_l1_kernel<<<a.Rpad / BLK_ROWS, BLK_THREADS, 0, st>>>(a);
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.a = a;
    g.njobs = 1;
    g.job[0] = GemmJob{a.w.H1h, a.w.H1l, a.w.W2h, a.w.W2l, KBH, KBH, KBH, rbs, NB, EPI_FWD2};
    train_gemm_kernel<<<rbs * NB, G_THREADS, G_SMEM, st>>>(g);
    train_loss_kernel<<<dim3(rbs, H / 64), 128, 0, st>>>(a);
    g.njobs = 2;
    g.job[0] = GemmJob{a.w.dZh, a.w.dZl, a.w.W2Th, a.w.W2Tl, KBH, KBH, KBH, rbs, NB, EPI_DGRAD};
    g.job[1] = GemmJob{a.w.dZTh, a.w.dZTl, a.w.H1Th, a.w.H1Tl, rmax / 64, rmax / 64, a.Rpad / 64, FB, NB, EPI_WGRAD};
    train_gemm_kernel<<<rbs * NB + FB * NB, G_THREADS, G_SMEM, st>>>(g);
    train_finalize_kernel<<<((LS_ACC + DG_ACC) * HID + 255) / 256, 256, 0, st>>>(a);
    return check_last();
}

int pg_mlp_train_update(const pg_mlp_dims_t* dims, float* params, float* state_sum, const float* bucket, float lr, float weight_decay,
                        float eps, void* workspace, int64_t workspace_bytes, int32_t max_rows, void* stream) {
    if (!dims_ok(dims) || !params || !state_sum || !bucket || !workspace || max_rows < 1) return PG_E_BADARG;
    const int rmax = (max_rows + 127) / 128 * 128;
    Ws w;
    if (ws_layout(&w, (uint8_t*)workspace, rmax) > workspace_bytes) return PG_E_BADARG;
    const Layout L = make_layout(dims);
    train_adagrad_kernel<<<(L.NP + 255) / 256, 256, 0, (cudaStream_t)stream>>>(params, state_sum, bucket, L, w, lr, weight_decay, eps);
    return check_last();
}

int64_t pg_mlp_train_p2p_bytes(const pg_mlp_dims_t* dims) {
    if (!dims_ok(dims)) return PG_E_BADARG;
    const int64_t npad = (make_layout(dims).NP + 1 + 255) / 256 * 256;
    return 2 * npad * 4 + 1024;  // two buckets + the flag words (+ a status word)
}

int pg_mlp_train_p2p_alloc(int64_t bytes, void** ptr, void* ipc_handle_out) {
    if (bytes <= 0 || !ptr || !ipc_handle_out) return PG_E_BADARG;
    cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*ptr, 0, (size_t)bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaIpcGetMemHandle((cudaIpcMemHandle_t*)ipc_handle_out, *ptr);
    return e == cudaSuccess ? 0 : (int)e;
}

int pg_mlp_train_p2p_open(const void* ipc_handle, void** ptr) {
    if (!ipc_handle || !ptr) return PG_E_BADARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    const cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    return e == cudaSuccess ? 0 : (int)e;
}

int pg_mlp_train_p2p_close(void* ptr, int32_t own) {
    if (!ptr) return 0;
    const cudaError_t e = own ? cudaFree(ptr) : cudaIpcCloseMemHandle(ptr);
    return e == cudaSuccess ? 0 : (int)e;
}

int pg_mlp_train_update_p2p(const pg_mlp_dims_t* dims, float* params, float* state_sum, void* const* blocks, int32_t world, int32_t rank,
                            int64_t step, float lr, float weight_decay, float eps, float* loss_out, void* workspace,
                            int64_t workspace_bytes, int32_t max_rows, void* stream) {
    if (!dims_ok(dims) || !params || !state_sum || !blocks || world < 1 || world > P2P_MAX_WORLD || rank < 0 || rank >= world || step < 0 ||
        !loss_out || !workspace || max_rows < 1)
        return PG_E_BADARG;
    const int rmax = (max_rows + 127) / 128 * 128;
    Ws w;
    if (ws_layout(&w, (uint8_t*)workspace, rmax) > workspace_bytes) return PG_E_BADARG;
    const Layout L = make_layout(dims);
    P2PTable t;
    memset(&t, 0, sizeof(t));
    t.world = world;
    t.rank = rank;
    t.npad = (L.NP + 1 + 255) / 256 * 256;
    for (int r = 0; r < world; r++) {
        if (!blocks[r]) return PG_E_BADARG;
        t.bucket[r] = (const float*)blocks[r];
        t.flags[r] = (uint32_t*)((char*)blocks[r] + 2 * t.npad * 4);
    }
    int32_t* status = (int32_t*)(t.flags[rank] + 64);
    const uint32_t step1 = (uint32_t)(step + 1);
    cudaStream_t st = (cudaStream_t)stream;
    train_p2p_signal_kernel<<<1, 32, 0, st>>>(t, step1);
    train_adagrad_p2p_kernel<<<(L.NP + 1 + 255) / 256, 256, 0, st>>>(params, state_sum, loss_out, status, L, w, t, step1, lr, weight_decay, eps);
    return check_last();
}

int pg_mlp_train_step(const pg_mlp_dims_t* dims, float* params, float* state_sum, const float* moments, const float* x_,
                      const float* x, const float* o_, const float* u, const int32_t* idx, int32_t R, float lr, float weight_decay,
                      float eps, void* workspace, int64_t workspace_bytes, int32_t max_rows, float* bucket, void* stream) {
    const int rc = pg_mlp_train_grad(dims, params, moments, x_, x, o_, u, idx, R, R, workspace, workspace_bytes, max_rows, bucket, stream);
    if (rc) return rc;
    return pg_mlp_train_update(dims, params, state_sum, bucket, lr, weight_decay, eps, workspace, workspace_bytes, max_rows, stream);
}

}  // extern "C"
