// Hand-written sm_100a kernels of the simulation-and-planning hot path, generic over the generated
// `Problem` (dynamics f, Jacobians A/B, cost and its expansion) and the scalar type T.
//
// Work decomposition: ONE THREAD OWNS ONE TRAJECTORY (forward kernels: one thread per
// (trajectory, line-search alpha)).  The n<=8 state, the m<=2 control, the Jacobians and the Riccati
// matrices live in registers; all loops over state indices are fully unrolled and skip structural
// zeros of A, B and the cost Hessians at compile time.  Arrays are trajectory-minor
// ([..., B] with b fastest), so every global access of a warp is one or two full 128-byte lines.
// The per-step nominal trajectory and gains are prefetched one control interval ahead into registers,
// which hides the L2/HBM latency behind the ~1-2k cycle RK4 step.
#pragma once
#include <stdint.h>

#include "pg_math.cuh"

namespace pgk {

constexpr int INTEG_RK4 = 0;
constexpr int INTEG_EULER = 1;

template <typename T, int N>
__device__ __forceinline__ void copy(T (&d)[N], const T (&s)[N]) {
#pragma unroll
    for (int i = 0; i < N; i++) d[i] = s[i];
}

// ---- integrators --------------------------------------------------------------------------------
// Classical RK4 with `S` substeps of h = dt/S and the control held constant: the fixed-step parity
// schedule for StateSpaceModel.step (pygent/environments.py:263).
template <typename P, typename T>
__device__ __forceinline__ void rk4_substep(const typename P::template Ctx<T>& m, T (&x)[P::N], const T (&u)[P::M], const T h) {
    constexpr int N = P::N;
    T k1[N], k2[N], k3[N], k4[N], xt[N];
    const T hh = T(0.5) * h;
    P::f(m, x, u, k1);
#pragma unroll
    for (int i = 0; i < N; i++) xt[i] = x[i] + hh * k1[i];
    P::f(m, xt, u, k2);
#pragma unroll
    for (int i = 0; i < N; i++) xt[i] = x[i] + hh * k2[i];
    P::f(m, xt, u, k3);
#pragma unroll
    for (int i = 0; i < N; i++) xt[i] = x[i] + h * k3[i];
    P::f(m, xt, u, k4);
    const T h6 = h / T(6);
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = x[i] + h6 * (k1[i] + T(2) * k2[i] + T(2) * k3[i] + k4[i]);
}

// The same substep for systems whose sin/cos arguments are all affine in the state (Problem::NTRIG > 0, every
// shipped model): (sn, cs) hold sin/cos of the arguments at x (the anchor); the three intermediate stages and,
// with ADVANCE, the new anchor at x_{n+1} are rotations of it through the argument increment w*k (the arguments
// are affine in x; the angle component of the stage state itself is then dead code).
// `big` collects the largest increment magnitude (see MathCtx::rotate).
// E trajectories per thread, interleaved stage by stage: the FP64 pipe of the B200 reaches its peak only with
// independent instructions from the SAME warp (profiles/fp_peak_b200.json: ILP 1 saturates at 72% of peak however
// many warps are resident, ILP 4 at 100% with one warp per scheduler), and one RK4 chain has little ILP.
template <typename P, typename T, int E, bool ADVANCE>
__device__ __forceinline__ void rk4_substep_anchored(const typename P::template Ctx<T>& m, T (&x)[E][P::N], const T (&u)[E][P::M], const T h,
                                                     T (&sn)[E][P::NTRIG_], T (&cs)[E][P::NTRIG_], int& big) {
    constexpr int N = P::N, NT = P::NTRIG_;
    T k1[E][N], k2[E][N], k3[E][N], k4[E][N], xt[E][N], d[E][NT], s1[E][NT], c1[E][NT];
    const T hh = T(0.5) * h;
    auto stage = [&](const T (&k)[E][N], const T w) {
#pragma unroll
        for (int e = 0; e < E; e++) {
#pragma unroll
            for (int i = 0; i < N; i++) xt[e][i] = x[e][i] + w * k[e][i];
            P::trig_inc(m, k[e], w, d[e]);
        }
#pragma unroll
        for (int e = 0; e < E; e++)
#pragma unroll
            for (int j = 0; j < NT; j++) {
                big = max(big, m.mag(d[e][j]));
                m.rotate(d[e][j], sn[e][j], cs[e][j], s1[e][j], c1[e][j]);
            }
    };
#pragma unroll
    for (int e = 0; e < E; e++) P::f_trig(m, x[e], u[e], sn[e], cs[e], k1[e]);
    stage(k1, hh);
#pragma unroll
    for (int e = 0; e < E; e++) P::f_trig(m, xt[e], u[e], s1[e], c1[e], k2[e]);
    stage(k2, hh);
#pragma unroll
    for (int e = 0; e < E; e++) P::f_trig(m, xt[e], u[e], s1[e], c1[e], k3[e]);
    stage(k3, h);
#pragma unroll
    for (int e = 0; e < E; e++) P::f_trig(m, xt[e], u[e], s1[e], c1[e], k4[e]);
    const T h6 = h / T(6);
    T ks[E][N];
#pragma unroll
    for (int e = 0; e < E; e++)
#pragma unroll
        for (int i = 0; i < N; i++) {
            ks[e][i] = k1[e][i] + T(2) * k2[e][i] + T(2) * k3[e][i] + k4[e][i];
            xt[e][i] = x[e][i] + h6 * ks[e][i];
        }
    if constexpr (ADVANCE) {
#pragma unroll
        for (int e = 0; e < E; e++) P::trig_inc(m, ks[e], h6, d[e]);
#pragma unroll
        for (int e = 0; e < E; e++)
#pragma unroll
            for (int j = 0; j < NT; j++) {
                big = max(big, m.mag(d[e][j]));
                m.rotate(d[e][j], sn[e][j], cs[e][j], sn[e][j], cs[e][j]);
            }
    }
#pragma unroll
    for (int e = 0; e < E; e++) copy(x[e], xt[e]);
}

template <typename P, typename T>
__device__ __forceinline__ void rk4_interval_plain(const typename P::template Ctx<T>& m, T (&x)[P::N], const T (&u)[P::M], const T h, const int S) {
    if (S == 4) {  // the parity schedule: fully unrolled, no loop-carried register copies
#pragma unroll
        for (int s = 0; s < 4; s++) rk4_substep<P, T>(m, x, u, h);
    } else {
        for (int s = 0; s < S; s++) rk4_substep<P, T>(m, x, u, h);
    }
}

// One control interval of E trajectories held by one thread.  INTEG_EULER is StateSpaceModel.fast_step
// (environments.py:313-315).
template <typename P, typename T, int INTEG, int E>
__device__ __forceinline__ void env_step_multi(const typename P::template Ctx<T>& m, T (&x)[E][P::N], const T (&u)[E][P::M], const T dt, const T h,
                                               const int S) {
    if constexpr (INTEG == INTEG_EULER) {
#pragma unroll
        for (int e = 0; e < E; e++) {
            T k1[P::N];
            P::f(m, x[e], u[e], k1);
#pragma unroll
            for (int i = 0; i < P::N; i++) x[e][i] = x[e][i] + dt * k1[i];
        }
    } else if constexpr (P::NTRIG > 0) {
        // anchor: one fresh sin/cos per argument and control interval, then rotations (15 of the 16 evaluations
        // of the S = 4 schedule).  Speculative: if any increment was outside the rotation's domain, the interval
        // is redone from the saved state with fresh evaluations (never on the example horizons; large |thetadot|).
        T x0[E][P::N], sn[E][P::NTRIG_], cs[E][P::NTRIG_], a[E][P::NTRIG_];
#pragma unroll
        for (int e = 0; e < E; e++) {
            copy(x0[e], x[e]);
            P::trig_arg(m, x[e], a[e]);
        }
#pragma unroll
        for (int e = 0; e < E; e++)
#pragma unroll
            for (int j = 0; j < P::NTRIG_; j++) m.sincos(a[e][j], sn[e][j], cs[e][j]);
        int big = 0;
        if (S == 4) {
            rk4_substep_anchored<P, T, E, true>(m, x, u, h, sn, cs, big);
            rk4_substep_anchored<P, T, E, true>(m, x, u, h, sn, cs, big);
            rk4_substep_anchored<P, T, E, true>(m, x, u, h, sn, cs, big);
            rk4_substep_anchored<P, T, E, false>(m, x, u, h, sn, cs, big);
        } else {
            for (int s = 0; s + 1 < S; s++) rk4_substep_anchored<P, T, E, true>(m, x, u, h, sn, cs, big);
            rk4_substep_anchored<P, T, E, false>(m, x, u, h, sn, cs, big);
        }
        if (big >= pg::MathCtx<T>::ROT_LIMIT) {
#pragma unroll
            for (int e = 0; e < E; e++) {
                copy(x[e], x0[e]);
                rk4_interval_plain<P, T>(m, x[e], u[e], h, S);
            }
        }
    } else {
#pragma unroll
        for (int e = 0; e < E; e++) rk4_interval_plain<P, T>(m, x[e], u[e], h, S);
    }
}

template <typename P, typename T, int INTEG>
__device__ __forceinline__ void env_step(const typename P::template Ctx<T>& m, T (&x)[P::N], const T (&u)[P::M], const T dt, const T h, const int S) {
    env_step_multi<P, T, INTEG, 1>(m, reinterpret_cast<T(&)[1][P::N]>(x), reinterpret_cast<const T(&)[1][P::M]>(u), dt, h, S);
}

// ---- in-kernel action stream ("u[k,b] ~ U(-uMax, uMax) ... or in-kernel Philox with the same counter scheme,
// validated against the host stream") --------------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11): counter
// (b mod 2^32, b div 2^32, control interval k, input j), key = the 64-bit seed.  The first two output words make one
// 53-bit uniform U in [0, 1); the action is (2 U - 1) * uMax, all in IEEE double (exact up to the two final roundings), so
// pygent_b200/random_actions.py reproduces the stream bit for bit on the host.  The fp32 kernels use the top 24 bits of the
// same draw and float arithmetic (random_actions(..., dtype=np.float32) on the host).
// Philox2x32-10 (Salmon et al., SC'11): ONE 64-bit draw per block -- exactly what an action needs -- for half the integer
// work of Philox4x32 (10 rounds of one 32 x 32 -> 64 multiply, two xors, one add); the block of four words cost the rollout
// kernel 19 % of its instructions.
__host__ __device__ __forceinline__ void philox2x32_10(uint32_t c0, uint32_t c1, uint32_t key, uint32_t& r0, uint32_t& r1) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const uint64_t p = (uint64_t)0xD256D193u * (uint64_t)c0;
        c0 = (uint32_t)(p >> 32) ^ key ^ c1;
        c1 = (uint32_t)p;
        key += 0x9E3779B9u;
    }
    r0 = c0;
    r1 = c1;
}

// key of a launch's stream from the caller's 64-bit seed (host side, pg_abi.cu): the first word of the block (seed_lo, seed_hi)
// under the fixed key 0x5047454E ("PGEN")
__host__ __device__ __forceinline__ uint32_t philox_stream_key(uint64_t seed) {
    uint32_t r0, r1;
    philox2x32_10((uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), 0x5047454Eu, r0, r1);
    return r0;
}

// action j of environment b (< 2^32) in control interval k (< 2^30): the block of counter (b, 4 k + j) under the stream key
template <typename T>
__device__ __forceinline__ T philox_action(uint32_t key, uint32_t /*unused*/, int64_t b, int k, int j, double u_max) {
    uint32_t r0, r1;
    philox2x32_10((uint32_t)((uint64_t)b & 0xffffffffu), ((uint32_t)k << 2) | (uint32_t)j, key, r0, r1);
    if constexpr (sizeof(T) == 8) {
        const uint64_t bits = (((uint64_t)r1 << 32) | (uint64_t)r0) >> 11;
        const double U = (double)bits * 1.1102230246251565e-16;  // 2^-53
        return (T)((2.0 * U - 1.0) * u_max);
    } else {  // the fp32 instantiation holds no fp64 arithmetic: the top 24 bits of the same 64-bit draw, in float
        const float U = (float)(r1 >> 8) * 5.9604644775390625e-08f;  // 2^-24
        return (T)((2.0f * U - 1.0f) * (float)u_max);
    }
}

// ---- K1: open-loop rollout ----------------------------------------------------------------------
template <typename T>
struct RolloutArgs {
    int64_t B;
    int T_steps, S;
    T dt, t0, terminal_cost;
    int add_final_cost, accumulate_cost;
    int rng;                      // 1: controls come from the Philox stream (U is ignored)
    uint32_t seed_lo, seed_hi;    // seed_lo: the stream key (philox_stream_key(seed), computed on the host); seed_hi unused
    int k_offset;                 // control interval of the stream that this launch's interval 0 is
    double u_max[2];
    const T* x0;
    const T* U;
    T* X;
    T* xN;
    T* cost;
    uint8_t* terminated;
};

// control j of trajectory b in control interval k: the action stream, the U array, or zero
template <typename T>
__device__ __forceinline__ T control_in(const RolloutArgs<T>& a, int k, int j, int64_t b, int M) {
    if (a.rng) return philox_action<T>(a.seed_lo, a.seed_hi, b, a.k_offset + k, j, a.u_max[j]);
    return a.U ? a.U[((int64_t)k * M + j) * a.B + b] : T(0);
}

template <typename P, typename T, int INTEG>
__global__ void __launch_bounds__(128) rollout_kernel(const RolloutArgs<T> a) {
    const typename P::template Ctx<T> m;
    constexpr int N = P::N, M = P::M;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const int64_t B = a.B;
    T x[N], u[M], un[M];
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = a.x0[i * B + b];
    if (a.X) {
#pragma unroll
        for (int i = 0; i < N; i++) a.X[i * B + b] = x[i];
    }
    const T h = a.dt / T(a.S);
    // accumulate_cost: this launch continues a rollout split into time segments (cost[b] carries the sum)
    T cost = (a.accumulate_cost && a.cost) ? a.cost[b] : T(0), t = a.t0;
    bool term = false;
#pragma unroll
    for (int j = 0; j < M; j++) un[j] = a.T_steps > 0 ? control_in(a, 0, j, b, M) : T(0);
    for (int k = 0; k < a.T_steps; k++) {
        copy(u, un);
        if ((a.U || a.rng) && k + 1 < a.T_steps) {  // prefetch next control
#pragma unroll
            for (int j = 0; j < M; j++) un[j] = control_in(a, k + 1, j, b, M);
        }
        T xp[N];
        copy(xp, x);
        env_step<P, T, INTEG>(m, x, u, a.dt, h, a.S);
        t += a.dt;  // env.tt.extend([tt[-1] + dt]) (environments.py:269)
        term = P::terminate(m, x);
        // c = (cost(x_, u, x, t, np) + terminal_cost*terminated)*dt (environments.py:274-275)
        cost += (P::cost(m, xp, u, x, t) + (term ? a.terminal_cost : T(0))) * a.dt;
        if (a.X) {
#pragma unroll
            for (int i = 0; i < N; i++) a.X[((int64_t)(k + 1) * N + i) * B + b] = x[i];
        }
    }
    if (a.add_final_cost) cost += P::fcost(m, x) * a.dt;  // ilqr.py:513
    if (a.xN) {
#pragma unroll
        for (int i = 0; i < N; i++) a.xN[i * B + b] = x[i];
    }
    if (a.cost) a.cost[b] = cost;
    if (a.terminated) a.terminated[b] = term ? 1 : 0;
}

// ---- K1b: the same rollout, load-balanced ---------------------------------------------------------
// At the benchmark size (65,536 envs = 2,048 warps on 592 warp schedulers) every scheduler holds 3 or 4
// warps and the kernel ends with the 4-warp ones: 13.5% of the FP64 pipe idles (measured: 65,536 envs take
// exactly as long as 75,776).  Here 3 persistent worker warps sit on EVERY scheduler and pull (group of 32
// trajectories, time chunk) units from a global ticket queue; a group is re-queued after each chunk, so
// every scheduler is fed the same share of the work.
//   ticket s < n_groups      : group s, first chunk (implicit, nothing to wait for)
//   ticket s >= n_groups     : wait until queue[s - n_groups] is published by whichever worker finished a chunk
// Chunk boundaries come from a table (uniform 20-step chunks by default, see queue_schedule() in pg_abi.cu); the
// per-unit hand-off costs 6-7 L2 round trips (~15% of the stall samples at 20 steps per unit).  Tried and measured
// worse at 65,536 x 500: requesting the next ticket and reserving the publish slot when a unit STARTS (hides two
// atomics, but binds workers to specific future units and serves consumers in start order: 1.35 ms instead of 0.97);
// carrying the chunk index in the queue entry and taking the next ticket just before the publish sequence (1.10 ms).
// The state of a group between chunks lives in xN / cost (L2: written with a release fence, read with
// ld.cg, because another SM's L1 may hold a stale copy), the group's clock in tq, its next chunk in progress.
constexpr int MAX_CHUNKS = 255;

template <typename T>
struct QueueArgs {
    RolloutArgs<T> r;
    int n_groups, n_chunks;
    int bounds[MAX_CHUNKS + 1];  // chunk c covers control intervals [bounds[c], bounds[c + 1])
    int* status;     // sticky error word: a worker that gives up waiting sets it; launches never clear it, the host reads it
    int* counters;   // [0] next ticket, [1] next free queue slot
    int* queue;      // [n_groups * (n_chunks - 1)], zero = not yet published, else group + 1
    int* progress;   // [n_groups] next chunk of the group
    T* tq;           // [n_groups] simulation time of the group
};

template <typename P, typename T, int INTEG, int E>
__global__ void __launch_bounds__(128, 3) rollout_queue_kernel(const QueueArgs<T> q) {
    // a group = 32 * E trajectories: lane l holds trajectories g*32*E + e*32 + l, e < E (every access coalesced)
    const typename P::template Ctx<T> m;
    constexpr int N = P::N, M = P::M;
    const RolloutArgs<T>& a = q.r;
    const int64_t B = a.B;
    const int lane = threadIdx.x & 31;
    const T h = a.dt / T(a.S);
    const int total = q.n_groups * q.n_chunks;
    for (;;) {
        int g = -1;
        if (lane == 0) {
            const int s = atomicAdd(&q.counters[0], 1);
            if (s < q.n_groups) {
                g = s;
            } else if (s < total) {
                volatile int* slot = q.queue + (s - q.n_groups);
                int v = 0, spins = 0;
                while ((v = *slot) == 0) {
                    __nanosleep(128);
                    if (++spins > (1 << 24)) {  // ~2 s: something is wrong, do not hang the GPU
                        atomicExch(q.status, 1);
                        break;
                    }
                }
                g = v - 1;
                __threadfence();  // acquire: the producer's state stores are visible after this
            }
        }
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g < 0) return;
        const int c = __ldcg(&q.progress[g]);
        int64_t b[E], bb[E];
        bool valid[E];
#pragma unroll
        for (int e = 0; e < E; e++) {
            b[e] = ((int64_t)g * E + e) * 32 + lane;
            valid[e] = b[e] < B;
            bb[e] = valid[e] ? b[e] : B - 1;  // ragged last group: idle slots shadow the last trajectory
        }
        const int k0 = q.bounds[c], k1 = q.bounds[c + 1];
        T x[E][N], u[E][M], un[E][M], cost[E];
        T t;
#pragma unroll
        for (int e = 0; e < E; e++) {
            if (c == 0) {
#pragma unroll
                for (int i = 0; i < N; i++) x[e][i] = a.x0[i * B + bb[e]];
                cost[e] = (a.accumulate_cost) ? __ldcg(&a.cost[bb[e]]) : T(0);
                if (a.X && valid[e]) {
#pragma unroll
                    for (int i = 0; i < N; i++) a.X[i * B + b[e]] = x[e][i];
                }
            } else {
#pragma unroll
                for (int i = 0; i < N; i++) x[e][i] = __ldcg(&a.xN[i * B + bb[e]]);
                cost[e] = __ldcg(&a.cost[bb[e]]);
            }
#pragma unroll
            for (int j = 0; j < M; j++) un[e][j] = k0 < k1 ? control_in(a, k0, j, bb[e], M) : T(0);
        }
        t = (c == 0) ? a.t0 : __ldcg(&q.tq[g]);
        bool term[E];
#pragma unroll
        for (int e = 0; e < E; e++) term[e] = false;
        for (int k = k0; k < k1; k++) {
            T xp[E][N];
#pragma unroll
            for (int e = 0; e < E; e++) {
                copy(u[e], un[e]);
                if ((a.U || a.rng) && k + 1 < k1) {
#pragma unroll
                    for (int j = 0; j < M; j++) un[e][j] = control_in(a, k + 1, j, bb[e], M);
                }
                copy(xp[e], x[e]);
            }
            env_step_multi<P, T, INTEG, E>(m, x, u, a.dt, h, a.S);
            t += a.dt;
#pragma unroll
            for (int e = 0; e < E; e++) {
                term[e] = P::terminate(m, x[e]);
                cost[e] += (P::cost(m, xp[e], u[e], x[e], t) + (term[e] ? a.terminal_cost : T(0))) * a.dt;
                if (a.X && valid[e]) {
#pragma unroll
                    for (int i = 0; i < N; i++) a.X[((int64_t)(k + 1) * N + i) * B + b[e]] = x[e][i];
                }
            }
        }
        const bool last = k1 >= a.T_steps;
#pragma unroll
        for (int e = 0; e < E; e++) {
            if (last && a.add_final_cost) cost[e] += P::fcost(m, x[e]) * a.dt;
            if (valid[e]) {
#pragma unroll
                for (int i = 0; i < N; i++) a.xN[i * B + b[e]] = x[e][i];
                a.cost[b[e]] = cost[e];
                if (last && a.terminated) a.terminated[b[e]] = term[e] ? 1 : 0;
            }
        }
        if (!last) {
            __threadfence();  // release: state before the ticket
            __syncwarp();
            if (lane == 0) {
                q.tq[g] = t;
                q.progress[g] = c + 1;
                __threadfence();
                const int p = atomicAdd(&q.counters[1], 1);
                atomicExch(&q.queue[p], g + 1);
            }
        }
    }
}

// U [T, m, B] <- the action stream (validation against the host stream; feeding kernels that read U)
template <typename T>
__global__ void __launch_bounds__(256) random_actions_kernel(int64_t B, int T_steps, int M, uint32_t seed_lo, uint32_t seed_hi, int k_offset,
                                                             double um0, double um1, T* __restrict__ U) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)T_steps * M * B) return;
    const int64_t b = i % B;
    const int j = (int)((i / B) % M), k = (int)(i / (B * M));
    U[i] = philox_action<T>(seed_lo, seed_hi, b, k_offset + k, j, j == 0 ? um0 : um1);
}

// ---- K2: closed-loop line-search rollouts (iLQR.forward_pass, ilqr.py:187-231) ---------------------
constexpr int MAX_ALPHAS = 16;

template <typename T>
struct ForwardArgs {
    int64_t B;
    int T_steps, S, Tk, n_alpha;
    T dt, terminal_cost;
    T alphas[MAX_ALPHAS];
    const T* alpha_dev;
    const T* x0;
    const T* xbar;
    const T* ubar;
    const T* KK;
    const T* kk;
    int clip;
    T ulim[2];
    T* X_out;
    T* U_out;
    T* cost_out;
    uint8_t* diverged_out;
    uint8_t* terminated_out;
    T* xN_out;  // [n_alpha, n, B] or NULL: terminal state of every candidate (what the planner's environment is left in)
    T* xS;      // [n_alpha, n_snap, n, B] or NULL: state of every candidate at the chunk boundaries snap[1..n_snap]
    int n_snap;
    int snap[MAX_CHUNKS + 1];  // snap[c] = first control interval of chunk c (queue_schedule() of pg_abi.cu)
    const uint8_t* active;
    const int32_t* index;  // optional work list: thread i handles trajectory index[i], i < *count
    const int32_t* count;
    // alpha window of a line search split over two launches (pg_alpha_window): decided on the device from *win_count
    int win_phase, win_first, win_small;
    const int32_t* win_count;
    // commit mode
    const uint8_t* accept;
    T* xbar_io;
    T* ubar_io;
    T* KK_acc;
    T* kk_acc;
};

// Alphas [lo, hi) this launch evaluates.  Launch 1 of a split search takes the first win_first alphas, launch 2 the
// others -- unless at most win_small trajectories are live: then launch 1 takes them all and launch 2 none (a second
// dependent chain of T intervals would cost more than it saves when the GPU is not full).
template <typename T>
__device__ __forceinline__ void alpha_window(const ForwardArgs<T>& a, int& lo, int& hi) {
    lo = 0;
    hi = a.n_alpha;
    if (a.win_phase) {
        const int split = (*a.win_count <= a.win_small) ? a.n_alpha : a.win_first;
        if (a.win_phase == 1) hi = split;
        else lo = split;
    }
}

template <typename P, typename T>
struct Nominal {  // gains and nominal point of one control interval
    T K[P::M * P::N], k[P::M], xb[P::N], ub[P::M];
    __device__ __forceinline__ void load(const T* KK, const T* kk, const T* xbar, const T* ubar, int j, int64_t B, int64_t b) {
        constexpr int N = P::N, M = P::M;
#pragma unroll
        for (int r = 0; r < M * N; r++) K[r] = KK[((int64_t)j * M * N + r) * B + b];
#pragma unroll
        for (int r = 0; r < M; r++) k[r] = kk[((int64_t)j * M + r) * B + b];
#pragma unroll
        for (int i = 0; i < N; i++) xb[i] = xbar[((int64_t)j * N + i) * B + b];
#pragma unroll
        for (int r = 0; r < M; r++) ub[r] = ubar[((int64_t)j * M + r) * B + b];
    }
};

// COMMIT = false: evaluate candidates (cost, flags, optional trajectories).
// COMMIT = true : re-run the accepted alpha and overwrite the nominal trajectory in place; the
//                 nominal point of interval k+1 is always loaded BEFORE x_{k+1} is stored, so the
//                 in-place update never reads its own output (the pointers are deliberately not
//                 __restrict__).
// Alphas per CTA: all of them while the state fits 128 registers; systems with n >= 8 (triple pendulum) get 8 warps per
// CTA and the full register file per thread instead (the alphas then span gridDim.y CTAs).
template <typename P>
constexpr int forward_alphas_per_cta() { return P::N >= 8 ? 8 : MAX_ALPHAS; }

// The closed-loop rollout of ONE (trajectory b, candidate ia) pair -- the body of K2, and what lane ia of a warp runs in the
// fused tail solver.  Trajectory outputs go to Xo / Uo, whose element (k, i) lives at Xo[(k * N + i) * os + oo]: the nominal
// arrays and the candidate arrays of K2 have os = B, oo = b; the fused solver keeps a warp's candidates lane-minor.
template <typename P, typename T, int INTEG, bool COMMIT>
__device__ __forceinline__ void forward_body(const typename P::template Ctx<T>& m, const ForwardArgs<T>& a, const int64_t b, const int ia,
                                             T* Xo, T* Uo, const int64_t os, const int64_t oo) {
    constexpr int N = P::N, M = P::M;
    const int64_t B = a.B;
    const T alpha = a.alpha_dev ? a.alpha_dev[b] : a.alphas[ia];
    const T* xbar = COMMIT ? a.xbar_io : a.xbar;
    const T* ubar = COMMIT ? a.ubar_io : a.ubar;
    T x[N], u[M];
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = a.x0[i * B + b];
    bool div = false;
#pragma unroll
    for (int i = 0; i < N; i++) div |= (x[i] > T(1e8));
    const T h = a.dt / T(a.S);
    T cost = T(0), t = T(0);  // environment.reset(x0) => tt = [0] (ilqr.py:199)
    bool term = false;
    Nominal<P, T> cur, nxt;
    if (a.T_steps > 0) nxt.load(a.KK, a.kk, xbar, ubar, 0, B, b);
    if (Xo) {
#pragma unroll
        for (int i = 0; i < N; i++) Xo[i * os + oo] = x[i];
    }
    int sc = 0;  // next snapshot
    for (int k = 0; k < a.T_steps; k++) {
        cur = nxt;
        if (k + 1 < a.T_steps) {
            const int j = (k + 1 < a.Tk) ? (k + 1) : (a.Tk - 1);
            nxt.load(a.KK, a.kk, xbar, ubar, j, B, b);
        }
        // u = KK[j] @ (x - xx_[j]) + alpha*kk[j] + uu_[j]   (ilqr.py:206)
#pragma unroll
        for (int r = 0; r < M; r++) {
            T acc = T(0);
#pragma unroll
            for (int i = 0; i < N; i++) acc += cur.K[r * N + i] * (x[i] - cur.xb[i]);
            acc = acc + alpha * cur.k[r];
            acc = acc + cur.ub[r];
            if (a.clip) acc = pg::clamp(acc, -a.ulim[r], a.ulim[r]);  // ilqr.py:208-209
            u[r] = acc;
        }
        T xp[N];
        copy(xp, x);
        env_step<P, T, INTEG>(m, x, u, a.dt, h, a.S);
        t += a.dt;
        if constexpr (!COMMIT) {
            term = P::terminate(m, x);
            cost += (P::cost(m, xp, u, x, t) + (term ? a.terminal_cost : T(0))) * a.dt;
#pragma unroll
            for (int i = 0; i < N; i++) div |= (x[i] > T(1e8));  // np.any(xx > 1e8) (ilqr.py:435)
        }
        if (Xo) {
#pragma unroll
            for (int i = 0; i < N; i++) Xo[((int64_t)(k + 1) * N + i) * os + oo] = x[i];
        }
        if constexpr (!COMMIT) {
            if (a.xS && sc < a.n_snap && k + 1 == a.snap[sc + 1]) {  // end of chunk sc
#pragma unroll
                for (int i = 0; i < N; i++) a.xS[(((int64_t)ia * a.n_snap + sc) * N + i) * B + b] = x[i];
                sc++;
            }
        }
        if (Uo) {
#pragma unroll
            for (int r = 0; r < M; r++) Uo[((int64_t)k * M + r) * os + oo] = u[r];
        }
        if constexpr (COMMIT) {
            if (k < a.Tk) {  // self.KK = KK; self.kk = kk (ilqr.py:463-464)
#pragma unroll
                for (int r = 0; r < M * N; r++) a.KK_acc[((int64_t)k * M * N + r) * B + b] = cur.K[r];
#pragma unroll
                for (int r = 0; r < M; r++) a.kk_acc[((int64_t)k * M + r) * B + b] = cur.k[r];
            }
        }
    }
    if constexpr (!COMMIT) {
        cost += P::fcost(m, x) * a.dt;  // ilqr.py:224
        const int64_t o = (int64_t)ia * B + b;
        a.cost_out[o] = cost;
        if (a.diverged_out) a.diverged_out[o] = div ? 1 : 0;
        if (a.terminated_out) a.terminated_out[o] = term ? 1 : 0;
        if (a.xN_out) {
#pragma unroll
            for (int i = 0; i < N; i++) a.xN_out[((int64_t)ia * N + i) * B + b] = x[i];
        }
    }
}

template <typename P, typename T, int INTEG, bool COMMIT>
__global__ void __launch_bounds__(32 * forward_alphas_per_cta<P>()) ilqr_forward_kernel(const ForwardArgs<T> a) {
    const typename P::template Ctx<T> m;
    constexpr int N = P::N, M = P::M;
    const int64_t B = a.B;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int ia = blockIdx.y * blockDim.y + threadIdx.y;
    int64_t b = tid;
    if (a.index) {  // compacted list of running trajectories (written by ilqr_select_kernel)
        if (tid >= *a.count) return;
        b = a.index[tid];
    } else if (b >= B) {
        return;
    }
    if constexpr (COMMIT) {
        if (!a.accept[b]) return;
    } else {
        if (a.active && !a.active[b]) return;
        int lo, hi;
        alpha_window(a, lo, hi);
        if (ia < lo || ia >= hi) return;
    }
    T* Xo = nullptr;
    T* Uo = nullptr;
    if constexpr (COMMIT) {
        Xo = a.xbar_io;
        Uo = a.ubar_io;
    } else {
        if (a.X_out) Xo = a.X_out + (int64_t)ia * (a.T_steps + 1) * N * B;
        if (a.U_out) Uo = a.U_out + (int64_t)ia * a.T_steps * M * B;
    }
    forward_body<P, T, INTEG, COMMIT>(m, a, b, ia, Xo, Uo, B, b);
}

// ---- K2b: the candidate rollouts, load-balanced -----------------------------------------------------------------
// 16,384 trajectories x 11 alphas are 5,632 warps: with ilqr_forward_kernel's 126 registers that is ONE 352-thread CTA
// per SM (2.75 warps per scheduler) in 3.46 -> 4 waves.  Here the K1b scheme: 3 persistent worker warps per scheduler
// pull (group of 32 work-list entries, alpha, time chunk) units from the ticket queue; between chunks a candidate lives
// in its output slots (xN_out, cost_out, diverged_out) and its clock in tq.  Evaluates costs only (no X_out / U_out);
// results are bit-identical to ilqr_forward_kernel.  The group count depends on the device-side work-list length, so
// every worker derives it from *count.
template <typename T>
struct ForwardQueueArgs {
    ForwardArgs<T> f;
    int n_chunks;
    int bounds[MAX_CHUNKS + 1];
    int* status;    // sticky error word (see QueueArgs)
    int* counters;  // [0] next ticket, [1] next free queue slot
    int* queue;
    int* progress;  // [n_alpha * ceil(B / 32)]
    T* tq;
};

template <typename P, typename T, int INTEG>
__global__ void __launch_bounds__(128, 3) ilqr_forward_queue_kernel(const ForwardQueueArgs<T> q) {
    const typename P::template Ctx<T> m;
    constexpr int N = P::N, M = P::M;
    const ForwardArgs<T>& a = q.f;
    const int64_t B = a.B;
    const int lane = threadIdx.x & 31;
    const int64_t cnt = a.index ? (int64_t)*a.count : B;
    const int ng = (int)((cnt + 31) / 32);
    int a_lo, a_hi;
    alpha_window(a, a_lo, a_hi);
    const int G = ng * (a_hi - a_lo), total = G * q.n_chunks;
    const T h = a.dt / T(a.S);
    for (;;) {
        int g = -1;
        if (lane == 0) {
            const int s = atomicAdd(&q.counters[0], 1);
            if (s < G) {
                g = s;
            } else if (s < total) {
                volatile int* slot = q.queue + (s - G);
                int v = 0, spins = 0;
                while ((v = *slot) == 0) {
                    __nanosleep(128);
                    if (++spins > (1 << 24)) {
                        atomicExch(q.status, 1);
                        break;
                    }
                }
                g = v - 1;
                __threadfence();
            }
        }
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g < 0) return;
        const int c = __ldcg(&q.progress[g]);
        const int tg = g % ng, ia = a_lo + g / ng;
        const int64_t w = (int64_t)tg * 32 + lane;
        const bool valid = w < cnt;
        const int64_t wv = valid ? w : cnt - 1;  // ragged last group: idle lanes shadow the last entry
        const int64_t b = a.index ? (int64_t)a.index[wv] : wv;
        const int64_t o = (int64_t)ia * B + b;
        const int k0 = q.bounds[c], k1 = q.bounds[c + 1];
        const T alpha = a.alpha_dev ? a.alpha_dev[b] : a.alphas[ia];
        T x[N], u[M];
        T cost, t;
        bool div = false, term = false;
        if (c == 0) {
#pragma unroll
            for (int i = 0; i < N; i++) x[i] = a.x0[i * B + b];
#pragma unroll
            for (int i = 0; i < N; i++) div |= (x[i] > T(1e8));
            cost = T(0);
            t = T(0);
        } else {
#pragma unroll
            for (int i = 0; i < N; i++) x[i] = __ldcg(&a.xN_out[((int64_t)ia * N + i) * B + b]);
            cost = __ldcg(&a.cost_out[o]);
            div = __ldcg(&a.diverged_out[o]) != 0;
            t = __ldcg(&q.tq[g]);
        }
        Nominal<P, T> cur, nxt;
        if (k0 < k1) nxt.load(a.KK, a.kk, a.xbar, a.ubar, (k0 < a.Tk) ? k0 : (a.Tk - 1), B, b);
        for (int k = k0; k < k1; k++) {
            cur = nxt;
            if (k + 1 < k1) {
                const int j = (k + 1 < a.Tk) ? (k + 1) : (a.Tk - 1);
                nxt.load(a.KK, a.kk, a.xbar, a.ubar, j, B, b);
            }
#pragma unroll
            for (int r = 0; r < M; r++) {  // u = KK[j] @ (x - xx_[j]) + alpha*kk[j] + uu_[j]   (ilqr.py:206)
                T acc = T(0);
#pragma unroll
                for (int i = 0; i < N; i++) acc += cur.K[r * N + i] * (x[i] - cur.xb[i]);
                acc = acc + alpha * cur.k[r];
                acc = acc + cur.ub[r];
                if (a.clip) acc = pg::clamp(acc, -a.ulim[r], a.ulim[r]);
                u[r] = acc;
            }
            T xp[N];
            copy(xp, x);
            env_step<P, T, INTEG>(m, x, u, a.dt, h, a.S);
            t += a.dt;
            term = P::terminate(m, x);
            cost += (P::cost(m, xp, u, x, t) + (term ? a.terminal_cost : T(0))) * a.dt;
#pragma unroll
            for (int i = 0; i < N; i++) div |= (x[i] > T(1e8));
        }
        const bool last = k1 >= a.T_steps;
        if (last) cost += P::fcost(m, x) * a.dt;
        if (valid) {
#pragma unroll
            for (int i = 0; i < N; i++) a.xN_out[((int64_t)ia * N + i) * B + b] = x[i];
            a.cost_out[o] = cost;
            a.diverged_out[o] = div ? 1 : 0;
            if (last && a.terminated_out) a.terminated_out[o] = term ? 1 : 0;
            if (a.xS && !last && c < a.n_snap) {
#pragma unroll
                for (int i = 0; i < N; i++) a.xS[(((int64_t)ia * a.n_snap + c) * N + i) * B + b] = x[i];
            }
        }
        if (!last) {
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                q.tq[g] = t;
                q.progress[g] = c + 1;
                __threadfence();
                const int p = atomicAdd(&q.counters[1], 1);
                atomicExch(&q.queue[p], g + 1);
            }
        }
    }
}

// ---- K2'': chunk-parallel commit -------------------------------------------------------------------------------------
// Re-integrating the accepted candidate as ONE rollout (ilqr_forward_kernel<..., COMMIT>) is a dependent chain of T
// control intervals run by B / 32 warps: 95 us for the cart-pole, 240-360 us for the n = 6 systems.  With the candidates'
// states at the chunk boundaries (xS, written by the line search) every chunk of the winner is integrated by its own
// thread: n_chunks times more warps, chains n_chunks times shorter.  In place: thread (b, c) reads the nominal points
// k0 .. k1-1 of its chunk and writes x_{k0} .. x_{k1-1} (its own start state included, AFTER loading the nominal there)
// but not x_{k1}, which thread (b, c+1) owns -- so no thread writes what another one reads; the last chunk also writes
// x_T.  Bit-identical to the single-rollout commit: the snapshots are the states that rollout passes through.
template <typename T>
struct CommitChunkArgs {
    ForwardArgs<T> f;   // B, T_steps, S, Tk, n_alpha, dt, alphas, x0, KK, kk, clip, ulim, xS, n_snap, snap, index, count
    const T* alpha_best;
    const uint8_t* accept;
    T* xbar_io;         // [T+1, n, B]
    T* ubar_io;         // [T, m, B]
    T* KK_acc;
    T* kk_acc;
};

template <typename P, typename T, int INTEG>
__global__ void __launch_bounds__(128) ilqr_commit_chunks_kernel(const CommitChunkArgs<T> q) {
    const typename P::template Ctx<T> m;
    constexpr int N = P::N, M = P::M;
    const ForwardArgs<T>& a = q.f;
    const int64_t B = a.B;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    int64_t b = tid;
    if (a.index) {
        if (tid >= *a.count) return;
        b = a.index[tid];
    } else if (b >= B) {
        return;
    }
    if (!q.accept[b]) return;
    const T alpha = q.alpha_best[b];
    int ia = 0;
    for (int i = 0; i < a.n_alpha; i++)
        if (a.alphas[i] == alpha) ia = i;
    const int k0 = a.snap[c], k1 = a.snap[c + 1];
    T x[N], u[M];
    if (c == 0) {
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = a.x0[i * B + b];
    } else {
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = a.xS[(((int64_t)ia * a.n_snap + (c - 1)) * N + i) * B + b];
    }
    const T h = a.dt / T(a.S);
    Nominal<P, T> cur, nxt;
    if (k0 < k1) nxt.load(a.KK, a.kk, q.xbar_io, q.ubar_io, (k0 < a.Tk) ? k0 : (a.Tk - 1), B, b);
#pragma unroll
    for (int i = 0; i < N; i++) q.xbar_io[((int64_t)k0 * N + i) * B + b] = x[i];  // own start state, nominal k0 is in registers
    for (int k = k0; k < k1; k++) {
        cur = nxt;
        if (k + 1 < k1) {
            const int j = (k + 1 < a.Tk) ? (k + 1) : (a.Tk - 1);
            nxt.load(a.KK, a.kk, q.xbar_io, q.ubar_io, j, B, b);
        }
#pragma unroll
        for (int r = 0; r < M; r++) {
            T acc = T(0);
#pragma unroll
            for (int i = 0; i < N; i++) acc += cur.K[r * N + i] * (x[i] - cur.xb[i]);
            acc = acc + alpha * cur.k[r];
            acc = acc + cur.ub[r];
            if (a.clip) acc = pg::clamp(acc, -a.ulim[r], a.ulim[r]);
            u[r] = acc;
        }
        env_step<P, T, INTEG>(m, x, u, a.dt, h, a.S);
        if (k + 1 < k1 || k1 >= a.T_steps) {  // x_{k1} belongs to the next chunk's thread, except at the very end
#pragma unroll
            for (int i = 0; i < N; i++) q.xbar_io[((int64_t)(k + 1) * N + i) * B + b] = x[i];
        }
#pragma unroll
        for (int r = 0; r < M; r++) q.ubar_io[((int64_t)k * M + r) * B + b] = u[r];
        if (k < a.Tk) {  // self.KK = KK; self.kk = kk (ilqr.py:463-464)
#pragma unroll
            for (int r = 0; r < M * N; r++) q.KK_acc[((int64_t)k * M * N + r) * B + b] = cur.K[r];
#pragma unroll
            for (int r = 0; r < M; r++) q.kk_acc[((int64_t)k * M + r) * B + b] = cur.k[r];
        }
    }
}

// ---- linearisation ------------------------------------------------------------------------------
// lin_mode 1: helpers.system_linearization (pygent/helpers.py:130-148), central differences, eps=1e-3.
template <typename P, typename T>
__device__ __forceinline__ void fd_AB(const typename P::template Ctx<T>& m, const T (&x)[P::N], const T (&u)[P::M], T (&A)[P::N * P::N], T (&Bm)[P::N * P::M]) {
    constexpr int N = P::N, M = P::M;
    const T eps = T(1e-3);
    T fp[N], fm[N], xt[N], ut[M];
#pragma unroll
    for (int j = 0; j < N; j++) {
        copy(xt, x);
        xt[j] = x[j] + eps;
        P::f(m, xt, u, fp);
        xt[j] = x[j] - eps;
        P::f(m, xt, u, fm);
#pragma unroll
        for (int i = 0; i < N; i++) A[i * N + j] = (fp[i] - fm[i]) / (T(2) * eps);
    }
#pragma unroll
    for (int j = 0; j < M; j++) {
        copy(ut, u);
        ut[j] = u[j] + eps;
        P::f(m, x, ut, fp);
        ut[j] = u[j] - eps;
        P::f(m, x, ut, fm);
#pragma unroll
        for (int i = 0; i < N; i++) Bm[i * M + j] = (fp[i] - fm[i]) / (T(2) * eps);
    }
}

// products that must not be contracted into the following subtraction (a*b - c*d as fma(a, b, -(c*d)) rounds one product
// and not the other: a one-sided error that the 1/(4 eps^2) of a second difference turns into 1e-10)
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }

// finite_diff=True: the cost expansion by central differences, eps = 1e-3 (pygent/helpers.py:41-128 as called from
// iLQR.cost_lin, ilqr.py:598-603, on cost_fnc = c(x, u, t) * dt, and from backward_pass, ilqr.py:240-242, on the raw final
// cost).  Every entry of the Hessians is formed on its own like the reference does -- H[i][j] and H[j][i] see the two middle
// evaluations in opposite order -- and the diagonal steps twice, (x + eps) + eps.  Returns the derivatives of c*dt.
template <typename P, typename T>
__device__ __forceinline__ void fd_cost_derivs(const typename P::template Ctx<T>& m, const T (&x)[P::N], const T (&u)[P::M], const T t, const T dt,
                                               T (&cx)[P::N], T (&cu)[P::M], T (&Cxx)[P::N * P::N], T (&Cuu)[P::M * P::M], T (&Cxu)[P::N * P::M]) {
    constexpr int N = P::N, M = P::M;
    const T eps = T(1e-3);
    const T d2 = T(2) * eps, d4 = T(4) * (eps * eps);
    T xa[N], xb[N], ua[M], ub[M];
    auto c = [&](const T (&xx)[N], const T (&uu)[M]) { return mul_rn(P::cost(m, xx, uu, xx, t), dt); };
#pragma unroll
    for (int i = 0; i < N; i++) {
        copy(xa, x);
        copy(xb, x);
        xa[i] = x[i] + eps;
        xb[i] = x[i] - eps;
        cx[i] = (c(xa, u) - c(xb, u)) / d2;
#pragma unroll
        for (int j = 0; j < N; j++) {
            T xpp[N], xmp[N], xpm[N], xmm[N];
            copy(xpp, xa);
            copy(xmp, xb);
            copy(xpm, xa);
            copy(xmm, xb);
            xpp[j] = xa[j] + eps;
            xmp[j] = xb[j] + eps;
            xpm[j] = xa[j] - eps;
            xmm[j] = xb[j] - eps;
            Cxx[i * N + j] = (((c(xpp, u) - c(xmp, u)) - c(xpm, u)) + c(xmm, u)) / d4;
        }
#pragma unroll
        for (int r = 0; r < M; r++) {
            copy(ua, u);
            copy(ub, u);
            ua[r] = u[r] + eps;
            ub[r] = u[r] - eps;
            Cxu[i * M + r] = (((c(xa, ua) - c(xb, ua)) - c(xa, ub)) + c(xb, ub)) / d4;
        }
    }
#pragma unroll
    for (int r = 0; r < M; r++) {
        copy(ua, u);
        copy(ub, u);
        ua[r] = u[r] + eps;
        ub[r] = u[r] - eps;
        cu[r] = (c(x, ua) - c(x, ub)) / d2;
#pragma unroll
        for (int s = 0; s < M; s++) {
            T upp[M], ump[M], upm[M], umm[M];
            copy(upp, ua);
            copy(ump, ub);
            copy(upm, ua);
            copy(umm, ub);
            upp[s] = ua[s] + eps;
            ump[s] = ub[s] + eps;
            upm[s] = ua[s] - eps;
            umm[s] = ub[s] - eps;
            Cuu[r * M + s] = (((c(x, upp) - c(x, ump)) - c(x, upm)) + c(x, umm)) / d4;
        }
    }
}

// helpers.fxN / fxxN on the final cost (ilqr.py:240-242); the caller scales by dt afterwards, like the reference.
template <typename P, typename T>
__device__ __forceinline__ void fd_fcost_derivs(const typename P::template Ctx<T>& m, const T (&x)[P::N], T (&cfx)[P::N], T (&Cfxx)[P::N * P::N]) {
    constexpr int N = P::N;
    const T eps = T(1e-3);
    const T d2 = T(2) * eps, d4 = T(4) * (eps * eps);
    T xa[N], xb[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        copy(xa, x);
        copy(xb, x);
        xa[i] = x[i] + eps;
        xb[i] = x[i] - eps;
        cfx[i] = (P::fcost(m, xa) - P::fcost(m, xb)) / d2;
#pragma unroll
        for (int j = 0; j < N; j++) {
            T xpp[N], xmp[N], xpm[N], xmm[N];
            copy(xpp, xa);
            copy(xmp, xb);
            copy(xpm, xa);
            copy(xmm, xb);
            xpp[j] = xa[j] + eps;
            xmp[j] = xb[j] + eps;
            xpm[j] = xa[j] - eps;
            xmm[j] = xb[j] - eps;
            Cfxx[i * N + j] = (((P::fcost(m, xpp) - P::fcost(m, xmp)) - P::fcost(m, xpm)) + P::fcost(m, xmm)) / d4;
        }
    }
}

// ---- K3: backward Riccati pass (iLQR.backward_pass, ilqr.py:233-346) --------------------------------
template <typename T>
struct BackwardArgs {
    int64_t B;
    int T_steps, reg_type, constrained;
    T dt;
    T ulim[2];
    const T* mu;
    const T* xbar;
    const T* ubar;
    T* KK;
    T* kk;
    T* dV;
    T* gnorm;
    uint8_t* success;
    const uint8_t* active;
    const int32_t* index;
    const int32_t* count;
    const T* A_ext;  // LIN == 2: continuous-time Jacobians supplied by the caller, [T, n, n, B] and [T, n, m, B]
    const T* B_ext;
    const T* Vxx_T;  // [n, n, B] or NULL: terminal value Hessian of the final backward pass (DARE solution, ilqr.py:246-260;
                     // the terminal gradient is zero there) instead of the final-cost expansion
};

template <typename T>
__device__ __forceinline__ bool np_isclose(T a, T b, T atol) {  // np.isclose(a, b, atol=atol), rtol=1e-5
    return fabs(a - b) <= atol + T(1e-5) * fabs(b);
}

// The backward sweep of ONE trajectory b (the body of K3; also lane 0's part of the fused tail solver ilqr_solve_kernel).
template <typename P, typename T, int LIN, bool CFD = false>
__device__ __forceinline__ void backward_body(const typename P::template Ctx<T>& m, const BackwardArgs<T>& a, const int64_t b) {
    constexpr int N = P::N, M = P::M;
    static_assert(M <= 2, "control dimension > 2 not supported by the closed-form Quu solve");
    const int64_t B = a.B;
    const T mu = a.mu[b];
    const T dt = a.dt;
    const int Ts = a.T_steps;
    T x[N], u[M];
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = a.xbar[((int64_t)Ts * N + i) * B + b];
    // vx = cfx(x_N)*dt, Vxx = Cfxx(x_N)*dt (ilqr.py:244-245)
    T vx[N], Vxx[N * N];
    if (a.Vxx_T) {  // vx = P x * 0, Vxx = P (ilqr.py:258-259)
#pragma unroll
        for (int i = 0; i < N; i++) vx[i] = T(0);
#pragma unroll
        for (int i = 0; i < N * N; i++) Vxx[i] = a.Vxx_T[(int64_t)i * B + b];
    } else {
        if constexpr (CFD) fd_fcost_derivs<P, T>(m, x, vx, Vxx);
        else P::fcost_derivs(m, x, vx, Vxx);
#pragma unroll
        for (int i = 0; i < N; i++) vx[i] *= dt;
#pragma unroll
        for (int i = 0; i < N * N; i++) Vxx[i] = (CFD || P::cfxx_nz(i)) ? Vxx[i] * dt : T(0);
    }
    T sV1 = T(0), sV2 = T(0);
    T gmax[M];
#pragma unroll
    for (int r = 0; r < M; r++) gmax[r] = T(0);
    bool ok = true;

    T xn_[N], un_[M];
    if (Ts > 0) {
#pragma unroll
        for (int i = 0; i < N; i++) xn_[i] = a.xbar[((int64_t)(Ts - 1) * N + i) * B + b];
#pragma unroll
        for (int r = 0; r < M; r++) un_[r] = a.ubar[((int64_t)(Ts - 1) * M + r) * B + b];
    }
    for (int k = Ts - 1; k >= 0; k--) {
        copy(x, xn_);
        copy(u, un_);
        if (k > 0) {  // prefetch the next (earlier) nominal point
#pragma unroll
            for (int i = 0; i < N; i++) xn_[i] = a.xbar[((int64_t)(k - 1) * N + i) * B + b];
#pragma unroll
            for (int r = 0; r < M; r++) un_[r] = a.ubar[((int64_t)(k - 1) * M + r) * B + b];
        }
        // Ad = A*dt + I, Bd = B*dt (ilqr.py:530-531) -- forward-Euler discretisation of the Jacobian
        T Ad[N * N], Bd[N * M];
        if constexpr (LIN == 0) {
            P::AB(m, x, u, Ad, Bd);
        } else if constexpr (LIN == 1) {
            fd_AB<P, T>(m, x, u, Ad, Bd);
        } else {  // learned dynamics: linearised by its own library (pg_mlp_linearize)
#pragma unroll
            for (int i = 0; i < N * N; i++) Ad[i] = a.A_ext[((int64_t)k * N * N + i) * B + b];
#pragma unroll
            for (int i = 0; i < N * M; i++) Bd[i] = a.B_ext[((int64_t)k * N * M + i) * B + b];
        }
#pragma unroll
        for (int i = 0; i < N; i++)
#pragma unroll
            for (int j = 0; j < N; j++) Ad[i * N + j] = Ad[i * N + j] * dt + (i == j ? T(1) : T(0));
#pragma unroll
        for (int i = 0; i < N * M; i++) Bd[i] = Bd[i] * dt;
        // structural sparsity known at compile time (analytic Jacobians only)
        auto adnz = [](int i, int j) constexpr { return LIN != 0 || i == j || P::a_nz(i * N + j); };
        auto bdnz = [](int i, int r) constexpr { return LIN != 0 || P::b_nz(i * M + r); };

        // cost expansion of c*dt at t_k = k*dt (ilqr.py:162, :236, :597-610)
        T cx[N], cu[M], Cxx[N * N], Cuu[M * M], Cxu[N * M];
        // (finite differences are taken of c*dt itself: cdt = 1 there, and no entry is structurally zero)
        const T cdt = CFD ? T(1) : dt;
        if constexpr (CFD) fd_cost_derivs<P, T>(m, x, u, T(k) * dt, dt, cx, cu, Cxx, Cuu, Cxu);
        else P::cost_derivs(m, x, u, T(k) * dt, cx, cu, Cxx, Cuu, Cxu);
#pragma unroll
        for (int i = 0; i < N; i++) cx[i] *= cdt;
#pragma unroll
        for (int r = 0; r < M; r++) cu[r] *= cdt;
#pragma unroll
        for (int i = 0; i < M * M; i++) Cuu[i] *= cdt;

        // W = Vxx Ad, g = Vxx Bd
        T W[N * N], g[N * M];
#pragma unroll
        for (int i = 0; i < N; i++) {
#pragma unroll
            for (int j = 0; j < N; j++) {
                T acc = T(0);
#pragma unroll
                for (int l = 0; l < N; l++)
                    if (adnz(l, j)) acc += Vxx[i * N + l] * Ad[l * N + j];
                W[i * N + j] = acc;
            }
#pragma unroll
            for (int r = 0; r < M; r++) {
                T acc = T(0);
#pragma unroll
                for (int l = 0; l < N; l++)
                    if (bdnz(l, r)) acc += Vxx[i * N + l] * Bd[l * M + r];
                g[i * M + r] = acc;
            }
        }
        // qx = cx + Ad' vx ; qu = cu + Bd' vx (ilqr.py:283-284)
        T qx[N], qu[M];
#pragma unroll
        for (int i = 0; i < N; i++) {
            T acc = T(0);
#pragma unroll
            for (int l = 0; l < N; l++)
                if (adnz(l, i)) acc += Ad[l * N + i] * vx[l];
            qx[i] = cx[i] + acc;
        }
#pragma unroll
        for (int r = 0; r < M; r++) {
            T acc = T(0);
#pragma unroll
            for (int l = 0; l < N; l++)
                if (bdnz(l, r)) acc += Bd[l * M + r] * vx[l];
            qu[r] = cu[r] + acc;
        }
        // Quu = Cuu + Bd' g ; Qux = Cxu' + Bd' W (ilqr.py:288-289)
        T Quu[M * M], Qux[M * N], QuuR[M * M], QuxR[M * N];
#pragma unroll
        for (int r = 0; r < M; r++) {
#pragma unroll
            for (int s = 0; s < M; s++) {
                T acc = T(0);
#pragma unroll
                for (int l = 0; l < N; l++)
                    if (bdnz(l, r)) acc += Bd[l * M + r] * g[l * M + s];
                Quu[r * M + s] = Cuu[r * M + s] + acc;
            }
#pragma unroll
            for (int j = 0; j < N; j++) {
                T acc = T(0);
#pragma unroll
                for (int l = 0; l < N; l++)
                    if (bdnz(l, r)) acc += Bd[l * M + r] * W[l * N + j];
                Qux[r * N + j] = ((CFD || P::cxu_nz(j * M + r)) ? Cxu[j * M + r] * cdt : T(0)) + acc;
            }
        }
        // regularisation (ilqr.py:291-295)
        if (a.reg_type == 1) {
            // VxxReg = Vxx + mu I  =>  Bd' VxxReg Bd = Bd'(g + mu Bd), Bd' VxxReg Ad = Bd'(W + mu Ad)
#pragma unroll
            for (int r = 0; r < M; r++) {
#pragma unroll
                for (int s = 0; s < M; s++) {
                    T acc = T(0);
#pragma unroll
                    for (int l = 0; l < N; l++)
                        if (bdnz(l, r)) acc += Bd[l * M + r] * (g[l * M + s] + mu * Bd[l * M + s]);
                    QuuR[r * M + s] = Cuu[r * M + s] + acc;
                }
#pragma unroll
                for (int j = 0; j < N; j++) {
                    T acc = T(0);
#pragma unroll
                    for (int l = 0; l < N; l++)
                        if (bdnz(l, r)) acc += Bd[l * M + r] * (W[l * N + j] + mu * Ad[l * N + j]);
                    QuxR[r * N + j] = ((CFD || P::cxu_nz(j * M + r)) ? Cxu[j * M + r] * cdt : T(0)) + acc;
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < M; r++)
#pragma unroll
                for (int s = 0; s < M; s++) QuuR[r * M + s] = Quu[r * M + s] + ((a.reg_type == 2 && r == s) ? mu : T(0));
#pragma unroll
            for (int i = 0; i < M * N; i++) QuxR[i] = Qux[i];
        }
        // positive-definiteness test = np.linalg.cholesky on the lower triangle (ilqr.py:297-302)
        T kt[M], Kt[M * N];
        if constexpr (M == 1) {
            if (!(QuuR[0] > T(0))) ok = false;
        } else {
            const T l11sq = QuuR[0];
            if (!(l11sq > T(0))) {
                ok = false;
            } else {
                const T l21 = QuuR[2] / pg::sqrt(l11sq);
                if (!(QuuR[3] - l21 * l21 > T(0))) ok = false;
            }
        }
        if (!ok) break;
        // kt = -QuuReg^-1 qu, Kt = -QuuReg^-1 QuxReg (ilqr.py:330-331)
        if constexpr (M == 1) {
            const T q = QuuR[0];
            kt[0] = -(qu[0] / q);
#pragma unroll
            for (int j = 0; j < N; j++) Kt[j] = -(QuxR[j] / q);
            if (a.constrained) {
                // box QP in one dimension (ilqr.py:304-327): kt = clip(-qu/Quu, -(lim+u), lim-u);
                // clamp detection keeps the reference's tolerance AND its lower-bound sign quirk
                const T up = a.ulim[0] - u[0], lo = a.ulim[0] + u[0];
                kt[0] = pg::clamp(kt[0], -lo, up);
                const bool clamped = np_isclose(kt[0], up, T(1e-3)) || np_isclose(kt[0], lo, T(1e-3));
                if (clamped) {
#pragma unroll
                    for (int j = 0; j < N; j++) Kt[j] = T(0);
                }
            }
        } else {
            // 2x2 solve by LU with partial pivoting (what LAPACK gesv does behind np.linalg.solve)
            T a00 = QuuR[0], a01 = QuuR[1], a10 = QuuR[2], a11 = QuuR[3];
            const bool swap = fabs(a10) > fabs(a00);
            if (swap) {
                T tmp = a00; a00 = a10; a10 = tmp;
                tmp = a01; a01 = a11; a11 = tmp;
            }
            const T l = a10 / a00;
            const T u11 = a11 - l * a01;
            {
                T b0 = swap ? qu[1] : qu[0], b1 = swap ? qu[0] : qu[1];
                b1 = b1 - l * b0;
                const T y1 = b1 / u11;
                const T y0 = (b0 - a01 * y1) / a00;
                kt[0] = -y0;
                kt[1] = -y1;
            }
#pragma unroll
            for (int j = 0; j < N; j++) {
                T b0 = swap ? QuxR[N + j] : QuxR[j], b1 = swap ? QuxR[j] : QuxR[N + j];
                b1 = b1 - l * b0;
                const T y1 = b1 / u11;
                const T y0 = (b0 - a01 * y1) / a00;
                Kt[j] = -y0;
                Kt[N + j] = -y1;
            }
            if (a.constrained) {
                // exact 2-D box QP by active-set enumeration; same clamp tolerance/quirk as above
                T up[2], lo[2];
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    up[r] = a.ulim[r] - u[r];
                    lo[r] = a.ulim[r] + u[r];
                }
                T best[2] = {kt[0], kt[1]};
                const bool inside = kt[0] <= up[0] && kt[0] >= -lo[0] && kt[1] <= up[1] && kt[1] >= -lo[1];
                if (!inside) {
                    T bestv = T(0);
                    bool have = false;
                    auto val = [&](T k0, T k1) {
                        return T(0.5) * (QuuR[0] * k0 * k0 + (QuuR[1] + QuuR[2]) * k0 * k1 + QuuR[3] * k1 * k1) + qu[0] * k0 + qu[1] * k1;
                    };
                    // one coordinate clamped, the other free (then clipped), and both clamped
#pragma unroll
                    for (int c = 0; c < 2; c++) {
#pragma unroll
                        for (int sgn = 0; sgn < 2; sgn++) {
                            T kc = sgn ? up[c] : -lo[c];
                            const int f = 1 - c;
                            T kf = -(qu[f] + T(0.5) * (QuuR[f * 2 + c] + QuuR[c * 2 + f]) * kc) / QuuR[f * 2 + f];
                            kf = pg::clamp(kf, -lo[f], up[f]);
                            const T k0 = c == 0 ? kc : kf, k1 = c == 0 ? kf : kc;
                            const T v = val(k0, k1);
                            if (!have || v < bestv) {
                                have = true;
                                bestv = v;
                                best[0] = k0;
                                best[1] = k1;
                            }
                        }
                    }
                    kt[0] = best[0];
                    kt[1] = best[1];
                }
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const bool clamped = np_isclose(kt[r], up[r], T(1e-3)) || np_isclose(kt[r], lo[r], T(1e-3));
                    if (clamped) {
#pragma unroll
                        for (int j = 0; j < N; j++) Kt[r * N + j] = T(0);
                    }
                }
            }
        }
        // dV1 = kt' qu ; dV2 = 0.5 kt' Quu kt (ilqr.py:337-338)
        T Qk[M];  // Quu kt
#pragma unroll
        for (int r = 0; r < M; r++) {
            T acc = T(0);
#pragma unroll
            for (int s = 0; s < M; s++) acc += Quu[r * M + s] * kt[s];
            Qk[r] = acc;
        }
        T dV1 = T(0), dV2 = T(0);
#pragma unroll
        for (int r = 0; r < M; r++) {
            dV1 += kt[r] * qu[r];
            dV2 += kt[r] * Qk[r];
        }
        sV1 += dV1;
        sV2 += T(0.5) * dV2;
        // vx = qx + Kt' Quu kt + Kt' qu + Qux' kt (ilqr.py:334)
#pragma unroll
        for (int i = 0; i < N; i++) {
            T t1 = T(0), t2 = T(0), t3 = T(0);
#pragma unroll
            for (int r = 0; r < M; r++) {
                t1 += Kt[r * N + i] * Qk[r];
                t2 += Kt[r * N + i] * qu[r];
                t3 += Qux[r * N + i] * kt[r];
            }
            vx[i] = ((qx[i] + t1) + t2) + t3;
        }
        // Vxx = Qxx + Kt' Quu Kt + Kt' Qux + Qux' Kt, then 0.5 (Vxx + Vxx') (ilqr.py:335-336);
        // Qxx = Cxx + Ad' W (ilqr.py:287)
        T QK[M * N];  // Quu Kt
#pragma unroll
        for (int r = 0; r < M; r++)
#pragma unroll
            for (int j = 0; j < N; j++) {
                T acc = T(0);
#pragma unroll
                for (int s = 0; s < M; s++) acc += Quu[r * M + s] * Kt[s * N + j];
                QK[r * N + j] = acc;
            }
        T Vn[N * N];
#pragma unroll
        for (int i = 0; i < N; i++)
#pragma unroll
            for (int j = 0; j < N; j++) {
                T acc = T(0);
#pragma unroll
                for (int l = 0; l < N; l++)
                    if (adnz(l, i)) acc += Ad[l * N + i] * W[l * N + j];
                T q = ((CFD || P::cxx_nz(i * N + j)) ? Cxx[i * N + j] * cdt : T(0)) + acc;
                T t1 = T(0), t2 = T(0), t3 = T(0);
#pragma unroll
                for (int r = 0; r < M; r++) {
                    t1 += Kt[r * N + i] * QK[r * N + j];
                    t2 += Kt[r * N + i] * Qux[r * N + j];
                    t3 += Qux[r * N + i] * Kt[r * N + j];
                }
                Vn[i * N + j] = ((q + t1) + t2) + t3;
            }
#pragma unroll
        for (int i = 0; i < N; i++)
#pragma unroll
            for (int j = 0; j < N; j++) Vxx[i * N + j] = T(0.5) * (Vn[i * N + j] + Vn[j * N + i]);
        // gradient norm ingredients: max_k |k| / (|u| + 1) (ilqr.py:418-420)
#pragma unroll
        for (int r = 0; r < M; r++) gmax[r] = fmax(gmax[r], fabs(kt[r]) / (fabs(u[r]) + T(1)));
        // KK.insert(0, Kt); kk.insert(0, kt)
#pragma unroll
        for (int r = 0; r < M * N; r++) a.KK[((int64_t)k * M * N + r) * B + b] = Kt[r];
#pragma unroll
        for (int r = 0; r < M; r++) a.kk[((int64_t)k * M + r) * B + b] = kt[r];
    }
    a.dV[b] = sV1;
    a.dV[B + b] = sV2;
    T gsum = T(0);
#pragma unroll
    for (int r = 0; r < M; r++) gsum += gmax[r];
    a.gnorm[b] = gsum / T(M);
    a.success[b] = ok ? 1 : 0;
}

template <typename P, typename T, int LIN, bool CFD = false>
__global__ void __launch_bounds__(128) ilqr_backward_kernel(const BackwardArgs<T> a) {
    const typename P::template Ctx<T> m;
    const int64_t B = a.B;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t b = tid;
    if (a.index) {
        if (tid >= *a.count) return;
        b = a.index[tid];
    } else if (b >= B) {
        return;
    }
    if (a.active && !a.active[b]) return;
    backward_body<P, T, LIN, CFD>(m, a, b);
}

// ---- the backward sweep of ONE trajectory by the lanes of a warp (ilqr_solve_kernel) ---------------------------------------
// The sweep of backward_body, value for value, laid over a warp in two phases per chunk of 32 control intervals:
//   A  the linearisation and the cost expansion do not depend on the value function, only on the nominal point: lane s
//      evaluates them for interval k_hi - s and leaves (Ad, Bd, the raw cost derivatives, u) in row s of the warp's
//      shared-memory staging area -- 32 intervals for the instructions of one;
//   B  the Riccati recursion walks the rows: lane j owns COLUMN j of W = Vxx Ad, of Qux, of K and of the new Vxx; the
//      scalars (Quu, k, dV) are computed by every lane.  Two exchanges per interval through shared memory: (K, Qux) of all
//      columns before the value update, the new columns (and vx) after it.
// Bit-identical to backward_body: every sum runs over the same terms in the same order (the structural zeros the
// single-thread code skips are exact zeros here: x + v * 0 == x), products meet their addends in the same expressions,
// so the compiler contracts the same multiply-adds.  M == 1 (the fused solver's scope).
template <typename P>
struct LaneSweep {
    static constexpr int N = P::N, M = P::M;
    static constexpr int oAd = 0, oCxx = oAd + N * N, oBd = oCxx + N * N, oCxu = oBd + N * M, oCx = oCxu + N * M, oCu = oCx + N,
                         oCuu = oCu + M, oU = oCuu + M * M, E = oU + M;
    // row stride in words of T: even (16-byte rows) and = 2 mod 4, so that the 32 rows phase A writes side by side fall on 8
    // distinct 16-byte bank groups (2-way conflicts at most)
    __host__ __device__ static constexpr int stride() {
        int s = (E + 1) & ~1;
        while (s % 4 != 2) s += 2;
        return s;
    }
    static constexpr int xVn = 0, xVx = xVn + N * N + (N * N & 1), xKt = xVx + N + (N & 1), xQux = xKt + N + (N & 1),
                         X = xQux + N + (N & 1);
    __host__ __device__ static constexpr int warp_words() { return 32 * stride() + X; }
};

template <typename T, int K, bool ALIGNED>
__device__ __forceinline__ void lds_vec(const T* p, T (&d)[K]) {
    if constexpr (ALIGNED && sizeof(T) == 8 && K % 2 == 0) {
#pragma unroll
        for (int q = 0; q < K / 2; q++) {
            const double2 v = reinterpret_cast<const double2*>(p)[q];
            d[2 * q] = v.x;
            d[2 * q + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int q = 0; q < K; q++) d[q] = p[q];
    }
}

template <typename P, typename T, int LIN>
__device__ __forceinline__ void backward_lanes(const typename P::template Ctx<T>& m, const BackwardArgs<T>& a, T* sm, const int lane) {
    using L = LaneSweep<P>;
    constexpr int N = P::N, M = P::M, S = L::stride();
    static_assert(M == 1, "lane-parallel sweep: one control input");
    const int j = lane % N;      // lanes >= N repeat the work of lane j (uniform control flow); only lanes < N publish
    const bool writer = lane < N;
    T* ex = sm + 32 * S;
    const T mu = a.mu[0];
    const T dt = a.dt;
    const int Ts = a.T_steps;
    T vx[N], Vxx[N * N];
    {
        T x[N];
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = a.xbar[(int64_t)Ts * N + i];
        P::fcost_derivs(m, x, vx, Vxx);
#pragma unroll
        for (int i = 0; i < N; i++) vx[i] *= dt;
#pragma unroll
        for (int i = 0; i < N * N; i++) Vxx[i] = P::cfxx_nz(i) ? Vxx[i] * dt : T(0);
    }
    T sV1 = T(0), sV2 = T(0);
    T gmax[M];
#pragma unroll
    for (int r = 0; r < M; r++) gmax[r] = T(0);
    bool ok = true;
    auto bdnz = [](int l, int r) constexpr { return LIN != 0 || P::b_nz(l * M + r); };
    for (int k_hi = Ts - 1; k_hi >= 0 && ok; k_hi -= 32) {
        __syncwarp();  // the rows of the previous chunk have been consumed
        {              // phase A: lane s <-> interval k_hi - s
            const int k = k_hi - lane;
            if (k >= 0) {
                T x[N], u[M];
#pragma unroll
                for (int i = 0; i < N; i++) x[i] = a.xbar[(int64_t)k * N + i];
#pragma unroll
                for (int r = 0; r < M; r++) u[r] = a.ubar[(int64_t)k * M + r];
                T Ad[N * N], Bd[N * M];
                if constexpr (LIN == 0) P::AB(m, x, u, Ad, Bd);
                else fd_AB<P, T>(m, x, u, Ad, Bd);
#pragma unroll
                for (int i = 0; i < N; i++)
#pragma unroll
                    for (int c = 0; c < N; c++) Ad[i * N + c] = Ad[i * N + c] * dt + (i == c ? T(1) : T(0));
#pragma unroll
                for (int i = 0; i < N * M; i++) Bd[i] = Bd[i] * dt;
                T cx[N], cu[M], Cxx[N * N], Cuu[M * M], Cxu[N * M];
                P::cost_derivs(m, x, u, T(k) * dt, cx, cu, Cxx, Cuu, Cxu);
                T* row = sm + lane * S;
#pragma unroll
                for (int i = 0; i < N * N; i++) row[L::oAd + i] = Ad[i];
#pragma unroll
                for (int i = 0; i < N * N; i++) row[L::oCxx + i] = Cxx[i];
#pragma unroll
                for (int i = 0; i < N * M; i++) row[L::oBd + i] = Bd[i];
#pragma unroll
                for (int i = 0; i < N * M; i++) row[L::oCxu + i] = Cxu[i];
#pragma unroll
                for (int i = 0; i < N; i++) row[L::oCx + i] = cx[i];
#pragma unroll
                for (int r = 0; r < M; r++) row[L::oCu + r] = cu[r];
#pragma unroll
                for (int i = 0; i < M * M; i++) row[L::oCuu + i] = Cuu[i];
#pragma unroll
                for (int r = 0; r < M; r++) row[L::oU + r] = u[r];
            }
        }
        __syncwarp();
        const int nk = k_hi + 1 < 32 ? k_hi + 1 : 32;
        for (int s = 0; s < nk; s++) {  // phase B: the recursion over the staged intervals
            const int k = k_hi - s;
            const T* e = sm + s * S;
            T Ad[N * N], Acol[N], Bd[N * M];
            lds_vec<T, N * N, true>(e + L::oAd, Ad);
            lds_vec<T, N * M, (L::oBd % 2 == 0)>(e + L::oBd, Bd);
#pragma unroll
            for (int l = 0; l < N; l++) Acol[l] = e[L::oAd + l * N + j];
            const T cxj = e[L::oCx + j] * dt;
            const T cu0 = e[L::oCu] * dt;
            const T Cuu0 = e[L::oCuu] * dt;
            const T u0 = e[L::oU];
            // column j of W = Vxx Ad; g = Vxx Bd
            T W[N], g[N];
#pragma unroll
            for (int i = 0; i < N; i++) {
                T acc = T(0);
#pragma unroll
                for (int l = 0; l < N; l++) acc += Vxx[i * N + l] * Acol[l];
                W[i] = acc;
                T acg = T(0);
#pragma unroll
                for (int l = 0; l < N; l++)
                    if (bdnz(l, 0)) acg += Vxx[i * N + l] * Bd[l];
                g[i] = acg;
            }
            // qx[j], qu
            T qxj, qu;
            {
                T acc = T(0);
#pragma unroll
                for (int l = 0; l < N; l++) acc += Acol[l] * vx[l];
                qxj = cxj + acc;
                T acu = T(0);
#pragma unroll
                for (int l = 0; l < N; l++)
                    if (bdnz(l, 0)) acu += Bd[l] * vx[l];
                qu = cu0 + acu;
            }
            // Quu, Qux[j]
            T Quu, Quxj, QuuR, QuxRj;
            {
                T acc = T(0);
#pragma unroll
                for (int l = 0; l < N; l++)
                    if (bdnz(l, 0)) acc += Bd[l] * g[l];
                Quu = Cuu0 + acc;
                T acx = T(0);
#pragma unroll
                for (int l = 0; l < N; l++)
                    if (bdnz(l, 0)) acx += Bd[l] * W[l];
                Quxj = e[L::oCxu + j] * dt + acx;
            }
            if (a.reg_type == 1) {
                T acc = T(0);
#pragma unroll
                for (int l = 0; l < N; l++)
                    if (bdnz(l, 0)) acc += Bd[l] * (g[l] + mu * Bd[l]);
                QuuR = Cuu0 + acc;
                T acx = T(0);
#pragma unroll
                for (int l = 0; l < N; l++)
                    if (bdnz(l, 0)) acx += Bd[l] * (W[l] + mu * Acol[l]);
                QuxRj = e[L::oCxu + j] * dt + acx;
            } else {
                QuuR = Quu + (a.reg_type == 2 ? mu : T(0));
                QuxRj = Quxj;
            }
            if (!(QuuR > T(0))) {
                ok = false;
                break;
            }
            T kt = -(qu / QuuR);
            T Ktj = -(QuxRj / QuuR);
            if (a.constrained) {
                const T up = a.ulim[0] - u0, lo = a.ulim[0] + u0;
                kt = pg::clamp(kt, -lo, up);
                const bool clamped = np_isclose(kt, up, T(1e-3)) || np_isclose(kt, lo, T(1e-3));
                if (clamped) Ktj = T(0);
            }
            if (writer) {
                ex[L::xKt + j] = Ktj;
                ex[L::xQux + j] = Quxj;
            }
            T Qk = T(0);
            Qk += Quu * kt;
            T dV1 = T(0), dV2 = T(0);
            dV1 += kt * qu;
            dV2 += kt * Qk;
            sV1 += dV1;
            sV2 += T(0.5) * dV2;
            T vxn;
            {
                T t1 = T(0), t2 = T(0), t3 = T(0);
                t1 += Ktj * Qk;
                t2 += Ktj * qu;
                t3 += Quxj * kt;
                vxn = ((qxj + t1) + t2) + t3;
            }
            T QKj = T(0);
            QKj += Quu * Ktj;
            __syncwarp();
            T Kt[N], Qux[N];
            lds_vec<T, N, true>(ex + L::xKt, Kt);
            lds_vec<T, N, true>(ex + L::xQux, Qux);
            T Vn[N];
#pragma unroll
            for (int i = 0; i < N; i++) {
                T acc = T(0);
#pragma unroll
                for (int l = 0; l < N; l++) acc += Ad[l * N + i] * W[l];
                T q = e[L::oCxx + i * N + j] * dt + acc;
                T t1 = T(0), t2 = T(0), t3 = T(0);
                t1 += Kt[i] * QKj;
                t2 += Kt[i] * Quxj;
                t3 += Qux[i] * Ktj;
                Vn[i] = ((q + t1) + t2) + t3;
            }
            if (writer) {
#pragma unroll
                for (int i = 0; i < N; i++) ex[L::xVn + i * N + j] = Vn[i];
                ex[L::xVx + j] = vxn;
                a.KK[(int64_t)k * N + j] = Ktj;
                if (lane == 0) a.kk[k] = kt;
            }
            gmax[0] = fmax(gmax[0], fabs(kt) / (fabs(u0) + T(1)));
            __syncwarp();
            T Vf[N * N];
            lds_vec<T, N * N, true>(ex + L::xVn, Vf);
            lds_vec<T, N, true>(ex + L::xVx, vx);
#pragma unroll
            for (int i = 0; i < N; i++)
#pragma unroll
                for (int l = 0; l < N; l++) Vxx[i * N + l] = T(0.5) * (Vf[i * N + l] + Vf[l * N + i]);
        }
    }
    if (lane == 0) {
        a.dV[0] = sV1;
        a.dV[1] = sV2;
        a.gnorm[0] = gmax[0] / T(M);
        a.success[0] = ok ? 1 : 0;
    }
}

// ---- terminal value of the final backward pass (iLQR.backward_pass(final=True), ilqr.py:246-260) ---------------------------
// Per trajectory, on the device: (1) the equilibrium of the final cost, x* = argmin fcost(x) started at x_N (the reference
// calls scipy.optimize.minimize; here Newton's method on the generated gradient / Hessian, exact in one step for the
// quadratic final costs of every shipped example); (2) Ad, Bd and the cost expansion at (x*, u = 0, t = t_N) exactly as the
// backward sweep forms them; (3) P = the stabilising solution of the discrete algebraic Riccati equation
//       P = Ad' P Ad - (Ad' P Bd + N)(R + Bd' P Bd)^-1 (Bd' P Ad + N') + Q         (scipy.linalg.solve_discrete_are(A, B, Q, R, s=N))
// by the structure-preserving doubling algorithm on (A0, G0, H0) = (Ad - Bd R^-1 N', Bd R^-1 Bd', Q - N R^-1 N'):
//       W = I + G H;  A+ = A W^-1 A;  G+ = G + A W^-1 G A';  H+ = H + A' H W^-1 A;      H_k -> P quadratically
// (Chu, Fan, Lin, Wang 2004).  One thread per trajectory, n <= 8 matrices in registers / local memory.
template <typename T>
struct TerminalArgs {
    int64_t B;
    int T_steps;
    T dt, t_final;
    const T* xbar;   // [T+1, n, B]
    T* P_out;        // [n, n, B]
    T* xeq_out;      // [n, B] or NULL
    uint8_t* ok_out; // [B]: 1 = Newton and the doubling iteration converged
    const int32_t* index;
    const int32_t* count;
};

// solve W Z = R in place (Gauss elimination with partial pivoting), W [n x n], R [n x nr]; returns false if singular
template <typename T, int N, int NR>
__device__ __forceinline__ bool gauss_solve(T (&W)[N * N], T (&R)[N * NR]) {
    for (int c = 0; c < N; c++) {
        int piv = c;
        T best = fabs(W[c * N + c]);
        for (int r = c + 1; r < N; r++) {
            const T v = fabs(W[r * N + c]);
            if (v > best) {
                best = v;
                piv = r;
            }
        }
        if (!(best > T(0))) return false;
        if (piv != c) {
            for (int j = 0; j < N; j++) {
                const T t = W[c * N + j];
                W[c * N + j] = W[piv * N + j];
                W[piv * N + j] = t;
            }
            for (int j = 0; j < NR; j++) {
                const T t = R[c * NR + j];
                R[c * NR + j] = R[piv * NR + j];
                R[piv * NR + j] = t;
            }
        }
        const T inv = T(1) / W[c * N + c];
        for (int r = c + 1; r < N; r++) {
            const T f = W[r * N + c] * inv;
            if (f != T(0)) {
                for (int j = c + 1; j < N; j++) W[r * N + j] -= f * W[c * N + j];
                for (int j = 0; j < NR; j++) R[r * NR + j] -= f * R[c * NR + j];
            }
        }
    }
    for (int c = N - 1; c >= 0; c--) {
        const T inv = T(1) / W[c * N + c];
        for (int j = 0; j < NR; j++) {
            T acc = R[c * NR + j];
            for (int k = c + 1; k < N; k++) acc -= W[c * N + k] * R[k * NR + j];
            R[c * NR + j] = acc * inv;
        }
    }
    return true;
}

template <typename P, typename T, int LIN>
__global__ void __launch_bounds__(64) terminal_value_kernel(const TerminalArgs<T> a) {
    const typename P::template Ctx<T> m;
    constexpr int N = P::N, M = P::M;
    const int64_t B = a.B;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t b = tid;
    if (a.index) {
        if (tid >= *a.count) return;
        b = a.index[tid];
    } else if (b >= B) {
        return;
    }
    T x[N], u[M];
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = a.xbar[((int64_t)a.T_steps * N + i) * B + b];
#pragma unroll
    for (int r = 0; r < M; r++) u[r] = T(0);
    bool ok = true;
    // (1) equilibrium of the final cost: Newton steps x <- x - Cfxx^-1 cfx
    for (int it = 0; it < 50; it++) {
        T g[N], Hm[N * N];
        P::fcost_derivs(m, x, g, Hm);
        T gmax = T(0);
#pragma unroll
        for (int i = 0; i < N; i++) gmax = fmax(gmax, fabs(g[i]));
        if (!(gmax > T(0))) break;
        T rhs[N * 1];
#pragma unroll
        for (int i = 0; i < N; i++) rhs[i] = g[i];
        // a final cost that does not weigh every state (the pendulum chains: velocities) has a singular Hessian; the gradient
        // vanishes along those directions and the minimiser keeps x_N's components there, as scipy's BFGS does: a relative
        // 1e-14 on the diagonal makes the step well defined (and costs one more iteration elsewhere)
        T hdiag = T(0);
#pragma unroll
        for (int i = 0; i < N; i++) hdiag = fmax(hdiag, fabs(Hm[i * N + i]));
#pragma unroll
        for (int i = 0; i < N; i++) Hm[i * N + i] += T(1e-14) * hdiag;
        if (!gauss_solve<T, N, 1>(Hm, rhs)) {
            ok = false;
            break;
        }
        T dmax = T(0), xmax = T(0);
#pragma unroll
        for (int i = 0; i < N; i++) {
            x[i] -= rhs[i];
            dmax = fmax(dmax, fabs(rhs[i]));
            xmax = fmax(xmax, fabs(x[i]));
        }
        if (dmax <= T(1e-14) * (xmax + T(1))) break;
        if (it == 49) ok = false;
    }
    // (2) linearisation and cost expansion at (x*, 0, t_N), as the backward sweep forms them
    T A[N * N], Bd[N * M];
    if constexpr (LIN == 0) P::AB(m, x, u, A, Bd);
    else fd_AB<P, T>(m, x, u, A, Bd);
    const T dt = a.dt;
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) A[i * N + j] = A[i * N + j] * dt + (i == j ? T(1) : T(0));
#pragma unroll
    for (int i = 0; i < N * M; i++) Bd[i] = Bd[i] * dt;
    T cx[N], cu[M], Q[N * N], R[M * M], Nx[N * M];
    P::cost_derivs(m, x, u, a.t_final, cx, cu, Q, R, Nx);
#pragma unroll
    for (int i = 0; i < N * N; i++) Q[i] *= dt;
#pragma unroll
    for (int i = 0; i < M * M; i++) R[i] *= dt;
#pragma unroll
    for (int i = 0; i < N * M; i++) Nx[i] *= dt;
    // R^-1 (m <= 2)
    T Ri[M * M];
    if constexpr (M == 1) {
        Ri[0] = T(1) / R[0];
    } else {
        const T det = R[0] * R[3] - R[1] * R[2];
        Ri[0] = R[3] / det;
        Ri[1] = -R[1] / det;
        Ri[2] = -R[2] / det;
        Ri[3] = R[0] / det;
    }
    // BR = Bd R^-1 [n x m], NR = N R^-1 [n x m]
    T BR[N * M], NR[N * M];
    for (int i = 0; i < N; i++)
        for (int r = 0; r < M; r++) {
            T s1 = T(0), s2 = T(0);
            for (int q = 0; q < M; q++) {
                s1 += Bd[i * M + q] * Ri[q * M + r];
                s2 += Nx[i * M + q] * Ri[q * M + r];
            }
            BR[i * M + r] = s1;
            NR[i * M + r] = s2;
        }
    // A0 = Ad - BR N', G0 = BR Bd', H0 = Q - NR N'
    T G[N * N], H[N * N];
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
            T s1 = T(0), s2 = T(0), s3 = T(0);
            for (int r = 0; r < M; r++) {
                s1 += BR[i * M + r] * Nx[j * M + r];
                s2 += BR[i * M + r] * Bd[j * M + r];
                s3 += NR[i * M + r] * Nx[j * M + r];
            }
            A[i * N + j] -= s1;
            G[i * N + j] = s2;
            H[i * N + j] = Q[i * N + j] - s3;
        }
    // (3) doubling
    bool conv = false;
    for (int it = 0; it < 60 && !conv; it++) {
        T W[N * N], Z[N * 2 * N];  // Z = [A | G] -> W^-1 [A | G]
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) {
                T s = (i == j) ? T(1) : T(0);
                for (int k = 0; k < N; k++) s += G[i * N + k] * H[k * N + j];
                W[i * N + j] = s;
                Z[i * 2 * N + j] = A[i * N + j];
                Z[i * 2 * N + N + j] = G[i * N + j];
            }
        if (!gauss_solve<T, N, 2 * N>(W, Z)) {
            ok = false;
            break;
        }
        // A+ = A Z1; G+ = G + A Z2 A'; H+ = H + A' H Z1
        T A1[N * N], HZ[N * N], AZ2[N * N];
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) {
                T s1 = T(0), s2 = T(0), s3 = T(0);
                for (int k = 0; k < N; k++) {
                    s1 += A[i * N + k] * Z[k * 2 * N + j];
                    s2 += H[i * N + k] * Z[k * 2 * N + j];
                    s3 += A[i * N + k] * Z[k * 2 * N + N + j];
                }
                A1[i * N + j] = s1;
                HZ[i * N + j] = s2;
                AZ2[i * N + j] = s3;
            }
        T dmax = T(0), hmax = T(0);
        T G1[N * N], H1[N * N];
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) {
                T s1 = T(0), s2 = T(0);
                for (int k = 0; k < N; k++) {
                    s1 += AZ2[i * N + k] * A[j * N + k];
                    s2 += A[k * N + i] * HZ[k * N + j];
                }
                G1[i * N + j] = G[i * N + j] + s1;
                H1[i * N + j] = H[i * N + j] + s2;
                dmax = fmax(dmax, fabs(s2));
                hmax = fmax(hmax, fabs(H1[i * N + j]));
            }
        for (int i = 0; i < N * N; i++) {
            A[i] = A1[i];
            G[i] = G1[i];
            H[i] = H1[i];
        }
        conv = dmax <= T(1e-15) * hmax;
        if (!(hmax == hmax) || hmax > T(1e300)) {
            ok = false;
            break;
        }
    }
    ok = ok && conv;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) a.P_out[((int64_t)(i * N + j)) * B + b] = T(0.5) * (H[i * N + j] + H[j * N + i]);
    if (a.xeq_out)
        for (int i = 0; i < N; i++) a.xeq_out[(int64_t)i * B + b] = x[i];
    a.ok_out[b] = ok ? 1 : 0;
}

// ---- K4: line-search decision, mu schedule, stopping tests (iLQR.run_optim, ilqr.py:407-485) -------
struct SelectParams {
    int n_alpha, parallel, iteration, max_iters;
    double alphas[MAX_ALPHAS];
    double zmin, tol_fun, tol_grad, mu_min, mu_max, mu_d0;
};

template <typename T>
struct SelectArgs {
    int64_t B;
    SelectParams p;
    const T* cost_cand;
    const uint8_t* diverged;
    const T* dV;
    const T* gnorm;
    const uint8_t* bw_success;
    T* cost;
    T* mu;
    T* mu_d;
    T* dcost;
    uint8_t* success_gradient;
    int32_t* status;
    int32_t* iters;
    T* alpha_best;
    uint8_t* accept;
    uint8_t* active;
    int32_t* n_active;
    const int32_t* index_in;  // optional work list of this iteration ...
    const int32_t* count_in;
    int32_t* index_out;       // ... and the compacted list of trajectories still running after it
    int32_t* last_tried;      // [B] or NULL: index of the last alpha the serial search evaluated, -1 if no forward pass ran
};

template <typename T>
__device__ __forceinline__ int select_one(const SelectArgs<T>& a, const int64_t b) {
    const int64_t B = a.B;
    const SelectParams& p = a.p;
    // Everything the decision may read is loaded up front, in one batch of independent loads: taken in program order -- status,
    // then mu, then the backward pass's flag, then one candidate after the other until the search stops -- the same reads are
    // a chain of ~15 dependent L2 round trips (17 us per launch, 10 us of every iteration of the fused solver).  Slots of
    // alphas a split line search did not evaluate hold older values; they are read, never used (same loop, same exits).
    const int status_in = a.status[b];
    const T mu_in = a.mu[b], mu_d_in = a.mu_d[b];
    const int iters_in = a.iters[b];
    const bool bw_ok = a.bw_success[b] != 0, sg_in = a.success_gradient[b] != 0;
    const T gnorm_in = a.gnorm[b], cost_in = a.cost[b], sV1_in = a.dV[b], sV2_in = a.dV[B + b], dcost_in = a.dcost[b];
    T cc[MAX_ALPHAS];
    bool dvg[MAX_ALPHAS];
#pragma unroll
    for (int i = 0; i < MAX_ALPHAS; i++) {
        const bool in = i < p.n_alpha;
        cc[i] = in ? a.cost_cand[(int64_t)i * B + b] : T(0);
        dvg[i] = in ? a.diverged[(int64_t)i * B + b] != 0 : true;
    }
    a.accept[b] = 0;
    if (status_in != 0) {
        a.active[b] = 0;
        return status_in;
    }
    T mu = mu_in, mu_d = mu_d_in;
    const T mu_d0 = T(p.mu_d0), mu_min = T(p.mu_min), mu_max = T(p.mu_max);
    auto increase_mu = [&]() {  // ilqr.py:618-622
        mu_d = fmax(mu_d0, mu_d0 * mu_d);
        mu = fmax(mu_min, mu * mu_d);
    };
    auto decrease_mu = [&]() {  // ilqr.py:612-616
        mu_d = fmin(mu_d / mu_d0, T(1) / mu_d0);
        mu = mu * mu_d * (mu > mu_min ? T(1) : T(0));
    };
    int status = 0;
    const int it = iters_in + 1;
    a.iters[b] = it;
    if (a.last_tried) a.last_tried[b] = -1;
    if (!bw_ok) {
        increase_mu();  // ilqr.py:411-413
        increase_mu();  // ... and again on the "forward not successful" branch (ilqr.py:478-479)
        if (mu > mu_max) status = 3;
    } else {
        bool sg = sg_in;
        if (gnorm_in < T(p.tol_grad) && mu < T(1e-5)) {  // ilqr.py:421-423
            decrease_mu();
            sg = true;
        }
        const T cost_old = cost_in;
        const T sV1 = sV1_in, sV2 = sV2_in;
        T dcost = dcost_in;
        bool success_fw = false, searching = true;
        int n_tried = 0;
#pragma unroll
        for (int i = 0; i < MAX_ALPHAS; i++) {
            if (searching && i < p.n_alpha) {
                n_tried = i + 1;
                if (dvg[i]) {  // 'forward diverged.' (ilqr.py:435-437)
                    searching = false;
                } else {
                    const T alpha = T(p.alphas[i]);
                    dcost = cost_old - cc[i];
                    const T expected = -alpha * (sV1 + alpha * sV2);
                    const T z = expected > T(0) ? dcost / expected : pg::sign(dcost);
                    if (z > T(p.zmin)) {
                        success_fw = true;
                        if (!p.parallel) searching = false;
                    }
                }
            }
        }
        if (a.last_tried) a.last_tried[b] = n_tried - 1;
        if (success_fw) {
            // best_idx = np.argmin(cost_list): first minimum (ilqr.py:458).  A NaN candidate never wins here: the reference
            // cannot produce one on this path (numpy's sin/cos are finite for any finite angle), while pg::sincos returns NaN
            // beyond |x| >= 2^30 -- a candidate that diverged towards -inf, which the one-sided test x > 1e8 does not flag.
            int best = -1;
            T bc = T(0);
#pragma unroll
            for (int i = 0; i < MAX_ALPHAS; i++) {
                const T c = cc[i];
                if (i < n_tried && c == c && (best < 0 || c < bc)) {
                    bc = c;
                    best = i;
                }
            }
            if (best < 0) {  // every tried candidate is NaN (cannot pass the z-test, so this is unreachable): keep alpha 0
                best = 0;
                bc = cc[0];
            }
            a.cost[b] = bc;
            a.alpha_best[b] = T(p.alphas[best]);
            a.accept[b] = 1;
            decrease_mu();
            if (dcost < T(p.tol_fun)) status = 1;
            else if (sg) status = 2;
        } else {
            increase_mu();
            if (mu > mu_max) status = 3;
        }
        a.dcost[b] = dcost;
        a.success_gradient[b] = sg ? 1 : 0;
    }
    if (status == 0 && it >= p.max_iters) status = 4;
    a.mu[b] = mu;
    a.mu_d[b] = mu_d;
    a.status[b] = status;
    a.active[b] = status == 0 ? 1 : 0;
    return status;
}

template <typename T>
__global__ void ilqr_select_kernel(const SelectArgs<T> a) {
    const int64_t B = a.B;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t b = tid;
    bool valid = true;
    if (a.index_in) {
        valid = tid < *a.count_in;
        b = valid ? a.index_in[tid] : 0;
    } else {
        valid = b < B;
    }
    int status = 1;
    if (valid) status = select_one(a, b);
    // compaction: trajectories still running are appended to index_out (warp-aggregated, order kept in a warp)
    const bool keep = valid && status == 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    if (ballot && a.n_active) {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == __ffs(ballot) - 1) base = atomicAdd(a.n_active, __popc(ballot));
        base = __shfl_sync(0xffffffffu, base, __ffs(ballot) - 1);
        if (keep && a.index_out) a.index_out[base + __popc(ballot & ((1u << lane) - 1))] = (int32_t)b;
    }
}

// ---- fused tail solver: one warp finishes one trajectory (iLQR.run_optim, ilqr.py:399-485) ---------------------------------
// Once few trajectories are still running, an iteration is a chain of dependent launches (backward sweep -> line search ->
// decision -> commit), each a latency-bound chain of T steps, with launch gaps between them and the slowest trajectory
// pacing everybody.  Trajectories are independent, so here ONE launch finishes them: a warp owns a trajectory and loops
//     lane 0        backward sweep (backward_body: exactly K3's arithmetic)
//     lanes 0..10   the 11 line-search candidates side by side (forward_body: exactly K2's arithmetic), each lane keeping
//                   its trajectory in the warp's candidate block (lane-minor: a warp store is one 128-byte line)
// on a compact private copy of the trajectory (nominal states and controls, gains), written back when it stops,
//     lane 0        the reference's decision logic (select_one: exactly K4's)
//     all lanes     winner -> nominal trajectory, candidate gains -> accepted gains (a copy: no re-integration)
// until the trajectory stops (converged / mu_max / max_iters), then takes the next one from the list.  No grid-wide
// synchronisation, no host in the loop, every decision bit-identical to the launch-per-step path (same device functions).
template <typename T>
struct SolveArgs {
    BackwardArgs<T> bw;
    ForwardArgs<T> fw;
    SelectArgs<T> sel;
    T* xbar;      // nominal trajectory, updated in place when a candidate is accepted
    T* ubar;
    T* KK_acc;    // accepted gains
    T* kk_acc;
    T* cand;      // [warps of the grid][solve_block_doubles(T)]: the private block of each warp (see ilqr_solve_kernel)
    int32_t* ticket;  // [0] next of the first n_first work-list entries, [1] next of the others (zeroed before the launch)
    int32_t* sm_rank; // [SMs] CTAs that have arrived on each SM (zeroed before the launch)
    int n_first;      // 4 x SMs: the entries that can each have a warp scheduler of their own
    int round_iters;  // iterations per trajectory in this launch; a trajectory still running after them goes on index_out
    const int32_t* index;
    const int32_t* count;
    int32_t* index_out;
    int32_t* count_out;
};

// Doubles of a warp's private block: the trajectory it is working on, compact (the batch-strided arrays give one useful word
// per 32-byte sector once the running trajectories are sparse in the batch: the tail of a 16,384-trajectory solve then
// misses L2 on every dependent load), its candidate and accepted gains, 16 scalars, and the 16 candidates lane-minor.
template <typename P>
__host__ __device__ constexpr size_t solve_block_doubles(int Ts) {
    return (size_t)P::N + (size_t)(Ts + 1) * P::N + (size_t)Ts * P::M + 2 * ((size_t)Ts * P::M * P::N + (size_t)Ts * P::M) + 64 +
           ((size_t)(Ts + 1) * P::N + (size_t)Ts * P::M) * 16;
}

template <typename P, typename T, int INTEG, int LIN>
__global__ void __launch_bounds__(128) ilqr_solve_kernel(const SolveArgs<T> a) {
    const typename P::template Ctx<T> m;
    constexpr int N = P::N, M = P::M;
    const int lane = threadIdx.x & 31;
    const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int64_t B = a.fw.B;
    const int Ts = a.fw.T_steps, nA = a.fw.n_alpha;
    const int cnt = *a.count;
    const int nX = (Ts + 1) * N, nU = Ts * M, nK = Ts * M * N;
    extern __shared__ __align__(16) unsigned char solve_smem[];  // M == 1: the staging rows of backward_lanes, per warp
    T* sm = reinterpret_cast<T*>(solve_smem) + (threadIdx.x >> 5) * LaneSweep<P>::warp_words();
    // the warp's private block
    T* px0 = a.cand + (size_t)warp * solve_block_doubles<P>(Ts);
    T* pX = px0 + N;
    T* pU = pX + nX;
    T* pKc = pU + nU;
    T* pkc = pKc + nK;
    T* pKa = pkc + nU;
    T* pka = pKa + nK;
    T* ps = pka + nU;  // [0] mu, [1..2] dV, [3] gnorm, [16..31] candidate costs
    uint8_t* pflag = reinterpret_cast<uint8_t*>(ps + 32);  // [0] backward success, [16..31] diverged, [32..47] terminated
    T* Xc = ps + 64;
    T* Uc = Xc + (size_t)nX * 16;
    // the kernels' argument blocks, redirected to the private copy: batch size 1, trajectory 0
    BackwardArgs<T> bw = a.bw;
    bw.B = 1;
    bw.mu = ps;
    bw.xbar = pX;
    bw.ubar = pU;
    bw.KK = pKc;
    bw.kk = pkc;
    bw.dV = ps + 1;
    bw.gnorm = ps + 3;
    bw.success = pflag;
    ForwardArgs<T> fw = a.fw;
    fw.B = 1;
    fw.x0 = px0;
    fw.xbar = pX;
    fw.ubar = pU;
    fw.KK = pKc;
    fw.kk = pkc;
    fw.cost_out = ps + 16;
    fw.diverged_out = pflag + 16;
    fw.terminated_out = pflag + 32;
    fw.xN_out = nullptr;
    // Placement: a 4-warp CTA puts one warp on each scheduler of its SM, and two CTAs fit an SM.  The first CTA to arrive on an
    // SM serves the first n_first = 4 x SMs entries of the work list, so that a list that short gives every trajectory a
    // scheduler of its own (two latency-bound warps on one scheduler run 1.5-2x slower each -- and the launch lasts as long
    // as its slowest trajectory); later arrivals serve the entries beyond, and so does everybody once the first are gone.
    __shared__ int s_rank;
    if (threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        s_rank = atomicAdd(a.sm_rank + smid, 1);
    }
    __syncthreads();
    bool first_open = s_rank == 0;
    const int n_first = cnt < a.n_first ? cnt : a.n_first;
    // a list that short needs no further dealing: this launch finishes it (and the launches queued behind it find it empty)
    const int round_iters = cnt > a.n_first ? a.round_iters : INT32_MAX;
    for (;;) {
        // trajectories are handed out one at a time: a warp whose trajectory stops early takes the next one, so the few
        // that run to max_iters never queue up behind each other on one warp
        int w = 0;
        if (lane == 0) {
            w = cnt;
            if (first_open) w = atomicAdd(a.ticket, 1);
            if (!first_open || w >= n_first) w = n_first + atomicAdd(a.ticket + 1, 1);
        }
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= n_first) first_open = false;
        if (w >= cnt) break;
        const int64_t b = a.index[w];
        for (int e = lane; e < N; e += 32) px0[e] = a.fw.x0[(int64_t)e * B + b];
        for (int e = lane; e < nX; e += 32) pX[e] = a.xbar[(int64_t)e * B + b];
        for (int e = lane; e < nU; e += 32) pU[e] = a.ubar[(int64_t)e * B + b];
        bool any_accept = false, running = false, last_is_accepted = false;
        __syncwarp();
        for (int it = 0;; it++) {
            if (it == round_iters) {  // this launch's share is done: the next launch deals the trajectory again
                running = true;
                break;
            }
            last_is_accepted = false;
            if constexpr (M == 1) {  // the sweep laid over the lanes (backward_lanes)
                bw.mu = a.sel.mu + b;
                backward_lanes<P, T, LIN>(m, bw, sm, lane);
            } else {
                if (lane == 0) {
                    ps[0] = a.sel.mu[b];
                    backward_body<P, T, LIN, false>(m, bw, 0);
                }
            }
            if (lane == 0) {
                a.bw.dV[b] = ps[1];
                a.bw.dV[B + b] = ps[2];
                a.bw.gnorm[b] = ps[3];
                a.bw.success[b] = pflag[0];
            }
            __syncwarp();
            if (pflag[0] && lane < nA) {
                forward_body<P, T, INTEG, false>(m, fw, 0, lane, Xc, Uc, 16, lane);
                a.fw.cost_out[(int64_t)lane * B + b] = ps[16 + lane];
                a.fw.diverged_out[(int64_t)lane * B + b] = pflag[16 + lane];
                if (a.fw.terminated_out) a.fw.terminated_out[(int64_t)lane * B + b] = pflag[32 + lane];
                if (a.fw.xN_out) {
#pragma unroll
                    for (int i = 0; i < N; i++) a.fw.xN_out[((int64_t)lane * N + i) * B + b] = Xc[((size_t)Ts * N + i) * 16 + lane];
                }
            }
            __syncwarp();
            int status = 0;
            if (lane == 0) status = select_one(a.sel, b);
            status = __shfl_sync(0xffffffffu, status, 0);
            if (a.sel.accept[b]) {  // self.xx, self.uu, self.KK, self.kk = the best candidate (ilqr.py:458-464)
                const T ab = a.sel.alpha_best[b];
                int best = 0;
                for (int i = 0; i < nA; i++)
                    if (T(a.sel.p.alphas[i]) == ab) best = i;
                for (int e = lane; e < nX; e += 32) pX[e] = Xc[(size_t)e * 16 + best];
                for (int e = lane; e < nU; e += 32) pU[e] = Uc[(size_t)e * 16 + best];
                {  // self.KK, self.kk = KK, kk: the candidate gains become the accepted ones by exchanging the two buffers
                    T* t1 = pKa;
                    pKa = pKc;
                    pKc = t1;
                    T* t2 = pka;
                    pka = pkc;
                    pkc = t2;
                    bw.KK = pKc;
                    bw.kk = pkc;
                    fw.KK = pKc;
                    fw.kk = pkc;
                    last_is_accepted = true;
                }
                any_accept = true;
            }
            __syncwarp();
            if (status != 0) break;
        }
        if (any_accept) {  // the result back into the batch arrays
            for (int e = lane; e < nX; e += 32) a.xbar[(int64_t)e * B + b] = pX[e];
            for (int e = lane; e < nU; e += 32) a.ubar[(int64_t)e * B + b] = pU[e];
            for (int e = lane; e < nK; e += 32) a.KK_acc[(int64_t)e * B + b] = pKa[e];
            for (int e = lane; e < nU; e += 32) a.kk_acc[(int64_t)e * B + b] = pka[e];
        }
        {  // the last candidate gains (as K3 leaves them): the accepted buffer if the last iteration was accepted
            const T* lK = last_is_accepted ? pKa : pKc;
            const T* lk = last_is_accepted ? pka : pkc;
            for (int e = lane; e < nK; e += 32) a.bw.KK[(int64_t)e * B + b] = lK[e];
            for (int e = lane; e < nU; e += 32) a.bw.kk[(int64_t)e * B + b] = lk[e];
        }
        if (running && lane == 0) a.index_out[atomicAdd(a.count_out, 1)] = (int32_t)b;
        __syncwarp();
    }
}

// ---- cost of given candidate trajectories (forward passes integrated elsewhere: learned dynamics) -----------------
// cost[a, b] = sum_k (c(x_k, u_k, x_{k+1}, t_{k+1}) + terminal_cost * terminated_{k+1}) dt + fcost(x_T) dt, with the
// time accumulated like the environment does (fast_step, environments.py:313-322; ilqr.py:213-224).
template <typename T>
struct TrajCostArgs {
    int64_t B;
    int T_steps, n_alpha;
    T dt, t0, terminal_cost;
    const T* X;  // [n_alpha, T+1, n, B]
    const T* U;  // [n_alpha, T, m, B]
    T* cost_out;
    uint8_t* diverged_out;
    uint8_t* terminated_out;
    const int32_t* index;
    const int32_t* count;
};

template <typename P, typename T>
__global__ void traj_cost_kernel(const TrajCostArgs<T> a) {
    const typename P::template Ctx<T> m;
    constexpr int N = P::N, M = P::M;
    const int64_t B = a.B;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int ia = blockIdx.y;
    int64_t b = tid;
    if (a.index) {
        if (tid >= *a.count) return;
        b = a.index[tid];
    } else if (b >= B) {
        return;
    }
    const T* X = a.X + (int64_t)ia * (a.T_steps + 1) * N * B;
    const T* U = a.U + (int64_t)ia * a.T_steps * M * B;
    T x[N], xn[N], u[M];
    bool div = false, term = false;
#pragma unroll
    for (int i = 0; i < N; i++) {
        xn[i] = X[i * B + b];
        div |= (xn[i] > T(1e8));
    }
    T cost = T(0), t = a.t0;
    for (int k = 0; k < a.T_steps; k++) {
        copy(x, xn);
#pragma unroll
        for (int i = 0; i < N; i++) xn[i] = X[((int64_t)(k + 1) * N + i) * B + b];
#pragma unroll
        for (int r = 0; r < M; r++) u[r] = U[((int64_t)k * M + r) * B + b];
        t += a.dt;
        term = P::terminate(m, xn);
        cost += (P::cost(m, x, u, xn, t) + (term ? a.terminal_cost : T(0))) * a.dt;
#pragma unroll
        for (int i = 0; i < N; i++) div |= (xn[i] > T(1e8));
    }
    cost += P::fcost(m, xn) * a.dt;
    const int64_t o = (int64_t)ia * B + b;
    a.cost_out[o] = cost;
    if (a.diverged_out) a.diverged_out[o] = div ? 1 : 0;
    if (a.terminated_out) a.terminated_out[o] = term ? 1 : 0;
}

// ---- commit by copy: the winning candidate becomes the nominal trajectory (ilqr.py:458-464) -------------------------
template <typename T>
struct CommitCopyArgs {
    int64_t B;
    int T_steps, n_alpha, N, M;
    T alphas[MAX_ALPHAS];
    const T* alpha_best;
    const uint8_t* accept;
    const T* X_cand;
    const T* U_cand;
    T* xbar;
    T* ubar;
    const T* KK;
    const T* kk;
    T* KK_acc;
    T* kk_acc;
    const int32_t* index;
    const int32_t* count;
};

// blockDim = (32 trajectories, COPY_ROWS row lanes): the ~(T+1) n + T m (n + 2) rows of a trajectory are spread over the
// row lanes so that a launch with a few hundred live trajectories still has thousands of loads in flight (a one-thread-
// per-trajectory loop is bound by ~1,000 dependent-issue memory round trips).
constexpr int COPY_ROWS = 16;

template <typename T>
__global__ void commit_copy_kernel(const CommitCopyArgs<T> a) {
    const int64_t B = a.B;
    const int64_t tid = (int64_t)blockIdx.x * 32 + threadIdx.x;
    int64_t b = tid;
    if (a.index) {
        if (tid >= *a.count) return;
        b = a.index[tid];
    } else if (b >= B) {
        return;
    }
    if (!a.accept[b]) return;
    const T ab = a.alpha_best[b];
    int ia = 0;
    for (int i = 0; i < a.n_alpha; i++)
        if (a.alphas[i] == ab) ia = i;
    const int N = a.N, M = a.M, Ts = a.T_steps;
    const T* X = a.X_cand + (int64_t)ia * (Ts + 1) * N * B;
    const T* U = a.U_cand + (int64_t)ia * Ts * M * B;
    const int64_t nx = (int64_t)(Ts + 1) * N, nu = (int64_t)Ts * M, nK = (int64_t)Ts * M * N;
    const int ry = threadIdx.y;
#pragma unroll 4
    for (int64_t r = ry; r < nx; r += COPY_ROWS) a.xbar[r * B + b] = X[r * B + b];
#pragma unroll 4
    for (int64_t r = ry; r < nu; r += COPY_ROWS) {
        a.ubar[r * B + b] = U[r * B + b];
        a.kk_acc[r * B + b] = a.kk[r * B + b];
    }
#pragma unroll 4
    for (int64_t r = ry; r < nK; r += COPY_ROWS) a.KK_acc[r * B + b] = a.KK[r * B + b];
}

// ---- receding-horizon control (MPCAgent, pygent/algorithms/nmpc.py:187-277) -------------------------------------------
// take_action (:187-211): u = K_0 (x - xbar_0) + ubar_0 + alpha k_0 (+ noise), clipped to +-uMax.
template <typename T>
struct MpcActionArgs {
    int64_t B;
    int N, M;
    const T* KK;
    const T* kk;
    const T* xbar;
    const T* ubar;
    const T* alpha;
    const T* x;
    const T* noise;
    T ulim[2];
    T* u_out;
};

template <typename T>
__global__ void mpc_action_kernel(const MpcActionArgs<T> a) {
    const int64_t B = a.B;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    for (int r = 0; r < a.M; r++) {
        T acc = T(0);
        for (int i = 0; i < a.N; i++) acc += a.KK[((int64_t)r * a.N + i) * B + b] * (a.x[(int64_t)i * B + b] - a.xbar[(int64_t)i * B + b]);
        acc = acc + a.ubar[(int64_t)r * B + b];
        acc = acc + a.alpha[b] * a.kk[(int64_t)r * B + b];
        if (a.noise) acc = acc + a.noise[(int64_t)r * B + b];
        a.u_out[(int64_t)r * B + b] = pg::clamp(acc, -a.ulim[r], a.ulim[r]);
    }
}

// shift_planner (:257-277): drop the first interval of the plan, append one interval with u = 0 integrated from the
// state the planner's environment was left in (the end of the last line-search candidate evaluated -- accepted or not),
// keep the cost consistent; the gains are NOT shifted, only their last entry is zeroed (as the reference does).
template <typename T>
struct MpcShiftArgs {
    int64_t B;
    int T_steps, S, n_alpha;
    T dt, t_env, terminal_cost;
    T* xbar;
    T* ubar;
    T* KK;
    T* kk;
    T* cost;
    T* xe;                      // [n, B] planner-environment state: in = fallback start, out = appended terminal state
    const T* xN_cand;           // [n_alpha, n, B] or NULL
    const int32_t* last_tried;  // [B] or NULL
    const T* x_new_ext;         // [n, B] or NULL: appended state integrated by the caller (learned dynamics)
};

template <typename P, typename T, int INTEG>
__global__ void mpc_shift_kernel(const MpcShiftArgs<T> a) {
    const typename P::template Ctx<T> m;
    constexpr int N = P::N, M = P::M;
    const int64_t B = a.B;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int Ts = a.T_steps;
    T x0[N], x1[N], xN[N], u0[M], xs[N], u[M];
#pragma unroll
    for (int i = 0; i < N; i++) {
        x0[i] = a.xbar[(int64_t)i * B + b];
        x1[i] = a.xbar[((int64_t)N + i) * B + b];
        xN[i] = a.xbar[((int64_t)Ts * N + i) * B + b];
    }
#pragma unroll
    for (int r = 0; r < M; r++) {
        u0[r] = a.ubar[(int64_t)r * B + b];
        u[r] = T(0);
    }
    // cost -= c(x0, u0, x1, t = None) dt + fcost(xN) dt   (the example costs do not use t: evaluated at 0)
    T cost = a.cost[b] - (P::cost(m, x0, u0, x1, T(0)) * a.dt + P::fcost(m, xN) * a.dt);
    // np.roll(xx, -1), np.roll(uu, -1), in place.  Eight intervals are loaded before they are stored one slot down: loads and
    // stores may alias, so an interval-by-interval loop is a chain of T dependent L2 round trips (116 us for T = 50: the
    // longest kernel of a control step); the blocked loop is T / 8 of them.
    constexpr int SB = 8;
    for (int k0 = 0; k0 < Ts; k0 += SB) {
        T xs8[SB][N], us8[SB][M];
#pragma unroll
        for (int s = 0; s < SB; s++) {
            const int k = k0 + s;
            if (k < Ts) {
#pragma unroll
                for (int i = 0; i < N; i++) xs8[s][i] = a.xbar[((int64_t)(k + 1) * N + i) * B + b];
                if (k + 1 < Ts) {
#pragma unroll
                    for (int r = 0; r < M; r++) us8[s][r] = a.ubar[((int64_t)(k + 1) * M + r) * B + b];
                }
            }
        }
#pragma unroll
        for (int s = 0; s < SB; s++) {
            const int k = k0 + s;
            if (k < Ts) {
#pragma unroll
                for (int i = 0; i < N; i++) a.xbar[((int64_t)k * N + i) * B + b] = xs8[s][i];
                if (k + 1 < Ts) {
#pragma unroll
                    for (int r = 0; r < M; r++) a.ubar[((int64_t)k * M + r) * B + b] = us8[s][r];
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < M; r++) {
        a.ubar[((int64_t)(Ts - 1) * M + r) * B + b] = T(0);
        a.kk[((int64_t)(Ts - 1) * M + r) * B + b] = T(0);
#pragma unroll
        for (int i = 0; i < N; i++) a.KK[(((int64_t)(Ts - 1) * M + r) * N + i) * B + b] = T(0);
    }
    const int lt = a.last_tried ? a.last_tried[b] : -1;
#pragma unroll
    for (int i = 0; i < N; i++)
        xs[i] = (lt >= 0 && a.xN_cand) ? a.xN_cand[((int64_t)lt * N + i) * B + b] : a.xe[(int64_t)i * B + b];
    T xn[N];
    copy(xn, xs);
    if (a.x_new_ext) {
#pragma unroll
        for (int i = 0; i < N; i++) xn[i] = a.x_new_ext[(int64_t)i * B + b];
    } else {
        env_step<P, T, INTEG>(m, xn, u, a.dt, a.dt / T(a.S), a.S);  // environment.step(u = 0)
    }
    const bool term = P::terminate(m, xn);
    cost += (P::cost(m, xs, u, xn, a.t_env + a.dt) + (term ? a.terminal_cost : T(0))) * a.dt + P::fcost(m, xn) * a.dt;
#pragma unroll
    for (int i = 0; i < N; i++) {
        a.xbar[((int64_t)Ts * N + i) * B + b] = xn[i];
        a.xe[(int64_t)i * B + b] = xn[i];
    }
    a.cost[b] = cost;
}

// ---- line-search work lists (serial search in two phases) ---------------------------------------------------------
// The serial line search stops at the first alpha with z > zmin (ilqr.py:453-454; ~4.8 of 11 alphas on average), so the
// candidates are evaluated in two launches: the first `n_first` alphas for every trajectory whose backward pass succeeded
// (phase 0 builds that list), the remaining ones only for trajectories that have neither succeeded nor diverged yet
// (phase 1 builds that list from the first costs, with the z-test of ilqr_select_kernel).  K4 reads only the alphas the
// serial search tries, so its decisions are unchanged.
template <typename T>
struct FilterArgs {
    int64_t B;
    int phase, n_first;
    double alphas[MAX_ALPHAS];
    double zmin;
    const T* cost;
    const T* dV;
    const uint8_t* bw_success;
    const T* cost_cand;
    const uint8_t* diverged;
    const int32_t* index_in;
    const int32_t* count_in;
    int32_t* index_out;
    int32_t* count_out;
    int small_count;  // phase 1: nothing to do when *count_in <= small_count (launch 1 took every alpha)
};

template <typename T>
__global__ void ilqr_linesearch_filter_kernel(const FilterArgs<T> a) {
    const int64_t B = a.B;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t b = tid;
    bool valid = true;
    if (a.index_in) {
        valid = tid < *a.count_in;
        b = valid ? a.index_in[tid] : 0;
    } else {
        valid = b < B;
    }
    bool keep = false;
    if (valid) {
        if (a.phase == 0) {
            keep = a.bw_success[b] != 0;
        } else if (a.count_in && *a.count_in <= a.small_count) {
            keep = false;
        } else {
            keep = true;
            const T cost_old = a.cost[b], sV1 = a.dV[b], sV2 = a.dV[B + b];
            for (int i = 0; i < a.n_first; i++) {
                if (a.diverged[(int64_t)i * B + b]) {  // 'forward diverged.': the search ends here
                    keep = false;
                    break;
                }
                const T alpha = T(a.alphas[i]);
                const T dcost = cost_old - a.cost_cand[(int64_t)i * B + b];
                const T expected = -alpha * (sV1 + alpha * sV2);
                const T z = expected > T(0) ? dcost / expected : pg::sign(dcost);
                if (z > T(a.zmin)) {
                    keep = false;
                    break;
                }
            }
        }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    if (ballot) {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == __ffs(ballot) - 1) base = atomicAdd(a.count_out, __popc(ballot));
        base = __shfl_sync(0xffffffffu, base, __ffs(ballot) - 1);
        if (keep) a.index_out[base + __popc(ballot & ((1u << lane) - 1))] = (int32_t)b;
    }
}

// ---- observation (helpers.observation, pygent/helpers.py:20-29) -----------------------------------
template <typename P, typename T>
__global__ void observe_kernel(int64_t B, const T* x, T* o) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int c = 0;
#pragma unroll
    for (int i = 0; i < P::N; i++) {
        const T v = x[i * B + b];
        if (P::x_is_angle(i)) {
            T s, co;
            pg::MathCtx<T>().sincos(v, s, co);
            o[(c++) * B + b] = co;
            o[(c++) * B + b] = s;
        } else {
            o[(c++) * B + b] = v;
        }
    }
}

// ---- batched point evaluation (parity tests) -------------------------------------------------------
template <typename T>
struct EvalArgs {
    int64_t B;
    int lin_mode;
    const T* x;
    const T* u;
    const T* t;
    T *f, *A, *Bm, *c, *cx, *cu, *Cxx, *Cuu, *Cxu, *cf, *cfx, *Cfxx;
};

template <typename P, typename T>
__global__ void eval_kernel(const EvalArgs<T> a) {
    const typename P::template Ctx<T> m;
    constexpr int N = P::N, M = P::M;
    const int64_t B = a.B;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    T x[N], u[M];
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = a.x[i * B + b];
#pragma unroll
    for (int r = 0; r < M; r++) u[r] = a.u[r * B + b];
    const T t = a.t ? a.t[b] : T(0);
    if (a.f) {
        T dx[N];
        P::f(m, x, u, dx);
#pragma unroll
        for (int i = 0; i < N; i++) a.f[i * B + b] = dx[i];
    }
    if (a.A || a.Bm) {
        T A[N * N], Bm[N * M];
        if ((a.lin_mode & 3) == 0) P::AB(m, x, u, A, Bm);
        else fd_AB<P, T>(m, x, u, A, Bm);
        if (a.A) {
#pragma unroll
            for (int i = 0; i < N * N; i++) a.A[i * B + b] = A[i];
        }
        if (a.Bm) {
#pragma unroll
            for (int i = 0; i < N * M; i++) a.Bm[i * B + b] = Bm[i];
        }
    }
    if (a.c) a.c[b] = P::cost(m, x, u, x, t);
    if (a.cx || a.cu || a.Cxx || a.Cuu || a.Cxu) {
        T cx[N], cu[M], Cxx[N * N], Cuu[M * M], Cxu[N * M];
        if (a.lin_mode & 4) fd_cost_derivs<P, T>(m, x, u, t, T(1), cx, cu, Cxx, Cuu, Cxu);  // finite_diff=True
        else P::cost_derivs(m, x, u, t, cx, cu, Cxx, Cuu, Cxu);
        if (a.cx) {
#pragma unroll
            for (int i = 0; i < N; i++) a.cx[i * B + b] = cx[i];
        }
        if (a.cu) {
#pragma unroll
            for (int i = 0; i < M; i++) a.cu[i * B + b] = cu[i];
        }
        if (a.Cxx) {
#pragma unroll
            for (int i = 0; i < N * N; i++) a.Cxx[i * B + b] = Cxx[i];
        }
        if (a.Cuu) {
#pragma unroll
            for (int i = 0; i < M * M; i++) a.Cuu[i * B + b] = Cuu[i];
        }
        if (a.Cxu) {
#pragma unroll
            for (int i = 0; i < N * M; i++) a.Cxu[i * B + b] = Cxu[i];
        }
    }
    if (a.cf) a.cf[b] = P::fcost(m, x);
    if (a.cfx || a.Cfxx) {
        T cfx[N], Cfxx[N * N];
        if (a.lin_mode & 4) fd_fcost_derivs<P, T>(m, x, cfx, Cfxx);
        else P::fcost_derivs(m, x, cfx, Cfxx);
        if (a.cfx) {
#pragma unroll
            for (int i = 0; i < N; i++) a.cfx[i * B + b] = cfx[i];
        }
        if (a.Cfxx) {
#pragma unroll
            for (int i = 0; i < N * N; i++) a.Cfxx[i * B + b] = Cfxx[i];
        }
    }
}

}  // namespace pgk
