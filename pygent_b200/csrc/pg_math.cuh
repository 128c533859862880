// Scalar math used by generated problem code (namespace pg), overloaded for double and float.
// No fast-math intrinsics: fp32 results are meant to track the fp64 path as closely as fp32 allows.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pg {

template <int K, typename T>
__device__ __forceinline__ T powi(T x) {
    static_assert(K >= 1, "powi exponent");
    if constexpr (K == 1) {
        return x;
    } else if constexpr (K % 2 == 0) {
        const T h = powi<K / 2>(x);
        return h * h;
    } else {
        return x * powi<K - 1>(x);
    }
}

// A value the compiler must keep in a register.  ptxas re-materialises every double literal with two
// UMOVs / IMAD.MOVs (or an LDC from the constant bank) right in front of its uses -- ~25% extra
// instructions, and their latency, on the dependent chain of the RK4 loop -- and it sees through
// `mov` of an immediate.  A volatile load from GLOBAL memory cannot be re-issued, so the value stays
// in a register for the lifetime of the thread; the 8-byte loads happen once per thread at kernel
// entry and hit L1/L2.
__device__ __forceinline__ double opaque_load(const double* p) {
    double r;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}

// literals of generated problem code: registers for double (no 64-bit immediates exist), plain
// immediates for float (FFMA takes a 32-bit immediate)
__device__ __forceinline__ double hoist(const double* table, int i, double) { return opaque_load(table + i); }
__device__ __forceinline__ float hoist(const double*, int, float v) { return v; }

// sin/cos coefficients: S1..S6, C1..C6 (fdlibm minimax on |r| <= pi/4), 2/pi, -pi/2 in three parts
__device__ const double PG_SC_TABLE[16] = {
    -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
    2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,
    4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
    -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11,
    0.63661977236758138, -1.5707963267948966e+00, -6.1232339957367574e-17, -8.4784276603688995e-32};

// ---- sin/cos ------------------------------------------------------------------------------------
// MathCtx<T> is created once per thread at kernel entry and handed to the generated problem code.
//
// fp64: one shared Cody-Waite reduction by pi/2 (3 constants, exact products via FMA) feeding the
// two minimax polynomials in Estrin form.  Everything numeric stays on the FP64 FMA pipe: no
// F2I/I2F conversions (the quadrant is the low word of the magic-number sum), quadrant swap and
// sign flips are integer selects / XORs.  22 FP64 instructions per sin+cos pair with a dependent
// chain of 10 (Horner would be 20 instructions but a chain of 15; the rollout kernels run at 3-4
// warps per scheduler and are bound by that chain, not by the pipe).  Coefficients are held in
// registers (see opaque_load()).  Accuracy <= 1.6 ulp, absolute error < 1.8e-16, for |x| < 2^30
// (checked against long-double libm on 2e7 points); outside that range, and for non-finite
// input, the result is NaN.  The line search's divergence test is one-sided (x > 1e8, as in the reference), so a state that
// diverges towards -inf gets here: its cost is NaN, fails the z-test and is skipped by the argmin (select_one).
// There is deliberately no call to the CUDA library slow path: its ABI spill sat in the hot loop.
template <typename T>
struct MathCtx;

template <>
struct MathCtx<double> {
    double S1, S2, S3, S4, S5, S6, C1, C2, C3, C4, C5, C6, TWO_OVER_PI, P1, P2, P3;
    __device__ __forceinline__ MathCtx() {
        const double* t = PG_SC_TABLE;
        S1 = opaque_load(t + 0);
        S2 = opaque_load(t + 1);
        S3 = opaque_load(t + 2);
        S4 = opaque_load(t + 3);
        S5 = opaque_load(t + 4);
        S6 = opaque_load(t + 5);
        C1 = opaque_load(t + 6);
        C2 = opaque_load(t + 7);
        C3 = opaque_load(t + 8);
        C4 = opaque_load(t + 9);
        C5 = opaque_load(t + 10);
        C6 = opaque_load(t + 11);
        TWO_OVER_PI = opaque_load(t + 12);
        P1 = opaque_load(t + 13);
        P2 = opaque_load(t + 14);
        P3 = opaque_load(t + 15);
    }
    __device__ __forceinline__ void sincos(double x, double& s, double& c) const {
        {  // out-of-domain (|x| >= 2^30, inf, nan) -> NaN, decided on the exponent bits (integer pipe)
            const int hi = __double2hiint(x);
            x = __hiloint2double(((hi & 0x7fffffff) < 0x41d00000) ? hi : 0x7ff80000, __double2loint(x));
        }
        const double magic = 6755399441055744.0;  // 1.5 * 2^52
        const double qd = fma(x, TWO_OVER_PI, magic);
        const int q = __double2loint(qd);
        const double j = qd - magic;
        double r = fma(j, P1, x);
        r = fma(j, P2, r);
        r = fma(j, P3, r);
        const double z = r * r;
        const double z2 = z * z, r3 = z * r;
        // sin r = r + r^3 ((S1 + S2 z) + z^2 ((S3 + S4 z) + z^2 (S5 + S6 z)))
        const double sa = fma(S2, z, S1), sb = fma(S4, z, S3), sc = fma(S6, z, S5);
        const double sr = fma(fma(fma(sc, z2, sb), z2, sa), r3, r);
        // cos r = (1 - z/2) + z^2 ((C1 + C2 z) + z^2 ((C3 + C4 z) + z^2 (C5 + C6 z)))
        const double c0 = fma(-0.5, z, 1.0), ca = fma(C2, z, C1), cb = fma(C4, z, C3), cc = fma(C6, z, C5);
        const double cr = fma(fma(fma(cc, z2, cb), z2, ca), z2, c0);
        const double a = (q & 1) ? cr : sr;  // sin of the full angle, up to sign
        const double b = (q & 1) ? sr : cr;  // cos of the full angle, up to sign
        s = __hiloint2double(__double2hiint(a) ^ ((q & 2) << 30), __double2loint(a));
        c = __hiloint2double(__double2hiint(b) ^ (((q + 1) & 2) << 30), __double2loint(b));
    }
    // sin/cos(a0 + d) from the "anchor" (s0, c0) = sin/cos(a0) by the angle-addition formulas, for the small
    // argument increments of the RK4 stages (|d| = h |thetadot|, a few 1e-2): 13 FP64 instructions and no integer
    // work instead of 22 + ~12.  sin d and cos d - 1 are the fdlibm polynomials cut after z^3 (their next terms
    // are < 4e-17 resp. 3e-19 for |d| < 2^-4), so the result is within ~1 ulp of a fresh evaluation plus the
    // anchor's own error.  Valid for |d| < 2^-4 only: callers collect mag(d) (integer pipe) and redo the control
    // interval with fresh sincos() when max mag >= ROT_LIMIT (also catches inf/nan); they re-anchor with
    // sincos() once per control interval.
    static constexpr int ROT_LIMIT = 0x3fb00000;  // high word of 2^-4
    __device__ __forceinline__ static int mag(double d) { return __double2hiint(d) & 0x7fffffff; }
    __device__ __forceinline__ void rotate(double d, double s0, double c0, double& s, double& c) const {
        const double z = d * d;
        const double sd = fma(fma(fma(S3, z, S2), z, S1), d * z, d);    // sin d
        const double cm = fma(fma(fma(C3, z, C2), z, C1), z, -0.5) * z;  // cos d - 1
        s = fma(c0, sd, fma(s0, cm, s0));
        c = fma(-s0, sd, fma(c0, cm, c0));
    }
    __device__ __forceinline__ double sin(double x) const {
        double s, c;
        sincos(x, s, c);
        return s;
    }
    __device__ __forceinline__ double cos(double x) const {
        double s, c;
        sincos(x, s, c);
        return c;
    }
};

template <>
struct MathCtx<float> {
    __device__ __forceinline__ void sincos(float x, float& s, float& c) const { ::sincosf(x, &s, &c); }
    // see MathCtx<double>::rotate; Taylor terms through d^5 / d^6 (next: 8e-13, 6e-15 for |d| < 2^-4)
    static constexpr int ROT_LIMIT = 0x3d800000;  // 2^-4f
    __device__ __forceinline__ static int mag(float d) { return __float_as_int(d) & 0x7fffffff; }
    __device__ __forceinline__ void rotate(float d, float s0, float c0, float& s, float& c) const {
        const float z = d * d;
        const float sd = fmaf(fmaf(8.3333333333e-3f, z, -1.6666666667e-1f), d * z, d);
        const float cm = fmaf(fmaf(-1.3888888889e-3f, z, 4.1666666667e-2f), z, -0.5f) * z;
        s = fmaf(c0, sd, fmaf(s0, cm, s0));
        c = fmaf(-s0, sd, fmaf(c0, cm, c0));
    }
    __device__ __forceinline__ float sin(float x) const { return ::sinf(x); }
    __device__ __forceinline__ float cos(float x) const { return ::cosf(x); }
};

// free-function forms (observation kernel, hand-written code)
__device__ __forceinline__ void sincos(double x, double& s, double& c) { MathCtx<double>().sincos(x, s, c); }
__device__ __forceinline__ void sincos(float x, float& s, float& c) { ::sincosf(x, &s, &c); }

#define PG_UNARY(name, fd, ff)                                   \
    __device__ __forceinline__ double name(double x) { return fd(x); } \
    __device__ __forceinline__ float name(float x) { return ff(x); }
PG_UNARY(tan, ::tan, ::tanf)
PG_UNARY(exp, ::exp, ::expf)
PG_UNARY(log, ::log, ::logf)
PG_UNARY(sqrt, ::sqrt, ::sqrtf)
PG_UNARY(asin, ::asin, ::asinf)
PG_UNARY(acos, ::acos, ::acosf)
PG_UNARY(atan, ::atan, ::atanf)
PG_UNARY(abs, ::fabs, ::fabsf)
PG_UNARY(tanh, ::tanh, ::tanhf)
PG_UNARY(sinh, ::sinh, ::sinhf)
PG_UNARY(cosh, ::cosh, ::coshf)
#undef PG_UNARY

__device__ __forceinline__ double pow(double x, double y) { return ::pow(x, y); }
__device__ __forceinline__ float pow(float x, float y) { return ::powf(x, y); }
__device__ __forceinline__ double atan2(double y, double x) { return ::atan2(y, x); }
__device__ __forceinline__ float atan2(float y, float x) { return ::atan2f(y, x); }
template <typename T>
__device__ __forceinline__ T sign(T x) {
    return x > T(0) ? T(1) : (x < T(0) ? T(-1) : x);  // np.sign: 0 -> 0, nan -> nan
}
template <typename T>
__device__ __forceinline__ T clamp(T v, T lo, T hi) {
    return fmin(fmax(v, lo), hi);  // np.clip
}

}  // namespace pg
