// FP64 / FP32 FMA-pipe peak and latency microbenchmark (roofline denominator of the RK4 / iLQR kernels).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp_peak tools/fp_peak.cu ; ./tools/fp_peak > out.json
// Each thread runs ILP independent dependent-FMA chains; sweeping warps/SMSP and ILP gives both the pipe
// peak (many warps, ILP 8) and the dependent-issue latency (1 warp per SMSP, ILP 1).
#include <cstdio>
#include <cuda_runtime.h>

template <typename T, int ILP>
__global__ void fma_kernel(T* out, int iters, T a, T b) {
    T v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = (T)(threadIdx.x + i) * (T)1e-3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) v[i] = fma(v[i], a, b);
        }
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += v[i];
    if (s == (T)12345.678) out[0] = s;
}

// packed fp32: fma.rn.f32x2 = one FFMA2 instruction doing two FMAs per lane
template <int ILP>
__global__ void fma2_kernel(float* out, float a, float b, int iters) {
    unsigned long long v[ILP], aa, bb;
    asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a), "f"(a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b), "f"(b));
#pragma unroll
    for (int i = 0; i < ILP; i++) {
        float x = (float)(threadIdx.x + i) * 1e-3f;
        asm("mov.b64 %0, {%1, %2};" : "=l"(v[i]) : "f"(x), "f"(x + 1.0f));
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(aa), "l"(bb));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[i]));
        s += lo + hi;
    }
    if (s == 12345.678f) out[0] = s;
}

template <int ILP>
double run2(int sms, int warps_per_sm, int iters, double* ms_out) {
    float* out;
    cudaMalloc(&out, 64);
    const int threads = 128;
    const int blocks = sms * warps_per_sm * 32 / threads;
    fma2_kernel<ILP><<<blocks, threads>>>(out, 0.999f, 1e-6f, 10);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    fma2_kernel<ILP><<<blocks, threads>>>(out, 0.999f, 1e-6f, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    cudaFree(out);
    *ms_out = ms;
    const double flops = 4.0 * 16 * ILP * (double)iters * (double)blocks * threads;
    return flops / (ms * 1e-3) / 1e12;
}

template <typename T, int ILP>
double run(int sms, int warps_per_sm, int iters, T* out, double* ms_out) {
    const int threads = 32 * (warps_per_sm < 32 ? warps_per_sm : 32);
    const int ctas_per_sm = (32 * warps_per_sm) / threads;
    dim3 grid(sms * ctas_per_sm), block(threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    fma_kernel<T, ILP><<<grid, block>>>(out, iters / 8, (T)0.999, (T)1e-3);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        fma_kernel<T, ILP><<<grid, block>>>(out, iters, (T)0.999, (T)1e-3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    *ms_out = best;
    const double fmas = (double)grid.x * threads * ILP * 8.0 * iters;
    return 2.0 * fmas / (best * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    void* out;
    cudaMalloc(&out, 64);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"runs\": [\n", p.name, sms, clk_khz);
    const int wps[] = {4, 8, 16, 32, 64};
    bool first = true;
    double best64 = 0, best32 = 0;
    for (int wi = 0; wi < 5; wi++) {
        const int w = wps[wi];
        double ms, tf;
#define RUN(T, NAME, ILP, ITERS, BEST)                                                              \
        tf = run<T, ILP>(sms, w, ITERS, (T*)out, &ms);                                              \
        if (tf > BEST) BEST = tf;                                                                   \
        printf("%s {\"type\": \"%s\", \"warps_per_sm\": %d, \"ilp\": %d, \"tflops\": %.3f, \"ms\": %.3f}", first ? "" : ",\n", NAME, w, ILP, tf, ms); \
        first = false;
        RUN(double, "fp64", 1, 4000, best64)
        RUN(double, "fp64", 2, 4000, best64)
        RUN(double, "fp64", 4, 4000, best64)
        RUN(double, "fp64", 8, 2000, best64)
        RUN(float, "fp32", 1, 16000, best32)
        RUN(float, "fp32", 2, 16000, best32)
        RUN(float, "fp32", 4, 16000, best32)
        RUN(float, "fp32", 8, 8000, best32)
    }
    double best2 = 0, ms2;
    for (int w : {8, 16, 32}) {
        double t = run2<4>(sms, w, 2000, &ms2);
        best2 = t > best2 ? t : best2;
        t = run2<8>(sms, w, 2000, &ms2);
        best2 = t > best2 ? t : best2;
    }
    printf("\n], \"fp64_fma_tflops\": %.3f, \"fp32_fma_tflops\": %.3f, \"fp32_fma2_tflops\": %.3f}\n", best64, best32, best2);
    return 0;
}
