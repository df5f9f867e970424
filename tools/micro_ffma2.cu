// micro_ffma2.cu -- FP32 FMA issue patterns on B200: scalar FFMA vs packed FFMA2 (fma.rn.f32x2)
// with and without operand reuse.  Decides how the inlier sweep (k_score) should be written.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_ffma2 tools/micro_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }

// MODE 0: scalar FFMA, acc_i = x_j * y_k + acc_i with rotating distinct x, y (12 accumulators)
// MODE 1: FFMA2, all three operands distinct register pairs, no reuse between neighbours
// MODE 2: FFMA2, three consecutive instructions share the B operand (like R0?*px for the 3 rows)
// MODE 3: FFMA2, acc = acc * a + b (two constant operands: the probe in lr_prof.cu)
template <int MODE>
__global__ void __launch_bounds__(128) k(float *out, int iters, float seed)
{
    float t = threadIdx.x * 1e-3f + seed;
    if (MODE == 0) {
        float x[4], y[4], a[12];
        for (int i = 0; i < 4; ++i) { x[i] = t + i; y[i] = t * 2 + i; }
        for (int i = 0; i < 12; ++i) a[i] = i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 12; ++i) a[i] = fmaf(x[(i + u) & 3], y[(i * 3 + u) & 3], a[i]);
        }
        float s = 0; for (int i = 0; i < 12; ++i) s += a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        u64 x[4], y[4], a[12];
        for (int i = 0; i < 4; ++i) { x[i] = pk(t + i, t - i); y[i] = pk(t * 2 + i, t * 3 - i); }
        for (int i = 0; i < 12; ++i) a[i] = pk((float)i, (float)-i);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (MODE == 1) {
#pragma unroll
                    for (int i = 0; i < 12; ++i) a[i] = fma2(x[(i + u) & 3], y[(i * 3 + u) & 3], a[i]);
                } else if (MODE == 2) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {  // 3 rows share the "point" operand y[g]
                        a[3 * g + 0] = fma2(x[(g + u) & 3], y[g], a[3 * g + 0]);
                        a[3 * g + 1] = fma2(x[(g + u + 1) & 3], y[g], a[3 * g + 1]);
                        a[3 * g + 2] = fma2(x[(g + u + 2) & 3], y[g], a[3 * g + 2]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 12; ++i) a[i] = fma2(a[i], x[0], y[0]);
                }
            }
        }
        float s = 0;
        for (int i = 0; i < 12; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i])); s += lo + hi; }
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}

template <int MODE> void run(const char *name, int sms, float *out)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int bps = 2; bps <= 8; bps *= 2) {
        const int iters = 20000;
        k<MODE><<<sms * bps, 128>>>(out, 100, 1.f);
        cudaEventRecord(a);
        k<MODE><<<sms * bps, 128>>>(out, iters, 1.f);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        double fmas = (double)sms * bps * 128 * iters * 48.0 * (MODE == 0 ? 1 : 2);
        printf("%-44s %d CTA/SM (%2d warps): %6.1f TFLOP/s  = %5.1f FMA/clk/SM @1965MHz\n", name, bps, bps * 4, 2 * fmas / (ms * 1e-3) / 1e12,
               fmas / (ms * 1e-3) / sms / 1.965e9);
    }
}
int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out; cudaMalloc(&out, sizeof(float) * sms * 8 * 128);
    run<0>("FFMA  distinct operands", sms, out);
    run<1>("FFMA2 three distinct operand pairs", sms, out);
    run<2>("FFMA2 B operand shared by 3 neighbours", sms, out);
    run<3>("FFMA2 acc*a+b (two constant operands)", sms, out);
    return 0;
}
