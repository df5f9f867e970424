// micro_const.cu -- does an FMA whose "point" operand comes from constant memory (uniform across the warp)
// escape the register-bandwidth limit measured by micro_ffma2.cu?  Mimics the inlier sweep: thread-owned
// [R|t] (12 registers), points streamed from __constant__ with a loop-counter index.
#include <cstdio>
#include <cuda_runtime.h>
__constant__ float4 cP[2048];   // (px, py, pz, qx)
__constant__ float2 cQ[2048];   // (qy, qz)

template <int MODE>
__global__ void __launch_bounds__(128) k(int *out, int iters, int npts, const float *__restrict__ models, float thr)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    float r[12];
    for (int i = 0; i < 12; ++i) r[i] = models[t * 12 + i];
    int cnt = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 8
        for (int i = 0; i < npts; ++i) {
            const float4 a = cP[i];
            const float2 b = cQ[i];
            float d0 = fmaf(r[0], a.x, fmaf(r[1], a.y, fmaf(r[2], a.z, r[3]))) - a.w;
            float d1 = fmaf(r[4], a.x, fmaf(r[5], a.y, fmaf(r[6], a.z, r[7]))) - b.x;
            float d2 = fmaf(r[8], a.x, fmaf(r[9], a.y, fmaf(r[10], a.z, r[11]))) - b.y;
            float rr = fmaf(d2, d2, fmaf(d1, d1, d0 * d0));
            cnt += rr < thr ? 1 : 0;
        }
    }
    out[t] = cnt;
}
int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = sms * 8 * 128;
    float *models; int *out;
    cudaMalloc(&models, sizeof(float) * 12 * threads); cudaMemset(models, 0, sizeof(float) * 12 * threads);
    cudaMalloc(&out, sizeof(int) * threads);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int bps = 4; bps <= 8; bps *= 2) {
        const int iters = 200, npts = 2048;
        k<0><<<sms * bps, 128>>>(out, 2, npts, models, 0.36f);
        cudaEventRecord(a);
        k<0><<<sms * bps, 128>>>(out, iters, npts, models, 0.36f);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        double evals = (double)sms * bps * 128 * iters * npts;
        printf("constant-operand sweep, %d CTA/SM: %.3e evals/s = %.1f FMA-equiv/clk/SM (15 per eval), %.1f cycles per eval-warp per SMSP\n", bps,
               evals / (ms * 1e-3), evals * 15 / (ms * 1e-3) / sms / 1.965e9, 1.965e9 * 4 * sms / (evals / 32 / (ms * 1e-3)));
    }
    return 0;
}
