// micro_b200.cu -- two B200 micro-benchmarks that decided the matching design:
//   (1) legacy mma.sync.m16n8k16 f16->f32 throughput (MAC/clk/SM)
//   (2) tcgen05.ld TMEM->register bandwidth for several shapes
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_b200 tools/micro_b200.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_hmma(float *out, int iters)
{
    unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD_X32(suffix)                                                                                               \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32" suffix ".b32 "                                                \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"  \
                 "%26,%27,%28,%29,%30,%31}, [%32];"                                                                  \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),  \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),         \
                   "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),       \
                   "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),       \
                   "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                             \
                 : "r"(taddr)                                                                                        \
                 : "memory")

// mode 0: 32x32b.x32 (32 columns, 4 B each); mode 1: 32x32b.x32.pack::16b (64 columns of 16-bit data);
// mode 2: 16x256b.x8 (each warp: 16 lanes x 8 x 256 bit)
template <int MODE>
__global__ void k_ldtm(unsigned long long *cycles, unsigned *sink, int iters, int inflight)
{
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t r[32];
    unsigned acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int k = 0; k < inflight; ++k) {
            const uint32_t taddr = base + ((it * inflight + k) % 8) * 32;
            if (MODE == 0) { LD_X32(""); }
            if (MODE == 1) { LD_X32(".pack::16b"); }
            if (MODE == 2) {
                asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 "
                             "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
                             "%26,%27,%28,%29,%30,%31}, [%32];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                               "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
                               "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
                               "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
                               "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                             : "r"(tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + ((it * inflight + k) % 4) * 64)
                             : "memory");
            }
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += r[it & 31];
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512u) : "memory");
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    float *out;
    cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int bps = 1; bps <= 4; bps *= 2) {
        const int iters = 20000;
        k_hmma<<<sms * bps, 256>>>(out, 100);
        cudaEventRecord(a);
        k_hmma<<<sms * bps, 256>>>(out, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        double macs = (double)sms * bps * 8 /*warps*/ * iters * 8.0 * 2048.0;
        printf("mma.sync m16n8k16 f16: %d CTA/SM x 8 warps: %.1f TFLOP/s dense, %.0f MAC/clk/SM at %.0f MHz nominal\n", bps,
               2 * macs / (ms * 1e-3) / 1e12, macs / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1e3);
    }
    unsigned long long *cyc;
    unsigned *sink;
    cudaMalloc(&cyc, sizeof(unsigned long long) * sms);
    cudaMalloc(&sink, sizeof(unsigned) * sms * 512);
    unsigned long long h[256];
    for (int mode = 0; mode < 3; ++mode)
        for (int warps = 4; warps <= 16; warps *= 2)
            for (int inflight = 1; inflight <= 4; inflight *= 2) {
                const int iters = 2000;
                if (mode == 0) k_ldtm<0><<<sms, warps * 32>>>(cyc, sink, iters, inflight);
                if (mode == 1) k_ldtm<1><<<sms, warps * 32>>>(cyc, sink, iters, inflight);
                if (mode == 2) k_ldtm<2><<<sms, warps * 32>>>(cyc, sink, iters, inflight);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
                double per = (double)h[0] / iters / inflight;
                printf("tcgen05.ld mode %d (%s) warps %2d inflight %d: %.1f cycles per warp-load, %.1f B/clk/SM (32-bit regs) [%s]\n", mode,
                       mode == 0 ? "32x32b.x32" : (mode == 1 ? "32x32b.x32.pack16" : "16x256b.x8"), warps, inflight, per,
                       warps * 4096.0 / per, cudaGetErrorString(e));
            }
    return 0;
}
