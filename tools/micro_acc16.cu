// micro_acc16.cu -- B200 experiments behind the fp16-accumulator matching epilogue:
//   (1) one 128 x 256 x 48 tcgen05.mma tile with fp32 and with fp16 accumulators on the operand image the
//       matching sweep uses; dumps both through tcgen05.ld (plain and .pack::16b) and reports the error of
//       the fp16 accumulators against the exact value, in fp16 ulps, plus which half of a packed register
//       holds the lower column
//   (2) issue rate of the max instructions an epilogue can use: FMNMX3 (fp32), VHMNMX (3-input f16x2),
//       VIMNMX3.S16x2
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_acc16 tools/micro_acc16.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int TM = 128, TN = 256, RG = 1024;
constexpr uint32_t IDESC32 = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
constexpr uint32_t IDESC16 = IDESC32 & ~(1u << 4);

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr)
{
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(RG >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u)
                 : "memory");
}
#define REGS32(r)                                                                                                    \
    "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),      \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),       \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),      \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
#define REGS16(r)                                                                                                    \
    "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),      \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
#define LIST32 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
#define LIST16 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"

__global__ void __launch_bounds__(128) k_tile(const uint4 *Aimg, const uint4 *Bimg, uint32_t *D32, uint32_t *D16raw,
                                              uint32_t *D16pk)
{
    extern __shared__ uint8_t dyn[];
    uint8_t *base = dyn + ((1024u - (smem_u32(dyn) & 1023u)) & 1023u);
    uint4 *sA = reinterpret_cast<uint4 *>(base);
    uint4 *sB = reinterpret_cast<uint4 *>(base + 16384);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    for (int i = threadIdx.x; i < 16384 / 16; i += 128) sA[i] = Aimg[i];
    for (int i = threadIdx.x; i < 32768 / 16; i += 128) sB[i] = Bimg[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_s;
    if (threadIdx.x == 0) {
        const uint32_t a = smem_u32(sA), b = smem_u32(sB);
        for (int k = 0; k < 3; ++k) mma(tm, smem_desc(a + k * 256), smem_desc(b + k * 256), IDESC32, k > 0);
        for (int k = 0; k < 3; ++k) mma(tm + 256, smem_desc(a + k * 256), smem_desc(b + k * 256), IDESC16, k > 0);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tq = tm + ((uint32_t)(warp * 32) << 16);
    const int row = warp * 32 + lane;
    for (int c = 0; c < 8; ++c) {
        uint32_t r[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " LIST32 : REGS32(r) : "r"(tq + c * 32) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 32; ++i) D32[row * 256 + c * 32 + i] = r[i];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " LIST32 : REGS32(r) : "r"(tq + 256 + c * 32) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 32; ++i) D16raw[row * 256 + c * 32 + i] = r[i];
        uint32_t h[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 " LIST16 : REGS16(h) : "r"(tq + 256 + c * 32) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) D16pk[row * 128 + c * 16 + i] = h[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

// ---- max-instruction issue rates
template <int MODE>
__global__ void __launch_bounds__(512) k_maxrate(uint32_t *out, int iters)
{
    uint32_t a[8], x = threadIdx.x * 2654435761u, y = x ^ 0x3c003c00u;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("max.f32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(x), "r"(y));
            if (MODE == 1) asm volatile("{\n\t.reg .b32 t;\n\tmax.f16x2 t, %0, %1;\n\tmax.f16x2 %0, t, %2;\n\t}" : "+r"(a[i]) : "r"(x), "r"(y));
            if (MODE == 2) a[i] = __vimax3_s16x2(a[i], x, y);
            if (MODE == 3) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(x));
            if (MODE == 4) asm volatile("max.f32 %0, %0, %1;" : "+r"(a[i]) : "r"(x));
            // pipe-sharing probes: alternate two instruction kinds on independent registers
            if (MODE == 5) {  // VHMNMX + HMNMX2
                if (i & 1) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(x));
                else asm volatile("{\n\t.reg .b32 t;\n\tmax.f16x2 t, %0, %1;\n\tmax.f16x2 %0, t, %2;\n\t}" : "+r"(a[i]) : "r"(x), "r"(y));
            }
            if (MODE == 6) {  // HMNMX2 + HFMA2 (fma pipe)
                if (i & 1) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(x));
                else asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(x), "r"(y));
            }
            if (MODE == 7) {  // VHMNMX + HFMA2
                if (i & 1) asm volatile("{\n\t.reg .b32 t;\n\tmax.f16x2 t, %0, %1;\n\tmax.f16x2 %0, t, %2;\n\t}" : "+r"(a[i]) : "r"(x), "r"(y));
                else asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(x), "r"(y));
            }
            if (MODE == 8) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(x), "r"(y));
        }
        x += 0x00010001u;
        if (MODE == 2) y ^= a[0] & 1;
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float h2f(uint16_t h) { __half_raw r; r.x = h; return __half2float(__half(r)); }
static uint16_t f2h(float f) { __half h = __float2half_rn(f); return __half_raw(h).x; }

int main()
{
    // operand images, matching-sweep layout: row r, core c at (r >> 3) * 1024 + c * 128 + (r & 7) * 16
    std::vector<uint16_t> A(TM * 64, 0), B(TN * 64, 0);
    std::vector<float> Af(TM * 48, 0.f), Bf(TN * 48, 0.f);
    srand(7);
    auto fill = [&](std::vector<uint16_t> &img, std::vector<float> &val, int rows, bool brole) {
        for (int r = 0; r < rows; ++r) {
            float f[32], n = 0.f;
            for (int k = 0; k < 32; ++k) { f[k] = (float)rand() / RAND_MAX - 0.5f; n += f[k] * f[k]; }
            const float s = (r % 5 == 0 ? 1.0f : 0.999f) / sqrtf(n);
            float v[48] = {0};
            for (int k = 0; k < 32; ++k) v[k] = h2f(f2h(f[k] * s));
            if (!brole) { v[32] = 1.f; v[33] = 1.f; }
            else {
                float nn = 0.f;
                for (int k = 0; k < 32; ++k) nn += f[k] * s * f[k] * s;
                const float c = -0.5f * nn, hi = h2f(f2h(c));
                v[32] = hi; v[33] = h2f(f2h(c - hi));
            }
            for (int k = 0; k < 48; ++k) {
                val[r * 48 + k] = v[k];
                const int core = k >> 3;
                img[((r >> 3) * 1024 + core * 128 + (r & 7) * 16) / 2 + (k & 7)] = f2h(v[k]);
            }
        }
    };
    fill(A, Af, TM, false);
    fill(B, Bf, TN, true);
    // make a few B rows near-duplicates of A rows so that values near +0.5 (the interesting range) exist
    uint4 *dA, *dB; uint32_t *d32, *d16r, *d16p;
    cudaMalloc(&dA, 16384); cudaMalloc(&dB, 32768);
    cudaMalloc(&d32, TM * TN * 4); cudaMalloc(&d16r, TM * TN * 4); cudaMalloc(&d16p, TM * TN * 2);
    cudaMemcpy(dA, A.data(), 16384, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), 32768, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 2048);
    k_tile<<<1, 128, 16384 + 32768 + 2048>>>(dA, dB, d32, d16r, d16p);
    cudaError_t e = cudaDeviceSynchronize();
    printf("k_tile: %s\n", cudaGetErrorString(e));
    std::vector<uint32_t> h32(TM * TN), h16r(TM * TN), h16p(TM * TN / 2);
    cudaMemcpy(h32.data(), d32, TM * TN * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(h16r.data(), d16r, TM * TN * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(h16p.data(), d16p, TM * TN * 2, cudaMemcpyDeviceToHost);
    double max32 = 0, max16 = 0, max16ulp = 0, sum16ulp = 0; long lo_first = 0, hi_first = 0, rawhi_nonzero = 0, below = 0, above = 0;
    for (int i = 0; i < TM; ++i)
        for (int j = 0; j < TN; ++j) {
            double ex = 0;
            for (int k = 0; k < 48; ++k) ex += (double)Af[i * 48 + k] * (double)Bf[j * 48 + k];
            float v32; memcpy(&v32, &h32[i * TN + j], 4);
            const uint32_t raw = h16r[i * TN + j];
            if (raw >> 16) ++rawhi_nonzero;
            const float v16 = h2f((uint16_t)(raw & 0xffff));
            const uint32_t pk = h16p[i * (TN / 2) + j / 2];
            const uint16_t lo = pk & 0xffff, hi = pk >> 16;
            if (((j & 1) ? hi : lo) == (uint16_t)(raw & 0xffff)) ++lo_first;
            if (((j & 1) ? lo : hi) == (uint16_t)(raw & 0xffff)) ++hi_first;
            max32 = fmax(max32, fabs(v32 - ex));
            const double err = fabs(v16 - ex);
            max16 = fmax(max16, err);
            int ee; frexp(fabs(ex) > 1e-6 ? fabs(ex) : 1e-6, &ee);
            const double ulp = ldexp(1.0, ee - 11);  // fp16 ulp at |ex|
            max16ulp = fmax(max16ulp, err / ulp); sum16ulp += err / ulp;
            if (v16 < ex) ++below; else if (v16 > ex) ++above;
        }
    printf("fp32 accumulators: max |err| %.3e | fp16 accumulators: max |err| %.3e = %.3f ulp(fp16), mean %.3f ulp; below/above exact %ld/%ld\n",
           max32, max16, max16ulp, sum16ulp / (TM * TN), below, above);
    printf("raw cell upper half non-zero in %ld cells; pack::16b: lower column in LOW half matches %ld / %d, in HIGH half %ld\n",
           rawhi_nonzero, lo_first, TM * TN, hi_first);

    int sms = 0, clk_khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    uint32_t *out; cudaMalloc(&out, sizeof(uint32_t) * sms * 512);
    cudaEvent_t ea, eb; cudaEventCreate(&ea); cudaEventCreate(&eb);
    const char *names[9] = {"FMNMX3 (max.f32 x3)", "VHMNMX (max.f16x2 x3)", "VIMNMX3.S16x2", "HMNMX2 (max.f16x2 x2)", "FMNMX (max.f32 x2)",
                             "VHMNMX + HMNMX2 mix", "HMNMX2 + HFMA2 mix", "VHMNMX + HFMA2 mix", "HFMA2"};
    for (int mode = 0; mode < 9; ++mode) {
        const int iters = 20000;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(ea);
            if (mode == 0) k_maxrate<0><<<sms, 512>>>(out, iters);
            if (mode == 1) k_maxrate<1><<<sms, 512>>>(out, iters);
            if (mode == 2) k_maxrate<2><<<sms, 512>>>(out, iters);
            if (mode == 3) k_maxrate<3><<<sms, 512>>>(out, iters);
            if (mode == 4) k_maxrate<4><<<sms, 512>>>(out, iters);
            if (mode == 5) k_maxrate<5><<<sms, 512>>>(out, iters);
            if (mode == 6) k_maxrate<6><<<sms, 512>>>(out, iters);
            if (mode == 7) k_maxrate<7><<<sms, 512>>>(out, iters);
            if (mode == 8) k_maxrate<8><<<sms, 512>>>(out, iters);
            cudaEventRecord(eb);
            cudaEventSynchronize(eb);
        }
        float ms; cudaEventElapsedTime(&ms, ea, eb);
        const double winst = 16.0 * iters * 8;  // per SM
        printf("%-24s %.2f warp-instructions / ns / SM  (%.2f per clk at the nominal %.0f MHz)\n", names[mode],
               winst / (ms * 1e6), winst / (ms * 1e-3) / (clk_khz * 1e3), clk_khz / 1e3);
    }
    return 0;
}
