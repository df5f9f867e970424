// lr_prof.cu -- optional device-time accounting for bench.py and an FP32 peak probe.
//
// When enabled, the library brackets its heavy kernels with CUDA events on the
// launching stream; lr_prof_read() resolves them after a synchronise.  Off by
// default: the product path records nothing.
#include <vector>

#include "lr_common.cuh"

namespace lr {

struct ProfEntry {
    int kind;
    cudaEvent_t a, b;
};
static bool g_prof_on = false;
static std::vector<ProfEntry> g_prof;
static std::vector<cudaEvent_t> g_pool;

static cudaEvent_t ev_get()
{
    cudaEvent_t e;
    if (!g_pool.empty()) {
        e = g_pool.back();
        g_pool.pop_back();
        return e;
    }
    cudaEventCreate(&e);
    return e;
}

bool prof_on() { return g_prof_on; }

int prof_begin(int kind, cudaStream_t st)
{
    if (!g_prof_on) return -1;
    ProfEntry p;
    p.kind = kind;
    p.a = ev_get();
    p.b = ev_get();
    cudaEventRecord(p.a, st);
    g_prof.push_back(p);
    return (int)g_prof.size() - 1;
}

void prof_end(int token, cudaStream_t st)
{
    if (token < 0) return;
    cudaEventRecord(g_prof[token].b, st);
}

}  // namespace lr

LR_EXPORT int lr_prof_enable(int on)
{
    lr::Lock lock;
    lr::g_prof_on = on != 0;
    return LR_OK;
}

// Sum of device milliseconds and launch count of kernel class `kind`
// (LR_PROF_*) since the last read; clears the class.  Synchronises the device.
LR_EXPORT int lr_prof_read(int kind, double *total_ms, int64_t *launches)
{
    lr::Lock lock;
    LR_CUDA_TRY(cudaDeviceSynchronize());
    double ms = 0.0;
    int64_t cnt = 0;
    std::vector<lr::ProfEntry> keep;
    for (auto &p : lr::g_prof) {
        if (p.kind != kind) {
            keep.push_back(p);
            continue;
        }
        float t = 0.f;
        if (cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess) {
            ms += t;
            ++cnt;
        }
        lr::g_pool.push_back(p.a);
        lr::g_pool.push_back(p.b);
    }
    lr::g_prof.swap(keep);
    if (total_ms) *total_ms = ms;
    if (launches) *launches = cnt;
    return LR_OK;
}

namespace {

// dependent-chain-free FMA storm: 8 independent accumulators per thread
template <int MODE>
__global__ void __launch_bounds__(256) k_fma_probe(float *out, int iters, float a, float b)
{
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    if (MODE == 0) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
                x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
            }
        }
    } else {
        // packed fp32x2 FMA (Blackwell FFMA2): same FLOPs in half the issue slots
        unsigned long long p0, p1, p2, p3, pa, pb;
        asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(x0), "f"(x1));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(x2), "f"(x3));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(x4), "f"(x5));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(x6), "f"(x7));
        asm("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
        asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pa), "l"(pb));
            }
        }
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(p0));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x2), "=f"(x3) : "l"(p1));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x4), "=f"(x5) : "l"(p2));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x6), "=f"(x7) : "l"(p3));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace

// Measured FP32 FMA throughput of the current device in TFLOP/s (2 flops per
// FMA): mode 0 = scalar FFMA, mode 1 = packed fma.rn.f32x2.  Used by bench.py
// as the roofline denominator of the inlier-sweep kernel.
LR_EXPORT int lr_peak_fp32(int mode, double *tflops)
{
    lr::Lock lock;
    LR_REQUIRE(tflops && (mode == 0 || mode == 1), "bad arguments");
    const int blocks = lr::sm_count() * 8, threads = 256, iters = 4096;
    float *out = (float *)lr::arena_get(lr::SLOT_MISC, sizeof(float) * blocks * threads);
    if (!out) return LR_ERR_ALLOC;
    cudaEvent_t a, b;
    LR_CUDA_TRY(cudaEventCreate(&a));
    LR_CUDA_TRY(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        LR_CUDA_TRY(cudaEventRecord(a, 0));
        if (mode == 0) k_fma_probe<0><<<blocks, threads>>>(out, iters, 1.0000001f, 1e-9f);
        else k_fma_probe<1><<<blocks, threads>>>(out, iters, 1.0000001f, 1e-9f);
        LR_CUDA_TRY(cudaEventRecord(b, 0));
        LR_CUDA_TRY(cudaEventSynchronize(b));
        float ms = 0.f;
        LR_CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    const double fmas = (double)blocks * threads * (double)iters * 64.0;
    *tflops = 2.0 * fmas / (best * 1e-3) / 1e12;
    return LR_OK;
}
