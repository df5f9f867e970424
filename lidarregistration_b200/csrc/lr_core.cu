// lr_core.cu -- error text, device scratch arenas, library-level entry points.
#include <mutex>
#include <stdarg.h>

#include "lr_common.cuh"

namespace lr {

int g_pdl = 1;
static thread_local char g_err[512] = "";
static std::recursive_mutex g_mu;
static const int kMaxDev = 64;
static Arena g_arena[kMaxDev][SLOT_COUNT];
static int g_sms[kMaxDev];

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

Lock::Lock() { g_mu.lock(); }
Lock::~Lock() { g_mu.unlock(); }

void *arena_get(int slot, size_t bytes)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) {
        set_error("no CUDA device (cudaGetDevice failed)");
        return nullptr;
    }
    Arena &a = g_arena[dev][slot];
    if (a.cap < bytes) {
        if (a.ptr) {
            cudaDeviceSynchronize();  // nobody may still be reading the old block
            cudaFree(a.ptr);
            a.ptr = nullptr;
            a.cap = 0;
        }
        size_t want = bytes + bytes / 4 + (size_t(1) << 20);
        cudaError_t e = cudaMalloc(&a.ptr, want);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            a.ptr = nullptr;
            return nullptr;
        }
        // control blocks at the head of an arena rely on zeroed "last block" tickets
        e = cudaMemset(a.ptr, 0, want < (size_t(1) << 16) ? want : (size_t(1) << 16));
        if (e != cudaSuccess) {
            set_error("cudaMemset of a fresh arena failed: %s", cudaGetErrorString(e));
            cudaFree(a.ptr);
            a.ptr = nullptr;
            return nullptr;
        }
        a.cap = want;
        ++a.gen;
    }
    return a.ptr;
}

uint64_t arena_gen(int slot)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return 0;
    return g_arena[dev][slot].gen;
}

void arena_release_all()
{
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) return;
    for (int d = 0; d < kMaxDev; ++d)
        for (int s = 0; s < SLOT_COUNT; ++s)
            if (g_arena[d][s].ptr) {
                cudaSetDevice(d);
                cudaFree(g_arena[d][s].ptr);
                const uint64_t gen = g_arena[d][s].gen + 1;
                g_arena[d][s] = Arena();
                g_arena[d][s].gen = gen;
            }
    cudaSetDevice(cur);
}

int sm_count()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return 148;
    if (g_sms[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        g_sms[dev] = v;
    }
    return g_sms[dev];
}

}  // namespace lr

LR_EXPORT const char *lr_last_error(void) { return lr::g_err; }

LR_EXPORT int lr_version(void) { return 140; }  // 140: lr_icp_refine, lr_nn3d_radius, lr_seeds_score, lr_kabsch_weighted_batch; 130: lr_gpf_filter; 120: lr_comm_*, lr_ransac_rigid_sharded, lr_ransac_tc_probe

LR_EXPORT int lr_device_info(int *sms, int *major, int *minor)
{
    int dev = 0;
    LR_CUDA_TRY(cudaGetDevice(&dev));
    int a = 0, b = 0, c = 0;
    LR_CUDA_TRY(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
    LR_CUDA_TRY(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
    LR_CUDA_TRY(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
    if (sms) *sms = a;
    if (major) *major = b;
    if (minor) *minor = c;
    return LR_OK;
}

// A/B switch of the programmatic dependent launch (lr_common.cuh): 0 = ordinary stream-ordered launches
LR_EXPORT int lr_debug_pdl(int on)
{
    lr::g_pdl = on ? 1 : 0;
    return LR_OK;
}

LR_EXPORT int lr_shutdown(void)
{
    lr::Lock lock;
    lr::arena_release_all();
    return LR_OK;
}
