// lr_common.cuh -- shared host/device plumbing of liblidarreg (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/lidarreg.h"

#define LR_EXPORT extern "C" __attribute__((visibility("default")))

namespace lr {

// thread-local error text returned by lr_last_error()
void set_error(const char *fmt, ...);

#define LR_CUDA_TRY(expr)                                                                          \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            lr::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));   \
            return LR_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

#define LR_REQUIRE(cond, msg)                                        \
    do {                                                             \
        if (!(cond)) {                                               \
            lr::set_error("%s:%d: %s", __FILE__, __LINE__, msg);     \
            return LR_ERR_ARG;                                       \
        }                                                            \
    } while (0)

// Grow-only device scratch, one arena per (device, slot).  The library owns
// nothing else on the device.  Not shared between concurrent calls: every entry
// point takes the arena lock for its duration (the reference path is called
// from one Python thread per process, Experiments/test.py:108-167).
enum Slot { SLOT_RANSAC = 0, SLOT_MATCH = 1, SLOT_MISC = 2, SLOT_RANSAC_B = 3, SLOT_COUNT = 4 };

struct Arena {
    void *ptr = nullptr;
    size_t cap = 0;
    uint64_t gen = 0;  // bumped whenever the block is (re)allocated or released: cached contents are gone
};

// returns nullptr (and sets the error) on failure
void *arena_get(int slot, size_t bytes);
void arena_release_all();
uint64_t arena_gen(int slot);
int sm_count();

struct Lock {
    Lock();
    ~Lock();
};

// optional device-time accounting (lr_prof.cu); tokens are -1 when disabled
enum ProfKind { PROF_SCORE = 0, PROF_GEN = 1, PROF_NN = 2, PROF_RECOUNT = 3, PROF_PACK = 4, PROF_END = 5, PROF_FIN = 6 };
bool prof_on();
int prof_begin(int kind, cudaStream_t st);
void prof_end(int token, cudaStream_t st);

// Programmatic dependent launch for the kernel chain of a run (pack -> gen -> Kabsch -> sweep -> round end -> finish):
// the next kernel's CTAs are scheduled while the previous kernel drains and block in pdl_wait() until it has completed
// and its writes are visible, so the ~2-3 us of launch latency per boundary leave the critical path of a single-pair call
// (and of every rank of a hypothesis-sharded one).  EVERY kernel launched through launch_pdl() starts with pdl_wait();
// launched the ordinary way the two instructions are no-ops.  lr_debug_pdl(0) / profiling runs launch the ordinary way.
extern int g_pdl;
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (g_pdl && !prof_on()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// carve 256-byte aligned pieces out of one arena block
struct Carver {
    char *base;
    size_t off = 0;
    explicit Carver(void *p) : base(reinterpret_cast<char *>(p)) {}
    template <class T>
    T *take(size_t count)
    {
        T *r = reinterpret_cast<T *>(base + off);
        off += (count * sizeof(T) + 255) & ~size_t(255);
        return r;
    }
};
inline size_t padded(size_t bytes) { return (bytes + 255) & ~size_t(255); }

}  // namespace lr
