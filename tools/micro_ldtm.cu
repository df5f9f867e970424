// micro_ldtm.cu -- how tcgen05.ld (LDTM) shares an SM sub-partition with ALU work on B200.
// Every warp loops over: load one 32-column chunk of its TMEM lane quadrant, then run `reps` 3-input-max
// trees over the registers.  Reported: cycles per loop iteration for 1 / 2 / 4 warps per sub-partition, with
//   LD    0 = none (ALU only), 1 = 32x32b.x32 (fp32 cells), 2 = 32x32b.x16.pack::16b, 3 = 32x32b.x32.pack::16b
//   PIPE  0 = load, wait, compute;  1 = the next load is issued before computing on the previous chunk
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_ldtm tools/micro_ldtm.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define R32(r) "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
#define P32(r) "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
#define L32 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
#define L16 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%32];"

template <int LD>
__device__ __forceinline__ void ld(uint32_t ta, uint32_t (&r)[32])
{
    if (LD == 1) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " L32 : R32(r) : "r"(ta) : "memory");
    if (LD == 2) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 " L16 : R32(r) : "r"(ta) : "memory");
    if (LD == 3) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 " L32 : R32(r) : "r"(ta) : "memory");
}
__device__ __forceinline__ void wait(uint32_t (&r)[32]) { asm volatile("tcgen05.wait::ld.sync.aligned;" : P32(r)::"memory"); }
__device__ __forceinline__ uint32_t tree(const uint32_t (&r)[32], uint32_t acc)
{
    uint32_t m[11];
#pragma unroll
    for (int k = 0; k < 10; ++k) asm volatile("max.f32 %0, %1, %2, %3;" : "=r"(m[k]) : "r"(r[3 * k]), "r"(r[3 * k + 1]), "r"(r[3 * k + 2]));
    asm volatile("max.f32 %0, %1, %2, %3;" : "=r"(m[10]) : "r"(r[30]), "r"(r[31]), "r"(acc));
    uint32_t a, b, c, d;
    asm volatile("max.f32 %0, %1, %2, %3;" : "=r"(a) : "r"(m[0]), "r"(m[1]), "r"(m[2]));
    asm volatile("max.f32 %0, %1, %2, %3;" : "=r"(b) : "r"(m[3]), "r"(m[4]), "r"(m[5]));
    asm volatile("max.f32 %0, %1, %2, %3;" : "=r"(c) : "r"(m[6]), "r"(m[7]), "r"(m[8]));
    asm volatile("max.f32 %0, %1, %2, %3;" : "=r"(d) : "r"(m[9]), "r"(m[10]), "r"(a));
    asm volatile("max.f32 %0, %1, %2, %3;" : "=r"(a) : "r"(b), "r"(c), "r"(d));
    return a;
}

template <int LD, int PIPE>
__global__ void __launch_bounds__(512) k(unsigned long long *cycles, uint32_t *sink, int iters, int reps)
{
    __shared__ uint32_t tmem_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_s + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 64;
    uint32_t a[32], b[32], acc = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) { a[i] = threadIdx.x + i; b[i] = threadIdx.x * 3 + i; }
    __syncthreads();
    const long long t0 = clock64();
    if (PIPE == 0 || LD == 0) {
        for (int it = 0; it < iters; ++it) {
            if (LD != 0) { ld<LD>(base + (it & 1) * 32, a); wait(a); }
            for (int k = 0; k < reps; ++k) acc = tree(a, acc);
        }
    } else {
        ld<LD>(base, a);
        for (int it = 0; it < iters; it += 2) {
            wait(a);
            ld<LD>(base + 32, b);
            for (int k = 0; k < reps; ++k) acc = tree(a, acc);
            wait(b);
            ld<LD>(base, a);
            for (int k = 0; k < reps; ++k) acc = tree(b, acc);
        }
        wait(a);
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) atomicMax(&cycles[blockIdx.x], (unsigned long long)(t1 - t0));
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + a[5] + b[7];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_s), "r"(512u) : "memory");
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long *cyc, h[256];
    uint32_t *sink;
    cudaMalloc(&cyc, sizeof(unsigned long long) * sms);
    cudaMalloc(&sink, sizeof(uint32_t) * sms * 512);
    const int iters = 4000;
    const char *ldn[4] = {"none", "x32 fp32", "x16.pack16", "x32.pack16"};
    for (int ldm = 0; ldm < 4; ++ldm)
        for (int pipe = 0; pipe < 2; ++pipe) {
            if (ldm == 0 && pipe) continue;
            for (int reps = 0; reps <= 4; reps += (reps == 0 ? 1 : reps)) {
                if (ldm == 0 && reps == 0) continue;
                printf("ld %-10s pipe %d trees/iter %d:", ldn[ldm], pipe, reps);
                for (int warps = 4; warps <= 16; warps *= 2) {
                    cudaMemset(cyc, 0, sizeof(unsigned long long) * sms);
#define RUN(L, P) if (ldm == L && pipe == P) k<L, P><<<sms, warps * 32>>>(cyc, sink, iters, reps)
                    RUN(0, 0); RUN(1, 0); RUN(1, 1); RUN(2, 0); RUN(2, 1); RUN(3, 0); RUN(3, 1);
                    cudaError_t e = cudaDeviceSynchronize();
                    cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
                    printf("  %d warps/SMSP: %7.1f cyc/iter%s", warps / 4, (double)h[0] / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
                }
                printf("\n");
            }
        }
    return 0;
}
