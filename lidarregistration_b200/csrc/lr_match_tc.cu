// lr_match_tc.cu -- tensor-core nearest-neighbour sweep (tcgen05 + TMEM + bulk-TMA) for D = 32.
//
// Replaces the inner loop of find_nn (Experiments/algorithms/matching.py:22-65 of the
// reference: 250-row SGEMM chunks + norms + sqrt + min) with a distance GEMM on the 5th-gen
// tensor cores whose epilogue never leaves the SM:
//
//   operands   fp16 copies of the (power-of-two scaled) features in the canonical no-swizzle
//              K-major "core matrix" layout, 128 B per row: 4 cores of 8 features, 2 cores of
//              A-role extension (1, 1, 0..) and 2 cores of B-role extension (c_hi, c_lo, 0..)
//              with c = -s^2 |b|^2 / 2, so that one K = 48 MMA chain yields
//              v_ij = s^2 (a_i . b_j - |b_j|^2 / 2)  =  -s^2 (d2_ij - |a_i|^2) / 2;
//              when the B side's norms are uniform (L2-normalised FCGF features) the extension
//              step is skipped (K = 32) and the band is widened by the norm spread
//   staging    cp.async.bulk (UBLKCP, the TMA engine) global -> shared, completion on mbarriers,
//              4-stage ring of 256-row target tiles, double-buffered 128-row query tile
//   MMA        a warp-uniform issue loop, elect.sync picks the lane: tcgen05.mma.cta_group::1
//              .kind::f16 (M 128, N 256, K 16) x 2-3 into one of two 256-column accumulators
//              in TMEM; fp16 accumulators by default (fp32 kept for A/B), one mbarrier per tile
//   epilogue   16 warps (TMEM lane quadrant = warp % 4).  fp16 accumulators: groups 0-1 drain
//              the even tiles, groups 2-3 the odd ones, 128 columns per warp and tile through
//              tcgen05.ld.32x32b.x32.pack::16b x 2; the TMEM buffer is released as soon as the
//              values are in registers; 3-input VHMNMX trees per 32-column chunk, ONE compare.
//              A chunk whose maximum reaches the row's threshold becomes an event
//              (chunk id | 6-bit column sub-group mask, chunk maximum) in the warp's private
//              region -- no atomics.  The running maximum of a row is shared by the groups
//              through shared memory, and every work item starts with a short seed phase
//              (maxima only) so thresholds are tight early.
//   re-rank    k_rerank (one warp per row) keeps the events within `beta` of the final
//              (second) maximum and evaluates their flagged columns with the reference's fp32
//              expression, same operation order as oracle/lr_oracle.c
//
// `beta` bounds the fp16 operand / accumulator error, so the evaluated set provably contains the
// argmin of the reference's fp32 expression, which makes the indices bit-exact.  Rows with more
// events than slots are redone by an exact scan.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include "lr_match_tc.cuh"

namespace lr_tc {

#ifdef LR_TC_TIMING
static inline float __uint_as_float_host(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
extern Prepared g_last; extern int g_last_regions; extern int64_t g_last_rows;
#endif

constexpr int TM = 128;                      // query rows per tile (UMMA M)
constexpr int TN = 256;                      // target rows per tile (UMMA N)
constexpr int KCORES = 8;                    // 16-byte K cores per row (4 data + 2 A-ext + 2 B-ext)
constexpr int RG_BYTES = KCORES * 128;       // one group of 8 rows
constexpr int A_TILE_BYTES = TM / 8 * RG_BYTES;  // 16 KB
constexpr int B_TILE_BYTES = TN / 8 * RG_BYTES;  // 32 KB
constexpr int STAGES = 4;
constexpr int CAND = 32;                     // event slots per (query row, column split, epilogue group)
constexpr int NGROUPS = 4;                   // epilogue groups of four warps (one TMEM lane quadrant each)
constexpr int NTHREADS = 64 + 128 * NGROUPS;  // the epilogue groups first, then the producer and the MMA warp
// The sub-partition arbiter favours the highest warp id among eligible warps (B300_MICROARCH.md, confirmed by the
// per-tile trace in tools/): the single MMA-issuing thread must never queue behind the issue-bound epilogue warps
// it shares a sub-partition with, or the tensor pipe idles between tiles.  Hence the last warp ids.
// (Two issuing threads on alternate tiles were tried: their instruction streams interleave in the tensor pipe,
// both tiles complete late, and the sweep got slower.)
constexpr int WARP_PRODUCER = 4 * NGROUPS, WARP_MMA = 4 * NGROUPS + 1;
constexpr uint32_t IDESC = (1u << 4) /*D = f32*/ | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
// A, B = f16 (format 0), both K-major (0), no negate, dense

struct Params {
    float scale;   // power of two applied to both feature sets before the fp16 conversion
    float beta;    // half-width of the candidate band in v units (fp32 accumulators)
    unsigned maxn0_bits, maxn1_bits;
    int ovf_count;
    float beta16;  // the same with fp16 accumulators
    unsigned minn0_bits, minn1_bits;
    // per sweep direction (0: f1 plays the B role, 1: f0 does): skip the norm-extension MMA when the B side's
    // squared norms are (nearly) uniform, and the band half-width that goes with that choice
    int k32[2];
    float beta_k32[2], beta16_k32[2];
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// Blocking wait on a phase, on a precomputed shared-space address (keeps address arithmetic out of the hot
// loops).  try_wait with a suspend-time hint: the thread sleeps in hardware instead of burning issue slots
// the epilogue warps need.  (Plain try_wait and a test_wait spin were measured: no difference / slower.)
__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(addr), "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) { mbar_wait_addr(smem_u32(bar), parity); }
__device__ __forceinline__ void mbar_arrive_addr(uint32_t addr)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ int lds_volatile_s32(uint32_t addr)
{
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine; completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// One tile's tensor-core work from one elected lane, as a single block: 2 (uniform-norm path) or 3 chained MMAs
// with the K-step descriptor adds, then the two commits.
__device__ __forceinline__ void tc_issue_tile(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool k48,
                                              uint32_t bar_stage, uint32_t bar_acc)
{
    asm volatile(
        "{\n\t"
        ".reg .pred e, e3, t, f;\n\t"
        ".reg .b64 a1, b1, a2, b2;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 e3, %4, 0;\n\t"
        "and.pred e3, e3, e;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "setp.ne.b32 f, 0, 0;\n\t"
        "add.u64 a1, %1, 16;\n\t"
        "add.u64 b1, %2, 16;\n\t"
        "add.u64 a2, %1, 32;\n\t"
        "add.u64 b2, %2, 48;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%7, %7, %7, %7}, f;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, {%7, %7, %7, %7}, t;\n\t"
        "@e3 tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, {%7, %7, %7, %7}, t;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)k48), "r"(bar_stage), "r"(bar_acc), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint32_t bar_addr)
{
    asm volatile(
        "{\n\t"
        ".reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(bar_addr)
        : "memory");
}
// K-major, no swizzle: 8-row groups SBO apart, K cores LBO apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr)
{
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(RG_BYTES >> 4) << 32) |
           (1ull << 46);
}
// tcgen05.ld is asynchronous: the registers are valid only after tcgen05.wait::ld.  The wait lists
// them as in/out operands so the compiler cannot schedule a use above it.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32_wait(uint32_t (&r)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                   "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                   "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// fp16 accumulators: .pack::16b puts two adjacent 16-bit columns in one register (lower column in the low half,
// tools/micro_acc16.cu), so 32 registers carry 64 columns
__device__ __forceinline__ void tmem_ld32p_issue(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// Scheduling fence for 32 live registers: the compiler may not move their uses above this point
// (used to keep a tcgen05.ld issue ahead of the ALU work on the previous chunk).
__device__ __forceinline__ void pin32(uint32_t (&r)[32])
{
    asm volatile(""
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                   "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                   "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// ---------------------------------------------------------------- operand preparation
__global__ void k_params_reset(Params *p)
{
    lr::pdl_wait();
    lr::pdl_launch();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        p->maxn0_bits = 0u;
        p->maxn1_bits = 0u;
        p->minn0_bits = 0x7f800000u;
        p->minn1_bits = 0x7f800000u;
        p->ovf_count = 0;
    }
}

// canonical squared norm of one 32-d row (8 strided lanes added in order, == oracle)
__device__ __forceinline__ float sqnorm32_canonical(const float *__restrict__ f)
{
    const float4 *x = reinterpret_cast<const float4 *>(f);
    float lane[8];
    float4 a = x[0], b = x[1];
    lane[0] = a.x * a.x; lane[1] = a.y * a.y; lane[2] = a.z * a.z; lane[3] = a.w * a.w;
    lane[4] = b.x * b.x; lane[5] = b.y * b.y; lane[6] = b.z * b.z; lane[7] = b.w * b.w;
#pragma unroll
    for (int k = 1; k < 4; ++k) {
        a = x[2 * k];
        b = x[2 * k + 1];
        lane[0] = lane[0] + a.x * a.x; lane[1] = lane[1] + a.y * a.y;
        lane[2] = lane[2] + a.z * a.z; lane[3] = lane[3] + a.w * a.w;
        lane[4] = lane[4] + b.x * b.x; lane[5] = lane[5] + b.y * b.y;
        lane[6] = lane[6] + b.z * b.z; lane[7] = lane[7] + b.w * b.w;
    }
    float s = lane[0];
#pragma unroll
    for (int l = 1; l < 8; ++l) s = s + lane[l];
    return s;
}

// both feature sets in one launch: the first nb0 blocks take f0, the rest f1
__global__ void k_sqnorms_both(const float *__restrict__ F0, int64_t N, float *__restrict__ out0, const float *__restrict__ F1,
                               int64_t M, float *__restrict__ out1, int nb0, Params *p)
{
    lr::pdl_wait();
    lr::pdl_launch();
    const bool second = (int)blockIdx.x >= nb0;
    const float *F = second ? F1 : F0;
    const int64_t n = second ? M : N;
    float *out = second ? out1 : out0;
    const int64_t i = (int64_t)(blockIdx.x - (second ? nb0 : 0)) * blockDim.x + threadIdx.x;
    float s = 0.f;
    if (i < n) {
        s = sqnorm32_canonical(F + i * 32);
        out[i] = s;
    }
    float m = s, lo = i < n ? s : INFINITY;  // non-negative floats order like their bit patterns
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (m > 0.f) atomicMax(second ? &p->maxn1_bits : &p->maxn0_bits, __float_as_uint(m));
        atomicMin(second ? &p->minn1_bits : &p->minn0_bits, __float_as_uint(lo));
    }
}

// scale = 2^-ceil(log2 sqrt(max |f|^2)): every scaled element and row norm is <= 1.
// beta (v units): fp16 rounding of both operands 2^-10 |a'||b'|, hi/lo split of the norm term,
// tensor-core accumulation (2^-20 of the term magnitudes, generous), rounding of the canonical
// fp32 expression itself (2^-18); doubled because the maximum and the true argmin both move.
__device__ __forceinline__ float params_scale(const Params *p)
{
    const float mx = fmaxf(__uint_as_float(p->maxn0_bits), __uint_as_float(p->maxn1_bits));
    float scale = 1.f;
    if (mx > 0.f) {
        int e;
        frexpf(sqrtf(mx) * 1.000001f, &e);  // sqrt(mx) <= 2^e
        e = e < -60 ? -60 : (e > 60 ? 60 : e);
        scale = ldexpf(1.f, -e);
    }
    return scale;
}
__device__ void params_finish(Params *p)
{
    const float scale = params_scale(p);
    const float an = scale * sqrtf(__uint_as_float(p->maxn0_bits)) * 1.000001f;
    const float bn = scale * sqrtf(__uint_as_float(p->maxn1_bits)) * 1.000001f;
    const float nm = fmaxf(an, bn);  // either side can play the B role (reverse sweep)
    const float e_dot = 9.9e-4f * an * bn + 3.5e-7f;
    const float e_norm = 2.4e-7f * nm * nm + 6e-8f;
    const float e_acc = 9.6e-7f * (an * bn + 0.5f * nm * nm);
    const float e_canon = 3.9e-6f * (an * bn + an * an + bn * bn);
    p->scale = scale;
    p->beta = 2.f * (e_dot + e_norm + e_acc) + 2.f * e_canon;
    // fp16 accumulators: |v| <= an bn + nm^2 / 2; each of the three chained MMAs leaves a result rounded to
    // fp16, taken as <= 1 ulp (round-to-nearest measured at <= 0.5, tools/micro_acc16.cu) of the largest
    // binade the partial sums can reach: 3 x 2^-10 x 2^ceil(log2 vmax) / 2 <= 3 x 2^-10 x vmax
    const float vmax = an * bn + 0.5f * nm * nm;
    const float e_acc16 = 3.f * 9.8e-4f * vmax + 6e-8f;
    p->beta16 = 2.f * (e_dot + e_norm + e_acc + e_acc16) + 2.f * e_canon;
    // Uniform-norm fast path (FCGF features are L2-normalised): without the extension the MMA yields
    // s^2 a.b = v + s^2 |b|^2 / 2, i.e. v up to a constant and a per-column deviation of at most
    // s^2 (max |b|^2 - min |b|^2) / 4; taken when that costs less than a quarter of the band.
    const float spread[2] = {0.25f * scale * scale * (__uint_as_float(p->maxn1_bits) - __uint_as_float(p->minn1_bits)),
                             0.25f * scale * scale * (__uint_as_float(p->maxn0_bits) - __uint_as_float(p->minn0_bits))};
    for (int r = 0; r < 2; ++r) {
        const bool ok = spread[r] >= 0.f && 2.f * spread[r] <= 0.25f * p->beta;
        p->k32[r] = ok ? 1 : 0;
        p->beta_k32[r] = ok ? p->beta + 2.f * spread[r] : p->beta;
        p->beta16_k32[r] = ok ? p->beta16 + 2.f * spread[r] : p->beta16;
    }
}

// fp32 [N,32] -> fp16 core-matrix image (128 B per row), padded to n_pad rows
// both sides in one launch (the first nb0 blocks take f0); thread 0 also derives the band parameters
__global__ void k_prep16(const float *__restrict__ F0, const float *__restrict__ nsq0, int64_t N0, int64_t n_pad0,
                         uint4 *__restrict__ out0, const float *__restrict__ F1, const float *__restrict__ nsq1, int64_t N1,
                         int64_t n_pad1, uint4 *__restrict__ out1, int nb0, Params *p)
{
    lr::pdl_wait();
    lr::pdl_launch();
    const bool second = (int)blockIdx.x >= nb0;
    const float *F = second ? F1 : F0, *nsq = second ? nsq1 : nsq0;
    const int64_t N = second ? N1 : N0, n_pad = second ? n_pad1 : n_pad0;
    uint4 *out = second ? out1 : out0;
    const float s = params_scale(p);  // (p->scale itself is written below, by one thread, for the later kernels)
    if (blockIdx.x == 0 && threadIdx.x == 0) params_finish(p);
    const int64_t gid = (int64_t)(blockIdx.x - (second ? nb0 : 0)) * blockDim.x + threadIdx.x;
    const int64_t row = gid >> 3;
    const int core = (int)(gid & 7);
    if (row >= n_pad) return;
    union {
        __half h[8];
        uint4 u;
    } v;
#pragma unroll
    for (int k = 0; k < 8; ++k) v.h[k] = __float2half_rn(0.f);
    if (core < 4) {
        if (row < N) {
            const float4 a = reinterpret_cast<const float4 *>(F + row * 32 + core * 8)[0];
            const float4 b = reinterpret_cast<const float4 *>(F + row * 32 + core * 8)[1];
            v.h[0] = __float2half_rn(a.x * s); v.h[1] = __float2half_rn(a.y * s);
            v.h[2] = __float2half_rn(a.z * s); v.h[3] = __float2half_rn(a.w * s);
            v.h[4] = __float2half_rn(b.x * s); v.h[5] = __float2half_rn(b.y * s);
            v.h[6] = __float2half_rn(b.z * s); v.h[7] = __float2half_rn(b.w * s);
        }
    } else if (core == 4) {  // A role: picks up the B-role norm term twice (hi + lo)
        v.h[0] = __float2half_rn(1.f);
        v.h[1] = __float2half_rn(1.f);
    } else if (core == 6) {  // B role: c = -s^2 |b|^2 / 2 as hi + lo halves
        if (row < N) {
            const float c = -0.5f * (s * s) * nsq[row];
            const __half hi = __float2half_rn(c);
            v.h[0] = hi;
            v.h[1] = __float2half_rn(c - __half2float(hi));
        } else {
            v.h[0] = __float2half_rn(-60000.f);  // padding rows can never be a maximum
        }
    }
    out[((row >> 3) * KCORES + core) * 8 + (row & 7)] = v.u;
}

// ---------------------------------------------------------------- the sweep
#ifdef LR_TC_TIMING
__device__ unsigned long long g_tc_timing[16];
#define TC_T0() const long long _t0 = clock64()
#define TC_ACC(var) var += clock64() - _t0
#else
#define TC_T0()
#define TC_ACC(var)
#endif

#ifdef LR_TC_TRACE
// per-tile event timeline of CTA 0 (all 16 epilogue warps + the MMA threads), tiles [256, 288)
__device__ long long g_tc_trace[18][32][4];
#define TC_TRACE(role, ev)                                                                                \
    do {                                                                                                  \
        if (blockIdx.x == 0 && lane == 0 && (role) >= 0 && t_it >= 256u && t_it < 288u)                   \
            g_tc_trace[role][t_it - 256u][ev] = clock64();                                                \
    } while (0)
#else
#define TC_TRACE(role, ev)
#endif

constexpr int SMAX_BUFS = 4;  // rotating per item; buffer (i + 1) % 4 is reset during item i (last used by item i - 3)
struct __align__(8) Smem {
    // go[s]: tile `it` (stage s = it % 4, accumulator buffer s & 1) may be issued -- its operand tile has landed
    // (one arrive.expect_tx + the bytes) AND the warps that drain buffer s & 1 have released tile it - 2.  One
    // barrier, so the MMA warp waits once per tile.
    uint64_t a_full[2], a_empty[2], go[STAGES], b_empty[STAGES], t_full[2];
    uint32_t tmem_base;
    int smax[SMAX_BUFS][TM];  // running maximum (2-NN variant: running second maximum) of every query row of the
                              // tile, shared by the epilogue groups, as order-preserving integer keys
};

// float <-> signed integer key with the same ordering (atomicMax on shared memory)
__device__ __forceinline__ int fkey(float v)
{
    const int i = __float_as_int(v);
    return i >= 0 ? i : (i ^ 0x7fffffff);
}
__device__ __forceinline__ float fkey_inv(int k) { return __int_as_float(k >= 0 ? k : (k ^ 0x7fffffff)); }
constexpr int KEY_NEG_INF = (int)(0xff800000u ^ 0x7fffffffu);

// tile sequence of one work item: a short seed phase over tiles spread across the item's range
// (running maxima only, no candidates) and then the full sweep.  Seeding makes the thresholds
// tight before the sweep starts: ~ln(columns) prefix-maximum records per row become ~ln(16).
__device__ __forceinline__ int seed_tiles(int ntl, int seed_cfg) { return ntl >= 32 ? seed_cfg : 0; }
__device__ __forceinline__ int tile_at(int sidx, int nseed, int t_lo, int ntl)
{
    return sidx < nseed ? t_lo + (int)(((unsigned)sidx * (unsigned)ntl) / (unsigned)nseed) : t_lo + (sidx - nseed);
}

__device__ __forceinline__ float max32(const uint32_t (&r)[32])
{
    float m[11];
#pragma unroll
    for (int k = 0; k < 10; ++k)
        m[k] = fmaxf(fmaxf(__uint_as_float(r[3 * k]), __uint_as_float(r[3 * k + 1])), __uint_as_float(r[3 * k + 2]));
    m[10] = fmaxf(__uint_as_float(r[30]), __uint_as_float(r[31]));
    const float a = fmaxf(fmaxf(m[0], m[1]), m[2]), b = fmaxf(fmaxf(m[3], m[4]), m[5]);
    const float c = fmaxf(fmaxf(m[6], m[7]), m[8]), d = fmaxf(m[9], m[10]);
    return fmaxf(fmaxf(a, b), fmaxf(c, d));
}

// 64 packed columns (32 f16x2 registers) -> the maxima of their two 32-column chunks
__device__ __forceinline__ float max16p_at(const uint32_t (&r)[32], int o)
{
    __half2 h[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) h[k] = *reinterpret_cast<const __half2 *>(&r[o + k]);
    __half2 m[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) m[k] = __hmax2(__hmax2(h[3 * k], h[3 * k + 1]), h[3 * k + 2]);
    const __half2 a = __hmax2(__hmax2(m[0], m[1]), m[2]);
    const __half2 b = __hmax2(__hmax2(m[3], m[4]), h[15]);
    const __half2 c = __hmax2(a, b);
    return fmaxf(__low2float(c), __high2float(c));
}
__device__ __forceinline__ void max32p(const uint32_t (&r)[32], float &lo, float &hi)
{
    lo = max16p_at(r, 0);
    hi = max16p_at(r, 16);
}

// The last column tile is padded up to 256 target rows.  With the norm extension the padding rows carry
// a -60000 bias and can never be a maximum; without it (uniform-norm path) their zero features give
// v = 0, which may exceed a row whose dot products are all negative.  So the columns >= M of the last
// tile are forced to -inf in registers (one warp-uniform branch per tile, taken once per work item).
__device__ __forceinline__ void mask_tail32(uint32_t (&v)[32], int col0, int M)
{
#pragma unroll
    for (int k = 0; k < 32; ++k)
        if (col0 + k >= M) v[k] = 0xff800000u;
}
__device__ __forceinline__ void mask_tail32p(uint32_t (&r)[32], int col0, int M)
{
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        if (col0 + 2 * k >= M) r[k] = (r[k] & 0xffff0000u) | 0xfc00u;
        if (col0 + 2 * k + 1 >= M) r[k] = (r[k] & 0x0000ffffu) | 0xfc000000u;
    }
}

// A 32-column chunk whose maximum reaches the row's threshold (running maximum, or running second
// maximum for the 2-NN variant, minus beta) is an EVENT: the running maxima absorb the chunk maximum
// and (chunk id, chunk maximum) is appended to the thread's private region of the candidate table
// (one writer per region: register counter, plain stores).  Nothing else happens in the sweep: a
// dozen instructions, so a warp that hits an event does not fall behind the fifteen others it
// shares every accumulator buffer with (the mask-building version of this path cost 300-1000 cycles
// and stalled the whole tile pipeline; per-tile trace in tools/).  k_rerank later drops the events
// whose maximum ended up below (final maximum - beta) and evaluates the survivors' 32 columns exactly.
// Maxima are tracked per chunk: the second largest CHUNK maximum is a lower bound of the row's
// second largest value, so thresholds derived from it stay safe.  The threshold never exceeds
// (final [second] maximum - beta), so the recorded set covers every column the re-rank has to see.
// An event also carries a 6-bit mask of the chunk's column sub-groups (columns 6k .. 6k+5 for k < 5, then 30-31)
// whose maximum exceeds the threshold as it stands after the event: only those can hold a column within
// beta of the final maximum, and the re-rank reads only them (its cost is the L2 traffic of target rows).
template <bool WANT2, class MaskFn>
__device__ __forceinline__ void note_chunk(float mx, int chunk, bool record, float beta, float shared_base, float &m1,
                                           float &m2, float &thr, int2 *__restrict__ slots, int &cnt, MaskFn &&submask)
{
    if (mx > thr) {
        if (WANT2) m2 = fmaxf(m2, fminf(m1, mx));
        m1 = fmaxf(m1, mx);
        thr = fmaxf(WANT2 ? m2 : m1, shared_base) - beta;
        if (record) {
            if (cnt < CAND) slots[cnt] = make_int2((int)(((unsigned)chunk << 6) | submask(thr)), __float_as_int(mx));
            ++cnt;
        }
    }
}
// sub-group masks of a chunk held as 32 fp32 registers / as 16 packed f16x2 registers starting at r[o]
__device__ __forceinline__ unsigned submask32(const uint32_t (&v)[32], float thr)
{
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        float x = __uint_as_float(v[6 * k]);
#pragma unroll
        for (int i = 1; i < 6; ++i) x = fmaxf(x, __uint_as_float(v[6 * k + i]));
        m |= x > thr ? 1u << k : 0u;
    }
    m |= fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31])) > thr ? 32u : 0u;
    return m;
}
__device__ __forceinline__ unsigned submask16p(const uint32_t (&r)[32], int o, float thr)
{
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const __half2 x = __hmax2(__hmax2(*reinterpret_cast<const __half2 *>(&r[o + 3 * k]),
                                          *reinterpret_cast<const __half2 *>(&r[o + 3 * k + 1])),
                                  *reinterpret_cast<const __half2 *>(&r[o + 3 * k + 2]));
        m |= fmaxf(__low2float(x), __high2float(x)) > thr ? 1u << k : 0u;
    }
    const __half2 y = *reinterpret_cast<const __half2 *>(&r[o + 15]);
    m |= fmaxf(__low2float(y), __high2float(y)) > thr ? 32u : 0u;
    return m;
}

// Work items of the persistent CTAs, taken round-robin: the first `full_rb` row blocks sweep all column tiles
// (whole waves of full items), the remaining row blocks -- the last, partial wave -- are cut into `nsplit`
// column ranges each so that the tail fills the machine too.
struct Sched {
    int full_rb, nsplit, tiles_per_split, nitems;
};
struct Item {
    int rb, cs, t_lo, t_hi, nsp;
};
__device__ __forceinline__ Item item_of(int item, const Sched &sc, int n_coltiles)
{
    Item it;
    if (item < sc.full_rb) {
        it.rb = item; it.cs = 0; it.t_lo = 0; it.t_hi = n_coltiles; it.nsp = 1;
    } else {
        const int j = item - sc.full_rb;
        it.rb = sc.full_rb + j / sc.nsplit;
        it.cs = j - (j / sc.nsplit) * sc.nsplit;
        it.t_lo = it.cs * sc.tiles_per_split;
        it.t_hi = min(n_coltiles, it.t_lo + sc.tiles_per_split);
        it.nsp = sc.nsplit;
    }
    return it;
}

template <bool WANT2, bool ACC16>
__global__ void __launch_bounds__(NTHREADS, 1)
k_nn_tc(const uint4 *__restrict__ Aop, const uint4 *__restrict__ Bop, int64_t N, int64_t M, int n_rowblocks,
        int n_coltiles, Sched sc, const Params *__restrict__ params, int role,
        int seed_cfg, int2 *__restrict__ cand, int *__restrict__ cand_cnt)
{
    extern __shared__ uint8_t smem_dyn[];
    uint8_t *smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    uint8_t *sA = smem_raw;                            // 2 x 16 KB
    uint8_t *sB = smem_raw + 2 * A_TILE_BYTES;         // STAGES x 32 KB
    Smem *sm = reinterpret_cast<Smem *>(smem_raw + 2 * A_TILE_BYTES + STAGES * B_TILE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nitems = sc.nitems, nsplit = sc.nsplit;  // nsplit = region sets per row
    (void)n_rowblocks;

    if (warp == WARP_MMA) {
        if (lane == 0) {
            for (int k = 0; k < 2; ++k) {
                mbar_init(&sm->a_full[k], 1);
                mbar_init(&sm->a_empty[k], 1);
                mbar_init(&sm->t_full[k], 1);
            }
            for (int k = 0; k < STAGES; ++k) {
                mbar_init(&sm->go[k], 1 + (ACC16 ? 2 * NGROUPS : 4 * NGROUPS));  // producer + warps that drain a buffer
                mbar_init(&sm->b_empty[k], 1);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int k = threadIdx.x; k < SMAX_BUFS * TM; k += NTHREADS) (&sm->smax[0][0])[k] = KEY_NEG_INF;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm->tmem_base;
    // programmatic dependent launch: barrier init, TMEM allocation and the shared-memory reset above ran while the
    // kernel before this one (operand preparation / the other direction's merge) was draining; its results are read from here on
    lr::pdl_wait();
    lr::pdl_launch();

    if (warp == WARP_PRODUCER) {
        // ===== producer: bulk-TMA copies of operand tiles =====
        if (lane == 0) {
            uint32_t a_it = 0, b_it = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const Item wi = item_of(item, sc, n_coltiles);
                const int rb = wi.rb, t_lo = wi.t_lo, t_hi = wi.t_hi;
                const int ab = a_it & 1;
                mbar_wait(&sm->a_empty[ab], ((a_it >> 1) & 1) ^ 1);
                mbar_expect_tx(&sm->a_full[ab], A_TILE_BYTES);
                bulk_g2s(sA + ab * A_TILE_BYTES, reinterpret_cast<const uint8_t *>(Aop) + (size_t)rb * A_TILE_BYTES,
                         A_TILE_BYTES, &sm->a_full[ab]);
                ++a_it;
                const int ntl = t_hi - t_lo, nseed = seed_tiles(ntl, seed_cfg);
                for (int sidx = 0; sidx < nseed + ntl; ++sidx) {
                    const int t = tile_at(sidx, nseed, t_lo, ntl);
                    const int s = b_it % STAGES;
                    mbar_wait(&sm->b_empty[s], ((b_it / STAGES) & 1) ^ 1);
                    mbar_expect_tx(&sm->go[s], B_TILE_BYTES);
                    const uint8_t *src = reinterpret_cast<const uint8_t *>(Bop) + (size_t)t * B_TILE_BYTES;
                    bulk_g2s(sB + s * B_TILE_BYTES, src, B_TILE_BYTES / 2, &sm->go[s]);
                    bulk_g2s(sB + s * B_TILE_BYTES + B_TILE_BYTES / 2, src + B_TILE_BYTES / 2, B_TILE_BYTES / 2,
                             &sm->go[s]);
                    ++b_it;
                }
            }
        }
    } else if (warp == WARP_MMA) {
        // ===== MMA issuer =====
        // The whole warp runs the loop (warp-uniform control flow, no divergence scaffolding around the
        // tcgen05 instructions); elect.sync picks the lane that issues.  The loop is unrolled over the four
        // smem stages, so stage, accumulator buffer, barrier addresses are static and
        // a tile costs one wait, the descriptor adds and one block of issues.  Measured on B200: the previous
        // one-thread loop (~110 instructions per tile) was the sweep's bottleneck at ~660 cycles per tile.
        // (Tried and dropped: two issuing warps on alternate tiles -- their instruction streams interleave in
        // the tensor pipe and both tiles complete late; awaiting the next tile's barriers before the current
        // tile's last MMA -- whenever that wait blocks the current tile completes late.)
        static_assert(STAGES == 4, "the issue loop is unrolled over four stages / two accumulator buffers");
        if ((int)blockIdx.x < nitems) {
            auto tiles_of = [&](int item) {
                const Item wi = item_of(item, sc, n_coltiles);
                const int ntl = wi.t_hi - wi.t_lo;
                return ntl + seed_tiles(ntl, seed_cfg);
            };
            const uint32_t a_full = smem_u32(&sm->a_full[0]), a_empty = smem_u32(&sm->a_empty[0]);
            const uint32_t go = smem_u32(&sm->go[0]), b_empty = smem_u32(&sm->b_empty[0]);
            const uint32_t t_full = smem_u32(&sm->t_full[0]);
            // the 14-bit address field of a descriptor counts 16-byte units: tiles and K steps are plain adds
            const uint64_t adesc0 = smem_desc(smem_u32(sA)), bdesc0 = smem_desc(smem_u32(sB));
            constexpr uint32_t idesc = ACC16 ? (IDESC & ~(1u << 4)) /*D = f16*/ : IDESC;
            const bool k48 = params->k32[role] == 0;
            uint32_t a_it = 0, pb = 0, ntiles = 0;
            int item = blockIdx.x, left = tiles_of(item);
            uint64_t adesc = adesc0;
#ifdef LR_TC_TIMING
            const long long t_start = clock64();
            unsigned long long g_start;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_start));
#endif
            mbar_wait_addr(a_full, 0u);
            bool done = false;
            while (!done) {
#pragma unroll
                for (int s = 0; s < STAGES; ++s) {
                    const uint32_t t_it = ntiles;  // (the trace macro's tile index)
                    (void)t_it;
                    TC_TRACE(16, 0);
                    mbar_wait_addr(go + s * 8u, pb);
                    TC_TRACE(16, 2);
                    tc_fence_after();
                    const uint64_t bdesc = bdesc0 + (uint64_t)(s * (B_TILE_BYTES >> 4));
                    const uint32_t d = tmem_base + (uint32_t)(s & 1) * TN;
                    // K = 16 per instruction = two 16-byte cores: features 0-15, 16-31, then (k48) the extension;
                    // then: smem stage reusable once these MMAs have read it, accumulator ready for the epilogue
                    tc_issue_tile(d, adesc, bdesc, idesc, k48, b_empty + s * 8u, t_full + (s & 1) * 8u);
                    TC_TRACE(16, 3);
                    ++ntiles;
                    if (--left == 0) {
                        tc_commit_elect(a_empty + (a_it & 1u) * 8u);
                        ++a_it;
                        item += (int)gridDim.x;
                        if (item >= nitems) {
                            done = true;
                            break;
                        }
                        left = tiles_of(item);
                        mbar_wait_addr(a_full + (a_it & 1u) * 8u, (a_it >> 1) & 1u);
                        adesc = adesc0 + (uint64_t)((a_it & 1u) * (A_TILE_BYTES >> 4));
                    }
                }
                pb ^= 1u;
            }
#ifdef LR_TC_TIMING
            if (lane == 0) {
                unsigned long long g_end;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_end));
                atomicAdd(&g_tc_timing[10], g_end - g_start);
                atomicAdd(&g_tc_timing[3], (unsigned long long)(clock64() - t_start));
                atomicAdd(&g_tc_timing[4], (unsigned long long)ntiles);
            }
#endif
            (void)ntiles;
        }
    } else {
        // ===== epilogue: TMEM -> registers, running max + candidate collection =====
        // NGROUPS epilogue groups of four warps share every tile: group g scans 256 / NGROUPS of its
        // columns (two 32-column chunks); a row's running maximum and candidate region are per group.
        // The sweep is bound by the ISSUE slots of the four sub-partitions (four epilogue warps each;
        // traced on B200: the tensor pipe needs ~420 cycles per tile, TMEM reads ~130), so the per-tile
        // path is kept to the max tree plus a few dozen instructions: addresses are precomputed, the
        // seed phase has its own loop, and one compare covers both chunks.
        static_assert((TN / 32) / NGROUPS == 2, "the tile body below is written for two chunks per group and tile");
        const int q = warp & 3;            // TMEM lane quadrant this warp may read
        const int grp = warp >> 2;         // 0 .. NGROUPS-1
        const float beta = ACC16 ? params->beta16_k32[role] : params->beta_k32[role];
        const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + grp * 64;
        const uint32_t full_addr = smem_u32(&sm->t_full[0]), go_addr = smem_u32(&sm->go[0]);
        // the accumulator buffers start out free: the release of "tile -2" and "tile -1"
        if (lane == 0 && (!ACC16 || (warp >> 3) == 0)) mbar_arrive_addr(go_addr);
        if (lane == 0 && (!ACC16 || (warp >> 3) == 1)) mbar_arrive_addr(go_addr + 8u);
        const int rowl = q * 32 + lane;
        uint32_t t_it = 0, item_it = 0;
        const int trole = warp;
        (void)trole;
        // fp16 accumulators: groups 0-1 serve the even tiles (accumulator buffer 0), groups 2-3 the odd ones,
        // 128 columns = four packed chunks per warp and tile.  A warp then has two tile periods for its
        // (latency-bound) wait -> load -> release -> scan chain, and the two buffers' chains no longer queue
        // behind each other in the same warps (fp32 accumulators: 128 registers per tile slice, so every
        // group shares every tile, 64 columns each).
        const uint32_t my_set = (uint32_t)(grp >> 1), my_half = (uint32_t)(grp & 1);
        (void)my_set; (void)my_half;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const Item wi = item_of(item, sc, n_coltiles);
            const int rb = wi.rb, cs = wi.cs, t_lo = wi.t_lo, t_hi = wi.t_hi;
            const int64_t row = (int64_t)rb * TM + rowl;
            const bool valid = row < N;
            float m1 = -INFINITY, m2 = -INFINITY, thr = -INFINITY;
            const int64_t region = valid ? (row * nsplit + cs) * NGROUPS + grp : 0;
            int2 *slots = cand + region * CAND;
            int cnt = 0;
            int *my_smax = &sm->smax[item_it % SMAX_BUFS][rowl];
            const uint32_t smax_addr = smem_u32(my_smax);
            if (grp == 0) sm->smax[(item_it + 1) % SMAX_BUFS][rowl] = KEY_NEG_INF;  // for the next item
            ++item_it;
            const int ntl = t_hi - t_lo, nseed = seed_tiles(ntl, seed_cfg);

            // fp32 accumulators: load the tile's two chunks, hand the TMEM buffer back, scan
            auto tile32 = [&](const int t, const bool seed) {
                uint32_t va[32], vb[32];
                const uint32_t acc = t_it & 1u;
                mbar_wait_addr(full_addr + acc * 8u, (t_it >> 1) & 1u);
                TC_TRACE(trole, 0);
                tc_fence_after();
                const uint32_t ta = tbase + acc * (uint32_t)TN;
                tmem_ld32_issue(ta, va);
                tmem_ld32_issue(ta + 32, vb);
                tmem_ld32_wait(va);  // waits for every outstanding load of the thread
                pin32(vb);           // ... so the second chunk only needs its uses ordered behind the wait
                TC_TRACE(trole, 1);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_addr(go_addr + ((t_it + 2u) % STAGES) * 8u);  // the tile that reuses the buffer
                TC_TRACE(trole, 2);
                if (t == n_coltiles - 1) {
                    mask_tail32(va, t * TN + grp * 64, (int)M);
                    mask_tail32(vb, t * TN + grp * 64 + 32, (int)M);
                }
                // what the other groups (and the seed phase) have established for this row
                const float shared_base = fkey_inv(lds_volatile_s32(smax_addr));
                thr = fmaxf(thr, shared_base - beta);
                const float mxa = max32(va), mxb = max32(vb);
                if (fmaxf(mxa, mxb) > thr) {
                    const int chunk0 = t * (TN / 32) + grp * 2;
                    const bool record = valid && !seed;
                    note_chunk<WANT2>(mxa, chunk0, record, beta, shared_base, m1, m2, thr, slots, cnt,
                                      [&](float th) { return submask32(va, th); });
                    note_chunk<WANT2>(mxb, chunk0 + 1, record, beta, shared_base, m1, m2, thr, slots, cnt,
                                      [&](float th) { return submask32(vb, th); });
                    // publish this thread's running maximum (2-NN: running second maximum): any group's value
                    // is a lower bound of the row's true one, so thresholds derived from it stay safe
                    const float mine = WANT2 ? m2 : m1;
                    if (mine > shared_base) atomicMax(my_smax, fkey(mine));
                }
                TC_TRACE(trole, 3);
            };
            // fp16 accumulators: this warp's tiles only; 2 x 64 packed columns
            auto tile16 = [&](const int t, const bool seed) {
                uint32_t va[32], vb[32];
                const uint32_t acc = my_set;
                mbar_wait_addr(full_addr + acc * 8u, (t_it >> 1) & 1u);
                TC_TRACE(trole, 0);
                tc_fence_after();
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)TN + my_half * 128u;
                tmem_ld32p_issue(ta, va);
                tmem_ld32p_issue(ta + 64, vb);
                const int key = lds_volatile_s32(smax_addr);  // overlaps the TMEM loads
                tmem_ld32_wait(va);
                pin32(vb);
                TC_TRACE(trole, 1);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_addr(go_addr + ((t_it + 2u) % STAGES) * 8u);  // the tile that reuses the buffer
                TC_TRACE(trole, 2);
                if (t == n_coltiles - 1) {
                    mask_tail32p(va, t * TN + (int)my_half * 128, (int)M);
                    mask_tail32p(vb, t * TN + (int)my_half * 128 + 64, (int)M);
                }
                const float shared_base = fkey_inv(key);
                thr = fmaxf(thr, shared_base - beta);
                float mx[4];
                max32p(va, mx[0], mx[1]);
                max32p(vb, mx[2], mx[3]);
                if (fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) > thr) {
                    const int chunk0 = t * (TN / 32) + (int)my_half * 4;
                    const bool record = valid && !seed;
                    note_chunk<WANT2>(mx[0], chunk0, record, beta, shared_base, m1, m2, thr, slots, cnt,
                                      [&](float th) { return submask16p(va, 0, th); });
                    note_chunk<WANT2>(mx[1], chunk0 + 1, record, beta, shared_base, m1, m2, thr, slots, cnt,
                                      [&](float th) { return submask16p(va, 16, th); });
                    note_chunk<WANT2>(mx[2], chunk0 + 2, record, beta, shared_base, m1, m2, thr, slots, cnt,
                                      [&](float th) { return submask16p(vb, 0, th); });
                    note_chunk<WANT2>(mx[3], chunk0 + 3, record, beta, shared_base, m1, m2, thr, slots, cnt,
                                      [&](float th) { return submask16p(vb, 16, th); });
                    const float mine = WANT2 ? m2 : m1;
                    if (mine > shared_base) atomicMax(my_smax, fkey(mine));
                }
                TC_TRACE(trole, 3);
            };
            auto tile = [&](const int t, const bool seed) {
                if constexpr (ACC16) {
                    if ((t_it & 1u) == my_set) tile16(t, seed);
                } else {
                    tile32(t, seed);
                }
                ++t_it;
            };

            if (nseed > 0) {
                for (int sidx = 0; sidx < nseed; ++sidx) tile(tile_at(sidx, nseed, t_lo, ntl), true);
                // end of the seed phase: its columns are visited again by the sweep, so the local
                // running maxima restart (a revisited column must not count twice towards the
                // second maximum); what the seed established lives on in the shared threshold base
                atomicMax(my_smax, fkey(WANT2 ? m2 : m1));
                m1 = -INFINITY;
                m2 = -INFINITY;
                thr = -INFINITY;
            }
#pragma unroll 1
            for (int t = t_lo; t < t_hi; ++t) tile(t, false);
            if (valid) {
                cand_cnt[region] = cnt;
                // a full item stands for all column ranges of its rows: the other region sets stay empty
                for (int c = wi.nsp; c < nsplit; ++c) cand_cnt[(row * nsplit + c) * NGROUPS + grp] = 0;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------- exact re-rank
struct Top2 {
    float s1;
    int j1;
    float s2;
    int j2;
};

__device__ __forceinline__ void top2_put(Top2 &c, float s, int j)
{
    if (s < c.s1 || (s == c.s1 && j < c.j1)) {
        c.s2 = c.s1;
        c.j2 = c.j1;
        c.s1 = s;
        c.j1 = j;
    } else if (s < c.s2 || (s == c.s2 && j < c.j2)) {
        c.s2 = s;
        c.j2 = j;
    }
}

// the reference's fp32 expression (matching.py:29-30), oracle operation order
__device__ __forceinline__ float canon_dist(const float *__restrict__ a, const float *__restrict__ b, float na, float nb)
{
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) acc = fmaf(a[k], b[k], acc);
    const float d2 = fmaf(-2.f, acc, na + nb);
    return __fsqrt_rn(fmaxf(d2, 1e-30f));
}

// lexicographic (dist, index) minimum across the warp
__device__ __forceinline__ void warp_lexmin(float &s, int &j)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
        const int j2 = __shfl_xor_sync(0xffffffffu, j, o);
        if (s2 < s || (s2 == s && j2 < j)) {
            s = s2;
            j = j2;
        }
    }
}

// One warp per query row.  Lane L reads the event counts of regions L and L + 32 (one round trip) and walks
// their few events itself: (chunk id | sub-group mask, chunk maximum).  The largest (2-NN: the two largest) recorded chunk maximum gives the
// cut: only events within beta of it can hold a (second) nearest neighbour -- on cfg 2 about 1.2 of the
// ~3.4 events a row records.  Their 32 columns (lane = column) are evaluated with the canonical fp32
// expression; per-lane lexicographic top-2, merged by shuffles.
// (four resident blocks per SM = 64 registers: the kernel is a chain of dependent loads per row -- counts, events, target
// rows -- and 32 instead of 24 warps per SM in flight cut find_nn by 5.7 %; five blocks = 48 registers spill and lose it again)
#ifndef LR_RERANK_MINB
#define LR_RERANK_MINB 4
#endif
__global__ void __launch_bounds__(256, LR_RERANK_MINB)
k_rerank(const float *__restrict__ F0, const float *__restrict__ F1, const float *__restrict__ n0,
         const float *__restrict__ n1, int64_t N, int64_t M, int nregions, bool acc16, int role,
         const int2 *__restrict__ cand, const int *__restrict__ cand_cnt, Params *p, int *__restrict__ ovf_rows,
         int64_t *__restrict__ idx1, int64_t *__restrict__ idx2)
{
    lr::pdl_wait();
    lr::pdl_launch();
    const int lane = threadIdx.x & 31;
    const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= N) return;  // warp-uniform
    const int *cc = cand_cnt + row * nregions;
    const int c0 = lane < nregions ? cc[lane] : 0;
    const int c1 = lane + 32 < nregions ? cc[lane + 32] : 0;
    // the query row, early: its latency hides behind the event bookkeeping
    float a[32];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float4 x = reinterpret_cast<const float4 *>(F0 + row * 32)[k];
        a[4 * k] = x.x; a[4 * k + 1] = x.y; a[4 * k + 2] = x.z; a[4 * k + 3] = x.w;
    }
    const float na = n0[row];
    const float beta = acc16 ? p->beta16_k32[role] : p->beta_k32[role];
    if (__any_sync(0xffffffffu, c0 > CAND || c1 > CAND)) {  // more events than slots somewhere: exact scan instead
        if (lane == 0) ovf_rows[atomicAdd(&p->ovf_count, 1)] = (int)row;
        return;
    }
    // lane L owns regions L and L + 32: it walks their (few) events itself, no cross-lane search
    int maxc = c0 > c1 ? c0 : c1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, o));
    const int2 *ev0 = cand + (row * nregions + lane) * CAND, *ev1 = cand + (row * nregions + lane + 32) * CAND;
    // pass 0: the two largest chunk maxima among the row's events
    float e1 = -INFINITY, e2 = -INFINITY;
    for (int k = 0; k < maxc; ++k) {
        if (k < c0) {
            const float mx = __int_as_float(ev0[k].y);
            e2 = fmaxf(e2, fminf(e1, mx));
            e1 = fmaxf(e1, mx);
        }
        if (k < c1) {
            const float mx = __int_as_float(ev1[k].y);
            e2 = fmaxf(e2, fminf(e1, mx));
            e1 = fmaxf(e1, mx);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float o1 = __shfl_xor_sync(0xffffffffu, e1, o), o2 = __shfl_xor_sync(0xffffffffu, e2, o);
        e2 = fmaxf(fmaxf(e2, o2), fminf(e1, o1));
        e1 = fmaxf(e1, o1);
    }
    const float keep = (idx2 ? e2 : e1) - beta;  // events below this cannot hold a (second) nearest neighbour
    // pass 1: the flagged columns of the surviving events, lane = column
    Top2 c;
    c.s1 = INFINITY; c.j1 = 0x7fffffff; c.s2 = INFINITY; c.j2 = 0x7fffffff;
    auto evaluate = [&](int2 ev, bool take) {
        unsigned todo = __ballot_sync(0xffffffffu, take);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const unsigned cx = (unsigned)__shfl_sync(0xffffffffu, ev.x, src);
            const int64_t j = (int64_t)(cx >> 6) * 32 + lane;
            const int sub = lane >= 30 ? 5 : lane / 6;  // the lane's column sub-group
            if (j < M && ((cx >> sub) & 1u)) {
                float b[32];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 x = reinterpret_cast<const float4 *>(F1 + j * 32)[q];
                    b[4 * q] = x.x; b[4 * q + 1] = x.y; b[4 * q + 2] = x.z; b[4 * q + 3] = x.w;
                }
                top2_put(c, canon_dist(a, b, na, n1[j]), (int)j);
            }
        }
    };
    for (int k = 0; k < maxc; ++k) {
        int2 ev = make_int2(0, 0);
        bool take = false;
        if (k < c0) {
            ev = ev0[k];
            take = __int_as_float(ev.y) >= keep;
        }
        evaluate(ev, take);
        if (nregions > 32) {
            take = false;
            if (k < c1) {
                ev = ev1[k];
                take = __int_as_float(ev.y) >= keep;
            }
            evaluate(ev, take);
        }
    }
    // merge the per-lane top-2 lists
    float s = c.s1;
    int j = c.j1;
    warp_lexmin(s, j);
    const int best = j;
    float t = c.j1 == best ? c.s2 : c.s1;
    int u = c.j1 == best ? c.j2 : c.j1;
    warp_lexmin(t, u);
    if (lane == 0) {
        idx1[row] = best == 0x7fffffff ? 0 : best;
        if (idx2) idx2[row] = u == 0x7fffffff ? 0 : u;
    }
}

// exact scan of the overflowed rows, canonical arithmetic over every column.  A work item is (row, column
// segment), so that a handful of rows still fills the machine; partial top-2 lists are merged by k_row_merge.
constexpr int OVF_SEGS = 64;
__device__ __forceinline__ int ovf_segments(int novf, int64_t partial_cap)
{
    int64_t s = partial_cap / (novf > 0 ? novf : 1);
    return (int)(s > OVF_SEGS ? OVF_SEGS : (s < 1 ? 1 : s));
}

__global__ void __launch_bounds__(256)
k_row_exact(const float *__restrict__ F0, const float *__restrict__ F1, const float *__restrict__ n0,
            const float *__restrict__ n1, int64_t M, const Params *__restrict__ p, const int *__restrict__ ovf_rows,
            Top2 *__restrict__ partial, int64_t partial_cap)
{
    lr::pdl_wait();
    lr::pdl_launch();
    __shared__ Top2 sh[256];
    __shared__ float a[32];
    const int novf = p->ovf_count;
    const int nseg = ovf_segments(novf, partial_cap);
    const int64_t seg_len = (M + nseg - 1) / nseg;
    for (int64_t w = blockIdx.x; w < (int64_t)novf * nseg; w += gridDim.x) {
        const int o = (int)(w / nseg), sg = (int)(w - (int64_t)o * nseg);
        const int64_t row = ovf_rows[o];
        __syncthreads();
        if (threadIdx.x < 32) a[threadIdx.x] = F0[row * 32 + threadIdx.x];
        __syncthreads();
        const float na = n0[row];
        Top2 c;
        c.s1 = INFINITY; c.j1 = 0x7fffffff; c.s2 = INFINITY; c.j2 = 0x7fffffff;
        const int64_t j_hi = (sg + 1) * seg_len < M ? (sg + 1) * seg_len : M;
        for (int64_t j = sg * seg_len + threadIdx.x; j < j_hi; j += blockDim.x) {
            float b[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 x = reinterpret_cast<const float4 *>(F1 + j * 32)[q];
                b[4 * q] = x.x; b[4 * q + 1] = x.y; b[4 * q + 2] = x.z; b[4 * q + 3] = x.w;
            }
            top2_put(c, canon_dist(a, b, na, n1[j]), (int)j);
        }
        sh[threadIdx.x] = c;
        __syncthreads();
        for (int stride = 128; stride > 0; stride >>= 1) {
            if ((int)threadIdx.x < stride) {
                Top2 m = sh[threadIdx.x];
                const Top2 o2 = sh[threadIdx.x + stride];
                if (o2.j1 != 0x7fffffff) top2_put(m, o2.s1, o2.j1);
                if (o2.j2 != 0x7fffffff) top2_put(m, o2.s2, o2.j2);
                sh[threadIdx.x] = m;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) partial[w] = sh[0];
    }
}

__global__ void __launch_bounds__(1024)
k_row_merge(Params *__restrict__ p, const int *__restrict__ ovf_rows, const Top2 *__restrict__ partial,
            int64_t partial_cap, int64_t *__restrict__ idx1, int64_t *__restrict__ idx2)
{
    lr::pdl_wait();
    lr::pdl_launch();
    const int novf = p->ovf_count;
    const int nseg = ovf_segments(novf, partial_cap);
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < novf; o += gridDim.x * blockDim.x) {
        Top2 m;
        m.s1 = INFINITY; m.j1 = 0x7fffffff; m.s2 = INFINITY; m.j2 = 0x7fffffff;
        for (int sg = 0; sg < nseg; ++sg) {
            const Top2 c = partial[(int64_t)o * nseg + sg];
            if (c.j1 != 0x7fffffff) top2_put(m, c.s1, c.j1);
            if (c.j2 != 0x7fffffff) top2_put(m, c.s2, c.j2);
        }
        const int64_t row = ovf_rows[o];
        idx1[row] = m.j1 == 0x7fffffff ? 0 : m.j1;
        if (idx2) idx2[row] = m.j2 == 0x7fffffff ? 0 : m.j2;
    }
    __syncthreads();  // single block: everyone has read the count
    if (threadIdx.x == 0) p->ovf_count = 0;
}

// ---------------------------------------------------------------- host side
static int64_t pad_rows(int64_t n) { return (n + TN - 1) / TN * TN; }
// candidate regions (row x column split x epilogue group) the scratch is sized for
static int64_t region_budget(int64_t rows)
{
    const int64_t floor_regions = (int64_t)1 << 20;  // lets small problems split their columns finely
    return rows * NGROUPS > floor_regions ? rows * NGROUPS : floor_regions;
}

// partial top-2 lists of the overflow scan: (overflowed rows) x (column segments) always fits
static int64_t partial_cap(int64_t rows) { return rows > ((int64_t)1 << 16) ? rows : ((int64_t)1 << 16); }

size_t scratch_bytes(int64_t N, int64_t M)
{
    const int64_t mx = N > M ? N : M;
    return lr::padded(sizeof(Params)) + lr::padded(pad_rows(N) * 128) + lr::padded(pad_rows(M) * 128) +
           lr::padded(sizeof(float) * N) + lr::padded(sizeof(float) * M) + lr::padded(sizeof(int2) * region_budget(mx) * CAND) +
           lr::padded(sizeof(int) * region_budget(mx)) + lr::padded(sizeof(int) * mx) +
           lr::padded(sizeof(Top2) * partial_cap(mx));
}

// norms, scale, band and the fp16 operand images of both feature sets (once per match call)
int prepare(const float *f0, int64_t N, const float *f1, int64_t M, char *scratch, Prepared &P, cudaStream_t st)
{
    const int64_t mx = N > M ? N : M;
    lr::Carver cv(scratch);
    P.params = cv.take<Params>(1);
    P.pad0 = pad_rows(N);
    P.pad1 = pad_rows(M);
    P.op0 = reinterpret_cast<uint4 *>(cv.take<char>(P.pad0 * 128));
    P.op1 = reinterpret_cast<uint4 *>(cv.take<char>(P.pad1 * 128));
    P.n0 = cv.take<float>(N);
    P.n1 = cv.take<float>(M);
    P.cand = cv.take<int2>(region_budget(mx) * CAND);
    P.cand_cnt = cv.take<int>(region_budget(mx));
    P.ovf_rows = cv.take<int>(mx);
    P.partial = cv.take<char>(sizeof(Top2) * partial_cap(mx));
    P.partial_cap = partial_cap(mx);
    LR_CUDA_TRY(lr::launch_pdl(k_params_reset, dim3(1), dim3(32), 0, st, P.params));
    const int nbs0 = (int)((N + 255) / 256), nbs1 = (int)((M + 255) / 256);
    LR_CUDA_TRY(lr::launch_pdl(k_sqnorms_both, dim3(nbs0 + nbs1), dim3(256), 0, st, f0, N, P.n0, f1, M, P.n1, nbs0, P.params));
    const int nbp0 = (int)((P.pad0 * 8 + 255) / 256), nbp1 = (int)((P.pad1 * 8 + 255) / 256);
    LR_CUDA_TRY(lr::launch_pdl(k_prep16, dim3(nbp0 + nbp1), dim3(256), 0, st, f0, P.n0, N, P.pad0, P.op0, f1, P.n1, M, P.pad1, P.op1, nbp0,
                               P.params));
    LR_CUDA_TRY(cudaGetLastError());
    return LR_OK;
}

// nearest (and second nearest) neighbour of every row of side `a` among the rows of side `b`
int sweep(const Prepared &P, bool swap, bool acc16, const float *f0, int64_t N, const float *f1, int64_t M,
          int64_t *idx1, int64_t *idx2, cudaStream_t st)
{
    const float *fa = swap ? f1 : f0, *fb = swap ? f0 : f1;
    const float *na = swap ? P.n1 : P.n0, *nb = swap ? P.n0 : P.n1;
    const uint4 *opa = swap ? P.op1 : P.op0, *opb = swap ? P.op0 : P.op1;
    const int64_t Na = swap ? M : N, Nb = swap ? N : M;
    const int n_rowblocks = (int)((Na + TM - 1) / TM);
    const int n_coltiles = (int)((Nb + TN - 1) / TN);
    const int sms = lr::sm_count();
    // Persistent CTAs take items round-robin (Sched): whole waves of full row blocks, then the row blocks of the
    // last partial wave cut into S column ranges.  S minimises rounds x (tiles per range + a per-item overhead of
    // ~12 tiles: pipeline fill, query tile, seed phase); 391 row blocks on 148 SMs: 296 full + 95 x 3 ranges =
    // 2.67 full-item times instead of 3 (one split for everything) or 8 x (1/3 + overhead) (three for everything).
    const int64_t region_cap = region_budget(N > M ? N : M) / (Na * NGROUPS);
    Sched sc;
    sc.full_rb = n_rowblocks / sms * sms;
    const int tail = n_rowblocks - sc.full_rb;
    sc.nsplit = 1;
    if (tail > 0) {
        double best_cost = 1e30;
        for (int S = 1; S <= 16; ++S) {
            if (S > n_coltiles || S > region_cap) break;
            const int tps_c = (n_coltiles + S - 1) / S;
            if (S > 1 && tps_c < 8) break;
            const int ns = (n_coltiles + tps_c - 1) / tps_c;
            const long long rounds = ((long long)tail * ns + sms - 1) / sms;
            const double cost = (double)rounds * (tps_c + 12.0);
            if (cost < best_cost - 1e-9) {
                best_cost = cost;
                sc.nsplit = ns;
            }
        }
    }
    if (const char *e = getenv("LR_TC_NSPLIT")) {  // experiments only
        const int v = atoi(e);
        if (v >= 1 && v <= n_coltiles && v <= region_cap) sc.nsplit = v;
    }
    if (getenv("LR_TC_NOFULL")) sc.full_rb = 0;  // experiments only: every row block split alike
    sc.tiles_per_split = (n_coltiles + sc.nsplit - 1) / sc.nsplit;
    sc.nsplit = (n_coltiles + sc.tiles_per_split - 1) / sc.tiles_per_split;
    sc.nitems = sc.full_rb + (n_rowblocks - sc.full_rb) * sc.nsplit;
    const int nsplit = sc.nsplit, nitems = sc.nitems;
    int seed_cfg = 8;
    if (const char *e = getenv("LR_TC_SEEDS")) seed_cfg = atoi(e);
    const int grid = nitems < sms ? nitems : sms;
    const size_t smem = 2 * A_TILE_BYTES + STAGES * B_TILE_BYTES + sizeof(Smem) + 1024 + 64;
    // (no clearing of cand_cnt: the sweep writes the count of every region of every valid row; ovf_count is zero
    // from k_params_reset and k_row_merge puts it back to zero)
    const int tok = lr::prof_begin(lr::PROF_NN, st);
    auto launch = [&](auto kern) -> int {
        LR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LR_CUDA_TRY(lr::launch_pdl(kern, dim3(grid), dim3(NTHREADS), smem, st, opa, opb, Na, Nb, n_rowblocks, n_coltiles, sc,
                                   P.params, swap ? 1 : 0, seed_cfg, P.cand, P.cand_cnt));
        return LR_OK;
    };
    int lrc;
    if (acc16) lrc = idx2 ? launch(k_nn_tc<true, true>) : launch(k_nn_tc<false, true>);
    else lrc = idx2 ? launch(k_nn_tc<true, false>) : launch(k_nn_tc<false, false>);
    if (lrc) return lrc;
    lr::prof_end(tok, st);
#ifdef LR_TC_TIMING
    g_last = P; g_last_regions = nsplit * NGROUPS; g_last_rows = Na;
#endif
    LR_CUDA_TRY(lr::launch_pdl(k_rerank, dim3((unsigned)((Na + 7) / 8)), dim3(256), 0, st, fa, fb, na, nb, Na, Nb, nsplit * NGROUPS,
                               acc16, swap ? 1 : 0, P.cand, P.cand_cnt, P.params, P.ovf_rows, idx1, idx2));
    LR_CUDA_TRY(lr::launch_pdl(k_row_exact, dim3(sms * 2), dim3(256), 0, st, fa, fb, na, nb, Nb, P.params, P.ovf_rows,
                               reinterpret_cast<Top2 *>(P.partial), P.partial_cap));
    LR_CUDA_TRY(lr::launch_pdl(k_row_merge, dim3(1), dim3(1024), 0, st, P.params, P.ovf_rows,
                               reinterpret_cast<const Top2 *>(P.partial), P.partial_cap, idx1, idx2));
    LR_CUDA_TRY(cudaGetLastError());
    return LR_OK;
}

#ifdef LR_TC_TIMING
void timing_dump()
{
    unsigned long long h[16];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(h, g_tc_timing, sizeof(h));
    const double t = (double)(h[4] ? h[4] : 1);
    fprintf(stderr, "[tc timing] SM clock during the sweep %.0f MHz | ", h[10] ? 1e3 * (double)h[3] / (double)h[10] : 0.0);
    fprintf(stderr, "[tc timing] tiles %llu | MMA per tile: wait b_full %.0f, wait t_empty %.0f, issue %.0f, total %.0f | "
            "epilogue per tile: wait t_full %.0f, scan %.0f, wait::ld %.0f, fence+arrive %.0f, try+issue %.0f\n", h[4], h[0] / t, h[1] / t, h[2] / t, h[3] / t, h[5] / t, h[6] / t, h[7] / t, h[8] / t, h[9] / t);
    memset(h, 0, sizeof(h));
    cudaMemcpyToSymbol(g_tc_timing, h, sizeof(h));
}
#endif

}  // namespace lr_tc

#ifdef LR_TC_TRACE
LR_EXPORT int lr_tc_trace_dump(void)
{
    static long long h[18][32][4];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(h, lr_tc::g_tc_trace, sizeof(h));
    long long t0 = h[16][0][0] ? h[16][0][0] : h[17][1][0];
    fprintf(stderr, "[tc trace] epilogue warp w (group w/4, sub-partition w%%4): t_full seen, chunks landed, released, scan done | "
                    "mma thread: top, b_full, t_empty, issued+committed\n");
    for (int t = 8; t < 20; ++t) {
        for (int r = 16; r < 17; ++r)
            if (h[r][t][0])
                fprintf(stderr, "tile %2d mma%d : %6lld %6lld %6lld %6lld\n", t, r - 16, h[r][t][0] - t0, h[r][t][1] - t0,
                        h[r][t][2] - t0, h[r][t][3] - t0);
        for (int sp = 0; sp < 4; ++sp) {
            fprintf(stderr, "tile %2d sp%d  :", t, sp);
            for (int g = 0; g < 4; ++g) {
                const int w = g * 4 + sp;
                fprintf(stderr, "  [g%d %6lld %6lld %6lld %6lld]", g, h[w][t][0] - t0, h[w][t][1] - t0, h[w][t][2] - t0, h[w][t][3] - t0);
            }
            fprintf(stderr, "\n");
        }
    }
    return 0;
}
#endif

#ifdef LR_TC_TIMING
namespace lr_tc { Prepared g_last; int g_last_regions = 0; int64_t g_last_rows = 0; }
LR_EXPORT int lr_tc_debug_stats(void)
{
    using namespace lr_tc;
    cudaDeviceSynchronize();
    Params hp;
    cudaMemcpy(&hp, g_last.params, sizeof(hp), cudaMemcpyDeviceToHost);
    const int64_t n = g_last_rows * g_last_regions;
    int *h = (int *)malloc(sizeof(int) * n);
    cudaMemcpy(h, g_last.cand_cnt, sizeof(int) * n, cudaMemcpyDeviceToHost);
    double sum = 0; int mx = 0; int64_t over = 0;
    for (int64_t i = 0; i < n; ++i) { sum += h[i]; if (h[i] > mx) mx = h[i]; if (h[i] > CAND) ++over; }
    fprintf(stderr, "[tc debug] scale %g beta %g maxn0 %g maxn1 %g ovf_rows %d | regions/row %d avg cand/region %.2f max %d regions over %lld\n",
            hp.scale, hp.beta, __uint_as_float_host(hp.maxn0_bits), __uint_as_float_host(hp.maxn1_bits), hp.ovf_count,
            g_last_regions, sum / n, mx, (long long)over);
    free(h);
    return 0;
}
LR_EXPORT int lr_tc_timing_dump(void)
{
    lr_tc::timing_dump();
    return 0;
}
#endif
