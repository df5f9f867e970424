// lr_score_tc.cuh -- the inlier sweep on the 5th-gen tensor cores (tcgen05 + TMEM + bulk-TMA).
// Included by lr_ransac.cu (needs its Ctl / helper definitions).
//
// Replaces the per-(hypothesis, correspondence) residual test of the reference engines
// (Open3D EvaluateRANSACBasedOnCorrespondence behind Experiments/algorithms/FR.py:122-139; the score loop of
// GCRANSAC::run behind GC-RANSAC/src/pygcransac/src/gcransac_python.cpp:507-533) for LR_SCORE_COUNT runs.
//
// The three components of a residual, d_a = sum_b R_ab p_b + t_a - q_a, are K <= 16 dot products between a
// per-hypothesis row and a per-correspondence column, i.e. three [128 hypotheses] x [16] x [64 correspondences]
// MMAs per tile.  Unlike the bilinear form of r^2 itself nothing cancels: every term is at most a coordinate
// (~100 m), so fp16 operand pieces and the fp32 accumulator leave d_a within E ~ 1e-4 m of the canonical fp64
// value, a band of ~4e-4 m^2 around thr^2 -- about one residual per hypothesis and 30k correspondences falls
// inside it and is decided in fp64 on the spot, so every count is exact (same contract as the fp32 sweep).
//
//   operands   fp16 pieces, canonical no-swizzle K-major core matrices (8 rows x 16 B):
//                hypothesis row a   core0 [R1_a0 R1_a0 R2_a0 R1_a1 R1_a1 R2_a1 t1_a t2_a]
//                                   core1 [R1_a2 R1_a2 R2_a2 t3_a  -1    -1    -1   0   ]
//                correspondence     core0 [p1_0  p2_0  p1_0  p1_1  p2_1  p1_1  1    1   ]   (shared by a = 0,1,2)
//                                   core1_a [p1_2 p2_2 p1_2  1     q1_a  q2_a  q3_a 0   ]
//              x1 = fp16(x), x2 = fp16(x - x1), x3 = fp16(x - x1 - x2).  q and t are exact in three pieces; of
//              R p the products R1 p1 + R1 p2 + R2 p1 are kept, the dropped ones are <= 3 * 2^-22 |p|_2.
//              The leading-byte-offset field of the smem descriptor selects core1_a, so the three B operands of
//              a tile share core0: 64 B per correspondence instead of 96.
//              Coordinates are taken relative to (c, c') = the first correspondence rounded to 1024 m
//              (t~ = t + R c - c'), so map-frame offsets do not eat the fp16 range.
//   staging    cp.async.bulk (UBLKCP): 12 KB hypothesis block per segment (double buffered), 8 KB stages of
//              128 correspondences (4-deep ring), completion on mbarriers
//   MMA        one elected lane: tcgen05.mma.cta_group::1.kind::f16, M 128, N 64, K 16, three per tile (a = 0,1,2)
//              into a 192-column TMEM buffer; two buffers (sub-tile j of every 128-correspondence stage -> buffer j).
//              The issue loop is unrolled over the four stages of the B ring: addresses, parities and descriptor
//              offsets are static.
//   epilogue   16 warps = 4 TMEM lane quadrants x 4 slices of 16 columns; EVERY warp takes its slice of EVERY tile,
//              so each warp has two tiles of look-ahead (the other buffer is filled while it computes).
//              Thread = hypothesis: tcgen05.ld 32x32b.x16 of the three regions, buffer handed back at once,
//              u = d0^2 + d1^2 + d2^2 - thr^2 as three packed fma.rn.f32x2 per two residuals,
//              count += sign bit (LEA.HI), running min |u| (FMNMX3): 3.25 issue slots per residual.
//              min |u| < delta_h somewhere in the warp -> the 16 columns are checked one by one and the in-band
//              ones decided with the canonical fp64 residual (rare: ~1e-4 of the residuals).
//              (First version: N 32, four single-buffered 96-column stages each owned by one class of warps -- the
//              arrive -> issue -> MMA -> commit -> wake -> tcgen05.ld round trip of a class was exposed every tile and
//              the issue loop cost ~240 cycles per tile: 41 % issue utilisation, tensor pipe 12 % active,
//              profiles/r2_ncu_k_score_tc_v1.txt.)
// Work of a CTA = an equal share (+-1 stage) of the linearised (128-survivor block, 128-correspondence stage) space,
// persistent CTAs, one per SM; partial counts merge through integer atomics.
#pragma once
// (cuda_fp16.h is included by lr_ransac.cu at file scope: this header sits inside its anonymous namespace)

namespace tcs {

constexpr int TM = 128;                 // hypotheses per row block (UMMA M)
#ifndef LR_TCS_TN
#define LR_TCS_TN 64   // 64: two 192-column TMEM buffers, two MMA warps, every epilogue warp reads 16 columns of every tile
                       // (6.85 ms for 1 M x 30 000 against 7.15 with 32 and 7.9 with 16: fewer, larger MMAs per stage)
#endif
// LR_TCS_WIDE (with TN 32): twelve epilogue warps = three groups of four (one per TMEM lane quadrant); a group owns a
// whole 32-column tile (96 accumulator registers per thread, 128 registers after setmaxnreg), so a TMEM buffer is waited
// for and handed back by four warps instead of eight or sixteen, the fixed work per barrier round trip covers twice the
// columns, and the three warps of a sub-partition belong to three different buffers' cycles (nothing re-locks them).
#ifndef LR_TCS_WIDE
#define LR_TCS_WIDE 0
#endif
#if LR_TCS_WIDE
#undef LR_TCS_TN
#define LR_TCS_TN 32
#endif
constexpr int TN = LR_TCS_TN;           // correspondences per MMA tile (UMMA N): 16, 32 or 64
constexpr int BROWS = 128;              // correspondences per shared-memory stage (== kChunk)
constexpr int TILES_PER_STAGE = BROWS / TN;   // == number of TMEM buffers: sub-tile j of every stage lives in buffer j
constexpr int B_GROUP_BYTES = 512;      // 8 correspondences: core0 + 3 x core1_a
constexpr int B_STAGE_BYTES = BROWS / 8 * B_GROUP_BYTES;  // 8 KB
constexpr int B_STAGES = 4;
constexpr int A_PART_BYTES = TM / 8 * 256;     // one residual component: 16 groups x (core0 + core1) = 4 KB
constexpr int A_BLOCK_BYTES = 3 * A_PART_BYTES;  // 12 KB per 128 hypotheses
constexpr int NBUF = TILES_PER_STAGE;
constexpr int TMEM_BUF_COLS = 3 * TN;   // the three residual components of a tile (NBUF x 3 x TN = 384 columns in all)
#if LR_TCS_WIDE
constexpr int NEPI = 12;                // epilogue warps: 3 tile groups x 4 TMEM lane quadrants, a warp takes all 32 columns
constexpr int NSLICE = 1;
constexpr int NGROUP = 3;               // group g takes the tiles t = g (mod 3) of the CTA's tile sequence (4 per stage)
constexpr int kRegsEpilogue = 112, kRegsOther = 48;  // 12 x 32 x epilogue + 8 x 32 x other < the 640 x 96 registers of the launch (with slack)
#else
constexpr int NEPI = 16;                // epilogue warps: 4 TMEM lane quadrants x NGROUP tile groups x NSLICE column slices
constexpr int NSLICE = TN / 16;         // a warp takes 16 columns of a tile (48 accumulator registers)
constexpr int NGROUP = 4 / NSLICE;      // group g takes the tiles j = g (mod NGROUP) of every stage: while one group
                                        // computes, another is free to pick the next tile up the moment it is ready
#endif
static_assert(TN == 16 || TN == 32 || TN == 64, "tile width");
constexpr int NMMA = TILES_PER_STAGE;    // MMA-issuing warps: warp WARP_MMA + j owns sub-tile j / TMEM buffer j of every stage
constexpr int WARP_PRODUCER = NEPI, WARP_MMA = NEPI + 1;
#if LR_TCS_WIDE
constexpr int NTHREADS = 32 * ((NEPI + 1 + NMMA + 3) / 4 * 4);  // whole warpgroups (setmaxnreg moves registers between them)
#else
constexpr int NTHREADS = 32 * (NEPI + 1 + NMMA);
#endif
constexpr uint32_t IDESC = (1u << 4) /*D = f32*/ | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
constexpr float kRangeLimit = 15000.f;  // |p~|, |q~| above this: the fp16 pieces could overflow -> fp64 fallback
constexpr double kAccKappa = 4.0;       // tensor-core accumulation error <= kappa * 2^-24 * sum |terms|
                                        // (measured through lr_ransac_tc_probe: tests/test_gpu_score_tc.py)

static_assert(BROWS == kChunk, "the correspondence padding of k_pack is the stage size of the tensor sweep");

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "TCS_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra.uni TCS_WAIT_DONE;\n\t"
        "bra.uni TCS_WAIT_LOOP;\n\t"
        "TCS_WAIT_DONE:\n\t"
        "}" ::"r"(addr), "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t addr, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
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
// K-major, no swizzle: K-adjacent cores `lbo` bytes apart, 8-row groups `sbo` bytes apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// tcgen05.ld is asynchronous: the registers are valid only after tcgen05.wait::ld; listing them as in/out
// operands keeps the compiler from scheduling a use above the wait
__device__ __forceinline__ void tmem_ld_wait3(uint32_t (&a)[16], uint32_t (&b)[16], uint32_t (&c)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                   "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]),
                   "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]),
                   "+r"(b[8]), "+r"(b[9]), "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15]),
                   "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7]),
                   "+r"(c[8]), "+r"(c[9]), "+r"(c[10]), "+r"(c[11]), "+r"(c[12]), "+r"(c[13]), "+r"(c[14]), "+r"(c[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
        "%25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
          "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
          "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// (tcgen05.wait::ld, then one ordering statement per register array: volatile asm statements keep their order)
__device__ __forceinline__ void tmem_touch32(uint32_t (&r)[32])
{
    asm volatile(""
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait3x32(uint32_t (&a)[32], uint32_t (&b)[32], uint32_t (&c)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    tmem_touch32(a);
    tmem_touch32(b);
    tmem_touch32(c);
}
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait3x8(uint32_t (&a)[8], uint32_t (&b)[8], uint32_t (&c)[8])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(b[0]),
                   "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]), "+r"(c[0]), "+r"(c[1]),
                   "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7])
                 :
                 : "memory");
}
// non-blocking probe of a barrier phase (the result is consumed later: its latency hides under the work in between)
__device__ __forceinline__ uint32_t mbar_test(uint32_t addr, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ u64 pk2(uint32_t lo, uint32_t hi)
{
    u64 v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(lo), "r"(hi));
    return v;
}
__device__ __forceinline__ float min3abs(float m, float a, float b)
{
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(m), "f"(fabsf(a)), "f"(fabsf(b)));
    return r;
}

// ---------------------------------------------------------------- operand images
// x = h1 + h2 + h3 up to 2^-33 |x| (+ 2^-25 where a piece is subnormal); every difference below is exact in fp64
__device__ __forceinline__ void split3(double x, __half &h1, __half &h2, __half &h3)
{
    h1 = __double2half(x);
    double r = x - (double)__half2float(h1);
    h2 = __double2half(r);
    r = r - (double)__half2float(h2);
    h3 = __double2half(r);
}
__device__ __forceinline__ void split2(double x, __half &h1, __half &h2)
{
    h1 = __double2half(x);
    h2 = __double2half(x - (double)__half2float(h1));
}

// the frame the operands are expressed in: the first correspondence, rounded to 1024 m (exactly representable,
// zero for sensor-centred scans)
__device__ __forceinline__ void tc_centre6(float px, float py, float pz, float qx, float qy, float qz, double (&c)[3],
                                           double (&cq)[3])
{
    c[0] = 1024.0 * rint((double)px * (1.0 / 1024.0));
    c[1] = 1024.0 * rint((double)py * (1.0 / 1024.0));
    c[2] = 1024.0 * rint((double)pz * (1.0 / 1024.0));
    cq[0] = 1024.0 * rint((double)qx * (1.0 / 1024.0));
    cq[1] = 1024.0 * rint((double)qy * (1.0 / 1024.0));
    cq[2] = 1024.0 * rint((double)qz * (1.0 / 1024.0));
}
__device__ __forceinline__ void tc_centre(const float4 *__restrict__ P8, double (&c)[3], double (&cq)[3])
{
    const float4 a = __ldg(P8), b = __ldg(P8 + 1);
    tc_centre6(a.x, a.y, a.z, a.w, b.x, b.y, c, cq);
}

union Core {
    __half h[8];
    uint4 u;
};

// correspondence i -> its four cores of the B image (group of 8 correspondences = 512 B)
__device__ __forceinline__ void tc_write_corr(uint4 *__restrict__ Bimg, int64_t i, const double (&pt)[3],
                                              const double (&qt)[3])
{
    __half p1[3], p2[3];
#pragma unroll
    for (int b = 0; b < 3; ++b) split2(pt[b], p1[b], p2[b]);
    const __half one = __float2half_rn(1.f), zero = __float2half_rn(0.f);
    uint4 *grp = Bimg + (i >> 3) * (B_GROUP_BYTES / 16) + (i & 7);
    Core c0;
    c0.h[0] = p1[0]; c0.h[1] = p2[0]; c0.h[2] = p1[0];
    c0.h[3] = p1[1]; c0.h[4] = p2[1]; c0.h[5] = p1[1];
    c0.h[6] = one; c0.h[7] = one;
    grp[0] = c0.u;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        Core c1;
        c1.h[0] = p1[2]; c1.h[1] = p2[2]; c1.h[2] = p1[2];
        c1.h[3] = one;
        split3(qt[a], c1.h[4], c1.h[5], c1.h[6]);
        c1.h[7] = zero;
        grp[(1 + a) * 8] = c1.u;
    }
}

// hypothesis slot -> its six cores of the A image (block of 128 slots = 12 KB: component a, group, core, row)
__device__ __forceinline__ void tc_write_model(uint4 *__restrict__ Aimg, int slot, const double (&T)[12],
                                               const double (&tt)[3])
{
    const int blk = slot / TM, g = (slot % TM) >> 3, r = slot & 7;
    const __half m1 = __float2half_rn(-1.f), zero = __float2half_rn(0.f);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        __half r1[3], r2[3], t1, t2, t3;
#pragma unroll
        for (int b = 0; b < 3; ++b) split2(T[4 * a + b], r1[b], r2[b]);
        split3(tt[a], t1, t2, t3);
        uint4 *base = Aimg + ((size_t)blk * A_BLOCK_BYTES + a * A_PART_BYTES + g * 256) / 16 + r;
        Core c0, c1;
        c0.h[0] = r1[0]; c0.h[1] = r1[0]; c0.h[2] = r2[0];
        c0.h[3] = r1[1]; c0.h[4] = r1[1]; c0.h[5] = r2[1];
        c0.h[6] = t1; c0.h[7] = t2;
        c1.h[0] = r1[2]; c1.h[1] = r1[2]; c1.h[2] = r2[2];
        c1.h[3] = t3; c1.h[4] = m1; c1.h[5] = m1; c1.h[6] = m1; c1.h[7] = zero;
        base[0] = c0.u;
        base[8] = c1.u;
    }
}

// |d_tc - d_64| <= E for every component of every correspondence of the run (header comment; P2 = max |p~|_2,
// Q = max |q~|_inf, tinf = |t~|_inf): dropped products, piece rounding, tensor-core accumulation
__device__ __forceinline__ double tc_err_bound(double P2, double Q, double tinf)
{
    const double u22 = 2.384185791015625e-07, u24 = 5.9604644775390625e-08;
    return 3.0 * u22 * P2 * 1.001 + kAccKappa * u24 * (1.7320508075688772 * P2 + tinf + Q + 1.0) + 1e-6;
}

// ---------------------------------------------------------------- the sweep
// -DLR_TCS_TRACE: CTA 0 records clock64() of its warps' pipeline events for stages [kTraceFrom, kTraceFrom + kTraceLen)
// (tools/tcs_trace.py reads them back through lr_debug_tcs_trace)
#ifdef LR_TCS_TRACE
constexpr int kTraceFrom = 160, kTraceLen = 24;
__device__ long long g_tcs_trace[NEPI + 1 + NMMA][kTraceLen][TILES_PER_STAGE][4];
#define TCS_TRACE(stage, j, ev)                                                                       \
    do {                                                                                              \
        const int ts_ = (int)(stage) - kTraceFrom;                                                    \
        if (blockIdx.x == 0 && lane == 0 && ts_ >= 0 && ts_ < kTraceLen) g_tcs_trace[warp][ts_][j][ev] = clock64(); \
    } while (0)
#else
#define TCS_TRACE(stage, j, ev)
#endif

struct __align__(8) Smem {
    // LR_TCS_WIDE: a group meets TMEM buffer j only every third stage, and a parity wait cannot tell phase s from phase
    // s - 2 -- so "tile (s, j) is ready" has one barrier per (j, s mod 3), each with one consumer group that sees every
    // one of its phases
    uint64_t a_full[2], a_empty[2], b_full[B_STAGES], b_empty[B_STAGES], t_full[NBUF * (LR_TCS_WIDE ? 3 : 1)], t_empty[NBUF];
    uint32_t tmem_base;
};

// Work of a CTA = a contiguous range of the linearised (hypothesis block, correspondence stage) space, so every CTA
// gets the same number of stages (+-1); a range is walked as segments = (one hypothesis block) x (stage range).
struct Range {
    long long pos, end;   // linear stage index hb * nchunks + c
};
__device__ __forceinline__ Range cta_range(int nsurv, int nchunks)
{
    const long long nhb = (nsurv + TM - 1) / TM;
    const long long W = nhb * nchunks;
    Range r;
    r.pos = W * blockIdx.x / gridDim.x;
    r.end = W * (blockIdx.x + 1) / gridDim.x;
    return r;
}
struct Seg {
    int hb, c_lo, c_hi;
};
__device__ __forceinline__ Seg seg_at(long long pos, long long end, int nchunks)
{
    Seg s;
    s.hb = (int)(pos / nchunks);
    s.c_lo = (int)(pos - (long long)s.hb * nchunks);
    const long long left = end - pos;
    s.c_hi = (long long)(nchunks - s.c_lo) <= left ? nchunks : s.c_lo + (int)left;
    return s;
}

__device__ __noinline__ int tc_exact_inlier(const float4 *__restrict__ P8, int64_t i, const double *__restrict__ m64s,
                                            double thr2)
{
    double p[3], q[3], T[12];
    load_pq(P8, i, p, q);
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = m64s[k];
    return res2_f64(T, p[0], p[1], p[2], q[0], q[1], q[2]) < thr2 ? 1 : 0;
}

// one tile (TN correspondences): D_a = A_a . B_a for a = 0,1,2 (no accumulate) into three TN-column regions of a TMEM
// buffer, then the accumulator-ready commit
__device__ __forceinline__ void tc_issue_tile(uint32_t d, uint64_t a0, uint64_t a1, uint64_t a2, uint64_t b0, uint64_t b1,
                                              uint64_t b2, uint32_t bar_full)
{
    asm volatile(
        "{\n\t"
        ".reg .pred e, f;\n\t"
        ".reg .b32 d1, d2;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 f, 0, 0;\n\t"
        "add.u32 d1, %0, %10;\n\t"
        "add.u32 d2, %0, %11;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %4, %7, {%8, %8, %8, %8}, f;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [d1], %2, %5, %7, {%8, %8, %8, %8}, f;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [d2], %3, %6, %7, {%8, %8, %8, %8}, f;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t"
        "}" ::"r"(d),
        "l"(a0), "l"(a1), "l"(a2), "l"(b0), "l"(b1), "l"(b2), "r"(IDESC), "r"(0u), "r"(bar_full), "n"(TN), "n"(2 * TN)
        : "memory");
}
// the MMA warps' wait: suspending by default; LR_TCS_SPIN=1 polls instead (A/B: polling steals issue slots from the
// epilogue warps of the same sub-partition, whose slowest member gates every buffer release)
#ifndef LR_TCS_SPIN
#define LR_TCS_SPIN 0
#endif
__device__ __forceinline__ void mbar_wait_mma(uint32_t addr, uint32_t parity)
{
#if LR_TCS_SPIN
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "TCS_SPIN_LOOP:\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni TCS_SPIN_DONE;\n\t"
        "bra.uni TCS_SPIN_LOOP;\n\t"
        "TCS_SPIN_DONE:\n\t"
        "}" ::"r"(addr), "r"(parity)
        : "memory");
#else
    mbar_wait(addr, parity);
#endif
}

// DUMP: write the three tensor-core residual components of every (slot, correspondence) to `dump`
// ([slot][n_pad][3] floats) instead of counting -- the error-bound probe of tests/test_gpu_score_tc.py
template <bool DUMP>
// (21 warps are allocated as 24 -- the register file is handed out four warps at a time -- so TN = 32 with its four MMA
// warps gets 80 registers per thread; __maxnreg__(96) compiles but the launch fails)
__global__ void __launch_bounds__(NTHREADS, 1)
k_score_tc(const uint4 *__restrict__ Aimg, const uint4 *__restrict__ Bimg, const float4 *__restrict__ P8, int64_t n,
           int64_t n_pad, Ctl *ctl, const double *__restrict__ m64, const float *__restrict__ band, int *__restrict__ cnt,
           double thr2, float *__restrict__ dump, int4 *__restrict__ events, unsigned ev_cap)
{
    lr::pdl_wait();
    lr::pdl_launch();
    if (ctl->done) return;
    const int nsurv = ctl->n_surv;
    if (nsurv <= 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = (int)(n_pad / BROWS);

    // operands out of the fp16 range (a cloud more than 15 km across): exact fp64 counts on the CUDA cores
    if (!(__uint_as_float(ctl->pt2max_bits) < kRangeLimit && __uint_as_float(ctl->qtmax_bits) < kRangeLimit)) {
        if (DUMP) return;
        unsigned long long evals = 0;
        for (int slot = blockIdx.x * NTHREADS + threadIdx.x; slot < nsurv; slot += gridDim.x * NTHREADS) {
            double T[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) T[k] = m64[(size_t)slot * 12 + k];
            int c = 0;
            for (int64_t i = 0; i < n; ++i) {
                double p[3], q[3];
                load_pq(P8, i, p, q);
                c += res2_f64(T, p[0], p[1], p[2], q[0], q[1], q[2]) < thr2 ? 1 : 0;
            }
            cnt[slot] = c;
            evals += (unsigned long long)n;
        }
        if (evals) atomicAdd(reinterpret_cast<unsigned long long *>(&ctl->n_rechecked), evals);
        return;
    }

    extern __shared__ uint8_t smem_dyn[];
    uint8_t *smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    uint8_t *sA = smem_raw;                             // 2 x 12 KB
    uint8_t *sB = smem_raw + 2 * A_BLOCK_BYTES;         // B_STAGES x 8 KB
    Smem *sm = reinterpret_cast<Smem *>(smem_raw + 2 * A_BLOCK_BYTES + B_STAGES * B_STAGE_BYTES);
    const Range rg = cta_range(nsurv, nchunks);

    if (warp == WARP_MMA) {
        if (lane == 0) {
            for (int k = 0; k < 2; ++k) {
                mbar_init(&sm->a_full[k], 1);
                mbar_init(&sm->a_empty[k], NMMA);
            }
            for (int k = 0; k < B_STAGES; ++k) {
                mbar_init(&sm->b_full[k], 1);
                mbar_init(&sm->b_empty[k], NMMA);
            }
            for (int k = NBUF; k < NBUF * (LR_TCS_WIDE ? 3 : 1); ++k) mbar_init(&sm->t_full[k], 1);
            for (int k = 0; k < NBUF; ++k) {
                mbar_init(&sm->t_full[k], 1);
                mbar_init(&sm->t_empty[k], LR_TCS_WIDE ? 4 : 4 * NSLICE);  // the warps (quadrant x slice) that read a tile
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm->tmem_base;

#if LR_TCS_WIDE
    // (each role executes its setmaxnreg inside its own branch: the register allocator budgets a region by the instruction
    // that dominates it)
#define TCS_REGS_MORE() asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpilogue))
#define TCS_REGS_LESS() asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsOther))
#else
#define TCS_REGS_MORE()
#define TCS_REGS_LESS()
#endif
    if (warp == WARP_PRODUCER) {
        TCS_REGS_LESS();
        // ===== producer: bulk copies of the hypothesis block and the correspondence stages =====
        if (lane == 0) {
            uint32_t a_it = 0, b_it = 0;
            for (long long pos = rg.pos; pos < rg.end;) {
                const Seg sg = seg_at(pos, rg.end, nchunks);
                const uint32_t ab = a_it & 1u;
                mbar_wait(smem_u32(&sm->a_empty[ab]), ((a_it >> 1) & 1u) ^ 1u);
                mbar_expect_tx(smem_u32(&sm->a_full[ab]), A_BLOCK_BYTES);
                bulk_g2s(smem_u32(sA + ab * A_BLOCK_BYTES), reinterpret_cast<const uint8_t *>(Aimg) + (size_t)sg.hb * A_BLOCK_BYTES,
                         A_BLOCK_BYTES, smem_u32(&sm->a_full[ab]));
                ++a_it;
                for (int c = sg.c_lo; c < sg.c_hi; ++c) {
                    const uint32_t s = b_it % B_STAGES;
                    mbar_wait(smem_u32(&sm->b_empty[s]), ((b_it / B_STAGES) & 1u) ^ 1u);
                    mbar_expect_tx(smem_u32(&sm->b_full[s]), B_STAGE_BYTES);
                    bulk_g2s(smem_u32(sB + s * B_STAGE_BYTES), reinterpret_cast<const uint8_t *>(Bimg) + (size_t)c * B_STAGE_BYTES,
                             B_STAGE_BYTES, smem_u32(&sm->b_full[s]));
                    ++b_it;
                }
                pos += sg.c_hi - sg.c_lo;
            }
        }
    } else if (warp >= WARP_MMA + NMMA) {
        TCS_REGS_LESS();  // (LR_TCS_WIDE: warps that only complete the last warpgroup)
    } else if (warp >= WARP_MMA) {
        TCS_REGS_LESS();
        // ===== MMA issuers: warp WARP_MMA + j issues sub-tile j of every stage into TMEM buffer j.  Warp-uniform loops,
        // elect.sync picks the lane; unrolled over the four stages of the B ring, so stage, barrier addresses, parities
        // and descriptors are static.  One issuing warp was the sweep's bottleneck (tools/tcs_trace.py,
        // profiles/r2_tcs_trace_*.txt): every barrier operation of a warp costs ~100 cycles of latency, and a stage
        // needs a b_full wait + per tile (t_empty wait, issue, commit) -- ~1100 cycles per stage in series whatever the
        // tile width.  One warp per tile runs those chains side by side.
        static_assert(B_STAGES == 4, "the issue loop is unrolled over the 4 stages of the B ring");
        const int j = warp - WARP_MMA;
        if (rg.pos < rg.end) {
            const uint32_t a_full = smem_u32(&sm->a_full[0]), a_empty = smem_u32(&sm->a_empty[0]);
            const uint32_t b_full = smem_u32(&sm->b_full[0]), b_empty = smem_u32(&sm->b_empty[0]);
#if LR_TCS_WIDE
            const uint32_t t_full0 = smem_u32(&sm->t_full[3 * j]), t_empty = smem_u32(&sm->t_empty[j]);
            uint32_t r3 = 0u;  // stages issued so far, modulo 3
#else
            const uint32_t t_full = smem_u32(&sm->t_full[j]), t_empty = smem_u32(&sm->t_empty[j]);
#endif
            const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB) + j * (TN / 8) * B_GROUP_BYTES;
            const uint32_t d_tmem = tmem_base + j * TMEM_BUF_COLS;
            // B: K-adjacent cores core0 -> core1_a are 128 (1 + a) bytes apart, 8-correspondence groups 512 bytes apart
            const uint64_t bd0 = smem_desc(b_addr, 128, B_GROUP_BYTES), bd1 = smem_desc(b_addr, 256, B_GROUP_BYTES),
                           bd2 = smem_desc(b_addr, 384, B_GROUP_BYTES);
            uint64_t ad0 = smem_desc(a_addr, 128, 256), ad1 = smem_desc(a_addr + A_PART_BYTES, 128, 256),
                     ad2 = smem_desc(a_addr + 2 * A_PART_BYTES, 128, 256);
            uint32_t a_it = 0, pb = 0;
            long long pos = rg.pos;
            Seg sg = seg_at(pos, rg.end, nchunks);
            int left = sg.c_hi - sg.c_lo;
#ifdef LR_TCS_TRACE
            int st_no = 0;
#endif
            mbar_wait_mma(a_full, 0u);
            bool done = false;
            while (!done) {
#pragma unroll
                for (int s = 0; s < B_STAGES; ++s) {
                    // descriptors of this stage: the 14-bit address field counts 16-byte units, a plain add
                    const uint64_t off = (uint64_t)((s * B_STAGE_BYTES) >> 4);
                    const uint64_t x0 = bd0 + off, x1 = bd1 + off, x2 = bd2 + off;
                    TCS_TRACE(st_no, j, 0);
                    mbar_wait_mma(b_full + s * 8u, pb);
                    // TMEM buffer j is free once its readers hold the previous stage's sub-tile j in registers
                    mbar_wait_mma(t_empty, (uint32_t)(s & 1) ^ 1u);
                    TCS_TRACE(st_no, j, 1);
                    tc_fence_after();
#if LR_TCS_WIDE
                    tc_issue_tile(d_tmem, ad0, ad1, ad2, x0, x1, x2, t_full0 + r3 * 8u);
                    r3 = r3 == 2u ? 0u : r3 + 1u;
#else
                    tc_issue_tile(d_tmem, ad0, ad1, ad2, x0, x1, x2, t_full);
#endif
                    tc_commit_elect(b_empty + s * 8u);  // the stage's bytes are reusable once the MMAs of all tiles have read them
                    TCS_TRACE(st_no, j, 2);
#ifdef LR_TCS_TRACE
                    ++st_no;
#endif
                    if (--left == 0) {
                        tc_commit_elect(a_empty + (a_it & 1u) * 8u);
                        ++a_it;
                        pos += sg.c_hi - sg.c_lo;
                        if (pos >= rg.end) {
                            done = true;
                            break;
                        }
                        sg = seg_at(pos, rg.end, nchunks);
                        left = sg.c_hi - sg.c_lo;
                        mbar_wait_mma(a_full + (a_it & 1u) * 8u, (a_it >> 1) & 1u);
                        const uint32_t aa = a_addr + (a_it & 1u) * A_BLOCK_BYTES;
                        ad0 = smem_desc(aa, 128, 256);
                        ad1 = smem_desc(aa + A_PART_BYTES, 128, 256);
                        ad2 = smem_desc(aa + 2 * A_PART_BYTES, 128, 256);
                    }
                }
                pb ^= 1u;
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> counts =====
        TCS_REGS_MORE();
#if LR_TCS_WIDE
        {
            const int q = warp & 3;    // TMEM lane quadrant this warp may read
            const int grp = warp >> 2; // this group's tiles: t = grp, grp + 3, ... of the CTA's sequence (4 tiles per stage)
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
            const uint32_t t_full = smem_u32(&sm->t_full[0]), t_empty = smem_u32(&sm->t_empty[0]);
            const float nthr2 = -(float)thr2;
            const u64 nthr2p = pk2(__float_as_uint(nthr2), __float_as_uint(nthr2));
            unsigned long long *n_rechecked = reinterpret_cast<unsigned long long *>(&ctl->n_rechecked);
            const int region = blockIdx.x * NEPI + warp;
            int4 *ev_mine = events + (size_t)region * ev_cap;
            unsigned ev_n = 0, evals_tail = 0;
            const long long ntiles = (rg.end - rg.pos) * TILES_PER_STAGE;
            // position of the current tile: stage s of the CTA's range (its parity is the barrier phase: every buffer is
            // used once per stage), sub-tile j, hypothesis block hb, correspondence stage c
            long long s = 0;
            uint32_t sr = 0u, sk = 0u;  // s mod 3 and the parity of s / 3: barrier (j, sr) is in its (s / 3)-th phase
            int j = grp;  // grp < 4: the first tile of the group lies in stage 0
            int hb = (int)(rg.pos / nchunks), c = (int)(rg.pos - (long long)hb * nchunks);
            int slot = hb * TM + q * 32 + lane;
            bool valid = slot < nsurv;
            float delta = valid ? band[slot] : -1.f;  // rows past the survivor count hold stale models
            int count = 0, count_b = 0;
            unsigned evals = 0;
            for (long long t = grp; t < ntiles; t += NGROUP) {
                TCS_TRACE((uint32_t)s, j, 0);
                mbar_wait(t_full + (uint32_t)(3 * j + (int)sr) * 8u, sk);
                TCS_TRACE((uint32_t)s, j, 1);
                tc_fence_after();
                uint32_t d0[32], d1[32], d2[32];
                tmem_ld32_issue(taddr + j * TMEM_BUF_COLS, d0);
                tmem_ld32_issue(taddr + j * TMEM_BUF_COLS + TN, d1);
                tmem_ld32_issue(taddr + j * TMEM_BUF_COLS + 2 * TN, d2);
                tmem_ld_wait3x32(d0, d1, d2);
                // the tile is in this group's registers: hand the TMEM buffer back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(t_empty + j * 8u);
                TCS_TRACE((uint32_t)s, j, 2);
                const int64_t col0 = (int64_t)c * BROWS + j * TN;
                if (DUMP) {
                    if (valid) {
#pragma unroll
                        for (int k = 0; k < 32; ++k) {
                            float *o = dump + ((size_t)slot * n_pad + (col0 + k)) * 3;
                            o[0] = __uint_as_float(d0[k]);
                            o[1] = __uint_as_float(d1[k]);
                            o[2] = __uint_as_float(d2[k]);
                        }
                    }
                } else {
                    u64 u[16];
                    float mn = INFINITY;
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const u64 e0 = pk2(d0[2 * k], d0[2 * k + 1]), e1 = pk2(d1[2 * k], d1[2 * k + 1]),
                                  e2 = pk2(d2[2 * k], d2[2 * k + 1]);
                        u[k] = fma2(e2, e2, fma2(e1, e1, fma2(e0, e0, nthr2p)));
                        float ua, ub;
                        upk2(u[k], ua, ub);
                        count += (int)(__float_as_uint(ua) >> 31);      // two chains: the LEA.HI adds of a tile do not
                        count_b += (int)(__float_as_uint(ub) >> 31);    // form one long dependency
                        mn = min3abs(mn, ua, ub);
                    }
                    if (__any_sync(0xffffffffu, mn < delta)) {
                        // rare: some residual of these 32 columns is within the error band of some hypothesis: the warp
                        // records which (bit k: |u_k| < delta) and what the tensor-core sign said; decided in fp64 after
                        // the warp's last tile (see the unwidened loop below for the why)
                        unsigned bm = 0u, sg_bits = 0u;
#pragma unroll
                        for (int k = 0; k < 32; ++k) {
                            float ua, ub;
                            upk2(u[k >> 1], ua, ub);
                            const float uk = (k & 1) ? ub : ua;
                            bm |= (fabsf(uk) < delta ? 1u : 0u) << k;
                            sg_bits |= (__float_as_uint(uk) >> 31) << k;
                        }
                        const unsigned m = __ballot_sync(0xffffffffu, bm != 0u);
                        if (bm != 0u) {
                            const unsigned e = ev_n + __popc(m & ((1u << lane) - 1u));
                            if (e < ev_cap) ev_mine[e] = make_int4(slot, (int)col0, (int)bm, (int)sg_bits);
                            else {  // list full: decide on the spot (same result, only slower)
                                for (unsigned b = bm; b; b &= b - 1u) {
                                    const int k = __ffs(b) - 1;
                                    count += tc_exact_inlier(P8, col0 + k, m64 + (size_t)slot * 12, thr2) - (int)((sg_bits >> k) & 1u);
                                    ++evals;
                                }
                            }
                        }
                        ev_n += __popc(m);
                    }
                }
                TCS_TRACE((uint32_t)s, j, 3);
                // advance three tiles: at most one stage further
                j += NGROUP;
                if (j >= TILES_PER_STAGE) {
                    j -= TILES_PER_STAGE;
                    ++s;
                    if (++sr == 3u) {
                        sr = 0u;
                        sk ^= 1u;
                    }
                    if (++c == nchunks) {  // the next tile belongs to the next hypothesis block: flush this one's counts
                        if (!DUMP && valid) {
                            count += count_b;
                            if (count) atomicAdd(&cnt[slot], count);
                            if (evals) atomicAdd(n_rechecked, (unsigned long long)evals);
                        }
                        count = count_b = 0;
                        evals = 0;
                        c = 0;
                        ++hb;
                        slot = hb * TM + q * 32 + lane;
                        valid = slot < nsurv;
                        delta = valid ? band[slot] : -1.f;
                    }
                }
            }
            if (!DUMP && valid) {
                count += count_b;
                if (count) atomicAdd(&cnt[slot], count);
                if (evals) atomicAdd(n_rechecked, (unsigned long long)evals);
            }
            if (!DUMP && ev_n) {
                __syncwarp();
                const unsigned ne = ev_n < ev_cap ? ev_n : ev_cap;
                for (unsigned e = lane; e < ne; e += 32) {
                    const int4 v = __ldcg(ev_mine + e);
                    int d = 0;
                    for (unsigned b = (unsigned)v.z; b; b &= b - 1u) {
                        const int k = __ffs(b) - 1;
                        d += tc_exact_inlier(P8, (int64_t)v.y + k, m64 + (size_t)v.x * 12, thr2) - (int)(((unsigned)v.w >> k) & 1u);
                        ++evals_tail;
                    }
                    if (d) atomicAdd(&cnt[v.x], d);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) evals_tail += __shfl_xor_sync(0xffffffffu, evals_tail, o);
                if (lane == 0 && evals_tail) atomicAdd(n_rechecked, (unsigned long long)evals_tail);
            }
        }
#else
        // (Tried and dropped: 8-column units whose registers ping-pong so that the next tcgen05.ld is in flight while
        // the current unit is counted, with an early non-blocking probe of the next t_full phase -- 0.57 ms instead of
        // 0.36 ms at cfg 3: the longer serial code per slice cost more than the hidden latencies saved.)
        const int q = warp & 3;                  // TMEM lane quadrant this warp may read
        const int grp = (warp >> 2) % NGROUP;    // which tiles of a stage
        const int cs = (warp >> 2) / NGROUP;     // which 16 of a tile's columns
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + cs * 16;
        const uint32_t t_full = smem_u32(&sm->t_full[0]), t_empty = smem_u32(&sm->t_empty[0]);
        const float nthr2 = -(float)thr2;
        const u64 nthr2p = pk2(__float_as_uint(nthr2), __float_as_uint(nthr2));
        unsigned long long *n_rechecked = reinterpret_cast<unsigned long long *>(&ctl->n_rechecked);
        uint32_t t_it = 0;  // stages consumed (each stage = one use of each TMEM buffer)
        // this warp's private list of in-band residuals (slot, correspondence)
        const int region = blockIdx.x * NEPI + warp;
        int4 *ev_mine = events + (size_t)region * ev_cap;
        unsigned ev_n = 0, evals_tail = 0;
        for (long long pos = rg.pos; pos < rg.end;) {
            const Seg sg = seg_at(pos, rg.end, nchunks);
            const int slot = sg.hb * TM + q * 32 + lane;
            const bool valid = slot < nsurv;
            const float delta = valid ? band[slot] : -1.f;  // rows past the survivor count hold stale models
            int count = 0, count_b = 0;
            unsigned evals = 0;
            for (int c = sg.c_lo; c < sg.c_hi; ++c) {
#pragma unroll
                for (int jj = 0; jj < TILES_PER_STAGE / NGROUP; ++jj) {
                    const int j = grp + jj * NGROUP;
                    TCS_TRACE(t_it, j, 0);
                    mbar_wait(t_full + j * 8u, t_it & 1u);
                    TCS_TRACE(t_it, j, 1);
                    tc_fence_after();
                    const int64_t col0 = (int64_t)c * BROWS + j * TN + cs * 16;
                    uint32_t d0[16], d1[16], d2[16];
                    tmem_ld16_issue(taddr + j * TMEM_BUF_COLS, d0);
                    tmem_ld16_issue(taddr + j * TMEM_BUF_COLS + TN, d1);
                    tmem_ld16_issue(taddr + j * TMEM_BUF_COLS + 2 * TN, d2);
                    tmem_ld_wait3(d0, d1, d2);
                    // this warp's share of the tile is in registers: hand the TMEM buffer back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(t_empty + j * 8u);
                    TCS_TRACE(t_it, j, 2);
                    if (DUMP) {
                        if (valid) {
#pragma unroll
                            for (int k = 0; k < 16; ++k) {
                                float *o = dump + ((size_t)slot * n_pad + (col0 + k)) * 3;
                                o[0] = __uint_as_float(d0[k]);
                                o[1] = __uint_as_float(d1[k]);
                                o[2] = __uint_as_float(d2[k]);
                            }
                        }
                        continue;
                    }
                    u64 u[8];
                    float mn = INFINITY;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const u64 e0 = pk2(d0[2 * k], d0[2 * k + 1]), e1 = pk2(d1[2 * k], d1[2 * k + 1]),
                                  e2 = pk2(d2[2 * k], d2[2 * k + 1]);
                        u[k] = fma2(e2, e2, fma2(e1, e1, fma2(e0, e0, nthr2p)));
                        float ua, ub;
                        upk2(u[k], ua, ub);
                        count += (int)(__float_as_uint(ua) >> 31);      // two chains: the LEA.HI adds of a tile do not
                        count_b += (int)(__float_as_uint(ub) >> 31);    // form one 16-long dependency
                        mn = min3abs(mn, ua, ub);
                    }
                    if (__any_sync(0xffffffffu, mn < delta)) {
                        // rare: some residual of these 16 columns is within the error band of some hypothesis.  The warp
                        // only RECORDS which (bit k: |u_k| < delta) and what the tensor-core sign said, one 16-byte
                        // record per affected lane in the warp's private list -- no loads, no atomics, no divergent
                        // loop: a warp that is late for its next tile stalls every warp behind the same TMEM buffer
                        // (58 % of the tiles had such a straggler at cfg 3 when this path decided in fp64 on the spot).
                        // The warp decides its listed residuals in fp64 after its last tile (end of this role).
                        unsigned bm = 0u, sg_bits = 0u;
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            float ua, ub;
                            upk2(u[k >> 1], ua, ub);
                            const float uk = (k & 1) ? ub : ua;
                            bm |= (fabsf(uk) < delta ? 1u : 0u) << k;
                            sg_bits |= (__float_as_uint(uk) >> 31) << k;
                        }
                        const unsigned m = __ballot_sync(0xffffffffu, bm != 0u);
                        if (bm != 0u) {
                            const unsigned e = ev_n + __popc(m & ((1u << lane) - 1u));
                            if (e < ev_cap) ev_mine[e] = make_int4(slot, (int)col0, (int)bm, (int)sg_bits);
                            else {  // list full: decide on the spot (same result, only slower)
                                for (unsigned b = bm; b; b &= b - 1u) {
                                    const int k = __ffs(b) - 1;
                                    count += tc_exact_inlier(P8, col0 + k, m64 + (size_t)slot * 12, thr2) - (int)((sg_bits >> k) & 1u);
                                    ++evals;
                                }
                            }
                        }
                        ev_n += __popc(m);
                    }
                    TCS_TRACE(t_it, j, 3);
                }
                ++t_it;
            }
            if (!DUMP && valid) {
                // the column slices of a quadrant (and the CTAs that share a hypothesis block) hold partial counts of the
                // same slot: always merge (k_kabsch / k_probe_install zero cnt[] before the sweep)
                count += count_b;
                if (count) atomicAdd(&cnt[slot], count);
                if (evals) atomicAdd(n_rechecked, (unsigned long long)evals);
            }
            pos += sg.c_hi - sg.c_lo;
        }
        if (!DUMP && ev_n) {
            // this warp's own list, decided with the canonical fp64 arithmetic: the count receives (exact - sign).
            // (A separate kernel did this first; as the tail of the sweep it costs one launch less per round.)
            __syncwarp();
            const unsigned ne = ev_n < ev_cap ? ev_n : ev_cap;
            for (unsigned e = lane; e < ne; e += 32) {
                const int4 v = __ldcg(ev_mine + e);
                int d = 0;
                for (unsigned b = (unsigned)v.z; b; b &= b - 1u) {
                    const int k = __ffs(b) - 1;
                    d += tc_exact_inlier(P8, (int64_t)v.y + k, m64 + (size_t)v.x * 12, thr2) - (int)(((unsigned)v.w >> k) & 1u);
                    ++evals_tail;
                }
                if (d) atomicAdd(&cnt[v.x], d);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) evals_tail += __shfl_xor_sync(0xffffffffu, evals_tail, o);
            if (lane == 0 && evals_tail) atomicAdd(n_rechecked, (unsigned long long)evals_tail);
        }
#endif
    }

    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

constexpr size_t kSmemBytes = 2 * A_BLOCK_BYTES + B_STAGES * B_STAGE_BYTES + sizeof(Smem) + 1024 + 64;

}  // namespace tcs
