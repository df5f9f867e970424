// lr_ransac.cu -- batched RANSAC rigid-motion estimation for sm_100a.
//
// Replaces (reference tree citations):
//   pygcransac.findRigidTransform        Experiments/algorithms/GC_RANSAC.py:46-49,
//                                        GC-RANSAC/src/pygcransac/src/gcransac_python.cpp:404-624
//   o3d registration_ransac_based_on_correspondence   Experiments/algorithms/FR.py:122-139
//   EdgeLenPreemptiveVerification::verifyModel
//                                        GC-RANSAC/src/pygcransac/include/preemption/preemption_edge_length.h:71-128
//   inlier refit                         Experiments/algorithms/FR.py:99-111
//
// Structure of one batch of hypotheses (DESIGN.md "RANSAC kernels"):
//   k_gen    one thread per hypothesis id: counter-based sample, edge-length
//            test in fp64; survivors are compacted (warp-aggregated)
//   k_kabsch dense over survivors: fixed-sweep Jacobi Kabsch in fp64 registers,
//            fp32 copy of [R|t] + the rigorous fp32 error band of the inlier test
//   k_score  thread-owns-two-hypotheses sweep (packed f32x2 FMAs) over all
//            correspondences staged through shared memory with cp.async;
//            warp-uniform early-out on the first residual component;
//            residuals inside the error band are decided in fp64 on the spot,
//            so every count is exact
//   k_resolve / k_round_end  packed (count, id) arg-max, confidence exit flag
// fp64 arithmetic follows oracle/lr_oracle.c operation for operation (this
// file is compiled with -fmad=false; FMAs are written explicitly where wanted).
#include <cuda_fp16.h>
#include <math.h>

#include <vector>

#include "lr_common.cuh"

namespace {

constexpr int kScoreThreads = 32;   // threads per score CTA (ONE warp: warps with different early-out rates never wait for
                                    // each other at a barrier); each thread owns TWO hypotheses (packed f32x2)
constexpr int kHypPerItem = 2 * kScoreThreads;
constexpr int kChunk = 128;         // correspondences per shared-memory stage (48 B each, double buffered)
constexpr int kGroup = 4;           // points whose first residual component is evaluated together (early-out votes)
constexpr int kBlock = 32;          // points between two checks of the "some residual is in the band" flag
constexpr int kGenThreads = 128;
constexpr int kGcMaxTrials = 64;    // inner LO draws scored by one launch
int g_dbg_rank = 0, g_dbg_world = 1;  // lr_debug_slice
int g_score_mode = 0;               // lr_ransac_set_mode: 0 = tensor-core sweep, 1 = fp32 sweep in full, 2 = fp32 sweep with
                                    // the first-component early-out (A/B)

struct Ctl {
    unsigned long long best_key;   // over finished rounds
    unsigned long long round_key;  // current round
    int n_surv;                    // survivors of the current round (compacted slots)
    unsigned int n_events;         // tensor sweep: in-band residuals listed for k_tc_events in the current round
    int done;                      // confidence exit reached
    int pad0;
    long long iters_run, n_scored, n_rechecked;
    unsigned int p1max_bits, qmax_bits;  // max |p|_1, max |q|_inf as float bits
    unsigned int pt2max_bits, qtmax_bits;  // max |p - c|_2, max |q - c'|_inf (lr_score_tc.cuh: operand frame)
    long long refit_count;
    double T[12], Tref[12];
    double csum[6];
    double H[9];
    double err2;  // sum of squared residuals over the inliers (ICP rmse)
    unsigned int ticket_end, ticket_fin;  // "last block finishes" counters of k_resolve_end / k_finish
    int comm_error;                       // hypothesis sharding: a peer did not answer in time
    unsigned int ticket_pack;             // k_pack's counter: zero whenever no k_pack is running (its last block re-zeroes
                                          // it; a fresh arena block is zero-filled), never touched by ctl_reset_fields
    // LR_SCORE_MSAC runs only (lr_ransac_gc.cuh); q = quantised MSAC score (include/lidarreg.h)
    struct Gc {
        unsigned long long round_q;     // highest q of the current round
        unsigned long long round_pick;  // lowest (id << 32 | slot) among the slots that reach it
        unsigned long long best_q;      // q of the selected minimal-sample hypothesis
        unsigned long long cur_q;       // q of cur[] (grows through LO and the iterated least squares)
        unsigned long long lo_q;        // cur_q when the local optimisation ended
        long long best_id, best_inl;    // selected hypothesis, its #(r^2 < tau^2)
        int has_model, lo_active, lsq_active, lo_improved, lsq_improved;
        int lo_I, lo_s;                 // inliers (at thr) of cur[] and the inner sample size of this LO round
        int mode;                       // 1 while the block belongs to an LR_SCORE_MSAC run
        double cur[12];
    } gc;
};

struct Ws {
    Ctl *ctl;
    float4 *P12;      // 3 x float4 per point: (px,px,py,py) (pz,pz,-qx,-qx) (-qy,-qy,-qz,-qz)
    float4 *P8;       // 2 x float4 per point: (px,py,pz,qx) (qy,qz,0,0) -- one 32-byte sector per sampled correspondence
    int32_t *samp;    // sample indices of the survivors, 4 per slot
    uint32_t *slot_id;
    float4 *m32;      // fp32 [R|t], lo, hi of two slots interleaved: see m32_index()
    double *m64;      // 12 per slot
    int *cnt;         // per slot: exact inlier count, accumulated over point splits
    int *need;        // per round: inlier count that triggers the confidence exit
    uint32_t *growth; // PROSAC growth function T'_n (null unless the sampler is PROSAC)
    double *scratchT; // 16 doubles of staging
    int64_t n_pad;
    // tensor-core sweep (lr_score_tc.cuh)
    uint4 *Aimg;      // fp16 operand image of the surviving hypotheses, 12 KB per 128 slots
    uint4 *Bimg;      // fp16 operand image of the correspondences, 64 B each
    float *band;      // per slot: half-width of the r^2 band that is decided in fp64
    double *partial;  // k_finish: kFinBlocksMax x kFinVals per-block sums
    int4 *events;     // records of the residuals inside the tensor sweep's error band, one list per epilogue warp and CTA
    unsigned long long *blockbest;  // k_resolve_end: per-block (key, slot)
    float *packmax;   // k_pack: per-block maxima (|p|_1, |p - c|_2, |q - c'|_inf), reduced by its last block
    float *stage;     // lr_ransac_rigid_batch: device copy of a pair that was handed over in HOST memory (2 x 3 n floats)
    // LR_SCORE_MSAC runs only (null otherwise)
    unsigned long long *q64;  // per slot: quantised MSAC score
    int32_t *lo_L;            // inlier index list of the current LO round, ascending
    double *tr_T;             // kGcMaxTrials x 12: models of one LO round / the least-squares candidate
    unsigned long long *tr_q; // their q
    int *tr_inl;              // their #(r^2 < tau^2)
};

// ------------------------------------------------------------------------
// canonical fp64 device arithmetic (mirrors oracle/lr_oracle.c)
// ------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

__device__ __forceinline__ uint32_t draw(uint64_t seed, uint64_t id, uint32_t d, uint32_t m)
{
    uint64_t r = mix64(mix64(seed ^ (id * 0xD1342543DE82EF95ULL)) + (uint64_t)d * 0x9E3779B97F4A7C15ULL);
    return (uint32_t)(((r >> 32) * (uint64_t)m) >> 32);
}

constexpr uint64_t kProsacTN = 100000;  // ProsacSampler(points, m, T_N = 100 000), SURVEY App. A

// K unique indices out of [0, n): draw d picks the r-th index not yet taken
template <int K>
__device__ __forceinline__ void unique_ids(uint64_t seed, uint64_t id, int64_t n, int32_t *out)
{
    int32_t taken[K > 0 ? K : 1];
#pragma unroll
    for (int d = 0; d < K; ++d) {
        int32_t r = (int32_t)draw(seed, id, d, (uint32_t)(n - d));
#pragma unroll
        for (int e = 0; e < K; ++e)
            if (e < d && r >= taken[e]) ++r;
        out[d] = r;
        taken[d] = r;
#pragma unroll
        for (int e = K - 1; e > 0; --e)
            if (e <= d && taken[e - 1] > taken[e]) {
                int32_t tmp = taken[e];
                taken[e] = taken[e - 1];
                taken[e - 1] = tmp;
            }
    }
}

// sampler ids of include/lidarreg.h (== oracle lro_sample): uniform-unique, PROSAC, with replacement
template <int M>
__device__ __forceinline__ void sample_ids(uint64_t seed, uint64_t id, int sampler, int64_t n,
                                           const uint32_t *__restrict__ growth, int32_t (&out)[M])
{
    if (sampler == LR_SAMPLER_REPLACE) {
#pragma unroll
        for (int d = 0; d < M; ++d) out[d] = (int32_t)draw(seed, id, d, (uint32_t)n);
        return;
    }
    if (sampler == LR_SAMPLER_PROSAC && growth != nullptr && id + 1 <= kProsacTN) {
        // draw k = id + 1 uses the n_k best correspondences, n_k = smallest n in [M, N] with T'_n >= k;
        // M - 1 unique indices below n_k - 1 plus correspondence n_k - 1 itself
        const uint64_t k = id + 1;
        int64_t lo = M, hi = n;
        while (lo < hi) {
            const int64_t mid = lo + (hi - lo) / 2;
            if ((uint64_t)growth[mid - 1] >= k) hi = mid;
            else lo = mid + 1;
        }
        unique_ids<M - 1>(seed, id, lo - 1, out);
        out[M - 1] = (int32_t)(lo - 1);
        return;
    }
    unique_ids<M>(seed, id, n, out);
}

__device__ __forceinline__ double len3(const double *a, const double *b)
{
    double dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
    return sqrt((dx * dx + dy * dy) + dz * dz);
}

template <int M>
__device__ __forceinline__ bool elc_pass(const double (&P)[M][3], const double (&Q)[M][3], double ratio)
{
    bool ok = true;
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = i + 1; j < M; ++j) {
            double ds = len3(P[i], P[j]);
            double dt = len3(Q[i], Q[j]);
            if ((ds < dt * ratio) || (dt < ds * ratio)) ok = false;
        }
    return ok;
}

__device__ __forceinline__ double coldot(const double (&H)[3][3], int p, int q)
{
    return (H[0][p] * H[0][q] + H[1][p] * H[1][q]) + H[2][p] * H[2][q];
}

__device__ __forceinline__ void colswap(double (&A)[3][3], int p, int q)
{
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double t = A[k][p];
        A[k][p] = A[k][q];
        A[k][q] = t;
    }
}

// H = sum (q - cq)(p - cp)^T  ->  proper rotation (see oracle lro_rot_from_H)
__device__ void rot_from_H(const double (&Hin)[3][3], double (&R)[3][3])
{
    double H[3][3], V[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            H[r][c] = Hin[r][c];
            V[r][c] = (r == c) ? 1.0 : 0.0;
        }
#pragma unroll 1
    for (int sweep = 0; sweep < 6; ++sweep) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int p = (r == 2) ? 1 : 0;
            const int q = (r == 0) ? 1 : 2;
            double alpha = coldot(H, p, p);
            double beta = coldot(H, q, q);
            double gamma = coldot(H, p, q);
            double c = 1.0, s = 0.0;
            if (gamma != 0.0) {
                double zeta = (beta - alpha) / (2.0 * gamma);
                double az = fabs(zeta);
                double tt = 1.0 / (az + sqrt(1.0 + zeta * zeta));
                if (zeta < 0.0) tt = -tt;
                c = 1.0 / sqrt(1.0 + tt * tt);
                s = c * tt;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double hp = H[k][p], hq = H[k][q];
                H[k][p] = c * hp - s * hq;
                H[k][q] = s * hp + c * hq;
                double vp = V[k][p], vq = V[k][q];
                V[k][p] = c * vp - s * vq;
                V[k][q] = s * vp + c * vq;
            }
        }
    }
    double s0 = coldot(H, 0, 0), s1 = coldot(H, 1, 1), s2 = coldot(H, 2, 2);
    // stable descending sort of the three columns (== oracle's first-argmax picks)
    if (s1 > s0) { colswap(H, 0, 1); colswap(V, 0, 1); double t = s0; s0 = s1; s1 = t; }
    if (s2 > s1) { colswap(H, 1, 2); colswap(V, 1, 2); double t = s1; s1 = s2; s2 = t; }
    if (s1 > s0) { colswap(H, 0, 1); colswap(V, 0, 1); double t = s0; s0 = s1; s1 = t; }
    if (!(s0 > 0.0)) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) R[r][c] = (r == c) ? 1.0 : 0.0;
        return;
    }
    double u1[3], u2[3], u3[3], v1[3], v2[3], v3[3];
    double sa = sqrt(s0);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        u1[k] = H[k][0] / sa;
        v1[k] = V[k][0];
        v2[k] = V[k][1];
    }
    if (s1 > s0 * 1e-30) {
        double sb = sqrt(s1);
#pragma unroll
        for (int k = 0; k < 3; ++k) u2[k] = H[k][1] / sb;
    } else {
        int e = 0;
        double m = fabs(u1[0]);
        if (fabs(u1[1]) < m) { e = 1; m = fabs(u1[1]); }
        if (fabs(u1[2]) < m) { e = 2; }
#pragma unroll
        for (int k = 0; k < 3; ++k) u2[k] = (k == e) ? 1.0 : 0.0;
    }
    double g = (u1[0] * u2[0] + u1[1] * u2[1]) + u1[2] * u2[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) u2[k] = u2[k] - g * u1[k];
    double nu = sqrt((u2[0] * u2[0] + u2[1] * u2[1]) + u2[2] * u2[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) u2[k] = u2[k] / nu;
    u3[0] = u1[1] * u2[2] - u1[2] * u2[1];
    u3[1] = u1[2] * u2[0] - u1[0] * u2[2];
    u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
    v3[0] = v1[1] * v2[2] - v1[2] * v2[1];
    v3[1] = v1[2] * v2[0] - v1[0] * v2[2];
    v3[2] = v1[0] * v2[1] - v1[1] * v2[0];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) R[r][c] = (u1[r] * v1[c] + u2[r] * v2[c]) + u3[r] * v3[c];
}

__device__ __forceinline__ void finish_T(const double (&R)[3][3], const double (&cp)[3], const double (&cq)[3],
                                         double (&T)[12])
{
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        T[4 * r + 0] = R[r][0];
        T[4 * r + 1] = R[r][1];
        T[4 * r + 2] = R[r][2];
        T[4 * r + 3] = cq[r] - ((R[r][0] * cp[0] + R[r][1] * cp[1]) + R[r][2] * cp[2]);
    }
}

template <int M>
__device__ void kabsch_small(const double (&P)[M][3], const double (&Q)[M][3], double (&T)[12])
{
    double cp[3] = {0, 0, 0}, cq[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            cp[c] = cp[c] + P[i][c];
            cq[c] = cq[c] + Q[i][c];
        }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        cp[c] = cp[c] / (double)M;
        cq[c] = cq[c] / (double)M;
    }
    double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int i = 0; i < M; ++i) {
        double dp[3], dq[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            dp[c] = P[i][c] - cp[c];
            dq[c] = Q[i][c] - cq[c];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) H[r][c] = H[r][c] + dq[r] * dp[c];
    }
    double R[3][3];
    rot_from_H(H, R);
    finish_T(R, cp, cq, T);
}

__device__ __forceinline__ double res2_f64(const double *T, double px, double py, double pz, double qx, double qy,
                                           double qz)
{
    double d0 = (((T[0] * px + T[1] * py) + T[2] * pz) + T[3]) - qx;
    double d1 = (((T[4] * px + T[5] * py) + T[6] * pz) + T[7]) - qy;
    double d2 = (((T[8] * px + T[9] * py) + T[10] * pz) + T[11]) - qz;
    return (d0 * d0 + d1 * d1) + d2 * d2;
}

__device__ __forceinline__ unsigned long long make_key(int count, uint32_t id)
{
    return ((unsigned long long)(uint32_t)(count + 1) << 32) | (unsigned long long)(0xFFFFFFFFu - id);
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    return v;
}

// one sampled correspondence = one 32-byte sector of the packed copy (the [n,3] arrays cost 2-4 sectors)
__device__ __forceinline__ void load_pq(const float4 *__restrict__ P8, int64_t k, double (&P)[3], double (&Q)[3])
{
    const float4 a = __ldg(P8 + 2 * k), b = __ldg(P8 + 2 * k + 1);
    P[0] = (double)a.x, P[1] = (double)a.y, P[2] = (double)a.z;
    Q[0] = (double)a.w, Q[1] = (double)b.x, Q[2] = (double)b.y;
}

typedef unsigned long long u64;

__device__ __forceinline__ void upk2(u64 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// same instruction, but `volatile`: ptxas keeps these in source order, which is written so that
// three consecutive FMAs share their B operand (register-reuse cache; tools/micro_ffma2.cu:
// 100 instead of 87 FMA/clk/SM)
__device__ __forceinline__ u64 fma2v(u64 a, u64 b, u64 c)
{
    u64 d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b)
{
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b)
{
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

#include "lr_score_tc.cuh"

// ------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------

// AoS [n,3] x2  ->  3 x float4 per correspondence with every value duplicated,
// so one LDS.128 yields two ready-made f32x2 operands (the target is stored
// negated: the sweep then only adds).  Padded to a multiple of kChunk with
// far-away points that can never be inliers.  Also the coordinate bound that
// enters the fp32 error band.
__device__ void ctl_reset_fields(Ctl *ctl)
{
    ctl->best_key = 0ULL;
    ctl->round_key = 0ULL;
    ctl->n_surv = 0;
    ctl->n_events = 0u;
    ctl->done = 0;
    ctl->iters_run = 0;
    ctl->n_scored = 0;
    ctl->n_rechecked = 0;
    ctl->p1max_bits = 0u;
    ctl->qmax_bits = 0u;
    ctl->pt2max_bits = 0u;
    ctl->qtmax_bits = 0u;
    ctl->ticket_end = ctl->ticket_fin = 0u;
    ctl->comm_error = 0;
    ctl->refit_count = 0;
    ctl->gc.round_q = ctl->gc.best_q = ctl->gc.cur_q = ctl->gc.lo_q = 0ULL;
    ctl->gc.round_pick = ~0ULL;
    ctl->gc.best_id = -1;
    ctl->gc.best_inl = 0;
    ctl->gc.has_model = ctl->gc.lo_active = ctl->gc.lsq_active = 0;
    ctl->gc.lo_improved = ctl->gc.lsq_improved = 0;
    ctl->gc.lo_I = ctl->gc.lo_s = 0;
    ctl->gc.mode = 0;
    for (int k = 0; k < 12; ++k) {
        ctl->gc.cur[k] = (k % 5 == 0) ? 1.0 : 0.0;
        ctl->T[k] = (k % 5 == 0) ? 1.0 : 0.0;  // no selection (or a 0-inlier one) leaves the identity
    }
}

// The first kernel of a run.  Block 0 resets the control block (nothing else of this kernel reads or writes it until
// the last block), every block leaves its coordinate maxima in packmax[], and the last block to finish (ticket)
// reduces them into the control block -- one launch instead of reset + pack, and no same-address atomics
// (3 per warp serialised in L2: the pair of launches took 15 us at cfg 3).
// WANT12: also the duplicated-value copy the fp32 sweeps read (lr_ransac_set_mode 1 / 2, LR_SCORE_MSAC).
__global__ void __launch_bounds__(256)
k_pack(const float *__restrict__ src, const float *__restrict__ tgt, int64_t n, int64_t n_pad,
       float4 *__restrict__ P12, float4 *__restrict__ P8, uint4 *__restrict__ Bimg, Ctl *ctl, float *__restrict__ packmax,
       int want12)
{
    lr::pdl_wait();  // (the previous run's k_finish may still be copying this control block out)
    lr::pdl_launch();
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl_reset_fields(ctl);
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    float p1 = 0.f, pt2 = 0.f, qt = 0.f;
    double c[3] = {0, 0, 0}, cq[3] = {0, 0, 0};
    if (n > 0) tcs::tc_centre6(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(tgt), __ldg(tgt + 1), __ldg(tgt + 2), c, cq);
    if (i < n) {
        float px = src[3 * i], py = src[3 * i + 1], pz = src[3 * i + 2];
        float qx = tgt[3 * i], qy = tgt[3 * i + 1], qz = tgt[3 * i + 2];
        if (want12) {
            P12[3 * i + 0] = make_float4(px, px, py, py);
            P12[3 * i + 1] = make_float4(pz, pz, -qx, -qx);
            P12[3 * i + 2] = make_float4(-qy, -qy, -qz, -qz);
        }
        P8[2 * i + 0] = make_float4(px, py, pz, qx);
        P8[2 * i + 1] = make_float4(qy, qz, 0.f, 0.f);
        p1 = fabsf(px) + fabsf(py) + fabsf(pz);
        // the tensor-core sweep's operand image, in the frame (c, c'): differences of fp32 values are exact in fp64
        const double pt[3] = {(double)px - c[0], (double)py - c[1], (double)pz - c[2]};
        const double qq[3] = {(double)qx - cq[0], (double)qy - cq[1], (double)qz - cq[2]};
        tcs::tc_write_corr(Bimg, i, pt, qq);
        pt2 = __double2float_ru(sqrt((pt[0] * pt[0] + pt[1] * pt[1]) + pt[2] * pt[2])) * 1.000001f;
        qt = __double2float_ru(fmax(fabs(qq[0]), fmax(fabs(qq[1]), fabs(qq[2]))));
    } else if (i < n_pad) {
        if (want12) {
            P12[3 * i + 0] = make_float4(0.f, 0.f, 0.f, 0.f);
            P12[3 * i + 1] = make_float4(0.f, 0.f, -1e18f, -1e18f);
            P12[3 * i + 2] = make_float4(-1e18f, -1e18f, -1e18f, -1e18f);
        }
        // padding: 60 km away from wherever a model can send the origin (|t~| <= 2 x tcs::kRangeLimit)
        const double pt[3] = {0.0, 0.0, 0.0}, qq[3] = {60000.0, 60000.0, 60000.0};
        tcs::tc_write_corr(Bimg, i, pt, qq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        p1 = fmaxf(p1, __shfl_xor_sync(0xffffffffu, p1, o));
        pt2 = fmaxf(pt2, __shfl_xor_sync(0xffffffffu, pt2, o));
        qt = fmaxf(qt, __shfl_xor_sync(0xffffffffu, qt, o));
    }
    __shared__ float s_max[3][8];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) {
        s_max[0][w] = p1;
        s_max[1][w] = pt2;
        s_max[2][w] = qt;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float m = 0.f;
        for (int k = 0; k < 8; ++k) m = fmaxf(m, s_max[threadIdx.x][k]);
        packmax[3 * blockIdx.x + threadIdx.x] = m;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&ctl->ticket_pack, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (!s_last || threadIdx.x >= 32) return;
    __threadfence();
    float m0 = 0.f, m1 = 0.f, m2 = 0.f;
    for (int b = lane; b < (int)gridDim.x; b += 32) {
        m0 = fmaxf(m0, __ldcg(packmax + 3 * b));
        m1 = fmaxf(m1, __ldcg(packmax + 3 * b + 1));
        m2 = fmaxf(m2, __ldcg(packmax + 3 * b + 2));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
        m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    }
    if (lane == 0) {
        // round up a little: the fp32 sum of |p|_1 is itself rounded
        ctl->p1max_bits = __float_as_uint(m0 * 1.000001f);
        ctl->pt2max_bits = __float_as_uint(m1);
        ctl->qtmax_bits = __float_as_uint(m2);
        ctl->ticket_pack = 0u;
    }
}

__global__ void k_ctl_reset(Ctl *ctl)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) ctl_reset_fields(ctl);
}

// fp32 models are stored so that the sweep can load ready-made f32x2 operands:
// slots s and s + W of the same 2W-slot block (W = kScoreThreads) share 16 float2
// {r00 r01 r02 t0 | r10 r11 r12 t1 | r20 r21 r22 t2 | lo hi c -}, value v of slot s
// sits at float index ((s / 2W * W + s % W) * 16 + v) * 2 + (s % 2W) / W.
__device__ __forceinline__ size_t m32_index(int slot, int v)
{
    const int blk = slot / kHypPerItem, r = slot % kHypPerItem;
    return ((size_t)(blk * kScoreThreads + (r % kScoreThreads)) * 16 + v) * 2 + (r / kScoreThreads);
}

// edge-length test on squared lengths; falls back to the reference's
// sqrt form (preemption_edge_length.h:116-123) only when a comparison is within
// 1e-14 (relative) of equality, so the decision is always the reference's.
template <int M>
__device__ __forceinline__ bool elc_pass_fast(const double (&P)[M][3], const double (&Q)[M][3], double ratio)
{
    const double c2 = ratio * ratio;
    const double c_lo = c2 * (1.0 - 1e-14), c_hi = c2 * (1.0 + 1e-14);
    bool ok = true, unsure = false;
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = i + 1; j < M; ++j) {
            double ax = P[j][0] - P[i][0], ay = P[j][1] - P[i][1], az = P[j][2] - P[i][2];
            double bx = Q[j][0] - Q[i][0], by = Q[j][1] - Q[i][1], bz = Q[j][2] - Q[i][2];
            double S = (ax * ax + ay * ay) + az * az;
            double T = (bx * bx + by * by) + bz * bz;
            if (S < c_lo * T || T < c_lo * S) ok = false;               // certainly rejected
            else if (!(S > c_hi * T && T > c_hi * S)) unsure = true;    // too close to call on squares
        }
    if (ok && unsure) ok = elc_pass<M>(P, Q, ratio);
    return ok;
}

// one thread per hypothesis id: counter-based sample -> ELC; survivors are
// compacted (block-aggregated) as (id, sample indices)
template <int M>
__global__ void __launch_bounds__(kGenThreads)
k_gen(const float4 *__restrict__ P8, int64_t n, uint64_t seed, int sampler,
      int use_elc, double elc_ratio, int64_t id_lo, int64_t id_hi, const int32_t *__restrict__ fed,
      const uint32_t *__restrict__ growth, Ctl *ctl, uint32_t *__restrict__ slot_id, int32_t *__restrict__ samp)
{
    lr::pdl_wait();
    lr::pdl_launch();
    if (ctl->done) return;
    int64_t id = id_lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    bool ok = id < id_hi;
    int32_t s[M];
    if (ok) {
        if (fed) {
#pragma unroll
            for (int d = 0; d < M; ++d) s[d] = fed[(id - id_lo) * M + d];
        } else {
            sample_ids<M>(seed, (uint64_t)id, sampler, n, growth, s);
        }
        if (use_elc) {
            double P[M][3], Q[M][3];
#pragma unroll
            for (int d = 0; d < M; ++d) load_pq(P8, (int64_t)s[d], P[d], Q[d]);
            ok = elc_pass_fast<M>(P, Q, elc_ratio);
        }
    }
    // block-aggregated compaction: ONE atomic on the survivor counter per CTA (a warp-level atomic per warp with
    // survivors means ~30k same-address atomics per million hypotheses, which serialise in L2)
    __shared__ int s_wcnt[kGenThreads / 32];
    __shared__ int s_base;
    const unsigned ballot = __ballot_sync(0xffffffffu, ok);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_wcnt[w] = __popc(ballot);
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int k = 0; k < kGenThreads / 32; ++k) tot += s_wcnt[k];
        s_base = tot ? atomicAdd(&ctl->n_surv, tot) : 0;
    }
    __syncthreads();
    if (!ok) return;
    int base = s_base;
#pragma unroll
    for (int k = 0; k < kGenThreads / 32; ++k) base += (k < w) ? s_wcnt[k] : 0;
    int slot = base + __popc(ballot & ((1u << lane) - 1u));
    slot_id[slot] = (uint32_t)id;
#pragma unroll
    for (int d = 0; d < M; ++d) samp[(size_t)slot * 4 + d] = s[d];
}

// dense over the survivors: fp64 Kabsch in registers, fp32 copy of [R|t] and the
// rigorous fp32 error band of the inlier test (DESIGN.md "fp32 bracket")
template <int M>
__global__ void __launch_bounds__(kGenThreads)
k_kabsch(const float4 *__restrict__ P8, double thr2, Ctl *ctl,
         const int32_t *__restrict__ samp, float4 *__restrict__ m32, double *__restrict__ m64,
         int *__restrict__ cnt, uint4 *__restrict__ Aimg, float *__restrict__ band)
{
    lr::pdl_wait();
    lr::pdl_launch();
    if (ctl->done) return;
    const int nsurv = ctl->n_surv;
    double cen[3], cenq[3];
    tcs::tc_centre(P8, cen, cenq);
    const double P2 = (double)__uint_as_float(ctl->pt2max_bits), Qt = (double)__uint_as_float(ctl->qtmax_bits);
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < nsurv; slot += gridDim.x * blockDim.x) {
        double P[M][3], Q[M][3], T[12];
#pragma unroll
        for (int d = 0; d < M; ++d) load_pq(P8, (int64_t)samp[(size_t)slot * 4 + d], P[d], Q[d]);
        kabsch_small<M>(P, Q, T);
#pragma unroll
        for (int k = 0; k < 12; ++k) m64[(size_t)slot * 12 + k] = T[k];
        cnt[slot] = 0;
        // |fp32 residual component - canonical fp64 one| <= E for every correspondence:
        // rounding of R (u |p|_1), of t (u |t|), three FMAs (3u (|p|_1 + |t|)); the final
        // add is exact to u |d| which is O(thr) near the threshold.  6u leaves slack.
        const double u = 5.9604644775390625e-08;  // 2^-24
        double tinf = fmax(fabs(T[3]), fmax(fabs(T[7]), fabs(T[11])));
        double E = 6.0 * u * ((double)__uint_as_float(ctl->p1max_bits) + tinf + 1.0);
        double thr = sqrt(thr2);
        double delta = 4.0 * E * thr + 4.0 * E * E + 8.0 * u * thr2 + 1e-9;
        float lo = __double2float_rd(thr2 - delta);
        float hi = __double2float_ru(thr2 + delta);
        if (Aimg) {
            // tensor-core sweep: the model in the operand frame, t~ = t + R c - c', and the band its error bound gives
            double tt[3];
#pragma unroll
            for (int a = 0; a < 3; ++a)
                tt[a] = (T[4 * a + 3] + ((T[4 * a] * cen[0] + T[4 * a + 1] * cen[1]) + T[4 * a + 2] * cen[2])) - cenq[a];
            tcs::tc_write_model(Aimg, slot, T, tt);
            const double Et = tcs::tc_err_bound(P2, Qt, fmax(fabs(tt[0]), fmax(fabs(tt[1]), fabs(tt[2]))));
            band[slot] = __double2float_ru(4.0 * Et * thr + 4.0 * Et * Et + 8.0 * u * thr2 + 1e-9);
        }
        float *mf = reinterpret_cast<float *>(m32);
#pragma unroll
        for (int k = 0; k < 12; ++k) mf[m32_index(slot, k)] = (float)T[k];
        mf[m32_index(slot, 12)] = lo;
        mf[m32_index(slot, 13)] = hi;
        // |d0| >= c  =>  fl(d0 * d0) >= hi (k_score's early-out on the first residual component)
        mf[m32_index(slot, 14)] = __fsqrt_ru(hi) * 1.000001f;
        // the partner slot of the last, half-filled block must never count anything
        const int partner = (slot % kHypPerItem) < kScoreThreads ? slot + kScoreThreads : -1;
        if (partner >= nsurv) {
            mf[m32_index(partner, 12)] = -1.f;
            mf[m32_index(partner, 13)] = -1.f;
            mf[m32_index(partner, 14)] = 0.f;
#pragma unroll
            for (int k = 0; k < 12; ++k) mf[m32_index(partner, k)] = 0.f;
        }
    }
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// cl += (r < lo); ch += (r < hi)  as two FSETP + two predicated IADD
__device__ __forceinline__ void count2(float r, float lo, float hi, int &cl, int &ch)
{
    asm("{\n\t.reg .pred p, q;\n\tsetp.lt.f32 p, %2, %3;\n\tsetp.lt.f32 q, %2, %4;\n\t"
        "@p add.s32 %0, %0, 1;\n\t@q add.s32 %1, %1, 1;\n\t}"
        : "+r"(cl), "+r"(ch)
        : "f"(r), "f"(lo), "f"(hi));
}

// Rare path of the sweep: some residual of this group of kGroup points fell
// inside the fp32 error band of hypothesis `slot`.  Recompute the group's fp32
// residuals (same operations, same order => same bits), and decide the in-band
// ones with the canonical fp64 arithmetic of the oracle.
__device__ __noinline__ int recheck_group(const float4 *grp, int npts, const float4 *__restrict__ m32, int slot,
                                          const double *__restrict__ m64s, double thr2, unsigned long long *n_rechecked)
{
    const float *mf = reinterpret_cast<const float *>(m32);
    float v[14];
#pragma unroll
    for (int k = 0; k < 14; ++k) v[k] = mf[m32_index(slot, k)];
    const float4 r0 = make_float4(v[0], v[1], v[2], v[3]), r1 = make_float4(v[4], v[5], v[6], v[7]);
    const float4 r2 = make_float4(v[8], v[9], v[10], v[11]), bw = make_float4(v[12], v[13], 0.f, 0.f);
    int add = 0, evals = 0;
    for (int g = 0; g < npts; ++g) {
        const float4 A = grp[3 * g + 0], B = grp[3 * g + 1], C = grp[3 * g + 2];
        const float px = A.x, py = A.z, pz = B.x;
        float d0 = fmaf(r0.x, px, fmaf(r0.y, py, fmaf(r0.z, pz, r0.w))) + B.z;
        float d1 = fmaf(r1.x, px, fmaf(r1.y, py, fmaf(r1.z, pz, r1.w))) + C.x;
        float d2 = fmaf(r2.x, px, fmaf(r2.y, py, fmaf(r2.z, pz, r2.w))) + C.z;
        float rr = fmaf(d2, d2, fmaf(d1, d1, d0 * d0));
        if (rr >= bw.x && rr < bw.y) {
            double T[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) T[k] = m64s[k];
            add += res2_f64(T, (double)px, (double)py, (double)pz, -(double)B.z, -(double)C.x, -(double)C.z) < thr2 ? 1 : 0;
            ++evals;
        }
    }
    if (evals) atomicAdd(n_rechecked, (unsigned long long)evals);
    return add;
}

// Inlier sweep -- the dominant kernel.  Each thread owns TWO surviving
// hypotheses whose [R|t] live in registers as packed f32x2 pairs; every
// correspondence is read from shared memory as three warp-wide broadcast
// LDS.128 whose halves are ready-made (v, v) operands, so R p + t - q and
// |.|^2 cost 15 packed FFMA2/FADD2/FMUL2 for two residuals (7.5 issue slots per
// residual instead of 15).  Correspondences are staged with cp.async in
// double-buffered chunks.  Counting is exact: residuals below `lo` are inliers,
// above `hi` outliers, and the rare ones in between are decided in fp64 by
// recheck_group().  Work item = (64 survivors) x (range of point chunks); the
// split over points is chosen on the device from the survivor count so that a
// round with few survivors (ELC rejects ~97 % at 70 % outliers) fills the chip.
// SKIP (default): the first residual component d0 of four points is evaluated first (4 of
// the 15 packed operations); if |d0| >= c (c^2 >= hi) for all 64 hypotheses of the warp the
// point is an outlier for every one of them (r^2 >= d0^2 >= hi, rounding is monotone) and
// the remaining 11 operations and the counting are skipped through a warp-uniform branch.
// 58 % of the (warp, point) pairs take the early-out at cfg 3 (ncu: 0.86 M of 2.05 M
// executions of the live path); counts are unchanged by construction.
template <bool SKIP>
__global__ void __launch_bounds__(kScoreThreads, 16)
k_score(const float4 *__restrict__ P12, int64_t n_pad, Ctl *ctl, const float4 *__restrict__ m32,
        const double *__restrict__ m64, int *__restrict__ cnt, double thr2)
{
    if (ctl->done) return;
    __shared__ __align__(16) float4 sP[2][3 * kChunk];
    const int tid = threadIdx.x;
    const int nsurv = ctl->n_surv;
    const int nhb = (nsurv + kHypPerItem - 1) / kHypPerItem;
    if (nhb == 0) return;
    const int nchunks = (int)(n_pad / kChunk);
    int nps = (8 * (int)gridDim.x + nhb - 1) / nhb;
    nps = nps < 1 ? 1 : (nps > nchunks ? nchunks : nps);
    const int cpp = (nchunks + nps - 1) / nps;  // chunks per point split
    nps = (nchunks + cpp - 1) / cpp;
    const int nitems = nhb * nps;
    unsigned long long *n_rechecked = reinterpret_cast<unsigned long long *>(&ctl->n_rechecked);
    int buf = 0;             // staging buffer of the chunk about to be consumed (runs across items)
    bool prefetched = false;  // the first chunk of this item was staged during the previous item's last chunk

    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int hb = item / nps, ps = item - hb * nps;
        const int c_lo = ps * cpp, c_hi = min(nchunks, c_lo + cpp);
        const int slotA = hb * kHypPerItem + tid, slotB = slotA + kScoreThreads;
        const bool vA = slotA < nsurv, vB = slotB < nsurv;
        // 16 float2 (slot A, slot B) = 8 x 16-byte loads, already paired for f32x2 math
        const ulonglong2 *mp = reinterpret_cast<const ulonglong2 *>(m32) + ((size_t)hb * kScoreThreads + tid) * 8;
        const ulonglong2 q0 = mp[0], q1 = mp[1], q2 = mp[2], q3 = mp[3], q4 = mp[4], q5 = mp[5], q6 = mp[6];
        const u64 R00 = q0.x, R01 = q0.y, R02 = q1.x, T0 = q1.y;
        const u64 R10 = q2.x, R11 = q2.y, R12 = q3.x, T1 = q3.y;
        const u64 R20 = q4.x, R21 = q4.y, R22 = q5.x, T2 = q5.y;
        float loA, loB, hiA, hiB, cA = 0.f, cB = 0.f;
        upk2(q6.x, loA, loB);
        upk2(q6.y, hiA, hiB);
        if (SKIP) upk2(mp[7].x, cA, cB);
        if (!vA) loA = hiA = -1.f, cA = 0.f;  // slots past the survivor count hold stale models: count nothing
        if (!vB) loB = hiB = -1.f, cB = 0.f;
        // running #(rr < lo), #(rr < hi) per hypothesis; their difference grows only when a residual
        // lands inside the error band, which sends the group to the fp64 recheck
        int loCntA = 0, hiCntA = 0, loCntB = 0, hiCntB = 0, seenA = 0, seenB = 0, exactA = 0, exactB = 0;

        auto stage = [&](int c, int buf) {
            const float4 *gp = P12 + (size_t)c * 3 * kChunk;
#pragma unroll
            for (int k = 0; k < 3 * kChunk / kScoreThreads; ++k)
                cp_async16(&sP[buf][tid + k * kScoreThreads], gp + tid + k * kScoreThreads);
            cp_async_commit();
        };

        if (!prefetched) stage(c_lo, buf);
        for (int c = c_lo; c < c_hi; ++c) {
            if (c + 1 < c_hi) {
                stage(c + 1, buf ^ 1);
                cp_async_wait<1>();
            } else {
                // last chunk of the item: stage the first chunk of this CTA's next item behind it
                const int nxt = item + gridDim.x;
                prefetched = nxt < nitems;
                if (prefetched) {
                    stage((nxt - (nxt / nps) * nps) * cpp, buf ^ 1);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
            }
            __syncthreads();
            const ulonglong2 *sp = reinterpret_cast<const ulonglong2 *>(&sP[buf][0]);
            // the "some residual fell into the error band" check runs once per kBlock points: it is rare
            // (~0.5 events per hypothesis and 30k points), so the hot loop carries no slow-path state
            for (int ib = 0; ib < kChunk; ib += kBlock) {
#pragma unroll 2
            for (int i = ib; i < ib + kBlock; i += kGroup) {
                if (SKIP) {
                    // Row 0 of the group's four points first (4 of the 15 packed operations of a residual).
                    // |d0| >= c for all 64 hypotheses of the warp  =>  every r^2 >= d0^2 >= hi: the point is
                    // an outlier for all of them and the other 11 operations and the counting are skipped
                    // (warp-uniform branch).  Outlier correspondences are far from every half-way sensible
                    // model, and most surviving hypotheses are all-inlier samples close to each other.
                    u64 d[kGroup];
#pragma unroll
                    for (int g = 0; g < kGroup; ++g) d[g] = fma2v(R02, sp[3 * (i + g) + 1].x, T0);
#pragma unroll
                    for (int g = 0; g < kGroup; ++g) d[g] = fma2v(R01, sp[3 * (i + g) + 0].y, d[g]);
#pragma unroll
                    for (int g = 0; g < kGroup; ++g) d[g] = fma2v(R00, sp[3 * (i + g) + 0].x, d[g]);
#pragma unroll
                    for (int g = 0; g < kGroup; ++g) d[g] = add2(d[g], sp[3 * (i + g) + 1].y);
                    bool live[kGroup];
#pragma unroll
                    for (int g = 0; g < kGroup; ++g) {
                        float a, b;
                        upk2(d[g], a, b);
                        live[g] = __any_sync(0xffffffffu, (fabsf(a) < cA) | (fabsf(b) < cB));
                    }
#pragma unroll
                    for (int g = 0; g < kGroup; ++g) {
                        if (!live[g]) continue;
                        const ulonglong2 A0 = sp[3 * (i + g) + 0], B0 = sp[3 * (i + g) + 1], C0 = sp[3 * (i + g) + 2];
                        u64 p1 = fma2v(R12, B0.x, T1), p2 = fma2v(R22, B0.x, T2);
                        p1 = fma2v(R11, A0.y, p1); p2 = fma2v(R21, A0.y, p2);
                        p1 = fma2v(R10, A0.x, p1); p2 = fma2v(R20, A0.x, p2);
                        p1 = add2(p1, C0.x); p2 = add2(p2, C0.y);
                        const u64 rr0 = fma2(p2, p2, fma2(p1, p1, mul2(d[g], d[g])));
                        float ra, rb;
                        upk2(rr0, ra, rb);
                        count2(ra, loA, hiA, loCntA, hiCntA);
                        count2(rb, loB, hiB, loCntB, hiCntB);
                    }
                } else
                // two points at a time; for each coordinate the three rows' FMAs are issued back to
                // back so they share the point operand (z, then y, then x)
#pragma unroll
                for (int g = 0; g < kGroup; g += 2) {
                    const ulonglong2 A0 = sp[3 * (i + g) + 0], B0 = sp[3 * (i + g) + 1], C0 = sp[3 * (i + g) + 2];
                    const ulonglong2 A1 = sp[3 * (i + g) + 3], B1 = sp[3 * (i + g) + 4], C1 = sp[3 * (i + g) + 5];
                    u64 p0 = fma2v(R02, B0.x, T0), p1 = fma2v(R12, B0.x, T1), p2 = fma2v(R22, B0.x, T2);
                    u64 s0 = fma2v(R02, B1.x, T0), s1 = fma2v(R12, B1.x, T1), s2 = fma2v(R22, B1.x, T2);
                    p0 = fma2v(R01, A0.y, p0); p1 = fma2v(R11, A0.y, p1); p2 = fma2v(R21, A0.y, p2);
                    s0 = fma2v(R01, A1.y, s0); s1 = fma2v(R11, A1.y, s1); s2 = fma2v(R21, A1.y, s2);
                    p0 = fma2v(R00, A0.x, p0); p1 = fma2v(R10, A0.x, p1); p2 = fma2v(R20, A0.x, p2);
                    s0 = fma2v(R00, A1.x, s0); s1 = fma2v(R10, A1.x, s1); s2 = fma2v(R20, A1.x, s2);
                    p0 = add2(p0, B0.y); p1 = add2(p1, C0.x); p2 = add2(p2, C0.y);
                    s0 = add2(s0, B1.y); s1 = add2(s1, C1.x); s2 = add2(s2, C1.y);
                    const u64 rr0 = fma2(p2, p2, fma2(p1, p1, mul2(p0, p0)));
                    const u64 rr1 = fma2(s2, s2, fma2(s1, s1, mul2(s0, s0)));
                    float ra, rb;
                    upk2(rr0, ra, rb);
                    count2(ra, loA, hiA, loCntA, hiCntA);
                    count2(rb, loB, hiB, loCntB, hiCntB);
                    upk2(rr1, ra, rb);
                    count2(ra, loA, hiA, loCntA, hiCntA);
                    count2(rb, loB, hiB, loCntB, hiCntB);
                }
            }
                if (hiCntA - loCntA != seenA) {
                    seenA = hiCntA - loCntA;
                    exactA += recheck_group(&sP[buf][3 * ib], kBlock, m32, slotA, m64 + (size_t)slotA * 12, thr2, n_rechecked);
                }
                if (hiCntB - loCntB != seenB) {
                    seenB = hiCntB - loCntB;
                    exactB += recheck_group(&sP[buf][3 * ib], kBlock, m32, slotB, m64 + (size_t)slotB * 12, thr2, n_rechecked);
                }
            }
            __syncthreads();
            buf ^= 1;
        }
        const int cntA = loCntA + exactA, cntB = loCntB + exactB;
        if (nps == 1) {
            if (vA) cnt[slotA] = cntA;
            if (vB) cnt[slotB] = cntB;
        } else {
            if (vA && cntA) atomicAdd(&cnt[slotA], cntA);
            if (vB && cntB) atomicAdd(&cnt[slotB], cntB);
        }
    }
}

// exact counts -> packed arg-max
__global__ void __launch_bounds__(256)
k_resolve(Ctl *ctl, const uint32_t *__restrict__ slot_id, const int *__restrict__ cnt,
          int32_t *__restrict__ counts_out, int64_t id_base)
{
    if (ctl->done) return;
    const int nsurv = ctl->n_surv;
    unsigned long long key = 0ULL;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < nsurv; slot += gridDim.x * blockDim.x) {
        const int c = cnt[slot];
        const uint32_t id = slot_id[slot];
        const unsigned long long k = make_key(c, id);
        key = k > key ? k : key;
        if (counts_out) counts_out[(int64_t)id - id_base] = c;
    }
    key = warp_max_u64(key);
    if ((threadIdx.x & 31) == 0 && key) atomicMax(&ctl->round_key, key);
}

__global__ void k_round_end(Ctl *ctl, int64_t round_len, const int *__restrict__ need, int round_idx,
                            unsigned long long *user_key)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (ctl->done) return;
    if (ctl->round_key > ctl->best_key) ctl->best_key = ctl->round_key;
    if (user_key && ctl->round_key > *user_key) *user_key = ctl->round_key;
    ctl->iters_run += round_len;
    ctl->n_scored += ctl->n_surv;
    ctl->round_key = 0ULL;
    ctl->n_surv = 0;
    ctl->n_events = 0u;
    if (need) {
        long long cnt = (long long)(ctl->best_key >> 32) - 1;
        if (ctl->best_key != 0ULL && cnt >= (long long)need[round_idx]) ctl->done = 1;
    }
}

// ------------------------------------------------------------------------
// hypothesis sharding over NVLink peer memory: every rank owns a mailbox that its peers write
// into directly (cudaIpc-mapped device memory, lr_comm_init / lr_comm_connect); the exchange of a round's
// packed (count, id) key is part of the kernel that ends the round -- no host round trip, no separate collective
// ------------------------------------------------------------------------
constexpr int kMaxRanks = 16;
struct Mail {
    unsigned long long key;  // packed (count + 1) << 32 | ~id of the sender's slice of the round
    unsigned long long aux;  // survivors the sender scored in the round
    unsigned long long seq;  // exchange number the entry belongs to (written last, release)
    unsigned long long pad;
    double T[12];            // the fp64 model behind `key` (the sender computed it in k_kabsch: nobody re-derives it)
};
struct Comm {
    int rank, world;
    unsigned long long epoch;  // exchanges completed (identical on every rank: the calls are collective)
    int error, pad;
    Mail *peer[kMaxRanks];     // peer[g] = rank g's mailbox [2][kMaxRanks] (parity of the exchange, sender)
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ double ld_acquire_sys_f64(const double *p)
{
    double v;
    asm volatile("ld.acquire.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_f64(double *p, double v)
{
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// one warp: all-gather of (key, aux, model) through the peers' mailboxes; key MAX / aux SUM reduced, T = the model of
// the winning key (valid in lane 0); every rank gets the same
__device__ __forceinline__ void comm_exchange(Comm *cm, unsigned long long &key, unsigned long long &aux, double (&T)[12],
                                              int &err)
{
    const int lane = threadIdx.x & 31;
    const int G = cm->world, me = cm->rank;
    const unsigned long long ep = cm->epoch + 1;
    unsigned long long k = 0ULL, a = 0ULL;
    int bad = 0;
    if (lane < G) {
        Mail *dst = cm->peer[lane] + (size_t)(ep & 1ULL) * kMaxRanks + me;
        st_relaxed_sys(&dst->key, key);
        st_relaxed_sys(&dst->aux, aux);
#pragma unroll
        for (int j = 0; j < 12; ++j) st_relaxed_sys_f64(&dst->T[j], T[j]);
        __threadfence_system();
        st_release_sys(&dst->seq, ep);
        const Mail *src = cm->peer[me] + (size_t)(ep & 1ULL) * kMaxRanks + lane;
        const long long t0 = clock64();
        while (ld_acquire_sys(&src->seq) != ep) {
            if (clock64() - t0 > 6000000000LL) {  // ~3 s: a rank is missing; give up instead of hanging the GPU
                bad = 1;
                break;
            }
        }
        k = ld_acquire_sys(&src->key);
        a = ld_acquire_sys(&src->aux);
    }
    bad = __any_sync(0xffffffffu, bad);
    const unsigned long long kmax = warp_max_u64(k);
    // keys are unique (the id is part of them) unless they are 0: the lowest lane holding the maximum wins
    const unsigned who = __ballot_sync(0xffffffffu, lane < G && k == kmax);
    const int wl = who ? __ffs(who) - 1 : 0;
    if (lane == wl) {
        const Mail *src = cm->peer[me] + (size_t)(ep & 1ULL) * kMaxRanks + lane;
#pragma unroll
        for (int j = 0; j < 12; ++j) T[j] = ld_acquire_sys_f64(&src->T[j]);
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) T[j] = __shfl_sync(0xffffffffu, T[j], wl);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    __syncwarp();
    if (lane == 0) {
        cm->epoch = ep;
        if (bad) cm->error = 1;
    }
    key = kmax;
    aux = a;
    err = bad;
}

struct EndArgs {
    int64_t round_len;     // hypotheses of the whole round (all ranks together)
    const int *need;       // confidence exit table (nullable)
    int round_idx;
    Comm *comm;            // hypothesis sharding: exchange the round's key with the peers (nullable)
};

// k_resolve + k_round_end (+ the peer exchange) in one launch: every block leaves its best (key, slot) in
// blockbest[], the last block to finish (ticket) ends the round.  The winner's model is COPIED from m64[] (k_kabsch
// computed it; across ranks it travels in the mailbox) whenever a round improves the selection -- re-deriving it from
// the id cost a single-thread fp64 Jacobi (~10 us) at the end of every run.
// Used by the count-scoring runs; the fed-sample hook keeps the two-kernel form.
constexpr int kEndBlocksMax = 512;
__global__ void __launch_bounds__(256)
k_resolve_end(Ctl *ctl, const uint32_t *__restrict__ slot_id, const int *__restrict__ cnt, const double *__restrict__ m64,
              unsigned long long *__restrict__ blockbest, EndArgs a)
{
    lr::pdl_wait();
    lr::pdl_launch();
    if (ctl->done) return;
    const int nsurv = ctl->n_surv;
    unsigned long long key = 0ULL;
    int bslot = 0;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < nsurv; slot += gridDim.x * blockDim.x) {
        const unsigned long long k = make_key(cnt[slot], slot_id[slot]);
        if (k > key) {
            key = k;
            bslot = slot;
        }
    }
    __shared__ unsigned long long s_key[8];
    __shared__ int s_slot[8];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    auto warp_argmax = [&](unsigned long long &k, int &sl) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, k, o);
            const int s2 = __shfl_xor_sync(0xffffffffu, sl, o);
            if (k2 > k) {
                k = k2;
                sl = s2;
            }
        }
    };
    warp_argmax(key, bslot);
    if (lane == 0) {
        s_key[w] = key;
        s_slot[w] = bslot;
    }
    __syncthreads();
    if (w == 0) {
        key = lane < 8 ? s_key[lane] : 0ULL;
        bslot = lane < 8 ? s_slot[lane] : 0;
        warp_argmax(key, bslot);
        if (lane == 0) {
            blockbest[2 * blockIdx.x] = key;
            blockbest[2 * blockIdx.x + 1] = (unsigned long long)bslot;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&ctl->ticket_end, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (!s_last || threadIdx.x >= 32) return;
    __threadfence();
    unsigned long long rkey = 0ULL;
    int rslot = 0;
    for (int b = lane; b < (int)gridDim.x; b += 32) {
        const unsigned long long k = __ldcg(blockbest + 2 * b);
        if (k > rkey) {
            rkey = k;
            rslot = (int)__ldcg(blockbest + 2 * b + 1);
        }
    }
    warp_argmax(rkey, rslot);
    double T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = rkey ? __ldcg(m64 + (size_t)rslot * 12 + k) : 0.0;
    unsigned long long scored = (unsigned long long)nsurv;
    int err = 0;
    if (a.comm) comm_exchange(a.comm, rkey, scored, T, err);
    if (threadIdx.x != 0) return;
    ctl->ticket_end = 0u;
    if (err) ctl->comm_error = 1;
    unsigned long long best = ctl->best_key;
    if (rkey > best) {
        best = rkey;
        // a 0-inlier selection never replaces the identity
        if ((long long)(rkey >> 32) - 1 > 0) {
#pragma unroll
            for (int k = 0; k < 12; ++k) ctl->T[k] = T[k];
        }
    }
    ctl->best_key = best;
    ctl->iters_run += a.round_len;
    ctl->n_scored += (long long)scored;
    ctl->round_key = 0ULL;
    ctl->n_surv = 0;
    ctl->n_events = 0u;
    int done = err;
    if (a.need) {
        const long long c = (long long)(best >> 32) - 1;
        if (best != 0ULL && c >= (long long)a.need[a.round_idx]) done = 1;
    }
    if (done) ctl->done = 1;
}

// key -> model of the selected hypothesis (identity when none / zero inliers)
template <int M>
__global__ void k_model_from_key(const float *__restrict__ src, const float *__restrict__ tgt, int64_t n,
                                 uint64_t seed, int sampler, const uint32_t *__restrict__ growth,
                                 unsigned long long key_in, int use_ctl_key, Ctl *ctl)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long key = use_ctl_key ? ctl->best_key : key_in;
    if (!use_ctl_key) ctl->best_key = key;
    double T[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    long long cnt = (long long)(key >> 32) - 1;
    if (key != 0ULL && cnt > 0) {
        uint64_t id = (uint64_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFULL));
        int32_t s[M];
        sample_ids<M>(seed, id, sampler, n, growth, s);
        double P[M][3], Q[M][3];
#pragma unroll
        for (int d = 0; d < M; ++d)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                P[d][c] = (double)src[3 * (int64_t)s[d] + c];
                Q[d][c] = (double)tgt[3 * (int64_t)s[d] + c];
            }
        kabsch_small<M>(P, Q, T);
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) ctl->T[k] = T[k];
    ctl->refit_count = 0;
    ctl->err2 = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) ctl->csum[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) ctl->H[k] = 0.0;
}

__device__ __forceinline__ double block_sum(double v, double *sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    return r;  // valid in thread 0
}

// fetch one correspondence either directly ([n,3] arrays) or through index lists
__device__ __forceinline__ void fetch_pair(const float *a, const float *b, const int64_t *ia, const int64_t *ib,
                                           int64_t i, double (&p)[3], double (&q)[3])
{
    if (a == nullptr) {  // packed records of the run (Ws::P8) passed through `b`: same floats, one sector per pair
        load_pq(reinterpret_cast<const float4 *>(b), i, p, q);
        return;
    }
    const int64_t ka = ia ? ia[i] : i, kb = ib ? ib[i] : i;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        p[c] = (double)a[3 * ka + c];
        q[c] = (double)b[3 * kb + c];
    }
}

// pass 1 of the refit: exact inlier mask of ctl->T, count, coordinate sums
__global__ void __launch_bounds__(256)
k_mask_sums(const float *__restrict__ a, const float *__restrict__ b, const int64_t *__restrict__ ia,
            const int64_t *__restrict__ ib, int64_t n, double thr2, Ctl *ctl, uint8_t *__restrict__ mask)
{
    __shared__ double sh[8];
    double T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = ctl->T[k];
    double s[6] = {0, 0, 0, 0, 0, 0};
    double e2 = 0.0;
    int cnt = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double p[3], q[3];
        fetch_pair(a, b, ia, ib, i, p, q);
        const double r2 = res2_f64(T, p[0], p[1], p[2], q[0], q[1], q[2]);
        bool in = r2 < thr2;
        if (mask) mask[i] = in ? 1 : 0;
        if (in) {
            ++cnt;
            e2 += r2;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                s[c] += p[c];
                s[3 + c] += q[c];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double r = block_sum(s[k], sh);
        if (threadIdx.x == 0 && r != 0.0) atomicAdd(&ctl->csum[k], r);
    }
    double c = block_sum((double)cnt, sh);
    if (threadIdx.x == 0 && c != 0.0) atomicAdd((unsigned long long *)&ctl->refit_count, (unsigned long long)c);
    double e = block_sum(e2, sh);
    if (threadIdx.x == 0 && e != 0.0) atomicAdd(&ctl->err2, e);
}

// pass 2: centred cross-covariance over the inliers
__global__ void __launch_bounds__(256)
k_refit_H(const float *__restrict__ a, const float *__restrict__ b, const int64_t *__restrict__ ia,
          const int64_t *__restrict__ ib, int64_t n, double thr2, Ctl *ctl)
{
    __shared__ double sh[8];
    const long long k = ctl->refit_count;
    if (k <= 0) return;
    double T[12], cp[3], cq[3];
#pragma unroll
    for (int j = 0; j < 12; ++j) T[j] = ctl->T[j];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        cp[c] = ctl->csum[c] / (double)k;
        cq[c] = ctl->csum[3 + c] / (double)k;
    }
    double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double p[3], q[3];
        fetch_pair(a, b, ia, ib, i, p, q);
        if (res2_f64(T, p[0], p[1], p[2], q[0], q[1], q[2]) < thr2) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) H[3 * r + c] += (q[r] - cq[r]) * (p[c] - cp[c]);
        }
    }
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        double r = block_sum(H[j], sh);
        if (threadIdx.x == 0 && r != 0.0) atomicAdd(&ctl->H[j], r);
    }
}

__global__ void k_refit_solve(Ctl *ctl)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const long long k = ctl->refit_count;
    double T[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    if (k > 0) {
        double H[3][3], R[3][3], cp[3], cq[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) H[r][c] = ctl->H[3 * r + c];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            cp[c] = ctl->csum[c] / (double)k;
            cq[c] = ctl->csum[3 + c] / (double)k;
        }
        rot_from_H(H, R);
        finish_T(R, cp, cq, T);
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) ctl->Tref[j] = T[j];
}

// Inlier mask of ctl->T, count, squared-error sum and the least-squares refit over the inliers in ONE launch
// (FR.py:99-111): per-block partial sums of {1, r^2, p', q', q' p'^T} (coordinates relative to the first pair
// rounded to 1024 m, so map-frame offsets do not cancel in the covariance), written to fixed slots; the last
// block to finish (ticket) adds them in block order -- a fixed-order two-stage reduction, so T_refit is
// bit-reproducible run to run -- and solves the Kabsch problem:  H = sum q' p'^T - (sum q')(sum p')^T / k.
// `host_out` (nullable, pinned + mapped): the control block is written there by the same block, so the caller
// needs no copy after the kernel, only the stream synchronisation.
constexpr unsigned kTcEventCap = 1024;  // in-band residuals an epilogue warp of the tensor sweep can list per round (one private
                                       // list per warp and CTA: 148 x 16 x 16 KB); beyond it the warp decides on the spot
constexpr int kFinVals = 17;
constexpr int kFinBlocksMax = 512;
__global__ void __launch_bounds__(256)
k_finish(const float *__restrict__ a, const float *__restrict__ b, const int64_t *__restrict__ ia,
         const int64_t *__restrict__ ib, int64_t n, double thr2, Ctl *ctl, uint8_t *__restrict__ mask,
         double *__restrict__ partial, int want_refit, Ctl *host_out)
{
    lr::pdl_wait();
    lr::pdl_launch();
    __shared__ int s_last;
    double T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = ctl->T[k];
    double o[3] = {0, 0, 0}, oq[3] = {0, 0, 0};
    if (n > 0) {
        double p0[3], q0[3];
        fetch_pair(a, b, ia, ib, 0, p0, q0);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            o[c] = 1024.0 * rint(p0[c] * (1.0 / 1024.0));
            oq[c] = 1024.0 * rint(q0[c] * (1.0 / 1024.0));
        }
    }
    double v[kFinVals];
#pragma unroll
    for (int k = 0; k < kFinVals; ++k) v[k] = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double p[3], q[3];
        fetch_pair(a, b, ia, ib, i, p, q);
        const double r2 = res2_f64(T, p[0], p[1], p[2], q[0], q[1], q[2]);
        const bool in = r2 < thr2;
        if (mask) mask[i] = in ? 1 : 0;
        if (in) {
            v[0] += 1.0;
            v[1] += r2;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                p[c] -= o[c];
                q[c] -= oq[c];
                v[2 + c] += p[c];
                v[5 + c] += q[c];
            }
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) v[8 + 3 * r + c] += q[r] * p[c];
        }
    }
    // first stage, fixed order: shuffle tree inside each warp, then thread k adds the 8 warp sums of value k in order
    // (one barrier instead of 34: the 17 values used to go through block_sum one after the other)
    __shared__ double s_part[8][kFinVals];
    {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < kFinVals; ++k) {
            double x = v[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) s_part[w][k] = x;
        }
        __syncthreads();
        if (threadIdx.x < kFinVals) {
            double r = 0.0;
            for (int ww = 0; ww < (int)(blockDim.x >> 5); ++ww) r += s_part[ww][threadIdx.x];
            partial[(size_t)blockIdx.x * kFinVals + threadIdx.x] = r;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&ctl->ticket_fin, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // second stage, fixed order: warp w owns the values k = w, w + 8, ...; lane l adds blocks l, l + 32, ... in order,
    // then the shuffle tree -- the 17 sums run side by side instead of 17 block-wide reductions in a row
    __shared__ double s_tot[kFinVals];
    {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        for (int k = w; k < kFinVals; k += (int)(blockDim.x >> 5)) {
            double x = 0.0;
            for (int blk = lane; blk < (int)gridDim.x; blk += 32) x += __ldcg(&partial[(size_t)blk * kFinVals + k]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) s_tot[k] = x;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot[kFinVals];
#pragma unroll
        for (int k = 0; k < kFinVals; ++k) tot[k] = s_tot[k];
        const long long k = (long long)(tot[0] + 0.5);
        double Tr[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
        if (k > 0 && want_refit) {
            double H[3][3], R[3][3], cp[3], cq[3];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) H[r][c] = tot[8 + 3 * r + c] - tot[5 + r] * tot[2 + c] / (double)k;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                cp[c] = o[c] + tot[2 + c] / (double)k;
                cq[c] = oq[c] + tot[5 + c] / (double)k;
            }
            rot_from_H(H, R);
            finish_T(R, cp, cq, Tr);
        }
#pragma unroll
        for (int j = 0; j < 12; ++j) ctl->Tref[j] = Tr[j];
        ctl->refit_count = k;
        ctl->err2 = tot[1];
        ctl->ticket_fin = 0u;
        __threadfence();
    }
    __syncthreads();
    if (host_out) {
        const unsigned long long *srcw = reinterpret_cast<const unsigned long long *>(ctl);
        unsigned long long *dstw = reinterpret_cast<unsigned long long *>(host_out);
        for (int w = threadIdx.x; w < (int)(sizeof(Ctl) / 8); w += blockDim.x) dstw[w] = __ldcg(srcw + w);
        __threadfence_system();
    }
}

__global__ void k_set_identity(Ctl *ctl)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int k = 0; k < 12; ++k) ctl->T[k] = (k % 5 == 0) ? 1.0 : 0.0;
}

__global__ void k_set_T(Ctl *ctl, const double *T12)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int k = 0; k < 12; ++k) ctl->T[k] = T12[k];
    ctl->ticket_fin = 0u;  // (these entries run on a bare control block: no k_ctl_reset before them)
    ctl->comm_error = 0;
    ctl->gc.mode = 0;
    ctl->refit_count = 0;
    ctl->err2 = 0.0;
    for (int k = 0; k < 6; ++k) ctl->csum[k] = 0.0;
    for (int k = 0; k < 9; ++k) ctl->H[k] = 0.0;
}

// out[i, 0:3] = fp32(R p_i + t) (canonical fp64 order), out[i, 3:8] = 0: the 8-wide rows the exact
// nearest-neighbour sweep (lr_match_nn, D = 8) searches in 3-D
__global__ void k_transform_pad8(const float *__restrict__ xyz, int64_t n, const double *__restrict__ T12,
                                 float *__restrict__ out)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
    float4 lo, hi = make_float4(0.f, 0.f, 0.f, 0.f);
    lo.x = (float)(((T12[0] * px + T12[1] * py) + T12[2] * pz) + T12[3]);
    lo.y = (float)(((T12[4] * px + T12[5] * py) + T12[6] * pz) + T12[7]);
    lo.z = (float)(((T12[8] * px + T12[9] * py) + T12[10] * pz) + T12[11]);
    lo.w = 0.f;
    reinterpret_cast<float4 *>(out)[2 * i] = lo;
    reinterpret_cast<float4 *>(out)[2 * i + 1] = hi;
}

template <int M>
__global__ void k_sample_only(uint64_t seed, int sampler, int64_t n, const uint32_t *__restrict__ growth, int64_t id_lo,
                              int64_t H, int32_t *out)
{
    int64_t h = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (h >= H) return;
    int32_t s[M];
    sample_ids<M>(seed, (uint64_t)(id_lo + h), sampler, n, growth, s);
#pragma unroll
    for (int d = 0; d < M; ++d) out[h * M + d] = s[d];
}

// compacted fp64 models -> sample order (fed-sample hook only)
__global__ void k_scatter_models(const Ctl *ctl, const uint32_t *slot_id, const double *m64, double *models)
{
    int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= ctl->n_surv) return;
    const size_t h = slot_id[slot];
    for (int k = 0; k < 12; ++k) models[h * 12 + k] = m64[(size_t)slot * 12 + k];
}

// ------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------

// PROSAC growth function T'_n (same expressions as the oracle's lro_prosac_growth).  The table is a
// function of (N, m) only: the host copy is cached, and the upload is skipped while the device copy of the same
// table still sits at the same address of the same arena block (a run of same-sized PROSAC pairs uploads once).
struct GrowthCache {
    std::vector<uint32_t> host;
    int64_t N = -1;
    int m = 0;
    const void *dev[lr::SLOT_COUNT] = {nullptr, nullptr, nullptr, nullptr};   // where the cached table was uploaded last
    int64_t devN[lr::SLOT_COUNT] = {-1, -1, -1, -1};
    int devm[lr::SLOT_COUNT] = {0, 0, 0, 0};
    // signature of the layout ws_setup carved last in each arena slot: any other layout may overwrite the table
    int64_t sig[lr::SLOT_COUNT][7] = {};
};
GrowthCache g_growth[64];

size_t tc_regions() { return (size_t)lr::sm_count() * tcs::NEPI; }

int ws_setup(int64_t n, int64_t round, int64_t nrounds, Ws &ws, bool prosac = false, int slot = lr::SLOT_RANSAC,
             bool gc = false)
{
    ws.n_pad = ((n + kChunk - 1) / kChunk) * kChunk;
    if (ws.n_pad == 0) ws.n_pad = kChunk;
    // the sweep walks slots in blocks of kHypPerItem: size the per-slot arrays for the padded count
    const int64_t slots = round + kHypPerItem;
    size_t bytes = lr::padded(sizeof(Ctl)) + lr::padded(sizeof(float4) * 3 * ws.n_pad) +
                   lr::padded(sizeof(float4) * 2 * ws.n_pad) +
                   lr::padded(sizeof(int32_t) * 4 * slots) + lr::padded(sizeof(uint32_t) * slots) +
                   lr::padded(sizeof(float4) * 4 * slots) + lr::padded(sizeof(double) * 12 * slots) +
                   lr::padded(sizeof(int) * slots) + lr::padded(sizeof(int) * (nrounds + 1)) +
                   lr::padded(sizeof(double) * 16) + lr::padded(sizeof(uint32_t) * (prosac ? n : 1)) +
                   lr::padded((size_t)((slots + tcs::TM - 1) / tcs::TM) * tcs::A_BLOCK_BYTES) +
                   lr::padded((size_t)ws.n_pad * 64) + lr::padded(sizeof(float) * slots) +
                   lr::padded(sizeof(double) * kFinBlocksMax * kFinVals) + lr::padded(sizeof(int4) * (size_t)kTcEventCap * tc_regions()) + lr::padded(sizeof(float) * 3 * (size_t)(ws.n_pad / 256 + 1)) + lr::padded(sizeof(unsigned long long) * 2 * kEndBlocksMax) +
                   lr::padded(sizeof(float) * 6 * (size_t)(n > 0 ? n : 1));
    if (gc)
        bytes += lr::padded(sizeof(unsigned long long) * slots) + lr::padded(sizeof(int32_t) * (n > 0 ? n : 1)) +
                 lr::padded(sizeof(double) * 12 * kGcMaxTrials) + lr::padded(sizeof(unsigned long long) * kGcMaxTrials) +
                 lr::padded(sizeof(int) * kGcMaxTrials);
    void *base = lr::arena_get(slot, bytes);
    if (!base) return LR_ERR_ALLOC;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) {
            const int64_t sig[7] = {n, round, nrounds, (int64_t)prosac, (int64_t)gc, (int64_t)(uintptr_t)base,
                                    (int64_t)lr::arena_gen(slot)};
            if (memcmp(sig, g_growth[dev].sig[slot], sizeof(sig)) != 0) {
                memcpy(g_growth[dev].sig[slot], sig, sizeof(sig));
                g_growth[dev].dev[slot] = nullptr;
            }
        }
    }
    lr::Carver cv(base);
    ws.ctl = cv.take<Ctl>(1);
    ws.P12 = cv.take<float4>(3 * ws.n_pad);
    ws.P8 = cv.take<float4>(2 * ws.n_pad);
    ws.samp = cv.take<int32_t>(4 * slots);
    ws.slot_id = cv.take<uint32_t>(slots);
    ws.m32 = cv.take<float4>(4 * slots);
    ws.m64 = cv.take<double>(12 * slots);
    ws.cnt = cv.take<int>(slots);
    ws.need = cv.take<int>(nrounds + 1);
    ws.scratchT = cv.take<double>(16);
    ws.growth = prosac ? cv.take<uint32_t>(n) : cv.take<uint32_t>(1);
    if (!prosac) ws.growth = nullptr;
    ws.Aimg = reinterpret_cast<uint4 *>(cv.take<char>((size_t)((slots + tcs::TM - 1) / tcs::TM) * tcs::A_BLOCK_BYTES));
    ws.Bimg = reinterpret_cast<uint4 *>(cv.take<char>((size_t)ws.n_pad * 64));
    ws.band = cv.take<float>(slots);
    ws.partial = cv.take<double>((size_t)kFinBlocksMax * kFinVals);
    ws.events = cv.take<int4>((size_t)kTcEventCap * tc_regions());
    ws.packmax = cv.take<float>(3 * (size_t)(ws.n_pad / 256 + 1));
    ws.blockbest = cv.take<unsigned long long>(2 * (size_t)kEndBlocksMax);
    ws.stage = cv.take<float>(6 * (size_t)(n > 0 ? n : 1));
    ws.q64 = gc ? cv.take<unsigned long long>(slots) : nullptr;
    ws.lo_L = gc ? cv.take<int32_t>(n > 0 ? n : 1) : nullptr;
    ws.tr_T = gc ? cv.take<double>(12 * kGcMaxTrials) : nullptr;
    ws.tr_q = gc ? cv.take<unsigned long long>(kGcMaxTrials) : nullptr;
    ws.tr_inl = gc ? cv.take<int>(kGcMaxTrials) : nullptr;
    return LR_OK;
}

int check_params(const LrRansacParams *p, int64_t n)
{
    LR_REQUIRE(p != nullptr, "params is null");
    LR_REQUIRE(p->sample_size == 3 || p->sample_size == 4, "sample_size must be 3 or 4");
    LR_REQUIRE(p->sampler == LR_SAMPLER_UNIFORM || p->sampler == LR_SAMPLER_REPLACE || p->sampler == LR_SAMPLER_PROSAC,
               "sampler must be one of LR_SAMPLER_*");
    LR_REQUIRE(p->threshold > 0.0, "threshold must be positive");
    LR_REQUIRE(p->max_iters >= 0 && p->max_iters < (int64_t)0xFFFFFFFFLL, "max_iters out of range");
    LR_REQUIRE(p->round_size > 0 && p->round_size <= (1 << 20), "round_size out of range");
    LR_REQUIRE(n >= 0 && n < (int64_t)1 << 31, "n out of range");
    LR_REQUIRE(p->scoring == LR_SCORE_COUNT || p->scoring == LR_SCORE_MSAC, "scoring must be one of LR_SCORE_*");
    if (p->scoring == LR_SCORE_MSAC)
        LR_REQUIRE(p->lo_rounds >= 0 && p->lo_rounds <= 64 && p->lo_trials >= 0 && p->lo_trials <= kGcMaxTrials &&
                       p->lsq_iters >= 0 && p->lsq_iters <= 64,
                   "lo_rounds / lsq_iters must be in [0, 64], lo_trials in [0, 64]");
    return LR_OK;
}

int64_t batch_len(int64_t total)
{
    const int64_t cap = (int64_t)1 << 20;
    return total < 1 ? 1 : (total < cap ? total : cap);
}

int upload_growth(const Ws &ws, int64_t N, int m, cudaStream_t st, int slot = lr::SLOT_RANSAC)
{
    if (!ws.growth) return LR_OK;
    int dev = 0;
    LR_CUDA_TRY(cudaGetDevice(&dev));
    LR_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
    GrowthCache &gc = g_growth[dev];
    if (gc.N != N || gc.m != m) {
        gc.host.resize((size_t)N);
        std::vector<uint32_t> &g = gc.host;
        double T_n = (double)kProsacTN;
        for (int i = 0; i < m; ++i) T_n *= (double)(m - i) / (double)(N - i);
        uint32_t T_prime = 1;
        for (int64_t i = 0; i < N; ++i) {
            if (i + 1 <= m) {
                g[i] = T_prime;
                continue;
            }
            double T_next = (double)(i + 1) * T_n / (double)(i + 1 - m);
            double inc = ceil(T_next - T_n);
            if (!(inc < 4.0e9)) inc = 4.0e9;
            uint64_t v = (uint64_t)T_prime + (uint64_t)inc;
            g[i] = v > 0xFFFFFFF0ULL ? 0xFFFFFFF0U : (uint32_t)v;
            T_n = T_next;
            T_prime = g[i];
        }
        gc.N = N;
        gc.m = m;
        for (int k = 0; k < lr::SLOT_COUNT; ++k) gc.dev[k] = nullptr;
    }
    if (gc.dev[slot] == (const void *)ws.growth && gc.devN[slot] == N && gc.devm[slot] == m) return LR_OK;
    // pageable source: the runtime stages it before returning; no stream synchronisation needed
    LR_CUDA_TRY(cudaMemcpyAsync(ws.growth, gc.host.data(), sizeof(uint32_t) * N, cudaMemcpyHostToDevice, st));
    gc.dev[slot] = ws.growth;
    gc.devN[slot] = N;
    gc.devm[slot] = m;
    return LR_OK;
}

// opt the tensor sweep's kernels into their dynamic shared memory once per device
cudaError_t tc_smem_attr()
{
    static bool done[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || done[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(tcs::k_score_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcs::kSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tcs::k_score_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcs::kSmemBytes);
    if (e != cudaSuccess) return e;
    done[dev] = true;
    return cudaSuccess;
}

int launch_pack(const float *src, const float *tgt, int64_t n, const Ws &ws, cudaStream_t st, bool want12 = true)
{
    const int tok = lr::prof_begin(lr::PROF_PACK, st);
    int blocks = (int)((ws.n_pad + 255) / 256);
    LR_CUDA_TRY(lr::launch_pdl(k_pack, dim3(blocks), dim3(256), 0, st, src, tgt, n, ws.n_pad, ws.P12, ws.P8, ws.Bimg, ws.ctl,
                               ws.packmax, want12 ? 1 : 0));
    lr::prof_end(tok, st);
    LR_CUDA_TRY(cudaGetLastError());
    return LR_OK;
}

#include "lr_ransac_gc.cuh"

// one round: ids [lo, hi) (or H fed samples), leaves the result in ctl->round_key (LR_SCORE_MSAC:
// ctl->gc.round_q / round_pick; counts_out then receives #(r^2 < tau^2) and scores_out the q values).
// With `end` (count scoring) the round is also closed by the same launch sequence: k_resolve_end.
int launch_round(const float *src, const float *tgt, int64_t n, const LrRansacParams &p, const Ws &ws, int64_t lo,
                 int64_t hi, const int32_t *fed, int32_t *counts_out, cudaStream_t st, int64_t *scores_out = nullptr,
                 const EndArgs *end = nullptr)
{
    const int64_t len = hi - lo;
    if (len <= 0 && !end) return LR_OK;
    const double thr2 = p.threshold * p.threshold;
    const int sms = lr::sm_count();
    if (len > 0) {
        int gblocks = (int)((len + kGenThreads - 1) / kGenThreads);
        int kblocks = gblocks < sms * 8 ? gblocks : sms * 8;
        // the tensor-core sweep serves count scoring; its operand image is only built when it will run
        const bool use_tc = p.scoring == LR_SCORE_COUNT && g_score_mode == 0;
        uint4 *aimg = use_tc ? ws.Aimg : nullptr;
        int tok = lr::prof_begin(lr::PROF_GEN, st);
        const float4 *P8c = ws.P8;
        const int32_t *sampc = ws.samp;
        const uint32_t *growthc = ws.growth;
        if (p.sample_size == 3) {
            LR_CUDA_TRY(lr::launch_pdl(k_gen<3>, dim3(gblocks), dim3(kGenThreads), 0, st, P8c, n, (uint64_t)p.seed, (int)p.sampler,
                                       (int)p.use_elc, p.elc_ratio, lo, hi, fed, growthc, ws.ctl, ws.slot_id, ws.samp));
            LR_CUDA_TRY(lr::launch_pdl(k_kabsch<3>, dim3(kblocks), dim3(kGenThreads), 0, st, P8c, thr2, ws.ctl, sampc, ws.m32,
                                       ws.m64, ws.cnt, aimg, ws.band));
        } else {
            LR_CUDA_TRY(lr::launch_pdl(k_gen<4>, dim3(gblocks), dim3(kGenThreads), 0, st, P8c, n, (uint64_t)p.seed, (int)p.sampler,
                                       (int)p.use_elc, p.elc_ratio, lo, hi, fed, growthc, ws.ctl, ws.slot_id, ws.samp));
            LR_CUDA_TRY(lr::launch_pdl(k_kabsch<4>, dim3(kblocks), dim3(kGenThreads), 0, st, P8c, thr2, ws.ctl, sampc, ws.m32,
                                       ws.m64, ws.cnt, aimg, ws.band));
        }
        lr::prof_end(tok, st);
        if (p.scoring == LR_SCORE_MSAC) return gc_launch_score(src, tgt, n, p, ws, lo, len, scores_out, counts_out, st);
        tok = lr::prof_begin(lr::PROF_SCORE, st);
        if (use_tc) {
            // persistent CTAs, one per SM (TMEM: 4 x 96 accumulator columns; 57 KB of operand staging)
            LR_CUDA_TRY(tc_smem_attr());
            LR_CUDA_TRY(lr::launch_pdl(tcs::k_score_tc<false>, dim3(sms), dim3(tcs::NTHREADS), tcs::kSmemBytes, st,
                                       (const uint4 *)ws.Aimg, (const uint4 *)ws.Bimg, P8c, n, ws.n_pad, ws.ctl,
                                       (const double *)ws.m64, (const float *)ws.band, ws.cnt, thr2, (float *)nullptr, ws.events,
                                       kTcEventCap));
        } else if (g_score_mode == 2) {
            // fp32 sweep: 16 resident one-warp CTAs per SM (128 registers per thread fill the register file; 12 KB of
            // staging each): no CTA-level barrier couples warps whose early-out rates differ
            k_score<true><<<sms * 16, kScoreThreads, 0, st>>>(ws.P12, ws.n_pad, ws.ctl, ws.m32, ws.m64, ws.cnt, thr2);
        } else {
            k_score<false><<<sms * 16, kScoreThreads, 0, st>>>(ws.P12, ws.n_pad, ws.ctl, ws.m32, ws.m64, ws.cnt, thr2);
        }
        lr::prof_end(tok, st);
    }
    int rblocks = (int)((len + 255) / 256);
    if (rblocks > sms * 2) rblocks = sms * 2;
    if (rblocks < 1) rblocks = 1;
    if (end) {
        const int tok = lr::prof_begin(lr::PROF_END, st);
        LR_CUDA_TRY(lr::launch_pdl(k_resolve_end, dim3(rblocks < kEndBlocksMax ? rblocks : kEndBlocksMax), dim3(256), 0, st,
                                   ws.ctl, (const uint32_t *)ws.slot_id, (const int *)ws.cnt, (const double *)ws.m64,
                                   ws.blockbest, *end));
        lr::prof_end(tok, st);
    } else {
        k_resolve<<<rblocks, 256, 0, st>>>(ws.ctl, ws.slot_id, ws.cnt, counts_out, lo);
    }
    LR_CUDA_TRY(cudaGetLastError());
    return LR_OK;
}

void T12_to_16(const double *T12, double *T16)
{
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) T16[4 * r + c] = T12[4 * r + c];
    T16[12] = T16[13] = T16[14] = 0.0;
    T16[15] = 1.0;
}

void identity16(double *T)
{
    for (int k = 0; k < 16; ++k) T[k] = (k % 5 == 0) ? 1.0 : 0.0;
}

int finish_blocks(int64_t n)
{
    int64_t blocks = (n + 255) / 256;
    const int cap = lr::sm_count() * 3 < kFinBlocksMax ? lr::sm_count() * 3 : kFinBlocksMax;
    if (blocks > cap) blocks = cap;
    return blocks < 1 ? 1 : (int)blocks;
}

// where the selected model comes from when a run is finished
enum FinHow {
    FIN_MODEL_READY = 0,  // count scoring through enqueue_run: k_resolve_end left it in ctl->T
    FIN_GC = 1,           // LR_SCORE_MSAC: kept in ctl->gc
    FIN_FROM_KEY = 2      // lr_ransac_finalize: decode the caller's key
};

// (model) + mask + refit + control block to `host_out` (pinned, nullable): launches only
int finish_launch(const float *src, const float *tgt, int64_t n, const LrRansacParams &p, const Ws &ws, FinHow how,
                  uint64_t key, bool want_refit, uint8_t *mask, cudaStream_t st, bool packed, Ctl *host_out)
{
    // after launch_pack the run's kernels read the packed device copy; src / tgt themselves (device or pinned
    // host memory) are then only touched by k_pack
    const float *fa = packed ? nullptr : src, *fb = packed ? reinterpret_cast<const float *>(ws.P8) : tgt;
    const double thr2 = p.threshold * p.threshold;
    if (how == FIN_GC)
        k_gc_commit<<<1, 32, 0, st>>>(ws.ctl);  // the model was kept in ctl->gc, there is no key to decode
    else if (how == FIN_FROM_KEY) {
        if (p.sample_size == 3)
            k_model_from_key<3><<<1, 32, 0, st>>>(src, tgt, n, p.seed, p.sampler, ws.growth, key, 0, ws.ctl);
        else
            k_model_from_key<4><<<1, 32, 0, st>>>(src, tgt, n, p.seed, p.sampler, ws.growth, key, 0, ws.ctl);
    }
    const int tok = lr::prof_begin(lr::PROF_FIN, st);
    LR_CUDA_TRY(lr::launch_pdl(k_finish, dim3(finish_blocks(n)), dim3(256), 0, st, fa, fb, (const int64_t *)nullptr,
                               (const int64_t *)nullptr, n, thr2, ws.ctl, mask, ws.partial, want_refit ? 1 : 0, host_out));
    lr::prof_end(tok, st);
    LR_CUDA_TRY(cudaGetLastError());
    return LR_OK;
}

// host copy of the control block -> the caller's outputs
void finish_read(const Ctl &h, bool want_refit, double *T_out, double *T_refit, LrRansacStats *stats)
{
    if (T_out) T12_to_16(h.T, T_out);
    if (T_refit) {
        if (want_refit) T12_to_16(h.Tref, T_refit);
        else identity16(T_refit);
    }
    if (stats) {
        stats->iters_run = h.iters_run;
        stats->n_scored = h.n_scored;
        stats->n_rechecked = h.n_rechecked;
        long long cnt = (long long)(h.best_key >> 32) - 1;
        if (h.best_key != 0ULL) {
            stats->best_id = (int64_t)(0xFFFFFFFFu - (uint32_t)(h.best_key & 0xFFFFFFFFULL));
            stats->best_count = cnt;
        } else {
            stats->best_id = -1;
            stats->best_count = -1;
        }
        stats->refit_count = h.refit_count;
        stats->best_score = stats->lo_score = stats->final_score = 0;
        stats->lo_improved = stats->lsq_improved = 0;
        if (h.gc.mode) {
            stats->best_id = h.gc.has_model ? (int64_t)h.gc.best_id : -1;
            stats->best_count = h.gc.has_model ? (int64_t)h.gc.best_inl : -1;
            stats->best_score = (int64_t)h.gc.best_q;
            stats->lo_score = (int64_t)h.gc.lo_q;
            stats->final_score = (int64_t)h.gc.cur_q;
            stats->lo_improved = h.gc.lo_improved;
            stats->lsq_improved = h.gc.lsq_improved;
        }
    }
}

// per device: two internal streams + pinned result staging (lr_ransac_rigid_batch), one pinned control block
// for the single-run entries, and the hypothesis-sharding communicator
struct CommHost {
    bool ready = false;
    int rank = 0, world = 1;
    Comm *dev = nullptr;    // device copy handed to the kernels
    Mail *box = nullptr;    // this rank's mailbox [2][kMaxRanks]
    void *peer[kMaxRanks] = {};
};
struct BatchCtx {
    cudaStream_t lane[2] = {nullptr, nullptr};
    cudaEvent_t start = nullptr, done[2] = {nullptr, nullptr};
    Ctl *host = nullptr;
    int host_cap = 0;
    Ctl *host1 = nullptr;
    CommHost comm;
};
BatchCtx g_batch[64];

int dev_ctx(BatchCtx **out)
{
    int dev = 0;
    LR_CUDA_TRY(cudaGetDevice(&dev));
    LR_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
    BatchCtx &c = g_batch[dev];
    if (!c.host1) LR_CUDA_TRY(cudaMallocHost(&c.host1, sizeof(Ctl)));
    *out = &c;
    return LR_OK;
}

int batch_ctx(int count, BatchCtx **out)
{
    int rc = dev_ctx(out);
    if (rc) return rc;
    BatchCtx &c = **out;
    if (!c.lane[0]) {
        for (int l = 0; l < 2; ++l) {
            LR_CUDA_TRY(cudaStreamCreateWithFlags(&c.lane[l], cudaStreamNonBlocking));
            LR_CUDA_TRY(cudaEventCreateWithFlags(&c.done[l], cudaEventDisableTiming));
        }
        LR_CUDA_TRY(cudaEventCreateWithFlags(&c.start, cudaEventDisableTiming));
    }
    if (c.host_cap < count) {
        if (c.host) cudaFreeHost(c.host);
        c.host = nullptr;
        c.host_cap = 0;
        LR_CUDA_TRY(cudaMallocHost(&c.host, sizeof(Ctl) * (size_t)(count + 16)));
        c.host_cap = count + 16;
    }
    return LR_OK;
}

// (model) + mask + refit, control block written into pinned host memory by the last kernel, one synchronisation
int finish(const float *src, const float *tgt, int64_t n, const LrRansacParams &p, const Ws &ws, FinHow how,
           uint64_t key, double *T_out, double *T_refit, uint8_t *mask, LrRansacStats *stats, cudaStream_t st,
           bool packed = true)
{
    const bool want_refit = (T_refit != nullptr) && p.refit;
    BatchCtx *ctx = nullptr;
    int rc = dev_ctx(&ctx);
    if (rc) return rc;
    rc = finish_launch(src, tgt, n, p, ws, how, key, want_refit, mask, st, packed, ctx->host1);
    if (rc) return rc;
    LR_CUDA_TRY(cudaStreamSynchronize(st));
    if (ctx->host1->comm_error) {
        lr::set_error("hypothesis sharding: a peer rank did not answer the key exchange (timeout)");
        return LR_ERR_CUDA;
    }
    finish_read(*ctx->host1, want_refit, T_out, T_refit, stats);
    return LR_OK;
}

// everything of one run up to (not including) the mask / refit / read-back, enqueued on `st`.  With a
// communicator (`world` > 1) this rank generates and scores only its contiguous slice of every round and the
// round's packed key is exchanged with the peers inside k_resolve_end; every rank then holds the same selection.
int enqueue_run(const float *src, const float *tgt, int64_t n, const LrRansacParams &p, int slot, Ws &ws, cudaStream_t st,
                Comm *comm = nullptr, int rank = 0, int world = 1)
{
    // With the confidence exit the round length is part of the result's
    // definition (the exit is evaluated at round ends); with a fixed budget the
    // result is a plain arg-max, so hypotheses are batched as large as the
    // scratch allows to keep every launch chip-filling.
    const bool use_conf = p.confidence < 1.0 && p.max_iters > 0;
    const int64_t R = use_conf ? (int64_t)p.round_size : batch_len(p.max_iters);
    const int64_t nrounds = (p.max_iters + R - 1) / R;
    const bool gc = p.scoring == LR_SCORE_MSAC;
    int rc = ws_setup(n, R, nrounds, ws, p.sampler == LR_SAMPLER_PROSAC, slot, gc);
    if (rc) return rc;
    rc = upload_growth(ws, n, p.sample_size, st, slot);
    if (rc) return rc;
    rc = launch_pack(src, tgt, n, ws, st, gc || g_score_mode != 0);  // the tensor sweep reads the fp16 image only
    if (rc) return rc;
    if (gc) k_gc_mode<<<1, 32, 0, st>>>(ws.ctl);
    // confidence exit: need[r] = smallest best-count that lets the loop stop
    // after round r (conf_iters is non-increasing in the count)
    if (use_conf) {
        std::vector<int> need((size_t)nrounds);
        for (int64_t r = 0; r < nrounds; ++r) {
            int64_t done = (r + 1) * R < p.max_iters ? (r + 1) * R : p.max_iters;
            int64_t lo = 1, hi = n + 1;  // smallest c in [1, n] with conf_iters(c) <= done, else n + 1
            while (lo < hi) {
                int64_t mid = lo + (hi - lo) / 2;
                if (lr_ransac_conf_iters(mid, n, p.sample_size, p.confidence, p.max_iters) <= done) hi = mid;
                else lo = mid + 1;
            }
            need[r] = (int)lo;
        }
        // pageable source: the runtime stages it before returning, `need` may go out of scope
        LR_CUDA_TRY(cudaMemcpyAsync(ws.need, need.data(), sizeof(int) * nrounds, cudaMemcpyHostToDevice, st));
    }
    if (nrounds == 0 && !gc) {  // max_iters == 0: nothing is selected, the identity is the model
        k_set_identity<<<1, 32, 0, st>>>(ws.ctl);
    }
    for (int64_t r = 0; r < nrounds; ++r) {
        int64_t lo = r * R, hi = (r + 1) * R < p.max_iters ? (r + 1) * R : p.max_iters;
        if (gc) {
            rc = launch_round(src, tgt, n, p, ws, lo, hi, nullptr, nullptr, st);
            if (rc) return rc;
            k_round_end_msac<<<1, 32, 0, st>>>(ws.ctl, hi - lo, use_conf ? ws.need : nullptr, (int)r, ws.m64, ws.cnt);
            continue;
        }
        EndArgs ea;
        ea.round_len = hi - lo;
        ea.need = use_conf ? ws.need : nullptr;
        ea.round_idx = (int)r;
        ea.comm = world > 1 ? comm : nullptr;
        // this rank's contiguous slice of the round's ids
        const int64_t a = lo + ((hi - lo) * rank) / world, b = lo + ((hi - lo) * (rank + 1)) / world;
        rc = launch_round(src, tgt, n, p, ws, a, b, nullptr, nullptr, st, nullptr, &ea);
        if (rc) return rc;
    }
    LR_CUDA_TRY(cudaGetLastError());
    if (gc) return gc_enqueue_polish(src, tgt, n, p, ws, st);
    return LR_OK;
}

int identity_result(int64_t n, double *T_out, double *T_refit, uint8_t *mask, LrRansacStats *stats, cudaStream_t st)
{
    identity16(T_out);
    if (T_refit) identity16(T_refit);
    if (mask && n > 0) LR_CUDA_TRY(cudaMemsetAsync(mask, 0, (size_t)n, st));
    if (stats) *stats = LrRansacStats{0, 0, 0, -1, -1, 0, 0, 0, 0, 0, 0};
    LR_CUDA_TRY(cudaStreamSynchronize(st));
    return LR_OK;
}

// one slot of the error-bound probe: the caller's fp64 model becomes survivor `h`
__global__ void k_probe_install(const double *__restrict__ models, int H, const float4 *__restrict__ P8, double thr2,
                                Ctl *ctl, double *__restrict__ m64, uint4 *__restrict__ Aimg, float *__restrict__ band,
                                int *__restrict__ cnt, double *__restrict__ E_out)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h == 0) {
        ctl->n_surv = H;
        ctl->n_events = 0u;
    }
    if (h >= H) return;
    double cen[3], cenq[3], T[12], tt[3];
    tcs::tc_centre(P8, cen, cenq);
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        T[k] = models[(size_t)h * 12 + k];
        m64[(size_t)h * 12 + k] = T[k];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
        tt[a] = (T[4 * a + 3] + ((T[4 * a] * cen[0] + T[4 * a + 1] * cen[1]) + T[4 * a + 2] * cen[2])) - cenq[a];
    tcs::tc_write_model(Aimg, h, T, tt);
    const double Et = tcs::tc_err_bound((double)__uint_as_float(ctl->pt2max_bits), (double)__uint_as_float(ctl->qtmax_bits),
                                        fmax(fabs(tt[0]), fmax(fabs(tt[1]), fabs(tt[2]))));
    const double u = 5.9604644775390625e-08, thr = sqrt(thr2);
    band[h] = __double2float_ru(4.0 * Et * thr + 4.0 * Et * Et + 8.0 * u * thr2 + 1e-9);
    cnt[h] = 0;
    if (E_out) E_out[h] = Et;
}

#include "lr_icp.cuh"

}  // namespace

#ifdef LR_TCS_TRACE
// trace builds only (tools/tcs_trace.py): the pipeline timestamps CTA 0 of the last tensor sweep recorded
LR_EXPORT int lr_debug_tcs_trace(long long *out, int64_t bytes)
{
    LR_REQUIRE(out && bytes == (int64_t)sizeof(tcs::g_tcs_trace), "buffer size mismatch");
    LR_CUDA_TRY(cudaMemcpyFromSymbol(out, tcs::g_tcs_trace, sizeof(tcs::g_tcs_trace)));
    return LR_OK;
}
#endif

LR_EXPORT int lr_ransac_set_mode(int mode)
{
    LR_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0, 1 or 2");
    g_score_mode = mode;
    return LR_OK;
}

LR_EXPORT int64_t lr_ransac_conf_iters(int64_t c, int64_t n, int m, double conf, int64_t max_iters)
{
    // Open3D stopping rule; same expression as the oracle's lro_conf_iters
    if (!(conf < 1.0) || c <= 0 || n <= 0) return max_iters;
    double fitness = (double)c / (double)n;
    double denom = log(1.0 - pow(fitness, (double)m));
    if (!(denom < 0.0)) return max_iters;
    double k = log(1.0 - conf) / denom;
    if (!(k < (double)max_iters)) return max_iters;
    int64_t ki = (int64_t)ceil(k);
    return ki < 1 ? 1 : ki;
}

LR_EXPORT int lr_ransac_rigid(const float *src, const float *tgt, int64_t n, const LrRansacParams *params,
                              double *T_out, double *T_refit, uint8_t *mask, LrRansacStats *stats, void *stream)
{
    lr::Lock lock;
    int rc = check_params(params, n);
    if (rc) return rc;
    LR_REQUIRE(T_out != nullptr, "T_out is null");
    LR_REQUIRE(n == 0 || (src && tgt), "src/tgt is null");
    cudaStream_t st = (cudaStream_t)stream;
    const LrRansacParams &p = *params;
    if (n < p.sample_size) return identity_result(n, T_out, T_refit, mask, stats, st);  // Open3D: |corres| < ransac_n (App. B)
    Ws ws;
    rc = enqueue_run(src, tgt, n, p, lr::SLOT_RANSAC, ws, st, nullptr, g_dbg_rank, g_dbg_world);
    if (rc) return rc;
    return finish(src, tgt, n, p, ws, p.scoring == LR_SCORE_MSAC ? FIN_GC : FIN_MODEL_READY, 0, T_out, T_refit, mask, stats, st);
}

// timing aid (tools/pair_breakdown.py): lr_ransac_rigid then generates and scores only rank's slice of every round, with no
// exchange -- what ONE rank of a hypothesis-sharded run executes, measurable on a single GPU.  (0, 1) restores the default.
LR_EXPORT int lr_debug_slice(int rank, int world)
{
    LR_REQUIRE(world >= 1 && rank >= 0 && rank < world, "rank / world out of range");
    g_dbg_rank = rank;
    g_dbg_world = world;
    return LR_OK;
}

// ---- hypothesis sharding with a library-owned communicator over NVLink peer memory --------------------------
LR_EXPORT int lr_comm_init(int rank, int world, void *handle_out)
{
    lr::Lock lock;
    LR_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "rank / world out of range (at most 16 ranks)");
    LR_REQUIRE(handle_out != nullptr, "handle_out is null");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the C ABI exchanges 64-byte handles");
    BatchCtx *ctx = nullptr;
    int rc = dev_ctx(&ctx);
    if (rc) return rc;
    CommHost &c = ctx->comm;
    LR_REQUIRE(!c.ready && !c.box, "communicator already initialised on this device (lr_comm_destroy first)");
    LR_CUDA_TRY(cudaMalloc(&c.box, sizeof(Mail) * 2 * kMaxRanks));
    LR_CUDA_TRY(cudaMemset(c.box, 0, sizeof(Mail) * 2 * kMaxRanks));
    LR_CUDA_TRY(cudaMalloc(&c.dev, sizeof(Comm)));
    c.rank = rank;
    c.world = world;
    cudaIpcMemHandle_t h;
    memset(&h, 0, sizeof(h));
    if (world > 1) LR_CUDA_TRY(cudaIpcGetMemHandle(&h, c.box));
    memcpy(handle_out, &h, sizeof(h));
    return LR_OK;
}

LR_EXPORT int lr_comm_connect(const void *all_handles)
{
    lr::Lock lock;
    BatchCtx *ctx = nullptr;
    int rc = dev_ctx(&ctx);
    if (rc) return rc;
    CommHost &c = ctx->comm;
    LR_REQUIRE(c.box && !c.ready, "lr_comm_init has not run on this device (or the communicator is already connected)");
    LR_REQUIRE(all_handles != nullptr || c.world == 1, "all_handles is null");
    Comm h;
    memset(&h, 0, sizeof(h));
    h.rank = c.rank;
    h.world = c.world;
    for (int g = 0; g < c.world; ++g) {
        if (g == c.rank) {
            c.peer[g] = c.box;
        } else {
            cudaIpcMemHandle_t hd;
            memcpy(&hd, reinterpret_cast<const char *>(all_handles) + 64 * (size_t)g, sizeof(hd));
            LR_CUDA_TRY(cudaIpcOpenMemHandle(&c.peer[g], hd, cudaIpcMemLazyEnablePeerAccess));
        }
        h.peer[g] = reinterpret_cast<Mail *>(c.peer[g]);
    }
    LR_CUDA_TRY(cudaMemcpy(c.dev, &h, sizeof(h), cudaMemcpyHostToDevice));
    LR_CUDA_TRY(cudaDeviceSynchronize());
    c.ready = true;
    return LR_OK;
}

LR_EXPORT int lr_comm_info(int *rank, int *world)
{
    lr::Lock lock;
    BatchCtx *ctx = nullptr;
    int rc = dev_ctx(&ctx);
    if (rc) return rc;
    if (rank) *rank = ctx->comm.ready ? ctx->comm.rank : 0;
    if (world) *world = ctx->comm.ready ? ctx->comm.world : 0;  // 0: no communicator on this device
    return LR_OK;
}

LR_EXPORT int lr_comm_destroy(void)
{
    lr::Lock lock;
    BatchCtx *ctx = nullptr;
    int rc = dev_ctx(&ctx);
    if (rc) return rc;
    CommHost &c = ctx->comm;
    cudaDeviceSynchronize();
    for (int g = 0; g < c.world; ++g)
        if (g != c.rank && c.peer[g]) cudaIpcCloseMemHandle(c.peer[g]);
    if (c.box) cudaFree(c.box);
    if (c.dev) cudaFree(c.dev);
    c = CommHost();
    return LR_OK;
}

LR_EXPORT int lr_ransac_rigid_sharded(const float *src, const float *tgt, int64_t n, const LrRansacParams *params,
                                      double *T_out, double *T_refit, uint8_t *mask, LrRansacStats *stats, void *stream)
{
    lr::Lock lock;
    int rc = check_params(params, n);
    if (rc) return rc;
    LR_REQUIRE(T_out != nullptr, "T_out is null");
    LR_REQUIRE(n == 0 || (src && tgt), "src/tgt is null");
    LR_REQUIRE(params->scoring == LR_SCORE_COUNT, "hypothesis sharding packs (count, id): LR_SCORE_COUNT only");
    BatchCtx *ctx = nullptr;
    rc = dev_ctx(&ctx);
    if (rc) return rc;
    LR_REQUIRE(ctx->comm.ready, "no communicator on this device: lr_comm_init + lr_comm_connect first");
    cudaStream_t st = (cudaStream_t)stream;
    const LrRansacParams &p = *params;
    if (n < p.sample_size) return identity_result(n, T_out, T_refit, mask, stats, st);
    Ws ws;
    rc = enqueue_run(src, tgt, n, p, lr::SLOT_RANSAC, ws, st, ctx->comm.dev, ctx->comm.rank, ctx->comm.world);
    if (rc) return rc;
    return finish(src, tgt, n, p, ws, FIN_MODEL_READY, 0, T_out, T_refit, mask, stats, st);
}

// ---- error-bound probe of the tensor-core sweep -------------------------------------------------------------
LR_EXPORT int lr_ransac_tc_probe(const float *src, const float *tgt, int64_t n, const double *models, int64_t H,
                                 double threshold, float *d_out, double *E_out, int32_t *counts_out, void *stream)
{
    lr::Lock lock;
    LR_REQUIRE(src && tgt && models, "null pointer");
    LR_REQUIRE(n > 0 && n < (int64_t)1 << 31 && H > 0 && H <= ((int64_t)1 << 20), "n / H out of range");
    LR_REQUIRE(threshold > 0.0, "threshold must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    Ws ws;
    int rc = ws_setup(n, H, 1, ws);
    if (rc) return rc;
    rc = launch_pack(src, tgt, n, ws, st);
    if (rc) return rc;
    const double thr2 = threshold * threshold;
    const int sms = lr::sm_count();
    LR_CUDA_TRY(tc_smem_attr());
    k_probe_install<<<(int)((H + 127) / 128), 128, 0, st>>>(models, (int)H, ws.P8, thr2, ws.ctl, ws.m64, ws.Aimg, ws.band,
                                                           ws.cnt, E_out);
    if (d_out)
        tcs::k_score_tc<true><<<sms, tcs::NTHREADS, tcs::kSmemBytes, st>>>(ws.Aimg, ws.Bimg, ws.P8, n, ws.n_pad, ws.ctl, ws.m64,
                                                                            ws.band, ws.cnt, thr2, d_out, ws.events, kTcEventCap);
    if (counts_out) {
        tcs::k_score_tc<false><<<sms, tcs::NTHREADS, tcs::kSmemBytes, st>>>(ws.Aimg, ws.Bimg, ws.P8, n, ws.n_pad, ws.ctl, ws.m64,
                                                                             ws.band, ws.cnt, thr2, nullptr, ws.events, kTcEventCap);
        LR_CUDA_TRY(cudaMemcpyAsync(counts_out, ws.cnt, sizeof(int32_t) * H, cudaMemcpyDeviceToDevice, st));
    }
    LR_CUDA_TRY(cudaGetLastError());
    LR_CUDA_TRY(cudaStreamSynchronize(st));
    return LR_OK;
}

LR_EXPORT int lr_ransac_rigid_batch(const float *const *src, const float *const *tgt, const int64_t *n, int count,
                                    const LrRansacParams *params, double *T_out, double *T_refit, LrRansacStats *stats,
                                    void *stream)
{
    lr::Lock lock;
    LR_REQUIRE(count >= 0 && (count == 0 || (src && tgt && n && T_out)), "null argument");
    if (count == 0) return LR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t n_max = 0;
    for (int i = 0; i < count; ++i) {
        int rc = check_params(params, n[i]);
        if (rc) return rc;
        LR_REQUIRE(n[i] == 0 || (src[i] && tgt[i]), "src/tgt is null");
        n_max = n[i] > n_max ? n[i] : n_max;
    }
    const LrRansacParams &p = *params;
    const bool want_refit = (T_refit != nullptr) && p.refit;
    BatchCtx *ctx = nullptr;
    int rc = batch_ctx(count, &ctx);
    if (rc) return rc;
    // Two pairs are in flight at any time, each with its own scratch and stream: the single-block tail of one
    // pair (round end, model, refit) runs under the other pair's chip-filling kernels, and there is no host
    // round trip between pairs.  Grow both arenas for the largest pair first: growing frees the old block.
    {
        const bool use_conf = p.confidence < 1.0 && p.max_iters > 0;
        const int64_t R = use_conf ? (int64_t)p.round_size : batch_len(p.max_iters);
        Ws tmp;
        for (int slot : {(int)lr::SLOT_RANSAC, (int)lr::SLOT_RANSAC_B}) {
            rc = ws_setup(n_max, R, (p.max_iters + R - 1) / R, tmp, p.sampler == LR_SAMPLER_PROSAC, slot,
                          p.scoring == LR_SCORE_MSAC);
            if (rc) return rc;
        }
    }
    LR_CUDA_TRY(cudaEventRecord(ctx->start, st));
    for (int l = 0; l < 2; ++l) LR_CUDA_TRY(cudaStreamWaitEvent(ctx->lane[l], ctx->start, 0));
    // any failure below leaves work in flight on the lanes (it writes into ctx->host): join them before returning
    auto bail = [&](int code) {
        for (int l = 0; l < 2; ++l) cudaStreamSynchronize(ctx->lane[l]);
        return code;
    };
    const bool use_conf_b = p.confidence < 1.0 && p.max_iters > 0;
    const int64_t R_b = use_conf_b ? (int64_t)p.round_size : batch_len(p.max_iters);
    for (int i = 0; i < count; ++i) {
        if (n[i] < p.sample_size) continue;  // identity, filled in below
        const int l = i & 1;
        const int slot = l ? lr::SLOT_RANSAC_B : lr::SLOT_RANSAC;
        const float *s_in = src[i], *t_in = tgt[i];
        // A pair handed over in HOST memory (pinned or pageable) is brought in by the copy engine on its lane, so the
        // transfer of pair i + 1 runs under the kernels of pair i on the other lane (k_pack reading pinned memory in place,
        // as the single-pair entry does, would hold SMs for the whole PCIe round trip).
        cudaPointerAttributes at_s, at_t;
        const bool host_s = cudaPointerGetAttributes(&at_s, s_in) != cudaSuccess || at_s.type == cudaMemoryTypeHost ||
                            at_s.type == cudaMemoryTypeUnregistered;
        const bool host_t = cudaPointerGetAttributes(&at_t, t_in) != cudaSuccess || at_t.type == cudaMemoryTypeHost ||
                            at_t.type == cudaMemoryTypeUnregistered;
        (void)cudaGetLastError();  // an unregistered pointer makes older runtimes report an error: it only means "host"
        if (host_s || host_t) {
            Ws tmp;
            rc = ws_setup(n[i], R_b, (p.max_iters + R_b - 1) / R_b, tmp, p.sampler == LR_SAMPLER_PROSAC, slot,
                          p.scoring == LR_SCORE_MSAC);
            if (rc) return bail(rc);
            const size_t bytes = sizeof(float) * 3 * (size_t)n[i];
            if (host_s) {
                if (cudaMemcpyAsync(tmp.stage, s_in, bytes, cudaMemcpyHostToDevice, ctx->lane[l]) != cudaSuccess) {
                    lr::set_error("lr_ransac_rigid_batch: host -> device copy of a source array failed");
                    return bail(LR_ERR_CUDA);
                }
                s_in = tmp.stage;
            }
            if (host_t) {
                if (cudaMemcpyAsync(tmp.stage + 3 * (size_t)n[i], t_in, bytes, cudaMemcpyHostToDevice, ctx->lane[l]) != cudaSuccess) {
                    lr::set_error("lr_ransac_rigid_batch: host -> device copy of a target array failed");
                    return bail(LR_ERR_CUDA);
                }
                t_in = tmp.stage + 3 * (size_t)n[i];
            }
        }
        Ws ws;
        rc = enqueue_run(s_in, t_in, n[i], p, slot, ws, ctx->lane[l]);
        if (rc) return bail(rc);
        rc = finish_launch(s_in, t_in, n[i], p, ws, p.scoring == LR_SCORE_MSAC ? FIN_GC : FIN_MODEL_READY, 0, want_refit,
                           nullptr, ctx->lane[l], true, &ctx->host[i]);
        if (rc) return bail(rc);
    }
    for (int l = 0; l < 2; ++l) {
        if (cudaEventRecord(ctx->done[l], ctx->lane[l]) != cudaSuccess || cudaStreamWaitEvent(st, ctx->done[l], 0) != cudaSuccess) {
            lr::set_error("lr_ransac_rigid_batch: joining the internal streams failed");
            return bail(LR_ERR_CUDA);
        }
    }
    LR_CUDA_TRY(cudaStreamSynchronize(st));
    for (int i = 0; i < count; ++i) {
        double *To = T_out + 16 * (size_t)i, *Tr = T_refit ? T_refit + 16 * (size_t)i : nullptr;
        LrRansacStats *s = stats ? stats + i : nullptr;
        if (n[i] < p.sample_size) {  // Open3D: |corres| < ransac_n -> identity (App. B)
            identity16(To);
            if (Tr) identity16(Tr);
            if (s) *s = LrRansacStats{0, 0, 0, -1, -1, 0, 0, 0, 0, 0, 0};
        } else {
            finish_read(ctx->host[i], want_refit, To, Tr, s);
        }
    }
    return LR_OK;
}

LR_EXPORT int lr_ransac_score_samples(const float *src, const float *tgt, int64_t n, const int32_t *samples, int64_t H,
                                      int m, double threshold, int use_elc, double elc_ratio, int32_t *counts,
                                      double *models, int64_t *best, void *stream)
{
    lr::Lock lock;
    LR_REQUIRE(m == 3 || m == 4, "m must be 3 or 4");
    LR_REQUIRE(src && tgt && samples && counts, "null pointer");
    LR_REQUIRE(n > 0 && n < (int64_t)1 << 31 && H > 0 && H < (int64_t)0xFFFFFFFFLL, "n/H out of range");
    LR_REQUIRE(threshold > 0.0, "threshold must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    LrRansacParams p;
    memset(&p, 0, sizeof(p));
    p.threshold = threshold;
    p.confidence = 1.0;
    p.elc_ratio = elc_ratio;
    p.max_iters = H;
    p.sample_size = m;
    p.sampler = LR_SAMPLER_UNIFORM;
    p.use_elc = use_elc;
    const int64_t R = batch_len(H);
    p.round_size = (int32_t)R;
    Ws ws;
    int rc = ws_setup(n, R, 1, ws);
    if (rc) return rc;
    rc = launch_pack(src, tgt, n, ws, st);
    if (rc) return rc;
    LR_CUDA_TRY(cudaMemsetAsync(counts, 0xFF, sizeof(int32_t) * H, st));  // -1 = rejected
    if (models) LR_CUDA_TRY(cudaMemsetAsync(models, 0, sizeof(double) * 12 * H, st));
    for (int64_t lo = 0; lo < H; lo += R) {
        int64_t hi = lo + R < H ? lo + R : H;
        rc = launch_round(src, tgt, n, p, ws, lo, hi, samples + lo * m, counts + lo, st);
        if (rc) return rc;
        if (models) {
            // scatter the compacted fp64 models back to sample order before the slots are reused
            k_scatter_models<<<(int)((hi - lo + 255) / 256), 256, 0, st>>>(ws.ctl, ws.slot_id, ws.m64, models);
        }
        k_round_end<<<1, 32, 0, st>>>(ws.ctl, hi - lo, nullptr, 0, nullptr);
    }
    LR_CUDA_TRY(cudaGetLastError());
    Ctl h;
    LR_CUDA_TRY(cudaMemcpyAsync(&h, ws.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
    LR_CUDA_TRY(cudaStreamSynchronize(st));
    if (best) *best = h.best_key ? (int64_t)(0xFFFFFFFFu - (uint32_t)(h.best_key & 0xFFFFFFFFULL)) : -1;
    return LR_OK;
}

LR_EXPORT int lr_ransac_score_samples_msac(const float *src, const float *tgt, int64_t n, const int32_t *samples,
                                           int64_t H, int m, double threshold, int use_elc, double elc_ratio,
                                           int64_t *scores, int32_t *inliers, int64_t *best, void *stream)
{
    lr::Lock lock;
    LR_REQUIRE(m == 3 || m == 4, "m must be 3 or 4");
    LR_REQUIRE(src && tgt && samples && scores, "null pointer");
    LR_REQUIRE(n > 0 && n < (int64_t)1 << 31 && H > 0 && H < (int64_t)0xFFFFFFFFLL, "n/H out of range");
    LR_REQUIRE(threshold > 0.0, "threshold must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    LrRansacParams p;
    memset(&p, 0, sizeof(p));
    p.threshold = threshold;
    p.confidence = 1.0;
    p.elc_ratio = elc_ratio;
    p.max_iters = H;
    p.sample_size = m;
    p.sampler = LR_SAMPLER_UNIFORM;
    p.use_elc = use_elc;
    p.scoring = LR_SCORE_MSAC;
    const int64_t R = batch_len(H);
    p.round_size = (int32_t)R;
    Ws ws;
    int rc = ws_setup(n, R, 1, ws, false, lr::SLOT_RANSAC, true);
    if (rc) return rc;
    rc = launch_pack(src, tgt, n, ws, st);
    if (rc) return rc;
    k_gc_mode<<<1, 32, 0, st>>>(ws.ctl);
    LR_CUDA_TRY(cudaMemsetAsync(scores, 0xFF, sizeof(int64_t) * H, st));  // -1 = rejected
    if (inliers) LR_CUDA_TRY(cudaMemsetAsync(inliers, 0xFF, sizeof(int32_t) * H, st));
    for (int64_t lo = 0; lo < H; lo += R) {
        int64_t hi = lo + R < H ? lo + R : H;
        rc = launch_round(src, tgt, n, p, ws, lo, hi, samples + lo * m, inliers ? inliers + lo : nullptr, st,
                          scores + lo);
        if (rc) return rc;
        k_round_end_msac<<<1, 32, 0, st>>>(ws.ctl, hi - lo, nullptr, 0, ws.m64, ws.cnt);
    }
    LR_CUDA_TRY(cudaGetLastError());
    Ctl h;
    LR_CUDA_TRY(cudaMemcpyAsync(&h, ws.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
    LR_CUDA_TRY(cudaStreamSynchronize(st));
    if (best) *best = h.gc.has_model ? (int64_t)h.gc.best_id : -1;
    return LR_OK;
}

LR_EXPORT int lr_ransac_shard(const float *src, const float *tgt, int64_t n, const LrRansacParams *params,
                              int64_t id_lo, int64_t id_hi, uint64_t *key, void *stream)
{
    lr::Lock lock;
    int rc = check_params(params, n);
    if (rc) return rc;
    LR_REQUIRE(src && tgt && key, "null pointer");
    LR_REQUIRE(params->scoring == LR_SCORE_COUNT, "hypothesis sharding packs (count, id): LR_SCORE_COUNT only");
    LR_REQUIRE(n >= params->sample_size, "fewer correspondences than the sample size");
    LR_REQUIRE(id_lo >= 0 && id_hi >= id_lo && id_hi < (int64_t)0xFFFFFFFFLL, "id range out of bounds");
    cudaStream_t st = (cudaStream_t)stream;
    const LrRansacParams &p = *params;
    const int64_t R = batch_len(id_hi - id_lo);
    Ws ws;
    rc = ws_setup(n, R, 1, ws, p.sampler == LR_SAMPLER_PROSAC);
    if (rc) return rc;
    rc = upload_growth(ws, n, p.sample_size, st);
    if (rc) return rc;
    rc = launch_pack(src, tgt, n, ws, st);
    if (rc) return rc;
    for (int64_t lo = id_lo; lo < id_hi; lo += R) {
        int64_t hi = lo + R < id_hi ? lo + R : id_hi;
        rc = launch_round(src, tgt, n, p, ws, lo, hi, nullptr, nullptr, st);
        if (rc) return rc;
        k_round_end<<<1, 32, 0, st>>>(ws.ctl, hi - lo, nullptr, 0, (unsigned long long *)key);
    }
    LR_CUDA_TRY(cudaGetLastError());
    return LR_OK;
}

LR_EXPORT int lr_ransac_finalize(const float *src, const float *tgt, int64_t n, const LrRansacParams *params,
                                 uint64_t key, double *T_out, double *T_refit, uint8_t *mask, LrRansacStats *stats,
                                 void *stream)
{
    lr::Lock lock;
    int rc = check_params(params, n);
    if (rc) return rc;
    LR_REQUIRE(src && tgt && T_out, "null pointer");
    LR_REQUIRE(params->scoring == LR_SCORE_COUNT, "hypothesis sharding packs (count, id): LR_SCORE_COUNT only");
    LR_REQUIRE(n >= params->sample_size, "fewer correspondences than the sample size");
    cudaStream_t st = (cudaStream_t)stream;
    Ws ws;
    rc = ws_setup(n, params->round_size, 1, ws, params->sampler == LR_SAMPLER_PROSAC);
    if (rc) return rc;
    rc = upload_growth(ws, n, params->sample_size, st);
    if (rc) return rc;
    k_ctl_reset<<<1, 32, 0, st>>>(ws.ctl);
    return finish(src, tgt, n, *params, ws, FIN_FROM_KEY, key, T_out, T_refit, mask, stats, st, /*packed=*/false);
}

LR_EXPORT int lr_ransac_sample(const LrRansacParams *params, int64_t n, int64_t id_lo, int64_t H, int32_t *samples,
                               void *stream)
{
    int rc = check_params(params, n);
    if (rc) return rc;
    LR_REQUIRE(samples && H >= 0 && n >= params->sample_size, "bad arguments");
    if (H == 0) return LR_OK;
    lr::Lock lock;
    cudaStream_t st = (cudaStream_t)stream;
    Ws ws;
    rc = ws_setup(n, 1, 1, ws, params->sampler == LR_SAMPLER_PROSAC);
    if (rc) return rc;
    rc = upload_growth(ws, n, params->sample_size, st);
    if (rc) return rc;
    int blocks = (int)((H + 255) / 256);
    if (params->sample_size == 3)
        k_sample_only<3><<<blocks, 256, 0, st>>>(params->seed, params->sampler, n, ws.growth, id_lo, H, samples);
    else
        k_sample_only<4><<<blocks, 256, 0, st>>>(params->seed, params->sampler, n, ws.growth, id_lo, H, samples);
    LR_CUDA_TRY(cudaGetLastError());
    return LR_OK;
}

LR_EXPORT int lr_transform_pad8(const float *xyz, int64_t n, const double *T_in, float *out, void *stream)
{
    lr::Lock lock;
    LR_REQUIRE(xyz && T_in && out && n >= 0, "bad arguments");
    if (n == 0) return LR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    Ws ws;
    int rc = ws_setup(1, 1, 1, ws);
    if (rc) return rc;
    double T12[12];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) T12[4 * r + c] = T_in[4 * r + c];
    LR_CUDA_TRY(cudaMemcpyAsync(ws.scratchT, T12, sizeof(T12), cudaMemcpyHostToDevice, st));
    LR_CUDA_TRY(cudaStreamSynchronize(st));  // T12 is a local
    k_transform_pad8<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(xyz, n, ws.scratchT, out);
    LR_CUDA_TRY(cudaGetLastError());
    return LR_OK;
}

LR_EXPORT int lr_icp_step(const float *xyz0, const float *xyz1, const int64_t *i0, const int64_t *i1, int64_t K,
                          const double *T_in, double threshold, double *T_out, int64_t *count, double *err2, void *stream)
{
    lr::Lock lock;
    LR_REQUIRE(xyz0 && xyz1 && T_in && T_out, "null pointer");
    LR_REQUIRE(K >= 0 && threshold > 0.0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    Ws ws;
    int rc = ws_setup(1, 1, 1, ws);
    if (rc) return rc;
    double T12[12];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) T12[4 * r + c] = T_in[4 * r + c];
    LR_CUDA_TRY(cudaMemcpyAsync(ws.scratchT, T12, sizeof(T12), cudaMemcpyHostToDevice, st));
    k_set_T<<<1, 32, 0, st>>>(ws.ctl, ws.scratchT);
    BatchCtx *ctx = nullptr;
    rc = dev_ctx(&ctx);
    if (rc) return rc;
    k_finish<<<finish_blocks(K), 256, 0, st>>>(xyz0, xyz1, i0, i1, K, threshold * threshold, ws.ctl, nullptr, ws.partial, 1,
                                               ctx->host1);
    LR_CUDA_TRY(cudaGetLastError());
    LR_CUDA_TRY(cudaStreamSynchronize(st));
    const Ctl &h = *ctx->host1;
    T12_to_16(h.Tref, T_out);
    if (count) *count = h.refit_count;
    if (err2) *err2 = h.err2;
    return LR_OK;
}

LR_EXPORT int lr_refit_indexed(const float *xyz0, const float *xyz1, const int64_t *i0, const int64_t *i1, int64_t K,
                               const double *T_in, double threshold, double *T_out, int64_t *count, void *stream)
{
    lr::Lock lock;
    LR_REQUIRE(xyz0 && xyz1 && T_in && T_out, "null pointer");
    LR_REQUIRE(K >= 0 && threshold > 0.0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    Ws ws;
    int rc = ws_setup(1, 1, 1, ws);
    if (rc) return rc;
    double *dT = ws.scratchT;
    double T12[12];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) T12[4 * r + c] = T_in[4 * r + c];
    LR_CUDA_TRY(cudaMemcpyAsync(dT, T12, sizeof(T12), cudaMemcpyHostToDevice, st));
    k_set_T<<<1, 32, 0, st>>>(ws.ctl, dT);
    BatchCtx *ctx = nullptr;
    rc = dev_ctx(&ctx);
    if (rc) return rc;
    k_finish<<<finish_blocks(K), 256, 0, st>>>(xyz0, xyz1, i0, i1, K, threshold * threshold, ws.ctl, nullptr, ws.partial, 1,
                                               ctx->host1);
    LR_CUDA_TRY(cudaGetLastError());
    LR_CUDA_TRY(cudaStreamSynchronize(st));
    const Ctl &h = *ctx->host1;
    T12_to_16(h.Tref, T_out);
    if (count) *count = h.refit_count;
    return LR_OK;
}

#include "lr_icp_api.cuh"
