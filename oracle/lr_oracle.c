/*
 * lr_oracle.c -- CPU ORACLE for the robust-registration hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package
 * (lidarregistration_b200/) may include, link, import or execute this file.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and there only as the checker / CPU baseline.
 *
 * What it restates (all citations relative to /root/reference):
 *   matching : Experiments/algorithms/matching.py:22-65   (find_nn, knn_dist)
 *              Experiments/algorithms/matching.py:67-87   (torch_intersect)
 *              Experiments/algorithms/matching.py:222-239 (nn_to_mutual)
 *              Experiments/algorithms/matching.py:89-98   (ratio quality)
 *   ELC      : GC-RANSAC/src/pygcransac/include/preemption/preemption_edge_length.h:71-128
 *              (checked against that header itself, compiled in place into oracle/_ref by `make ref`)
 *   RANSAC   : Experiments/algorithms/FR.py:99-111,122-139 (Open3D branch: parameters + refit)
 *              Experiments/algorithms/GC_RANSAC.py:8-55, gcransac_python.cpp:404-624 (GC branch wiring)
 *              The loop arithmetic itself lives in un-vendored third-party code
 *              (open3d==0.13.0, pygcransac==0.1); it is restated from SURVEY.md
 *              Appendix A/B.  PARITY STATUS: matching = PINNED against the
 *              reference's own matching.py run in the build container
 *              (tests/golden/, made by tests/golden/make_golden.py);
 *              RANSAC = "parity unpinned" by any reference fixture (the reference
 *              ships none); pinned only against authored known-answer cases and
 *              an independent numpy/torch Kabsch (Experiments/models/common.py:7-45).
 *
 * Canonical arithmetic (DESIGN.md "Canonical arithmetic"): every floating
 * point operation below is an individually rounded IEEE-754 operation in a
 * fixed order (compile with -ffp-contract=off); the CUDA product follows the
 * same order, so integer results (indices, inlier counts, selected hypothesis)
 * are bit-exact and fp64 models are identical.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LRO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* Matching                                                                  */
/* ------------------------------------------------------------------------- */

/* torch.sum(f**2, dim=1) on CPU (matching.py:29): squares rounded to fp32,
 * accumulated in 8 strided lanes (lane l sums x[l], x[8+l], ... in order) and
 * the 8 lanes are then added sequentially.  Verified bit-for-bit against
 * torch 2.11 CPU for D = 32 (tests/test_oracle_matching.py). */
static float lro_sqnorm(const float *x, int D)
{
    float lane[8];
    for (int l = 0; l < 8; ++l) lane[l] = 0.0f;
    int first = 1;
    int k = 0;
    for (; k + 8 <= D; k += 8) {
        for (int l = 0; l < 8; ++l) {
            float sq = x[k + l] * x[k + l];
            lane[l] = first ? sq : lane[l] + sq;
        }
        first = 0;
    }
    float s = lane[0];
    for (int l = 1; l < 8; ++l) s = s + lane[l];
    for (; k < D; ++k) s = s + x[k] * x[k]; /* D % 8 tail: not used by the path (D = 32) */
    return s;
}

LRO_API void lro_sqnorms(const float *F, int64_t N, int D, float *out)
{
    for (int64_t i = 0; i < N; ++i) out[i] = lro_sqnorm(F + i * D, D);
}

/* One reference distance, matching.py:29-30:
 *   dist2 = (|f0|^2 + |f1|^2) - 2*dot ;  dist = sqrt(max(dist2, 1e-30))
 * dot = sequential fused multiply-add over k (what MKL's sgemm micro-kernel
 * does for K = 32; verified bit-for-bit against torch.einsum on CPU). */
static inline float lro_dist(float n0, float n1, float dot)
{
    float d2 = (n0 + n1) - 2.0f * dot;
    if (!(d2 >= 1e-30f)) d2 = 1e-30f; /* clamp_min (NaN cannot occur on finite input) */
    return sqrtf(d2);
}

/* find_nn (matching.py:22-65).  idx1[i] = argmin_j dist(i,j), lowest j on ties
 * (torch.min returns the first minimum on CPU); idx2 (nullable) = argmin with
 * entry idx1[i] masked to +inf (matching.py:36-37).  d1/d2 (nullable) receive
 * the distances themselves (used by tests only). */
LRO_API void lro_find_nn(const float *F0, int64_t N, const float *F1, int64_t M, int D,
                         int64_t *idx1, int64_t *idx2, float *d1out, float *d2out)
{
    float *n0 = (float *)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
    float *n1 = (float *)malloc(sizeof(float) * (size_t)(M > 0 ? M : 1));
    /* transposed copy of F1 so the j loop vectorises: F1t[k][j] */
    float *F1t = (float *)malloc(sizeof(float) * (size_t)(M > 0 ? M : 1) * (size_t)D);
    lro_sqnorms(F0, N, D, n0);
    lro_sqnorms(F1, M, D, n1);
    for (int64_t j = 0; j < M; ++j)
        for (int k = 0; k < D; ++k) F1t[(size_t)k * M + j] = F1[j * D + k];

#pragma omp parallel
    {
        float *acc = (float *)malloc(sizeof(float) * (size_t)(M > 0 ? M : 1));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            const float *a = F0 + i * D;
            for (int64_t j = 0; j < M; ++j) acc[j] = 0.0f;
            for (int k = 0; k < D; ++k) {
                const float ak = a[k];
                const float *row = F1t + (size_t)k * M;
                for (int64_t j = 0; j < M; ++j) acc[j] = fmaf(ak, row[j], acc[j]);
            }
            float best = INFINITY, second = INFINITY;
            int64_t bj = 0, sj = 0;
            /* torch.min over an all-equal / empty row returns index 0 */
            for (int64_t j = 0; j < M; ++j) {
                float d = lro_dist(n0[i], n1[j], acc[j]);
                if (d < best) {
                    second = best; sj = bj;
                    best = d; bj = j;
                } else if (d < second) {
                    second = d; sj = j;
                }
            }
            /* second-best under "mask the best then first-min" semantics:
             * ties with the best value at a higher index are legitimately
             * second (d < second covers d == best since second > best or inf). */
            if (M == 1) { sj = 0; second = INFINITY; } /* all-inf row -> index 0 */
            idx1[i] = bj;
            if (idx2) idx2[i] = sj;
            if (d1out) d1out[i] = best;
            if (d2out) d2out[i] = second;
        }
        free(acc);
    }
    free(n0); free(n1); free(F1t);
}

/* nn_to_mutual (matching.py:222-239) + torch_intersect (:67-87).
 * Reverse NN is computed for the rows of F1 in unique(idx1) against all of F0;
 * (i, j) survives iff j == idx1[i] and i == NN_0(j).  Output sorted by i
 * ascending (coalesce order).  Returns K. */
LRO_API int64_t lro_mutual(const float *F0, int64_t N, const float *F1, int64_t M, int D,
                           const int64_t *idx1, int64_t *out_i, int64_t *out_j)
{
    int64_t *rev = (int64_t *)malloc(sizeof(int64_t) * (size_t)(M > 0 ? M : 1));
    /* reverse NN for every row of F1: identical on the unique subset */
    lro_find_nn(F1, M, F0, N, D, rev, NULL, NULL, NULL);
    int64_t K = 0;
    for (int64_t i = 0; i < N; ++i) {
        int64_t j = idx1[i];
        if (j >= 0 && j < M && rev[j] == i) { out_i[K] = i; out_j[K] = j; ++K; }
    }
    free(rev);
    return K;
}

/* calc_distance_ratio_in_feature_space (matching.py:89-98):
 *   d = sqrt(sum((A-B)^2)) ; ratio = d1 / (d2 + 1e-6)     (fp32)
 * torch.sum over the inner dim uses the same 8-lane order as lro_sqnorm. */
static float lro_diffnorm(const float *a, const float *b, int D)
{
    float tmp[512];
    for (int k = 0; k < D; ++k) tmp[k] = a[k] - b[k];
    return sqrtf(lro_sqnorm(tmp, D));
}

LRO_API void lro_ratio(const float *F0, const float *F1, int D, int64_t K,
                       const int64_t *i0, const int64_t *i1, const int64_t *i2, float *out)
{
    const float eps = 1e-6f; /* 10**-6 promoted to the tensor dtype */
    for (int64_t k = 0; k < K; ++k) {
        float da = lro_diffnorm(F0 + i0[k] * D, F1 + i1[k] * D, D);
        float db = lro_diffnorm(F0 + i0[k] * D, F1 + i2[k] * D, D);
        out[k] = da / (db + eps);
    }
}

/* ------------------------------------------------------------------------- */
/* Counter-based sampling (pure function of (seed, hypothesis id))           */
/* ------------------------------------------------------------------------- */

static inline uint64_t lro_mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

/* draw d of hypothesis id: uniform integer in [0, m) (multiply-shift on the
 * high 32 bits of the mixed counter) */
static inline uint32_t lro_draw(uint64_t seed, uint64_t id, uint32_t d, uint32_t m)
{
    uint64_t r = lro_mix64(lro_mix64(seed ^ (id * 0xD1342543DE82EF95ULL)) + (uint64_t)d * 0x9E3779B97F4A7C15ULL);
    return (uint32_t)(((r >> 32) * (uint64_t)m) >> 32);
}

/* k unique indices out of [0, n), uniform over k-subsets in draw order: draw d
 * picks the r-th index not yet taken (taken list kept sorted) */
static void lro_unique(uint64_t seed, uint64_t id, int k, int64_t n, int32_t *out)
{
    int32_t taken[4];
    for (int d = 0; d < k; ++d) {
        int32_t r = (int32_t)lro_draw(seed, id, (uint32_t)d, (uint32_t)(n - d));
        for (int e = 0; e < d; ++e)
            if (r >= taken[e]) ++r;
        out[d] = r;
        int e = d;
        while (e > 0 && taken[e - 1] > r) { taken[e] = taken[e - 1]; --e; }
        taken[e] = r;
    }
}

#define LRO_PROSAC_TN 100000 /* ProsacSampler(points, m, T_N = 100 000), SURVEY App. A */

/* PROSAC growth function: growth[n-1] = T'_n, the last draw (1-based)
 * made from the n best correspondences; T_n = T_N prod_{i<m} (n-i)/(N-i),
 * T'_{n+1} = T'_n + ceil(T_{n+1} - T_n); T'_n = 1 for n <= m. */
LRO_API void lro_prosac_growth(int64_t N, int m, uint32_t *growth)
{
    double T_n = (double)LRO_PROSAC_TN;
    for (int i = 0; i < m; ++i) T_n *= (double)(m - i) / (double)(N - i);
    uint32_t T_prime = 1;
    for (int64_t i = 0; i < N; ++i) {
        if (i + 1 <= m) { growth[i] = T_prime; continue; }
        double T_next = (double)(i + 1) * T_n / (double)(i + 1 - m);
        double inc = ceil(T_next - T_n);
        if (!(inc < 4.0e9)) inc = 4.0e9;
        uint64_t g = (uint64_t)T_prime + (uint64_t)inc;
        growth[i] = g > 0xFFFFFFF0ULL ? 0xFFFFFFF0U : (uint32_t)g;
        T_n = T_next;
        T_prime = growth[i];
    }
}

/* sampler ids as in include/lidarreg.h:
 *   0 uniform: m unique indices (GC UniformSampler semantics);
 *   1 PROSAC : draw k = id + 1 takes m-1 unique indices from the first n_k - 1
 *              correspondences plus correspondence n_k - 1, n_k = smallest n with
 *              k <= T'_n; uniform after T_N draws (needs `growth`, inputs sorted best first);
 *   2 replace: m indices with replacement (Open3D semantics). */
LRO_API void lro_sample(uint64_t seed, uint64_t id, int sampler, int m, int64_t n, const uint32_t *growth,
                        int32_t *out)
{
    if (sampler == 2) {
        for (int d = 0; d < m; ++d) out[d] = (int32_t)lro_draw(seed, id, (uint32_t)d, (uint32_t)n);
        return;
    }
    if (sampler == 1 && growth && id + 1 <= LRO_PROSAC_TN) {
        const uint64_t k = id + 1;
        int64_t lo = m, hi = n; /* smallest subset size in [m, n] whose T' covers draw k */
        while (lo < hi) {
            int64_t mid = lo + (hi - lo) / 2;
            if ((uint64_t)growth[mid - 1] >= k) hi = mid;
            else lo = mid + 1;
        }
        lro_unique(seed, id, m - 1, lo - 1, out);
        out[m - 1] = (int32_t)(lo - 1);
        return;
    }
    lro_unique(seed, id, m, n, out);
}

/* ------------------------------------------------------------------------- */
/* Edge-length constraint (preemption_edge_length.h:82-127)                  */
/* ------------------------------------------------------------------------- */

static inline double lro_len3(const double *a, const double *b)
{
    double dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
    return sqrt((dx * dx + dy * dy) + dz * dz);
}

/* P, Q: m x 3 fp64 sample (source / target).  Returns 1 = passes. */
LRO_API int lro_elc(const double *P, const double *Q, int m, double ratio)
{
    for (int i = 0; i < m; ++i)
        for (int j = i + 1; j < m; ++j) {
            double ds = lro_len3(P + 3 * i, P + 3 * j);
            double dt = lro_len3(Q + 3 * i, Q + 3 * j);
            if ((ds < dt * ratio) || (dt < ds * ratio)) return 0;
        }
    return 1;
}

/* ------------------------------------------------------------------------- */
/* Kabsch (canonical): least-squares rigid motion Q ~ R P + t                */
/* ------------------------------------------------------------------------- */

#define LRO_JACOBI_SWEEPS 6

static inline double lro_dot3c(const double *H, int p, int q) /* columns p,q of row-major 3x3 */
{
    return (H[p] * H[q] + H[3 + p] * H[3 + q]) + H[6 + p] * H[6 + q];
}

/* From the 3x3 cross-covariance H = sum (q - cq)(p - cp)^T (row-major) to the
 * proper rotation R maximising trace(R^T H) ... i.e. the Kabsch / umeyama
 * rotation R = U diag(1,1,det(U)det(V)) V^T, computed with a
 * fixed-sweep one-sided Jacobi SVD (H V = U S) and the identity
 *   R = u1 v1^T + u2 v2^T + (u1 x u2)(v1 x v2)^T
 * over the two largest singular pairs, which equals the formula above for
 * every rank >= 2 input and needs no sign logic. */
static void lro_rot_from_H(const double Hin[9], double R[9])
{
    double H[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    memcpy(H, Hin, sizeof(H));
    static const int PP[3] = {0, 0, 1}, QQ[3] = {1, 2, 2};
    for (int sweep = 0; sweep < LRO_JACOBI_SWEEPS; ++sweep)
        for (int r = 0; r < 3; ++r) {
            const int p = PP[r], q = QQ[r];
            double alpha = lro_dot3c(H, p, p);
            double beta = lro_dot3c(H, q, q);
            double gamma = lro_dot3c(H, p, q);
            double c = 1.0, s = 0.0;
            if (gamma != 0.0) {
                double zeta = (beta - alpha) / (2.0 * gamma);
                double az = fabs(zeta);
                double tt = 1.0 / (az + sqrt(1.0 + zeta * zeta));
                if (zeta < 0.0) tt = -tt;
                c = 1.0 / sqrt(1.0 + tt * tt);
                s = c * tt;
            }
            for (int k = 0; k < 3; ++k) {
                double hp = H[3 * k + p], hq = H[3 * k + q];
                H[3 * k + p] = c * hp - s * hq;
                H[3 * k + q] = s * hp + c * hq;
                double vp = V[3 * k + p], vq = V[3 * k + q];
                V[3 * k + p] = c * vp - s * vq;
                V[3 * k + q] = s * vp + c * vq;
            }
        }
    double s2[3] = {lro_dot3c(H, 0, 0), lro_dot3c(H, 1, 1), lro_dot3c(H, 2, 2)};
    int a = 0;
    if (s2[1] > s2[a]) a = 1;
    if (s2[2] > s2[a]) a = 2;
    int b = (a == 0) ? 1 : 0;
    for (int k = 0; k < 3; ++k)
        if (k != a && k != b && s2[k] > s2[b]) b = k;
    if (!(s2[a] > 0.0)) { /* H == 0: all points coincide */
        for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
        return;
    }
    double u1[3], u2[3], u3[3], v1[3], v2[3], v3[3];
    double sa = sqrt(s2[a]);
    for (int k = 0; k < 3; ++k) { u1[k] = H[3 * k + a] / sa; v1[k] = V[3 * k + a]; v2[k] = V[3 * k + b]; }
    if (s2[b] > s2[a] * 1e-30) {
        double sb = sqrt(s2[b]);
        for (int k = 0; k < 3; ++k) u2[k] = H[3 * k + b] / sb;
    } else { /* rank 1 (collinear sample): any direction orthogonal to u1, chosen deterministically */
        int e = 0;
        if (fabs(u1[1]) < fabs(u1[e])) e = 1;
        if (fabs(u1[2]) < fabs(u1[e])) e = 2;
        for (int k = 0; k < 3; ++k) u2[k] = (k == e) ? 1.0 : 0.0;
    }
    /* Gram-Schmidt u2 against u1, renormalise */
    double g = (u1[0] * u2[0] + u1[1] * u2[1]) + u1[2] * u2[2];
    for (int k = 0; k < 3; ++k) u2[k] = u2[k] - g * u1[k];
    double nu = sqrt((u2[0] * u2[0] + u2[1] * u2[1]) + u2[2] * u2[2]);
    for (int k = 0; k < 3; ++k) u2[k] = u2[k] / nu;
    u3[0] = u1[1] * u2[2] - u1[2] * u2[1];
    u3[1] = u1[2] * u2[0] - u1[0] * u2[2];
    u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
    v3[0] = v1[1] * v2[2] - v1[2] * v2[1];
    v3[1] = v1[2] * v2[0] - v1[0] * v2[2];
    v3[2] = v1[0] * v2[1] - v1[1] * v2[0];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            R[3 * r + c] = (u1[r] * v1[c] + u2[r] * v2[c]) + u3[r] * v3[c];
}

/* P, Q: k x 3 fp64.  T: 3x4 row-major [R | t]. */
LRO_API void lro_kabsch(const double *P, const double *Q, int64_t k, double T[12])
{
    double cp[3] = {0, 0, 0}, cq[3] = {0, 0, 0};
    if (k <= 0) { /* empty set -> identity (Open3D ComputeTransformation, App. B) */
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) T[4 * r + c] = (r == c) ? 1.0 : 0.0;
        return;
    }
    for (int64_t i = 0; i < k; ++i)
        for (int c = 0; c < 3; ++c) { cp[c] = cp[c] + P[3 * i + c]; cq[c] = cq[c] + Q[3 * i + c]; }
    for (int c = 0; c < 3; ++c) { cp[c] = cp[c] / (double)k; cq[c] = cq[c] / (double)k; }
    double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = 0; i < k; ++i) {
        double dp[3], dq[3];
        for (int c = 0; c < 3; ++c) { dp[c] = P[3 * i + c] - cp[c]; dq[c] = Q[3 * i + c] - cq[c]; }
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) H[3 * r + c] = H[3 * r + c] + dq[r] * dp[c];
    }
    double R[9];
    lro_rot_from_H(H, R);
    for (int r = 0; r < 3; ++r) {
        T[4 * r + 0] = R[3 * r + 0];
        T[4 * r + 1] = R[3 * r + 1];
        T[4 * r + 2] = R[3 * r + 2];
        T[4 * r + 3] = cq[r] - ((R[3 * r + 0] * cp[0] + R[3 * r + 1] * cp[1]) + R[3 * r + 2] * cp[2]);
    }
}

/* canonical squared residual |R p + t - q|^2 */
static inline double lro_res2(const double T[12], const float *p, const float *q)
{
    double px = p[0], py = p[1], pz = p[2];
    double d0 = (((T[0] * px + T[1] * py) + T[2] * pz) + T[3]) - (double)q[0];
    double d1 = (((T[4] * px + T[5] * py) + T[6] * pz) + T[7]) - (double)q[1];
    double d2 = (((T[8] * px + T[9] * py) + T[10] * pz) + T[11]) - (double)q[2];
    return (d0 * d0 + d1 * d1) + d2 * d2;
}

/* Inlier count of T over n correspondences: dis^2 < thr^2 strict (App. B). */
LRO_API int64_t lro_count_inliers(const float *src, const float *tgt, int64_t n, const double T[12],
                                  double thr, uint8_t *mask)
{
    const double thr2 = thr * thr;
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i) {
        int in = lro_res2(T, src + 3 * i, tgt + 3 * i) < thr2;
        cnt += in;
        if (mask) mask[i] = (uint8_t)in;
    }
    return cnt;
}

/* "Faithful" Open3D evaluation (FR.py:122-139 -> RegistrationRANSACBasedOnCorrespondence): for
 * every hypothesis that passes the checkers Open3D copies the WHOLE source cloud, transforms the copy and then
 * walks the correspondences (EvaluateRANSACBasedOnCorrespondence also accumulates the squared error for the rmse).
 * Same arithmetic per residual as lro_res2, so the counts are those of lro_count_inliers; what changes is the
 * memory traffic per hypothesis -- this variant exists for the CPU timing of bench.py ("faithful" vs "lean").
 * cloud: xyz of the source cloud [cloud_n,3]; src_idx (nullable = identity): row of the cloud of correspondence i. */
static int g_o3d_faithful = 0;
LRO_API void lro_set_o3d_faithful(int on) { g_o3d_faithful = on; }
LRO_API int lro_get_o3d_faithful(void) { return g_o3d_faithful; }

static int64_t lro_count_inliers_faithful(const float *cloud, int64_t cloud_n, const float *tgt, int64_t n,
                                          const double T[12], double thr, double *buf, double *err2_out)
{
    for (int64_t i = 0; i < cloud_n; ++i) { /* pcd = source; pcd.Transform(transformation) */
        double px = cloud[3 * i], py = cloud[3 * i + 1], pz = cloud[3 * i + 2];
        buf[3 * i + 0] = ((T[0] * px + T[1] * py) + T[2] * pz) + T[3];
        buf[3 * i + 1] = ((T[4] * px + T[5] * py) + T[6] * pz) + T[7];
        buf[3 * i + 2] = ((T[8] * px + T[9] * py) + T[10] * pz) + T[11];
    }
    const double thr2 = thr * thr;
    int64_t cnt = 0;
    double err2 = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double d0 = buf[3 * i + 0] - (double)tgt[3 * i + 0];
        double d1 = buf[3 * i + 1] - (double)tgt[3 * i + 1];
        double d2 = buf[3 * i + 2] - (double)tgt[3 * i + 2];
        double r2 = (d0 * d0 + d1 * d1) + d2 * d2;
        if (r2 < thr2) { ++cnt; err2 += r2; }
    }
    if (err2_out) *err2_out = err2;
    return cnt;
}

/* MSAC score (App. A): sum over r^2 < tau^2 of (1 - r^2/tau^2), tau = 1.5*thr */
LRO_API double lro_msac(const float *src, const float *tgt, int64_t n, const double T[12], double thr,
                        int64_t *inliers_out)
{
    const double tau2 = (1.5 * thr) * (1.5 * thr);
    double v = 0.0;
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i) {
        double r2 = lro_res2(T, src + 3 * i, tgt + 3 * i);
        if (r2 < tau2) { v = v + (1.0 - r2 / tau2); ++cnt; }
    }
    if (inliers_out) *inliers_out = cnt;
    return v;
}

/* model of one minimal sample: returns 1 if it passes ELC (or ELC off) */
static int lro_model_from_sample(const float *src, const float *tgt, const int32_t *s, int m,
                                 int use_elc, double elc_ratio, double T[12])
{
    double P[12], Q[12];
    for (int d = 0; d < m; ++d)
        for (int c = 0; c < 3; ++c) {
            P[3 * d + c] = (double)src[3 * (int64_t)s[d] + c];
            Q[3 * d + c] = (double)tgt[3 * (int64_t)s[d] + c];
        }
    if (use_elc && !lro_elc(P, Q, m, elc_ratio)) return 0;
    lro_kabsch(P, Q, m, T);
    return 1;
}

/* Fed-sample parity hook (north_star: "same fed hypothesis triplets").
 * samples: H x m int32.  counts[h] = inlier count, or -1 when ELC rejects.
 * models (nullable): H x 12 fp64.  Returns the selected hypothesis
 * (max count, ties -> lowest h), or -1 if none passed. */
LRO_API int64_t lro_score_samples(const float *src, const float *tgt, int64_t n, const int32_t *samples,
                                  int64_t H, int m, double thr, int use_elc, double elc_ratio,
                                  int32_t *counts, double *models)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t h = 0; h < H; ++h) {
        double T[12];
        int ok = lro_model_from_sample(src, tgt, samples + h * m, m, use_elc, elc_ratio, T);
        counts[h] = ok ? (int32_t)lro_count_inliers(src, tgt, n, T, thr, NULL) : -1;
        if (models) {
            if (ok) memcpy(models + 12 * h, T, sizeof(T));
            else memset(models + 12 * h, 0, sizeof(T));
        }
    }
    int64_t best = -1;
    int32_t bc = -1;
    for (int64_t h = 0; h < H; ++h)
        if (counts[h] > bc) { bc = counts[h]; best = h; }
    return best;
}

/* Confidence stopping rule (Open3D App. B): number of hypotheses after which a
 * best inlier count c lets the loop stop. */
LRO_API int64_t lro_conf_iters(int64_t c, int64_t n, int m, double conf, int64_t max_iters)
{
    if (!(conf < 1.0) || c <= 0 || n <= 0) return max_iters;
    double fitness = (double)c / (double)n;
    double denom = log(1.0 - pow(fitness, (double)m));
    if (!(denom < 0.0)) return max_iters; /* fitness^m underflowed to 0 */
    double k = log(1.0 - conf) / denom;
    if (!(k < (double)max_iters)) return max_iters;
    int64_t ki = (int64_t)ceil(k);
    return ki < 1 ? 1 : ki;
}

typedef struct {
    int64_t iters_run;   /* hypotheses generated */
    int64_t n_passed;    /* hypotheses that passed ELC (= scored) */
    int64_t best_id;     /* selected hypothesis id, -1 if none */
    int64_t best_count;  /* its inlier count */
    int64_t refit_count; /* inliers used by the final refit */
} LroStats;

/* The full loop, "count" scoring (Open3D semantics, the graded criterion of
 * SURVEY 8(a)): hypotheses id = 0,1,... are pure functions of (seed, id);
 * selected = max inlier count, ties -> lowest id; the confidence exit is
 * evaluated at round boundaries (every `round` hypotheses) so the result does
 * not depend on thread or GPU count.  T: 3x4 of the selected model (identity
 * if none); Trefit (nullable): Kabsch over its inliers (FR.py:99-111 applied
 * to the same correspondence set). */
LRO_API void lro_ransac(const float *src, const float *tgt, int64_t n, int m, int sampler, int use_elc,
                        double elc_ratio, double thr, double conf, int64_t max_iters, int64_t round,
                        uint64_t seed, double T[12], double *Trefit, uint8_t *mask, LroStats *st)
{
    int64_t best_id = -1, best_cnt = -1, passed = 0, done = 0;
    double bestT[12];
    uint32_t *growth = NULL;
    if (sampler == 1 && n >= m) {
        growth = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n);
        lro_prosac_growth(n, m, growth);
    }
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) bestT[4 * r + c] = (r == c) ? 1.0 : 0.0;
    if (n >= m && round > 0) {
        while (done < max_iters) {
            int64_t lo = done, hi = done + round < max_iters ? done + round : max_iters;
            int64_t r_id = -1, r_cnt = -1, r_pass = 0;
#pragma omp parallel
            {
                int64_t t_id = -1, t_cnt = -1, t_pass = 0;
                double *buf = g_o3d_faithful ? (double *)malloc(sizeof(double) * 3 * (size_t)n) : NULL;
#pragma omp for schedule(dynamic, 256) nowait
                for (int64_t id = lo; id < hi; ++id) {
                    int32_t s[4];
                    double Th[12];
                    lro_sample(seed, (uint64_t)id, sampler, m, n, growth, s);
                    if (!lro_model_from_sample(src, tgt, s, m, use_elc, elc_ratio, Th)) continue;
                    ++t_pass;
                    int64_t c = buf ? lro_count_inliers_faithful(src, n, tgt, n, Th, thr, buf, NULL)
                                    : lro_count_inliers(src, tgt, n, Th, thr, NULL);
                    if (c > t_cnt || (c == t_cnt && id < t_id)) { t_cnt = c; t_id = id; }
                }
                free(buf);
#pragma omp critical
                {
                    r_pass += t_pass;
                    if (t_id >= 0 && (t_cnt > r_cnt || (t_cnt == r_cnt && t_id < r_id))) { r_cnt = t_cnt; r_id = t_id; }
                }
            }
            passed += r_pass;
            if (r_id >= 0 && r_cnt > best_cnt) { best_cnt = r_cnt; best_id = r_id; }
            done = hi;
            if (best_id >= 0 && done >= lro_conf_iters(best_cnt, n, m, conf, max_iters)) break;
        }
    }
    if (best_id >= 0 && best_cnt > 0) { /* a 0-inlier model never replaces the identity (App. B: fitness must improve on 0) */
        int32_t s[4];
        lro_sample(seed, (uint64_t)best_id, sampler, m, n, growth, s);
        lro_model_from_sample(src, tgt, s, m, 0, elc_ratio, bestT);
    }
    memcpy(T, bestT, sizeof(bestT));
    int64_t refit_cnt = 0;
    if (Trefit || mask) {
        uint8_t *mk = mask ? mask : (uint8_t *)malloc((size_t)(n > 0 ? n : 1));
        refit_cnt = lro_count_inliers(src, tgt, n, bestT, thr, mk);
        if (Trefit) {
            double *P = (double *)malloc(sizeof(double) * 3 * (size_t)(refit_cnt > 0 ? refit_cnt : 1));
            double *Q = (double *)malloc(sizeof(double) * 3 * (size_t)(refit_cnt > 0 ? refit_cnt : 1));
            int64_t k = 0;
            for (int64_t i = 0; i < n; ++i)
                if (mk[i]) {
                    for (int c = 0; c < 3; ++c) { P[3 * k + c] = src[3 * i + c]; Q[3 * k + c] = tgt[3 * i + c]; }
                    ++k;
                }
            lro_kabsch(P, Q, k, Trefit);
            free(P); free(Q);
        }
        if (!mask) free(mk);
    }
    free(growth);
    if (st) {
        st->iters_run = done; st->n_passed = passed; st->best_id = best_id;
        st->best_count = best_cnt; st->refit_count = refit_cnt;
    }
}

/* ------------------------------------------------------------------------- */
/* GC-RANSAC semantics: MSAC score, local         */
/* optimisation, iterated least squares                                      */
/* ------------------------------------------------------------------------- */

/* The MSAC value sum_{r^2 < tau^2} (1 - r^2/tau^2), tau = 1.5*thr (App. A), is a floating-point sum whose
 * value depends on the order of the terms.  To make "best model = highest score" a well-defined, order-free
 * (hence thread- and GPU-count independent, bit-reproducible) criterion, every term is quantised to
 * 2^-16 and the score is the exact INTEGER sum  q = sum trunc((1 - r^2/tau^2) * 65536).  n < 2^31 terms of
 * at most 2^16 fit an int64 with room to spare. */
#define LRO_MSAC_SCALE 65536.0

static inline int64_t lro_msac_term(double r2, double tau2)
{
    return (int64_t)((1.0 - r2 / tau2) * LRO_MSAC_SCALE);
}

LRO_API int64_t lro_msac_q(const float *src, const float *tgt, int64_t n, const double T[12], double thr,
                           int64_t *inliers_out)
{
    const double tau2 = (1.5 * thr) * (1.5 * thr);
    int64_t q = 0, cnt = 0;
    for (int64_t i = 0; i < n; ++i) {
        double r2 = lro_res2(T, src + 3 * i, tgt + 3 * i);
        if (r2 < tau2) { q += lro_msac_term(r2, tau2); ++cnt; }
    }
    if (inliers_out) *inliers_out = cnt;
    return q;
}

/* Fed-sample hook, MSAC flavour: scores[h] = q of sample h's Kabsch model (-1 where ELC rejects it),
 * inliers[h] = #(r^2 < tau^2).  Returns the selected hypothesis: max q, ties -> lowest h; -1 if none scored above 0. */
LRO_API int64_t lro_score_samples_msac(const float *src, const float *tgt, int64_t n, const int32_t *samples,
                                       int64_t H, int m, double thr, int use_elc, double elc_ratio,
                                       int64_t *scores, int32_t *inliers)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t h = 0; h < H; ++h) {
        double T[12];
        int64_t inl = -1;
        int ok = lro_model_from_sample(src, tgt, samples + h * m, m, use_elc, elc_ratio, T);
        scores[h] = ok ? lro_msac_q(src, tgt, n, T, thr, &inl) : -1;
        if (inliers) inliers[h] = (int32_t)inl;
    }
    int64_t best = -1, bq = 0; /* a model that scores 0 is never selected */
    for (int64_t h = 0; h < H; ++h)
        if (scores[h] > bq) { bq = scores[h]; best = h; }
    return best;
}

/* k <= 32 unique indices out of [0, n): same rule as lro_unique (draw d picks the r-th index not yet taken) */
static void lro_unique_k(uint64_t seed, uint64_t id, int k, int64_t n, int32_t *out)
{
    int32_t taken[32];
    for (int d = 0; d < k; ++d) {
        int32_t r = (int32_t)lro_draw(seed, id, (uint32_t)d, (uint32_t)(n - d));
        for (int e = 0; e < d; ++e)
            if (r >= taken[e]) ++r;
        out[d] = r;
        int e = d;
        while (e > 0 && taken[e - 1] > r) { taken[e] = taken[e - 1]; --e; }
        taken[e] = r;
    }
}

#define LRO_LO_SEED_SALT 0x4C4F43414C4F5054ULL /* "LOCALOPT" */
#define LRO_LO_SAMPLE_FACTOR 7                /* inner samples of min(7 m, #inliers) points (App. A) */

typedef struct {
    int64_t iters_run;    /* hypotheses generated */
    int64_t n_passed;     /* hypotheses that passed ELC (= scored) */
    int64_t best_id;      /* minimal-sample hypothesis with the highest q (-1: none) */
    int64_t best_score;   /* its q */
    int64_t best_inliers; /* its #(r^2 < tau^2), the I of the confidence rule */
    int64_t lo_score;     /* q after the local optimisation */
    int64_t final_score;  /* q after the iterated least squares (= q of T) */
    int64_t refit_count;  /* inliers (at thr) of T used by Trefit */
    int32_t lo_improved;  /* LO rounds that improved q */
    int32_t lsq_improved; /* least-squares iterations that improved q */
} LroGcStats;

/* GC-RANSAC loop (App. A) restated for a batched evaluation order:
 *  1. global search: hypotheses id = 0,1,... (pure functions of (seed, id)) -> ELC -> Kabsch -> q; selected =
 *     highest q, ties -> lowest id, q = 0 never replaces the identity; the confidence exit
 *     log(1-conf)/log(1-(I/n)^m) uses the selected model's I = #(r^2 < tau^2) and is evaluated at round ends;
 *  2. local optimisation (lo_rounds > 0; graph cut with spatial_coherence_weight = 0 is thresholding at thr,
 *     App. A): up to lo_rounds rounds of { L = inliers of the current model at thr, stop if |L| <= m;
 *     lo_trials inner draws of min(7 m, |L|) unique members of L -> non-minimal Kabsch -> q; take the best
 *     trial (ties -> lowest trial) if its q is higher, else stop };
 *  3. iterated least squares (lsq_iters > 0): refit on the inliers at thr, keep while q improves.
 * Upstream runs 2. whenever the so-far-best improves inside its sequential loop; the batched order runs it
 * once on the winner ("if no LO ever ran, run one", App. A) -- the difference is confined to which
 * intermediate models get polished, not to the definition of any step.
 * T = final model; Trefit (nullable) = Kabsch over its inliers at thr; mask (nullable) = those inliers. */
LRO_API void lro_ransac_gc(const float *src, const float *tgt, int64_t n, int m, int sampler, int use_elc,
                           double elc_ratio, double thr, double conf, int64_t max_iters, int64_t round,
                           uint64_t seed, int lo_rounds, int lo_trials, int lsq_iters, double T[12],
                           double *Trefit, uint8_t *mask, LroGcStats *st)
{
    int64_t best_id = -1, best_q = 0, best_inl = 0, passed = 0, done = 0;
    double cur[12];
    uint32_t *growth = NULL;
    if (sampler == 1 && n >= m) {
        growth = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n);
        lro_prosac_growth(n, m, growth);
    }
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) cur[4 * r + c] = (r == c) ? 1.0 : 0.0;
    if (n >= m && round > 0) {
        while (done < max_iters) {
            int64_t lo = done, hi = done + round < max_iters ? done + round : max_iters;
            int64_t r_id = -1, r_q = 0, r_inl = 0, r_pass = 0;
#pragma omp parallel
            {
                int64_t t_id = -1, t_q = 0, t_inl = 0, t_pass = 0;
#pragma omp for schedule(dynamic, 256) nowait
                for (int64_t id = lo; id < hi; ++id) {
                    int32_t s[4];
                    double Th[12];
                    int64_t inl = 0;
                    lro_sample(seed, (uint64_t)id, sampler, m, n, growth, s);
                    if (!lro_model_from_sample(src, tgt, s, m, use_elc, elc_ratio, Th)) continue;
                    ++t_pass;
                    int64_t q = lro_msac_q(src, tgt, n, Th, thr, &inl);
                    if (q > t_q || (q == t_q && q > 0 && id < t_id)) { t_q = q; t_id = id; t_inl = inl; }
                }
#pragma omp critical
                {
                    r_pass += t_pass;
                    if (t_id >= 0 && (t_q > r_q || (t_q == r_q && t_id < r_id))) { r_q = t_q; r_id = t_id; r_inl = t_inl; }
                }
            }
            passed += r_pass;
            if (r_id >= 0 && r_q > best_q) { best_q = r_q; best_id = r_id; best_inl = r_inl; }
            done = hi;
            if (best_id >= 0 && done >= lro_conf_iters(best_inl, n, m, conf, max_iters)) break;
        }
    }
    int64_t cur_q = 0;
    if (best_id >= 0) {
        int32_t s[4];
        lro_sample(seed, (uint64_t)best_id, sampler, m, n, growth, s);
        lro_model_from_sample(src, tgt, s, m, 0, elc_ratio, cur);
        cur_q = best_q;
    }
    const double thr2 = thr * thr;
    int32_t lo_improved = 0, lsq_improved = 0;
    int32_t *L = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    const uint64_t lo_seed = lro_mix64(seed ^ LRO_LO_SEED_SALT);
    if (lo_trials > 64) lo_trials = 64; /* the CUDA side scores one round's trials in one launch */
    for (int rd = 0; best_id >= 0 && rd < lo_rounds; ++rd) {
        int64_t I = 0;
        for (int64_t i = 0; i < n; ++i)
            if (lro_res2(cur, src + 3 * i, tgt + 3 * i) < thr2) L[I++] = (int32_t)i;
        if (I <= m) break;
        const int s_n = (int)(I < LRO_LO_SAMPLE_FACTOR * m ? I : LRO_LO_SAMPLE_FACTOR * m);
        int t_best = -1;
        int64_t q_best = 0;
        double T_best[12];
        for (int t = 0; t < lo_trials; ++t) {
            int32_t pos[32];
            double P[96], Q[96], Tt[12];
            lro_unique_k(lo_seed, (uint64_t)rd * (uint64_t)lo_trials + (uint64_t)t, s_n, I, pos);
            for (int d = 0; d < s_n; ++d)
                for (int c = 0; c < 3; ++c) {
                    P[3 * d + c] = (double)src[3 * (int64_t)L[pos[d]] + c];
                    Q[3 * d + c] = (double)tgt[3 * (int64_t)L[pos[d]] + c];
                }
            lro_kabsch(P, Q, s_n, Tt);
            int64_t q = lro_msac_q(src, tgt, n, Tt, thr, NULL);
            if (q > q_best) { q_best = q; t_best = t; memcpy(T_best, Tt, sizeof(Tt)); }
        }
        if (t_best >= 0 && q_best > cur_q) { cur_q = q_best; memcpy(cur, T_best, sizeof(cur)); ++lo_improved; }
        else break;
    }
    const int64_t lo_q = cur_q;
    for (int it = 0; best_id >= 0 && it < lsq_iters; ++it) {
        double Tn[12];
        int64_t k = 0;
        {
            double *P = (double *)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1));
            double *Q = (double *)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1));
            for (int64_t i = 0; i < n; ++i)
                if (lro_res2(cur, src + 3 * i, tgt + 3 * i) < thr2) {
                    for (int c = 0; c < 3; ++c) { P[3 * k + c] = src[3 * i + c]; Q[3 * k + c] = tgt[3 * i + c]; }
                    ++k;
                }
            if (k >= m) lro_kabsch(P, Q, k, Tn);
            free(P); free(Q);
        }
        if (k < m) break;
        int64_t q = lro_msac_q(src, tgt, n, Tn, thr, NULL);
        if (q > cur_q) { cur_q = q; memcpy(cur, Tn, sizeof(cur)); ++lsq_improved; }
        else break;
    }
    free(L);
    memcpy(T, cur, sizeof(cur));
    int64_t refit_cnt = 0;
    if (Trefit || mask) {
        uint8_t *mk = mask ? mask : (uint8_t *)malloc((size_t)(n > 0 ? n : 1));
        refit_cnt = lro_count_inliers(src, tgt, n, cur, thr, mk);
        if (Trefit) {
            double *P = (double *)malloc(sizeof(double) * 3 * (size_t)(refit_cnt > 0 ? refit_cnt : 1));
            double *Q = (double *)malloc(sizeof(double) * 3 * (size_t)(refit_cnt > 0 ? refit_cnt : 1));
            int64_t k = 0;
            for (int64_t i = 0; i < n; ++i)
                if (mk[i]) {
                    for (int c = 0; c < 3; ++c) { P[3 * k + c] = src[3 * i + c]; Q[3 * k + c] = tgt[3 * i + c]; }
                    ++k;
                }
            lro_kabsch(P, Q, k, Trefit);
            free(P); free(Q);
        }
        if (!mask) free(mk);
    }
    free(growth);
    if (st) {
        st->iters_run = done; st->n_passed = passed; st->best_id = best_id; st->best_score = best_q;
        st->best_inliers = best_inl; st->lo_score = lo_q; st->final_score = cur_q; st->refit_count = refit_cnt;
        st->lo_improved = lo_improved; st->lsq_improved = lsq_improved;
    }
}

/* Refit over an arbitrary correspondence set given as index pairs into two
 * clouds (FR.py:99-111: inliers of the ORIGINAL NN set under T, then Kabsch). */
LRO_API int64_t lro_refit_indexed(const float *xyz0, const float *xyz1, const int64_t *i0, const int64_t *i1,
                                  int64_t K, const double T[12], double thr, double Tout[12])
{
    const double thr2 = thr * thr;
    double *P = (double *)malloc(sizeof(double) * 3 * (size_t)(K > 0 ? K : 1));
    double *Q = (double *)malloc(sizeof(double) * 3 * (size_t)(K > 0 ? K : 1));
    int64_t k = 0;
    for (int64_t e = 0; e < K; ++e) {
        const float *p = xyz0 + 3 * i0[e], *q = xyz1 + 3 * i1[e];
        if (lro_res2(T, p, q) < thr2) {
            for (int c = 0; c < 3; ++c) { P[3 * k + c] = p[c]; Q[3 * k + c] = q[c]; }
            ++k;
        }
    }
    lro_kabsch(P, Q, k, Tout);
    free(P); free(Q);
    return k;
}

/* ------------------------------------------------------------------------- */
/* the consumers right after the path                          */
/* ------------------------------------------------------------------------- */

/* Weighted Kabsch of one neighbourhood -- Experiments/models/common.py:7-45 (rigid_transform_3d, the call of
 * Experiments/models/PointDSC.py:318 with total_weight): centroids = sum w x / (sum w + 1e-6) (:24-25),
 * H = Am^T diag(w) Bm (:32-33), R = V diag(1,1,det(V U^T)) U^T (:36-41), t = cB - R cA (:42).  Sums in index
 * order, fp64 on the fp32 inputs.  A, B: k x 3; w: k (nullable = ones). */
LRO_API void lro_kabsch_weighted(const float *A, const float *B, const float *w, int64_t k, double T[12])
{
    double sw = 0.0, ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0};
    for (int64_t i = 0; i < k; ++i) {
        const double wi = w ? (double)w[i] : 1.0;
        sw = sw + wi;
        for (int c = 0; c < 3; ++c) {
            ca[c] = ca[c] + wi * (double)A[3 * i + c];
            cb[c] = cb[c] + wi * (double)B[3 * i + c];
        }
    }
    const double den = sw + 1e-6;
    for (int c = 0; c < 3; ++c) { ca[c] = ca[c] / den; cb[c] = cb[c] / den; }
    double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; /* sum w (b - cb)(a - ca)^T: the transpose of the reference's H */
    for (int64_t i = 0; i < k; ++i) {
        const double wi = w ? (double)w[i] : 1.0;
        double da[3], db[3];
        for (int c = 0; c < 3; ++c) { da[c] = (double)A[3 * i + c] - ca[c]; db[c] = (double)B[3 * i + c] - cb[c]; }
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) H[3 * r + c] = H[3 * r + c] + (wi * db[r]) * da[c];
    }
    double R[9];
    lro_rot_from_H(H, R);
    for (int r = 0; r < 3; ++r) {
        T[4 * r + 0] = R[3 * r + 0];
        T[4 * r + 1] = R[3 * r + 1];
        T[4 * r + 2] = R[3 * r + 2];
        T[4 * r + 3] = cb[r] - ((R[3 * r + 0] * ca[0] + R[3 * r + 1] * ca[1]) + R[3 * r + 2] * ca[2]);
    }
}

/* Seed scoring -- Experiments/models/PointDSC.py:319-336: every seed's transform is applied to all n
 * correspondences, fitness = #(|T p - q| < thr) / n (:323-324), the best seed is the arg-max (first maximum, :326),
 * final labels = its inlier mask (:330-331).  models: S x 12 ([R|t] rows).  Returns the best seed. */
LRO_API int64_t lro_seeds_score(const float *src, const float *tgt, int64_t n, const double *models, int64_t S,
                                double thr, int32_t *counts, uint8_t *labels)
{
    int64_t best = 0, best_c = -1;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < S; ++s) counts[s] = (int32_t)lro_count_inliers(src, tgt, n, models + 12 * s, thr, NULL);
    for (int64_t s = 0; s < S; ++s)
        if ((int64_t)counts[s] > best_c) { best_c = counts[s]; best = s; }
    if (labels && S > 0) lro_count_inliers(src, tgt, n, models + 12 * best, thr, labels);
    return S > 0 ? best : -1;
}

/* Nearest target point of every transformed source point within a radius -- what Open3D's registration_icp
 * (Experiments/test.py:183-188) asks of its KD-tree (SearchHybrid(point, max_correspondence_distance, 1)): brute
 * force over all m targets, canonical squared distance of lro_res2, strict d^2 < radius^2, ties -> lowest index.
 * idx[i] = -1 when no target is inside the radius. */
LRO_API void lro_nn3d_radius(const float *src, int64_t n, const float *tgt, int64_t m, const double T[12],
                             double radius, int64_t *idx, double *d2)
{
    const double r2 = radius * radius;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double best = r2;
        int64_t bj = -1;
        for (int64_t j = 0; j < m; ++j) {
            const double d = lro_res2(T, src + 3 * i, tgt + 3 * j);
            if (d < best) { best = d; bj = j; }
        }
        idx[i] = bj;
        if (d2) d2[i] = bj >= 0 ? best : 0.0;
    }
}

LRO_API int lro_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

LRO_API void lro_set_threads(int t)
{
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}
