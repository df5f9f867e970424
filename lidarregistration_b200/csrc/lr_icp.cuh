// lr_icp.cuh -- the two consumers right after the path, device side.  Included by lr_ransac.cu inside its
// anonymous namespace (uses its canonical fp64 helpers: res2_f64, rot_from_H, finish_T, make_key).
//
//   * ICP refinement -- o3d.pipelines.registration.registration_icp(src, tgt, 0.6, T_init, PointToPoint)
//     (Experiments/test.py:183-188): per iteration the nearest target point of every transformed source point inside
//     the correspondence distance (Open3D: KD-tree SearchHybrid(point, max_dist, 1)), Kabsch over those pairs,
//     relative fitness / rmse stopping rule.  Here the target is binned once into a hashed uniform grid of
//     max_dist-sized cells (a point inside the radius lies in one of the 27 cells around the query), and ONE kernel per
//     iteration does transform + search + sums + solve + stopping rule; all iterations are enqueued up front and
//     switch themselves off through IcpCtl::done -- no host round trip inside the refinement.
//   * PointDSC seed scoring (Experiments/models/PointDSC.py:319-336): the per-seed weighted Kabsch of
//     rigid_transform_3d (models/common.py:7-45) and the (seed transform) x (correspondence) inlier sweep, which is the
//     RANSAC path's own tensor-core sweep fed with the caller's models.
#pragma once

// ---------------------------------------------------------------- hashed uniform grid over the target cloud
constexpr unsigned long long kGridEmpty = ~0ULL;
constexpr int kGridBias = 1 << 20;  // cell coordinates are clamped to (-2^20, 2^20): monotone, so neighbours stay neighbours

struct Grid {
    unsigned long long *keys;  // [size] packed cell coordinates, kGridEmpty = free
    unsigned int *cnt;         // [size] points in the cell
    unsigned int *start;       // [size] first point of the cell in pts[]
    float4 *pts;               // [m] (x, y, z, original index as bits), grouped by cell
    unsigned int *pt_slot, *pt_rank;  // [m] build scratch
    unsigned int mask;         // size - 1 (size is a power of two >= 2 m)
    double inv_cell;           // 1 / (max_dist * (1 + 1e-7)): |x - y| < max_dist  =>  cell indices differ by at most 1
};

__device__ __forceinline__ int grid_coord(double x, double inv_cell)
{
    double c = floor(x * inv_cell);
    c = fmin(fmax(c, -(double)(kGridBias - 2)), (double)(kGridBias - 2));
    return (int)c;
}
__device__ __forceinline__ unsigned long long grid_key(int ix, int iy, int iz)
{
    return ((unsigned long long)(unsigned)(ix + kGridBias) << 42) | ((unsigned long long)(unsigned)(iy + kGridBias) << 21) |
           (unsigned long long)(unsigned)(iz + kGridBias);
}

__global__ void __launch_bounds__(256)
k_grid_insert(const float *__restrict__ tgt, int64_t m, Grid g)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    const unsigned long long key = grid_key(grid_coord((double)tgt[3 * i], g.inv_cell), grid_coord((double)tgt[3 * i + 1], g.inv_cell),
                                            grid_coord((double)tgt[3 * i + 2], g.inv_cell));
    unsigned int h = (unsigned int)mix64(key) & g.mask;
    for (;;) {
        const unsigned long long prev = atomicCAS(&g.keys[h], kGridEmpty, key);
        if (prev == kGridEmpty || prev == key) break;
        h = (h + 1u) & g.mask;
    }
    g.pt_slot[i] = h;
    g.pt_rank[i] = atomicAdd(&g.cnt[h], 1u);  // order inside a cell is arbitrary: the search breaks ties by index
}

// exclusive scan of cnt[] -> start[] (one block; the table has at most a few hundred thousand slots)
__global__ void __launch_bounds__(1024)
k_grid_scan(Grid g)
{
    __shared__ unsigned int s_sum[1024];
    const unsigned int size = g.mask + 1u, per = (size + 1023u) / 1024u;
    const unsigned int lo = threadIdx.x * per, hi = lo + per < size ? lo + per : size;
    unsigned int t = 0;
    for (unsigned int k = lo; k < hi; ++k) t += g.cnt[k];
    s_sum[threadIdx.x] = t;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        unsigned int v = threadIdx.x >= (unsigned)o ? s_sum[threadIdx.x - o] : 0u;
        __syncthreads();
        s_sum[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned int run = s_sum[threadIdx.x] - t;
    for (unsigned int k = lo; k < hi; ++k) {
        g.start[k] = run;
        run += g.cnt[k];
    }
}

__global__ void __launch_bounds__(256)
k_grid_scatter(const float *__restrict__ tgt, int64_t m, Grid g)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    g.pts[g.start[g.pt_slot[i]] + g.pt_rank[i]] = make_float4(tgt[3 * i], tgt[3 * i + 1], tgt[3 * i + 2], __uint_as_float((unsigned)i));
}

// nearest target of the transformed point mv (canonical squared distance: the tail of res2_f64) with d^2 < r2,
// ties -> lowest target index; returns -1 when the ball is empty
__device__ __forceinline__ long long grid_nearest(const Grid &g, const double (&mv)[3], double r2, double &best_d2)
{
    const int cx = grid_coord(mv[0], g.inv_cell), cy = grid_coord(mv[1], g.inv_cell), cz = grid_coord(mv[2], g.inv_cell);
    double best = r2;
    long long bj = -1;
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const unsigned long long key = grid_key(cx + dx, cy + dy, cz + dz);
                unsigned int h = (unsigned int)mix64(key) & g.mask;
                unsigned long long k;
                while ((k = __ldg(&g.keys[h])) != kGridEmpty && k != key) h = (h + 1u) & g.mask;
                if (k != key) continue;
                const unsigned int a = __ldg(&g.start[h]), b = a + __ldg(&g.cnt[h]);
                for (unsigned int e = a; e < b; ++e) {
                    const float4 q = __ldg(&g.pts[e]);
                    const double d0 = mv[0] - (double)q.x, d1 = mv[1] - (double)q.y, d2 = mv[2] - (double)q.z;
                    const double d = (d0 * d0 + d1 * d1) + d2 * d2;
                    const long long j = (long long)__float_as_uint(q.w);
                    if (d < best || (d == best && bj >= 0 && j < bj)) {
                        best = d;
                        bj = j;
                    }
                }
            }
    best_d2 = best;
    return bj;
}

// the same search by LANES consecutive lanes of a warp (the 27 cells dealt round-robin, then a lexicographic (d^2, index)
// minimum across the lanes): the chain of dependent key -> start / count -> point loads is what bounds the search, and this
// cuts it LANES-fold.  Every lane of the group returns the result; the minimum does not depend on the order it is taken in.
template <int LANES>
__device__ __forceinline__ long long grid_nearest_coop(const Grid &g, const double (&mv)[3], double r2, double &best_d2)
{
    const int sub = (threadIdx.x & 31) % LANES;
    const int cx = grid_coord(mv[0], g.inv_cell), cy = grid_coord(mv[1], g.inv_cell), cz = grid_coord(mv[2], g.inv_cell);
    double best = r2;
    long long bj = -1;
    for (int c = sub; c < 27; c += LANES) {
        const int dz = c / 9 - 1, dy = (c / 3) % 3 - 1, dx = c % 3 - 1;
        const unsigned long long key = grid_key(cx + dx, cy + dy, cz + dz);
        unsigned int h = (unsigned int)mix64(key) & g.mask;
        unsigned long long k;
        while ((k = __ldg(&g.keys[h])) != kGridEmpty && k != key) h = (h + 1u) & g.mask;
        if (k != key) continue;
        const unsigned int a = __ldg(&g.start[h]), b = a + __ldg(&g.cnt[h]);
        for (unsigned int e = a; e < b; ++e) {
            const float4 q = __ldg(&g.pts[e]);
            const double d0 = mv[0] - (double)q.x, d1 = mv[1] - (double)q.y, d2 = mv[2] - (double)q.z;
            const double d = (d0 * d0 + d1 * d1) + d2 * d2;
            const long long j = (long long)__float_as_uint(q.w);
            if (d < best || (d == best && bj >= 0 && j < bj)) {
                best = d;
                bj = j;
            }
        }
    }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, best, o);
        const long long oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (oj >= 0 && (od < best || (od == best && (bj < 0 || oj < bj)))) {
            best = od;
            bj = oj;
        }
    }
    best_d2 = best;
    return bj;
}

__device__ __forceinline__ void icp_move(const double *T, const float *__restrict__ p, double (&mv)[3])
{
    const double px = (double)p[0], py = (double)p[1], pz = (double)p[2];
    mv[0] = ((T[0] * px + T[1] * py) + T[2] * pz) + T[3];
    mv[1] = ((T[4] * px + T[5] * py) + T[6] * pz) + T[7];
    mv[2] = ((T[8] * px + T[9] * py) + T[10] * pz) + T[11];
}

// the search on its own (parity hook: lr_nn3d_radius)
__global__ void __launch_bounds__(256)
k_nn3d_query(const float *__restrict__ src, int64_t n, Grid g, const double *__restrict__ T12, double r2,
             int64_t *__restrict__ idx, double *__restrict__ d2)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double T[12], mv[3], bd;
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = T12[k];
    icp_move(T, src + 3 * i, mv);
    const long long j = grid_nearest(g, mv, r2, bd);
    idx[i] = j;
    if (d2) d2[i] = j >= 0 ? bd : 0.0;
}

// ---------------------------------------------------------------- one ICP iteration per launch
constexpr int kIcpLanes = 8;  // lanes that share the 27-cell search of one source point
struct IcpCtl {
    double T[12];      // transform the next launch evaluates
    double Tres[12];   // transform of the last evaluation (the result)
    double fitness, rmse;
    long long count;
    int it, done;
    unsigned int ticket, pad;
};

__global__ void k_icp_begin(IcpCtl *c, const double *T12)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int k = 0; k < 12; ++k) c->T[k] = c->Tres[k] = T12[k];
    c->fitness = c->rmse = 0.0;
    c->count = 0;
    c->it = 0;
    c->done = 0;
    c->ticket = 0u;
}

// evaluation e of the refinement (e = 0: the initial transform): correspondences of c->T, their count / squared error,
// Kabsch over them (fixed-order two-stage reduction as k_finish), Open3D's stopping rule in the last block
#ifndef LR_ICP_MINB
#define LR_ICP_MINB 3   // 80 registers, three blocks per SM: the search is a chain of dependent loads (0.756 -> 0.702 ms per 25k-point pair)
#endif
__global__ void __launch_bounds__(256, LR_ICP_MINB)
k_icp_eval(const float *__restrict__ src, int64_t n, const float *__restrict__ tgt, Grid g, double r2, IcpCtl *c,
           double *__restrict__ partial, int e, double rel_fitness, double rel_rmse)
{
    lr::pdl_wait();  // (launched with programmatic stream serialisation: iteration e + 1 is resident when e finishes)
    lr::pdl_launch();
    if (c->done) return;
    __shared__ int s_last;
    __shared__ double s_part[8][kFinVals];
    __shared__ double s_tot[kFinVals];
    double T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = c->T[k];
    double o[3], oq[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        o[k] = 1024.0 * rint((double)src[k] * (1.0 / 1024.0));
        oq[k] = 1024.0 * rint((double)tgt[k] * (1.0 / 1024.0));
    }
    double v[kFinVals];
#pragma unroll
    for (int k = 0; k < kFinVals; ++k) v[k] = 0.0;
    // kIcpLanes lanes per source point; the group's first lane carries the point's terms of the sums.  (The trip count is
    // warp-uniform: the shuffles inside the search need every lane.)
    const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, gstride = (int64_t)gridDim.x * blockDim.x;
    const bool lead = (threadIdx.x % kIcpLanes) == 0;
    const int64_t n_ceil = (n + (32 / kIcpLanes) - 1) / (32 / kIcpLanes) * (32 / kIcpLanes);
    for (int64_t i = gtid / kIcpLanes; i < n_ceil; i += gstride / kIcpLanes) {
        double mv[3] = {0, 0, 0}, bd = 0.0;
        long long j = -1;
        if (i < n) icp_move(T, src + 3 * i, mv);
        j = grid_nearest_coop<kIcpLanes>(g, mv, i < n ? r2 : -1.0, bd);
        if (j < 0 || !lead || i >= n) continue;
        double p[3], q[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            p[k] = (double)src[3 * i + k] - o[k];
            q[k] = (double)tgt[3 * j + k] - oq[k];
        }
        v[0] += 1.0;
        v[1] += bd;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            v[2 + k] += p[k];
            v[5 + k] += q[k];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) v[8 + 3 * r + k] += q[r] * p[k];
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kFinVals; ++k) {
        double x = v[k];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
        if (lane == 0) s_part[w][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < kFinVals) {
        double r = 0.0;
        for (int ww = 0; ww < (int)(blockDim.x >> 5); ++ww) r += s_part[ww][threadIdx.x];
        partial[(size_t)blockIdx.x * kFinVals + threadIdx.x] = r;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&c->ticket, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int k = w; k < kFinVals; k += (int)(blockDim.x >> 5)) {
        double x = 0.0;
        for (int blk = lane; blk < (int)gridDim.x; blk += 32) x += __ldcg(&partial[(size_t)blk * kFinVals + k]);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
        if (lane == 0) s_tot[k] = x;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    double tot[kFinVals];
#pragma unroll
    for (int k = 0; k < kFinVals; ++k) tot[k] = s_tot[k];
    const long long cnt = (long long)(tot[0] + 0.5);
    double Tn[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    if (cnt > 0) {
        double H[3][3], R[3][3], cp[3], cq[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) H[r][k] = tot[8 + 3 * r + k] - tot[5 + r] * tot[2 + k] / (double)cnt;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            cp[k] = o[k] + tot[2 + k] / (double)cnt;
            cq[k] = oq[k] + tot[5 + k] / (double)cnt;
        }
        rot_from_H(H, R);
        finish_T(R, cp, cq, Tn);
    }
    const double f2 = n > 0 ? (double)cnt / (double)n : 0.0;
    const double r2m = cnt > 0 ? sqrt(tot[1] / (double)cnt) : 0.0;
    if (e > 0 && fabs(c->fitness - f2) < rel_fitness && fabs(c->rmse - r2m) < rel_rmse) c->done = 1;
    c->fitness = f2;
    c->rmse = r2m;
    c->count = cnt;
    c->it = e;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        c->Tres[k] = T[k];
        c->T[k] = Tn[k];
    }
    c->ticket = 0u;
}

// ---------------------------------------------------------------- PointDSC: per-seed weighted Kabsch, seed scoring
// thread = seed: sums in index order (== oracle lro_kabsch_weighted, bit for bit)
__global__ void __launch_bounds__(128)
k_kabsch_weighted(const float *__restrict__ A, const float *__restrict__ B, const float *__restrict__ wgt, int64_t S, int k,
                  double *__restrict__ T_out)
{
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= S) return;
    const float *a = A + (size_t)s * k * 3, *b = B + (size_t)s * k * 3, *w = wgt ? wgt + (size_t)s * k : nullptr;
    double sw = 0.0, ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0};
    for (int i = 0; i < k; ++i) {
        const double wi = w ? (double)w[i] : 1.0;
        sw = sw + wi;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            ca[c] = ca[c] + wi * (double)a[3 * i + c];
            cb[c] = cb[c] + wi * (double)b[3 * i + c];
        }
    }
    const double den = sw + 1e-6;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        ca[c] = ca[c] / den;
        cb[c] = cb[c] / den;
    }
    double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int i = 0; i < k; ++i) {
        const double wi = w ? (double)w[i] : 1.0;
        double da[3], db[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            da[c] = (double)a[3 * i + c] - ca[c];
            db[c] = (double)b[3 * i + c] - cb[c];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) H[r][c] = H[r][c] + (wi * db[r]) * da[c];
    }
    double R[3][3], T[12];
    rot_from_H(H, R);
    finish_T(R, ca, cb, T);
    double *out = T_out + (size_t)s * 16;
#pragma unroll
    for (int j = 0; j < 12; ++j) out[j] = T[j];
    out[12] = out[13] = out[14] = 0.0;
    out[15] = 1.0;
}

// the seeds' 4x4 transforms become the survivors of one round of the tensor sweep (cf. k_probe_install)
__global__ void k_seed_install(const double *__restrict__ models16, int S, const float4 *__restrict__ P8, double thr2,
                               Ctl *ctl, double *__restrict__ m64, uint4 *__restrict__ Aimg, float *__restrict__ band,
                               int *__restrict__ cnt, uint32_t *__restrict__ slot_id)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h == 0) {
        ctl->n_surv = S;
        ctl->n_events = 0u;
    }
    if (h >= S) return;
    double cen[3], cenq[3], T[12], tt[3];
    tcs::tc_centre(P8, cen, cenq);
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        T[k] = models16[(size_t)h * 16 + k];
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
    slot_id[h] = (uint32_t)h;
}

// arg-max of the seeds' counts (first maximum, as torch.argmax of PointDSC.py:326); unlike a RANSAC round a
// 0-inlier winner is still the selection (final_trans = seedwise_trans[argmax], :329)
__global__ void __launch_bounds__(256)
k_seed_end(Ctl *ctl, const int *__restrict__ cnt, const double *__restrict__ m64, int S, int32_t *__restrict__ counts_out)
{
    __shared__ unsigned long long s_key[8];
    unsigned long long key = 0ULL;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const int c = cnt[s];
        if (counts_out) counts_out[s] = c;
        const unsigned long long k = make_key(c, (uint32_t)s);
        key = k > key ? k : key;
    }
    key = warp_max_u64(key);
    if ((threadIdx.x & 31) == 0) s_key[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (int k = 1; k < 8; ++k) key = s_key[k] > key ? s_key[k] : key;
    ctl->best_key = key;
    ctl->iters_run = S;
    ctl->n_scored = S;
    ctl->n_surv = 0;
    if (key) {
        const uint32_t s = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFULL);
        for (int k = 0; k < 12; ++k) ctl->T[k] = m64[(size_t)s * 12 + k];
    }
}
