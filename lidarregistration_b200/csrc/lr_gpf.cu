// lr_gpf.cu -- Grid-Prioritised Filter (--mode GPF) on the device.
//
// Replaces Grid_Prioritized_Filter of the reference, Experiments/algorithms/matching.py:100-205:
//   normalise the ratio quality to [0,1] (:119-124), best buddies get -1 so they sort first (:126-134),
//   cell of a pair = floor(G (x - min) / (max - min + 1e-3)) on the source x and y (:136-146),
//   per-cell population (:147-152), water-filling height by bisection (:154-179),
//   per cell: everything if under quota, else the `quota` smallest normalised distances (:184-195).
// Every step follows the reference's arithmetic operation for operation: fp32 for the tensors torch holds in fp32
// (IEEE sub / div / mul / floor are correctly rounded on both sides), fp64 for the numpy water-filling.  The per-cell
// argsort becomes a rank-by-counting selection: pair i of cell c is kept iff fewer than quota_c pairs of the cell have
// a smaller (distance, index) -- O(sum n_c^2) comparisons, a few million at 25-50k pairs, no sort, no host round trip.
#include <math.h>

#include "lr_common.cuh"

namespace {

constexpr int kMaxCells = 64 * 64;

struct GpfCtl {
    unsigned int fmin_key, fmax_key, xmin_key, xmax_key, ymin_key, ymax_key;  // order-preserving keys of fp32 values
    unsigned int pad[2];
    double total_num;
};

// monotone map fp32 -> uint32 (and back)
__device__ __forceinline__ unsigned int f2key(float f)
{
    const unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned int k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

__global__ void k_gpf_reset(GpfCtl *ctl, int *count, int *cursor, int ncell, double total_num)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) {
        ctl->fmin_key = ctl->xmin_key = ctl->ymin_key = 0xFFFFFFFFu;
        ctl->fmax_key = ctl->xmax_key = ctl->ymax_key = 0u;
        ctl->total_num = total_num;
    }
    if (t < ncell) {
        count[t] = 0;
        cursor[t] = 0;
    }
}

// min / max of the ratio and of the source x, y over the pairs (torch.min / torch.max of :120-121, :139-140)
__global__ void __launch_bounds__(256)
k_gpf_minmax(const float *__restrict__ ratio, const float *__restrict__ xyz0, const int64_t *__restrict__ idx0, int64_t n,
             GpfCtl *ctl)
{
    unsigned int lo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, hi[3] = {0u, 0u, 0u};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = idx0 ? idx0[i] : i;
        const unsigned int k[3] = {f2key(ratio[i]), f2key(xyz0[3 * s]), f2key(xyz0[3 * s + 1])};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            lo[c] = min(lo[c], k[c]);
            hi[c] = max(hi[c], k[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = min(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = max(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&ctl->fmin_key, lo[0]);
        atomicMax(&ctl->fmax_key, hi[0]);
        atomicMin(&ctl->xmin_key, lo[1]);
        atomicMax(&ctl->xmax_key, hi[1]);
        atomicMin(&ctl->ymin_key, lo[2]);
        atomicMax(&ctl->ymax_key, hi[2]);
    }
}

// normalised distance (best buddies - 1), cell id, per-cell population
__global__ void __launch_bounds__(256)
k_gpf_cells(const float *__restrict__ ratio, const uint8_t *__restrict__ is_bb, const float *__restrict__ xyz0,
            const int64_t *__restrict__ idx0, int64_t n, int G, const GpfCtl *__restrict__ ctl, float *__restrict__ norm,
            int *__restrict__ cell, int *__restrict__ count)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float fm = key2f(ctl->fmin_key), fM = key2f(ctl->fmax_key);
    const float xm = key2f(ctl->xmin_key), xM = key2f(ctl->xmax_key);
    const float ym = key2f(ctl->ymin_key), yM = key2f(ctl->ymax_key);
    // (T - m) / (M - m), then -1 for best buddies (:119-134); every operation individually rounded like torch's
    float v = __fdiv_rn(__fsub_rn(ratio[i], fm), __fsub_rn(fM, fm));
    if (is_bb && is_bb[i]) v = __fsub_rn(v, 1.0f);
    norm[i] = v;
    const int64_t s = idx0 ? idx0[i] : i;
    const float eps = 1e-3f;  // EPS = 10**-3 promoted to the tensor dtype (:137)
    const float gx = floorf(__fmul_rn((float)G, __fdiv_rn(__fsub_rn(xyz0[3 * s], xm), __fadd_rn(__fsub_rn(xM, xm), eps))));
    const float gy = floorf(__fmul_rn((float)G, __fdiv_rn(__fsub_rn(xyz0[3 * s + 1], ym), __fadd_rn(__fsub_rn(yM, ym), eps))));
    int ci = (int)gx, cj = (int)gy;
    ci = ci < 0 ? 0 : (ci >= G ? G - 1 : ci);  // (cannot leave [0, G) by construction; NaN input guards)
    cj = cj < 0 ? 0 : (cj >= G ? G - 1 : cj);
    const int c = ci * G + cj;
    cell[i] = c;
    atomicAdd(&count[c], 1);
}

// water-filling (:154-179) in fp64 like the reference's numpy, one thread; also the cells' offsets in the bucket list
__global__ void k_gpf_quota(const int *__restrict__ count, int ncell, const GpfCtl *__restrict__ ctl, int *__restrict__ quota,
                            int *__restrict__ start)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double total = ctl->total_num;
    auto filled = [&](double h) {
        double s = 0.0;
        for (int c = 0; c < ncell; ++c) {
            const double m = (double)count[c];
            s += (m < h) ? m : h;
        }
        return s;
    };
    double max_h = total, min_h = 0.0, cur = (max_h + min_h) / 2;
    while (fabs(max_h - min_h) > 2) {
        const double t = filled(cur);
        if (t == total) break;
        else if (t < total) min_h = cur;
        else max_h = cur;
        cur = (max_h + min_h) / 2;
    }
    const double h = rint(cur);  // np.round: half to even, like rint in the default rounding mode
    int off = 0;
    for (int c = 0; c < ncell; ++c) {
        const double m = (double)count[c];
        const double per = (m < h) ? m : h;
        quota[c] = (int)per;  // int(per_quad[qi, qj]) (:186)
        start[c] = off;
        off += count[c];
    }
    start[ncell] = off;
}

// bucket the pair ids by cell (order inside a bucket is irrelevant to the rank count)
__global__ void __launch_bounds__(256)
k_gpf_bucket(const int *__restrict__ cell, int64_t n, const int *__restrict__ start, int *__restrict__ cursor,
             int *__restrict__ bucket)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cell[i];
    bucket[start[c] + atomicAdd(&cursor[c], 1)] = (int)i;
}

// keep[i] = pair i is among the quota_c smallest (distance, index) of its cell; block = cell (blockIdx.y = slice of the
// cell's members that this block ranks)
__global__ void __launch_bounds__(256)
k_gpf_select(const float *__restrict__ norm, const int *__restrict__ bucket, const int *__restrict__ start,
             const int *__restrict__ quota, uint8_t *__restrict__ keep)
{
    const int c = blockIdx.x;
    const int lo = start[c], hi = start[c + 1], m = hi - lo, q = quota[c];
    if (m == 0) return;
    __shared__ float s_v[1024];
    __shared__ int s_i[1024];
    for (int base = blockIdx.y * blockDim.x; base < m; base += gridDim.y * blockDim.x) {
        const int t = base + threadIdx.x;
        const bool live = t < m;
        const int me = live ? bucket[lo + t] : -1;
        const float v = live ? norm[me] : 0.f;
        int rank = 0;
        const bool need = live && q > 0 && q < m;  // quota == population keeps everything (:188-189), 0 keeps nothing
        for (int tile = 0; tile < m; tile += 1024) {
            __syncthreads();
            for (int k = threadIdx.x; k < 1024 && tile + k < m; k += blockDim.x) {
                const int j = bucket[lo + tile + k];
                s_i[k] = j;
                s_v[k] = norm[j];
            }
            __syncthreads();
            if (need) {
                const int cnt = min(1024, m - tile);
                for (int k = 0; k < cnt; ++k) {
                    const float w = s_v[k];
                    rank += (w < v || (w == v && s_i[k] < me)) ? 1 : 0;
                }
            }
        }
        if (live) keep[me] = (q >= m) ? 1 : (q <= 0 ? 0 : (rank < q ? 1 : 0));
    }
}

}  // namespace

// [device] ratio[n] = d1 / (d2 + 1e-6) of the candidate pairs (lr_match_ratio), is_bb[n] (nullable: the TEASER variant,
// BB_first, has no offset), xyz0[.,3] source cloud, idx0[n] (nullable = identity) source index of each pair.
// total_num = GPF_factor x #best buddies (or GPF_max_matches).  Out: keep[n] (0 / 1), norm[n] the normalised distance the
// reference returns for the kept pairs (it becomes the PROSAC quality, FR.py:75-76).
LR_EXPORT int lr_gpf_filter(const float *ratio, const uint8_t *is_bb, const float *xyz0, const int64_t *idx0, int64_t n,
                            int grid_wid, double total_num, uint8_t *keep, float *norm, void *stream)
{
    lr::Lock lock;
    LR_REQUIRE(ratio && xyz0 && keep && norm, "null pointer");
    LR_REQUIRE(n > 0 && n < ((int64_t)1 << 31), "n out of range");
    LR_REQUIRE(grid_wid >= 1 && grid_wid * grid_wid <= kMaxCells, "GPF_grid_wid out of range (1..64)");
    cudaStream_t st = (cudaStream_t)stream;
    const int ncell = grid_wid * grid_wid;
    const size_t bytes = lr::padded(sizeof(GpfCtl)) + 4 * lr::padded(sizeof(int) * (ncell + 1)) + 2 * lr::padded(sizeof(int) * n);
    char *scratch = (char *)lr::arena_get(lr::SLOT_MISC, bytes);
    if (!scratch) return LR_ERR_ALLOC;
    lr::Carver cv(scratch);
    GpfCtl *ctl = cv.take<GpfCtl>(1);
    int *count = cv.take<int>(ncell + 1), *cursor = cv.take<int>(ncell + 1), *quota = cv.take<int>(ncell + 1),
        *start = cv.take<int>(ncell + 1);
    int *cell = cv.take<int>(n), *bucket = cv.take<int>(n);
    const unsigned nb = (unsigned)((n + 255) / 256);
    k_gpf_reset<<<(ncell + 255) / 256, 256, 0, st>>>(ctl, count, cursor, ncell, total_num);
    k_gpf_minmax<<<nb < 296u ? nb : 296u, 256, 0, st>>>(ratio, xyz0, idx0, n, ctl);
    k_gpf_cells<<<nb, 256, 0, st>>>(ratio, is_bb, xyz0, idx0, n, grid_wid, ctl, norm, cell, count);
    k_gpf_quota<<<1, 32, 0, st>>>(count, ncell, ctl, quota, start);
    k_gpf_bucket<<<nb, 256, 0, st>>>(cell, n, start, cursor, bucket);
    // a cell holds n / ncell pairs on average; a few blocks per cell keep the skewed ones from serialising
    const int slices = (int)((n / ncell + 1023) / 1024) + 1;
    k_gpf_select<<<dim3((unsigned)ncell, (unsigned)(slices < 8 ? slices : 8)), 256, 0, st>>>(norm, bucket, start, quota, keep);
    LR_CUDA_TRY(cudaGetLastError());
    return LR_OK;
}
