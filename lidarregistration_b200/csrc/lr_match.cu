// lr_match.cu -- feature-space correspondence search for sm_100a (exact fp32 path).
//
// Replaces (reference tree citations):
//   find_nn / knn_dist      Experiments/algorithms/matching.py:22-65
//   nn_to_mutual            Experiments/algorithms/matching.py:222-239
//   torch_intersect         Experiments/algorithms/matching.py:67-87
//   calc_distance_ratio_in_feature_space   Experiments/algorithms/matching.py:89-98
//
// The reference materialises 250 x M distance tiles (SGEMM + norms + clamp +
// sqrt + min, ~6 launches per 250 rows).  Here one kernel sweeps 128 x 128
// tiles with the distance, the argmin and the second argmin fused into the
// GEMM epilogue; nothing of size N x M ever reaches HBM.
//
// Canonical arithmetic (== oracle/lr_oracle.c, == torch CPU bit for bit):
//   dot   = sequential fp32 FMA over k = 0..D-1
//   |f|^2 = 8 strided lanes, lanes added in order
//   d2    = (|a|^2 + |b|^2) - 2 dot  == fma(-2, dot, |a|^2 + |b|^2)
//   dist  = sqrt(max(d2, 1e-30)), argmin with lowest index on ties
// The sqrt is only evaluated when a running minimum changes: the kernel
// compares d2 against the smallest d2 whose sqrt reaches the current value.
#include <math.h>

#include "lr_match_tc.cuh"

namespace {

int g_match_mode = 0;  // 0: auto; 1: always the exact CUDA-core sweep; 2 / 3: tensor-core sweep, fp32 / fp16 accumulators
constexpr bool kAutoAcc16 = true;   // what mode 0 picks for D == 32
inline bool tc_acc16() { return g_match_mode == 3 || (g_match_mode == 0 && kAutoAcc16); }


constexpr int BM = 128, BN = 128, NT = 256;  // block tile and threads (8 x 8 outputs per thread)

struct Cand {  // running best / second best of one row: lexicographic (dist, index)
    float s1;
    int j1;
    float s2;
    int j2;
};

__device__ __forceinline__ float f_prev(float x) { return __uint_as_float(__float_as_uint(x) - 1u); }
__device__ __forceinline__ float f_next(float x) { return __uint_as_float(__float_as_uint(x) + 1u); }

// smallest clamped d2 whose distance sqrt(d2) is >= s  (s = +inf -> +inf)
__device__ __noinline__ float lo_bound(float s)
{
    if (!(s < INFINITY)) return INFINITY;
    float x = __fmul_rn(s, s);
    for (int it = 0; it < 8; ++it) {
        float xp = f_prev(x);
        if (__fsqrt_rn(xp) >= s) x = xp;
        else break;
    }
    for (int it = 0; it < 8; ++it) {
        if (__fsqrt_rn(x) < s) x = f_next(x);
        else break;
    }
    return fmaxf(x, 1e-30f);
}

// rare path: candidate column j with raw d2 below the row's threshold
__device__ __noinline__ void cand_update(Cand &c, float &thr, float d2, int j, bool want2)
{
    const float s = __fsqrt_rn(fmaxf(d2, 1e-30f));
    if (s < c.s1) {
        c.s2 = c.s1;
        c.j2 = c.j1;
        c.s1 = s;
        c.j1 = j;
    } else if (s < c.s2) {
        c.s2 = s;
        c.j2 = j;
    } else {
        return;
    }
    thr = lo_bound(want2 ? c.s2 : c.s1);
}

__device__ __forceinline__ bool lex_less(float sa, int ja, float sb, int jb)
{
    return sa < sb || (sa == sb && ja < jb);
}

// merge candidate (s, j) into a top-2 list (arbitrary arrival order)
__device__ __forceinline__ void top2_insert(Cand &c, float s, int j)
{
    if (j < 0) return;
    if (lex_less(s, j, c.s1, c.j1)) {
        c.s2 = c.s1;
        c.j2 = c.j1;
        c.s1 = s;
        c.j1 = j;
    } else if (!(s == c.s1 && j == c.j1) && lex_less(s, j, c.s2, c.j2)) {
        c.s2 = s;
        c.j2 = j;
    }
}

__global__ void k_sqnorms(const float *__restrict__ F, int64_t N, int D, float *__restrict__ out)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float *x = F + i * D;
    float lane[8];
#pragma unroll
    for (int l = 0; l < 8; ++l) lane[l] = x[l] * x[l];
    for (int k = 8; k < D; k += 8) {
#pragma unroll
        for (int l = 0; l < 8; ++l) lane[l] = lane[l] + x[k + l] * x[k + l];
    }
    float s = lane[0];
#pragma unroll
    for (int l = 1; l < 8; ++l) s = s + lane[l];
    out[i] = s;
}

__device__ __forceinline__ void cp_async16_zfill(void *smem, const void *gmem, bool valid)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// rows [r0, r0+rows) of F[total, D] -> smem tile[rows][D + 4] (zero rows past `total`)
template <int D>
__device__ __forceinline__ void stage_tile(float *tile, const float *__restrict__ F, int64_t r0, int64_t total, int tid)
{
    constexpr int LD = D + 4;
    constexpr int V = D / 4;  // 16-byte pieces per row
    for (int p = tid; p < BM * V; p += NT) {
        const int r = p / V, v = p % V;
        const int64_t g = r0 + r;
        const bool ok = g < total;
        cp_async16_zfill(tile + r * LD + v * 4, F + (ok ? g : 0) * D + v * 4, ok);
    }
}

// One block: BM query rows against the column tiles [tile_lo, tile_hi) of F1.
// Thread (ty, tx) owns rows ty*8 + i and columns tx + 16*j of each tile.
template <int D, bool WANT2>
__global__ void __launch_bounds__(NT)
k_nn_exact(const float *__restrict__ F0, int64_t N, const float *__restrict__ F1, int64_t M,
           const float *__restrict__ n0, const float *__restrict__ n1, int tiles_per_split, int nsplit,
           Cand *__restrict__ part)
{
    extern __shared__ __align__(16) float smem[];
    constexpr int LD = D + 4;
    float *As = smem;                 // [BM][LD]
    float *Bs0 = As + BM * LD;        // [BN][LD] x 2
    float *Bs1 = Bs0 + BN * LD;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t row0 = (int64_t)blockIdx.x * BM;
    const int split = blockIdx.y;
    const int ntiles = (int)((M + BN - 1) / BN);
    const int t_lo = split * tiles_per_split;
    const int t_hi = min(ntiles, t_lo + tiles_per_split);

    Cand c[8];
    float thr[8], na[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        c[i].s1 = INFINITY;
        c[i].j1 = -1;
        c[i].s2 = INFINITY;
        c[i].j2 = -1;
        thr[i] = INFINITY;
        const int64_t r = row0 + ty * 8 + i;
        na[i] = r < N ? n0[r] : 0.f;
    }

    stage_tile<D>(As, F0, row0, N, tid);
    if (t_lo < t_hi) stage_tile<D>(Bs0, F1, (int64_t)t_lo * BN, M, tid);
    cp_async_commit();

    for (int t = t_lo; t < t_hi; ++t) {
        float *Bs = ((t - t_lo) & 1) ? Bs1 : Bs0;
        float *Bn = ((t - t_lo) & 1) ? Bs0 : Bs1;
        if (t + 1 < t_hi) {
            stage_tile<D>(Bn, F1, (int64_t)(t + 1) * BN, M, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

#pragma unroll
        for (int kq = 0; kq < D / 4; ++kq) {
            float4 a4[8], b4[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a4[i] = *reinterpret_cast<const float4 *>(As + (ty * 8 + i) * LD + kq * 4);
#pragma unroll
            for (int j = 0; j < 8; ++j) b4[j] = *reinterpret_cast<const float4 *>(Bs + (tx + 16 * j) * LD + kq * 4);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    acc[i][j] = fmaf(a4[i].x, b4[j].x, acc[i][j]);
                    acc[i][j] = fmaf(a4[i].y, b4[j].y, acc[i][j]);
                    acc[i][j] = fmaf(a4[i].z, b4[j].z, acc[i][j]);
                    acc[i][j] = fmaf(a4[i].w, b4[j].w, acc[i][j]);
                }
        }

        // fused epilogue: distances, running argmin / second argmin
        const int64_t col0 = (int64_t)t * BN;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t col = col0 + tx + 16 * j;
            const float nb = col < M ? n1[col] : INFINITY;  // columns past M can never win
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float d2 = fmaf(-2.f, acc[i][j], na[i] + nb);
                if (d2 < thr[i]) cand_update(c[i], thr[i], d2, (int)col, WANT2);
            }
        }
        __syncthreads();
    }

    // merge the 16 column-owners of each row, one thread per row
    Cand *cs = reinterpret_cast<Cand *>(smem);  // [BM][16], reuses the tiles (all reads are done)
#pragma unroll
    for (int i = 0; i < 8; ++i) cs[(ty * 8 + i) * 16 + tx] = c[i];
    __syncthreads();
    if (tid < BM) {
        const int64_t r = row0 + tid;
        if (r < N) {
            Cand m;
            m.s1 = INFINITY; m.j1 = 0x7fffffff; m.s2 = INFINITY; m.j2 = 0x7fffffff;
            for (int k = 0; k < 16; ++k) {
                const Cand e = cs[tid * 16 + k];
                top2_insert(m, e.s1, e.j1);
                top2_insert(m, e.s2, e.j2);
            }
            part[r * nsplit + split] = m;
        }
    }
}

// merge the column splits; indices widen to int64 like the reference's
__global__ void k_nn_merge(const Cand *__restrict__ part, int64_t N, int nsplit, int64_t *__restrict__ idx1,
                           int64_t *__restrict__ idx2)
{
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= N) return;
    Cand m;
    m.s1 = INFINITY; m.j1 = 0x7fffffff; m.s2 = INFINITY; m.j2 = 0x7fffffff;
    for (int s = 0; s < nsplit; ++s) {
        const Cand e = part[r * nsplit + s];
        top2_insert(m, e.s1, e.j1);
        top2_insert(m, e.s2, e.j2);
    }
    // torch.min over an all-inf row returns index 0 (M == 1 second-NN case)
    idx1[r] = (m.j1 == 0x7fffffff) ? 0 : m.j1;
    if (idx2) idx2[r] = (m.j2 == 0x7fffffff) ? 0 : m.j2;
}

// mutual check + order-preserving compaction in three small launches:
// per-block survivor counts, exclusive scan of the (few) block counts, ordered scatter
constexpr int kCompactBlock = 1024;

__device__ __forceinline__ bool is_mutual(const int64_t *__restrict__ idx1, const int64_t *__restrict__ rev, int64_t i,
                                          int64_t N, int64_t M, int64_t &j)
{
    if (i >= N) return false;
    j = idx1[i];
    return j >= 0 && j < M && rev[j] == i;
}

__global__ void __launch_bounds__(kCompactBlock)
k_mutual_count(const int64_t *__restrict__ idx1, const int64_t *__restrict__ rev, int64_t N, int64_t M,
               int *__restrict__ block_cnt)
{
    lr::pdl_wait();
    lr::pdl_launch();
    __shared__ int warp_tot[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int64_t j;
    const bool keep = is_mutual(idx1, rev, blockIdx.x * (int64_t)kCompactBlock + threadIdx.x, N, M, j);
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_tot[w] = __popc(b);
    __syncthreads();
    if (w == 0) {
        int v = warp_tot[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) block_cnt[blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(1024)
k_block_scan(int *__restrict__ block_cnt, int nblocks, int64_t *__restrict__ K)
{
    lr::pdl_wait();
    lr::pdl_launch();
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + tid;
        const int v = i < nblocks ? block_cnt[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[w] = incl;
        __syncthreads();
        int before = 0, total = 0;
        for (int k = 0; k < 32; ++k) {
            before += k < w ? warp_tot[k] : 0;
            total += warp_tot[k];
        }
        if (i < nblocks) block_cnt[i] = carry_s + before + incl - v;  // exclusive prefix
        __syncthreads();
        if (tid == 0) carry_s += total;
        __syncthreads();
    }
    if (tid == 0) *K = carry_s;
}

__global__ void __launch_bounds__(kCompactBlock)
k_mutual_scatter(const int64_t *__restrict__ idx1, const int64_t *__restrict__ rev, int64_t N, int64_t M,
                 const int *__restrict__ block_off, int64_t *__restrict__ out_i, int64_t *__restrict__ out_j)
{
    lr::pdl_wait();
    lr::pdl_launch();
    __shared__ int warp_tot[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t i = blockIdx.x * (int64_t)kCompactBlock + threadIdx.x;
    int64_t j = -1;
    const bool keep = is_mutual(idx1, rev, i, N, M, j);
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_tot[w] = __popc(b);
    __syncthreads();
    int before = 0;
    for (int k = 0; k < w; ++k) before += warp_tot[k];
    if (keep) {
        const int64_t off = (int64_t)block_off[blockIdx.x] + before + __popc(b & ((1u << lane) - 1u));
        out_i[off] = i;
        out_j[off] = j;
    }
}

__device__ __forceinline__ float diffnorm(const float *a, const float *b, int D)
{
    float lane[8];
#pragma unroll
    for (int l = 0; l < 8; ++l) {
        const float d = a[l] - b[l];
        lane[l] = d * d;
    }
    for (int k = 8; k < D; k += 8) {
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            const float d = a[k + l] - b[k + l];
            lane[l] = lane[l] + d * d;
        }
    }
    float s = lane[0];
#pragma unroll
    for (int l = 1; l < 8; ++l) s = s + lane[l];
    return __fsqrt_rn(s);
}

__global__ void k_ratio(const float *__restrict__ f0, const float *__restrict__ f1, int D, int64_t K,
                        const int64_t *__restrict__ i0, const int64_t *__restrict__ i1,
                        const int64_t *__restrict__ i2, float *__restrict__ out)
{
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= K) return;
    const float *a = f0 + i0[k] * D;
    const float da = diffnorm(a, f1 + i1[k] * D, D);
    const float db = diffnorm(a, f1 + i2[k] * D, D);
    out[k] = __fdiv_rn(da, db + 1e-6f);
}

__global__ void k_gather_xyz(const float *__restrict__ xyz, const int64_t *__restrict__ idx, int64_t K,
                             float *__restrict__ out)
{
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int64_t s = idx[k];
    out[3 * k + 0] = xyz[3 * s + 0];
    out[3 * k + 1] = xyz[3 * s + 1];
    out[3 * k + 2] = xyz[3 * s + 2];
}

template <int D>
int launch_nn_d(const float *f0, int64_t N, const float *f1, int64_t M, const float *n0, const float *n1,
                Cand *part, int tps, int nsplit, bool want2, cudaStream_t st)
{
    size_t smem = sizeof(float) * (size_t)(BM + 2 * BN) * (D + 4);
    if (smem < sizeof(Cand) * BM * 16) smem = sizeof(Cand) * BM * 16;  // the merge reuses the tiles
    dim3 grid((unsigned)((N + BM - 1) / BM), (unsigned)nsplit);
    if (want2) {
        LR_CUDA_TRY(cudaFuncSetAttribute(k_nn_exact<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_nn_exact<D, true><<<grid, NT, smem, st>>>(f0, N, f1, M, n0, n1, tps, nsplit, part);
    } else {
        LR_CUDA_TRY(cudaFuncSetAttribute(k_nn_exact<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_nn_exact<D, false><<<grid, NT, smem, st>>>(f0, N, f1, M, n0, n1, tps, nsplit, part);
    }
    LR_CUDA_TRY(cudaGetLastError());
    return LR_OK;
}

// NN (+ second NN) of every row of f0 in f1; scratch from the MATCH arena at `offset`
int nn_sweep(const float *f0, int64_t N, const float *f1, int64_t M, int D, int64_t *idx1, int64_t *idx2,
             char *scratch, cudaStream_t st)
{
    const int ntiles = (int)((M + BN - 1) / BN);
    const int rowblocks = (int)((N + BM - 1) / BM);
    // enough blocks for >= ~6 waves at 2 blocks/SM, but never split finer than 8 tiles
    int nsplit = (lr::sm_count() * 12 + rowblocks - 1) / rowblocks;
    int max_split = (ntiles + 7) / 8;
    if (nsplit > max_split) nsplit = max_split;
    if (nsplit < 1) nsplit = 1;
    const int tps = (ntiles + nsplit - 1) / nsplit;
    nsplit = (ntiles + tps - 1) / tps;
    lr::Carver cv(scratch);
    float *n0 = cv.take<float>(N);
    float *n1 = cv.take<float>(M);
    Cand *part = cv.take<Cand>((size_t)N * nsplit);
    k_sqnorms<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(f0, N, D, n0);
    k_sqnorms<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(f1, M, D, n1);
    int rc;
    const bool want2 = idx2 != nullptr;
    const int tok = lr::prof_begin(lr::PROF_NN, st);
    switch (D) {
        case 8: rc = launch_nn_d<8>(f0, N, f1, M, n0, n1, part, tps, nsplit, want2, st); break;
        case 16: rc = launch_nn_d<16>(f0, N, f1, M, n0, n1, part, tps, nsplit, want2, st); break;
        case 32: rc = launch_nn_d<32>(f0, N, f1, M, n0, n1, part, tps, nsplit, want2, st); break;
        case 64: rc = launch_nn_d<64>(f0, N, f1, M, n0, n1, part, tps, nsplit, want2, st); break;
        default: lr::set_error("unsupported feature dimension D=%d (8, 16, 32 or 64)", D); return LR_ERR_ARG;
    }
    if (rc) return rc;
    lr::prof_end(tok, st);
    k_nn_merge<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(part, N, nsplit, idx1, idx2);
    LR_CUDA_TRY(cudaGetLastError());
    return LR_OK;
}

size_t nn_scratch_bytes(int64_t N, int64_t M)
{
    const int ntiles = (int)((M + BN - 1) / BN);
    int max_split = (ntiles + 7) / 8;
    if (max_split < 1) max_split = 1;
    return lr::padded(sizeof(float) * N) + lr::padded(sizeof(float) * M) + lr::padded(sizeof(Cand) * (size_t)N * max_split);
}

}  // namespace

LR_EXPORT int lr_match_nn(const float *f0, int64_t N, const float *f1, int64_t M, int D, int64_t *idx1,
                          int64_t *idx1_2nd, void *stream)
{
    lr::Lock lock;
    LR_REQUIRE(f0 && f1 && idx1, "null pointer");
    LR_REQUIRE(N > 0 && M > 0 && N < ((int64_t)1 << 31) && M < ((int64_t)1 << 31), "N/M out of range");
    cudaStream_t st = (cudaStream_t)stream;
    if (D == 32 && g_match_mode != 1) {
        char *scratch = (char *)lr::arena_get(lr::SLOT_MATCH, lr_tc::scratch_bytes(N, M));
        if (!scratch) return LR_ERR_ALLOC;
        lr_tc::Prepared P;
        int rc = lr_tc::prepare(f0, N, f1, M, scratch, P, st);
        if (rc) return rc;
        return lr_tc::sweep(P, false, tc_acc16(), f0, N, f1, M, idx1, idx1_2nd, st);
    }
    char *scratch = (char *)lr::arena_get(lr::SLOT_MATCH, nn_scratch_bytes(N, M));
    if (!scratch) return LR_ERR_ALLOC;
    return nn_sweep(f0, N, f1, M, D, idx1, idx1_2nd, scratch, st);
}

LR_EXPORT int lr_match_set_mode(int mode)
{
    lr::Lock lock;
    LR_REQUIRE(mode >= 0 && mode <= 3, "mode must be 0 (auto), 1 (exact CUDA-core sweep), 2 / 3 (tensor-core sweep, fp32 / fp16 accumulators)");
    g_match_mode = mode;
    return LR_OK;
}

LR_EXPORT int lr_match_mutual(const float *f0, int64_t N, const float *f1, int64_t M, int D, const int64_t *idx1,
                              int64_t *out_i, int64_t *out_j, int64_t *K, void *stream)
{
    lr::Lock lock;
    LR_REQUIRE(f0 && f1 && idx1 && out_i && out_j && K, "null pointer");
    LR_REQUIRE(N > 0 && M > 0 && N < ((int64_t)1 << 31) && M < ((int64_t)1 << 31), "N/M out of range");
    cudaStream_t st = (cudaStream_t)stream;
    // reverse sweep: nearest neighbour in f0 of every row of f1.  The reference
    // does it for unique(idx1) only (matching.py:224-225); rows outside that
    // set are never consulted by the intersection, so the result is identical.
    const int nblocks = (int)((N + kCompactBlock - 1) / kCompactBlock);
    const size_t rev_bytes = lr::padded(sizeof(int64_t) * M) + lr::padded(sizeof(int) * nblocks);
    const bool tc = D == 32 && g_match_mode != 1;
    char *scratch = (char *)lr::arena_get(lr::SLOT_MATCH,
                                          rev_bytes + (tc ? lr_tc::scratch_bytes(N, M) : nn_scratch_bytes(M, N)));
    if (!scratch) return LR_ERR_ALLOC;
    int64_t *rev = reinterpret_cast<int64_t *>(scratch);
    int *block_cnt = reinterpret_cast<int *>(scratch + lr::padded(sizeof(int64_t) * M));
    int rc;
    if (tc) {
        lr_tc::Prepared P;
        rc = lr_tc::prepare(f0, N, f1, M, scratch + rev_bytes, P, st);
        if (rc) return rc;
        rc = lr_tc::sweep(P, true, tc_acc16(), f0, N, f1, M, rev, nullptr, st);
    } else {
        rc = nn_sweep(f1, M, f0, N, D, rev, nullptr, scratch + rev_bytes, st);
    }
    if (rc) return rc;
    LR_CUDA_TRY(lr::launch_pdl(k_mutual_count, dim3(nblocks), dim3(kCompactBlock), 0, st, idx1, rev, N, M, block_cnt));
    LR_CUDA_TRY(lr::launch_pdl(k_block_scan, dim3(1), dim3(1024), 0, st, block_cnt, nblocks, K));
    LR_CUDA_TRY(lr::launch_pdl(k_mutual_scatter, dim3(nblocks), dim3(kCompactBlock), 0, st, idx1, rev, N, M, block_cnt, out_i, out_j));
    LR_CUDA_TRY(cudaGetLastError());
    return LR_OK;
}

LR_EXPORT int lr_match_ratio(const float *f0, const float *f1, int D, int64_t K, const int64_t *i0, const int64_t *i1,
                             const int64_t *i2, float *out, void *stream)
{
    LR_REQUIRE(f0 && f1 && i0 && i1 && i2 && out, "null pointer");
    LR_REQUIRE(D > 0 && D % 8 == 0 && K >= 0, "D must be a positive multiple of 8");
    if (K == 0) return LR_OK;
    k_ratio<<<(unsigned)((K + 127) / 128), 128, 0, (cudaStream_t)stream>>>(f0, f1, D, K, i0, i1, i2, out);
    LR_CUDA_TRY(cudaGetLastError());
    return LR_OK;
}

LR_EXPORT int lr_gather_xyz(const float *xyz, const int64_t *idx, int64_t K, float *out, void *stream)
{
    LR_REQUIRE(xyz && idx && out && K >= 0, "bad arguments");
    if (K == 0) return LR_OK;
    k_gather_xyz<<<(unsigned)((K + 255) / 256), 256, 0, (cudaStream_t)stream>>>(xyz, idx, K, out);
    LR_CUDA_TRY(cudaGetLastError());
    return LR_OK;
}
