// lr_ransac_gc.cuh -- GC-RANSAC semantics on top of the batched hypothesis pipeline:
// MSAC selection, local optimisation, iterated least squares.  Included by lr_ransac.cu inside its anonymous
// namespace (shares Ctl / Ws / the canonical fp64 helpers); selected with LrRansacParams.scoring = LR_SCORE_MSAC.
//
// Replaces, of the reference's native glue (GC-RANSAC/src/pygcransac/src/gcransac_python.cpp):
//   MSACScoringFunction<Estimator>                         :507-510   -> k_score_msac
//   settings.max_local_optimization_number = 20            :517       -> lo_trials (k_lo_gen, k_trial_score)
//   settings.max_graph_cut_number (0 when --GC_LO False)   :518-521   -> lo_rounds
//   statistics.inliers / model.descriptor read-back        :594-611   -> k_gc_commit + the common finish path
// The engine behind those settings is un-vendored (danini/graph-cut-ransac); its steps are restated in
// oracle/lr_oracle.c (lro_ransac_gc), which this file follows operation for operation.
//
// Exactness: residuals are the canonical fp64 expression (res2_f64); a term of the score is
// trunc((1 - r^2/tau^2) * 65536) and the score is their INTEGER sum, so any partition of the correspondences
// over threads, CTAs or launches gives the same q, and "highest q, lowest id" is bit-exact against the oracle.
// The selected hypothesis, every LO model (small-sample Kabsch in draw order) and all q values are bit-exact;
// only the least-squares candidates (block-reduced sums over thousands of inliers) differ from the oracle's
// sequential sums, by ~1e-15.
//
// Everything is enqueued without a host round trip: the LO rounds and least-squares passes are launched
// unconditionally and switch themselves off through ctl->gc.lo_active / lsq_active.
#pragma once

constexpr int kMsacThreads = 128;   // one hypothesis per thread
constexpr int kMsacChunk = 256;     // correspondences per shared-memory stage (6 doubles each = 12 KB)
constexpr double kMsacScale = 65536.0;
constexpr uint64_t kLoSeedSalt = 0x4C4F43414C4F5054ULL;  // == LRO_LO_SEED_SALT
constexpr int kLoSampleFactor = 7;  // inner samples of min(7 m, #inliers) points (App. A)
constexpr int kLoMaxSample = 32;

__device__ __forceinline__ long long msac_term(double r2, double tau2)
{
    return (long long)((1.0 - r2 / tau2) * kMsacScale);
}

// q and #(r^2 < tau^2) of every surviving hypothesis of the round.  Work item = 128 slots x a range of
// 256-correspondence chunks (chosen from the survivor count so that every CTA gets work); points are converted
// to fp64 once per stage and read as shared-memory broadcasts.
__global__ void __launch_bounds__(kMsacThreads)
k_score_msac(const float4 *__restrict__ P8, int64_t n, Ctl *ctl,
             const double *__restrict__ m64, unsigned long long *__restrict__ q64, int *__restrict__ cnt, double tau2)
{
    if (ctl->done) return;
    const int nsurv = ctl->n_surv;
    if (nsurv <= 0) return;
    __shared__ double sP[kMsacChunk][6];
    const int slot_blocks = (nsurv + kMsacThreads - 1) / kMsacThreads;
    const int n_chunks = (int)((n + kMsacChunk - 1) / kMsacChunk);
    int nps = (2 * (int)gridDim.x + slot_blocks - 1) / slot_blocks;
    nps = nps < 1 ? 1 : (nps > n_chunks ? n_chunks : nps);
    const int chunks_per = (n_chunks + nps - 1) / nps;
    const int n_items = slot_blocks * nps;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int sb = item / nps, ps = item % nps;
        const int slot = sb * kMsacThreads + threadIdx.x;
        const bool valid = slot < nsurv;
        double T[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) T[k] = valid ? m64[(size_t)slot * 12 + k] : 0.0;
        long long q = 0;
        int c = 0;
        const int c_lo = ps * chunks_per;
        const int c_hi = c_lo + chunks_per < n_chunks ? c_lo + chunks_per : n_chunks;
        for (int ch = c_lo; ch < c_hi; ++ch) {
            __syncthreads();
            const int64_t base = (int64_t)ch * kMsacChunk;
            const int len = (int)(n - base < kMsacChunk ? n - base : kMsacChunk);
            for (int e = threadIdx.x; e < len; e += kMsacThreads) {
                double pp[3], qq[3];
                load_pq(P8, base + e, pp, qq);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    sP[e][k] = pp[k];
                    sP[e][3 + k] = qq[k];
                }
            }
            __syncthreads();
            // first residual component of four points at a time; d0^2 >= tau^2 for the whole warp => the point
            // scores nothing for any of its 32 hypotheses (r^2 >= d0^2, rounding is monotone) and the other two
            // components are skipped (warp-uniform branch, as in k_score)
            for (int j0 = 0; j0 < len; j0 += 4) {
                double d0[4];
                bool live[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + u < len ? j0 + u : len - 1;
                    d0[u] = (((T[0] * sP[j][0] + T[1] * sP[j][1]) + T[2] * sP[j][2]) + T[3]) - sP[j][3];
                    live[u] = __any_sync(0xffffffffu, valid && j0 + u < len && d0[u] * d0[u] < tau2);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (!live[u]) continue;
                    const int j = j0 + u;
                    const double d1 = (((T[4] * sP[j][0] + T[5] * sP[j][1]) + T[6] * sP[j][2]) + T[7]) - sP[j][4];
                    const double d2 = (((T[8] * sP[j][0] + T[9] * sP[j][1]) + T[10] * sP[j][2]) + T[11]) - sP[j][5];
                    const double r2 = (d0[u] * d0[u] + d1 * d1) + d2 * d2;  // == res2_f64
                    if (valid && r2 < tau2) {
                        q += msac_term(r2, tau2);
                        ++c;
                    }
                }
            }
        }
        if (valid && c_lo < c_hi) {
            if (nps == 1) {
                q64[slot] = (unsigned long long)q;
                cnt[slot] = c;
            } else {
                if (q) atomicAdd(&q64[slot], (unsigned long long)q);
                if (c) atomicAdd(&cnt[slot], c);
            }
        }
    }
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w < v ? w : v;
    }
    return v;
}

// pass 1 of the arg-max: highest q of the round (+ the fed-sample hook's per-sample outputs)
__global__ void __launch_bounds__(256)
k_resolve_msac_q(Ctl *ctl, const uint32_t *__restrict__ slot_id, const unsigned long long *__restrict__ q64,
                 const int *__restrict__ cnt, int64_t *__restrict__ scores_out, int32_t *__restrict__ inl_out,
                 int64_t id_base)
{
    if (ctl->done) return;
    const int nsurv = ctl->n_surv;
    unsigned long long best = 0ULL;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < nsurv; slot += gridDim.x * blockDim.x) {
        const unsigned long long q = q64[slot];
        best = q > best ? q : best;
        if (scores_out) {
            const int64_t h = (int64_t)slot_id[slot] - id_base;
            scores_out[h] = (int64_t)q;
            if (inl_out) inl_out[h] = cnt[slot];
        }
    }
    best = warp_max_u64(best);
    if ((threadIdx.x & 31) == 0 && best) atomicMax(&ctl->gc.round_q, best);
}

// pass 2: lowest hypothesis id among the slots that reach it
__global__ void __launch_bounds__(256)
k_resolve_msac_id(Ctl *ctl, const uint32_t *__restrict__ slot_id, const unsigned long long *__restrict__ q64)
{
    if (ctl->done) return;
    const unsigned long long rq = ctl->gc.round_q;
    if (rq == 0ULL) return;
    const int nsurv = ctl->n_surv;
    unsigned long long pick = ~0ULL;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < nsurv; slot += gridDim.x * blockDim.x)
        if (q64[slot] == rq) {
            const unsigned long long k = ((unsigned long long)slot_id[slot] << 32) | (unsigned long long)(uint32_t)slot;
            pick = k < pick ? k : pick;
        }
    pick = warp_min_u64(pick);
    if ((threadIdx.x & 31) == 0 && pick != ~0ULL) atomicMin(&ctl->gc.round_pick, pick);
}

// end of a round: a higher q replaces the selection (rounds ascend in id, so ties keep the lower id); the
// model is copied out of the slot array before the next round reuses it; confidence exit on the
// selection's #(r^2 < tau^2)
__global__ void k_round_end_msac(Ctl *ctl, int64_t round_len, const int *__restrict__ need, int round_idx,
                                 const double *__restrict__ m64, const int *__restrict__ cnt)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (ctl->done) return;
    Ctl::Gc &g = ctl->gc;
    if (g.round_q > g.best_q && g.round_pick != ~0ULL) {
        const uint32_t slot = (uint32_t)(g.round_pick & 0xFFFFFFFFULL);
        g.best_q = g.cur_q = g.round_q;
        g.best_id = (long long)(g.round_pick >> 32);
        g.best_inl = cnt[slot];
        g.has_model = 1;
        for (int k = 0; k < 12; ++k) g.cur[k] = m64[(size_t)slot * 12 + k];
    }
    ctl->iters_run += round_len;
    ctl->n_scored += ctl->n_surv;
    g.round_q = 0ULL;
    g.round_pick = ~0ULL;
    ctl->n_surv = 0;
    ctl->n_events = 0u;
    if (need && g.has_model && g.best_inl >= (long long)need[round_idx]) ctl->done = 1;
}

__global__ void k_gc_mode(Ctl *ctl)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) ctl->gc.mode = 1;
}

// ---- local optimisation ----------------------------------------------------------------------------------

__global__ void k_lo_begin(Ctl *ctl, int lo_rounds)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ctl->gc.lo_active = (ctl->gc.has_model && lo_rounds > 0) ? 1 : 0;
}

// L = { i : |cur p_i - q_i|^2 < thr^2 } in ascending order (graph-cut labelling with weight 0); one block
__global__ void __launch_bounds__(1024)
k_lo_label(const float4 *__restrict__ P8, int64_t n, double thr2, int m, Ctl *ctl, int32_t *__restrict__ L)
{
    if (!ctl->gc.lo_active) return;
    __shared__ int s_wsum[32];
    __shared__ int s_base;
    double T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = ctl->gc.cur[k];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int64_t tile = 0; tile < n; tile += blockDim.x) {
        const int64_t i = tile + threadIdx.x;
        bool in = false;
        if (i < n) {
            double pp[3], qq[3];
            load_pq(P8, i, pp, qq);
            in = res2_f64(T, pp[0], pp[1], pp[2], qq[0], qq[1], qq[2]) < thr2;
        }
        const unsigned b = __ballot_sync(0xffffffffu, in);
        if (lane == 0) s_wsum[w] = __popc(b);
        __syncthreads();
        int before = 0, total = 0;
        for (int k = 0; k < nw; ++k) {
            const int v = s_wsum[k];
            before += (k < w) ? v : 0;
            total += v;
        }
        const int base = s_base;
        if (in) L[base + before + __popc(b & ((1u << lane) - 1u))] = (int32_t)i;
        __syncthreads();
        if (threadIdx.x == 0) s_base = base + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int I = s_base;
        ctl->gc.lo_I = I;
        ctl->gc.lo_s = I < kLoSampleFactor * m ? I : kLoSampleFactor * m;
        if (I <= m) ctl->gc.lo_active = 0;
    }
}

// k <= kLoMaxSample unique positions out of [0, n): draw d picks the r-th position not yet taken (== lro_unique_k)
__device__ void unique_ids_rt(uint64_t seed, uint64_t id, int k, int64_t n, int32_t *out)
{
    int32_t taken[kLoMaxSample];
    for (int d = 0; d < k; ++d) {
        int32_t r = (int32_t)draw(seed, id, (uint32_t)d, (uint32_t)(n - d));
        for (int e = 0; e < d; ++e)
            if (r >= taken[e]) ++r;
        out[d] = r;
        int e = d;
        while (e > 0 && taken[e - 1] > r) {
            taken[e] = taken[e - 1];
            --e;
        }
        taken[e] = r;
    }
}

// Kabsch over k points in the given order (== lro_kabsch: sequential sums)
__device__ void kabsch_rt(const double (*P)[3], const double (*Q)[3], int k, double (&T)[12])
{
    double cp[3] = {0, 0, 0}, cq[3] = {0, 0, 0};
    for (int i = 0; i < k; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            cp[c] = cp[c] + P[i][c];
            cq[c] = cq[c] + Q[i][c];
        }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        cp[c] = cp[c] / (double)k;
        cq[c] = cq[c] / (double)k;
    }
    double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int i = 0; i < k; ++i) {
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

// one thread per inner draw of the round: sample of L -> non-minimal Kabsch
__global__ void __launch_bounds__(kGcMaxTrials)
k_lo_gen(const float4 *__restrict__ P8, uint64_t lo_seed, int round, int trials,
         Ctl *ctl, const int32_t *__restrict__ L, double *__restrict__ tr_T, unsigned long long *__restrict__ tr_q,
         int *__restrict__ tr_inl)
{
    if (!ctl->gc.lo_active) return;
    const int t = threadIdx.x;
    if (t >= trials) return;
    const int I = ctl->gc.lo_I, s = ctl->gc.lo_s;
    int32_t pos[kLoMaxSample];
    double P[kLoMaxSample][3], Q[kLoMaxSample][3], T[12];
    unique_ids_rt(lo_seed, (uint64_t)round * (uint64_t)trials + (uint64_t)t, s, I, pos);
    for (int d = 0; d < s; ++d) load_pq(P8, (int64_t)L[pos[d]], P[d], Q[d]);
    kabsch_rt(P, Q, s, T);
#pragma unroll
    for (int k = 0; k < 12; ++k) tr_T[t * 12 + k] = T[k];
    tr_q[t] = 0ULL;
    tr_inl[t] = 0;
}

// q and #(r^2 < tau^2) of `trials` candidate models over all correspondences: thread = correspondence, the
// models sit in shared memory, integer warp / block reductions
__global__ void __launch_bounds__(256)
k_trial_score(const float4 *__restrict__ P8, int64_t n, double tau2, int trials,
              const int *__restrict__ active, const double *__restrict__ tr_T, unsigned long long *__restrict__ tr_q,
              int *__restrict__ tr_inl)
{
    if (!*active) return;
    __shared__ double sT[kGcMaxTrials * 12];
    __shared__ unsigned long long sq[kGcMaxTrials];
    __shared__ int sc[kGcMaxTrials];
    for (int e = threadIdx.x; e < trials * 12; e += blockDim.x) sT[e] = tr_T[e];
    for (int e = threadIdx.x; e < trials; e += blockDim.x) {
        sq[e] = 0ULL;
        sc[e] = 0;
    }
    __syncthreads();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool have = i < n;
    double px = 0, py = 0, pz = 0, qx = 0, qy = 0, qz = 0;
    if (have) {
        double pp[3], qq[3];
        load_pq(P8, i, pp, qq);
        px = pp[0], py = pp[1], pz = pp[2];
        qx = qq[0], qy = qq[1], qz = qq[2];
    }
    const int lane = threadIdx.x & 31;
    for (int t = 0; t < trials; ++t) {
        long long q = 0;
        int c = 0;
        if (have) {
            const double r2 = res2_f64(&sT[12 * t], px, py, pz, qx, qy, qz);
            if (r2 < tau2) {
                q = msac_term(r2, tau2);
                c = 1;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            q += __shfl_xor_sync(0xffffffffu, q, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (lane == 0 && c) {
            atomicAdd(&sq[t], (unsigned long long)q);
            atomicAdd(&sc[t], c);
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < trials; e += blockDim.x)
        if (sc[e]) {
            atomicAdd(&tr_q[e], sq[e]);
            atomicAdd(&tr_inl[e], sc[e]);
        }
}

// best inner draw (highest q, lowest trial on ties) replaces cur if it scores higher, else LO stops
__global__ void k_lo_select(Ctl *ctl, int trials, const double *__restrict__ tr_T,
                            const unsigned long long *__restrict__ tr_q)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Ctl::Gc &g = ctl->gc;
    if (!g.lo_active) return;
    int tb = -1;
    unsigned long long qb = 0ULL;
    for (int t = 0; t < trials; ++t)
        if (tr_q[t] > qb) {
            qb = tr_q[t];
            tb = t;
        }
    if (tb >= 0 && qb > g.cur_q) {
        g.cur_q = qb;
        for (int k = 0; k < 12; ++k) g.cur[k] = tr_T[tb * 12 + k];
        ++g.lo_improved;
    } else {
        g.lo_active = 0;
    }
}

// ---- iterated least squares --------------------------------------------------------------------------------

// first = 1: LO is over (snapshot its score, arm the least squares).  first = 0: judge the candidate of the
// previous pass.  Either way, if still active, stage cur as ctl->T for the next refit pass.
__global__ void k_lsq_step(Ctl *ctl, int first, int lsq_iters, const double *__restrict__ tr_T,
                           const unsigned long long *__restrict__ tr_q)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Ctl::Gc &g = ctl->gc;
    if (first) {
        g.lo_active = 0;
        g.lo_q = g.cur_q;
        g.lsq_active = (g.has_model && lsq_iters > 0) ? 1 : 0;
    } else if (g.lsq_active) {
        if (tr_q[0] > g.cur_q) {
            g.cur_q = tr_q[0];
            for (int k = 0; k < 12; ++k) g.cur[k] = tr_T[k];
            ++g.lsq_improved;
        } else {
            g.lsq_active = 0;
        }
    }
    for (int k = 0; k < 12; ++k) ctl->T[k] = g.cur[k];
    ctl->refit_count = 0;
    ctl->err2 = 0.0;
    for (int k = 0; k < 6; ++k) ctl->csum[k] = 0.0;
    for (int k = 0; k < 9; ++k) ctl->H[k] = 0.0;
}

// the refit of this pass (ctl->Tref over ctl->refit_count inliers at thr) becomes the candidate
__global__ void k_lsq_stage(Ctl *ctl, int m, double *__restrict__ tr_T, unsigned long long *__restrict__ tr_q,
                            int *__restrict__ tr_inl)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Ctl::Gc &g = ctl->gc;
    if (!g.lsq_active) return;
    if (ctl->refit_count < (long long)m) {
        g.lsq_active = 0;
        return;
    }
    for (int k = 0; k < 12; ++k) tr_T[k] = ctl->Tref[k];
    tr_q[0] = 0ULL;
    tr_inl[0] = 0;
}

// final model -> ctl->T (identity when nothing scored above 0), accumulators of the common finish path cleared
__global__ void k_gc_commit(Ctl *ctl)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Ctl::Gc &g = ctl->gc;
    g.lo_active = g.lsq_active = 0;
    for (int k = 0; k < 12; ++k) ctl->T[k] = g.has_model ? g.cur[k] : ((k % 5 == 0) ? 1.0 : 0.0);
    ctl->refit_count = 0;
    ctl->err2 = 0.0;
    for (int k = 0; k < 6; ++k) ctl->csum[k] = 0.0;
    for (int k = 0; k < 9; ++k) ctl->H[k] = 0.0;
}

// ---- host side -----------------------------------------------------------------------------------------------

inline double gc_tau2(double threshold) { return (1.5 * threshold) * (1.5 * threshold); }

// scoring + arg-max of one round in LR_SCORE_MSAC mode (after k_gen / k_kabsch)
int gc_launch_score(const float *src, const float *tgt, int64_t n, const LrRansacParams &p, const Ws &ws, int64_t lo,
                    int64_t len, int64_t *scores_out, int32_t *inl_out, cudaStream_t st)
{
    const int sms = lr::sm_count();
    const int64_t slots = len + kHypPerItem;
    // split items add into q64 (k_kabsch already cleared cnt)
    LR_CUDA_TRY(cudaMemsetAsync(ws.q64, 0, sizeof(unsigned long long) * (size_t)(len < slots ? len : slots), st));
    int tok = lr::prof_begin(lr::PROF_SCORE, st);
    k_score_msac<<<sms * 4, kMsacThreads, 0, st>>>(ws.P8, n, ws.ctl, ws.m64, ws.q64, ws.cnt, gc_tau2(p.threshold));
    lr::prof_end(tok, st);
    int rblocks = (int)((len + 255) / 256);
    if (rblocks > sms * 4) rblocks = sms * 4;
    k_resolve_msac_q<<<rblocks, 256, 0, st>>>(ws.ctl, ws.slot_id, ws.q64, ws.cnt, scores_out, inl_out, lo);
    k_resolve_msac_id<<<rblocks, 256, 0, st>>>(ws.ctl, ws.slot_id, ws.q64);
    LR_CUDA_TRY(cudaGetLastError());
    return LR_OK;
}

// local optimisation + iterated least squares on the selection left in ctl->gc (launches only)
int gc_enqueue_polish(const float *src, const float *tgt, int64_t n, const LrRansacParams &p, const Ws &ws,
                      cudaStream_t st)
{
    const double thr2 = p.threshold * p.threshold, tau2 = gc_tau2(p.threshold);
    const int trials = p.lo_trials < 1 ? 1 : (p.lo_trials > kGcMaxTrials ? kGcMaxTrials : p.lo_trials);
    // no inner draws = no local optimisation (the oracle's trial loop is then empty and the stage ends at once)
    const int lo_rounds = (p.lo_rounds < 0 || p.lo_trials < 1) ? 0 : p.lo_rounds, lsq_iters = p.lsq_iters < 0 ? 0 : p.lsq_iters;
    const uint64_t lo_seed = mix64(p.seed ^ kLoSeedSalt);
    const float *P8f = reinterpret_cast<const float *>(ws.P8);  // fetch_pair's packed-record convention
    const int sblocks = (int)((n + 255) / 256) > 0 ? (int)((n + 255) / 256) : 1;
    int rblocks = sblocks;
    if (rblocks > lr::sm_count() * 8) rblocks = lr::sm_count() * 8;
    k_lo_begin<<<1, 32, 0, st>>>(ws.ctl, lo_rounds);
    for (int rd = 0; rd < lo_rounds; ++rd) {
        k_lo_label<<<1, 1024, 0, st>>>(ws.P8, n, thr2, p.sample_size, ws.ctl, ws.lo_L);
        k_lo_gen<<<1, kGcMaxTrials, 0, st>>>(ws.P8, lo_seed, rd, trials, ws.ctl, ws.lo_L, ws.tr_T, ws.tr_q, ws.tr_inl);
        k_trial_score<<<sblocks, 256, 0, st>>>(ws.P8, n, tau2, trials, &ws.ctl->gc.lo_active, ws.tr_T, ws.tr_q,
                                               ws.tr_inl);
        k_lo_select<<<1, 32, 0, st>>>(ws.ctl, trials, ws.tr_T, ws.tr_q);
    }
    for (int it = 0; it <= lsq_iters; ++it) {
        k_lsq_step<<<1, 32, 0, st>>>(ws.ctl, it == 0, lsq_iters, ws.tr_T, ws.tr_q);
        if (it == lsq_iters) break;
        k_mask_sums<<<rblocks, 256, 0, st>>>(nullptr, P8f, nullptr, nullptr, n, thr2, ws.ctl, nullptr);
        k_refit_H<<<rblocks, 256, 0, st>>>(nullptr, P8f, nullptr, nullptr, n, thr2, ws.ctl);
        k_refit_solve<<<1, 32, 0, st>>>(ws.ctl);
        k_lsq_stage<<<1, 32, 0, st>>>(ws.ctl, p.sample_size, ws.tr_T, ws.tr_q, ws.tr_inl);
        k_trial_score<<<sblocks, 256, 0, st>>>(ws.P8, n, tau2, 1, &ws.ctl->gc.lsq_active, ws.tr_T, ws.tr_q, ws.tr_inl);
    }
    LR_CUDA_TRY(cudaGetLastError());
    return LR_OK;
}
