// lr_icp_api.cuh -- C-ABI entries of SURVEY 8(f4) (kernels: lr_icp.cuh).  Included at the end of lr_ransac.cu.
#pragma once

namespace {

struct IcpWs {
    IcpCtl *ctl;
    double *T12;
    double *partial;
    Grid g;
};

// carve the grid + control block out of the MISC arena and bin the target cloud (three launches)
int icp_setup(const float *tgt, int64_t m, double max_dist, IcpWs &w, cudaStream_t st)
{
    unsigned int size = 1024;
    while ((int64_t)size < 2 * m) size <<= 1;
    const size_t bytes = lr::padded(sizeof(IcpCtl)) + lr::padded(sizeof(double) * 16) +
                         lr::padded(sizeof(double) * kFinBlocksMax * kFinVals) + lr::padded(sizeof(unsigned long long) * size) +
                         2 * lr::padded(sizeof(unsigned int) * size) + lr::padded(sizeof(float4) * (size_t)(m > 0 ? m : 1)) +
                         2 * lr::padded(sizeof(unsigned int) * (size_t)(m > 0 ? m : 1));
    void *base = lr::arena_get(lr::SLOT_MISC, bytes);
    if (!base) return LR_ERR_ALLOC;
    lr::Carver cv(base);
    w.ctl = cv.take<IcpCtl>(1);
    w.T12 = cv.take<double>(16);
    w.partial = cv.take<double>((size_t)kFinBlocksMax * kFinVals);
    w.g.keys = cv.take<unsigned long long>(size);
    w.g.cnt = cv.take<unsigned int>(size);
    w.g.start = cv.take<unsigned int>(size);
    w.g.pts = cv.take<float4>((size_t)(m > 0 ? m : 1));
    w.g.pt_slot = cv.take<unsigned int>((size_t)(m > 0 ? m : 1));
    w.g.pt_rank = cv.take<unsigned int>((size_t)(m > 0 ? m : 1));
    w.g.mask = size - 1u;
    w.g.inv_cell = 1.0 / (max_dist * (1.0 + 1e-7));
    LR_CUDA_TRY(cudaMemsetAsync(w.g.keys, 0xFF, sizeof(unsigned long long) * size, st));
    LR_CUDA_TRY(cudaMemsetAsync(w.g.cnt, 0, sizeof(unsigned int) * size, st));
    if (m > 0) {
        const unsigned blocks = (unsigned)((m + 255) / 256);
        k_grid_insert<<<blocks, 256, 0, st>>>(tgt, m, w.g);
        k_grid_scan<<<1, 1024, 0, st>>>(w.g);
        k_grid_scatter<<<blocks, 256, 0, st>>>(tgt, m, w.g);
    } else {
        k_grid_scan<<<1, 1024, 0, st>>>(w.g);
    }
    LR_CUDA_TRY(cudaGetLastError());
    return LR_OK;
}

void T16_to_12(const double *T16, double *T12)
{
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) T12[4 * r + c] = T16 ? T16[4 * r + c] : (r == c ? 1.0 : 0.0);
}

}  // namespace

LR_EXPORT int lr_nn3d_radius(const float *src, int64_t n, const float *tgt, int64_t m, const double *T_in, double radius,
                             int64_t *idx, double *d2, void *stream)
{
    lr::Lock lock;
    LR_REQUIRE(n >= 0 && m >= 0 && n < ((int64_t)1 << 31) && m < ((int64_t)1 << 31), "n / m out of range");
    LR_REQUIRE(radius > 0.0, "radius must be positive");
    if (n == 0) return LR_OK;
    LR_REQUIRE(src && idx && (m == 0 || tgt), "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    IcpWs w;
    int rc = icp_setup(tgt, m, radius, w, st);
    if (rc) return rc;
    double T12[12];
    T16_to_12(T_in, T12);
    LR_CUDA_TRY(cudaMemcpyAsync(w.T12, T12, sizeof(T12), cudaMemcpyHostToDevice, st));  // pageable source: staged before return
    k_nn3d_query<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, n, w.g, w.T12, radius * radius, idx, d2);
    LR_CUDA_TRY(cudaGetLastError());
    LR_CUDA_TRY(cudaStreamSynchronize(st));
    return LR_OK;
}

LR_EXPORT int lr_icp_refine(const float *src, int64_t n, const float *tgt, int64_t m, double max_dist, const double *T_init,
                            int max_iteration, double rel_fitness, double rel_rmse, double *T_out, double *fitness,
                            double *inlier_rmse, int *iterations, void *stream)
{
    lr::Lock lock;
    LR_REQUIRE(T_out != nullptr, "T_out is null");
    LR_REQUIRE(n >= 0 && m >= 0 && n < ((int64_t)1 << 31) && m < ((int64_t)1 << 31), "n / m out of range");
    LR_REQUIRE(max_dist > 0.0 && max_iteration >= 0 && max_iteration <= 100000, "max_dist / max_iteration out of range");
    LR_REQUIRE((n == 0 || src) && (m == 0 || tgt), "null pointer");
    double T12[12];
    T16_to_12(T_init, T12);
    if (n == 0 || m == 0) {  // nothing to match: the initial transform, fitness 0
        T12_to_16(T12, T_out);
        if (fitness) *fitness = 0.0;
        if (inlier_rmse) *inlier_rmse = 0.0;
        if (iterations) *iterations = 0;
        return LR_OK;
    }
    cudaStream_t st = (cudaStream_t)stream;
    IcpWs w;
    int rc = icp_setup(tgt, m, max_dist, w, st);
    if (rc) return rc;
    LR_CUDA_TRY(cudaMemcpyAsync(w.T12, T12, sizeof(T12), cudaMemcpyHostToDevice, st));
    k_icp_begin<<<1, 32, 0, st>>>(w.ctl, w.T12);
    int blocks = (int)((n * kIcpLanes + 255) / 256);
    const int cap = lr::sm_count() * 4 < kFinBlocksMax ? lr::sm_count() * 4 : kFinBlocksMax;
    if (blocks > cap) blocks = cap;
    // evaluation 0 scores T_init, evaluation e >= 1 is Open3D's iteration e; a converged run turns the rest into no-ops
    for (int e = 0; e <= max_iteration; ++e)
        LR_CUDA_TRY(lr::launch_pdl(k_icp_eval, dim3(blocks), dim3(256), 0, st, src, n, tgt, w.g, max_dist * max_dist, w.ctl,
                                   w.partial, e, rel_fitness, rel_rmse));
    LR_CUDA_TRY(cudaGetLastError());
    IcpCtl h;
    LR_CUDA_TRY(cudaMemcpyAsync(&h, w.ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
    LR_CUDA_TRY(cudaStreamSynchronize(st));
    T12_to_16(h.Tres, T_out);
    if (fitness) *fitness = h.fitness;
    if (inlier_rmse) *inlier_rmse = h.rmse;
    if (iterations) *iterations = h.it;
    return LR_OK;
}

LR_EXPORT int lr_kabsch_weighted_batch(const float *A, const float *B, const float *w, int64_t S, int k, double *T_out,
                                       void *stream)
{
    lr::Lock lock;
    LR_REQUIRE(S >= 0 && k >= 0, "S / k out of range");
    if (S == 0) return LR_OK;
    LR_REQUIRE(T_out && (k == 0 || (A && B)), "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    k_kabsch_weighted<<<(unsigned)((S + 127) / 128), 128, 0, st>>>(A, B, w, S, k, T_out);
    LR_CUDA_TRY(cudaGetLastError());
    return LR_OK;
}

LR_EXPORT int lr_seeds_score(const float *src, const float *tgt, int64_t n, const double *models, int64_t S,
                             double threshold, int32_t *counts, uint8_t *labels, int64_t *best, int64_t *best_count,
                             double *T_best, double *T_refit, void *stream)
{
    lr::Lock lock;
    LR_REQUIRE(src && tgt && models && T_best, "null pointer");
    LR_REQUIRE(n > 0 && n < ((int64_t)1 << 31) && S > 0 && S <= ((int64_t)1 << 20), "n / S out of range");
    LR_REQUIRE(threshold > 0.0, "threshold must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    LrRansacParams p;
    memset(&p, 0, sizeof(p));
    p.threshold = threshold;
    p.confidence = 1.0;
    p.max_iters = S;
    p.sample_size = 3;
    p.sampler = LR_SAMPLER_UNIFORM;
    p.round_size = 1;
    p.refit = 1;
    p.scoring = LR_SCORE_COUNT;
    Ws ws;
    int rc = ws_setup(n, S, 1, ws);
    if (rc) return rc;
    rc = launch_pack(src, tgt, n, ws, st, false);
    if (rc) return rc;
    const double thr2 = threshold * threshold;
    LR_CUDA_TRY(tc_smem_attr());
    k_seed_install<<<(int)((S + 127) / 128), 128, 0, st>>>(models, (int)S, ws.P8, thr2, ws.ctl, ws.m64, ws.Aimg, ws.band, ws.cnt,
                                                          ws.slot_id);
    tcs::k_score_tc<false><<<lr::sm_count(), tcs::NTHREADS, tcs::kSmemBytes, st>>>(ws.Aimg, ws.Bimg, ws.P8, n, ws.n_pad, ws.ctl,
                                                                                    ws.m64, ws.band, ws.cnt, thr2, nullptr,
                                                                                    ws.events, kTcEventCap);
    k_seed_end<<<1, 256, 0, st>>>(ws.ctl, ws.cnt, ws.m64, (int)S, counts);
    LR_CUDA_TRY(cudaGetLastError());
    LrRansacStats stats;
    rc = finish(src, tgt, n, p, ws, FIN_MODEL_READY, 0, T_best, T_refit, labels, &stats, st);
    if (rc) return rc;
    if (best) *best = stats.best_id;
    if (best_count) *best_count = stats.best_count;
    return LR_OK;
}
