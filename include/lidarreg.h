/*
 * lidarreg.h -- C ABI of the B200-native robust-registration hot path.
 *
 * Drop-in boundary for AmnonDrory/LidarRegistration's
 *   Experiments/test.py --algo RANSAC --mode {MNN|MMN|no_filter}
 * i.e. the work below Experiments/algorithms/FR.py:16 (citations relative to
 * the reference tree).  Each entry point names the reference interface it
 * replaces.  All array arguments are DEVICE pointers owned by the caller
 * unless marked [host]; `stream` is a cudaStream_t passed as void*.  Every
 * function returns 0 on success and a non-zero LrStatus otherwise, with a
 * message available from lr_last_error(); nothing throws across the boundary.
 * There is no CPU fallback: without a CUDA device every compute entry fails
 * with LR_ERR_CUDA.
 *
 * Transforms are 4x4 row-major doubles in the column-vector convention
 * (q ~ T[:3,:3] p + T[:3,3]), i.e. what FR() returns (FR.py:119) -- the
 * transpose of pygcransac's row-vector pose (GC_RANSAC.py:55).
 */
#ifndef LIDARREG_H
#define LIDARREG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum LrStatus {
    LR_OK = 0,
    LR_ERR_ARG = 1,   /* bad argument (null pointer, unsupported D / sample size, ...) */
    LR_ERR_CUDA = 2,  /* CUDA runtime error, or no device */
    LR_ERR_ALLOC = 3  /* workspace allocation failed */
} LrStatus;

/* sampler ids follow gcransac_python.cpp:462-465 (0 uniform, 1 PROSAC) and add
 * Open3D's with-replacement draw */
enum { LR_SAMPLER_UNIFORM = 0, LR_SAMPLER_PROSAC = 1, LR_SAMPLER_REPLACE = 2 };

/* Model selection criterion (SURVEY 8(a) "two scoring semantics"):
 *   LR_SCORE_COUNT  inlier iff r^2 < thr^2, best = highest count, ties -> lowest hypothesis id (Open3D's
 *                   EvaluateRANSACBasedOnCorrespondence, FR.py:122-139; the graded, bit-exact criterion);
 *   LR_SCORE_MSAC   GC-RANSAC's MSACScoringFunction (gcransac_python.cpp:507-510): inlier iff
 *                   r^2 < tau^2, tau = 1.5 thr, best = highest sum (1 - r^2/tau^2).  Every term is quantised
 *                   to 2^-16 and summed as an integer, q = sum trunc((1 - r^2/tau^2) * 65536), so that the
 *                   winner does not depend on the summation order (thread / GPU count); ties -> lowest id. */
enum { LR_SCORE_COUNT = 0, LR_SCORE_MSAC = 1 };

/* Parameters of one RANSAC run.  Mirrors the kwargs of
 * pygcransac.findRigidTransform (GC_RANSAC.py:12-22) and of Open3D's
 * registration_ransac_based_on_correspondence (FR.py:128-137). */
typedef struct LrRansacParams {
    double threshold;     /* inlier distance, metres (FR.py:85,95: 2*voxel = 0.6) */
    double confidence;    /* stopping confidence; >= 1 disables the early exit (fixed budget) */
    double elc_ratio;     /* edge-length similarity (preemption_edge_length.h:82: 0.9) */
    int64_t max_iters;    /* hypothesis budget (--iters) */
    uint64_t seed;        /* hypothesis h is a pure function of (seed, h) */
    int32_t sample_size;  /* 3 (GC minimal sample) or 4 (FR.py:134 ransac_n) */
    int32_t sampler;      /* LR_SAMPLER_* */
    int32_t use_elc;      /* edge-length pre-rejection on/off (--fast_rejection ELC|NONE) */
    int32_t round_size;   /* hypotheses per round; the confidence exit is evaluated at round ends */
    int32_t refit;        /* also return the least-squares refit over the selected model's inliers */
    int32_t scoring;      /* LR_SCORE_COUNT (graded criterion, Open3D semantics) or LR_SCORE_MSAC (GC semantics) */
    /* GC-RANSAC finishing steps; used with LR_SCORE_MSAC only */
    int32_t lo_rounds;    /* local-optimisation rounds (settings.max_graph_cut_number, default 10; 0 = --GC_LO False,
                             gcransac_python.cpp:518-521) */
    int32_t lo_trials;    /* inner draws per round (settings.max_local_optimization_number = 20,
                             gcransac_python.cpp:517), at most 64 */
    int32_t lsq_iters;    /* iterated least-squares passes over the inliers (upstream: at most 10) */
    int32_t reserved;
} LrRansacParams;

typedef struct LrRansacStats {
    int64_t iters_run;    /* hypotheses generated (multiple of round_size unless capped by max_iters) */
    int64_t n_scored;     /* hypotheses that passed ELC and were scored against all correspondences */
    int64_t n_rechecked;  /* scored hypotheses whose fp32 count was ambiguous and was redone in fp64 */
    int64_t best_id;      /* selected hypothesis (-1: none) */
    int64_t best_count;   /* its inlier count (exact, fp64 semantics) */
    int64_t refit_count;  /* inliers used by the refit */
    /* LR_SCORE_MSAC only (0 otherwise); with it best_count = #(r^2 < tau^2) of hypothesis best_id */
    int64_t best_score;   /* q of the selected minimal-sample hypothesis */
    int64_t lo_score;     /* q after the local optimisation */
    int64_t final_score;  /* q of T_out (after the iterated least squares) */
    int32_t lo_improved;  /* local-optimisation rounds that raised q */
    int32_t lsq_improved; /* least-squares passes that raised q */
} LrRansacStats;

/* ---- library ---------------------------------------------------------- */
const char *lr_last_error(void);
int lr_version(void); /* 130: lr_gpf_filter, LR_PROF_PACK / END / FIN; 120: lr_comm_*, lr_ransac_rigid_sharded, lr_ransac_tc_probe, three sweep modes (110: LR_SCORE_MSAC) */
/* [host] outputs; number of SMs and compute capability of the current device */
int lr_device_info(int *sm_count, int *cc_major, int *cc_minor);
/* release every device workspace held by the library */
int lr_shutdown(void);

/* ---- measurement hooks (used by bench.py only) -------------------------- */
enum { LR_PROF_SCORE = 0, LR_PROF_GEN = 1, LR_PROF_NN = 2, LR_PROF_RECOUNT = 3,
       LR_PROF_PACK = 4 /* reset + pack */, LR_PROF_END = 5 /* round end (+ key exchange) */, LR_PROF_FIN = 6 /* mask + refit */ };
/* bracket the heavy kernels with CUDA events on their stream (off by default) */
int lr_prof_enable(int on);
/* [host] device milliseconds and launch count of one kernel class since the last read */
int lr_prof_read(int kind, double *total_ms, int64_t *launches);
/* [host] measured FP32 FMA throughput (TFLOP/s): mode 0 = FFMA, 1 = packed fma.rn.f32x2 */
int lr_peak_fp32(int mode, double *tflops);

/* ---- correspondence search (Experiments/algorithms/matching.py) ------- */

/* find_nn (matching.py:22-65).  f0[N,D], f1[M,D] fp32 row-major; D multiple of
 * 8, D <= 128.  idx1[N] = argmin_j sqrt(max(|f0_i|^2+|f1_j|^2-2 f0_i.f1_j,1e-30)),
 * lowest j on ties; idx1_2nd[N] (nullable) = same with column idx1[i] masked
 * (matching.py:36-37).  Indices are int64 like the reference's. */
int lr_match_nn(const float *f0, int64_t N, const float *f1, int64_t M, int D, int64_t *idx1,
                int64_t *idx1_2nd, void *stream);

/* Implementation switch for lr_match_nn / lr_match_mutual (both give identical indices):
 * 0 = tensor-core sweep (tcgen05, fp16 operands) + exact fp32 re-rank when D == 32 [default];
 * 1 = exact fp32 CUDA-core sweep for every D;
 * 2 / 3 = the tensor-core sweep with fp32 / fp16 accumulators in TMEM, explicitly (A/B tests). */
int lr_match_set_mode(int mode);

/* nn_to_mutual (matching.py:222-239) incl. torch_intersect (:67-87): keeps
 * (i, idx1[i]) iff i is the nearest neighbour of idx1[i] in f0.  out_i/out_j
 * have room for N entries and come out sorted by i; *K [device] = count. */
int lr_match_mutual(const float *f0, int64_t N, const float *f1, int64_t M, int D, const int64_t *idx1,
                    int64_t *out_i, int64_t *out_j, int64_t *K, void *stream);

/* calc_distance_ratio_in_feature_space (matching.py:89-98): out[k] =
 * |f0[i0[k]]-f1[i1[k]]| / (|f0[i0[k]]-f1[i2[k]]| + 1e-6), fp32. */
int lr_match_ratio(const float *f0, const float *f1, int D, int64_t K, const int64_t *i0, const int64_t *i1,
                   const int64_t *i2, float *out, void *stream);

/* xyz[idx] gather that builds the correspondence arrays (FR.py:72-73):
 * out[k,:] = xyz[idx[k],:], fp32 [.,3]. */
int lr_gather_xyz(const float *xyz, const int64_t *idx, int64_t K, float *out, void *stream);

/* Grid_Prioritized_Filter (matching.py:100-205, --mode GPF), the part after best-buddy marking and the
 * ratio quality.  ratio[n] = lr_match_ratio of the candidate pairs, is_bb[n] 1 for mutual pairs (nullable: the BB_first
 * variant has no offset), xyz0[.,3] source cloud, idx0[n] source index per pair (nullable = identity), total_num =
 * GPF_factor * #best buddies (or GPF_max_matches).  Out: keep[n] 0 / 1, norm[n] = (ratio - min) / (max - min), - 1 for
 * best buddies: the value the reference returns for the kept pairs (PROSAC quality, FR.py:75-76).  All device. */
int lr_gpf_filter(const float *ratio, const uint8_t *is_bb, const float *xyz0, const int64_t *idx0, int64_t n, int grid_wid,
                  double total_num, uint8_t *keep, float *norm, void *stream);

/* ---- RANSAC rigid motion ---------------------------------------------- */

/* Replaces pygcransac.findRigidTransform (GC_RANSAC.py:46-49,
 * gcransac_python.cpp:404-624) and Open3D's
 * registration_ransac_based_on_correspondence (FR.py:128-137): src[n,3],
 * tgt[n,3] fp32 correspondences -- device memory, or pinned host memory under
 * unified addressing: both arrays are read once by the pack kernel (plus the
 * three rows of the selected sample), every sweep works on packed device copies.  T_out[16] [host] = selected model (identity
 * if none); T_refit[16] [host, nullable] = Kabsch over its inliers
 * (FR.py:99-111); mask[n] (device, nullable) = inlier mask of the selected
 * model; stats [host, nullable].  Synchronises `stream` before returning.
 * With params->scoring == LR_SCORE_MSAC the run follows GC-RANSAC: MSAC selection, then
 * lo_rounds rounds of local optimisation (inner draws of min(7 m, #inliers) inliers -> non-minimal Kabsch; the
 * graph cut with spatial_coherence_weight = 0, test.py:306, is thresholding at `threshold`), then lsq_iters
 * passes of iterated least squares, each kept only while q improves; T_out = the final model. */
int lr_ransac_rigid(const float *src, const float *tgt, int64_t n, const LrRansacParams *params,
                    double *T_out, double *T_refit, uint8_t *mask, LrRansacStats *stats, void *stream);

/* Implementation switch of the LR_SCORE_COUNT inlier sweep (identical results):
 * 0 = tensor-core sweep [default]: the three residual components as tcgen05 MMAs over fp16 operand pieces,
 *     residuals inside the proven error band decided in fp64 (csrc/lr_score_tc.cuh);
 * 1 = fp32 CUDA-core sweep, every residual in full; 2 = the same with the warp-uniform early-out on the first
 *     residual component (round 1's default).  A/B measurements and parity tests. */
int lr_ransac_set_mode(int mode);

/* The same for `count` independent pairs (the reference's per-pair loop over a registration set,
 * Experiments/test.py:108-167, sharded by rank as in data_loaders.py:111-116): src[i] / tgt[i] point to
 * [n[i],3] fp32 correspondences in device memory OR in host memory (pinned for full PCIe rate; what the
 * reference's loop holds, GC_RANSAC.py:10-11) -- host arrays are brought in by the copy engine on the pair's
 * stream, under the kernels of the pair before it.  Two pairs are in flight at a time on two internal streams with
 * separate scratch, so one pair's single-block tail kernels run under the next pair's chip-filling ones and
 * there is no host round trip between pairs.  T_out[16*count] [host], T_refit[16*count] [host, nullable],
 * stats[count] [host, nullable]; results are those of `count` lr_ransac_rigid calls.  The work is ordered
 * after what `stream` holds at the call; synchronises before returning. */
int lr_ransac_rigid_batch(const float *const *src, const float *const *tgt, const int64_t *n, int count,
                          const LrRansacParams *params, double *T_out, double *T_refit, LrRansacStats *stats,
                          void *stream);

/* Fed-sample parity hook (BASELINE.json: "same fed hypothesis triplets").
 * samples[H,m] int32 (m = 3 or 4).  counts[H] = exact inlier count of each
 * sample's Kabsch model, -1 where ELC rejects it; models[H,12] (nullable) =
 * the fp64 [R|t] rows; *best [host] = argmax count, lowest h on ties (-1 if
 * every sample is rejected).  Synchronises `stream`. */
int lr_ransac_score_samples(const float *src, const float *tgt, int64_t n, const int32_t *samples, int64_t H,
                            int m, double threshold, int use_elc, double elc_ratio, int32_t *counts,
                            double *models, int64_t *best, void *stream);

/* The same hook for LR_SCORE_MSAC: scores[H] = quantised MSAC score q of each sample's Kabsch model (-1 where
 * ELC rejects it), inliers[H] (nullable) = #(r^2 < (1.5 threshold)^2); *best [host] = argmax q, lowest h on
 * ties (-1 if none scored above 0).  Synchronises `stream`. */
int lr_ransac_score_samples_msac(const float *src, const float *tgt, int64_t n, const int32_t *samples, int64_t H,
                                 int m, double threshold, int use_elc, double elc_ratio, int64_t *scores,
                                 int32_t *inliers, int64_t *best, void *stream);

/* Multi-GPU hypothesis sharding, caller-owned collective (the transport-agnostic form: the caller
 * all-reduces the key with NCCL / gloo / MPI; lr_ransac_rigid_sharded is the fused form): score hypotheses [id_lo, id_hi)
 * of the run described by `params` and max-merge the packed result
 *   key = (count + 1) << 32 | (0xFFFFFFFF - id)
 * into *key [device, uint64].  Asynchronous on `stream`.  The caller
 * all-reduces (MAX) the keys across ranks between rounds. */
int lr_ransac_shard(const float *src, const float *tgt, int64_t n, const LrRansacParams *params, int64_t id_lo,
                    int64_t id_hi, uint64_t *key, void *stream);

/* Turn a (reduced) key back into the model: regenerates hypothesis id from
 * (seed, id), returns T_out / T_refit / mask / stats as lr_ransac_rigid does.
 * `key` [host].  Synchronises `stream`. */
int lr_ransac_finalize(const float *src, const float *tgt, int64_t n, const LrRansacParams *params, uint64_t key,
                       double *T_out, double *T_refit, uint8_t *mask, LrRansacStats *stats, void *stream);

/* ---- hypothesis sharding with a library-owned communicator (SURVEY 8(b) lr_comm_init, 8(e)) ----------------
 * One process per GPU of one node.  Every rank owns a small mailbox in device memory that its peers write into
 * directly over NVLink / NVSwitch (cudaIpc-mapped peer memory): the 8-byte packed (count, id) key of a round is
 * exchanged inside the kernel that ends the round, so a sharded run has no host round trip and no separate
 * collective launch.  Set-up: every rank calls lr_comm_init (allocates the mailbox, returns its 64-byte IPC
 * handle in handle_out [host]), the caller all-gathers the handles (torch.distributed, MPI, a file, ...) and
 * every rank calls lr_comm_connect with all `world` handles in rank order (all_handles [host], world x 64 B;
 * may be null when world == 1).  At most 16 ranks.  lr_comm_info: *world = 0 when no communicator is connected. */
int lr_comm_init(int rank, int world, void *handle_out);
int lr_comm_connect(const void *all_handles);
int lr_comm_info(int *rank, int *world);
int lr_comm_destroy(void);

/* lr_ransac_rigid with the hypotheses of every round split across the ranks of the communicator (a COLLECTIVE:
 * every rank calls it with the same correspondences and parameters).  Rank g generates and scores the g-th
 * contiguous slice of each round; the winner is the arg-max over all ranks of the packed key (ties -> lowest id),
 * so the outputs are those of the single-GPU call for any number of ranks; every rank regenerates the selected
 * model from its id, computes mask / refit itself and returns the same values.  stats->n_scored counts all
 * ranks.  LR_SCORE_COUNT only.  An unanswered exchange (a rank missing) times out after ~3 s with LR_ERR_CUDA. */
int lr_ransac_rigid_sharded(const float *src, const float *tgt, int64_t n, const LrRansacParams *params,
                            double *T_out, double *T_refit, uint8_t *mask, LrRansacStats *stats, void *stream);

/* Error-bound probe of the tensor-core sweep (parity tests / micro-benchmark): models[H,12] (device, fp64 rows
 * [R|t]) are installed as the surviving hypotheses of one round over the correspondences src/tgt[n,3].
 * d_out [H, pad128(n), 3] fp32 (device, nullable) = the three residual components as the tensor cores deliver
 * them; E_out[H] (device, nullable) = the bound |d_tc - d_fp64| <= E the sweep's band is derived from;
 * counts_out[H] int32 (device, nullable) = exact inlier counts of the models through the production sweep.
 * pad128(n) = n rounded up to a multiple of 128.  H <= 2^20.  Synchronises `stream`. */
int lr_ransac_tc_probe(const float *src, const float *tgt, int64_t n, const double *models, int64_t H,
                       double threshold, float *d_out, double *E_out, int32_t *counts_out, void *stream);

/* Confidence stopping rule shared by every rank: hypotheses needed once the
 * best inlier count is c (Open3D: log(1-conf)/log(1-(c/n)^m)). */
int64_t lr_ransac_conf_iters(int64_t c, int64_t n, int m, double confidence, int64_t max_iters);

/* The hypothesis sampler itself, for parity tests: ids[H] int64 (device) ->
 * samples[H,m] int32 (device). */
int lr_ransac_sample(const LrRansacParams *params, int64_t n, int64_t id_lo, int64_t H, int32_t *samples,
                     void *stream);

/* Refit over an indexed correspondence set (FR.py:99-111): inliers of
 * (xyz0[i0[k]], xyz1[i1[k]]) under T_in [host] at `threshold`, then Kabsch.
 * T_out[16] [host]; *count [host, nullable].  Synchronises `stream`. */
int lr_refit_indexed(const float *xyz0, const float *xyz1, const int64_t *i0, const int64_t *i1, int64_t K,
                     const double *T_in, double threshold, double *T_out, int64_t *count, void *stream);

/* ---- ICP refinement, SURVEY 8(f4) (Experiments/test.py:183-188: open3d registration_icp,
 * point-to-point, 0.6 m) composed from the path's own kernels ------------------------------ */

/* out[n,8] = (fp32(T_in * xyz[i]), 0, 0, 0, 0, 0): rows for a 3-D nearest-neighbour search with
 * lr_match_nn(..., D = 8).  T_in[16] [host]. */
int lr_transform_pad8(const float *xyz, int64_t n, const double *T_in, float *out, void *stream);

/* One ICP update: over the pairs (xyz0[i0[k]], xyz1[i1[k]]) (i0 nullable = identity) keep those with
 * |T_in p - q| < threshold, return their count, the sum of their squared residuals under T_in and
 * the Kabsch transform of the kept pairs.  [host] outputs.  Synchronises `stream`. */
int lr_icp_step(const float *xyz0, const float *xyz1, const int64_t *i0, const int64_t *i1, int64_t K,
                const double *T_in, double threshold, double *T_out, int64_t *count, double *err2, void *stream);

/* The whole refinement on the device (replaces the Open3D call of Experiments/test.py:183-188,
 * registration_icp(src, tgt, max_dist, T_init, TransformationEstimationPointToPoint()), Open3D defaults
 * max_iteration 30, relative_fitness = relative_rmse = 1e-6): the target is binned once into a hashed uniform grid of
 * max_dist-sized cells; every iteration is ONE kernel (transform, nearest target inside max_dist -- strict, ties ->
 * lowest index --, sums, Kabsch, stopping rule) and all iterations are enqueued up front.
 * src [n,3], tgt [m,3] fp32 [device]; T_init[16] [host, nullable = identity]; T_out[16], *fitness (= pairs / n),
 * *inlier_rmse, *iterations [host, nullable except T_out].  Synchronises `stream`. */
int lr_icp_refine(const float *src, int64_t n, const float *tgt, int64_t m, double max_dist, const double *T_init,
                  int max_iteration, double rel_fitness, double rel_rmse, double *T_out, double *fitness,
                  double *inlier_rmse, int *iterations, void *stream);

/* The search of one ICP iteration on its own (parity hook): idx[i] [device, int64] = nearest row of tgt to
 * T_in * src[i] with squared distance < radius^2 (-1 = none), d2[i] [device, nullable] that squared distance. */
int lr_nn3d_radius(const float *src, int64_t n, const float *tgt, int64_t m, const double *T_in, double radius,
                   int64_t *idx, double *d2, void *stream);

/* ---- PointDSC seed scoring, SURVEY 8(f4) (Experiments/models/PointDSC.py:293-336) --------------------- */

/* rigid_transform_3d over S neighbourhoods (Experiments/models/common.py:7-45, the call of PointDSC.py:318):
 * A, B [S,k,3] fp32, w [S,k] fp32 (nullable = ones) -> T_out [S,16] fp64 (4x4 row-major, column-vector
 * convention), all [device].  Centroids divide by (sum w + 1e-6) as the reference does.  Asynchronous. */
int lr_kabsch_weighted_batch(const float *A, const float *B, const float *w, int64_t S, int k, double *T_out,
                             void *stream);

/* PointDSC.py:319-331: every seed transform (models [S,16] fp64 [device]) against all n correspondences on the
 * tensor-core inlier sweep: counts[S] [device, nullable] = #(|T p - q| < threshold) (fitness = counts / n),
 * *best [host] = arg-max (first maximum), T_best[16] [host] = its transform, labels[n] [device, nullable] =
 * its inlier mask, T_refit[16] [host, nullable] = Kabsch over those inliers.  Synchronises `stream`. */
int lr_seeds_score(const float *src, const float *tgt, int64_t n, const double *models, int64_t S, double threshold,
                   int32_t *counts, uint8_t *labels, int64_t *best, int64_t *best_count, double *T_best,
                   double *T_refit, void *stream);

/* ---- measurement switches (tools/, never needed by a caller) ------------------------------------------ */

/* lr_ransac_rigid then generates and scores only rank's contiguous slice of every round, with no exchange: what ONE
 * rank of a hypothesis-sharded run executes, measurable on a single GPU (tools/pair_breakdown.py).  (0, 1) = default. */
int lr_debug_slice(int rank, int world);

/* 0: launch the kernels of a run the ordinary, fully stream-ordered way instead of with programmatic dependent launch
 * (tools/pdl_ab.py).  Results are identical either way. */
int lr_debug_pdl(int on);

#ifdef __cplusplus
}
#endif
#endif /* LIDARREG_H */
