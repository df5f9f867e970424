// ref_elc_wrap.cpp -- TEST INFRASTRUCTURE.  Compiles the REFERENCE's own edge-length pre-verification
//   /root/reference/GC-RANSAC/src/pygcransac/include/preemption/preemption_edge_length.h:71-128
// unmodified and in place (never copied into this repo) behind a C entry, so that the oracle's restatement
// (lro_elc) and through it the CUDA kernel can be checked against the real code.  The header's other includes
// (GCRANSAC "model.h", OpenCV, Eigen) are un-vendored; oracle/ref_stubs/ declares the two members of cv::Mat and the
// two type names the class touches.  The GC-RANSAC engine itself remains unbuildable here.
// Built by `make -C oracle ref` into oracle/_ref/ (git-ignored; travels to the GPU box like the other built .so files).
#include <cstddef>
#include <vector>

#include "preemption/preemption_edge_length.h"

struct DummyEstimator {};

extern "C" __attribute__((visibility("default"))) int ref_elc_verify(const double *points_n_by_6, long n,
                                                                     const size_t *minimal_sample, size_t sample_number)
{
    cv::Mat points;
    points.data = reinterpret_cast<unsigned char *>(const_cast<double *>(points_n_by_6));
    points.rows = (int)n;
    points.cols = 6;  // [x1 y1 z1 x2 y2 z2] per row, as gcransac_python.cpp:426-435 packs them
    DummyEstimator est;
    gcransac::preemption::EdgeLenPreemptiveVerification<DummyEstimator> v(points, est);
    gcransac::Model model;
    gcransac::Score best, score;
    std::vector<size_t> inliers;
    const double threshold = 0.0;
    const size_t iteration = 0;
    return v.verifyModel(model, est, threshold, iteration, best, points, minimal_sample, sample_number, inliers, score) ? 1 : 0;
}
