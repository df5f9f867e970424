// STUB for <opencv2/core.hpp>: the reference's EdgeLenPreemptiveVerification reads cv::Mat::data and cv::Mat::cols
// (preemption_edge_length.h:84-85) and nothing else of OpenCV.  See oracle/ref_stubs/model.h.
#pragma once
namespace cv
{
struct Mat {
    unsigned char *data;
    int rows, cols;
};
}  // namespace cv
