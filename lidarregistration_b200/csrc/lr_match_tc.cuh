// lr_match_tc.cuh -- interface of the tensor-core matching sweep (lr_match_tc.cu)
#pragma once
#include "lr_common.cuh"

namespace lr_tc {

struct Params;
struct Prepared {
    uint4 *op0, *op1;   // fp16 operand images of f0 / f1
    float *n0, *n1;     // canonical squared norms
    Params *params;
    int2 *cand;         // events: (32-column chunk id, chunk maximum bits)
    int *cand_cnt, *ovf_rows;
    char *partial;      // partial top-2 lists of the overflow scan
    int64_t partial_cap;
    int64_t pad0, pad1;
};

size_t scratch_bytes(int64_t N, int64_t M);
int prepare(const float *f0, int64_t N, const float *f1, int64_t M, char *scratch, Prepared &P, cudaStream_t st);
// swap = false: neighbours of f0's rows in f1; swap = true: neighbours of f1's rows in f0
// acc16: fp16 accumulators in TMEM (packed tcgen05.ld, half the epilogue work, wider candidate band)
int sweep(const Prepared &P, bool swap, bool acc16, const float *f0, int64_t N, const float *f1, int64_t M,
          int64_t *idx1, int64_t *idx2, cudaStream_t st);

}  // namespace lr_tc
