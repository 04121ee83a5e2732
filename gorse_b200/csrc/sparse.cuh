// sparse.cuh -- state of the sparse Dot index (sparse.cu), shared with the vector-collection layer (vecdb.cu).
#pragma once
#include <mutex>
#include <vector>

#include "common.cuh"

struct gorse_b200_sparse_index {
    gorse_b200_ctx *ctx = nullptr;
    std::mutex mu;
    // host copy (append-friendly); device mirrors are rebuilt lazily after an add
    std::vector<int64_t> h_off{0};
    std::vector<uint32_t> h_ind;
    std::vector<float> h_val;
    uint32_t n_features = 0;
    bool dirty = true;
    gb::DevBuf<int64_t> off, foff;
    gb::DevBuf<uint32_t> ind;
    gb::DevBuf<float> val, fval;
    gb::DevBuf<int32_t> frow;
    gb::DevBuf<float> acc;
    int64_t acc_slots = 0;
};


namespace gb {
int32_t sparse_search_host(gorse_b200_sparse_index *ix, int64_t q0, int64_t nq, const int64_t *q_off, const uint32_t *q_ind,
                           const float *q_val, const uint8_t *d_allow, int32_t k, int32_t *idx_out, float *dot_out, int32_t *count_out);
}
