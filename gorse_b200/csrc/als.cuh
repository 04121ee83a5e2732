// als.cuh -- pieces of the eALS epoch that live in their own translation units.
#pragma once
#include <algorithm>

#include "cf.cuh"

namespace gb {

// als_thread.cu: one thread per row, rows with at most nmax (<= 16) entries, d in {32, 64, 96, 128}
int32_t als_thread_rows(gorse_b200_ctx *c, int d, int nmax, float *X, const float *Y, const int64_t *off, const int32_t *idx,
                        const float *S, float reg, float w, const int32_t *rows, int32_t n_rows);

// als_gram_tc.cu: partial (G, h) of every chunk of the long rows on the tensor cores (3xTF32), d = 128 only
int32_t als_chunk_gram_tc(gorse_b200_ctx *c, const float *Y, const int32_t *idx, const int64_t *chunk_begin, const int32_t *chunk_len,
                          int32_t n_chunks, float *partial);

}  // namespace gb
