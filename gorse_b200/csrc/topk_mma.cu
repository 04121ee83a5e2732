// topk_mma.cu -- stage 1 of the brute-force search on the tensor cores (tcgen05 + TMA), see topk.cu.
#include "topk.cuh"

namespace gb {

bool mma_path_eligible(const gorse_b200_index *, int64_t, int) { return false; }

int32_t search_mma(gorse_b200_index *, const float *, const int64_t *, int64_t, int64_t, int, int, int32_t *, float *, int32_t *, int *)
{
    set_error("tensor-core search path not built");
    return GORSE_B200_ERR_UNSUPPORTED;
}

}  // namespace gb
