// logics.cu -- host side of SURVEY 8(a) row J: what logics.item_to_item / user_to_user do around the vector
// store: vector construction (bf16-truncated dense embeddings, sqrt(idf) sparse vectors) and the score
// post-processing of QueryItemToItem / QueryUserToUser.  Plain host code; the neighbour search itself is
// gorse_b200_index_search_range (topk.cu).
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "topk.cuh"

using namespace gb;

extern "C" {

// bfloats.FromFloat32 + bfloats.ToFloat32 (common/bfloats/bfloats.go:23-37): keep the high 16 bits.
// A matrix stored this way is exactly representable in bf16, i.e. the index's tensor-core mirror of it is lossless.
int32_t gorse_b200_bf16_truncate(const float *in, int64_t n, float *out)
{
    GB_CHECK_ARG(n >= 0, "negative n");
    if (n == 0) return GORSE_B200_OK;
    GB_CHECK_ARG(in != nullptr && out != nullptr, "NULL argument");
    for (int64_t i = 0; i < n; i++) {
        uint32_t u;
        std::memcpy(&u, &in[i], 4);
        u &= 0xffff0000u;
        std::memcpy(&out[i], &u, 4);
    }
    return GORSE_B200_OK;
}

// appendSparseVector (logics/vector_writer.go:200-209)
int32_t gorse_b200_sparse_vector(const int32_t *ids, int32_t n_ids, const float *idf, int32_t n_idf, uint32_t offset,
                                 uint32_t *indices_out, float *values_out, int32_t *count_out)
{
    GB_CHECK_ARG(n_ids >= 0 && n_idf >= 0, "negative length");
    GB_CHECK_ARG(count_out != nullptr, "count_out is NULL");
    GB_CHECK_ARG(n_ids == 0 || (ids != nullptr && indices_out != nullptr && values_out != nullptr), "NULL argument");
    GB_CHECK_ARG(n_idf == 0 || idf != nullptr, "idf is NULL");
    int32_t m = 0;
    for (int32_t t = 0; t < n_ids; t++) {
        const int32_t id = ids[t];
        if (id < 0 || id >= n_idf || idf[id] <= 0) continue;
        indices_out[m] = offset + (uint32_t)id;
        values_out[m] = (float)std::sqrt((double)idf[id]);
        m++;
    }
    *count_out = m;
    return GORSE_B200_OK;
}

// QueryItemToItem (logics/item_to_item.go:63-85) on neighbours in THIS library's convention: ascending distance,
// distance = -dot for GORSE_B200_METRIC_NEG_DOT, so the vector store's "higher is closer" score is -distance
// for both metrics (storage/vectors/database.go:101).
int32_t gorse_b200_similar_scores(int32_t metric, double score_scale, int32_t self_id, int32_t n, const int32_t *nbr_ids,
                                  const float *nbr_dist, int32_t n_nbr, int32_t *ids_out, double *scores_out, int32_t *count_out)
{
    GB_CHECK_ARG(metric == GORSE_B200_METRIC_EUCLIDEAN || metric == GORSE_B200_METRIC_NEG_DOT, "bad metric %d", metric);
    GB_CHECK_ARG(n >= 0 && n_nbr >= 0, "negative length");
    GB_CHECK_ARG(count_out != nullptr, "count_out is NULL");
    GB_CHECK_ARG(n_nbr == 0 || (nbr_ids != nullptr && nbr_dist != nullptr), "NULL neighbours");
    GB_CHECK_ARG(n == 0 || (ids_out != nullptr && scores_out != nullptr), "NULL output");
    const bool dot = metric == GORSE_B200_METRIC_NEG_DOT;
    int32_t m = 0;
    for (int32_t t = 0; t < n_nbr && m < n; t++) {
        if (nbr_ids[t] < 0) break;  // padding of a short result row
        const float s = -nbr_dist[t];
        if (nbr_ids[t] == self_id || (dot && s <= 0)) continue;
        double score = (double)s * score_scale;
        if (!dot) score = 1 / (1 - score);
        ids_out[m] = nbr_ids[t];
        scores_out[m] = score;
        m++;
    }
    *count_out = m;
    return GORSE_B200_OK;
}

// QueryItemToItem / QueryUserToUser for stored vectors [q0, q1): ids_out/scores_out are (q1-q0) x n, rows padded
// with id -1.  The reference asks the store for n+1 neighbours of the vector and drops its own id; asking for n
// neighbours that exclude the query (SearchIndex, bruteforce.go:47) is the same set in the same order.
int32_t gorse_b200_index_query_similar(gorse_b200_index *ix, int64_t q0, int64_t q1, int32_t n, double score_scale,
                                       int32_t *ids_out, double *scores_out, int32_t *count_out)
{
    GB_CHECK_ARG(ix != nullptr, "index is NULL");
    GB_CHECK_ARG(n > 0, "n must be positive");
    GB_CHECK_ARG(q1 >= q0, "empty range");
    const int64_t nq = q1 - q0;
    if (nq == 0) return GORSE_B200_OK;
    GB_CHECK_ARG(ids_out != nullptr && scores_out != nullptr && count_out != nullptr, "NULL output");
    const int32_t metric = ix->metric;
    std::vector<int32_t> idx((size_t)nq * n), cnt((size_t)nq);
    std::vector<float> dist((size_t)nq * n);
    GB_TRY(gorse_b200_index_search_range(ix, q0, q1, n, 0, idx.data(), dist.data(), cnt.data()));
    for (int64_t r = 0; r < nq; r++) {
        int32_t m = 0;
        GB_TRY(gorse_b200_similar_scores(metric, score_scale, -1, n, idx.data() + r * n, dist.data() + r * n, cnt[(size_t)r],
                                         ids_out + r * n, scores_out + r * n, &m));
        for (int32_t t = m; t < n; t++) { ids_out[r * n + t] = -1; scores_out[r * n + t] = 0; }
        count_out[r] = m;
    }
    return GORSE_B200_OK;
}

// ---- model hand-off: cf.BaseMatrixFactorization.Marshal's factor block from the flat mirror ----------------
// model/cf/model.go:212-245 writes, for users and then items,  int64 LE count of predictable rows  followed by one
// pbutil.WriteDelimited(protocol.LatentFactor{Id, Data}) per predictable row (protocol/encoding.proto:27-30:
// string id = 1; repeated float data = 2 -> packed): a varint message length, then  0x0A len id-bytes  (omitted for an empty
// id, proto3)  and  0x12 varint(4 d) d little-endian floats  (omitted for d = 0).  The reference builds one protobuf object
// and one slice header per row; this writes the same bytes in one pass over the flat [rows x d] table.
static size_t put_varint(uint8_t *out, uint64_t v)
{
    size_t n = 0;
    while (v >= 0x80) { if (out) out[n] = (uint8_t)(v | 0x80); n++; v >>= 7; }
    if (out) out[n] = (uint8_t)v;
    return n + 1;
}
static size_t varint_len(uint64_t v) { return put_varint(nullptr, v); }

int32_t gorse_b200_marshal_latent_factors(const float *factors, int32_t rows, int32_t d, const uint8_t *predictable, const char *const *ids,
                                          uint8_t *out, size_t cap, size_t *len_out)
{
    GB_CHECK_ARG(len_out != nullptr, "len_out is NULL");
    GB_CHECK_ARG(rows >= 0 && d >= 0, "negative size");
    GB_CHECK_ARG(rows == 0 || (factors != nullptr && ids != nullptr), "NULL factors/ids");
    // pass 1: size;  pass 2 (only when the buffer is large enough): bytes
    int64_t count = 0;
    size_t need = 8;
    for (int32_t r = 0; r < rows; r++) {
        if (predictable && !predictable[r]) continue;
        GB_CHECK_ARG(ids[r] != nullptr, "ids[%d] is NULL", r);
        const size_t idl = strlen(ids[r]);
        const size_t body = (idl ? 1 + varint_len(idl) + idl : 0) + (d ? 1 + varint_len(4ull * d) + 4ull * d : 0);
        need += varint_len(body) + body;
        count++;
    }
    *len_out = need;
    if (out == nullptr || cap < need) return out == nullptr ? GORSE_B200_OK : GORSE_B200_ERR_RANGE;
    uint8_t *p = out;
    for (int b = 0; b < 8; b++) *p++ = (uint8_t)((uint64_t)count >> (8 * b));   // binary.Write(w, LittleEndian, int64(count))
    for (int32_t r = 0; r < rows; r++) {
        if (predictable && !predictable[r]) continue;
        const size_t idl = strlen(ids[r]);
        const size_t body = (idl ? 1 + varint_len(idl) + idl : 0) + (d ? 1 + varint_len(4ull * d) + 4ull * d : 0);
        p += put_varint(p, body);
        if (idl) { *p++ = 0x0A; p += put_varint(p, idl); memcpy(p, ids[r], idl); p += idl; }
        if (d) { *p++ = 0x12; p += put_varint(p, 4ull * d); memcpy(p, factors + (int64_t)r * d, 4ull * d); p += 4ull * d; }   // x86 / arm64 are little-endian
    }
    return GORSE_B200_OK;
}

}  // extern "C"
