// sparse.cu -- brute-force search over SPARSE vectors with the Dot metric: what logics.item_to_item / user_to_user of the
// types "tags", "users" and "auto" ask their vector store for (storage/vectors/xvec.go:244-248 flat sparse index, :405
// query; vectors built by appendSparseVector, logics/vector_writer.go:200-209).  SURVEY 8f-1, the row after the dense index.
//
// First hardware run in round 2 (tests/test_sparse_gpu.py green).  The arithmetic it must reproduce: merge-join dot of two
// sparse vectors summed in ascending index order, fp32, multiply and add unfused (the test suite's CPU restatement).
//
// Layout: the vectors as CSR (row -> ascending feature indices + values) and the same entries as CSC (feature -> ascending
// rows + values, the posting lists).  One warp answers one query row r:
//   1. for every feature f of r in ascending order: acc[c] += v_rf * v_cf over the posting list of f (the lanes take distinct
//      rows c, a __syncwarp orders the features) -- every candidate's dot is summed in ascending feature order, i.e. exactly
//      the merge-join order of the oracle, bit for bit;
//   2. one coalesced pass over the warp's private accumulator row acc[0..N): keep the k largest positive dots with the
//      warp-cooperative sorted list (order: dot descending, then row ascending) and reset the accumulator to zero.
// Only rows that share a feature with the query can have a non-zero dot; rows with dot <= 0 are never returned (the
// reference drops them right after the query, logics/item_to_item.go:73).  Roofline class: HBM/L2 (N floats scanned per query).
#include <algorithm>

#include "sparse.cuh"

using namespace gb;

namespace {

// (dist, idx) ordering of topk.cu: ascending distance (= descending dot), then ascending index
__device__ __forceinline__ bool before(float da, int32_t ia, float db, int32_t ib) { return da < db || (da == db && ia < ib); }

// topk.cu list_insert_warp: sorted insertion done by the whole warp; all arguments warp-uniform, k <= 128
__device__ __forceinline__ void list_insert_warp(float *ld, int32_t *li, int &len, int k, float dv, int32_t iv, int lane)
{
    if (len == k && !before(dv, iv, ld[k - 1], li[k - 1])) return;
    const int n = len;
    float od[4];
    int32_t oi[4];
    int pos = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int e = lane + 32 * r;
        const bool has = e < n;
        od[r] = has ? ld[e] : 0.f;
        oi[r] = has ? li[e] : 0;
        pos += __popc(__ballot_sync(0xffffffffu, has && before(od[r], oi[r], dv, iv)));
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int e = lane + 32 * r;
        if (e < n && e >= pos && e + 1 < k) { ld[e + 1] = od[r]; li[e + 1] = oi[r]; }
    }
    if (lane == 0) { ld[pos] = dv; li[pos] = iv; }
    if (len < k) len++;
    __syncwarp();
}

__global__ void __launch_bounds__(128)
sparse_search_kernel(const int64_t *off, const uint32_t *ind, const float *val, const int64_t *foff, const int32_t *frow,
                     const float *fval, int64_t N, uint32_t n_features, int64_t q0, int64_t nq, int k, float *acc, int32_t *out_idx,
                     float *out_dot, int32_t *out_count, const int64_t *q_off, const uint32_t *q_ind, const float *q_val,
                     const uint8_t *allow)
{
    // queries: stored rows q0 + q (never returned themselves), or -- q_off != nullptr -- rows of an external CSR
    // (vecdb.cu: QueryVectors passes the vector, not its id); allow: optional per-row filter
    extern __shared__ unsigned char sm_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float *ld = reinterpret_cast<float *>(sm_raw) + (size_t)wid * k;
    int32_t *li = reinterpret_cast<int32_t *>(reinterpret_cast<float *>(sm_raw) + (size_t)nw * k) + (size_t)wid * k;
    float *a = acc + ((int64_t)blockIdx.x * nw + wid) * N;   // this warp's accumulator row, all zero between queries
    for (int64_t q = (int64_t)blockIdx.x * nw + wid; q < nq; q += (int64_t)gridDim.x * nw) {
        const int64_t row = q_off ? -1 : q0 + q;
        const int64_t tb = q_off ? q_off[q] : off[row], te = q_off ? q_off[q + 1] : off[row + 1];
        const uint32_t *qi = q_off ? q_ind : ind;
        const float *qv = q_off ? q_val : val;
        for (int64_t t = tb; t < te; t++) {
            const uint32_t f = qi[t];
            const float v = qv[t];
            if (f >= n_features) continue;   // a feature no stored vector has (warp-uniform)
            for (int64_t p = foff[f] + lane; p < foff[f + 1]; p += 32) {
                const int32_t c = frow[p];
                a[c] = a[c] + v * fval[p];        // unfused multiply and add, like the oracle (-fmad=false)
            }
            __syncwarp();
        }
        int len = 0;
        for (int64_t base = 0; base < N; base += 32) {
            const int64_t c = base + lane;
            float x = 0.f;
            if (c < N) {
                x = a[c];
                if (x != 0.f) a[c] = 0.f;
            }
            const bool valid = c < N && c != row && x > 0.f && (allow == nullptr || allow[c] != 0);
            // cheap pre-filter against the current k-th best before the warp-wide insertion
            const bool cand = valid && (len < k || before(-x, (int32_t)c, ld[k - 1], li[k - 1]));
            unsigned m = __ballot_sync(0xffffffffu, cand);
            while (m) {
                const int s = __ffs(m) - 1;
                m &= m - 1;
                const float dd = __shfl_sync(0xffffffffu, -x, s);
                const int32_t vv = (int32_t)(base + s);
                list_insert_warp(ld, li, len, k, dd, vv, lane);
            }
        }
        __syncwarp();
        for (int e = lane; e < k; e += 32) {
            const bool has = e < len;
            out_idx[q * k + e] = has ? li[e] : -1;
            out_dot[q * k + e] = has ? -ld[e] : 0.f;
        }
        if (lane == 0) out_count[q] = len;
        __syncwarp();
    }
}

int32_t sync_device(gorse_b200_sparse_index *ix)
{
    if (!ix->dirty) return GORSE_B200_OK;
    gorse_b200_ctx *c = ix->ctx;
    const int64_t n = (int64_t)ix->h_off.size() - 1, nnz = (int64_t)ix->h_ind.size();
    // CSC by counting sort: postings of a feature in ascending row order
    std::vector<int64_t> foff((size_t)ix->n_features + 1, 0);
    for (int64_t t = 0; t < nnz; t++) foff[(size_t)ix->h_ind[(size_t)t] + 1]++;
    for (size_t f = 0; f < ix->n_features; f++) foff[f + 1] += foff[f];
    std::vector<int64_t> cur(foff.begin(), foff.end() - 1);
    std::vector<int32_t> frow((size_t)nnz);
    std::vector<float> fval((size_t)nnz);
    for (int64_t r = 0; r < n; r++)
        for (int64_t t = ix->h_off[(size_t)r]; t < ix->h_off[(size_t)r + 1]; t++) {
            const int64_t p = cur[ix->h_ind[(size_t)t]]++;
            frow[(size_t)p] = (int32_t)r;
            fval[(size_t)p] = ix->h_val[(size_t)t];
        }
    GB_TRY(ix->off.alloc((size_t)n + 1));
    GB_TRY(ix->foff.alloc(foff.size()));
    GB_TRY(ix->ind.alloc((size_t)nnz));
    GB_TRY(ix->val.alloc((size_t)nnz));
    GB_TRY(ix->frow.alloc((size_t)nnz));
    GB_TRY(ix->fval.alloc((size_t)nnz));
    cudaStream_t s = c->stream;
    GB_CUDA(cudaMemcpyAsync(ix->off.p, ix->h_off.data(), sizeof(int64_t) * ((size_t)n + 1), cudaMemcpyHostToDevice, s));
    GB_CUDA(cudaMemcpyAsync(ix->foff.p, foff.data(), sizeof(int64_t) * foff.size(), cudaMemcpyHostToDevice, s));
    if (nnz) {
        GB_CUDA(cudaMemcpyAsync(ix->ind.p, ix->h_ind.data(), sizeof(uint32_t) * (size_t)nnz, cudaMemcpyHostToDevice, s));
        GB_CUDA(cudaMemcpyAsync(ix->val.p, ix->h_val.data(), sizeof(float) * (size_t)nnz, cudaMemcpyHostToDevice, s));
        GB_CUDA(cudaMemcpyAsync(ix->frow.p, frow.data(), sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, s));
        GB_CUDA(cudaMemcpyAsync(ix->fval.p, fval.data(), sizeof(float) * (size_t)nnz, cudaMemcpyHostToDevice, s));
    }
    GB_CUDA(cudaStreamSynchronize(s));   // the staging vectors die here
    ix->acc.free();                      // sized for the old N
    ix->acc_slots = 0;
    ix->dirty = false;
    return GORSE_B200_OK;
}

}  // namespace

extern "C" {

int32_t gorse_b200_sparse_index_create(gorse_b200_ctx *ctx, gorse_b200_sparse_index **out)
{
    GB_CHECK_ARG(ctx != nullptr && out != nullptr, "NULL argument");
    gorse_b200_sparse_index *ix = new gorse_b200_sparse_index();
    ix->ctx = ctx;
    *out = ix;
    return GORSE_B200_OK;
}

int32_t gorse_b200_sparse_index_destroy(gorse_b200_sparse_index *ix)
{
    if (!ix) return GORSE_B200_OK;
    {
        ScopedDevice sd(ix->ctx->device);
        cudaStreamSynchronize(ix->ctx->stream);
        ix->off.free(); ix->foff.free(); ix->ind.free(); ix->val.free(); ix->frow.free(); ix->fval.free(); ix->acc.free();
    }
    delete ix;
    return GORSE_B200_OK;
}

int32_t gorse_b200_sparse_index_add(gorse_b200_sparse_index *ix, const int64_t *off, const uint32_t *indices, const float *values,
                                    int64_t n, int64_t *count_out)
{
    GB_CHECK_ARG(ix != nullptr, "index is NULL");
    GB_CHECK_ARG(n >= 0, "negative n");
    std::lock_guard<std::mutex> lk(ix->mu);
    if (n > 0) {
        GB_CHECK_ARG(off != nullptr, "off is NULL");
        GB_CHECK_ARG(off[0] == 0, "off[0] must be 0");
        for (int64_t r = 0; r < n; r++) {
            GB_CHECK_ARG(off[r + 1] >= off[r], "offsets must not decrease (row %lld)", (long long)r);
            for (int64_t t = off[r]; t < off[r + 1]; t++) {
                GB_CHECK_ARG(indices != nullptr && values != nullptr, "NULL indices/values");
                GB_CHECK_ARG(t == off[r] || indices[t] > indices[t - 1], "indices of row %lld are not strictly ascending", (long long)r);
                GB_CHECK_ARG(values[t] == values[t], "NaN value in row %lld", (long long)r);
                GB_CHECK_ARG(indices[t] < 0x7fffffffu, "feature index too large");
            }
        }
        if ((int64_t)ix->h_off.size() - 1 + n >= (1ll << 31)) { set_error("more than 2^31 vectors"); return GORSE_B200_ERR_UNSUPPORTED; }
        const int64_t base = ix->h_off.back();
        for (int64_t r = 0; r < n; r++) ix->h_off.push_back(base + off[r + 1]);
        ix->h_ind.insert(ix->h_ind.end(), indices, indices + off[n]);
        ix->h_val.insert(ix->h_val.end(), values, values + off[n]);
        for (int64_t t = 0; t < off[n]; t++) ix->n_features = std::max(ix->n_features, indices[t] + 1);
        ix->dirty = true;
    }
    if (count_out) *count_out = (int64_t)ix->h_off.size() - 1;
    return GORSE_B200_OK;
}

int32_t gorse_b200_sparse_index_len(const gorse_b200_sparse_index *ix, int64_t *count_out)
{
    GB_CHECK_ARG(ix != nullptr && count_out != nullptr, "NULL argument");
    *count_out = (int64_t)ix->h_off.size() - 1;
    return GORSE_B200_OK;
}

int32_t gorse_b200_sparse_index_search_range(gorse_b200_sparse_index *ix, int64_t q0, int64_t q1, int32_t k, int32_t *idx_out,
                                             float *dot_out, int32_t *count_out)
{
    GB_CHECK_ARG(ix != nullptr, "index is NULL");
    GB_CHECK_ARG(k > 0 && k <= 128, "k must be in 1..128 (got %d)", k);
    std::lock_guard<std::mutex> lk(ix->mu);
    const int64_t n = (int64_t)ix->h_off.size() - 1;
    if (q0 < 0 || q1 < q0 || q1 > n) {
        set_error("index out of range: [%lld, %lld)", (long long)q0, (long long)q1);
        return GORSE_B200_ERR_RANGE;
    }
    const int64_t nq = q1 - q0;
    if (nq == 0) return GORSE_B200_OK;
    GB_CHECK_ARG(idx_out != nullptr && dot_out != nullptr && count_out != nullptr, "NULL output");
    return gb::sparse_search_host(ix, q0, nq, nullptr, nullptr, nullptr, nullptr, k, idx_out, dot_out, count_out);
}

}  // extern "C"

namespace gb {

// queries = stored rows [q0, q0 + nq) (q_off == nullptr) or nq rows of a HOST CSR (q_off/q_ind/q_val); allow = optional DEVICE
// per-row filter; results to host buffers.  The caller holds ix->mu.
int32_t sparse_search_host(gorse_b200_sparse_index *ix, int64_t q0, int64_t nq, const int64_t *q_off, const uint32_t *q_ind,
                           const float *q_val, const uint8_t *d_allow, int32_t k, int32_t *idx_out, float *dot_out, int32_t *count_out)
{
    const int64_t n = (int64_t)ix->h_off.size() - 1;
    gorse_b200_ctx *c = ix->ctx;
    ScopedDevice sd(c->device);
    GB_TRY(sync_device(ix));
    if (n == 0) {
        for (int64_t q = 0; q < nq; q++) count_out[q] = 0;
        return GORSE_B200_OK;
    }
    const int warps = 4;
    // one accumulator row per resident warp, at most ~1 GB of them
    int64_t slots = std::min<int64_t>((nq + warps - 1) / warps * warps, (int64_t)c->sm_count * 4 * warps);
    slots = std::max<int64_t>(warps, std::min<int64_t>(slots, ((1ll << 30) / 4 / std::max<int64_t>(1, n)) / warps * warps));
    if (ix->acc_slots < slots) {
        GB_TRY(ix->acc.alloc((size_t)slots * (size_t)n));
        GB_CUDA(cudaMemsetAsync(ix->acc.p, 0, sizeof(float) * (size_t)slots * (size_t)n, c->stream));
        ix->acc_slots = slots;
    }
    DevBuf<int32_t> d_idx, d_cnt;
    DevBuf<float> d_dot, dq_val;
    DevBuf<int64_t> dq_off;
    DevBuf<uint32_t> dq_ind;
    int32_t st = GORSE_B200_OK;
    auto done = [&](int32_t s) { cudaStreamSynchronize(c->stream); d_idx.free(); d_cnt.free(); d_dot.free(); dq_val.free(); dq_off.free(); dq_ind.free(); return s; };
    if ((st = d_idx.alloc((size_t)nq * k)) || (st = d_dot.alloc((size_t)nq * k)) || (st = d_cnt.alloc((size_t)nq))) return done(st);
    cudaError_t e = cudaSuccess;
    if (q_off) {
        const int64_t qn = q_off[nq];
        if ((st = dq_off.alloc((size_t)nq + 1)) || (st = dq_ind.alloc((size_t)qn)) || (st = dq_val.alloc((size_t)qn))) return done(st);
        e = cudaMemcpyAsync(dq_off.p, q_off, sizeof(int64_t) * ((size_t)nq + 1), cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess && qn) e = cudaMemcpyAsync(dq_ind.p, q_ind, sizeof(uint32_t) * (size_t)qn, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess && qn) e = cudaMemcpyAsync(dq_val.p, q_val, sizeof(float) * (size_t)qn, cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) { set_error("sparse search upload: %s", cudaGetErrorString(e)); return done(GORSE_B200_ERR_CUDA); }
    }
    const int grid = (int)(slots / warps);
    const size_t sm = (size_t)warps * 2 * (size_t)k * 4;
    sparse_search_kernel<<<grid, 32 * warps, sm, c->stream>>>(ix->off.p, ix->ind.p, ix->val.p, ix->foff.p, ix->frow.p, ix->fval.p, n,
                                                             ix->n_features, q0, nq, k, ix->acc.p, d_idx.p, d_dot.p, d_cnt.p,
                                                             q_off ? dq_off.p : nullptr, dq_ind.p, dq_val.p, d_allow);
    c->launches++;
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(idx_out, d_idx.p, sizeof(int32_t) * (size_t)nq * k, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dot_out, d_dot.p, sizeof(float) * (size_t)nq * k, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(count_out, d_cnt.p, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { set_error("sparse search: %s", cudaGetErrorString(e)); return done(GORSE_B200_ERR_CUDA); }
    return done(GORSE_B200_OK);
}

}  // namespace gb
