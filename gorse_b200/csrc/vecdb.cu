// vecdb.cu -- one vector COLLECTION of a vectors.Database on the GPU: what item-to-item, user-to-user and
// CF retrieval query in this snapshot (storage/vectors/database.go:107-120; backend semantics restated from the reference's
// default backend storage/vectors/xvec.go:288-449).
//
//   AddVectors (:288-318, an upsert by id)   gorse_b200_vecdb_add      append; the previous version of an id is tombstoned
//   GetVectors (:320-363)                    gorse_b200_vecdb_get
//   DeleteVectors(timestamp) (:365-371)      gorse_b200_vecdb_delete_before   tombstone ts < timestamp
//   CountVectors (:274-286)                  gorse_b200_vecdb_count
//   QueryVectors(q, categories, topK) (:373-449)  gorse_b200_vecdb_query
//       filter  hidden = false AND categories CONTAIN_ALL (...)  (:381-389)  -> a per-vector allow mask built on the device
//       dense   brute force over the stored fp32 vectors (exact: the reference's DiskANN is approximate), in the
//               summation order of floats.Dot / floats.Euclidean; sparse (dimension 0, Dot only, :244-248) through sparse.cu
//       score   "higher = more similar" (database.go:101): Dot -> the dot product; Euclidean / Cosine -> the NEGATED distance
//               (:425-427); sparse results with score 0 are dropped (:421-423)
// Ids are strings in the reference; the shim keeps the id <-> slot map (go/storage/vectors/b200.go) and the category
// string <-> int32 map, the library works on slots (int64, in insertion order) and category ids.
#include <algorithm>

#include "cf.cuh"
#include "sparse.cuh"
#include "topk.cuh"

extern "C" {
int32_t gorse_b200_sparse_index_create(gorse_b200_ctx *ctx, gorse_b200_sparse_index **out);
int32_t gorse_b200_sparse_index_destroy(gorse_b200_sparse_index *ix);
}

struct gorse_b200_vecdb {
    gorse_b200_ctx *ctx = nullptr;
    int32_t dim = 0, distance = 0;
    gorse_b200_index *dense = nullptr;           // dim > 0
    gorse_b200_sparse_index *sparse = nullptr;   // dim == 0
    std::mutex mu;
    // host metadata (authoritative) + lazily refreshed device copies for the filter kernel
    std::vector<uint8_t> hidden, dead;
    std::vector<int64_t> ts;
    std::vector<int64_t> cat_off{0};
    std::vector<int32_t> cats;
    int64_t live = 0;
    bool meta_dirty = true;
    gb::DevBuf<uint8_t> d_blocked, d_allow;   // blocked = hidden | dead
    gb::DevBuf<int64_t> d_cat_off;
    gb::DevBuf<int32_t> d_cats, d_want;
};

namespace gb {

// allow[i] = !blocked[i] && categories(i) contains every wanted category
__global__ void vecdb_allow_kernel(const uint8_t *blocked, const int64_t *cat_off, const int32_t *cats, const int32_t *want, int n_want,
                                   int64_t n, uint8_t *allow)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool ok = blocked[i] == 0;
    for (int w = 0; w < n_want && ok; w++) {
        bool has = false;
        for (int64_t t = cat_off[i]; t < cat_off[i + 1] && !has; t++) has = cats[t] == want[w];
        ok = has;
    }
    allow[i] = ok ? 1 : 0;
}

// flag[item] = 1 for every item with >= 1 training feedback (IsItemPredictable, model/cf/model.go:129-146)
__global__ void vecdb_mark_items_kernel(const int32_t *user_items, int64_t n, int32_t *flag)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t st = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) flag[user_items[i]] = 1;
}
// X[first + s] = Q[items[s]]  (device to device, float4 per thread)
__global__ void vecdb_copy_rows_kernel(const float *Q, int d4, const int32_t *items, int64_t n, float *X)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = n * d4, st = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += st) {
        const int64_t s = i / d4, c = i - s * d4;
        reinterpret_cast<float4 *>(X)[s * d4 + c] = reinterpret_cast<const float4 *>(Q)[(int64_t)items[s] * d4 + c];
    }
}

static int32_t sync_meta(gorse_b200_vecdb *db)
{
    const int64_t n = (int64_t)db->ts.size();
    if (!db->meta_dirty) return GORSE_B200_OK;
    cudaStream_t s = db->ctx->stream;
    std::vector<uint8_t> blocked((size_t)n);
    for (int64_t i = 0; i < n; i++) blocked[(size_t)i] = db->hidden[(size_t)i] | db->dead[(size_t)i];
    GB_TRY(db->d_blocked.alloc((size_t)n));
    GB_TRY(db->d_allow.alloc((size_t)n));
    GB_TRY(db->d_cat_off.alloc((size_t)n + 1));
    GB_TRY(db->d_cats.alloc(db->cats.size()));
    if (n) GB_CUDA(cudaMemcpyAsync(db->d_blocked.p, blocked.data(), (size_t)n, cudaMemcpyHostToDevice, s));
    GB_CUDA(cudaMemcpyAsync(db->d_cat_off.p, db->cat_off.data(), sizeof(int64_t) * ((size_t)n + 1), cudaMemcpyHostToDevice, s));
    if (!db->cats.empty()) GB_CUDA(cudaMemcpyAsync(db->d_cats.p, db->cats.data(), sizeof(int32_t) * db->cats.size(), cudaMemcpyHostToDevice, s));
    GB_CUDA(cudaStreamSynchronize(s));
    db->meta_dirty = false;
    return GORSE_B200_OK;
}

}  // namespace gb

using namespace gb;

extern "C" {

int32_t gorse_b200_vecdb_create(gorse_b200_ctx *ctx, int32_t dim, int32_t distance, gorse_b200_vecdb **out)
{
    GB_CHECK_ARG(ctx != nullptr && out != nullptr, "NULL ctx/out");
    *out = nullptr;
    GB_CHECK_ARG(dim >= 0 && dim <= 16384, "invalid vector dimension %d", dim);
    GB_CHECK_ARG(distance == GORSE_B200_DISTANCE_EUCLIDEAN || distance == GORSE_B200_DISTANCE_DOT || distance == GORSE_B200_DISTANCE_COSINE,
                 "unknown distance %d", distance);
    if (dim == 0 && distance != GORSE_B200_DISTANCE_DOT) {
        set_error("distance method for sparse vector not supported");   // xvec.go:243-245
        return GORSE_B200_ERR_UNSUPPORTED;
    }
    gorse_b200_vecdb *db = new (std::nothrow) gorse_b200_vecdb();
    if (!db) { set_error("host allocation failed"); return GORSE_B200_ERR_OOM; }
    db->ctx = ctx; db->dim = dim; db->distance = distance;
    const int32_t metric = distance == GORSE_B200_DISTANCE_DOT ? GORSE_B200_METRIC_NEG_DOT
                           : distance == GORSE_B200_DISTANCE_EUCLIDEAN ? GORSE_B200_METRIC_EUCLIDEAN : GORSE_B200_METRIC_COSINE;
    int32_t st = dim > 0 ? gorse_b200_index_create(ctx, dim, metric, &db->dense)
                         : gorse_b200_sparse_index_create(ctx, &db->sparse);
    if (st) { delete db; return st; }
    *out = db;
    return GORSE_B200_OK;
}

int32_t gorse_b200_vecdb_destroy(gorse_b200_vecdb *db)
{
    if (!db) return GORSE_B200_OK;
    {
        ScopedDevice sd(db->ctx->device);
        cudaStreamSynchronize(db->ctx->stream);
        db->d_blocked.free(); db->d_allow.free(); db->d_cat_off.free(); db->d_cats.free(); db->d_want.free();
    }
    gorse_b200_index_destroy(db->dense);
    gorse_b200_sparse_index_destroy(db->sparse);
    delete db;
    return GORSE_B200_OK;
}

int32_t gorse_b200_vecdb_count(gorse_b200_vecdb *db, int64_t *live_out, int64_t *slots_out)
{
    GB_CHECK_ARG(db != nullptr, "collection is NULL");
    std::lock_guard<std::mutex> lk(db->mu);
    if (live_out) *live_out = db->live;
    if (slots_out) *slots_out = (int64_t)db->ts.size();
    return GORSE_B200_OK;
}

int32_t gorse_b200_vecdb_add(gorse_b200_vecdb *db, int64_t n, const float *values, const int64_t *sp_off, const uint32_t *sp_indices,
                             const uint8_t *hidden, const int64_t *timestamp_ms, const int64_t *cat_off, const int32_t *cats,
                             const int64_t *replace, int64_t *first_slot_out)
{
    GB_CHECK_ARG(db != nullptr, "collection is NULL");
    GB_CHECK_ARG(n >= 0, "negative n");
    std::lock_guard<std::mutex> lk(db->mu);
    const int64_t first = (int64_t)db->ts.size();
    if (first_slot_out) *first_slot_out = first;
    if (n == 0) return GORSE_B200_OK;
    GB_CHECK_ARG(values != nullptr || (db->dim == 0 && sp_off && sp_off[n] == 0), "values is NULL");
    GB_CHECK_ARG(db->dim > 0 || sp_off != nullptr, "a sparse collection needs sp_off / sp_indices");
    GB_CHECK_ARG(cat_off == nullptr || cat_off[0] == 0, "cat_off[0] must be 0");
    for (int64_t i = 0; i < n; i++) {
        GB_CHECK_ARG(cat_off == nullptr || cat_off[i + 1] >= cat_off[i], "cat_off not non-decreasing at %lld", (long long)i);
        GB_CHECK_ARG(replace == nullptr || replace[i] < first, "replace[%lld] = %lld is not an existing slot", (long long)i, (long long)replace[i]);
    }
    GB_CHECK_ARG(cat_off == nullptr || cat_off[n] == 0 || cats != nullptr, "cats is NULL");
    int64_t cnt = 0;
    int32_t st = db->dim > 0 ? gorse_b200_index_add(db->dense, values, n, &cnt)
                             : gorse_b200_sparse_index_add(db->sparse, sp_off, sp_indices, values, n, &cnt);
    if (st) return st;
    for (int64_t i = 0; i < n; i++) {
        db->hidden.push_back(hidden ? (hidden[i] != 0) : 0);
        db->dead.push_back(0);
        db->ts.push_back(timestamp_ms ? timestamp_ms[i] : 0);
        if (cat_off) db->cats.insert(db->cats.end(), cats + cat_off[i], cats + cat_off[i + 1]);
        db->cat_off.push_back((int64_t)db->cats.size());
        db->live++;
        if (replace && replace[i] >= 0 && !db->dead[(size_t)replace[i]]) {   // upsert: the old version disappears
            db->dead[(size_t)replace[i]] = 1;
            db->live--;
        }
    }
    db->meta_dirty = true;
    return GORSE_B200_OK;
}

int32_t gorse_b200_vecdb_get(gorse_b200_vecdb *db, const int64_t *slots, int64_t n, float *values_out, uint8_t *hidden_out,
                             int64_t *timestamp_out, uint8_t *live_out)
{
    GB_CHECK_ARG(db != nullptr, "collection is NULL");
    GB_CHECK_ARG(n >= 0 && (n == 0 || slots != nullptr), "bad slots");
    std::lock_guard<std::mutex> lk(db->mu);
    const int64_t total = (int64_t)db->ts.size();
    ScopedDevice sd(db->ctx->device);
    for (int64_t i = 0; i < n; i++) {
        const int64_t s = slots[i];
        const bool ok = s >= 0 && s < total && !db->dead[(size_t)s];
        if (live_out) live_out[i] = ok;
        if (hidden_out) hidden_out[i] = ok ? db->hidden[(size_t)s] : 0;
        if (timestamp_out) timestamp_out[i] = ok ? db->ts[(size_t)s] : 0;
        if (values_out && db->dim > 0) {
            if (ok) GB_CUDA(cudaMemcpyAsync(values_out + i * db->dim, db->dense->X.p + s * db->dim, sizeof(float) * db->dim, cudaMemcpyDeviceToHost, db->ctx->stream));
            else memset(values_out + i * db->dim, 0, sizeof(float) * db->dim);
        }
    }
    GB_CUDA(cudaStreamSynchronize(db->ctx->stream));
    return GORSE_B200_OK;
}

int32_t gorse_b200_vecdb_get_sparse(gorse_b200_vecdb *db, int64_t slot, uint32_t *indices_out, float *values_out, int32_t cap, int32_t *nnz_out)
{
    GB_CHECK_ARG(db != nullptr && nnz_out != nullptr, "NULL argument");
    GB_CHECK_ARG(db->dim == 0, "not a sparse collection");
    std::lock_guard<std::mutex> lk(db->mu);
    GB_CHECK_ARG(slot >= 0 && slot < (int64_t)db->ts.size(), "slot out of range");
    const int64_t b = db->sparse->h_off[(size_t)slot], e = db->sparse->h_off[(size_t)slot + 1];
    *nnz_out = (int32_t)(e - b);
    for (int64_t t = b; t < e && t - b < cap; t++) {
        if (indices_out) indices_out[t - b] = db->sparse->h_ind[(size_t)t];
        if (values_out) values_out[t - b] = db->sparse->h_val[(size_t)t];
    }
    return GORSE_B200_OK;
}

int32_t gorse_b200_vecdb_delete_before(gorse_b200_vecdb *db, int64_t timestamp_ms, int64_t *slots_out, int64_t cap, int64_t *count_out)
{
    GB_CHECK_ARG(db != nullptr, "collection is NULL");
    std::lock_guard<std::mutex> lk(db->mu);
    int64_t cnt = 0;
    for (size_t i = 0; i < db->ts.size(); i++)
        if (!db->dead[i] && db->ts[i] < timestamp_ms) {   // "timestamp < t", xvec.go:370
            db->dead[i] = 1;
            db->live--;
            if (slots_out && cnt < cap) slots_out[cnt] = (int64_t)i;
            cnt++;
        }
    if (cnt) db->meta_dirty = true;
    if (count_out) *count_out = cnt;
    return GORSE_B200_OK;
}

int32_t gorse_b200_vecdb_query(gorse_b200_vecdb *db, int64_t nq, const float *q_values, const int64_t *q_sp_off, const uint32_t *q_sp_indices,
                               const int32_t *categories, int32_t n_categories, int32_t topk, int64_t *slots_out, float *scores_out,
                               int32_t *count_out)
{
    GB_CHECK_ARG(db != nullptr, "collection is NULL");
    GB_CHECK_ARG(nq >= 0 && n_categories >= 0, "negative count");
    if (nq == 0) return GORSE_B200_OK;
    GB_CHECK_ARG(count_out != nullptr, "count_out is NULL");
    if (topk <= 0) {   // xvec.go:374-376
        for (int64_t q = 0; q < nq; q++) count_out[q] = 0;
        return GORSE_B200_OK;
    }
    GB_CHECK_ARG(slots_out != nullptr && scores_out != nullptr, "NULL output");
    GB_CHECK_ARG(n_categories == 0 || categories != nullptr, "categories is NULL");
    GB_CHECK_ARG(db->dim == 0 || q_values != nullptr, "q_values is NULL");
    GB_CHECK_ARG(db->dim > 0 || q_sp_off != nullptr, "a sparse collection is queried with q_sp_off / q_sp_indices");
    std::lock_guard<std::mutex> lk(db->mu);
    const int64_t n = (int64_t)db->ts.size();
    gorse_b200_ctx *c = db->ctx;
    ScopedDevice sd(c->device);
    if (n == 0) {
        for (int64_t q = 0; q < nq; q++) count_out[q] = 0;
        return GORSE_B200_OK;
    }
    GB_TRY(sync_meta(db));
    if (n_categories) {
        GB_TRY(db->d_want.alloc((size_t)n_categories));
        GB_CUDA(cudaMemcpyAsync(db->d_want.p, categories, sizeof(int32_t) * (size_t)n_categories, cudaMemcpyHostToDevice, c->stream));
    }
    vecdb_allow_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(db->d_blocked.p, db->d_cat_off.p, db->d_cats.p, db->d_want.p, n_categories, n, db->d_allow.p);
    GB_LAUNCHED(c);
    std::vector<int32_t> idx((size_t)nq * topk);
    int32_t st = GORSE_B200_OK;
    if (db->dim > 0) {
        gorse_b200_index *ix = db->dense;
        std::lock_guard<std::mutex> lk2(ix->mu);
        DevBuf<float> d_q, d_dist;
        DevBuf<int32_t> d_idx, d_cnt;
        DevBuf<int> d_nan;
        auto done = [&](int32_t s) { cudaStreamSynchronize(c->stream); d_q.free(); d_dist.free(); d_idx.free(); d_cnt.free(); d_nan.free(); return s; };
        if ((st = d_q.alloc((size_t)nq * db->dim)) || (st = d_dist.alloc((size_t)nq * topk)) || (st = d_idx.alloc((size_t)nq * topk)) ||
            (st = d_cnt.alloc((size_t)nq)) || (st = d_nan.alloc(1)))
            return done(st);
        cudaError_t e = cudaMemsetAsync(d_nan.p, 0, sizeof(int), c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_q.p, q_values, sizeof(float) * (size_t)nq * db->dim, cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) { set_error("vecdb query upload: %s", cudaGetErrorString(e)); return done(GORSE_B200_ERR_CUDA); }
        if ((st = launch_exact_split(ix, d_q.p, nullptr, 0, (int32_t)nq, topk, nullptr, d_idx.p, d_dist.p, d_cnt.p, 0, d_nan.p, db->d_allow.p))) return done(st);
        int h_nan = 0;
        if ((e = cudaMemcpyAsync(idx.data(), d_idx.p, sizeof(int32_t) * (size_t)nq * topk, cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess ||
            (e = cudaMemcpyAsync(scores_out, d_dist.p, sizeof(float) * (size_t)nq * topk, cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess ||
            (e = cudaMemcpyAsync(count_out, d_cnt.p, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess ||
            (e = cudaMemcpyAsync(&h_nan, d_nan.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess ||
            (e = cudaStreamSynchronize(c->stream)) != cudaSuccess) {
            set_error("vecdb query: %s", cudaGetErrorString(e));
            return done(GORSE_B200_ERR_CUDA);
        }
        if (h_nan) { set_error("NaN distance"); return done(GORSE_B200_ERR_ARG); }
        // both metrics of the index are distances (-dot / Euclidean): the score is their negation (xvec.go:425-427; Dot -> dot)
        for (int64_t t = 0; t < nq * topk; t++) scores_out[t] = -scores_out[t];
        st = done(GORSE_B200_OK);
    } else {
        GB_CHECK_ARG(topk <= 128, "sparse collections answer topK <= 128 (got %d)", topk);
        for (int64_t q = 0; q < nq; q++)
            for (int64_t t = q_sp_off[q]; t < q_sp_off[q + 1]; t++)
                GB_CHECK_ARG(t == q_sp_off[q] || q_sp_indices[t] > q_sp_indices[t - 1], "query %lld: indices not strictly ascending", (long long)q);
        std::lock_guard<std::mutex> lk2(db->sparse->mu);
        st = sparse_search_host(db->sparse, 0, nq, q_sp_off, q_sp_indices, q_values, db->d_allow.p, topk, idx.data(), scores_out, count_out);
    }
    if (st) return st;
    for (int64_t t = 0; t < nq * topk; t++) slots_out[t] = idx[(size_t)t];
    return GORSE_B200_OK;
}

// master/tasks.go:925-961: after Fit the item factors of the predictable items become the "collaborative_filtering_<id>"
// collection (distance Dot).  The reference walks GetItemFactor(i) row by row into []vectors.Vector batches; here the rows go
// from the model's item table to the collection device to device.  slot_of_item_out[n_items]: the slot of each item, -1 for
// items without training feedback (they are skipped, :943).  Single-GPU models (the replicated Q of a distributed model is
// complete on every rank, but predictability needs all ranks' rows).
int32_t gorse_b200_vecdb_add_item_factors(gorse_b200_vecdb *db, gorse_b200_cf *cf, const uint8_t *hidden, int64_t timestamp_ms,
                                          const int64_t *cat_off, const int32_t *cats, int64_t *slot_of_item_out)
{
    GB_CHECK_ARG(db != nullptr && cf != nullptr, "NULL collection / model");
    GB_CHECK_ARG(db->dim == cf->d, "collection dimension %d != model factors %d", db->dim, cf->d);
    GB_CHECK_ARG(cf->ctx->world == 1, "hand-off from a distributed model is not supported");
    GB_CHECK_ARG(db->ctx->device == cf->ctx->device, "collection and model live on different devices");
    GB_CHECK_ARG(cat_off == nullptr || cat_off[0] == 0, "cat_off[0] must be 0");
    std::lock_guard<std::mutex> lk(db->mu);
    gorse_b200_ctx *c = db->ctx;
    ScopedDevice sd(c->device);
    const int32_t I = cf->n_items;
    DevBuf<int32_t> d_flag, d_items;
    auto done = [&](int32_t s) { cudaStreamSynchronize(c->stream); d_flag.free(); d_items.free(); return s; };
    int32_t st;
    if ((st = d_flag.alloc((size_t)std::max(I, 1)))) return done(st);
    cudaError_t e = cudaMemsetAsync(d_flag.p, 0, sizeof(int32_t) * (size_t)std::max(I, 1), c->stream);
    if (e == cudaSuccess && cf->n_feedback > 0) {
        vecdb_mark_items_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(cf->user_items.p, cf->n_feedback, d_flag.p);
        c->launches++;
        e = cudaGetLastError();
    }
    std::vector<int32_t> flag((size_t)I);
    if (e == cudaSuccess && I) e = cudaMemcpyAsync(flag.data(), d_flag.p, sizeof(int32_t) * (size_t)I, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { set_error("add_item_factors: %s", cudaGetErrorString(e)); return done(GORSE_B200_ERR_CUDA); }
    std::vector<int32_t> items;
    const int64_t first = (int64_t)db->ts.size();
    for (int32_t i = 0; i < I; i++) {
        if (slot_of_item_out) slot_of_item_out[i] = flag[(size_t)i] ? first + (int64_t)items.size() : -1;
        if (flag[(size_t)i]) items.push_back(i);
    }
    const int64_t n = (int64_t)items.size();
    if (n == 0) return done(GORSE_B200_OK);
    gorse_b200_index *ix = db->dense;
    {
        std::lock_guard<std::mutex> lk2(ix->mu);
        if ((st = index_reserve(ix, ix->n + n))) return done(st);
        if ((st = d_items.alloc((size_t)n))) return done(st);
        e = cudaMemcpyAsync(d_items.p, items.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) {
            if (cf->d % 4 == 0) vecdb_copy_rows_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(cf->Q.p, cf->d / 4, d_items.p, n, ix->X.p + ix->n * ix->d);
            else for (int64_t s2 = 0; s2 < n && e == cudaSuccess; s2++)
                e = cudaMemcpyAsync(ix->X.p + (ix->n + s2) * ix->d, cf->Q.p + (int64_t)items[(size_t)s2] * cf->d, sizeof(float) * cf->d, cudaMemcpyDeviceToDevice, c->stream);
            c->launches++;
            if (e == cudaSuccess) e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { set_error("add_item_factors: %s", cudaGetErrorString(e)); return done(GORSE_B200_ERR_CUDA); }
        ix->n += n;
        ix->mma_ready = false;
    }
    for (int64_t s2 = 0; s2 < n; s2++) {
        const int32_t it = items[(size_t)s2];
        db->hidden.push_back(hidden ? (hidden[it] != 0) : 0);
        db->dead.push_back(0);
        db->ts.push_back(timestamp_ms);
        if (cat_off) db->cats.insert(db->cats.end(), cats + cat_off[it], cats + cat_off[it + 1]);
        db->cat_off.push_back((int64_t)db->cats.size());
        db->live++;
    }
    db->meta_dirty = true;
    return done(GORSE_B200_OK);
}

}  // extern "C"
