// evaluate.cu -- cf.Evaluate on device: model/cf/evaluator.go:35-72 with scorers NDCG, Precision, Recall.
// Kernel 1 scores every (user, candidate) pair with floats.Dot order (internalPredict); kernel 2 runs the
// reference's TopKFilter (Go container/heap, common/heap/filter.go:35-59) per user and the metrics;
// the host adds the per-user metrics in user order, which is the reference with Jobs = 1.
#include <cmath>

#include "cf.cuh"

namespace gb {

#define GB_EVAL_MAX_TOPK 128

__global__ void eval_score_kernel(const float *P, const float *Q, int d, int32_t u_lo, const int32_t *cand_user,
                                  const int32_t *cand_item, int64_t n, float *score)
{
    if (d % 16 == 0) {
        int lane4 = threadIdx.x & 3;
        unsigned mask = quad_mask();
        int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
        int64_t ng = ((int64_t)gridDim.x * blockDim.x) >> 2;
        for (; g < n; g += ng) {
            float v = quad_dot_global(P + (int64_t)(cand_user[g] - u_lo) * d, Q + (int64_t)cand_item[g] * d, d / 16, lane4, mask);
            if (lane4 == 0) score[g] = v;
        }
    } else {
        int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        int64_t ng = (int64_t)gridDim.x * blockDim.x;
        for (; g < n; g += ng) score[g] = dot_any(P + (int64_t)(cand_user[g] - u_lo) * d, Q + (int64_t)cand_item[g] * d, d);
    }
}

// Go container/heap on a min-heap of (value, weight): Less = strict weight compare (common/heap/pq.go:42-48)
struct GoMinHeap {
    int32_t v[GB_EVAL_MAX_TOPK + 1];
    float w[GB_EVAL_MAX_TOPK + 1];
    int n;
    __device__ void swap(int i, int j)
    {
        int32_t tv = v[i]; v[i] = v[j]; v[j] = tv;
        float tw = w[i]; w[i] = w[j]; w[j] = tw;
    }
    __device__ void up(int j)
    {
        for (;;) {
            int i = (j - 1) / 2;
            if (i == j || !(w[j] < w[i])) break;
            swap(i, j);
            j = i;
        }
    }
    __device__ void down(int i0, int nn)
    {
        int i = i0;
        for (;;) {
            int j1 = 2 * i + 1;
            if (j1 >= nn || j1 < 0) break;
            int j = j1, j2 = j1 + 1;
            if (j2 < nn && w[j2] < w[j1]) j = j2;
            if (!(w[j] < w[i])) break;
            swap(i, j);
            i = j;
        }
    }
    __device__ void push(int32_t val, float wt)
    {
        v[n] = val; w[n] = wt; n++;
        up(n - 1);
    }
    __device__ int32_t pop()
    {
        int nn = n - 1;
        swap(0, nn);
        down(0, nn);
        n--;
        return v[n];
    }
};

// one thread per user: Rank (evaluator.go:162-169) + NDCG/Precision/Recall (:75-116)
__global__ void eval_rank_kernel(const int64_t *cand_off, const int32_t *cand_item, const float *score,
                                 const int64_t *test_off, const int32_t *test_items, int32_t n_users, int32_t topk,
                                 const float *inv_log2, float *metrics /* n_users x 4: ndcg, prec, recall, valid */)
{
    int32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_users) return;
    int64_t t0 = test_off[u], nt = test_off[u + 1] - t0;
    float *m = metrics + 4 * (int64_t)u;
    if (nt <= 0) { m[0] = m[1] = m[2] = m[3] = 0.f; return; }
    GoMinHeap h;
    h.n = 0;
    int64_t c0 = cand_off[u], nc = cand_off[u + 1] - c0;
    for (int64_t c = 0; c < nc; c++) {
        h.push(cand_item[c0 + c], score[c0 + c]);
        if (h.n > topk) (void)h.pop();
    }
    int nr = h.n;
    int32_t rank[GB_EVAL_MAX_TOPK];
    for (int i = nr - 1; i >= 0; i--) rank[i] = h.pop();  // PopAllValues: decreasing weight
    const int32_t *tgt = test_items + t0;
    // mapset cardinality = distinct test items
    int card = 0;
    for (int64_t a = 0; a < nt; a++) {
        bool dup = false;
        for (int64_t b = 0; b < a; b++) if (tgt[b] == tgt[a]) { dup = true; break; }
        if (!dup) card++;
    }
    float idcg = 0.f, dcg = 0.f, hit = 0.f;
    for (int i = 0; i < card && i < nr; i++) idcg = idcg + inv_log2[i];
    for (int i = 0; i < nr; i++) {
        bool in = false;
        for (int64_t a = 0; a < nt; a++) if (tgt[a] == rank[i]) { in = true; break; }
        if (in) { dcg = dcg + inv_log2[i]; hit = hit + 1.0f; }
    }
    m[0] = dcg / idcg;
    m[1] = hit / (float)nr;
    m[2] = hit / (float)card;
    m[3] = 1.f;
}

// math32.Log2 as restated in DESIGN.md (frexp, Log(frac) * (1/Ln2) + exp)
static float log2_math32(float x)
{
    int e;
    float frac = frexpf(x, &e);
    if (frac == 0.5f) return (float)(e - 1);
    volatile float l = logf(frac);
    volatile float inv = (float)(1.0 / 0.693147180559945309417232121458176568);
    volatile float p = l * inv;
    return p + (float)e;
}

}  // namespace gb

using namespace gb;

extern "C" int32_t gorse_b200_cf_evaluate(gorse_b200_cf *cf, const int64_t *test_off, const int32_t *test_items,
                                          const int64_t *neg_off, const int32_t *neg_items, int32_t topk, float *out)
{
    GB_CHECK_ARG(cf != nullptr && out != nullptr, "NULL cf/out");
    GB_CHECK_ARG(test_off != nullptr && neg_off != nullptr, "NULL offsets");
    GB_CHECK_ARG(topk >= 1 && topk <= GB_EVAL_MAX_TOPK, "topk %d out of range [1, %d]", topk, GB_EVAL_MAX_TOPK);
    GB_CHECK_ARG(cf->ctx->world == 1, "evaluate needs all user rows on one device (world = 1)");
    const int32_t U = cf->n_users;
    out[0] = out[1] = out[2] = 0.f;
    // candidates = test positives ++ negatives, evaluator.go:51-53
    std::vector<int64_t> cand_off((size_t)U + 1, 0);
    for (int32_t u = 0; u < U; u++) {
        int64_t nt = test_off[u + 1] - test_off[u], nn = neg_off[u + 1] - neg_off[u];
        GB_CHECK_ARG(nt >= 0 && nn >= 0, "offsets not non-decreasing at user %d", u);
        cand_off[(size_t)u + 1] = cand_off[u] + (nt > 0 ? nt + nn : 0);
    }
    int64_t nc = cand_off[U];
    if (nc == 0) { out[0] = out[1] = out[2] = NAN; return GORSE_B200_OK; }  // 0 * (1/0), like the reference
    GB_CHECK_ARG(test_items != nullptr, "test_items is NULL");
    std::vector<int32_t> cu((size_t)nc), ci((size_t)nc);
    for (int32_t u = 0; u < U; u++) {
        int64_t nt = test_off[u + 1] - test_off[u], nn = neg_off[u + 1] - neg_off[u];
        if (nt <= 0) continue;
        int64_t p = cand_off[u];
        for (int64_t a = 0; a < nt; a++, p++) { cu[p] = u; ci[p] = test_items[test_off[u] + a]; }
        for (int64_t a = 0; a < nn; a++, p++) { cu[p] = u; ci[p] = neg_items[neg_off[u] + a]; }
    }
    for (int64_t p = 0; p < nc; p++)
        GB_CHECK_ARG(ci[p] >= 0 && ci[p] < cf->n_items, "candidate item %d out of range", ci[p]);
    std::vector<float> inv_log2((size_t)topk);
    for (int i = 0; i < topk; i++) inv_log2[i] = 1.0f / log2_math32((float)i + 2.0f);

    ScopedDevice sd(cf->ctx->device);
    gorse_b200_ctx *c = cf->ctx;
    DevBuf<int64_t> d_coff, d_toff;
    DevBuf<int32_t> d_cu, d_ci, d_titems;
    DevBuf<float> d_score, d_inv, d_metrics;
    int64_t n_test = test_off[U];
    int32_t st = GORSE_B200_OK;
    auto done = [&](int32_t s) {
        cudaStreamSynchronize(c->stream);
        d_coff.free(); d_toff.free(); d_cu.free(); d_ci.free(); d_titems.free(); d_score.free(); d_inv.free(); d_metrics.free();
        return s;
    };
    if ((st = d_coff.alloc(U + 1)) || (st = d_toff.alloc(U + 1)) || (st = d_cu.alloc(nc)) || (st = d_ci.alloc(nc)) ||
        (st = d_titems.alloc(n_test)) || (st = d_score.alloc(nc)) || (st = d_inv.alloc(topk)) ||
        (st = d_metrics.alloc((size_t)4 * U)))
        return done(st);
    cudaStream_t s = c->stream;
    cudaError_t e = cudaSuccess;
    auto up = [&](void *dst, const void *src, size_t bytes) {
        if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
    };
    up(d_coff.p, cand_off.data(), sizeof(int64_t) * (U + 1));
    up(d_toff.p, test_off, sizeof(int64_t) * (U + 1));
    up(d_cu.p, cu.data(), sizeof(int32_t) * nc);
    up(d_ci.p, ci.data(), sizeof(int32_t) * nc);
    up(d_titems.p, test_items, sizeof(int32_t) * n_test);
    up(d_inv.p, inv_log2.data(), sizeof(float) * topk);
    if (e != cudaSuccess) { set_error("evaluate upload: %s", cudaGetErrorString(e)); return done(GORSE_B200_ERR_CUDA); }
    int threads = 256;
    int64_t work = cf->d % 16 == 0 ? nc * 4 : nc;
    int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((work + threads - 1) / threads, (int64_t)c->sm_count * 16));
    eval_score_kernel<<<blocks, threads, 0, s>>>(cf->P.p, cf->Q.p, cf->d, cf->u_lo, d_cu.p, d_ci.p, nc, d_score.p);
    c->launches++;
    eval_rank_kernel<<<(U + 127) / 128, 128, 0, s>>>(d_coff.p, d_ci.p, d_score.p, d_toff.p, d_titems.p, U, topk, d_inv.p, d_metrics.p);
    c->launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) { set_error("evaluate launch: %s", cudaGetErrorString(e)); return done(GORSE_B200_ERR_CUDA); }
    std::vector<float> metrics((size_t)4 * U);
    if ((e = cudaMemcpyAsync(metrics.data(), d_metrics.p, sizeof(float) * 4 * U, cudaMemcpyDeviceToHost, s)) != cudaSuccess ||
        (e = cudaStreamSynchronize(s)) != cudaSuccess) {
        set_error("evaluate: %s", cudaGetErrorString(e));
        return done(GORSE_B200_ERR_CUDA);
    }
    // evaluator.go:63-71 with one worker: fp32 sums in user order, then * (1/count)
    volatile float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, count = 0.f;
    for (int32_t u = 0; u < U; u++) {
        const float *m = metrics.data() + 4 * (size_t)u;
        if (m[3] != 0.f) {
            count = count + 1.0f;
            sum0 = sum0 + m[0]; sum1 = sum1 + m[1]; sum2 = sum2 + m[2];
        }
    }
    volatile float inv = 1.0f / count;
    out[0] = sum0 * inv; out[1] = sum1 * inv; out[2] = sum2 * inv;
    return done(GORSE_B200_OK);
}
