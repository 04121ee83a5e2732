// bpr_hot.cuh -- the hot-item half of a BPR epoch: queue, counting sort, capped-concurrency apply.
//
// Why it exists: with a popularity head (Zipf(1.0): the top item is the positive of ~7 % of all
// triples) free-running Hogwild on a GPU has two problems the reference's 8-128 goroutines do not have:
//   throughput  all updates of one row serialise on the atomic unit of one L2 slice and back-pressure every SM;
//   stability   ~4*10^4 triples are in flight, i.e. thousands of stale updates of the SAME hot row overlap; once
//               lr * (sum of sigma'(x) |p_u|^2 + reg) * (#overlapping updates) exceeds ~2 the row oscillates and the
//               factors blow up to NaN after 30-60 epochs (measured; the sequential reference is stable).
// So triples whose POSITIVE item is hot are not applied by the free-running kernel.  It appends them to a queue;
// the queue is grouped by item (counting sort) and a second kernel applies them with a HARD CAP on how many
// triples of the same hot item are in flight: item h gets k_h quads (k_h proportional to its share of the queue,
// at most GB_HOT_ROW_CONCURRENCY), and only those quads ever touch row h, each one triple at a time.  Hot rows live
// in a striped side table for the epoch (every 16-byte piece in its own 128-byte line, the 16 pieces of a row in 16
// planes) so that the L2 atomic units of many slices share one row (tools/l2_atomic_probe.cu: 2.7x).
//
// Tried and rejected (kept in git history, "BPR v4 owner"): one CTA per hot item with the row in shared memory and
// 128-triple batches.  Correct and stable, but the top item's 700 K triples are ~6*10^7 warp instructions, i.e.
// 16.7 ms on the single SM that owns it (measured) -- a hot item needs many SMs.
#pragma once
#include "cf.cuh"

namespace gb {

#define GB_HOTQ_REGIONS 64      // independent append counters (one per group of warps)
#define GB_HOT_ROW_CONCURRENCY 768   // most triples of one hot item in flight at any time (at lr = 0.05; scaled by 0.05/lr)
#define GB_HOT_SLOT_FLOATS 32        // one 128-byte line per 16-byte piece of a hot row

struct HotQueue {
    int32_t *entries;                  // [GB_HOTQ_REGIONS][region_cap][3]  (slot, u, j)
    unsigned long long *counts;        // [GB_HOTQ_REGIONS]
    int64_t region_cap;
};

// Opportunistically warp-aggregated append.  Returns false when the region is full (caller applies the triple itself).
// The aggregation runs over whatever lanes happen to be converged (`__activemask()`, the coalesced-group pattern): each such
// group elects a leader for ONE atomicAdd.  The four lanes of a quad always call this together (a quad owns one triple), but
// independent thread scheduling does not promise that they sit in the same converged group, so the quad's verdict is
// broadcast afterwards with the QUAD's own mask -- `__shfl_sync` then waits for exactly those four lanes (the
// first version read lane4 == 0 through the opportunistic mask).
__device__ __forceinline__ bool hotq_append(const HotQueue &hq, bool want, int lane4, int32_t slot, int32_t u, int32_t j)
{
    const int lane = threadIdx.x & 31;
    const unsigned act = __activemask();
    const unsigned m = __ballot_sync(act, want && lane4 == 0);
    bool ok = true;
    if (m) {
        const int leader = __ffs(m) - 1;
        const int region = (int)((((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) % GB_HOTQ_REGIONS);
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(hq.counts + region, (unsigned long long)__popc(m));
        base = __shfl_sync(act, base, leader);
        if (want && lane4 == 0) {
            const unsigned long long pos = base + __popc(m & ((1u << lane) - 1));
            if (pos < (unsigned long long)hq.region_cap) {
                int32_t *e = hq.entries + ((size_t)region * hq.region_cap + pos) * 3;
                e[0] = slot; e[1] = u; e[2] = j;
            } else ok = false;
        }
    }
    // every lane of a quad must agree with its lane 0
    return __shfl_sync(0xFu << (lane & ~3), ok, lane & ~3) != 0;
}

// ---- counting sort of the queue by slot ---------------------------------------------------------
__global__ void hot_hist_kernel(HotQueue hq, int n_hot, unsigned *hist)
{
    extern __shared__ unsigned s_hist[];
    for (int t = threadIdx.x; t < n_hot; t += blockDim.x) s_hist[t] = 0;
    __syncthreads();
    const int region = blockIdx.y;
    const int64_t n = min((int64_t)hq.counts[region], hq.region_cap);
    const int32_t *e = hq.entries + (size_t)region * hq.region_cap * 3;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&s_hist[e[3 * t]], 1u);
    __syncthreads();
    for (int t = threadIdx.x; t < n_hot; t += blockDim.x)
        if (s_hist[t]) atomicAdd(&hist[t], s_hist[t]);
}

// exclusive scan of <= 1024 counts by one CTA of 1024 threads; begin[n_hot] = total; cursor = begin.
// Also the quad assignment: slot t gets k_t = clamp(ceil(count_t * quad_budget / total), 1, cap) quads,
// first_quad = exclusive scan of k, first_quad[n_hot] = quads in use.
__global__ void hot_scan_kernel(const unsigned *hist, int n_hot, unsigned *begin, unsigned *cursor, unsigned quad_budget,
                                unsigned cap, unsigned *first_quad)
{
    __shared__ unsigned s[1024];
    const int t = threadIdx.x;
    unsigned v = t < n_hot ? hist[t] : 0;
    s[t] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        unsigned a = t >= off ? s[t - off] : 0;
        __syncthreads();
        s[t] += a;
        __syncthreads();
    }
    const unsigned total = s[1023];
    if (t < n_hot) { begin[t] = s[t] - v; cursor[t] = s[t] - v; }
    if (t == n_hot - 1) begin[n_hot] = s[t];
    __syncthreads();
    unsigned k = 0;
    if (t < n_hot && v > 0) {
        k = (unsigned)(((unsigned long long)v * quad_budget + total - 1) / max(total, 1u));
        k = min(max(k, 1u), min(cap, v));
        k = (k + 7u) & ~7u;  // whole warps: the 8 quads of a warp serve one hot row and combine their updates
    }
    s[t] = k;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        unsigned a = t >= off ? s[t - off] : 0;
        __syncthreads();
        s[t] += a;
        __syncthreads();
    }
    if (t < n_hot) first_quad[t] = s[t] - k;
    if (t == n_hot - 1) first_quad[n_hot] = s[t];
}

__global__ void hot_fill_kernel(HotQueue hq, int n_hot, unsigned *cursor, int32_t *sorted /* [total][2] (u, j) */)
{
    extern __shared__ unsigned s_cnt[];  // [n_hot] local counts, then local bases
    for (int t = threadIdx.x; t < n_hot; t += blockDim.x) s_cnt[t] = 0;
    __syncthreads();
    const int region = blockIdx.y;
    const int64_t n = min((int64_t)hq.counts[region], hq.region_cap);
    const int32_t *e = hq.entries + (size_t)region * hq.region_cap * 3;
    // this CTA owns a contiguous tile of the region
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * per, t1 = min(n, t0 + per);
    for (int64_t t = t0 + threadIdx.x; t < t1; t += blockDim.x) atomicAdd(&s_cnt[e[3 * t]], 1u);
    __syncthreads();
    for (int t = threadIdx.x; t < n_hot; t += blockDim.x) {
        unsigned c = s_cnt[t];
        s_cnt[t] = c ? atomicAdd(&cursor[t], c) : 0;  // reserve a range per (CTA, slot); now a running cursor
    }
    __syncthreads();
    for (int64_t t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
        const int32_t slot = e[3 * t];
        const unsigned pos = atomicAdd(&s_cnt[slot], 1u);
        sorted[2 * (size_t)pos] = e[3 * t + 1];
        sorted[2 * (size_t)pos + 1] = e[3 * t + 2];
    }
}

// hot rows <-> striped side table, one 16-byte piece per thread
__global__ void hot_gather_kernel(const float *Q, int d, const int32_t *hot_items, int n_hot, int hot_pad, float *hot)
{
    int pieces = d / 4;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_hot * pieces) return;
    int s = (int)(t / pieces), pc = (int)(t % pieces);
    float4 v = *reinterpret_cast<const float4 *>(Q + (int64_t)hot_items[s] * d + 4 * pc);
    *reinterpret_cast<float4 *>(hot + ((int64_t)pc * hot_pad + s) * GB_HOT_SLOT_FLOATS) = v;
}
__global__ void hot_scatter_kernel(float *Q, int d, const int32_t *hot_items, int n_hot, int hot_pad, const float *hot)
{
    int pieces = d / 4;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_hot * pieces) return;
    int s = (int)(t / pieces), pc = (int)(t % pieces);
    float4 v = *reinterpret_cast<const float4 *>(hot + ((int64_t)pc * hot_pad + s) * GB_HOT_SLOT_FLOATS);
    *reinterpret_cast<float4 *>(Q + (int64_t)hot_items[s] * d + 4 * pc) = v;
}

}  // namespace gb
