// Drop-in for gorse-io/gorse @ 5404aefa, package common/ann: a GPU-resident ann.Index (common/ann/ann.go:21-25)
// with ann.Bruteforce semantics (common/ann/bruteforce.go:24-83).  NOT COMPILED HERE (no Go toolchain).

//go:build b200 && cgo

package ann

/*
#cgo LDFLAGS: -lgorse_b200
#include "gorse_b200.h"
*/
import "C"

import (
	"runtime"
	"unsafe"

	"github.com/pkg/errors"
	"github.com/samber/lo"
)

type B200Metric int32

const (
	B200Euclidean B200Metric = C.GORSE_B200_METRIC_EUCLIDEAN // floats.Euclidean
	B200NegDot    B200Metric = C.GORSE_B200_METRIC_NEG_DOT   // -floats.Dot (logics/cf.go:32-34)
)

// B200Bruteforce implements Index for []float32 vectors.
type B200Bruteforce struct {
	ctx *C.gorse_b200_ctx
	ix  *C.gorse_b200_index
	dim int
}

var _ Index = (*B200Bruteforce)(nil)

func NewB200Bruteforce(dim int, metric B200Metric) (*B200Bruteforce, error) {
	b := &B200Bruteforce{dim: dim}
	if st := C.gorse_b200_ctx_create(0, &b.ctx); st != 0 {
		return nil, errors.New(C.GoString(C.gorse_b200_last_error()))
	}
	if st := C.gorse_b200_index_create(b.ctx, C.int32_t(dim), C.int32_t(metric), &b.ix); st != 0 {
		C.gorse_b200_ctx_destroy(b.ctx)
		return nil, errors.New(C.GoString(C.gorse_b200_last_error()))
	}
	runtime.SetFinalizer(b, func(b *B200Bruteforce) {
		C.gorse_b200_index_destroy(b.ix)
		C.gorse_b200_ctx_destroy(b.ctx)
	})
	return b, nil
}

// Add appends one vector and returns the length after the append, like Bruteforce.Add (bruteforce.go:33-37).
func (b *B200Bruteforce) Add(v []float32) int {
	defer runtime.KeepAlive(b) // the finalizer must not destroy the handles while a C call is using them
	var n C.int64_t
	if st := C.gorse_b200_index_add(b.ix, (*C.float)(unsafe.Pointer(&v[0])), 1, &n); st != 0 {
		panic(C.GoString(C.gorse_b200_last_error())) // the reference's Add cannot fail
	}
	return int(n)
}

// AddBatch uploads n vectors stored contiguously (one PCIe copy instead of n).
func (b *B200Bruteforce) AddBatch(flat []float32) int {
	defer runtime.KeepAlive(b) // the finalizer must not destroy the handles while a C call is using them
	var n C.int64_t
	if st := C.gorse_b200_index_add(b.ix, (*C.float)(unsafe.Pointer(&flat[0])), C.int64_t(len(flat)/b.dim), &n); st != 0 {
		panic(C.GoString(C.gorse_b200_last_error()))
	}
	return int(n)
}

func collect(idx []int32, dist []float32, n int32) []lo.Tuple2[int, float32] {
	out := make([]lo.Tuple2[int, float32], 0, n)
	for i := int32(0); i < n; i++ {
		out = append(out, lo.Tuple2[int, float32]{A: int(idx[i]), B: dist[i]})
	}
	return out
}

// SearchIndex never returns q itself; out-of-range q is an error (bruteforce.go:39-63).
func (b *B200Bruteforce) SearchIndex(q, k int, prune0 bool) ([]lo.Tuple2[int, float32], error) {
	defer runtime.KeepAlive(b) // the finalizer must not destroy the handles while a C call is using them
	idx, dist := make([]int32, max(k, 1)), make([]float32, max(k, 1))
	var cnt C.int32_t
	qi := C.int64_t(q)
	st := C.gorse_b200_index_search_indices(b.ix, &qi, 1, C.int32_t(k), C.int32_t(lo.Ternary(prune0, 1, 0)),
		(*C.int32_t)(unsafe.Pointer(&idx[0])), (*C.float)(unsafe.Pointer(&dist[0])), &cnt)
	if st == C.GORSE_B200_ERR_RANGE {
		return nil, errors.Errorf("index out of range: %v", q)
	} else if st != 0 {
		return nil, errors.New(C.GoString(C.gorse_b200_last_error()))
	}
	return collect(idx, dist, int32(cnt)), nil
}

// SearchVector mirrors bruteforce.go:65-83 (a NaN distance panics like heap.PriorityQueue.Push, pq.go:82-83).
func (b *B200Bruteforce) SearchVector(q []float32, k int, prune0 bool) []lo.Tuple2[int, float32] {
	defer runtime.KeepAlive(b) // the finalizer must not destroy the handles while a C call is using them
	idx, dist := make([]int32, max(k, 1)), make([]float32, max(k, 1))
	var cnt C.int32_t
	st := C.gorse_b200_index_search_vectors(b.ix, (*C.float)(unsafe.Pointer(&q[0])), 1, C.int32_t(k),
		C.int32_t(lo.Ternary(prune0, 1, 0)), (*C.int32_t)(unsafe.Pointer(&idx[0])), (*C.float)(unsafe.Pointer(&dist[0])), &cnt)
	if st != 0 {
		panic(C.GoString(C.gorse_b200_last_error()))
	}
	return collect(idx, dist, int32(cnt))
}

// AllNeighbors returns the k nearest neighbours of every stored vector in [q0, q1): what item-to-item / user-to-user
// ask their vector store for (logics/item_to_item.go:50-62), in one call.
func (b *B200Bruteforce) AllNeighbors(q0, q1, k int, prune0 bool) ([][]lo.Tuple2[int, float32], error) {
	defer runtime.KeepAlive(b) // the finalizer must not destroy the handles while a C call is using them
	nq := q1 - q0
	idx, dist, cnt := make([]int32, max(nq*k, 1)), make([]float32, max(nq*k, 1)), make([]int32, max(nq, 1))
	st := C.gorse_b200_index_search_range(b.ix, C.int64_t(q0), C.int64_t(q1), C.int32_t(k), C.int32_t(lo.Ternary(prune0, 1, 0)),
		(*C.int32_t)(unsafe.Pointer(&idx[0])), (*C.float)(unsafe.Pointer(&dist[0])), (*C.int32_t)(unsafe.Pointer(&cnt[0])))
	if st != 0 {
		return nil, errors.New(C.GoString(C.gorse_b200_last_error()))
	}
	out := make([][]lo.Tuple2[int, float32], nq)
	for r := range out {
		out[r] = collect(idx[r*k:], dist[r*k:], cnt[r])
	}
	return out, nil
}

// SimilarScore is one (row, score) of QuerySimilar; the caller maps rows back to item / user ids.
type SimilarScore struct {
	Row   int
	Score float64
}

// QuerySimilar is logics.QueryItemToItem / logics.QueryUserToUser (logics/item_to_item.go:50-86,
// logics/user_to_user.go:50-86) for every stored vector in [q0, q1) in one call: the n nearest neighbours without the
// vector itself, Dot scores <= 0 dropped, Euclidean mapped to 1/(1+dist); scoreScale is .5 for type "auto".
func (b *B200Bruteforce) QuerySimilar(q0, q1, n int, scoreScale float64) ([][]SimilarScore, error) {
	defer runtime.KeepAlive(b) // the finalizer must not destroy the handles while a C call is using them
	nq := q1 - q0
	ids, scores, cnt := make([]int32, max(nq*n, 1)), make([]float64, max(nq*n, 1)), make([]int32, max(nq, 1))
	st := C.gorse_b200_index_query_similar(b.ix, C.int64_t(q0), C.int64_t(q1), C.int32_t(n), C.double(scoreScale),
		(*C.int32_t)(unsafe.Pointer(&ids[0])), (*C.double)(unsafe.Pointer(&scores[0])), (*C.int32_t)(unsafe.Pointer(&cnt[0])))
	if st != 0 {
		return nil, errors.New(C.GoString(C.gorse_b200_last_error()))
	}
	out := make([][]SimilarScore, nq)
	for r := range out {
		out[r] = make([]SimilarScore, cnt[r])
		for t := range out[r] {
			out[r][t] = SimilarScore{Row: int(ids[r*n+t]), Score: scores[r*n+t]}
		}
	}
	return out, nil
}

// TruncateBF16 stores a dense embedding the way the reference does (bfloats.FromFloat32 + ToFloat32,
// logics/item_to_item.go:151-164); vectors added this way are represented exactly by the index's bf16 tensor-core mirror.
func TruncateBF16(v []float32) {
	if len(v) > 0 {
		C.gorse_b200_bf16_truncate((*C.float)(unsafe.Pointer(&v[0])), C.int64_t(len(v)), (*C.float)(unsafe.Pointer(&v[0])))
	}
}
