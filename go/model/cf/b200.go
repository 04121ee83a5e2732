// Copyright 2026 gorse-b200 authors. Drop-in for gorse-io/gorse @ 5404aefa, package model/cf.
//
// NOT COMPILED IN THIS REPOSITORY: the build image has no Go toolchain (see INTEGRATION.md).  Every C call below is
// mirrored one-to-one by the ctypes binding in gorse_b200/_lib.py, which the GPU test-suite exercises.
//
// Build:  go build -tags "b200 cgo" ./...   with  CGO_CFLAGS=-I<repo>/include  CGO_LDFLAGS="-L<repo>/gorse_b200 -lgorse_b200"
// With the tag set, master.newCollaborativeFilteringModel (master/tasks.go:1036-1045) keeps calling NewBPR / NewALS;
// bpr_b200.go / als_b200.go in this directory replace Fit on the two types (see INTEGRATION.md for the 6-line patch).

//go:build b200 && cgo

package cf

/*
#cgo LDFLAGS: -lgorse_b200
#include <stdlib.h>
#include "gorse_b200.h"

// cgo cannot call Go closures from C directly: the progress callback goes through this exported trampoline.
extern int32_t gorseB200Progress(void *user, int32_t epoch, int32_t nEpochs, float ndcg);
static int32_t gorse_b200_progress_tramp(void *user, int32_t epoch, int32_t n, float ndcg) {
    return gorseB200Progress(user, epoch, n, ndcg);
}
static gorse_b200_progress_fn gorse_b200_progress_ptr(void) { return gorse_b200_progress_tramp; }
*/
import "C"

import (
	"context"
	"errors"
	"runtime/cgo"
	"sync"
	"unsafe"

	"github.com/gorse-io/gorse/common/log"
	"github.com/gorse-io/gorse/common/monitor"
	"github.com/gorse-io/gorse/dataset"
	"go.uber.org/zap"
)

// b200Error turns a status code into an error carrying gorse_b200_last_error().
func b200Error(st C.int32_t) error {
	if st == C.GORSE_B200_OK {
		return nil
	}
	return errors.New(C.GoString(C.gorse_b200_last_error()))
}

// one context per process and device; cf.Fit is single-goroutine (master/master.go:457-484)
var (
	b200Once sync.Once
	b200Ctx  *C.gorse_b200_ctx
	b200Err  error
)

func b200Context() (*C.gorse_b200_ctx, error) {
	b200Once.Do(func() {
		b200Err = b200Error(C.gorse_b200_ctx_create(0, &b200Ctx))
	})
	return b200Ctx, b200Err
}

// flattenCSR turns dataset.CFSplit's [][]int32 (dataset/dataset.go:40-60) into offsets + indices.
func flattenCSR(rows [][]int32) ([]int64, []int32) {
	off := make([]int64, len(rows)+1)
	for r, row := range rows {
		off[r+1] = off[r] + int64(len(row))
	}
	idx := make([]int32, off[len(rows)])
	for r, row := range rows {
		copy(idx[off[r]:], row)
	}
	return off, idx
}

func i64ptr(s []int64) *C.int64_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.int64_t)(unsafe.Pointer(&s[0]))
}
func i32ptr(s []int32) *C.int32_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.int32_t)(unsafe.Pointer(&s[0]))
}

// flatMatrix is the host mirror behind UserFactor / ItemFactor: ONE flat Go-owned []float32 per table with [][]float32
// row views into it, so master's row-by-row reads (master/tasks.go:946,969) need no per-row allocation (SURVEY 8b
// "ownership").  The backing array is ordinary Go memory: the row slices stored in BaseMatrixFactorization keep it alive
// for exactly as long as the model is reachable, Clear() (model/cf/model.go:294-307) drops it with the rest of the model,
// and no finalizer or C allocation is involved (round 1 kept the rows in cudaHostAlloc memory behind a finalizer that
// could fire while the rows were still in use).  gorse_b200_cf_get_factors copies into it during the call only.
func newFlatMatrix(n, d int) (flat []float32, rows [][]float32) {
	flat = make([]float32, n*d)
	rows = make([][]float32, n)
	for i := range rows {
		rows[i] = flat[i*d : (i+1)*d : (i+1)*d]
	}
	return flat, rows
}

func f32ptr(s []float32) *C.float {
	if len(s) == 0 {
		return nil
	}
	return (*C.float)(unsafe.Pointer(&s[0]))
}

type fitProgress struct {
	ctx  context.Context
	span *monitor.Span
}

//export gorseB200Progress
func gorseB200Progress(user unsafe.Pointer, epoch, nEpochs C.int32_t, ndcg C.float) C.int32_t {
	p := (*(*cgo.Handle)(user)).Value().(*fitProgress) // user = &handle, the documented way to carry a Handle in a void*
	p.span.Add(1) // model/cf/model.go:519
	if p.ctx.Err() != nil {
		return 1 // cancelled -> Fit returns Score{} (:491-493)
	}
	return 0
}

// fitB200 is the shared body of BPR.Fit / ALS.Fit on the GPU.
func fitB200(ctx context.Context, base *BaseMatrixFactorization, als bool, params C.gorse_b200_fit_params,
	trainSet, valSet dataset.CFSplit, config *FitConfig, spanName string) Score {
	c, err := b200Context()
	if err != nil {
		log.Logger().Error("gorse_b200 context", zap.Error(err))
		return Score{}
	}
	base.Init(trainSet) // UserIndex/ItemIndex + predictable bitsets, model/cf/model.go:129-146
	nUsers, nItems, d := trainSet.CountUsers(), trainSet.CountItems(), int(params.n_factors)
	uOff, uIdx := flattenCSR(trainSet.GetUserFeedback())
	iOff, iIdx := flattenCSR(trainSet.GetItemFeedback())
	tOff, tIdx := flattenCSR(valSet.GetUserFeedback())
	// negatives: NULL -> sampled on the device by the library (dataset.SampleUserNegatives semantics, params.candidates per
	// user; the Go math/rand stream is not reproducible either way, SURVEY F9)

	var cf *C.gorse_b200_cf
	// the library copies during the call and keeps no Go pointer (cgo rule)
	if err := b200Error(C.gorse_b200_cf_create(c, C.int32_t(nUsers), C.int32_t(nItems), C.int32_t(d),
		i64ptr(uOff), i32ptr(uIdx), i64ptr(iOff), i32ptr(iIdx), &cf)); err != nil {
		log.Logger().Error("gorse_b200 cf_create", zap.Error(err))
		return Score{}
	}
	defer C.gorse_b200_cf_destroy(cf)

	_, span := monitor.Start(ctx, spanName, int(params.n_epochs)) // :442 / :639
	defer span.End()
	h := cgo.NewHandle(&fitProgress{ctx: ctx, span: span})
	defer h.Delete()

	params.verbose, params.candidates = C.int32_t(config.Verbose), C.int32_t(config.Candidates)
	params.topk, params.patience = C.int32_t(config.TopK), C.int32_t(config.Patience)
	var res C.gorse_b200_fit_result
	var st C.int32_t
	if als {
		st = C.gorse_b200_als_fit(cf, &params, i64ptr(tOff), i32ptr(tIdx), nil, nil,
			C.gorse_b200_progress_ptr(), unsafe.Pointer(&h), &res)
	} else {
		st = C.gorse_b200_bpr_fit(cf, &params, i64ptr(tOff), i32ptr(tIdx), nil, nil,
			C.gorse_b200_progress_ptr(), unsafe.Pointer(&h), &res)
	}
	if err := b200Error(st); err != nil {
		log.Logger().Error("gorse_b200 fit", zap.Error(err))
		return Score{}
	}
	if res.cancelled != 0 {
		log.Logger().Info("fit canceled", zap.Int("epoch", int(res.epochs_run)))
		return Score{}
	}
	// host mirror: Predict == floats.Dot(GetUserFactor, GetItemFactor) bit for bit (model_test.go:53-54)
	pFlat, pRows := newFlatMatrix(nUsers, d)
	qFlat, qRows := newFlatMatrix(nItems, d)
	if err := b200Error(C.gorse_b200_cf_get_factors(cf, f32ptr(pFlat), f32ptr(qFlat))); err != nil {
		log.Logger().Error("gorse_b200 get_factors", zap.Error(err))
		return Score{}
	}
	base.UserFactor, base.ItemFactor = pRows, qRows
	return Score{NDCG: float32(res.ndcg), Precision: float32(res.precision), Recall: float32(res.recall)}
}
