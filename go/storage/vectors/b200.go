// Copyright 2026 gorse-b200 authors. Drop-in for gorse-io/gorse @ 5404aefa, package storage/vectors.
//
// NOT COMPILED IN THIS REPOSITORY (no Go toolchain in the build image, see INTEGRATION.md).  Every C call below is mirrored
// one-to-one by gorse_b200.VectorCollection (gorse_b200/__init__.py), which tests/test_vecdb_gpu.py drives through the same
// ABI on the reference's own test values (storage/vectors/database_test.go).
//
// A vectors.Database (database.go:107-120) whose collections live in GPU memory: item-to-item, user-to-user and CF retrieval
// reach it with NO change to master / worker / logics -- set `database.vector_store = "b200://0"` (device ordinal) and the
// init() below makes vectors.Open (database.go:168-175) return it.  The library works on slots and int32 category ids; this
// file owns the id <-> slot and category string <-> id maps, which is all the reference's backends keep on the host too.

//go:build b200 && cgo

package vectors

/*
#cgo LDFLAGS: -lgorse_b200
#include <stdlib.h>
#include "gorse_b200.h"
*/
import "C"

import (
	"context"
	"fmt"
	"runtime"
	"sort"
	"strconv"
	"strings"
	"sync"
	"time"
	"unsafe"

	"github.com/gorse-io/gorse/storage"
	"github.com/juju/errors"
)

const b200Prefix = "b200://"

func init() {
	Register([]string{b200Prefix}, func(path, tablePrefix string, opts ...storage.Option) (Database, error) {
		device, err := strconv.Atoi(strings.TrimPrefix(path, b200Prefix))
		if err != nil {
			device = 0
		}
		return &B200{device: device, tablePrefix: tablePrefix, collections: map[string]*b200Collection{}}, nil
	})
}

type b200Collection struct {
	info    CollectionInfo
	handle  *C.gorse_b200_vecdb
	slotOf  map[string]int64 // id -> live slot
	idOf    []string         // slot -> id
	catsOf  [][]string       // slot -> categories (returned with query results, xvec.go:428-430)
	catID   map[string]int32
	catName []string
}

// B200 implements Database on one GPU.
type B200 struct {
	mu          sync.RWMutex
	device      int
	tablePrefix string
	ctx         *C.gorse_b200_ctx
	collections map[string]*b200Collection
}

var _ Database = (*B200)(nil)

func b200Err(st C.int32_t) error {
	if st == C.GORSE_B200_OK {
		return nil
	}
	return errors.New(C.GoString(C.gorse_b200_last_error()))
}

func (db *B200) Init() error {
	db.mu.Lock()
	defer db.mu.Unlock()
	if db.ctx != nil {
		return nil
	}
	return b200Err(C.gorse_b200_ctx_create(C.int32_t(db.device), &db.ctx))
}

func (db *B200) Optimize(ctx context.Context, name string) error { return nil } // brute force: nothing to build

func (db *B200) Close() error {
	db.mu.Lock()
	defer db.mu.Unlock()
	for name, c := range db.collections {
		C.gorse_b200_vecdb_destroy(c.handle)
		delete(db.collections, name)
	}
	if db.ctx != nil {
		C.gorse_b200_ctx_destroy(db.ctx)
		db.ctx = nil
	}
	return nil
}

func (db *B200) ListCollections(ctx context.Context) ([]string, error) {
	db.mu.RLock()
	defer db.mu.RUnlock()
	names := make([]string, 0, len(db.collections))
	for name := range db.collections {
		names = append(names, name)
	}
	sort.Strings(names)
	return names, nil
}

func (db *B200) collection(name string) (*b200Collection, error) {
	c, ok := db.collections[name]
	if !ok {
		return nil, fmt.Errorf("collection %s: %w", name, storage.ErrNotFound)
	}
	return c, nil
}

func (db *B200) DescribeCollection(ctx context.Context, name string) (*CollectionInfo, error) {
	db.mu.RLock()
	defer db.mu.RUnlock()
	c, err := db.collection(name)
	if err != nil {
		return nil, err
	}
	info := c.info
	return &info, nil
}

func (db *B200) AddCollection(ctx context.Context, name string, dimensions int, distance Distance, config VectorConfig) error {
	if config.Type != QuantizationNone {
		return fmt.Errorf("quantization type %s for b200 %w", config.Type, storage.ErrNotSupported)
	}
	db.mu.Lock()
	defer db.mu.Unlock()
	if _, exists := db.collections[name]; exists {
		return errors.Errorf("collection %s already exists", name)
	}
	c := &b200Collection{info: CollectionInfo{Name: name, Dimension: dimensions, Distance: distance},
		slotOf: map[string]int64{}, catID: map[string]int32{}}
	// vectors.Distance and gorse_b200_distance share their numbering (database.go:27-33)
	if err := b200Err(C.gorse_b200_vecdb_create(db.ctx, C.int32_t(dimensions), C.int32_t(distance), &c.handle)); err != nil {
		return err
	}
	db.collections[name] = c
	return nil
}

func (db *B200) DeleteCollection(ctx context.Context, name string) error {
	db.mu.Lock()
	defer db.mu.Unlock()
	c, err := db.collection(name)
	if err != nil {
		return err
	}
	delete(db.collections, name)
	return b200Err(C.gorse_b200_vecdb_destroy(c.handle))
}

func (db *B200) CountVectors(ctx context.Context, name string) (int64, error) {
	db.mu.RLock()
	defer db.mu.RUnlock()
	c, err := db.collection(name)
	if err != nil {
		return 0, err
	}
	var live C.int64_t
	if err := b200Err(C.gorse_b200_vecdb_count(c.handle, &live, nil)); err != nil {
		return 0, err
	}
	return int64(live), nil
}

// AddVectors is an upsert by id (xvec.go:288-318): one flat upload per call.
func (db *B200) AddVectors(ctx context.Context, name string, vectors []Vector) error {
	if len(vectors) == 0 {
		return nil
	}
	db.mu.Lock()
	defer db.mu.Unlock()
	c, err := db.collection(name)
	if err != nil {
		return err
	}
	n, dim := len(vectors), c.info.Dimension
	var values []float32
	var spOff []int64
	var spIdx []uint32
	if dim == 0 {
		spOff = make([]int64, n+1)
	}
	hidden, ts, replace := make([]uint8, n), make([]int64, n), make([]int64, n)
	catOff, cats := make([]int64, n+1), []int32{}
	seen := map[string]int{}
	for i, v := range vectors {
		if dim > 0 {
			if len(v.Values) != dim {
				return errors.Errorf("vector %s has dimension %d, want %d", v.Id, len(v.Values), dim)
			}
			values = append(values, v.Values...)
		} else {
			values, spIdx = append(values, v.Values...), append(spIdx, v.Indices...)
			spOff[i+1] = int64(len(spIdx))
		}
		if v.IsHidden {
			hidden[i] = 1
		}
		ts[i] = v.Timestamp.UnixMilli()
		for _, cat := range v.Categories {
			id, ok := c.catID[cat]
			if !ok {
				id = int32(len(c.catName))
				c.catID[cat], c.catName = id, append(c.catName, cat)
			}
			cats = append(cats, id)
		}
		catOff[i+1] = int64(len(cats))
		replace[i] = -1
		if old, ok := c.slotOf[v.Id]; ok {
			replace[i] = old
		}
		if j, dup := seen[v.Id]; dup { // the same id twice in one call: the later one wins, like an upsert
			hidden[j] = 1
		}
		seen[v.Id] = i
	}
	var first C.int64_t
	st := C.gorse_b200_vecdb_add(c.handle, C.int64_t(n), f32p(values), i64p(spOff), u32p(spIdx), u8p(hidden), i64p(ts),
		i64p(catOff), i32p(cats), i64p(replace), &first)
	runtime.KeepAlive(values)
	if err := b200Err(st); err != nil {
		return err
	}
	for i, v := range vectors {
		c.slotOf[v.Id] = int64(first) + int64(i)
		c.idOf = append(c.idOf, v.Id)
		c.catsOf = append(c.catsOf, v.Categories)
	}
	return nil
}

func (db *B200) GetVectors(ctx context.Context, name string, ids []string) ([]Vector, error) {
	if len(ids) == 0 {
		return []Vector{}, nil
	}
	db.mu.RLock()
	defer db.mu.RUnlock()
	c, err := db.collection(name)
	if err != nil {
		return nil, err
	}
	found := make([]Vector, 0, len(ids))
	for _, id := range ids {
		slot, ok := c.slotOf[id]
		if !ok {
			continue
		}
		v, err := db.getSlot(c, slot)
		if err != nil {
			return nil, err
		}
		found = append(found, v)
	}
	return orderVectors(ids, found), nil // database.go:138-157
}

func (db *B200) getSlot(c *b200Collection, slot int64) (Vector, error) {
	v := Vector{Id: c.idOf[slot], Categories: c.catsOf[slot]}
	var hidden, live C.uint8_t
	var ts C.int64_t
	s := C.int64_t(slot)
	if c.info.Dimension > 0 {
		v.Values = make([]float32, c.info.Dimension)
		if err := b200Err(C.gorse_b200_vecdb_get(c.handle, &s, 1, f32p(v.Values), &hidden, &ts, &live)); err != nil {
			return v, err
		}
	} else {
		if err := b200Err(C.gorse_b200_vecdb_get(c.handle, &s, 1, nil, &hidden, &ts, &live)); err != nil {
			return v, err
		}
		var nnz C.int32_t
		if err := b200Err(C.gorse_b200_vecdb_get_sparse(c.handle, s, nil, nil, 0, &nnz)); err != nil {
			return v, err
		}
		v.Indices, v.Values = make([]uint32, nnz), make([]float32, nnz)
		if err := b200Err(C.gorse_b200_vecdb_get_sparse(c.handle, s, u32p(v.Indices), f32p(v.Values), nnz, &nnz)); err != nil {
			return v, err
		}
	}
	v.IsHidden, v.Timestamp = hidden != 0, time.UnixMilli(int64(ts)).UTC()
	return v, nil
}

func (db *B200) DeleteVectors(ctx context.Context, name string, timestamp time.Time) error {
	db.mu.Lock()
	defer db.mu.Unlock()
	c, err := db.collection(name)
	if err != nil {
		return err
	}
	slots := make([]int64, len(c.idOf))
	var count C.int64_t
	if err := b200Err(C.gorse_b200_vecdb_delete_before(c.handle, C.int64_t(timestamp.UnixMilli()), i64p(slots), C.int64_t(len(slots)), &count)); err != nil {
		return err
	}
	for _, slot := range slots[:count] {
		if c.slotOf[c.idOf[slot]] == slot {
			delete(c.slotOf, c.idOf[slot])
		}
	}
	return nil
}

// QueryVectors: filter "hidden = false AND categories CONTAIN_ALL (...)" (xvec.go:381-389) evaluated on the device.
func (db *B200) QueryVectors(ctx context.Context, name string, q Vector, categories []string, topK int) ([]ScoredVector, error) {
	if topK <= 0 {
		return []ScoredVector{}, nil
	}
	db.mu.RLock()
	defer db.mu.RUnlock()
	c, err := db.collection(name)
	if err != nil {
		return nil, err
	}
	want := make([]int32, 0, len(categories))
	for _, cat := range categories {
		id, ok := c.catID[cat]
		if !ok {
			return []ScoredVector{}, nil // nobody carries this category
		}
		want = append(want, id)
	}
	slots, scores := make([]int64, topK), make([]float32, topK)
	var count C.int32_t
	var qOff []int64
	if c.info.Dimension == 0 {
		qOff = []int64{0, int64(len(q.Indices))}
	}
	st := C.gorse_b200_vecdb_query(c.handle, 1, f32p(q.Values), i64p(qOff), u32p(q.Indices), i32p(want), C.int32_t(len(want)),
		C.int32_t(topK), i64p(slots), f32p(scores), &count)
	if err := b200Err(st); err != nil {
		return nil, err
	}
	results := make([]ScoredVector, 0, count)
	for i := 0; i < int(count); i++ {
		v, err := db.getSlot(c, slots[i])
		if err != nil {
			return nil, err
		}
		results = append(results, ScoredVector{Vector: v, Score: scores[i]})
	}
	return results, nil
}

func f32p(s []float32) *C.float {
	if len(s) == 0 {
		return nil
	}
	return (*C.float)(unsafe.Pointer(&s[0]))
}
func i64p(s []int64) *C.int64_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.int64_t)(unsafe.Pointer(&s[0]))
}
func i32p(s []int32) *C.int32_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.int32_t)(unsafe.Pointer(&s[0]))
}
func u32p(s []uint32) *C.uint32_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.uint32_t)(unsafe.Pointer(&s[0]))
}
func u8p(s []uint8) *C.uint8_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.uint8_t)(unsafe.Pointer(&s[0]))
}
