"""Sparse Dot index (csrc/sparse.cu) against the oracle's sparse brute force and the reference's known neighbour
orders (logics/item_to_item_test.go:212-316).  First hardware run: round 2 (4 passed)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(gb):
    c = gb.Context(0)
    yield c
    c.close()


def _pack(vectors):
    off = np.concatenate([[0], np.cumsum([len(v[0]) for v in vectors])]).astype(np.int64)
    ind = np.concatenate([v[0] for v in vectors]).astype(np.uint32)
    val = np.concatenate([v[1] for v in vectors]).astype(np.float32)
    return off, ind, val


def test_reference_known_answers(gb, orc, ctx):
    idf = np.ones(101, np.float32)
    off, ind, val = _pack([gb.sparse_vector(list(range(1, 101 - i)), idf) for i in range(100)])   # TestTags / TestUsers
    with gb.SparseIndex(ctx) as ix:
        assert ix.add(off, ind, val) == 100 and len(ix) == 100
        idx, dot, cnt = ix.search_range(0, 1, 10)
        assert cnt[0] == 10 and idx[0].tolist() == list(range(1, 11)) and dot[0].tolist() == [float(100 - i) for i in range(1, 11)]
        ids, sc = gb.similar_scores(gb.METRIC_NEG_DOT, 1.0, 0, 10, idx[0], -dot[0])
        assert ids.tolist() == list(range(1, 11))
    auto = []
    for i in range(100):                                                                            # TestAuto
        ti, tv = gb.sparse_vector(list(range(1, 101 - i)) if i % 2 == 0 else [], idf)
        ui, uv = gb.sparse_vector(list(range(1, 101 - i)) if i % 2 == 1 else [], idf, offset=len(idf))
        auto.append((np.concatenate([ti, ui]), np.concatenate([tv, uv])))
    with gb.SparseIndex(ctx) as ix:
        ix.add(*_pack(auto))
        idx, dot, cnt = ix.search_range(0, 2, 10)
        assert idx[0].tolist() == [2 * i for i in range(1, 11)] and idx[1].tolist() == [2 * i + 1 for i in range(1, 11)]


@pytest.mark.parametrize("N,F,nnz,k", [(400, 300, 12, 20), (3000, 5000, 40, 100), (1500, 200, 60, 37)])
def test_matches_oracle(gb, orc, ctx, N, F, nnz, k):
    rng = np.random.default_rng(N + F)
    idf = (rng.random(F) * 3 + 0.01).astype(np.float32)
    pop = 1.0 / np.arange(1, F + 1)
    vecs = []
    for _ in range(N):
        m = int(rng.integers(0, 2 * nnz))
        ids = np.unique(rng.choice(F, size=m, p=pop / pop.sum())) if m else np.zeros(0, np.int64)
        vecs.append(gb.sparse_vector(ids, idf))
    off, ind, val = _pack(vecs)
    with gb.SparseIndex(ctx) as ix:
        ix.add(off[:N // 2 + 1], ind[:off[N // 2]], val[:off[N // 2]])                   # two batches
        ix.add(off[N // 2:] - off[N // 2], ind[off[N // 2]:], val[off[N // 2]:])
        idx, dot, cnt = ix.search_range(0, N, k)
    for q in range(0, N, max(1, N // 150)):
        oi, od = orc.sparse_bruteforce_search(off, ind, val, q, k)
        keep = od > 0
        oi, od = oi[keep], od[keep]
        assert cnt[q] == len(oi), q
        assert dot[q, :cnt[q]].tobytes() == od.tobytes(), q                # dots bit-identical, best first
        if idx[q, :cnt[q]].tolist() != oi.tolist():                          # only exact ties may differ in order
            for v in np.unique(od[od > od[-1]]) if len(od) else []:
                assert set(idx[q, :cnt[q]][dot[q, :cnt[q]] == v].tolist()) == set(oi[od == v].tolist())
        assert (idx[q, cnt[q]:] == -1).all() and q not in idx[q].tolist()
