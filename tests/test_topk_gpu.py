"""GPU parity of the brute-force index against the oracle's ann.Bruteforce (common/ann/bruteforce.go:24-83):
indices bit-exact, distances bit-exact (the exact re-rank uses the reference's fp32 summation order)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(gb):
    c = gb.Context(0)
    yield c
    c.close()


def test_reference_known_answers(gb, orc, ctx):
    # logics/cf_test.go:26-58 (-Dot): items k*(1,1,1), query (1,1,1), n=3 -> ids 5,4,3 with scores 15,12,9
    with gb.BruteforceIndex(ctx, 3, gb.METRIC_NEG_DOT) as ix:
        for k in range(1, 6):
            assert ix.add(np.full((1, 3), k, np.float32)) == k  # Add returns len after append (bruteforce.go:33-37)
        idx, dist, cnt = ix.search_vectors([[1, 1, 1]], 3)
        assert cnt[0] == 3 and (idx[0] + 1).tolist() == [5, 4, 3] and (-dist[0]).tolist() == [15, 12, 9]
    # worker/worker_test.go:194-221: item i = (i,1), user (1,0) -> 9,8,7,6
    with gb.BruteforceIndex(ctx, 2, gb.METRIC_NEG_DOT) as ix:
        ix.add(np.array([[i, 1] for i in range(10)], np.float32))
        idx, dist, cnt = ix.search_vectors([[1, 0]], 4)
        assert idx[0].tolist() == [9, 8, 7, 6]
        # prune0 keeps score > 0 only (bruteforce.go:78): every -dot here is <= 0
        idx, dist, cnt = ix.search_vectors([[1, 0]], 10, prune0=True)
        assert cnt[0] == 0 and (idx[0] == -1).all()


@pytest.mark.parametrize("metric", ["euclid", "negdot"])
@pytest.mark.parametrize("d,N", [(16, 500), (128, 2000), (64, 1000), (10, 300), (784, 200), (24, 257)])
def test_search_matches_oracle(gb, orc, ctx, metric, d, N):
    rng = np.random.default_rng(d * 7 + N)
    X = rng.standard_normal((N, d)).astype(np.float32)
    gm = gb.METRIC_EUCLIDEAN if metric == "euclid" else gb.METRIC_NEG_DOT
    om = orc.METRIC_EUCLIDEAN if metric == "euclid" else orc.METRIC_NEG_DOT
    k = 100 if N >= 500 else 17
    with gb.BruteforceIndex(ctx, d, gm) as ix:
        ix.add(X[: N // 2])
        assert ix.add(X[N // 2:]) == N and len(ix) == N
        Qv = rng.standard_normal((9, d)).astype(np.float32)
        idx, dist, cnt = ix.search_vectors(Qv, k)
        for q in range(9):
            oi, od = orc.bruteforce_search(X, Qv[q], k, metric=om)
            assert cnt[q] == len(oi) == min(k, N)
            assert idx[q, :cnt[q]].tolist() == oi.tolist()
            assert dist[q, :cnt[q]].tobytes() == od.tobytes()
        # SearchIndex: never returns the query itself (bruteforce.go:47)
        qs = np.array([0, 5, N - 1], np.int64)
        idx, dist, cnt = ix.search_indices(qs, k, prune0=True)
        for r, q in enumerate(qs):
            oi, od = orc.bruteforce_search(X, X[q], k, prune0=True, metric=om, self_index=int(q))
            assert cnt[r] == len(oi) and idx[r, :cnt[r]].tolist() == oi.tolist()
            assert dist[r, :cnt[r]].tobytes() == od.tobytes() and q not in idx[r]
        # all-pairs over a range == per-query SearchIndex
        idx, dist, cnt = ix.search_range(3, 11, 5)
        oi, od, oc, _ = orc.bruteforce_all(X, 3, 11, 5, metric=om)
        assert idx.tolist() == oi.tolist() and dist.tobytes() == od.tobytes() and cnt.tolist() == oc.tolist()


def test_edge_cases(gb, orc, ctx):
    with gb.BruteforceIndex(ctx, 4, gb.METRIC_EUCLIDEAN) as ix:
        idx, dist, cnt = ix.search_vectors(np.zeros((2, 4), np.float32), 3)  # empty index
        assert cnt.tolist() == [0, 0]
        ix.add(np.eye(4, dtype=np.float32)[:3])
        idx, dist, cnt = ix.search_vectors(np.zeros((1, 4), np.float32), 10)  # k > N
        assert cnt[0] == 3 and sorted(idx[0, :3].tolist()) == [0, 1, 2] and (idx[0, 3:] == -1).all()
        idx, dist, cnt = ix.search_vectors(np.zeros((1, 4), np.float32), 0)  # k = 0
        assert cnt[0] == 0
        with pytest.raises(gb.GorseB200Error) as e:  # bruteforce.go:41-43
            ix.search_indices(np.array([3], np.int64), 2)
        assert e.value.status == -5 and "out of range" in e.value.message
        with pytest.raises(gb.GorseB200Error) as e:  # pq.go:82-83 panics on NaN
            ix.search_vectors(np.full((1, 4), np.nan, np.float32), 2)
        assert "NaN" in e.value.message
    with pytest.raises(gb.GorseB200Error):
        gb.BruteforceIndex(ctx, 4, 7)


def test_ties_are_ordered_by_index(gb, orc, ctx):
    # the reference's tie order is a container/heap artefact no reference test pins; ours is (distance, index)
    X = np.zeros((40, 16), np.float32)
    X[:, 0] = np.repeat(np.arange(10), 4)
    with gb.BruteforceIndex(ctx, 16, gb.METRIC_EUCLIDEAN) as ix:
        ix.add(X)
        idx, dist, cnt = ix.search_vectors(np.zeros((1, 16), np.float32), 10)
    assert idx[0].tolist() == list(range(10)) and dist[0].tolist() == [0, 0, 0, 0, 1, 1, 1, 1, 2, 2]
    oi, od = orc.bruteforce_search(X, np.zeros(16, np.float32), 10, metric=orc.METRIC_EUCLIDEAN)
    assert od.tolist() == dist[0].tolist()  # same distance multiset as the Go-heap order
