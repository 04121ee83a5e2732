"""GPU vector collection (gorse_b200_vecdb_*) against the reference's own vector-store tests
(storage/vectors/database_test.go:86-330, lifted values) and against the oracle's brute force on random collections with
hidden flags, category filters (CONTAIN_ALL), upserts and timestamp deletes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(gb):
    c = gb.Context(0)
    yield c
    c.close()


def unit(d, *kv):
    v = np.zeros(d, np.float32)
    for k, x in kv:
        v[k] = x
    return v


def test_reference_TestVectors_and_TestHidden(gb, ctx):
    # database_test.go:86-158 (Cosine collection, categories) with ids a = slot 0, b = slot 1; categories cat-a=0 cat-b=1 common=2
    d = 32
    with gb.VectorCollection(ctx, d, gb.DISTANCE_COSINE) as col:
        assert col.count() == (0, 0)
        a, b = unit(d, (0, 1.0)), unit(d, (0, 0.9), (1, 0.1))
        assert col.add(np.stack([a, b]), categories=[[0, 2], [1, 2]]) == 0
        assert col.count() == (2, 2)
        s, sc, n = col.query(a, [0], 10)
        assert n[0] == 1 and s[0, 0] == 0
        s, sc, n = col.query(a, [2], 10)
        assert n[0] == 2 and s[0, :2].tolist() == [0, 1] and sc[0, 0] > sc[0, 1]
        s, sc, n = col.query(a, [0, 2], 10)
        assert n[0] == 1 and s[0, 0] == 0
        s, sc, n = col.query(a, [], 1)
        assert n[0] == 1
        vals, hid, ts, live = col.get([0])
        s, sc, n = col.query(vals[0], [], 10)
        assert n[0] == 2 and s[0, :2].tolist() == [0, 1]
    # :246-277 TestHidden: a hidden vector is never returned, with or without a category filter
    with gb.VectorCollection(ctx, 4, gb.DISTANCE_COSINE) as col:
        q = np.array([1, 0, 0, 0], np.float32)
        col.add(np.stack([np.array([0.9, 0.1, 0, 0], np.float32), q]), hidden=[0, 1], categories=[[0, 1], [0, 1]])
        assert col.count()[0] == 2
        for cats in ([], [0], [1]):
            s, sc, n = col.query(q, cats, 10)
            assert n[0] == 1 and s[0, 0] == 0


def test_reference_TestDot_TestGetVectors_TestDelete(gb, ctx):
    # :279-298 TestDot: score = dot, a (2,0,0,0) before b (1,1,0,0)
    with gb.VectorCollection(ctx, 4, gb.DISTANCE_DOT) as col:
        col.add(np.array([[2, 0, 0, 0], [1, 1, 0, 0]], np.float32))
        s, sc, n = col.query(np.array([1, 0, 0, 0], np.float32), [], 2)
        assert n[0] == 2 and s[0].tolist() == [0, 1] and sc[0].tolist() == [2.0, 1.0]
        s, sc, n = col.query(np.array([1, 0, 0, 0], np.float32), [], 0)   # topK <= 0 -> nothing (xvec.go:374-376)
        assert n[0] == 0
    # :160-190 TestGetVectors: values, hidden flag and timestamps come back; unknown slots are reported dead
    with gb.VectorCollection(ctx, 4, gb.DISTANCE_EUCLIDEAN) as col:
        col.add(np.array([[1, 0, 0, 0], [0, 1, 0, 0]], np.float32), hidden=[0, 1], timestamps=[1000, 2000])
        vals, hid, ts, live = col.get([1, 7, 0, 1])
        assert live.tolist() == [1, 0, 1, 1] and hid.tolist() == [1, 0, 0, 1] and ts.tolist() == [2000, 0, 1000, 2000]
        assert vals[0].tolist() == [0, 1, 0, 0] and vals[2].tolist() == [1, 0, 0, 0]
        # Euclidean: score = -distance (xvec.go:425-427)
        s, sc, n = col.query(np.array([1, 0, 0, 0], np.float32), [], 5)
        assert n[0] == 1 and s[0, 0] == 0 and sc[0, 0] == 0.0
        # :300-330 TestDeleteVectors: strictly-older-than semantics, upsert replaces
        assert col.delete_before(1000).tolist() == [] and col.delete_before(1001).tolist() == [0]
        assert col.count() == (1, 2)
        first = col.add(np.array([[0, 0, 1, 0]], np.float32), timestamps=[3000], replace=[1])   # new version of the id in slot 1
        assert first == 2 and col.count() == (1, 3)
        assert col.get([1])[3].tolist() == [0]


def test_reference_TestSparse(gb, ctx):
    # :192-244: old (1,100 -> 1,1), match (1,100 -> 1,2), other (2,200 -> 1,2); query = match's vector
    with gb.VectorCollection(ctx, 0, gb.DISTANCE_DOT) as col:
        cutoff = 10_000
        col.add(sparse=[([1, 100], [1, 1]), ([1, 100], [1, 2]), ([2, 200], [1, 2])], timestamps=[cutoff - 3600_000, cutoff, cutoff])
        assert col.count()[0] == 3
        ind, val = col.get_sparse(1)
        assert ind.tolist() == [1, 100] and val.tolist() == [1.0, 2.0]
        s, sc, n = col.query([([1, 100], [1, 2])], [], 10)
        assert n[0] == 2 and s[0, :2].tolist() == [1, 0] and sc[0, :2].tolist() == [5.0, 3.0]   # "other" has dot 0: dropped (:421-423)
        assert col.delete_before(cutoff).tolist() == [0] and col.count()[0] == 2
        s, sc, n = col.query([([1, 100], [1, 2])], [], 10)
        assert n[0] == 1 and s[0, 0] == 1
    with pytest.raises(gb.GorseB200Error):      # distance method for sparse vector not supported (xvec.go:243-245)
        gb.VectorCollection(ctx, 0, gb.DISTANCE_EUCLIDEAN)


@pytest.mark.parametrize("dist,d", [("dot", 64), ("euclid", 40), ("cosine", 16)])
def test_filtered_queries_match_the_oracle(gb, orc, ctx, dist, d):
    rng = np.random.default_rng(d)
    N, NQ, k = 5000, 40, 25
    X = rng.standard_normal((N, d)).astype(np.float32)
    hidden = (rng.random(N) < 0.2).astype(np.uint8)
    cats = [sorted(rng.choice(6, size=int(rng.integers(0, 4)), replace=False).tolist()) for _ in range(N)]
    ts = rng.integers(0, 1000, N)
    gd = {"dot": gb.DISTANCE_DOT, "euclid": gb.DISTANCE_EUCLIDEAN, "cosine": gb.DISTANCE_COSINE}[dist]
    with gb.VectorCollection(ctx, d, gd) as col:
        col.add(X[:3000], hidden=hidden[:3000], timestamps=ts[:3000], categories=cats[:3000])
        col.add(X[3000:], hidden=hidden[3000:], timestamps=ts[3000:], categories=cats[3000:])
        dead = np.zeros(N, bool)
        dead[col.delete_before(150)] = True
        assert dead.sum() == (ts < 150).sum()
        Q = rng.standard_normal((NQ, d)).astype(np.float32)
        for want in ([], [2], [1, 4]):
            slots, scores, cnt = col.query(Q, want, k)
            allow = np.array([not hidden[i] and not dead[i] and set(want) <= set(cats[i]) for i in range(N)])
            ids = np.nonzero(allow)[0]
            for q in range(NQ):
                if dist == "cosine":
                    dots = X[ids].astype(np.float64) @ Q[q].astype(np.float64)
                    sim = dots / (np.linalg.norm(X[ids].astype(np.float64), axis=1) * np.linalg.norm(Q[q].astype(np.float64)))
                    order = np.argsort(-sim, kind="stable")[:k]
                    assert cnt[q] == min(k, len(ids))
                    # fp32 vs float64: compare the scores, and the ids wherever neighbours are not within rounding of each other
                    np.testing.assert_allclose(scores[q, :cnt[q]], (sim[order] - 1.0), atol=2e-6)
                    assert set(slots[q, :cnt[q]].tolist()) ^ set(ids[order].tolist()) == set() or np.min(np.diff(-sim[np.argsort(-sim)][:k + 1])) < 1e-6
                    continue
                om = orc.METRIC_NEG_DOT if dist == "dot" else orc.METRIC_EUCLIDEAN
                oi, od = orc.bruteforce_search(np.ascontiguousarray(X[ids]), Q[q], k, metric=om)
                assert cnt[q] == len(oi)
                assert scores[q, :cnt[q]].tobytes() == (-od).astype(np.float32).tobytes()        # scores bit-exact
                if slots[q, :cnt[q]].tolist() != ids[oi].tolist():                                # only exact ties may reorder
                    for v in np.unique(od):
                        assert set(slots[q, :cnt[q]][scores[q, :cnt[q]] == -v].tolist()) == set(ids[oi][od == v].tolist())


def test_sparse_filtered_queries_match_the_oracle(gb, orc, ctx):
    rng = np.random.default_rng(3)
    N, F, k = 1200, 400, 30
    idf = (rng.random(F) * 3 + 0.01).astype(np.float32)
    pop = 1.0 / np.arange(1, F + 1)
    vecs = []
    for _ in range(N):
        m = int(rng.integers(1, 30))
        vecs.append(gb.sparse_vector(np.unique(rng.choice(F, size=m, p=pop / pop.sum())), idf))
    hidden = (rng.random(N) < 0.25).astype(np.uint8)
    cats = [[int(rng.integers(0, 3))] for _ in range(N)]
    with gb.VectorCollection(ctx, 0, gb.DISTANCE_DOT) as col:
        col.add(sparse=vecs, hidden=hidden, categories=cats)
        qs = [vecs[i] for i in range(0, N, 60)] + [(np.array([F + 5], np.uint32), np.array([1.0], np.float32))]   # + a feature nobody has
        for want in ([], [1]):
            slots, scores, cnt = col.query(qs, want, k)
            allow = np.array([not hidden[i] and set(want) <= set(cats[i]) for i in range(N)])
            for qi, (qind, qval) in enumerate(qs):
                dots = np.array([orc.sparse_dot(qind, qval, v[0], v[1]) if allow[j] else 0.0 for j, v in enumerate(vecs)], np.float32)
                order = sorted([j for j in range(N) if dots[j] > 0], key=lambda j: (-dots[j], j))[:k]
                assert cnt[qi] == len(order) and slots[qi, :cnt[qi]].tolist() == order
                assert scores[qi, :cnt[qi]].tobytes() == dots[order].tobytes()


def test_item_factors_hand_off(gb, orc, ctx):
    """master/tasks.go:925-961 without the host round trip: the predictable items' factor rows go device to device into a Dot
    collection; CF retrieval then == the oracle's brute force over those rows (worker/worker_test.go:194-221 pattern)."""
    import ctypes as C

    from gorse_b200 import _lib, synth

    U, I, d = 300, 120, 32
    # feedback only touches items 0..99: items 100..119 are not predictable
    o2, it2 = synth.make_feedback(U, 100, 2500, seed=3)
    rng = np.random.default_rng(0)
    P = rng.standard_normal((U, d)).astype(np.float32)
    Q = rng.standard_normal((I, d)).astype(np.float32)
    hidden = (rng.random(I) < 0.1).astype(np.uint8)
    with gb.CFModel(ctx, U, I, d, o2, it2) as m, gb.VectorCollection(ctx, d, gb.DISTANCE_DOT) as col:
        m.set_factors(P, Q)
        slot = np.zeros(I, np.int64)
        gb.check(_lib.lib.gorse_b200_vecdb_add_item_factors(col.h, m.h, gb.ptr(hidden), 1234, None, None, gb.ptr(slot)))
        present = np.zeros(I, bool)
        present[np.unique(it2)] = True
        assert ((slot >= 0) == present).all() and slot[present].tolist() == list(range(present.sum()))
        assert col.count() == (int(present.sum()), int(present.sum()))
        vals, hid, ts, live = col.get(slot[present][:5])
        assert vals.tobytes() == Q[np.nonzero(present)[0][:5]].tobytes() and (ts == 1234).all()
        s, sc, n = col.query(P[:20], [], 10)
        ids = np.nonzero(present & (hidden == 0))[0]
        for u in range(20):
            oi, od = orc.bruteforce_search(np.ascontiguousarray(Q[ids]), P[u], 10, metric=orc.METRIC_NEG_DOT)
            assert sc[u, :n[u]].tobytes() == (-od).astype(np.float32).tobytes()
            assert [int(np.nonzero(slot == x)[0][0]) for x in s[u, :n[u]]] == ids[oi].tolist() or len(set(od.tolist())) < len(od)
