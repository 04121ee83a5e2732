"""BASELINE.json's FULL sizes, checked through size-independent properties and spot checks against the oracle
(the oracle cannot run whole configs in seconds, so it is used on samples):
  C2  BPR 1M x 100K x 10M, d=64: sampler validity on 10^6 triples, a sampled triple stream applied in order is bit-exact
      vs the oracle on the full tables, one free-running epoch keeps the factors finite and moves hot and cold rows
  C3  eALS 1M x 100K x 10M, d=128: two epochs keep the factors finite and strictly decrease the objective the sweep
      minimises (a property of exact coordinate descent); 40 user rows (the longest, short ones, random ones) within 1e-4
      of the oracle's half-sweep on a reduced problem with the same S^q (helpers validated on the oracle in
      tests/test_oracle_als.py)
  C4  top-100 over 1M x 128: every row sorted, self-free, k results; 12 rows bit-exact vs the oracle's Bruteforce
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_c2_bpr_full_size(gb, orc):
    from gorse_b200 import synth

    U, I, R, d = 1_000_000, 100_000, 10_000_000, 64
    off, items = synth.make_feedback(U, I, R, seed=1000, zipf_s=1.0, exact=True)
    assert off[-1] == R
    rng = np.random.default_rng(0)
    P = (rng.standard_normal((U, d)) * 0.01).astype(np.float32)
    Q = (rng.standard_normal((I, d)) * 0.01).astype(np.float32)
    with gb.Context(0) as ctx, gb.CFModel(ctx, U, I, d, off, items) as m:
        # integer path at full size: every sampled triple is valid, users are uniform
        t = m.bpr_sample_triples(seed=5, first_step=123456789, n=1_000_000)
        u, i, j = t[:, 0].astype(np.int64), t[:, 1].astype(np.int64), t[:, 2].astype(np.int64)
        key = np.sort(np.repeat(np.arange(U, dtype=np.int64), np.diff(off)) * I + items)  # all (u, item) pairs
        assert np.isin(u * I + i, key, assume_unique=False).all()          # i in R_u
        assert not np.isin(u * I + j, key, assume_unique=False).any()      # j not in R_u
        assert j.min() >= 0 and j.max() < I
        cnt = np.bincount(u, minlength=U)
        assert abs(cnt.mean() - 1.0) < 1e-9 and cnt.max() < 12              # uniform over users (model.go:452-458)
        # the oracle's restatement of the sampler agrees on a prefix
        want = orc.bpr_sample_triples(I, off, items, np.arange(U, dtype=np.int32), 5, 123456789, 20000)
        assert t[:20000].tobytes() == want.tobytes()
        # float path at full table size: in-order application == reference with Jobs = 1, bit for bit
        m.set_factors(P, Q)
        m.bpr_apply_triples(t[:20000], 0.05, 0.01, gb.SCATTER_STORE, gb.ORDER_SEQUENTIAL)
        Pg, Qg = m.get_factors()
        Po, Qo = P.copy(), Q.copy()
        orc.bpr_apply_triples(Po, Qo, t[:20000], 0.05, 0.01)
        assert Pg.tobytes() == Po.tobytes() and Qg.tobytes() == Qo.tobytes()
        # one free-running epoch (hot queue + capped hot apply): finite, and both the hottest and a cold item moved
        m.set_factors(P, Q)
        m.bpr_epoch(0.05, 0.01, R, 77)
        P1, Q1 = m.get_factors()
    assert np.isfinite(P1).all() and np.isfinite(Q1).all()
    pop = np.bincount(items, minlength=I)
    hot, cold = int(pop.argmax()), int(np.argsort(pop)[I // 2])
    assert not np.array_equal(Q1[hot], Q[hot]) and not np.array_equal(Q1[cold], Q[cold])
    moved = (np.abs(P1 - P).max(axis=1) > 0).mean()
    assert moved > 0.99  # ~every user is drawn ~10 times per epoch


def test_c3_als_full_size(gb, orc):
    from gorse_b200 import synth

    from als_checks import als_objective, oracle_user_rows

    U, I, R, d, reg, w = 1_000_000, 100_000, 10_000_000, 128, 0.06, 0.001
    off, items = synth.make_feedback(U, I, R, seed=1000, zipf_s=1.0, exact=True)
    ioff, iusers = gb.transpose_csr(off, items, I)
    rng = np.random.default_rng(3)
    P0 = (rng.standard_normal((U, d), dtype=np.float32) * np.float32(0.1))   # ALS init std, model/cf/model.go:582-583
    Q0 = (rng.standard_normal((I, d), dtype=np.float32) * np.float32(0.1))
    with gb.Context(0) as ctx, gb.CFModel(ctx, U, I, d, off, items, ioff, iusers) as m:
        m.set_factors(P0, Q0)
        m.als_epoch(reg, w)
        P1, Q1 = m.get_factors()
        m.als_epoch(reg, w)
        P2, Q2 = m.get_factors()
    assert np.isfinite(P1).all() and np.isfinite(Q1).all() and np.isfinite(P2).all() and np.isfinite(Q2).all()
    l0, l1, l2 = (als_objective(orc, P, Q, off, items, w, reg) for P, Q in ((P0, Q0), (P1, Q1), (P2, Q2)))
    assert l1 < 0.9 * l0 and l2 < l1, (l0, l1, l2)
    # P after an epoch is the output of the user half-sweep (the item half-sweep only reads it): spot rows vs the oracle
    deg = np.diff(off)
    users = np.unique(np.concatenate([[int(deg.argmax()), int(deg.argmin()), 0, U - 1], np.nonzero(deg == 9)[0][:4], np.nonzero(deg == 17)[0][:4],
                                      np.nonzero(deg == 40)[0][:4], np.nonzero(deg > 96)[0][:4], rng.integers(0, U, 20)])).astype(np.int64)
    want = oracle_user_rows(orc, gb.transpose_csr, P0, Q0, off, items, users, reg, w)
    got = P1[users]
    err = (np.abs(got - want) / np.maximum(np.abs(want).max(axis=1, keepdims=True), 1e-12)).max()
    assert err < 1e-4, err


def test_c4_topk_full_size(gb, orc):
    N, d, k, NQ = 1_000_000, 128, 100, 4096
    rng = np.random.default_rng(0)
    X = rng.standard_normal((N, d), dtype=np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    with gb.Context(0) as ctx, gb.BruteforceIndex(ctx, d, gb.METRIC_NEG_DOT) as ix:
        ix.add(X)
        idx, dist, cnt = ix.search_range(500_000, 500_000 + NQ, k)
    assert (cnt == k).all()
    assert (np.diff(dist, axis=1) >= 0).all()                                   # ascending distances
    assert not (idx == np.arange(500_000, 500_000 + NQ)[:, None]).any()         # never the query itself
    assert all(len(set(r.tolist())) == k for r in idx[::97])                    # no duplicates
    for q in range(0, NQ, 350):
        oi, od = orc.bruteforce_search(X, X[500_000 + q], k, metric=orc.METRIC_NEG_DOT, self_index=500_000 + q)
        assert dist[q].tobytes() == od.tobytes()
        if idx[q].tolist() != oi.tolist():  # only exact fp32 ties may differ in order
            for v in np.unique(od[od < od[-1]]):
                assert set(idx[q][dist[q] == v].tolist()) == set(oi[od == v].tolist())
