"""BASELINE.json's C3 at FULL size (eALS 1M x 100K x 10M, d=128) through size-independent properties: two epochs keep the
factors finite and strictly decrease the objective the sweep minimises (a property of exact coordinate descent); 40 user
rows (the longest, short ones, random ones) are within 1e-4 of the oracle's half-sweep on a reduced problem with the same S^q.
The helpers are validated on the oracle in tests/test_oracle_als.py.  (Its own file, collected last: it was added after
round 1's GPU budget was spent, so its first run on hardware is the round-end run.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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
