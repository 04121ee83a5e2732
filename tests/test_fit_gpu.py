"""gorse_b200_bpr_fit / _als_fit: the host mirror of cf.BPR.Fit / cf.ALS.Fit (model/cf/model.go:408-530, 609-775):
Verbose cadence, early stopping on Patience, cancellation, returned Score == Evaluate of the final factors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def data(gb):
    from gorse_b200 import synth

    U, I = 1200, 300
    off, items = synth.make_feedback(U, I, 24000, seed=4, n_clusters=6)
    train, test = synth.leave_one_out(off, items, seed=1)
    neg = synth.sample_negatives(I, train, test, 100, seed=2)
    ioff, iusers = gb.transpose_csr(train[0], train[1], I)
    return U, I, train, test, neg, ioff, iusers


@pytest.mark.parametrize("kind", ["bpr", "als"])
def test_fit_scores_and_cadence(gb, orc, data, kind):
    U, I, train, test, neg, ioff, iusers = data
    seen = []
    with gb.Context(0) as ctx, gb.CFModel(ctx, U, I, 16, train[0], train[1], ioff, iusers) as m:
        kw = dict(n_epochs=12, verbose=5, seed=3)
        if kind == "als":
            kw.update(reg=0.015, alpha=0.05)
        res = m.fit(kind, test[0], test[1], neg[0], neg[1], progress=lambda ep, n, ndcg: seen.append((ep, n, ndcg)) or False, **kw)
        P, Q = m.get_factors()
        dev = m.evaluate(test[0], test[1], neg[0], neg[1], 10)
    assert res.epochs_run == 12 and not res.cancelled and not res.early_stopped
    assert [e for e, _, _ in seen] == list(range(1, 13)) and all(n == 12 for _, n, _ in seen)
    # evaluated at epochs 5, 10 (verbose) and 12 (last), model.go:496
    assert [e for e, _, s in seen if s >= 0] == [5, 10, 12]
    want = orc.evaluate(P, Q, test[0], test[1], neg[0], neg[1], 10)
    assert (res.ndcg, res.precision, res.recall) == tuple(want) == tuple(dev)
    assert res.ndcg > 0.15  # it learned (random is ~0.05)


def test_early_stopping_and_cancel(gb, data):
    U, I, train, test, neg, ioff, iusers = data
    with gb.Context(0) as ctx, gb.CFModel(ctx, U, I, 16, train[0], train[1], ioff, iusers) as m:
        # lr = 0: NDCG never improves on epoch 0 -> stop at the first evaluation after `patience` epochs (model.go:508-517)
        res = m.fit("bpr", test[0], test[1], neg[0], neg[1], n_epochs=50, verbose=2, patience=4, lr=0.0, reg=0.0, seed=1)
        assert res.early_stopped and res.best_epoch == 0 and res.epochs_run == 6
        # cancellation (ctx.Err() != nil): zero Score like the reference (model.go:491-493)
        res = m.fit("bpr", test[0], test[1], neg[0], neg[1], progress=lambda ep, n, s: ep >= 3, n_epochs=50, verbose=10)
        assert res.cancelled and res.epochs_run == 3 and (res.ndcg, res.precision, res.recall) == (0.0, 0.0, 0.0)


def test_fit_samples_its_own_negatives(gb, orc, data):
    """neg_off = NULL: Fit draws the negatives itself like the reference's Evaluate (evaluator.go:42, seed 0)."""
    U, I, train, test, neg, ioff, iusers = data
    with gb.Context(0) as ctx, gb.CFModel(ctx, U, I, 16, train[0], train[1]) as m:
        res = m.fit("bpr", test[0], test[1], None, None, n_epochs=10, verbose=5, seed=3, candidates=50)
        P, Q = m.get_factors()
    noff, nitems = orc.sample_user_negatives(I, train[0], train[1], test[0], test[1], 50, seed=0)
    want = orc.evaluate(P, Q, test[0], test[1], noff, nitems, 10)
    assert (res.ndcg, res.precision, res.recall) == tuple(want) and res.ndcg > 0.15
