"""Device cf.Evaluate (model/cf/evaluator.go:35-72) == oracle Evaluate, bit for bit (Jobs = 1 summation order)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d", [16, 8, 64, 20])
def test_evaluate_bit_exact(gb, orc, d):
    from gorse_b200 import synth

    U, I = 700, 250
    off, items = synth.make_feedback(U, I, 9000, seed=d, n_clusters=5)
    train, test = synth.leave_one_out(off, items, seed=1)
    # some users without test items, one with two (duplicates allowed by the CSR contract)
    to, ti = test[0].copy(), test[1].copy()
    neg = synth.sample_negatives(I, train, (to, ti), 100, seed=2)
    rng = np.random.default_rng(0)
    P = rng.standard_normal((U, d)).astype(np.float32)
    Q = rng.standard_normal((I, d)).astype(np.float32)
    Q[::7] = Q[3]  # equal scores exercise the Go-heap tie mechanics inside TopKFilter
    with gb.Context(0) as ctx, gb.CFModel(ctx, U, I, d, train[0], train[1]) as m:
        m.set_factors(P, Q)
        for topk in (10, 4, 1):
            got = m.evaluate(to, ti, neg[0], neg[1], topk)
            want = orc.evaluate(P, Q, to, ti, neg[0], neg[1], topk)
            assert got.tobytes() == want.tobytes(), (topk, got, want)


def test_reference_known_answer(gb, orc):
    # model/cf/evaluator_test.go:137-171 TestEvaluate -> Precision@4 == 0.625
    pos = [{0, 1, 2, 3}, {4, 5, 6}, {8, 9}, {12}]
    neg = [set(), {7}, {10, 11}, {13, 14, 15}]
    U, I = 4, 16
    P = np.eye(U, dtype=np.float32)
    Q = np.zeros((I, U), np.float32)
    for u in range(U):
        for i in pos[u]:
            Q[i, u] = 1
        for i in neg[u]:
            Q[i, u] = -1
    test_off = np.arange(0, 17, 4, dtype=np.int64)
    test_items = np.arange(16, dtype=np.int32)
    neg_items, neg_off = [], [0]
    for u in range(U):
        neg_items += [i for i in range(I) if not (4 * u <= i < 4 * u + 4)]
        neg_off.append(len(neg_items))
    with gb.Context(0) as ctx, gb.CFModel(ctx, U, I, U, test_off, test_items) as m:
        m.set_factors(P, Q)
        out = m.evaluate(test_off, test_items, np.array(neg_off, np.int64), np.array(neg_items, np.int32), 4)
    assert out[1] == np.float32(0.625)


@pytest.mark.parametrize("I,n_cand", [(250, 100), (40, 100), (120, 30)])
def test_device_negatives_equal_the_oracle_sampler(gb, orc, I, n_cand):
    """gorse_b200_eval_create(neg_off = NULL) samples dataset.SampleUserNegatives on the device: identical int32 lists to the
    oracle's restatement on the shared counter RNG, incl. the "fewer than n_cand items left -> all of them, ascending"
    branch (common/util/random.go:116-122), users without test items, duplicate test items."""
    from gorse_b200 import synth

    U, d = 500, 16
    off, items = synth.make_feedback(U, I, 6000, seed=I, n_clusters=4)
    train, test = synth.leave_one_out(off, items, seed=1)
    to, ti = test[0].copy(), test[1].copy()
    rng = np.random.default_rng(1)
    P = rng.standard_normal((U, d)).astype(np.float32)
    Q = rng.standard_normal((I, d)).astype(np.float32)
    with gb.Context(0) as ctx, gb.CFModel(ctx, U, I, d, train[0], train[1]) as m:
        m.set_factors(P, Q)
        with m.eval_plan(to, ti, None, None, n_candidates=n_cand, seed=5, topk=10) as plan:
            noff, nitems = plan.negatives()
            got = plan.run()
            m.set_factors(Q[:1].repeat(U, 0) * 0 + P[::-1], Q)     # the plan scores the CURRENT factors
            got2 = plan.run()
    ooff, oitems = orc.sample_user_negatives(I, train[0], train[1], to, ti, n_cand, seed=5)
    assert noff.tolist() == ooff.tolist() and nitems.tolist() == oitems.tolist()
    for u in range(0, U, 17):   # distinct, outside train and test
        row = nitems[noff[u]:noff[u + 1]]
        excl = set(train[1][train[0][u]:train[0][u + 1]].tolist()) | set(ti[to[u]:to[u + 1]].tolist())
        assert len(set(row.tolist())) == len(row) == min(n_cand, I - len(excl)) and not (set(row.tolist()) & excl)
    assert got.tobytes() == orc.evaluate(P, Q, to, ti, noff, nitems, 10).tobytes()
    assert got2.tobytes() == orc.evaluate(np.ascontiguousarray(P[::-1]), Q, to, ti, noff, nitems, 10).tobytes()
