"""Pins the oracle's Go container/heap, TopKFilter, PriorityQueue and ranking metrics on the
reference's own known answers (values lifted from common/heap/*_test.go, model/cf/evaluator_test.go)."""
import numpy as np


def test_topk_filter_known_answers(orc):
    # common/heap/filter_test.go:23-47
    v, w = orc.topk_filter([10, 20, 30], [2, 8, 1], 3)
    assert v.tolist() == [20, 10, 30]
    v, w = orc.topk_filter([10, 20, 30, 40, 50, 12, 67, 32], [2, 8, 1, 2, 5, 10, 7, 9], 3)
    assert v.tolist() == [12, 32, 20] and w.tolist() == [10, 9, 8]


def test_priority_queue_known_answers(orc):
    # common/heap/pq_test.go:28-60: ascending pops, Reverse() pops descending
    els = [5, 3, 7, 8, 6, 2, 9]
    v, w = orc.pq_push_pop_all(els, els, desc=False)
    assert v.tolist() == sorted(els) and w.tolist() == sorted(els)
    v, w = orc.pq_push_pop_all(els, els, desc=False, reverse_first=True)
    assert v.tolist() == sorted(els, reverse=True)
    # duplicates are dropped silently (pq.go:84)
    v, w = orc.pq_push_pop_all([1, 1, 2], [3.0, 4.0, 5.0], desc=False)
    assert v.tolist() == [1, 2] and w.tolist() == [3.0, 5.0]


def test_metrics_known_answers(orc):
    # model/cf/evaluator_test.go:33-67, eps 1e-5
    rank = list(range(10))
    assert abs(orc.metric("ndcg", [1, 3, 5, 7], rank) - 0.6766372989) < 1e-5
    assert abs(orc.metric("precision", [1, 3, 5, 7], rank) - 0.4) < 1e-5
    assert abs(orc.metric("recall", [1, 3, 15, 17, 19], rank) - 0.4) < 1e-5
    assert abs(orc.metric("map", [1, 3, 7, 9], rank) - 0.44375) < 1e-5
    assert abs(orc.metric("mrr", [3], rank) - 0.25) < 1e-5
    assert orc.metric("hr", [3], rank) == 1 and orc.metric("hr", [30], rank) == 0


def test_evaluate_known_answer(orc):
    # model/cf/evaluator_test.go:137-171 TestEvaluate: user u's test items are 4u..4u+3, the mock scores
    # +1/-1/0 are realised here with d=1 factors; numCandidates = all items -> Precision@4 == 0.625
    pos = [{0, 1, 2, 3}, {4, 5, 6}, {8, 9}, {12}]
    neg = [set(), {7}, {10, 11}, {13, 14, 15}]
    U, I = 4, 16
    # d = U: P = identity rows, Q[i][u] = score(u, i)
    P = np.eye(U, dtype=np.float32)
    Q = np.zeros((I, U), np.float32)
    for u in range(U):
        for i in pos[u]:
            Q[i, u] = 1
        for i in neg[u]:
            Q[i, u] = -1
    test_off = np.arange(0, 17, 4, dtype=np.int64)
    test_items = np.arange(16, dtype=np.int32)
    # SampleUserNegatives with numCandidates >= remaining items returns all other items ascending
    neg_items, neg_off = [], [0]
    for u in range(U):
        neg_items += [i for i in range(I) if not (4 * u <= i < 4 * u + 4)]
        neg_off.append(len(neg_items))
    out = orc.evaluate(P, Q, test_off, test_items, np.array(neg_off, np.int64), np.array(neg_items, np.int32), 4)
    assert out[1] == np.float32(0.625)


def test_bruteforce_cf_known_answers(orc):
    # logics/cf_test.go:26-58: -Dot distance, items k*(1,1,1) k=1..5, query (1,1,1), n=3 -> ids 5,4,3 / 15,12,9
    X = np.array([[k, k, k] for k in range(1, 6)], np.float32)
    idx, sc = orc.bruteforce_search(X, [1, 1, 1], 3, metric=orc.METRIC_NEG_DOT)
    assert (idx + 1).tolist() == [5, 4, 3] and (-sc).tolist() == [15, 12, 9]
    # worker/worker_test.go:194-221: item i = (i, 1), user (1, 0) -> descending i
    X = np.array([[i, 1] for i in range(10)], np.float32)
    idx, sc = orc.bruteforce_search(X, [1, 0], 4, metric=orc.METRIC_NEG_DOT)
    assert idx.tolist() == [9, 8, 7, 6]
    # SearchIndex never returns the query itself (bruteforce.go:47); prune0 drops score <= 0 (:58)
    idx, sc = orc.bruteforce_search(X, X[3], 3, metric=orc.METRIC_EUCLIDEAN, self_index=3)
    assert 3 not in idx.tolist() and sorted(idx.tolist()) == [2, 4, 5] or sorted(idx.tolist()) == [1, 2, 4]
    idx, sc = orc.bruteforce_search(X, [1, 0], 10, prune0=True, metric=orc.METRIC_NEG_DOT)
    assert len(idx) == 0  # all distances are <= 0
