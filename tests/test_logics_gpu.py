"""QueryItemToItem / QueryUserToUser through the C ABI (logics/item_to_item.go:50-86) against the oracle
(ann.Bruteforce + the score post-processing) and the reference's own known answer."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(gb):
    c = gb.Context(0)
    yield c
    c.close()


def test_reference_embedding_known_answer(gb, ctx):
    # logics/item_to_item_test.go:117-141 (TestEmbedding): item i has embedding i*(0.1, 0.2, 0.3), stored bf16-truncated
    # (item_to_item.go:151-164), Euclidean; neighbours of item "0" with n = 10 are items 1..10 in that order
    X = gb.bf16_truncate(np.array([[np.float32(0.1) * np.float32(i), np.float32(0.2) * np.float32(i), np.float32(0.3) * np.float32(i)]
                                   for i in range(100)], np.float32))
    with gb.BruteforceIndex(ctx, 3, gb.METRIC_EUCLIDEAN) as ix:
        ix.add(X)
        ids, sc, cnt = ix.query_similar(0, 1, 10)
        assert cnt[0] == 10 and ids[0].tolist() == list(range(1, 11))
        assert np.all(np.diff(sc[0]) < 0) and np.all(sc[0] > 0) and np.all(sc[0] < 1)
        want = 1.0 / (1.0 + np.sqrt(((X[1:11].astype(np.float64) - X[0]) ** 2).sum(1)))
        assert np.allclose(sc[0], want, rtol=1e-6)


@pytest.mark.parametrize("metric,scale", [("euclid", 1.0), ("negdot", 1.0), ("negdot", 0.5)])
def test_query_similar_matches_oracle(gb, orc, ctx, metric, scale):
    rng = np.random.default_rng(11)
    N, d, n = 700, 32, 20
    X = gb.bf16_truncate(rng.standard_normal((N, d)).astype(np.float32))
    gm = gb.METRIC_EUCLIDEAN if metric == "euclid" else gb.METRIC_NEG_DOT
    om = orc.METRIC_EUCLIDEAN if metric == "euclid" else orc.METRIC_NEG_DOT
    with gb.BruteforceIndex(ctx, d, gm) as ix:
        ix.add(X)
        ids, sc, cnt = ix.query_similar(100, 164, n, scale)
    # the reference: n+1 neighbours of the stored vector (the query itself included), then the post-processing
    for r, q in enumerate(range(100, 164)):
        oi, od = orc.bruteforce_search(X, X[q], n + 1, metric=om)
        o_ids, o_sc = orc.similar_scores(metric == "euclid", scale, q, n, oi, -od)
        assert cnt[r] == len(o_ids)
        assert ids[r, :cnt[r]].tolist() == o_ids.tolist()
        assert sc[r, :cnt[r]].tobytes() == o_sc.tobytes()
        assert (ids[r, cnt[r]:] == -1).all()
