"""The NCF-format loader behind `gorse-bench cf` (gorse_b200_ncf_*; dataset.LoadDataFromBuiltIn, dataset/dataset.go:398-490)
and the gorse-bench-cf binary (BASELINE configs[0]: an ml-100k-shaped BPR d=16 fit)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_loader_semantics(gb, tmp_path):
    tr, te = tmp_path / "t.train.rating", tmp_path / "t.test.negative"
    tr.write_text("0\t1\t5\t1\n0\t3\t4\t2\n2\t0\t1\t3\n0\t1\t2\t9\r\n")       # a duplicate (kept, dataset.go:231-240), user 1 has no rows, CRLF
    te.write_text("(0,2)\t4\t5\n(2,3)\t1\t6\n")                                   # negative 6 is an item nobody rated: it joins the dictionary (:485)
    U, I, train, test, neg = gb.load_ncf(tr, te)
    assert (U, I) == (3, 7)
    assert train[0].tolist() == [0, 3, 3, 4] and train[1].tolist() == [1, 3, 1, 0]
    assert test[0].tolist() == [0, 1, 1, 2] and test[1].tolist() == [2, 3]
    assert neg[0].tolist() == [0, 2, 2, 4] and neg[1].tolist() == [4, 5, 1, 6]
    U, I, train, test, neg = gb.load_ncf(tr)                                      # train only
    assert (U, I) == (3, 4) and test[0].tolist() == [0, 0, 0, 0]
    for bad in ("0 1\n", "x\t1\n", "0\t-1\n"):
        tr.write_text(bad)
        with pytest.raises(gb.GorseB200Error):
            gb.load_ncf(tr)
    tr.write_text("0\t1\n")
    te.write_text("0,2\t4\n")                                                     # wrong format: no parentheses (:468-470)
    with pytest.raises(gb.GorseB200Error):
        gb.load_ncf(tr, te)
    with pytest.raises(gb.GorseB200Error):
        gb.load_ncf(tmp_path / "missing")


def test_round_trip_of_a_synthetic_dataset(gb, tmp_path):
    from gorse_b200 import synth

    off, items = synth.make_feedback(300, 120, 4000, seed=2, n_clusters=4)
    train, test = synth.leave_one_out(off, items, seed=1)
    neg = synth.sample_negatives(120, train, test, 20, seed=3)
    a, b = synth.write_ncf(str(tmp_path / "s"), train, test, neg)
    U, I, tr2, te2, ng2 = gb.load_ncf(a, b)
    assert (U, I) == (300, 120)
    for x, y in ((train, tr2), (test, te2), (neg, ng2)):
        assert x[0].tolist() == y[0].tolist() and x[1].tolist() == y[1].tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["bpr", "als"])
def test_gorse_bench_cf_on_an_ml100k_surrogate(gb, orc, tmp_path, model):
    """BASELINE configs[0] through the binary: 943 x 1682 x 100K (the real ml-100k is not available offline), d = 16; the
    table's scores equal a fit through the Python binding of the same ABI with the same seed, and the model learned."""
    from gorse_b200 import synth

    U, I = 943, 1682
    off, items = synth.make_feedback(U, I, 100_000 + U, seed=7, n_clusters=10, exact=True)
    train, test = synth.leave_one_out(off, items, seed=1)
    neg = synth.sample_negatives(I, train, test, 99, seed=2)
    a, b = synth.write_ncf(str(tmp_path / "ml-100k-surrogate"), train, test, neg)
    exe = os.path.join(ROOT, "gorse_b200", "gorse-bench-cf")
    extra = ["--reg", "0.015", "--alpha", "0.05"] if model == "als" else []
    out = subprocess.run([exe, "--train", a, "--test", b, "--model", model, "--factors", "16", "--epochs", "30", "--seed", "5", *extra],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    Uf, If, trf, tef, ngf = gb.load_ncf(a, b)     # the dictionary sizes are what the FILES say (ids never drawn do not exist)
    assert Uf == U and If <= I and trf[0][-1] == 100_000
    assert f"| train | {Uf:6d} | {If:6d} | {100000:13d} |" in out.stdout
    I = If
    row = re.search(r"\| (BPR|ALS)\s+\|\s+([0-9.]+) \|\s+([0-9.]+) \|\s+([0-9.]+) \|\s+(\d+)", out.stdout)
    assert row and row.group(1) == model.upper() and int(row.group(5)) == 30
    ndcg = float(row.group(2))
    ioff, iusers = gb.transpose_csr(train[0], train[1], I)
    kw = dict(n_epochs=30, seed=5)
    if model == "als":
        kw.update(reg=0.015, alpha=0.05)
    with gb.Context(0) as ctx, gb.CFModel(ctx, U, I, 16, train[0], train[1], ioff, iusers) as m:
        res = m.fit(model, test[0], test[1], neg[0], neg[1], **kw)
    assert abs(ndcg - res.ndcg) < 5e-4 + (0.02 if model == "bpr" else 0.0)   # BPR: Hogwild atomics are not bit-reproducible run to run
    assert ndcg > 0.15
