"""Host-side helpers of the benchmark and of the synthetic data (no GPU, no library calls)."""
import importlib.util
import os
import stat
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exact_feedback_count_on_a_dense_matrix():
    """ml-100k's shape (943 x 1682, 6 % filled, Zipf head saturated): de-duplication removes far more than the first
    over-draw; exact=True keeps drawing until the requested count exists, every user keeps >= 1 item."""
    from gorse_b200 import synth

    U, I = 943, 1682
    off, items = synth.make_feedback(U, I, 100_000 + U, seed=7, n_clusters=10, exact=True)
    assert off[-1] == 100_000 + U == items.size
    assert np.diff(off).min() >= 1 and items.min() >= 0 and items.max() < I
    for u in (0, 1, 500, U - 1):
        row = items[off[u]:off[u + 1]]
        assert (np.diff(row) > 0).all()            # sorted, no duplicates
    (tr_off, tr_items), (te_off, te_items) = synth.leave_one_out(off, items, seed=1)
    assert tr_off[-1] == 100_000 and te_off[-1] == U
    # a sparse shape is untouched by the extra rounds (same data as before they existed: first round suffices)
    off2, items2 = synth.make_feedback(20_000, 5_000, 400_000, seed=3, exact=True)
    assert off2[-1] == 400_000


def _load_bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_clock_sampler_never_returns_empty(tmp_path, monkeypatch):
    """A fake nvidia-smi that needs 0.25 s for its first line: a timed region shorter than that still gets a sample (the
    nearest one), a warm sampler gets the ones inside the region, throttle reasons are reported."""
    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/bin/bash\nsleep 0.25\nwhile true; do echo '1800, 1965, 900.0, Not Active, Not Active, Not Active, Active'; sleep 0.05; done\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    bench = _load_bench()
    s = bench.ClockSampler(0)
    t0 = time.time()
    time.sleep(0.03)
    c = s.stop(t0, time.time())
    assert c is not None and c["samples"] >= 1 and c["sm_mhz"] == 1800.0 and c["sm_max_mhz"] == 1965.0
    assert c["reasons"] == ["sw_power_cap"]
    s = bench.ClockSampler(0)
    time.sleep(0.5)
    t0 = time.time()
    time.sleep(0.1)
    c = s.stop(t0, time.time())
    assert c["samples"] >= 2
