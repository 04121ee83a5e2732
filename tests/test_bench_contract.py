"""bench.py's host logic with the device mocked out: the control flow of the default BPR leg (device-timed steps, roofline
leg, end-to-end leg with its cleanup paths) runs on the CPU box and the JSON line carries every key of the bench contract.
No number produced here means anything; the point is that a typo in the host code cannot surface only on the GPU box."""
import json
import sys
import types

import numpy as np


class _FakeCtx:
    def __init__(self, *a):
        self._launches = 0
        self.h = None

    def launch_count(self):
        self._launches += 7
        return self._launches

    def timer_begin(self):
        pass

    def timer_end(self):
        return 4.0

    def sync(self):
        pass

    def barrier(self):
        pass

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass


class _FakeModel:
    def __init__(self, ctx, n_users, n_items, d, off, items, *a):
        self.h = 1

    def init_normal(self, *a):
        pass

    def bpr_epoch(self, *a):
        pass

    def close(self):
        self.h = None

    def als_epoch(self, *a):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class _FakePinned:
    def __init__(self, shape, dtype=np.float32):
        self.array = np.zeros(shape, dtype)

    def free(self):
        self.array = None


def _fake_gb(real):
    import importlib

    synth = importlib.import_module(real.__name__ + ".synth")
    m = types.ModuleType("gorse_b200")
    for k in dir(real):
        if not k.startswith("__"):
            setattr(m, k, getattr(real, k))
    m.Context, m.CFModel, m.PinnedArray = _FakeCtx, _FakeModel, _FakePinned
    m.check = lambda st: None
    m.ptr = lambda a: None
    m.lib = types.SimpleNamespace(gorse_b200_cf_set_factors=lambda *a: 0, gorse_b200_cf_get_factors=lambda *a: 0)
    m.synth = synth
    return m


def test_default_bench_leg_with_a_mocked_device(gb, monkeypatch, capsys):
    import importlib

    from gorse_b200 import synth  # noqa: F401  (bench imports it lazily through the package)
    bench = importlib.import_module("bench")
    fake = _fake_gb(gb)
    monkeypatch.setitem(sys.modules, "gorse_b200", fake)
    monkeypatch.setattr(bench, "cpu_baseline", lambda wl: ({"value": 1.0, "unit": "triples/s", "cores": 1, "kind": "reference", "sample": "mock"}, None, None))
    monkeypatch.setattr(bench, "ClockSampler", lambda *a: types.SimpleNamespace(stop=lambda t0, t1: {"sm_mhz": 1.0, "sm_max_mhz": 1.0, "reasons": [], "samples": 1}))
    for var in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(var, raising=False)
    args = bench.build_parser().parse_args(["--workload", "small", "--steps", "2", "--warmup", "1", "--e2e-epochs", "2"])
    bench.run_ours(args)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["config"]["workload"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in line["roofline"], key
    assert line["e2e"] is not None and "e2e_error" not in line
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in line["e2e"], key

    # a failing end-to-end leg must still print the contract line, with the reason
    class _Boom(_FakeModel):
        made = 0

        def __init__(self, *a):
            _Boom.made += 1
            super().__init__(*a)

        def bpr_epoch(self, *a):
            if _Boom.made >= 2:
                raise RuntimeError("boom")

    fake.CFModel = _Boom
    bench.run_ours(args)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["e2e"] is None and "boom" in line["e2e_error"]


class _FakeIndex:
    def __init__(self, ctx, d, metric):
        self.d = d

    def add(self, X):
        self.n = len(X)
        return self.n

    def search_range(self, q0, q1, k, prune0=False, out=None):
        idx, dist, cnt = out if out is not None else (np.zeros((q1 - q0, k), np.int32), np.zeros((q1 - q0, k), np.float32), np.zeros(q1 - q0, np.int32))
        cnt[:] = k
        return idx, dist, cnt

    def debug_stage1(self):
        return 10.0, 1.0e12

    def debug_fallback_rows(self):
        return 3

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass


def test_other_workloads_with_a_mocked_device(gb, monkeypatch, capsys):
    import importlib

    bench = importlib.import_module("bench")
    fake = _fake_gb(gb)
    fake.BruteforceIndex = _FakeIndex
    monkeypatch.setitem(sys.modules, "gorse_b200", fake)
    monkeypatch.setattr(bench, "ClockSampler", lambda *a: types.SimpleNamespace(stop=lambda t0, t1: None))
    for var in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(var, raising=False)
    for argv, metric in ((["--workload", "c4", "--small", "--queries", "2048", "--steps", "2", "--no-cpu"], "top-k"),
                         (["--workload", "c3", "--small", "--steps", "1", "--warmup", "1", "--no-cpu"], "eALS")):
        args = bench.build_parser().parse_args(argv)
        (bench.run_topk if args.workload == "c4" else bench.run_als)(args)
        line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
        assert metric in line["metric"] and line["n_gpus"] == 1 and line["value"] > 0
        for key in ("roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "config", "ms_per_step"):
            assert key in line, key
