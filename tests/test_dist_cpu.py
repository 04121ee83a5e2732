"""The N > 1 plumbing of bench.py on CPU (gloo, world_size 2): rendezvous on 127.0.0.1, nccl-id broadcast buffer,
max-over-ranks reduction, per-rank shard construction.  The CUDA/NCCL data path itself runs on the GPU box
(tests/test_dist_gpu.py via torchrun)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import bench
rank, world, local = bench.dist_env()
dist.init_process_group("gloo", rank=rank, world_size=world)
# the 128-byte id travels from rank 0 to everyone
idbuf = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    idbuf = torch.frombuffer(bytearray(range(128)), dtype=torch.uint8).clone()
dist.broadcast(idbuf, 0)
assert bytes(idbuf.numpy().tobytes()) == bytes(range(128))
# weak-scaling shard: a rank holds (and hands to the C ABI) only its own users' rows
bench.WORKLOADS["tiny"] = (300, 50, 2000, 16, "tiny")
n_users, n_items, d, off, items = bench.make_shard("tiny", rank, world)
assert n_users == 300 * world and len(off) == 301 and off[0] == 0 and 300 <= off[-1] == items.size <= 2000
# max over ranks
t = torch.tensor([float(rank + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == world
dist.barrier()
dist.destroy_process_group()
open(os.path.join(%r, f"ok_{rank}"), "w").write("ok")
'''


def test_gloo_world2_plumbing(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % (ROOT, str(tmp_path)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert (tmp_path / "ok_0").exists() and (tmp_path / "ok_1").exists()


def test_reference_arm_prints_contract_line():
    env = dict(os.environ)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr
    import json

    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "triples/s" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["cores"] >= 1


BENCH_WORKER = r'''
import os, sys, types, json, io, contextlib
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
import gorse_b200 as real
import bench
from test_bench_contract import _fake_gb
fake = _fake_gb(real)
fake.nccl_unique_id = lambda: bytes(range(128))
sys.modules["gorse_b200"] = fake
bench.ClockSampler = lambda *a: types.SimpleNamespace(stop=lambda t0, t1: None)
bench.WORKLOADS["tiny"] = (3000, 500, 20000, 16, "tiny")
args = bench.build_parser().parse_args(["--workload", "tiny", "--gpus", "2", "--steps", "2", "--warmup", "1", "--e2e-epochs", "2"])
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    bench.run_ours(args)
rank = int(os.environ["RANK"])
out = buf.getvalue().strip()
if rank == 0:
    line = json.loads(out.splitlines()[-1])
    assert line["n_gpus"] == 2 and line["e2e"] is not None and line["cpu_baseline"] is None and "x 2 ranks" in line["config"]["workload"], line
else:
    assert out == "", out
open(os.path.join(%r, f"bench_ok_{rank}"), "w").write("ok")
'''


def test_bench_two_ranks_with_a_mocked_device(tmp_path):
    """bench.py's N > 1 control flow end to end over gloo (shards, step count agreement, max over ranks, the collective
    verdict of the end-to-end leg, rank 0 alone prints) with the device calls mocked out."""
    script = tmp_path / "bench_worker.py"
    script.write_text(BENCH_WORKER % (ROOT, ROOT, str(tmp_path)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29534", str(script)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert (tmp_path / "bench_ok_0").exists() and (tmp_path / "bench_ok_1").exists()


TOPK_WORKER = r'''
import os, sys, types, json, io, contextlib
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import gorse_b200 as real
import bench
from test_bench_contract import _fake_gb, _FakeIndex
fake = _fake_gb(real)
fake.BruteforceIndex = _FakeIndex
sys.modules["gorse_b200"] = fake
bench.ClockSampler = lambda *a: types.SimpleNamespace(stop=lambda t0, t1: None)
args = bench.build_parser().parse_args(["--workload", "c4", "--small", "--gpus", "2", "--queries", "1024", "--steps", "2", "--no-cpu"])
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    bench.run_topk(args)
rank = int(os.environ["RANK"])
out = buf.getvalue().strip()
if rank == 0:
    line = json.loads(out.splitlines()[-1])
    assert line["n_gpus"] == 2 and abs(line["value"] - 2 * 1024 / (line["ms_per_step"] * 1e-3)) < 1e-6 * line["value"], line
else:
    assert out == "", out
open(os.path.join(%r, f"topk_ok_{rank}"), "w").write("ok")
'''


def test_bench_topk_two_ranks_with_a_mocked_device(tmp_path):
    script = tmp_path / "topk_worker.py"
    script.write_text(TOPK_WORKER % (ROOT, ROOT, str(tmp_path)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29535", str(script)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert (tmp_path / "topk_ok_0").exists() and (tmp_path / "topk_ok_1").exists()
