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
    # both arms build `config` with the same function: the driver's same_config check compares the dicts
    sys.path.insert(0, ROOT)
    import bench

    assert line["config"] == bench.bpr_config("small", 1) == bench.bpr_config("small", 8)
