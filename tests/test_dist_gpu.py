"""Two ranks on two GPUs through torchrun: user-sharded BPR with the per-epoch NCCL all-reduce of item deltas
(SURVEY 8e).  Skipped unless the box has >= 2 GPUs.  Checks: both ranks end with bit-identical item tables, the
run learns (NDCG), and a 1-rank run on the same data is close."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import gorse_b200 as gb
from gorse_b200 import synth
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
idbuf = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    idbuf = torch.frombuffer(bytearray(gb.nccl_unique_id()), dtype=torch.uint8).clone()
dist.broadcast(idbuf, 0)
ctx = gb.Context(rank, rank, world, bytes(idbuf.numpy().tobytes()))
U, I, d = 4000, 600, 32
off, items = synth.make_feedback(U, I, 80000, seed=5, n_clusters=8)
train, test = synth.leave_one_out(off, items, seed=1)
m = gb.CFModel(ctx, U, I, d, train[0], train[1])
m.init_normal(0.0, 0.001, 7)
n = int(train[0][-1])
for ep in range(20):
    m.bpr_epoch(0.05, 0.01, n, 100 + ep)
ctx.barrier()
P = np.zeros((U, d), np.float32); Q = np.zeros((I, d), np.float32)
m.get_factors(P, Q)
lo, hi = U * rank // world, U * (rank + 1) // world
np.save(os.path.join(%r, f"P_{rank}.npy"), P[lo:hi]); np.save(os.path.join(%r, f"Q_{rank}.npy"), Q)
m.close(); ctx.close()
dist.barrier(); dist.destroy_process_group()
'''


def test_two_rank_bpr(gb, orc, tmp_path):
    if gb.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from gorse_b200 import synth

    script = tmp_path / "w.py"
    script.write_text(WORKER % (ROOT, str(tmp_path), str(tmp_path)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29577", str(script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    Q0, Q1 = np.load(tmp_path / "Q_0.npy"), np.load(tmp_path / "Q_1.npy")
    assert Q0.tobytes() == Q1.tobytes()  # replicas agree after the all-reduce
    P = np.concatenate([np.load(tmp_path / "P_0.npy"), np.load(tmp_path / "P_1.npy")])
    U, I = 4000, 600
    off, items = synth.make_feedback(U, I, 80000, seed=5, n_clusters=8)
    train, test = synth.leave_one_out(off, items, seed=1)
    neg = synth.sample_negatives(I, train, test, 100, seed=2)
    ndcg2 = orc.evaluate(P, Q0, test[0], test[1], neg[0], neg[1], 10)[0]
    with gb.Context(0) as ctx, gb.CFModel(ctx, U, I, 32, train[0], train[1]) as m:
        m.init_normal(0.0, 0.001, 7)
        for ep in range(20):
            m.bpr_epoch(0.05, 0.01, int(train[0][-1]), 100 + ep)
        ndcg1 = m.evaluate(test[0], test[1], neg[0], neg[1], 10)[0]
    assert ndcg2 > 0.2 and abs(ndcg2 - ndcg1) < 0.03, (ndcg1, ndcg2)


ALS_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import gorse_b200 as gb
from gorse_b200 import synth
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
idbuf = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    idbuf = torch.frombuffer(bytearray(gb.nccl_unique_id()), dtype=torch.uint8).clone()
dist.broadcast(idbuf, 0)
ctx = gb.Context(rank, rank, world, bytes(idbuf.numpy().tobytes()))
U, I, d = 2001, 301, 32
off, items = synth.make_feedback(U, I, 30000, seed=6, zipf_s=1.1)
ioff, iusers = gb.transpose_csr(off, items, I)
rng = np.random.default_rng(2)
P = (rng.standard_normal((U, d)) * 0.1).astype(np.float32)
Q = (rng.standard_normal((I, d)) * 0.1).astype(np.float32)
m = gb.CFModel(ctx, U, I, d, off, items, ioff, iusers)
m.set_factors(P, Q)            # every rank passes the full tables; it keeps its own user rows and all of Q
for ep in range(2):
    m.als_epoch(0.06, 0.001)
ctx.barrier()
P1 = np.zeros((U, d), np.float32); Q1 = np.zeros((I, d), np.float32)
m.get_factors(P1, Q1)
lo, hi = U * rank // world, U * (rank + 1) // world
np.save(os.path.join(%r, f"alsP_{rank}.npy"), P1[lo:hi]); np.save(os.path.join(%r, f"alsQ_{rank}.npy"), Q1)
m.close(); ctx.close()
dist.barrier(); dist.destroy_process_group()
'''


def test_two_rank_als(gb, orc, tmp_path):
    """eALS over 2 ranks (users and items range-sharded, ranges exchanged with grouped NCCL broadcasts):
    the epoch is deterministic, so the result must equal the 1-rank run bit for bit and the oracle within 1e-4."""
    if gb.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from gorse_b200 import synth

    script = tmp_path / "w_als.py"
    script.write_text(ALS_WORKER % (ROOT, str(tmp_path), str(tmp_path)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29578", str(script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    Q0, Q1 = np.load(tmp_path / "alsQ_0.npy"), np.load(tmp_path / "alsQ_1.npy")
    assert Q0.tobytes() == Q1.tobytes()
    P2 = np.concatenate([np.load(tmp_path / "alsP_0.npy"), np.load(tmp_path / "alsP_1.npy")])
    U, I, d = 2001, 301, 32
    off, items = synth.make_feedback(U, I, 30000, seed=6, zipf_s=1.1)
    ioff, iusers = gb.transpose_csr(off, items, I)
    rng = np.random.default_rng(2)
    P = (rng.standard_normal((U, d)) * 0.1).astype(np.float32)
    Q = (rng.standard_normal((I, d)) * 0.1).astype(np.float32)
    with gb.Context(0) as ctx, gb.CFModel(ctx, U, I, d, off, items, ioff, iusers) as m:
        m.set_factors(P, Q)
        for ep in range(2):
            m.als_epoch(0.06, 0.001)
        P1, Q1g = m.get_factors()
    assert P2.tobytes() == P1.tobytes() and Q0.tobytes() == Q1g.tobytes()
    Po, Qo = P.copy(), Q.copy()
    for ep in range(2):
        orc.als_epoch(Po, Qo, off, items, ioff, iusers, 0.06, 0.001)
    rel = lambda a, b: (np.abs(a - b) / np.maximum(np.abs(b).max(axis=1, keepdims=True), 1e-12)).max()  # noqa: E731
    assert rel(P2, Po) < 1e-4 and rel(Q0, Qo) < 1e-4
