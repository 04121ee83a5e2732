"""W ranks on W GPUs through torchrun.  Every rank hands the C ABI ITS OWN rows only (users
[U*r/W, U*(r+1)/W) of the user CSR, items [I*r/W, I*(r+1)/W) of the item CSR for eALS).

BPR: user-sharded, per-epoch NCCL exchange of item deltas; checks: replicas end with bit-identical item tables, the
sharded Evaluate (per-rank sums all-reduced) agrees with the oracle's Evaluate of the gathered factors, and NDCG is
within 0.03 of the 1-GPU fit at W = 2, 4 and 8.
eALS: row ranges exchanged with grouped NCCL broadcasts, Grams all-reduced; the result must equal the 1-GPU epoch up to
the reassociation of the Gram sums (1e-5) and the oracle within 1e-4.
Each test is skipped unless the box has >= W GPUs."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PRELUDE = r'''
import os, sys, json
sys.path.insert(0, %r)
OUT = %r
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
def rows(off, idx, lo, hi):
    """this rank's slice of a global CSR: offsets keep their global values, the index pointer starts at off[lo]"""
    return off[lo:hi + 1], idx[off[lo]:off[hi]]
'''

BPR_WORKER = PRELUDE + r'''
U, I, d, EPOCHS = 8000, 600, 32, 20
off, items = synth.make_feedback(U, I, 160000, seed=5, n_clusters=8)
train, test = synth.leave_one_out(off, items, seed=1)
neg = synth.sample_negatives(I, train, test, 100, seed=2)
lo, hi = U * rank // world, U * (rank + 1) // world
m = gb.CFModel(ctx, U, I, d, *rows(train[0], train[1], lo, hi))
res = m.fit("bpr", *rows(test[0], test[1], lo, hi), *rows(neg[0], neg[1], lo, hi), n_epochs=EPOCHS, verbose=10, seed=7)
ndcg_fit = res.ndcg
ndcg_eval = float(m.evaluate(*rows(test[0], test[1], lo, hi), *rows(neg[0], neg[1], lo, hi), 10)[0])
ctx.barrier()
P = np.zeros((U, d), np.float32); Q = np.zeros((I, d), np.float32)
m.get_factors(P, Q)
np.save(os.path.join(OUT, f"P_{rank}.npy"), P[lo:hi]); np.save(os.path.join(OUT, f"Q_{rank}.npy"), Q)
json.dump({"ndcg_fit": ndcg_fit, "ndcg_eval": ndcg_eval, "epochs": res.epochs_run}, open(os.path.join(OUT, f"r_{rank}.json"), "w"))
m.close(); ctx.close()
dist.barrier(); dist.destroy_process_group()
'''


def _torchrun(script, world, port):
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_rank_bpr(gb, orc, tmp_path, world):
    if gb.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from gorse_b200 import synth

    script = tmp_path / "w.py"
    script.write_text(BPR_WORKER % (ROOT, str(tmp_path)))
    _torchrun(script, world, 29570 + world)
    Qs = [np.load(tmp_path / f"Q_{r}.npy") for r in range(world)]
    for r in range(1, world):
        assert Qs[0].tobytes() == Qs[r].tobytes()  # replicas agree after the exchange
    P = np.concatenate([np.load(tmp_path / f"P_{r}.npy") for r in range(world)])
    rs = [json.load(open(tmp_path / f"r_{r}.json")) for r in range(world)]
    U, I, d, EPOCHS = 8000, 600, 32, 20
    off, items = synth.make_feedback(U, I, 160000, seed=5, n_clusters=8)
    train, test = synth.leave_one_out(off, items, seed=1)
    neg = synth.sample_negatives(I, train, test, 100, seed=2)
    ndcg_w = orc.evaluate(P, Qs[0], test[0], test[1], neg[0], neg[1], 10)[0]
    for r in rs:   # the sharded Evaluate: every rank reports the global score (sums all-reduced)
        assert r["epochs"] == EPOCHS and abs(r["ndcg_eval"] - ndcg_w) < 1e-5 and abs(r["ndcg_fit"] - ndcg_w) < 1e-5, (r, ndcg_w)
    with gb.Context(0) as ctx, gb.CFModel(ctx, U, I, d, train[0], train[1]) as m:
        res = m.fit("bpr", test[0], test[1], neg[0], neg[1], n_epochs=EPOCHS, verbose=10, seed=7)
        ndcg1 = res.ndcg
    print(f"NDCG@10 world={world}: {ndcg_w:.4f}   world=1: {ndcg1:.4f}")
    assert ndcg_w > 0.2 and abs(ndcg_w - ndcg1) < 0.03, (ndcg1, ndcg_w)


ALS_WORKER = PRELUDE + r'''
U, I, d = 2001, 301, 32
off, items = synth.make_feedback(U, I, 30000, seed=6, zipf_s=1.1)
ioff, iusers = gb.transpose_csr(off, items, I)
rng = np.random.default_rng(2)
P = (rng.standard_normal((U, d)) * 0.1).astype(np.float32)
Q = (rng.standard_normal((I, d)) * 0.1).astype(np.float32)
lo, hi = U * rank // world, U * (rank + 1) // world
ilo, ihi = I * rank // world, I * (rank + 1) // world
m = gb.CFModel(ctx, U, I, d, *rows(off, items, lo, hi), *rows(ioff, iusers, ilo, ihi))
m.set_factors(P, Q)            # every rank passes the base of the full tables; it keeps its own user rows and all of Q
for ep in range(2):
    m.als_epoch(0.06, 0.001)
ctx.barrier()
P1 = np.zeros((U, d), np.float32); Q1 = np.zeros((I, d), np.float32)
m.get_factors(P1, Q1)
np.save(os.path.join(OUT, f"alsP_{rank}.npy"), P1[lo:hi]); np.save(os.path.join(OUT, f"alsQ_{rank}.npy"), Q1)
m.close(); ctx.close()
dist.barrier(); dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 4])
def test_multi_rank_als(gb, orc, tmp_path, world):
    if gb.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from gorse_b200 import synth

    script = tmp_path / "w_als.py"
    script.write_text(ALS_WORKER % (ROOT, str(tmp_path)))
    _torchrun(script, world, 29580 + world)
    Qs = [np.load(tmp_path / f"alsQ_{r}.npy") for r in range(world)]
    for r in range(1, world):
        assert Qs[0].tobytes() == Qs[r].tobytes()
    P2 = np.concatenate([np.load(tmp_path / f"alsP_{r}.npy") for r in range(world)])
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
    rel = lambda a, b: (np.abs(a - b) / np.maximum(np.abs(b).max(axis=1, keepdims=True), 1e-12)).max()  # noqa: E731
    # the Gram sums are taken per rank and all-reduced: reassociation only
    assert rel(P2, P1) < 1e-5 and rel(Qs[0], Q1g) < 1e-5, (rel(P2, P1), rel(Qs[0], Q1g))
    Po, Qo = P.copy(), Q.copy()
    for ep in range(2):
        orc.als_epoch(Po, Qo, off, items, ioff, iusers, 0.06, 0.001)
    assert rel(P2, Po) < 1e-4 and rel(Qs[0], Qo) < 1e-4
