"""CPU simulation (oracle epochs) of the multi-GPU item-factor exchange rules: W ranks each run one BPR epoch on
their own user shard against a replica of Q, then the replicas are combined.

    sum    Q <- Q0 + sum_r (Q_r - Q0)                      (unstable once a row moves most of the way to its local
                                                           fixed point within one epoch: error x (1 - W*a))
    mean   Q <- Q0 + mean_r (Q_r - Q0)                     (stable, but cold rows learn W x slower)
    damped Q <- Q0 + s_i * sum_r (Q_r - Q0),  s_i = (1 - prod_r (1-a_ir)) / sum_r a_ir,
           a_ir = 1 - (1 - lr*kappa_r)^{n_ir}: what fraction of the way rank r's n_ir updates carry row i
           (kappa_r = reg + 2 E|p|^2/d; n_ir = expected updates of item i on rank r per epoch) -- the rule in csrc/bpr.cu;
           the extra argument scales the curvature term (1 = shipped, 0.125 = isotropic estimate without safety factor)

usage: python tools/dist_rule_sim.py [world] [epochs]
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import oracle  # noqa: E402
from gorse_b200 import synth  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
EPOCHS = int(sys.argv[2]) if len(sys.argv) > 2 else 40
U, I, F, D = 20000, 2000, 400000, 32
LR, REG = 0.05, 0.01

shards = []
for r in range(W):
    off, items = synth.make_feedback(U, I, F, seed=100 + r, zipf_s=1.0, n_clusters=8)
    tr, te = synth.leave_one_out(off, items, seed=r)
    shards.append((tr, te))


def expected_updates(tr, n_steps):
    off, items = tr
    deg = np.diff(off)
    act = np.nonzero(deg > 0)[0]
    w = np.repeat(1.0 / np.maximum(deg, 1), deg) / len(act)
    pos = np.bincount(items, weights=w, minlength=I)
    return n_steps * (pos + 1.0 / I)


def auc(P, Q, tr, te, rng):
    toff, titems = te
    us = np.nonzero(np.diff(toff) > 0)[0][:4000]
    pos = titems[toff[us]]
    neg = rng.integers(0, I, len(us))
    return float(np.mean(np.einsum("ij,ij->i", P[us], Q[pos]) > np.einsum("ij,ij->i", P[us], Q[neg])))


def run(rule, kappa_mult=1.0):
    rng = np.random.default_rng(0)
    Q = (rng.standard_normal((I, D)) * 0.001).astype(np.float32)
    Ps = [(np.random.default_rng(10 + r).standard_normal((U, D)) * 0.001).astype(np.float32) for r in range(W)]
    acts = [np.nonzero(np.diff(s[0][0]) > 0)[0].astype(np.int32) for s in shards]
    steps = [len(s[0][1]) for s in shards]
    n_ir = np.stack([expected_updates(shards[r][0], steps[r]) for r in range(W)])
    hist = []
    for ep in range(EPOCHS):
        deltas = []
        for r in range(W):
            Qr = Q.copy()
            oracle.bpr_epoch_threads(Ps[r], Qr, shards[r][0][0], shards[r][0][1], acts[r], 1000 + ep, steps[r], LR, REG, 8)
            deltas.append(Qr - Q)
        S = np.sum(deltas, axis=0, dtype=np.float64)
        if rule == "sum":
            s = np.ones(I)
        elif rule == "mean":
            s = np.full(I, 1.0 / W)
        else:
            # "damped" = what csrc/bpr.cu does (xchg_prepare_kernel / xchg_scale): kappa from the rank's own user factors
            p2 = np.array([np.mean(np.einsum("ij,ij->i", P_, P_)) for P_ in Ps])
            kappa = REG + kappa_mult * 2.0 * p2 / D
            L = n_ir * np.log1p(-np.minimum(LR * kappa, 0.5))[:, None]
            a = -np.expm1(L)
            tot = a.sum(0)
            s = np.where(tot > 1e-12, np.minimum(1.0, -np.expm1(L.sum(0)) / np.maximum(tot, 1e-12)), 1.0)
        Q = (Q + s[:, None] * S).astype(np.float32)
        if not np.isfinite(Q).all():
            hist.append((ep, "NaN"))
            break
        if ep % 5 == 4 or ep == EPOCHS - 1:
            a_ = np.mean([auc(Ps[r], Q, *shards[r], np.random.default_rng(5)) for r in range(W)])
            hist.append((ep, round(a_, 4), round(float(np.abs(Q).max()), 3)))
    return hist


if __name__ == "__main__":
    # the single-replica run every rule approximates: the W shards processed one after another on ONE Q
    def seq():
        rng = np.random.default_rng(0)
        Q = (rng.standard_normal((I, D)) * 0.001).astype(np.float32)
        Ps = [(np.random.default_rng(10 + r).standard_normal((U, D)) * 0.001).astype(np.float32) for r in range(W)]
        acts = [np.nonzero(np.diff(s_[0][0]) > 0)[0].astype(np.int32) for s_ in shards]
        h = []
        for ep in range(EPOCHS):
            for r in range(W):
                oracle.bpr_epoch_threads(Ps[r], Q, shards[r][0][0], shards[r][0][1], acts[r], 1000 + ep, len(shards[r][0][1]), LR, REG, 8)
            if ep % 5 == 4 or ep == EPOCHS - 1:
                h.append((ep, round(float(np.mean([auc(Ps[r], Q, *shards[r], np.random.default_rng(5)) for r in range(W)])), 4), round(float(np.abs(Q).max()), 3),
                          round(float(np.mean(np.einsum("ij,ij->i", Ps[0], Ps[0]))), 3)))
        return h
    print("sequential", seq(), flush=True)
    for rule, km in (("sum", 1), ("mean", 1), ("damped", 1), ("damped", 0.125), ("damped", 4)):
        t = time.time()
        print(rule, km, run(rule, km), f"{time.time() - t:.0f}s", flush=True)
