import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
import gorse_b200 as gb
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
NQ = int(sys.argv[2]) if len(sys.argv) > 2 else 151552
d, k = 128, 100
rng = np.random.default_rng(0)
X = rng.standard_normal((N, d), dtype=np.float32)
X /= np.linalg.norm(X, axis=1, keepdims=True)
with gb.Context(0) as ctx, gb.BruteforceIndex(ctx, d, gb.METRIC_NEG_DOT) as ix:
    t = time.time(); ix.add(X); print("add s", time.time() - t)
    ix.search_range(0, 512, k)  # builds the bf16 mirror, warms up
    for rep in range(2):
        ctx.timer_begin(); t = time.time()
        idx, dist, cnt = ix.search_range(0, NQ, k)
        ms = ctx.timer_end(); wall = time.time() - t
        fb = ix.debug_fallback_rows()
        flop = 2.0 * NQ * N * d
        print(f"N={N} nq={NQ} device {ms:.1f} ms wall {wall*1e3:.1f} ms -> {NQ/ms*1e3/1e6:.3f} M vectors/s, {flop/ms/1e9:.1f} TFLOP/s (algorithmic), fallback rows {fb}, min cnt {cnt.min()}")
