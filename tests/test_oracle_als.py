"""The oracle's eALS epoch (C) against a second, independent transcription of model/cf/model.go:641-738 in pure
Python with one np.float32 rounding per Go operator, on a small case: bit-identical.  (Go itself cannot run here, so
this guards the transcription, not the Go compiler: SURVEY App. A explains why every operator rounds separately.)"""
import numpy as np

f32 = np.float32


def als_epoch_py(orc, P, Q, user_fb, item_fb, reg, w):
    U, d = P.shape
    n_items = Q.shape[0]
    reg, w, one = f32(reg), f32(w), f32(1)

    def gram(X, fb):                       # :645-658 / :693-706
        s = np.zeros((d, d), f32)
        for r in range(X.shape[0]):
            if len(fb[r]) > 0:
                for i in range(d):
                    for j in range(d):
                        s[i, j] = f32(s[i, j] + f32(X[r, i] * X[r, j]))
        return s

    def rows(X, Y, fb, s):                 # :659-687 / :707-735
        for r in range(X.shape[0]):
            pred = {i: orc.dot(X[r], Y[i]) for i in fb[r]}      # internalPredict = floats.Dot (:195-203)
            for f in range(d):
                res = {i: f32(pred[i] - f32(X[r, f] * Y[i, f])) for i in fb[r]}
                a = b = c = f32(0)
                for i in fb[r]:
                    a = f32(a + f32(f32(one - f32(f32(one - w) * res[i])) * Y[i, f]))
                    c = f32(c + f32(f32(f32(one - w) * Y[i, f]) * Y[i, f]))
                for k in range(d):
                    if k != f:
                        b = f32(b + f32(f32(w * X[r, k]) * s[k, f]))
                X[r, f] = f32(f32(a - b) / f32(f32(c + f32(w * s[f, f])) + reg))
                for i in fb[r]:
                    pred[i] = f32(res[i] + f32(X[r, f] * Y[i, f]))

    rows(P, Q, user_fb, gram(Q, item_fb))
    rows(Q, P, item_fb, gram(P, user_fb))
    assert n_items == len(item_fb) and U == len(user_fb)


def test_oracle_als_epoch_matches_line_by_line_transcription(orc):
    import gorse_b200 as gb  # host-only helpers (CSR transpose); no device call

    rng = np.random.default_rng(5)
    U, I, d = 14, 9, 16
    user_fb = [sorted(rng.choice(I, size=rng.integers(0, 6), replace=False).tolist()) for _ in range(U)]
    user_fb[3] = []                                              # a user without feedback is still visited
    off = np.concatenate([[0], np.cumsum([len(r) for r in user_fb])]).astype(np.int64)
    items = np.array([i for r in user_fb for i in r], np.int32)
    ioff, iusers = gb.transpose_csr(off, items, I)
    item_fb = [iusers[ioff[i]:ioff[i + 1]].tolist() for i in range(I)]
    P = (rng.standard_normal((U, d)) * 0.1).astype(np.float32)
    Q = (rng.standard_normal((I, d)) * 0.1).astype(np.float32)
    Po, Qo = P.copy(), Q.copy()
    with np.errstate(all="ignore"):
        for _ in range(2):
            als_epoch_py(orc, P, Q, user_fb, item_fb, 0.06, 0.001)
            orc.als_epoch(Po, Qo, off, items, ioff, iusers, 0.06, 0.001)
            assert P.tobytes() == Po.tobytes() and Q.tobytes() == Qo.tobytes()


def test_objective_is_monotone_and_reduced_problem_reproduces_rows(orc):
    """Validates, on the oracle, the two checks the full-size GPU test (tests/test_xl_als_gpu.py) relies on."""
    import gorse_b200 as gb
    from gorse_b200 import synth

    from als_checks import als_objective, oracle_user_rows

    U, I, d = 500, 180, 32
    off, items = synth.make_feedback(U, I, 6000, seed=4, zipf_s=1.2)
    assert (np.bincount(items, minlength=I) == 0).any()          # some items without feedback, as at C3
    ioff, iusers = gb.transpose_csr(off, items, I)
    rng = np.random.default_rng(1)
    P = (rng.standard_normal((U, d)) * 0.1).astype(np.float32)
    Q = (rng.standard_normal((I, d)) * 0.1).astype(np.float32)
    P0, Q0 = P.copy(), Q.copy()
    losses = [als_objective(orc, P, Q, off, items, 0.001, 0.06)]
    for ep in range(3):
        orc.als_epoch(P, Q, off, items, ioff, iusers, 0.06, 0.001)
        losses.append(als_objective(orc, P, Q, off, items, 0.001, 0.06))
        if ep == 0:
            users = np.array([0, 7, 123, 499, int(np.diff(off).argmax())])
            rows = oracle_user_rows(orc, gb.transpose_csr, P0, Q0, off, items, users, 0.06, 0.001)
            assert rows.tobytes() == P[users].tobytes()
    assert all(b < a for a, b in zip(losses, losses[1:])), losses
