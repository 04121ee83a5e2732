"""Helpers shared by the CPU and the full-size GPU checks of the eALS epoch (not a test module)."""
import numpy as np


def als_objective(orc, P, Q, user_off, user_items, w, reg):
    """The quantity one eALS epoch cannot increase when every user has feedback:
         sum_obs [(1 - r)^2 - w r^2]  +  w * sum_{u, i in A_I} r_ui^2  +  reg * (|P|^2 + |Q_{A_I}|^2),
    A_I = items with feedback: model/cf/model.go:645-658 builds S^q from those only, so the user half-sweep minimises exactly
    this in every coordinate, and the item half-sweep (S^p over all users, :693-706) does so for every item of A_I.
    Data part in double (oracle helper, threaded); Grams by sgemm then double."""
    assert (np.diff(user_off) > 0).all(), "monotonicity needs every user to have feedback"
    act = np.bincount(user_items, minlength=Q.shape[0]) > 0
    obs = orc.als_observed_loss(P, Q, user_off, user_items, w)
    Qa = np.ascontiguousarray(Q[act])
    allsq = float(np.sum((P.T @ P).astype(np.float64) * (Qa.T @ Qa).astype(np.float64)))
    return obs + w * allsq + reg * (float(np.square(P, dtype=np.float64).sum()) + float(np.square(Qa, dtype=np.float64).sum()))


def oracle_user_rows(orc, transpose_csr, P0, Q0, user_off, user_items, users, reg, w):
    """Rows `users` of P after the user half-sweep of one epoch, from the oracle, without running the whole data set:
    a reduced problem holding only those users plus one dummy user whose row lists every item that has feedback in the
    FULL data, so that S^q = sum over the same items in the same order (model.go:651).  Returns P rows [len(users), d]."""
    n_items = Q0.shape[0]
    act_items = np.nonzero(np.bincount(user_items, minlength=n_items) > 0)[0].astype(np.int32)
    rows = [user_items[user_off[u]:user_off[u + 1]] for u in users] + [act_items]
    off = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    items = np.concatenate(rows).astype(np.int32)
    ioff, iusers = transpose_csr(off, items, n_items)
    P = np.concatenate([P0[users], np.zeros((1, P0.shape[1]), np.float32)]).astype(np.float32)
    Q = Q0.copy()
    orc.als_epoch(P, Q, off, items, ioff, iusers, reg, w)
    return P[:len(users)]
