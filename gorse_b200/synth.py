"""Synthetic implicit-feedback data of the shapes BASELINE.json names (no datasets are available offline).

numpy only; used by tests/ and bench.py.  Rows are sorted and de-duplicated; every user has >= 1 item.
"""
import numpy as np


def zipf_weights(n, s=1.0):
    w = 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), s)
    return w / w.sum()


def make_feedback(n_users, n_items, n_feedback, seed=0, zipf_s=1.0, n_clusters=0, in_cluster=0.8, exact=False):
    """Returns (user_off int64[U+1], user_items int32[|R|]) with |R| <= n_feedback after de-duplication
    (exact=True over-draws and trims so that |R| == n_feedback whenever enough distinct pairs exist).

    Item popularity is Zipf(zipf_s) (uniform when zipf_s == 0).  With n_clusters > 0 a planted block
    structure is added: a user in cluster c draws `in_cluster` of its items from cluster c's items.
    """
    rng = np.random.default_rng(seed)
    want = n_feedback
    if exact:
        n_feedback = int(n_feedback * 1.12) + 16
    extra = max(0, n_feedback - n_users)
    act = rng.lognormal(0.0, 1.0, n_users)
    deg = 1 + rng.multinomial(extra, act / act.sum())
    users = np.repeat(np.arange(n_users, dtype=np.int64), deg)
    total = users.size
    # popularity rank -> item id through a fixed permutation so that hot items are spread over the table
    perm = rng.permutation(n_items)
    cdf = np.cumsum(zipf_weights(n_items, zipf_s))
    ranks = np.minimum(np.searchsorted(cdf, rng.random(total)), n_items - 1)
    if n_clusters > 0:
        # restrict the rank to the user's cluster for most draws: rank r belongs to cluster r % n_clusters
        uc = users % n_clusters
        inside = rng.random(total) < in_cluster
        r_in = (ranks // n_clusters) * n_clusters + uc
        r_in = np.where(r_in >= n_items, uc % n_items, r_in)
        ranks = np.where(inside, r_in, ranks)
        items = ranks  # keep rank == id so the planted structure is item % n_clusters
    else:
        items = perm[ranks]
    key = users * np.int64(n_items) + items.astype(np.int64)
    key.sort()  # (np.unique's hash path is ~6x slower at 10^7 keys)
    if key.size > 1:
        key = key[np.concatenate(([True], key[1:] != key[:-1]))]
    # a dense matrix (ml-100k's shape: 6 % filled, Zipf head saturated) loses more than the 12 % over-draw to duplicates:
    # keep drawing (same user activity, same item law) until enough distinct pairs exist
    for _ in range(64 if exact else 0):
        if key.size >= want or key.size >= n_users * n_items:
            break
        m = max(want - key.size, 1024) * 2
        u2 = rng.choice(n_users, size=m, p=act / act.sum()).astype(np.int64)
        r2 = np.minimum(np.searchsorted(cdf, rng.random(m)), n_items - 1)
        if n_clusters > 0:
            uc2 = u2 % n_clusters
            rin2 = (r2 // n_clusters) * n_clusters + uc2
            rin2 = np.where(rin2 >= n_items, uc2 % n_items, rin2)
            i2 = np.where(rng.random(m) < in_cluster, rin2, r2)
        else:
            i2 = perm[r2]
        key = np.unique(np.concatenate((key, u2 * np.int64(n_items) + i2.astype(np.int64))))
    if exact and key.size > want:
        # trim the surplus at random, never a user's first item (every user keeps >= 1)
        u_of = key // n_items
        first = np.ones(key.size, bool)
        first[1:] = u_of[1:] != u_of[:-1]
        cand = np.nonzero(~first)[0]
        k = min(key.size - want, cand.size)
        r = rng.random(cand.size)
        thr = np.partition(r, k - 1)[k - 1]
        keep = np.ones(key.size, bool)
        keep[cand[r <= thr]] = False
        key = key[keep]
    users = (key // n_items).astype(np.int64)
    items = (key % n_items).astype(np.int32)
    off = np.zeros(n_users + 1, np.int64)
    np.cumsum(np.bincount(users, minlength=n_users), out=off[1:])
    return off, items


def leave_one_out(user_off, user_items, seed=0):
    """dataset.SplitCF(0, seed) (dataset/dataset.go:258-319): one random item of every user goes to the test
    split (Go's math/rand stream is not reproducible; the distribution is)."""
    rng = np.random.default_rng(seed)
    U = len(user_off) - 1
    deg = np.diff(user_off)
    pick = user_off[:-1] + np.floor(rng.random(U) * np.maximum(deg, 1)).astype(np.int64)
    has = deg > 0
    mask = np.ones(user_items.size, bool)
    mask[pick[has]] = False
    test_off = np.zeros(U + 1, np.int64)
    np.cumsum(has.astype(np.int64), out=test_off[1:])
    test_items = user_items[pick[has]].astype(np.int32)
    train_items = user_items[mask]
    train_off = np.zeros(U + 1, np.int64)
    np.cumsum(deg - has.astype(np.int64), out=train_off[1:])
    return (train_off, train_items), (test_off, test_items)


def sample_negatives(n_items, train, test, n_candidates, seed=0):
    """dataset.SampleUserNegatives (dataset/dataset.go:242-256 -> util.SampleInt32, common/util/random.go:108-132):
    per user, n_candidates distinct items outside train(u) and test(u) (all of them, ascending, if fewer remain)."""
    rng = np.random.default_rng(seed)
    (tr_off, tr_items), (te_off, te_items) = train, test
    U = len(tr_off) - 1
    neg_off = np.zeros(U + 1, np.int64)
    out = []
    for u in range(U):
        excl = set(tr_items[tr_off[u]:tr_off[u + 1]].tolist()) | set(te_items[te_off[u]:te_off[u + 1]].tolist())
        if n_candidates >= n_items - len(excl):
            s = [i for i in range(n_items) if i not in excl]
        else:
            s = []
            while len(s) < n_candidates:
                v = int(rng.integers(0, n_items))
                if v not in excl:
                    s.append(v)
                    excl.add(v)
        out.append(np.asarray(s, np.int32))
        neg_off[u + 1] = neg_off[u] + len(s)
    return neg_off, (np.concatenate(out) if out else np.zeros(0, np.int32))


def conflict_free_triples(user_off, user_items, n_items, n, seed=0):
    """n triples (u, i, j) with i in R_u, j not in R_u, and no row shared between two triples."""
    rng = np.random.default_rng(seed)
    U = len(user_off) - 1
    users = rng.permutation(np.nonzero(np.diff(user_off) > 0)[0])
    used = set()
    out = []
    for u in users:
        if len(out) == n:
            break
        row = user_items[user_off[u]:user_off[u + 1]]
        cand_i = [int(x) for x in row if int(x) not in used]
        if not cand_i:
            continue
        i = cand_i[int(rng.integers(0, len(cand_i)))]
        rowset = set(int(x) for x in row)
        j = -1
        for _ in range(64):
            c = int(rng.integers(0, n_items))
            if c not in rowset and c not in used and c != i:
                j = c
                break
        if j < 0:
            continue
        used.add(i)
        used.add(j)
        out.append((int(u), i, j))
    return np.asarray(out, np.int32).reshape(-1, 3)


def write_ncf(prefix, train, test, neg):
    """Write (train, test, negatives) CSRs as the NCF files dataset.LoadDataFromBuiltIn reads (dataset/dataset.go:398-490):
    `<prefix>.train.rating` ("user<TAB>item<TAB>rating<TAB>timestamp") and `<prefix>.test.negative`
    ("(user,item)<TAB>neg<TAB>neg...")."""
    (tr_off, tr_items), (te_off, te_items), (ng_off, ng_items) = train, test, neg
    with open(prefix + ".train.rating", "w") as f:
        for u in range(len(tr_off) - 1):
            for i in tr_items[tr_off[u]:tr_off[u + 1]]:
                f.write(f"{u}\t{int(i)}\t1\t0\n")
    with open(prefix + ".test.negative", "w") as f:
        for u in range(len(te_off) - 1):
            for i in te_items[te_off[u]:te_off[u + 1]]:
                f.write(f"({u},{int(i)})" + "".join(f"\t{int(x)}" for x in ng_items[ng_off[u]:ng_off[u + 1]]) + "\n")
    return prefix + ".train.rating", prefix + ".test.negative"
