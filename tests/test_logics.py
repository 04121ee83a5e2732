"""Host logic around the neighbour search: dense bf16 embeddings, sqrt(idf) sparse vectors and
the QueryItemToItem / QueryUserToUser score post-processing.  CPU only: the oracle's restatement against values
lifted from the reference, and the library's host functions against the oracle (bit-exact)."""
import numpy as np


def test_oracle_bf16_matches_the_reference_definition(orc):
    # bfloats.FromFloat32 keeps bits 31..16 (common/bfloats/bfloats.go:23-29); ToFloat32 shifts them back (:31-37)
    a = np.array([0.1, 0.2, 0.3, 1.0, -2.7, 3.0e38, 1e-40, 0.0, -0.0], np.float32)
    want = (a.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)
    got = orc.bf16_truncate(a)
    assert got.tobytes() == want.tobytes()
    assert got[:3].tolist() == [0.099609375, 0.19921875, 0.298828125]


def test_oracle_sparse_vector(orc):
    # logics/vector_writer.go:200-209: skip id < 0, id >= len(idf), idf <= 0; value = float32(sqrt(float64(idf)))
    idf = np.array([4.0, 0.0, 2.0, 9.0, -1.0], np.float32)
    ind, val = orc.sparse_vector([0, 3, -1, 9, 2, 1, 4], idf)
    assert ind.tolist() == [0, 3, 2] and val.tolist() == [2.0, 3.0, np.float32(np.sqrt(2.0))]
    # "auto": users appended after tags with offset = len(tagsIDF) (logics/item_to_item.go:238-239)
    ind2, _ = orc.sparse_vector([2, 0], idf, offset=7)
    assert ind2.tolist() == [9, 7]


def test_oracle_similar_scores(orc):
    # logics/item_to_item.go:63-85.  Euclidean: store score = -distance -> 1/(1+dist); own id dropped; stops at n
    ids, sc = orc.similar_scores(True, 1.0, 0, 3, [0, 5, 6, 7, 8], np.float32([0, -1, -2, -3, -4]))
    assert ids.tolist() == [5, 6, 7] and sc.tolist() == [0.5, 1 / 3, 0.25]
    # Dot, type "auto" (x0.5): score <= 0 dropped
    ids, sc = orc.similar_scores(False, 0.5, 0, 3, [5, 0, 6, 7, 8], np.float32([9, 8, 0, -3, 4]))
    assert ids.tolist() == [5, 8] and sc.tolist() == [4.5, 2.0]


def test_library_host_functions_match_the_oracle(gb, orc):
    rng = np.random.default_rng(0)
    a = (rng.standard_normal(5000) * 10.0 ** rng.integers(-20, 20, 5000)).astype(np.float32)
    assert gb.bf16_truncate(a).tobytes() == orc.bf16_truncate(a).tobytes()
    idf = rng.random(300).astype(np.float32) * (rng.random(300) > 0.2)
    ids = rng.integers(-5, 320, 500).astype(np.int32)
    gi, gv = gb.sparse_vector(ids, idf, offset=123)
    oi, ov = orc.sparse_vector(ids, idf, offset=123)
    assert gi.tobytes() == oi.tobytes() and gv.tobytes() == ov.tobytes()
    for metric, euclid in ((gb.METRIC_EUCLIDEAN, True), (gb.METRIC_NEG_DOT, False)):
        for scale in (1.0, 0.5):
            nbr = rng.permutation(50).astype(np.int32)[:21]
            dist = np.sort(rng.standard_normal(21).astype(np.float32) + (1.5 if euclid else 0.0))
            if euclid:
                dist = np.abs(dist)
                dist.sort()
            g_ids, g_sc = gb.similar_scores(metric, scale, int(nbr[3]), 10, nbr, dist)
            o_ids, o_sc = orc.similar_scores(euclid, scale, int(nbr[3]), 10, nbr, -dist)
            assert g_ids.tolist() == o_ids.tolist() and g_sc.tobytes() == o_sc.tobytes()


def _pack(vectors):
    off = np.concatenate([[0], np.cumsum([len(v[0]) for v in vectors])]).astype(np.int64)
    ind = np.concatenate([v[0] for v in vectors]).astype(np.uint32)
    val = np.concatenate([v[1] for v in vectors]).astype(np.float32)
    return off, ind, val


def test_oracle_sparse_search_reference_known_answers(orc):
    """logics/item_to_item_test.go:212-316 (TestTags, TestUsers, TestAuto): idf = 1, item i carries ids 1..100-i;
    the 10 neighbours of item 0 are items 1..10; in "auto" even items carry tags and odd items users (offset
    len(tagsIDF) = 101), so item 0 -> 2,4,..,20 and item 1 -> 3,5,..,21 (score x0.5, item_to_item.go:70-72)."""
    idf = np.ones(101, np.float32)
    vecs = [orc.sparse_vector(list(range(1, 101 - i)), idf) for i in range(100)]          # TestTags == TestUsers
    off, ind, val = _pack(vecs)
    ids, dots = orc.sparse_bruteforce_search(off, ind, val, 0, 11)
    s_ids, s_sc = orc.similar_scores(False, 1.0, 0, 10, ids, dots)
    assert s_ids.tolist() == list(range(1, 11)) and s_sc.tolist() == [float(100 - i) for i in range(1, 11)]
    auto = []
    for i in range(100):
        tags = list(range(1, 101 - i)) if i % 2 == 0 else []
        users = list(range(1, 101 - i)) if i % 2 == 1 else []
        ti, tv = orc.sparse_vector(tags, idf)
        ui, uv = orc.sparse_vector(users, idf, offset=len(idf))
        auto.append((np.concatenate([ti, ui]), np.concatenate([tv, uv])))
    off, ind, val = _pack(auto)
    for q, want in ((0, [2 * i for i in range(1, 11)]), (1, [2 * i + 1 for i in range(1, 11)])):
        ids, dots = orc.sparse_bruteforce_search(off, ind, val, q, 11)
        s_ids, s_sc = orc.similar_scores(False, 0.5, q, 10, ids, dots)
        assert s_ids.tolist() == want
        assert s_sc.tolist() == [0.5 * (100 - j) for j in want]


def test_marshal_latent_factors_is_the_reference_wire_format(gb):
    """gorse_b200_marshal_latent_factors == the factor block of cf.BaseMatrixFactorization.Marshal (model/cf/model.go:212-245):
    int64 LE count + varint-delimited protocol.LatentFactor{id = 1, data = 2 packed} (protocol/encoding.proto:27-30).
    Checked against hand-assembled protobuf bytes and by decoding with an independent varint parser."""
    import ctypes as C

    from gorse_b200 import _lib

    F = np.arange(12, dtype=np.float32).reshape(4, 3) * 0.5 - 1
    ids = [b"u0", b"", b"user-with-a-much-longer-identifier-" + b"x" * 120, b"u3"]
    pred = np.array([1, 1, 1, 0], np.uint8)
    arr = (C.c_char_p * 4)(*ids)
    n = C.c_size_t(0)
    assert _lib.lib.gorse_b200_marshal_latent_factors(gb.ptr(F), 4, 3, gb.ptr(pred), arr, None, 0, C.byref(n)) == 0
    buf = np.zeros(n.value, np.uint8)
    assert _lib.lib.gorse_b200_marshal_latent_factors(gb.ptr(F), 4, 3, gb.ptr(pred), arr, gb.ptr(buf), buf.size, C.byref(n)) == 0
    assert _lib.lib.gorse_b200_marshal_latent_factors(gb.ptr(F), 4, 3, gb.ptr(pred), arr, gb.ptr(buf), buf.size - 1, C.byref(n)) == _lib.ERR_RANGE
    raw = buf.tobytes()

    def varint(b, p):
        v = s = 0
        while True:
            v |= (b[p] & 0x7F) << s
            p += 1
            if b[p - 1] < 0x80:
                return v, p
            s += 7

    assert int.from_bytes(raw[:8], "little") == 3
    p, rows = 8, []
    while p < len(raw):
        ln, p = varint(raw, p)
        msg, p = raw[p:p + ln], p + ln
        q, rid, data = 0, b"", None
        while q < len(msg):
            tag, q = varint(msg, q)
            ln2, q = varint(msg, q)
            if tag == 0x0A:
                rid = msg[q:q + ln2]
            else:
                assert tag == 0x12
                data = np.frombuffer(msg[q:q + ln2], "<f4")
            q += ln2
        rows.append((rid, data))
    assert [r[0] for r in rows] == ids[:3] and all(np.array_equal(rows[i][1], F[i]) for i in range(3))
    # first message by hand: len 18 = (0x0A 2 'u0') + (0x12 12 <3 floats>)
    assert raw[8:12] == bytes([18, 0x0A, 2]) + b"u"[:1] and raw[12:14] == b"0" + bytes([0x12]) and raw[14] == 12
    # empty id: the field is omitted (proto3 default); long id: two-byte varint lengths
    assert raw[8 + 19] == 14 and raw[8 + 20] == 0x12
