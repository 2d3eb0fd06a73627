"""Pin the CPU oracle against the reference's own golden outputs.

Every expected value below comes from a .reference file of the reference's SQL test
suite (tests/golden/reference_goldens.json, produced by tests/golden/make_golden.py).
Float goldens are printed by ClickHouse with shortest-round-trip formatting, so
`np.float32(x) == np.float32(golden)` is a bit-exact check where asserted with ==.
"""
import numpy as np
import pytest

import oracle as orc

F32 = np.float32


def _corpus_nnn(lo, hi, d=3):
    return np.repeat(np.arange(lo, hi, dtype=np.float32)[:, None], d, axis=1)


def test_00001_flat_l2_bitexact(goldens):
    g = goldens["00001_flat_l2"]
    y = _corpus_nnn(0, 100)
    dis, ids = orc.knn_flat(orc.L2, np.array([g["query"]], F32), y, g["k"])
    for exp in (g["expect"], g["expect_after_reload"]):
        assert ids[0].tolist() == [e[0] for e in exp]
        assert dis[0].tolist() == [float(F32(e[1])) for e in exp]


def test_00012_bruteforce_with_empty_rows_bitexact(goldens):
    g = goldens["00012_bruteforce_l2"]
    # helper 00000_prepare_index_2.sh: ids 0..9 = [n]*3, ids 10..29 empty vectors, ids 30..10029 = [n]*3
    y = _corpus_nnn(0, 10030)
    y[10:30] = np.finfo(np.float32).max  # empty rows padded FLT_MAX (MergeTreeVSManager.cpp:1380)
    with np.errstate(over="ignore"):
        dis, ids = orc.part_scan(orc.L2, np.array([g["query"]], F32), y, g["k"], block_rows=128)
    assert ids[0].tolist() == [e[0] for e in g["expect"]]
    assert dis[0].tolist() == [float(F32(e[1])) for e in g["expect"]]


@pytest.mark.parametrize("name,metric", [("00002_batch_l2", orc.L2), ("00002_batch_ip", orc.IP)])
def test_00002_batch_two_parts(goldens, name, metric):
    g = goldens[name]
    q = np.array(g["queries"], F32)
    parts = [_corpus_nnn(0, 50), _corpus_nnn(50, 100)]
    k = g["k"]
    got = []
    for qi in range(len(q)):
        sc, pa, la = [], [], []
        for pi, y in enumerate(parts):
            d, i = orc.part_scan(metric, q[qi:qi + 1], y, k)
            ok = i[0] >= 0
            sc += d[0][ok].tolist(); pa += [pi] * int(ok.sum()); la += i[0][ok].tolist()
        s, p, l = orc.merge_parts(sc, pa, la, k, desc=(metric == orc.IP))
        got += [[int(p[j]) * 50 + int(l[j]), qi, float(s[j])] for j in range(len(s))]
    exp = [[e[0], e[1], float(F32(e[2]))] for e in g["expect"]]
    assert got == exp


def test_00014_cosine_bruteforce_bitexact(goldens):
    g = goldens["00014_cosine_bruteforce"]
    n = np.arange(1000, dtype=np.float32)
    y = np.stack([n, n + 3, n + 1], axis=1)
    dis, ids = orc.part_scan(orc.COSINE, np.array([g["query"]], F32), y, g["k"], block_rows=1024)
    assert ids[0].tolist() == [e[0] for e in g["expect"]]
    np.testing.assert_allclose(dis[0], [e[1] for e in g["expect"]], rtol=0, atol=6e-8)


def _bin_corpus():
    return np.repeat((np.arange(1024) % 256).astype(np.uint8)[:, None], 4, axis=1)


@pytest.mark.parametrize("metric,prefix", [(orc.HAMMING, "hamming"), (orc.JACCARD, "jaccard")])
def test_00038_binary(goldens, metric, prefix):
    g = goldens["00038_binary"]
    y = _bin_corpus()
    q = np.array([g["query"]], np.uint8)
    dis, ids = orc.part_scan(metric, q, y, 20)
    exp = g[prefix + "_brute"]
    assert ids[0].tolist() == [e[0] for e in exp]
    assert dis[0].tolist() == [float(F32(e[1])) for e in exp]
    # WHERE id > 100 and id < 120 -> prefilter bitmap
    mask = np.zeros(1024, bool); mask[101:120] = True
    dis, ids = orc.part_scan(metric, q, y, 20, filter_bits=orc.pack_bits(mask))
    exp = g[prefix + "_filter"]
    assert ids[0][:len(exp)].tolist() == [e[0] for e in exp]
    assert (ids[0][len(exp):] == -1).all()
    assert dis[0][:len(exp)].tolist() == [float(F32(e[1])) for e in exp]
    # batch
    bq = np.array(g["batch_queries"], np.uint8)
    dis, ids = orc.part_scan(metric, bq, y, 10)
    got = [[int(ids[qi, j]), qi, float(dis[qi, j])] for qi in range(3) for j in range(10)]
    assert got == [[e[0], e[1], float(F32(e[2]))] for e in g[prefix + "_batch"]]


def test_00038_hamming_after_lightweight_delete(goldens):
    g = goldens["00038_binary"]
    y = _bin_corpus()
    row_exists = np.ones(1024, np.uint8); row_exists[:200] = 0
    dis, ids = orc.part_scan(orc.HAMMING, np.array([g["query"]], np.uint8), y, 10, row_exists=row_exists)
    exp = g["hamming_after_lwd_lt200"]
    assert ids[0].tolist() == [e[0] for e in exp]
    assert dis[0].tolist() == [e[1] for e in exp]


def _mstg_corpus():
    n = np.arange(1000, dtype=np.float64)[:, None]
    x = np.arange(768, dtype=np.float64)[None, :]
    sign = np.where(x % 2 == 0, -1.0, 1.0)
    return (0.00001 * (n * 768 + x + 1) * sign).astype(np.float32)


def test_00028_mstg_768_exact_small_n(goldens):
    """ANN index goldens at N=1000 equal the exact answer; distances within 1e-4 rel
    (fp32 reduction order differs, SURVEY.md 7 'hard parts')."""
    g = goldens["00028_mstg_768"]
    y = _mstg_corpus()
    q = np.array([g["query"]], F32)
    dis, ids = orc.search_without_index(orc.L2, q, y, 5)
    assert ids[0].tolist() == [e[0] for e in g["expect_l2"]]
    np.testing.assert_allclose(dis[0], [e[1] for e in g["expect_l2"]], rtol=1e-6)
    dis, ids = orc.search_without_index(orc.COSINE, q, y, 5)
    assert ids[0].tolist() == [e[0] for e in g["expect_cosine"]]
    np.testing.assert_allclose(dis[0], [e[1] for e in g["expect_cosine"]], rtol=1e-6)
    alive = np.ones(1000, bool); alive[0] = False
    dis, ids = orc.search_without_index(orc.COSINE, q, y, 5, alive=orc.pack_bits(alive))
    assert ids[0].tolist() == [e[0] for e in g["expect_cosine_where_id_gt0"]]
    alive[2] = False
    dis, ids = orc.search_without_index(orc.COSINE, q, y, 5, alive=orc.pack_bits(alive))
    assert ids[0].tolist() == [e[0] for e in g["expect_cosine_after_delete_id2"]]


def test_00035_ties_smaller_id_first(goldens):
    g = goldens["00035_ties"]
    idv = np.array([n for n in range(1001) if n != 1])
    y = np.repeat(idv.astype(np.float32)[:, None], 16, axis=1)
    q = np.array([g["query"]], F32)
    dis, ids = orc.part_scan(orc.L2, q, y, 10)
    assert [[int(idv[i]), float(d)] for i, d in zip(ids[0], dis[0])] == g["expect_unfiltered"]
    dis, ids = orc.part_scan(orc.L2, q, y, 10, filter_bits=orc.pack_bits(idv < 11))
    assert [[int(idv[i]), float(d)] for i, d in zip(ids[0], dis[0])] == g["expect_filtered"]


# ----------------------------------------------------------------------------- BM25
def _index20(goldens, ids=None):
    ix = orc.BM25Index(1)
    for rid, _, doc in goldens["00040_hybrid"]["docs"]:
        if ids is None or rid in ids:
            ix.add_doc(rid, [doc])
    return ix


def test_fieldnorm_table():
    assert [orc.id_to_fieldnorm(i) for i in (0, 39, 40, 41, 47, 48, 56, 64, 72, 80, 88)] == \
        [0, 39, 40, 42, 54, 56, 88, 152, 280, 536, 1048]
    assert orc.fieldnorm_to_id(7) == 7 and orc.fieldnorm_to_id(41) == 40 and orc.fieldnorm_to_id(43) == 41
    assert orc.id_to_fieldnorm(255) == 2013265944  # tantivy FIELD_NORMS_TABLE last entry


def test_00040_bm25_bitexact(goldens):
    g = goldens["00040_hybrid"]
    ix = _index20(goldens)
    assert ix.total_docs == 20 and ix.doc_freq("ancient") == 2
    rows, sc = ix.search(g["query_text"], 5)
    assert [[int(r), float(s)] for r, s in zip(rows, sc)] == [[e[0], float(F32(e[1]))] for e in g["text_search"]]
    alive = np.zeros(20, bool); alive[:10] = True
    rows, sc = ix.search(g["query_text"], 5, alive=orc.pack_bits(alive))
    assert [[int(r), float(s)] for r, s in zip(rows, sc)] == [[e[0], float(F32(e[1]))] for e in g["text_search_where_id_lt10"]]


def test_00040_bm25_array_and_map_columns_bitexact(goldens):
    g = goldens["00040_hybrid"]
    ix = orc.BM25Index(1)
    for rid, pieces in g["array_docs"]:
        ix.add_doc(rid, [pieces])
    rows, sc = ix.search(g["array_query"], 5)
    assert [[int(r), float(s)] for r, s in zip(rows, sc)] == [[e[0], float(F32(e[1]))] for e in g["array_text_search"]]
    ix = orc.BM25Index(1)
    for rid, key in g["map_docs"]:
        ix.add_doc(rid, [[key]])
    rows, sc = ix.search(g["map_query"], 5)
    assert [[int(r), float(s)] for r, s in zip(rows, sc)] == [[e[0], float(F32(e[1]))] for e in g["map_text_search"]]


def test_00041_two_parts_global_stats_equal_one_part(goldens):
    g = goldens["00041_multi_parts"]
    p0, p1 = _index20(goldens, set(range(10))), _index20(goldens, set(range(10, 20)))
    # getStatisticForTextSearch: sum over parts (BM25InfoInDataParts.cpp:40-94)
    stats = dict(total_docs=p0.total_docs + p1.total_docs,
                 total_tokens={0: p0.total_tokens() + p1.total_tokens()},
                 doc_freq={(0, "ancient"): p0.doc_freq("ancient") + p1.doc_freq("ancient")})
    sc, pa, la = [], [], []
    for pi, ix in enumerate((p0, p1)):
        rows, s = ix.search("Ancient", 5, stats=stats)
        sc += s.tolist(); pa += [pi] * len(s); la += rows.tolist()
    s, p, l = orc.merge_parts(sc, pa, la, 5, desc=True)
    got = [[int(l[j]), float(s[j])] for j in range(len(s))]
    assert got == [[e[0], float(F32(e[1]))] for e in g["text_2parts"]] == [[e[0], float(F32(e[1]))] for e in g["text_1part"]]


# --------------------------------------------------------------------------- fusion
def _lists_5(goldens):
    """5 candidates per modality (the goldens predate num_candidates = 3 x LIMIT; SURVEY.md 8c caveat)."""
    ix = _index20(goldens)
    rows, sc = ix.search("Ancient", 5)
    txt = [(0, 0, int(r), float(s)) for r, s in zip(rows, sc)]
    y = _corpus_nnn(0, 20)
    dis, ids = orc.part_scan(orc.L2, np.array([[1, 1, 1]], F32), y, 5)
    vec = [(0, 0, int(i), float(d)) for i, d in zip(ids[0], dis[0])]
    return vec, txt


def _order_by_score_desc_id(res, k=5):
    return sorted(((r[2], r[3]) for r in res), key=lambda t: (-t[1], t[0]))[:k]


def test_00040_rsf_rrf_formula_known_answers(goldens):
    g = goldens["00040_hybrid"]
    vec, txt = _lists_5(goldens)
    assert [v[2] for v in vec] == [1, 0, 2, 3, 4] and [v[3] for v in vec] == [0.0, 3.0, 3.0, 12.0, 27.0]
    rsf = orc.hybrid_fusion("rsf", vec, txt, 10, fusion_weight=0.5, vector_scan_direction=1)
    assert [[i, float(F32(s))] for i, s in _order_by_score_desc_id(rsf)] == [[e[0], float(F32(e[1]))] for e in g["rsf"]]
    rrf = orc.hybrid_fusion("rrf", vec, txt, 10, fusion_k=60)
    assert [[i, float(F32(s))] for i, s in _order_by_score_desc_id(rrf)] == [[e[0], float(F32(e[1]))] for e in g["rrf"]]


def test_00041_rsf_one_part(goldens):
    g = goldens["00041_multi_parts"]
    vec, txt = _lists_5(goldens)
    rsf = orc.hybrid_fusion("rsf", vec, txt, 10)
    assert [[i, float(F32(s))] for i, s in _order_by_score_desc_id(rsf)] == [[e[0], float(F32(e[1]))] for e in g["rsf_1part"]]


# ------------------------------------------------------- timed baseline == checker
@pytest.mark.parametrize("metric", [orc.L2, orc.IP])
def test_cpu_baseline_matches_checker(metric):
    rng = np.random.default_rng(7)
    x = rng.standard_normal((37, 96)).astype(F32)
    y = rng.standard_normal((5003, 96)).astype(F32)
    d0, i0 = orc.knn_flat(metric, x, y, 10)
    d1, i1 = orc.knn_flat_parts(metric, x, y, 10, n_parts=3)
    assert (i0 == i1).mean() > 0.999
    np.testing.assert_allclose(d0, d1, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("metric", [orc.L2, orc.IP])
def test_cpu_baseline_blas_matches_checker(metric):
    rng = np.random.default_rng(8)
    x = rng.standard_normal((45, 96)).astype(F32)
    y = rng.standard_normal((7003, 96)).astype(F32)
    d0, i0 = orc.knn_flat(metric, x, y, 10)
    res = orc.knn_flat_parts_blas(metric, x, y, 10, n_parts=3)
    if res is None:
        pytest.skip("no OpenBLAS next to numpy")
    d1, i1 = res
    assert (i0 == i1).mean() > 0.999
    np.testing.assert_allclose(d0, d1, rtol=2e-4, atol=2e-4)


# ---------------------------------------------------------------- filters, empty rows, lightweight delete (more goldens)
def _helper2_corpus():
    """helpers/00000_prepare_index_2.sh: ids 0..9 = [n]*3, ids 10..29 empty vectors, ids 30..10029 = [n]*3."""
    y = _corpus_nnn(0, 10030)
    y[10:30] = np.finfo(np.float32).max  # empty rows padded FLT_MAX (MergeTreeVSManager.cpp:1380)
    return y


def _check_exact(dis, ids, exp, k):
    n = len(exp)
    assert ids[0, :n].tolist() == [e[0] for e in exp]
    assert dis[0, :n].tolist() == [float(F32(e[1])) for e in exp]
    assert (ids[0, n:k] == -1).all()


def test_00003_prewhere_filter_bitmap_bitexact(goldens):
    g = goldens["00003_prewhere"]
    idv = np.arange(100)
    bits = orc.pack_bits((idv < 10) | (idv > 60))
    dis, ids = orc.part_scan(orc.L2, np.array([g["query"]], F32), _corpus_nnn(0, 100), g["k"], block_rows=1024, filter_bits=bits)
    _check_exact(dis, ids, g["expect"], g["k"])
    dis, ids = orc.knn_flat(orc.L2, np.array([g["query"]], F32), _corpus_nnn(0, 100), g["k"], alive=bits)  # FLAT index + filter
    _check_exact(dis, ids, g["expect"], g["k"])


def test_00008_empty_vectors_bitexact(goldens):
    g = goldens["00008_empty_vectors"]
    y = _corpus_nnn(0, 430)
    y[10:30] = np.finfo(np.float32).max
    with np.errstate(over="ignore"):
        dis, ids = orc.part_scan(orc.L2, np.array([g["query"]], F32), y, g["k"], block_rows=1024)
    _check_exact(dis, ids, g["expect_flat"], g["k"])
    # an index never receives the empty rows (VIPartReader skips them): same answer from the rows that exist
    keep = np.r_[0:10, 30:430]
    dis, pos = orc.knn_flat(orc.L2, np.array([g["query"]], F32), y[keep], g["k"])
    assert keep[pos[0]].tolist() == [e[0] for e in g["expect_ivfflat"]]
    assert dis[0].tolist() == [float(F32(e[1])) for e in g["expect_ivfflat"]]


@pytest.mark.parametrize("name", ["00009_bruteforce_prewhere", "00011_bruteforce_prewhere_sparse"])
def test_00009_00011_bruteforce_with_prewhere_bitexact(goldens, name):
    g = goldens[name]
    idv = np.arange(10030)
    if name.startswith("00009"):
        m = (idv > 5000) | np.isin(idv, [9, 31, 999, 1])
    else:
        m = (idv < 50) | np.isin(idv, [51, 55, 99, 100, 9999])
    with np.errstate(over="ignore"):
        dis, ids = orc.part_scan(orc.L2, np.array([g["query"]], F32), _helper2_corpus(), g["k"], block_rows=128,
                                 filter_bits=orc.pack_bits(m))
    _check_exact(dis, ids, g["expect"], g["k"])


def test_00016_lightweight_delete_bitexact(goldens):
    g = goldens["00016_lightweight_delete"]
    row_exists = np.ones(2100, np.uint8)
    row_exists[2] = 0
    dis, ids = orc.part_scan(orc.L2, np.array([g["query"]], F32), _corpus_nnn(0, 2100), g["k"], block_rows=1024, row_exists=row_exists)
    _check_exact(dis, ids, g["expect"], g["k"])


def test_00040_hybrid_with_lightweight_delete_bitexact(goldens):
    """DELETE id = 13: the document stays in the segment (statistics unchanged), only the alive bitmap drops it;
    a part without an FTS index contributes an EMPTY text list to the fusion."""
    g = goldens["00040_hybrid_with_lwd"]
    ix = _index20(goldens)
    rows, sc = ix.search(g["query_text"], 1)
    assert [[int(rows[0]), float(sc[0])]] == [[e[0], float(F32(e[1]))] for e in g["text_before_lwd_top1"]]
    alive = np.ones(20, bool); alive[g["deleted_id"]] = False
    rows, sc = ix.search(g["query_text"], 2, alive=orc.pack_bits(alive))
    assert [[int(r), float(x)] for r, x in zip(rows, sc)] == [[e[0], float(F32(e[1]))] for e in g["text_after_lwd_top2"]]
    # vector side: brute force over [n,n,n], num_candidates = 5 (as in the other 00040 goldens)
    y = _corpus_nnn(0, 20)
    q = np.array([g["query_vector"]], F32)
    dis, ids = orc.part_scan(orc.L2, q, y, 5)
    vec = [(0, 0, int(i), float(d)) for i, d in zip(ids[0], dis[0])]
    rsf = orc.hybrid_fusion("rsf", vec, [], 5)
    assert [[i, float(F32(x))] for i, x in _order_by_score_desc_id(rsf)] == [[e[0], float(F32(e[1]))] for e in g["rsf_no_text_index"]]
    row_exists = alive.astype(np.uint8)
    dis, ids = orc.part_scan(orc.L2, q, y, 5, row_exists=row_exists)
    vec = [(0, 0, int(i), float(d)) for i, d in zip(ids[0], dis[0])]
    rows, sc = ix.search(g["query_text"], 5, alive=orc.pack_bits(alive))
    txt = [(0, 0, int(r), float(x)) for r, x in zip(rows, sc)]
    rsf = orc.hybrid_fusion("rsf", vec, txt, 5)
    assert [[i, float(F32(x))] for i, x in _order_by_score_desc_id(rsf)] == [[e[0], float(F32(e[1]))] for e in g["rsf_after_lwd"]]


def test_00029_fallback_to_flat_cosine_small_distances(goldens):
    """Cosine distances of 1e-4 are a difference of two numbers next to 1.0: ids must match exactly, values to the last
    ulp or two of 1.0 (the closed library's summation order inside an 8-float row is not known)."""
    g = goldens["00029_fallback_to_flat_cosine"]
    n = np.arange(1000, dtype=F32)[:, None]
    y = n + np.array([0, 7, 6, 5, 4, 3, 2, 1], F32)[None, :]
    dis, ids = orc.search_without_index(orc.COSINE, np.array([g["query"]], F32), y, g["k"])
    for exp in (g["expect"], g["expect_after_reload"]):
        assert ids[0].tolist() == [e[0] for e in exp]
        np.testing.assert_allclose(dis[0], [e[1] for e in exp], rtol=0, atol=2.4e-7)


def test_bm25_tokenizer_and_multi_field_and_semantics():
    """(1) RemoveLongFilter::limit(40) keeps tokens of fewer than 40 bytes; (2) SimpleTokenizer splits on every
    non-alphanumeric CHAR, so Unicode punctuation / spaces separate tokens and LowerCaser lowercases non-ASCII letters;
    (3) AND over several columns = every term found in at least one column (tantivy QueryParser), not every (column, term)."""
    ix = orc.BM25Index(2)
    ix.add_doc(0, ["a" * 39 + " " + "b" * 40 + " tail", ""])
    ix.add_doc(1, ["alpha，beta—gamma delta", "ÉCOLE Ünïcode ПРИВЕТ"])
    ix.add_doc(2, ["alpha only here", "beta only there"])
    ix.add_doc(3, ["alpha beta together", ""])
    assert ix.doc_freq("a" * 39, 0) == 1 and ix.doc_freq("b" * 40, 0) == 0 and ix.doc_len(0, 0) == 2
    for t in ("alpha", "beta", "gamma", "delta"):
        assert ix.doc_freq(t, 0) >= 1, t
    assert ix.doc_len(1, 0) == 4
    for t in ("école", "ünïcode", "привет"):
        assert ix.doc_freq(t, 1) == 1, t
    assert orc.BM25Index.query_terms("Alpha，BETA") == ["alpha", "beta"]
    rows, _ = ix.search("alpha beta", 10, fields=(0, 1), operator_or=False)
    assert sorted(int(r) for r in rows) == [1, 2, 3]      # doc 2 has alpha in column 0 and beta in column 1
    rows, _ = ix.search("alpha beta", 10, fields=(0,), operator_or=False)
    assert sorted(int(r) for r in rows) == [1, 3]
    rows, _ = ix.search("alpha nosuchterm", 10, fields=(0, 1), operator_or=False)
    assert len(rows) == 0
    rows, _ = ix.search("alpha nosuchterm", 10, fields=(0, 1), operator_or=True)
    assert sorted(int(r) for r in rows) == [1, 2, 3]


def test_simd_small_batch_cpu_arm_agrees_with_the_checker():
    """cpu_baseline.c::orc_knn_flat_simd (the timed cfg-1 CPU arm) returns what vs_oracle.c returns."""
    rng = np.random.default_rng(8)
    y = rng.standard_normal((10000, 128)).astype(np.float32)
    x = rng.standard_normal((3, 128)).astype(np.float32)
    for metric in (orc.L2, orc.IP):
        d0, i0 = orc.knn_flat(metric, x, y, 10)
        d1, i1 = orc.knn_flat_simd(metric, x, y, 10)
        assert (i0 == i1).all()
        np.testing.assert_allclose(d1, d0, rtol=2e-6, atol=1e-5)
