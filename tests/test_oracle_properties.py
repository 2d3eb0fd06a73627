"""Property tests of the CPU oracle itself (hypothesis): the checker must be right before it checks anything.
Independent restatements in numpy / pure Python on small integer-valued inputs, where ties are everywhere and
fp32 arithmetic is exact, so equality is bit-exact."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle as orc

F32 = np.float32
FMAX = float(np.finfo(np.float32).max)


def _brute(metric, q, y, k, alive=None):
    """(score, then smaller id) top-k in plain numpy on exactly representable integers."""
    q = q.astype(np.float64); y = y.astype(np.float64)
    if metric == orc.L2:
        s = ((y - q) ** 2).sum(1)
        order = np.lexsort((np.arange(len(y)), s))
    else:
        s = (y * q).sum(1)
        order = np.lexsort((np.arange(len(y)), -s))
    if alive is not None:
        order = [i for i in order if alive[i]]
    order = list(order)[:k]
    return order, [s[i] for i in order]


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 120), st.integers(1, 6), st.integers(1, 12), st.integers(1, 40), st.integers(0, 2 ** 31), st.sampled_from([0, 1]))
def test_part_scan_in_blocks_equals_one_pass_with_ties(n, d, k, block_rows, seed, metric):
    """searchWrapper's block-by-block running merge (strict compare, earlier block wins) == one global
    (score, smaller id) selection, for any block size, with heavy ties."""
    rng = np.random.default_rng(seed)
    y = rng.integers(-3, 4, (n, d)).astype(F32)
    q = rng.integers(-3, 4, (1, d)).astype(F32)
    dis, ids = orc.part_scan(metric, q, y, k, block_rows=block_rows)
    exp_ids, exp_s = _brute(metric, q[0], y, k)
    if metric == orc.IP:  # vectorScanWithoutIndex IP quirk: scores <= FLT_MIN never enter
        keep = [j for j, s in enumerate(exp_s) if s > 0]
        exp_ids, exp_s = [exp_ids[j] for j in keep], [exp_s[j] for j in keep]
    got = [int(i) for i in ids[0] if i >= 0]
    assert got == exp_ids
    assert [float(x) for x in dis[0][:len(got)]] == [float(F32(s)) for s in exp_s]


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 150), st.integers(1, 5), st.integers(1, 10), st.integers(0, 2 ** 31), st.floats(0.0, 1.0))
def test_filter_bitmap_equals_search_over_the_surviving_rows(n, d, k, seed, p_alive):
    rng = np.random.default_rng(seed)
    y = rng.integers(-4, 5, (n, d)).astype(F32)
    q = rng.integers(-4, 5, (1, d)).astype(F32)
    alive = rng.random(n) < p_alive
    dis, ids = orc.knn_flat(orc.L2, q, y, k, alive=orc.pack_bits(alive))
    exp_ids, exp_s = _brute(orc.L2, q[0], y, k, alive=alive)
    got = [int(i) for i in ids[0] if i >= 0]
    assert got == exp_ids and all(alive[i] for i in got)
    assert [float(x) for x in dis[0][:len(got)]] == [float(F32(s)) for s in exp_s]
    assert (ids[0][len(got):] == -1).all()


@settings(max_examples=80, deadline=None)
@given(st.lists(st.tuples(st.integers(-5, 5), st.integers(0, 3), st.integers(0, 50)), min_size=0, max_size=60, unique_by=lambda t: (t[1], t[2])),
       st.integers(1, 20), st.booleans())
def test_merge_parts_is_a_stable_multimap_walk(entries, top_k, desc):
    """getTotalTopSearchResultImpl: multimap<score> filled in (part, position) order; ascending walk keeps insertion
    order among equal scores, the descending walk (reverse iterator) yields the LATER-inserted equal key first."""
    entries = sorted(entries, key=lambda t: t[1])  # parts arrive in part_index order
    score = [float(e[0]) for e in entries]; part = [e[1] for e in entries]; label = [e[2] for e in entries]
    s, p, l = orc.merge_parts(score, part, label, top_k, desc=desc)
    idx = list(range(len(entries)))
    exp = sorted(idx, key=lambda i: (score[i], i))
    if desc:
        exp = exp[::-1]
    exp = exp[:top_k]
    assert [(float(a), int(b), int(c)) for a, b, c in zip(s, p, l)] == [(score[i], part[i], label[i]) for i in exp]


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 25), st.integers(0, 25), st.integers(0, 2 ** 31), st.integers(1, 60))
def test_rank_fusion_formula(nv, nt, seed, fusion_k):
    """RankFusion: score[id] += 1 / (fusion_k + rank + 1), ranks 0-based per list, fp32 accumulation."""
    rng = np.random.default_rng(seed)
    vl = rng.permutation(40)[:nv]; tl = rng.permutation(40)[:nt]
    vec = [(0, 0, int(l), float(i)) for i, l in enumerate(vl)]            # ascending distances
    txt = [(0, 0, int(l), float(100 - i)) for i, l in enumerate(tl)]      # descending bm25
    if not vec and not txt:
        return
    got = orc.hybrid_fusion("rrf", vec, txt, 100, fusion_k=fusion_k)
    exp = {}
    for lst in (vec, txt):
        for rank, (_, _, l, _) in enumerate(lst):
            exp[l] = F32(exp.get(l, F32(0)) + F32(1.0) / F32(fusion_k + rank + 1))
    assert {r[2]: F32(r[3]) for r in got} == exp
    assert [r[3] for r in got] == sorted((r[3] for r in got), reverse=True)


@settings(max_examples=40, deadline=None)
@given(st.integers(2, 30), st.integers(0, 2 ** 31))
def test_bm25_two_shards_with_summed_statistics_equal_one_index(n_docs, seed):
    """Scores must not depend on how documents are split: per-shard scoring with table-wide statistics (the sums of
    total_docs / total_tokens / doc_freq) is bit-identical to scoring in one index."""
    rng = np.random.default_rng(seed)
    vocab = [f"w{i}" for i in range(12)]
    docs = [" ".join(rng.choice(vocab, size=int(rng.integers(1, 15)))) for _ in range(n_docs)]
    one = orc.BM25Index(1)
    for i, t in enumerate(docs):
        one.add_doc(i, [t])
    cut = int(rng.integers(1, n_docs))
    a, b = orc.BM25Index(1), orc.BM25Index(1)
    for i, t in enumerate(docs):
        (a if i < cut else b).add_doc(i, [t])
    query = " ".join(rng.choice(vocab, size=2, replace=False))
    terms = orc.BM25Index.query_terms(query)
    stats = {"total_docs": a.total_docs + b.total_docs, "total_tokens": {0: a.total_tokens() + b.total_tokens()},
             "doc_freq": {(0, t): a.doc_freq(t) + b.doc_freq(t) for t in terms}}
    rows1, sc1 = one.search(query, n_docs)
    merged = {}
    for ix in (a, b):
        rows, sc = ix.search(query, n_docs, stats=stats)
        merged.update({int(r): float(s) for r, s in zip(rows, sc)})
    assert merged == {int(r): float(s) for r, s in zip(rows1, sc1)}


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 20), st.integers(1, 20), st.integers(0, 2 ** 31), st.sampled_from([1, -1]), st.floats(0.0, 1.0))
def test_relative_score_fusion_formula(nv, nt, seed, direction, w):
    """RelativeScoreFusion: each list min-max normalised ((s - min) / (max - min), all-equal -> 1.0), then
    text * w + (direction == -1 ? vec : 1 - vec) * (1 - w), fp32 throughout (HybridSearchUtils.cpp:212-314)."""
    rng = np.random.default_rng(seed)
    vs = np.sort(rng.integers(0, 50, nv).astype(F32))
    if direction == -1:
        vs = vs[::-1]
    ts = np.sort(rng.integers(0, 50, nt).astype(F32))[::-1]
    vl = rng.permutation(60)[:nv]; tl = rng.permutation(60)[:nt]
    vec = [(0, 0, int(l), float(s)) for l, s in zip(vl, vs)]
    txt = [(0, 0, int(l), float(s)) for l, s in zip(tl, ts)]
    got = {r[2]: F32(r[3]) for r in orc.hybrid_fusion("rsf", vec, txt, 200, fusion_weight=w, vector_scan_direction=direction)}

    def norm(scores):
        lo, hi = F32(min(scores)), F32(max(scores))
        return [F32(1.0) if hi == lo else F32(F32(s - lo) / F32(hi - lo)) for s in map(F32, scores)]
    w32 = F32(w)
    exp = {}
    for l, s in zip(tl, norm(ts)):
        exp[int(l)] = F32(exp.get(int(l), F32(0)) + F32(s * w32))
    for l, s in zip(vl, norm(vs)):
        v = s if direction == -1 else F32(F32(1.0) - s)
        exp[int(l)] = F32(exp.get(int(l), F32(0)) + F32(v * F32(F32(1.0) - w32)))
    assert set(got) == set(exp)
    for l in exp:
        assert abs(float(got[l]) - float(exp[l])) <= 1.2e-7 * max(1.0, abs(float(exp[l]))), (l, got[l], exp[l])


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 200), st.integers(1, 8), st.integers(1, 15), st.integers(0, 2 ** 31), st.sampled_from([3, 4]))
def test_binary_metrics_against_python_popcounts(n, nbytes, k, seed, metric):
    """Hamming = popcount(x ^ y); Jaccard distance = (|x or y| - |x and y|) / |x or y| (0 for two empty sets);
    ties -> smaller id."""
    rng = np.random.default_rng(seed)
    y = rng.integers(0, 256, (n, nbytes), dtype=np.uint8)
    x = rng.integers(0, 256, (1, nbytes), dtype=np.uint8)
    dis, ids = orc.knn_binary(metric, x, y, k)
    pc = lambda a: int(np.unpackbits(a).sum())
    if metric == orc.HAMMING:
        s = [float(pc(x[0] ^ r)) for r in y]
    else:
        s = [0.0 if pc(x[0] | r) == 0 else float(F32(F32(pc(x[0] | r) - pc(x[0] & r)) / F32(pc(x[0] | r)))) for r in y]
    order = sorted(range(n), key=lambda i: (s[i], i))[:k]
    got = [int(i) for i in ids[0] if i >= 0]
    assert got == order
    assert [float(v) for v in dis[0][:len(got)]] == [s[i] for i in order]
