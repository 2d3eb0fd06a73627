"""GPU BM25 parity: reference goldens bit-exact through the C ABI, then random corpora vs the oracle."""
import numpy as np
import pytest

import myscaledb_b200 as b2
import oracle as orc

pytestmark = pytest.mark.gpu
F32 = np.float32


def _pairs(res):
    rows, sc = res
    return [[int(r), float(s)] for r, s in zip(rows, sc)]


def _gold(lst):
    return [[e[0], float(F32(e[1]))] for e in lst]


def test_goldens_00040_00041_bitexact(goldens):
    g = goldens["00040_hybrid"]
    ix = b2.BM25Index(1)
    for rid, _, doc in g["docs"]:
        ix.add_doc(rid, [doc])
    ix.commit()
    assert ix.total_docs == 20 and ix.doc_freq("ancient") == 2 and ix.total_tokens() == 146
    assert _pairs(ix.search("Ancient", 5)) == _gold(g["text_search"])
    alive = np.zeros(20, bool); alive[:10] = True
    assert _pairs(ix.search("Ancient", 5, alive_bits=orc.pack_bits(alive))) == _gold(g["text_search_where_id_lt10"])
    ia = b2.BM25Index(1)
    for rid, pieces in g["array_docs"]:
        ia.add_doc(rid, [pieces])
    ia.commit()
    assert _pairs(ia.search(g["array_query"], 5)) == _gold(g["array_text_search"])
    im = b2.BM25Index(1)
    for rid, key in g["map_docs"]:
        im.add_doc(rid, [[key]])
    im.commit()
    assert _pairs(im.search(g["map_query"], 5)) == _gold(g["map_text_search"])
    # 00041: two parts with table-wide statistics == one part
    p0, p1 = b2.BM25Index(1), b2.BM25Index(1)
    for rid, _, doc in g["docs"]:
        (p0 if rid < 10 else p1).add_doc(rid, [doc])
    p0.commit(); p1.commit()
    stats = dict(total_docs=p0.total_docs + p1.total_docs, total_tokens={0: p0.total_tokens() + p1.total_tokens()},
                 doc_freq={(0, "ancient"): p0.doc_freq("ancient") + p1.doc_freq("ancient")})
    sc, pa, la = [], [], []
    for pi, part in enumerate((p0, p1)):
        rows, s = part.search("Ancient", 5, stats=stats)
        sc += s.tolist(); pa += [pi] * len(s); la += rows.tolist()
    s, p, l = orc.merge_parts(sc, pa, la, 5, desc=True)
    assert [[int(l[j]), float(s[j])] for j in range(len(s))] == _gold(goldens["00041_multi_parts"]["text_2parts"])


def _random_corpus(n_docs, vocab, seed, n_fields=1):
    rng = np.random.default_rng(seed)
    words = [f"w{i}" for i in range(vocab)]
    pz = 1.0 / np.arange(1, vocab + 1) ** 1.1
    pz /= pz.sum()
    docs = []
    for d in range(n_docs):
        fields = []
        for f in range(n_fields):
            ln = 1 + rng.poisson(12)
            fields.append(" ".join(words[i] for i in rng.choice(vocab, size=ln, p=pz)))
        docs.append(fields)
    return docs, words, rng


@pytest.mark.parametrize("operator_or", [True, False])
@pytest.mark.parametrize("n_fields", [1, 2])
def test_random_corpus_matches_oracle(operator_or, n_fields):
    docs, words, rng = _random_corpus(6000, 400, 3 + n_fields, n_fields)
    g, o = b2.BM25Index(n_fields), orc.BM25Index(n_fields)
    for d, fields in enumerate(docs):
        g.add_doc(d * 2 + 1, fields)  # row ids != doc ordinals
        o.add_doc(d * 2 + 1, fields)
    g.commit()
    alive = rng.random(2 * len(docs) + 2) < 0.7
    queries = [" ".join(words[i] for i in rng.integers(0, 60, size=rng.integers(1, 5))) for _ in range(40)]
    queries += ["w1 zzzunknown", "zzzunknown", "W3,w3;w5"]
    flds = tuple(range(n_fields))
    for ab in (None, orc.pack_bits(alive)):
        res = g.search_batch(queries, 15, fields=flds, alive_bits=ab, operator_or=operator_or)
        for qi, qs in enumerate(queries):
            ro, so = o.search(qs, 15, fields=flds, alive=ab, operator_or=operator_or)
            rg, sg = res[qi]
            assert rg.tolist() == ro.tolist(), (qs, operator_or)
            # 1-2 clause sums are bit-exact; longer sums follow the same clause order -> also exact
            assert sg.tolist() == so.tolist(), (qs, operator_or)


def test_large_corpus_property_monotone_and_filter():
    docs, words, rng = _random_corpus(120000, 3000, 9)
    g = b2.BM25Index(1)
    for d, fields in enumerate(docs):
        g.add_doc(d, fields)
    g.commit()
    qs = ["w0 w1 w2", "w10 w500", "w2999"]
    full = g.search_batch(qs, 100)
    for rows, sc in full:
        assert (np.diff(sc) <= 0).all() and len(set(rows.tolist())) == len(rows)
    alive = np.ones(len(docs), bool); alive[::2] = False
    filt = g.search_batch(qs, 100, alive_bits=orc.pack_bits(alive))
    for (rows, sc), (r0, s0) in zip(filt, full):
        assert (rows % 2 == 1).all()
        keep = [(r, s) for r, s in zip(r0.tolist(), s0.tolist()) if r % 2 == 1]
        assert list(zip(rows.tolist(), sc.tolist()))[:len(keep)] == keep  # filtering == masking the unfiltered ranking


def test_tokenizer_unicode_long_tokens_and_multi_field_and_match_the_oracle(tmp_path):
    """Same documents through the GPU index and the oracle: 40-byte tokens dropped, Unicode punctuation separates, non-ASCII
    capitals are lowercased, AND over two columns = every term in at least one column; and a save / load round trip."""
    docs = [["a" * 39 + " " + "b" * 40 + " tail", ""], ["alpha，beta—gamma delta", "ÉCOLE Ünïcode ПРИВЕТ"],
            ["alpha only here", "beta only there"], ["alpha beta together", ""], ["gamma gamma gamma beta", "alpha"]]
    g, o = b2.BM25Index(2), orc.BM25Index(2)
    for i, dd in enumerate(docs):
        g.add_doc(i, dd); o.add_doc(i, dd)
    g.commit()
    for term, f in (("a" * 39, 0), ("b" * 40, 0), ("alpha", 0), ("beta", 1), ("école", 1), ("привет", 1)):
        assert g.doc_freq(term, f) == o.doc_freq(term, f), term
    assert b2.BM25Index.query_terms("Alpha，BETA ÉCOLE") == orc.BM25Index.query_terms("Alpha，BETA ÉCOLE") == ["alpha", "beta", "école"]
    g.save(tmp_path / "t.b2tx")
    g2 = b2.BM25Index.load(tmp_path / "t.b2tx", 2)
    for ix in (g, g2):
        for sentence in ("alpha beta", "gamma alpha", "école alpha", "alpha nosuchterm", "beta"):
            for fields in ((0, 1), (0,), (1,)):
                for op_or in (True, False):
                    rows, sc = ix.search(sentence, 10, fields=fields, operator_or=op_or)
                    er, es = o.search(sentence, 10, fields=fields, operator_or=op_or)
                    assert [int(r) for r in rows] == [int(r) for r in er], (sentence, fields, op_or)
                    assert [float(s) for s in sc] == [float(s) for s in es], (sentence, fields, op_or)
    with pytest.raises(b2.B200Error):
        b2.BM25Index.load(tmp_path / "missing.b2tx", 2)
    raw = (tmp_path / "t.b2tx").read_bytes()
    (tmp_path / "cut.b2tx").write_bytes(raw[: len(raw) - 20])
    with pytest.raises(b2.B200Error):
        b2.BM25Index.load(tmp_path / "cut.b2tx", 2)
