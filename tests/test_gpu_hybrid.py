"""HybridSearch on the GPU: vector scan + BM25 + fusion through the C ABI reproduce the reference's
goldens (fed 5 candidates per modality, see SURVEY.md 8c stale-golden caveat) and the oracle on
random lists."""
import numpy as np
import pytest

import myscaledb_b200 as b2
import oracle as orc

pytestmark = pytest.mark.gpu
F32 = np.float32


def _lists_5_gpu(goldens):
    g = goldens["00040_hybrid"]
    ix = b2.BM25Index(1)
    for rid, _, doc in g["docs"]:
        ix.add_doc(rid, [doc])
    ix.commit()
    rows, sc = ix.search("Ancient", 5)
    txt = [(0, 0, int(r), float(s)) for r, s in zip(rows, sc)]
    y = np.repeat(np.arange(20, dtype=np.float32)[:, None], 3, axis=1)
    dis, ids = b2.part_scan(b2.L2, np.array([[1, 1, 1]], F32), y, 5)
    vec = [(0, 0, int(i), float(d)) for i, d in zip(ids[0], dis[0])]
    return vec, txt


def _order(res, k=5):
    return sorted(((r[2], float(F32(r[3]))) for r in res), key=lambda t: (-t[1], t[0]))[:k]


def test_goldens_rsf_rrf_end_to_end_on_gpu(goldens):
    g = goldens["00040_hybrid"]
    vec, txt = _lists_5_gpu(goldens)
    rsf, rrf = (b2.hybrid_fusion_batch(ft, [vec], [txt], 10, fusion_weight=0.5, fusion_k=60, vector_scan_direction=1)[0]
                for ft in ("rsf", "rrf"))
    assert [[i, s] for i, s in _order(rsf)] == [[e[0], float(F32(e[1]))] for e in g["rsf"]]
    assert [[i, s] for i, s in _order(rrf)] == [[e[0], float(F32(e[1]))] for e in g["rrf"]]
    assert [[i, s] for i, s in _order(rsf)] == [[e[0], float(F32(e[1]))] for e in goldens["00041_multi_parts"]["rsf_1part"]]


def test_golden_00040_hybrid_with_lightweight_delete_on_gpu(goldens):
    """DELETE id = 13: statistics unchanged, alive bitmap drops the doc; a part without an FTS index gives the fusion
    an EMPTY text list (00040_mqvs_hybrid_search_with_lwd)."""
    g = goldens["00040_hybrid_with_lwd"]
    ix = b2.BM25Index(1)
    for rid, _, doc in goldens["00040_hybrid"]["docs"]:
        ix.add_doc(rid, [doc])
    ix.commit()
    rows, sc = ix.search(g["query_text"], 1)
    assert [[int(rows[0]), float(sc[0])]] == [[e[0], float(F32(e[1]))] for e in g["text_before_lwd_top1"]]
    alive = np.ones(20, bool); alive[g["deleted_id"]] = False
    rows, sc = ix.search(g["query_text"], 2, alive_bits=orc.pack_bits(alive))
    assert [[int(r), float(x)] for r, x in zip(rows, sc)] == [[e[0], float(F32(e[1]))] for e in g["text_after_lwd_top2"]]
    y = np.repeat(np.arange(20, dtype=np.float32)[:, None], 3, axis=1)
    q = np.array([g["query_vector"]], F32)
    dis, ids = b2.part_scan(b2.L2, q, y, 5)
    vec = [(0, 0, int(i), float(d)) for i, d in zip(ids[0], dis[0])]
    rsf = b2.hybrid_fusion_batch("rsf", [vec], [[]], 5)[0]
    assert [[i, x] for i, x in _order(rsf)] == [[e[0], float(F32(e[1]))] for e in g["rsf_no_text_index"]]
    dis, ids = b2.part_scan(b2.L2, q, y, 5, row_exists=alive.astype(np.uint8))
    vec = [(0, 0, int(i), float(d)) for i, d in zip(ids[0], dis[0])]
    rows, sc = ix.search(g["query_text"], 5, alive_bits=orc.pack_bits(alive))
    txt = [(0, 0, int(r), float(x)) for r, x in zip(rows, sc)]
    rsf = b2.hybrid_fusion_batch("rsf", [vec], [txt], 5)[0]
    assert [[i, x] for i, x in _order(rsf)] == [[e[0], float(F32(e[1]))] for e in g["rsf_after_lwd"]]


@pytest.mark.parametrize("ft", ["rsf", "rrf"])
@pytest.mark.parametrize("direction", [1, -1])
def test_random_lists_match_oracle_bitexact(ft, direction):
    rng = np.random.default_rng(17)
    vecs, txts = [], []
    for q in range(64):
        nv, nt = int(rng.integers(0, 31)), int(rng.integers(0, 31))
        labels = rng.permutation(80)
        vs = np.sort(rng.random(nv).astype(F32))
        if direction == -1:
            vs = vs[::-1]
        if q % 7 == 0 and nv:
            vs[:] = vs[0]  # all-equal scores -> normalised to 1.0
        ts = np.sort(rng.random(nt).astype(F32))[::-1]
        vecs.append([(int(l % 2), int(l % 3), int(l), float(s)) for l, s in zip(labels[:nv], vs)])
        txts.append([(int(l % 2), int(l % 3), int(l), float(s)) for l, s in zip(rng.permutation(80)[:nt], ts)])
    got = b2.hybrid_fusion_batch(ft, vecs, txts, 20, fusion_weight=0.3, fusion_k=60, vector_scan_direction=direction)
    for q in range(64):
        exp = orc.hybrid_fusion(ft, vecs[q], txts[q], 20, fusion_weight=0.3, fusion_k=60, vector_scan_direction=direction) \
            if (vecs[q] or txts[q]) else []
        assert [(a, b, c, float(F32(d))) for a, b, c, d in got[q]] == [(a, b, c, float(F32(d))) for a, b, c, d in exp], q


def test_array_form_of_the_fusion_equals_the_list_form():
    from myscaledb_b200 import search as S
    rng = np.random.default_rng(23)
    nq, kv, kt = 40, 30, 30
    v_ids = np.full((nq, kv), -1, np.int64); t_ids = np.full((nq, kt), -1, np.int64)
    v_sc = np.zeros((nq, kv), F32); t_sc = np.full((nq, kt), -np.inf, F32)
    vecs, txts = [], []
    for q in range(nq):
        nv, nt = int(rng.integers(0, kv + 1)), int(rng.integers(0, kt + 1))
        v_ids[q, :nv] = rng.permutation(100)[:nv]; t_ids[q, :nt] = rng.permutation(100)[:nt]
        v_sc[q, :nv] = np.sort(rng.random(nv).astype(F32)); t_sc[q, :nt] = np.sort(rng.random(nt).astype(F32))[::-1]
        vecs.append([(0, 0, int(i), float(s)) for i, s in zip(v_ids[q, :nv], v_sc[q, :nv])])
        txts.append([(0, 0, int(i), float(s)) for i, s in zip(t_ids[q, :nt], t_sc[q, :nt])])
    for ft in ("rrf", "rsf"):
        want = b2.hybrid_fusion_batch(ft, vecs, txts, 10, fusion_weight=0.4, fusion_k=60, vector_scan_direction=1)
        ids, sc, cnt = S.hybrid_fusion_arrays(ft, v_ids, v_sc, t_ids, t_sc, 10, fusion_weight=0.4, fusion_k=60, vector_scan_direction=1)
        for q in range(nq):
            assert [(int(i), float(x)) for i, x in zip(ids[q, :cnt[q]], sc[q, :cnt[q]])] == [(c, float(F32(d))) for _, _, c, d in want[q]], (ft, q)
            assert (ids[q, cnt[q]:] == -1).all()


def test_reference_call_sites_run_on_gpu():
    """The reference's own call expressions (createVectorIndex, reader-driven build, search with a filter, serialize / load
    through the stream classes, computeTopDistanceSubset, faiss::knn_*, TANTIVY::ffi_*) end to end on the GPU."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "callsite_compile")
    if not os.path.exists(exe):
        pytest.skip("callsite_compile not built (run __graft_entry__.build())")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "CALLSITES OK" in r.stdout, r.stdout + r.stderr
