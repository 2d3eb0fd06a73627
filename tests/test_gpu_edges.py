"""Edge cases of the vector path through the C ABI: empty / tiny inputs, k > n, extreme k, ties,
NaN rows, odd dimensions, tile-boundary sizes -- each checked against the oracle."""
import numpy as np
import pytest

import myscaledb_b200 as b2
import oracle as orc
from myscaledb_b200 import search as S
from tests.util import check_topk, to_bf16_values

pytestmark = pytest.mark.gpu
F32 = np.float32


def test_empty_and_tiny_inputs():
    y = np.arange(12, dtype=F32).reshape(3, 4)
    x = np.ones((2, 4), F32)
    dis, ids = b2.flat_knn(b2.L2, x, y, 5)                       # k > n: unfilled slots are -1 / FLT_MAX
    do, io = orc.knn_flat(orc.L2, x, y, 5)
    assert (ids == io).all() and (ids[:, 3:] == -1).all() and np.array_equal(dis, do)
    dis, ids = b2.flat_knn(b2.IP, x, y, 5)
    do, io = orc.knn_flat(orc.IP, x, y, 5)
    assert (ids == io).all() and np.array_equal(dis, do)
    dis, ids = b2.flat_knn(b2.L2, np.zeros((0, 4), F32), y, 3)   # no queries
    assert dis.shape == (0, 3)
    dis, ids = b2.flat_knn(b2.L2, x, np.zeros((0, 4), F32), 3)   # empty part
    assert (ids == -1).all()
    dis, ids = b2.part_scan(b2.IP, x, np.zeros((0, 4), F32), 3)
    assert (ids == -1).all()
    alive = np.zeros(3, bool)                                     # everything filtered out
    dis, ids = b2.flat_knn(b2.L2, x, y, 2, alive_bits=orc.pack_bits(alive))
    assert (ids == -1).all()


@pytest.mark.parametrize("k", [1, 1000, 2048])
def test_extreme_k_on_scan_path(k):
    rng = np.random.default_rng(k)
    y = rng.standard_normal((5000, 24)).astype(F32)
    x = rng.standard_normal((3, 24)).astype(F32)
    dg, ig = b2.flat_knn(b2.L2, x, y, k)
    do, io = orc.knn_flat(orc.L2, x, y, k)
    check_topk(b2.L2, x, y, dg, ig, do, io)


def test_all_rows_identical_ties_go_to_smaller_ids():
    y = np.ones((10000, 16), F32)
    x = np.zeros((2, 16), F32)
    for metric in (b2.L2, b2.IP):
        dis, ids = b2.flat_knn(metric, x + (metric == b2.IP), y, 10)
        assert ids.tolist() == [list(range(10))] * 2
    c = b2.Corpus(b2.IP, 64, dtype=S.BF16).append(np.ones((70000, 64), F32))
    c.set_path(2)
    dis, ids = c.search(np.ones((200, 64), F32), 10)             # every score ties, across CTAs and tiles
    assert (ids == np.arange(10)[None, :]).all() and (dis == 64).all()


def test_nan_and_inf_rows_never_returned_for_l2():
    rng = np.random.default_rng(2)
    y = rng.standard_normal((1500, 8)).astype(F32)
    y[5] = np.nan
    y[7] = np.finfo(F32).max                                      # FLT_MAX padded "empty" row -> inf distance
    x = rng.standard_normal((2, 8)).astype(F32)
    with np.errstate(all="ignore"):
        dg, ig = b2.flat_knn(b2.L2, x, y, 1499)                   # 1498 rows are eligible
        do, io = orc.knn_flat(orc.L2, x, y, 1499)
    assert 5 not in ig and 7 not in ig and (ig[:, -1] == -1).all() and (ig[:, -2] >= 0).all()
    assert (ig == io).mean() > 0.99
    with pytest.raises(b2.B200Error):                             # documented limit of the fused scan top-k
        b2.flat_knn(b2.L2, x, y, 4096)


@pytest.mark.parametrize("d", [1, 2, 5, 63, 100, 2049])
def test_odd_dimensions(d):
    rng = np.random.default_rng(d)
    y = rng.standard_normal((4000, d)).astype(F32)
    x = rng.standard_normal((4, d)).astype(F32)
    for metric in (b2.L2, b2.COSINE):
        dg, ig = b2.flat_knn(metric, x, y, 7)
        do, io = orc.search_without_index(metric, x, y, 7)
        # d <= 2 under cosine: many rows share a direction to within 1e-7 -> near-tie swaps are legitimate
        check_topk(metric, x, y, dg, ig, do, io, rtol=2e-4, atol=2e-5, min_exact=0.99 if d > 2 else 0.8)


@pytest.mark.parametrize("n", [1, 255, 256, 257, 511, 513])
@pytest.mark.parametrize("nq,k", [(16, 5), (129, 100), (300, 1000)])
def test_gemm_tile_boundaries_and_large_k(n, nq, k):
    rng = np.random.default_rng(n + nq)
    y = to_bf16_values(rng.standard_normal((n, 100)).astype(F32))  # d = 100 -> padded to 128
    x = to_bf16_values(rng.standard_normal((nq, 100)).astype(F32))
    c = b2.Corpus(b2.L2, 100, dtype=S.BF16).append(y)
    c.set_path(2)
    dg, ig = c.search(x, k)
    c.close()
    do, io = orc.knn_flat(orc.L2, x, y, k)
    check_topk(b2.L2, x, y, dg, ig, do, io, rtol=3e-4, atol=3e-4, min_exact=0.98)


@pytest.mark.parametrize("nbytes", [1, 3, 33, 128])
def test_binary_odd_sizes(nbytes):
    rng = np.random.default_rng(nbytes)
    y = rng.integers(0, 256, (3000, nbytes), dtype=np.uint8)
    x = rng.integers(0, 256, (3, nbytes), dtype=np.uint8)
    for metric in (b2.HAMMING, b2.JACCARD):
        dg, ig = b2.binary_knn(metric, x, y, 12)
        do, io = orc.knn_binary(metric, x, y, 12)
        assert (ig == io).all() and np.array_equal(dg, do)


def test_residency_cache_holds_real_corpora_and_evicts_by_hbm_bytes():
    """VICacheManager on the device: parts' FLAT corpora cached by CacheKey string, LRU by their HBM bytes."""
    rng = np.random.default_rng(3)
    ya = rng.standard_normal((20000, 64)).astype(F32)
    yb = rng.standard_normal((30000, 64)).astype(F32)
    q = rng.standard_normal((3, 64)).astype(F32)
    S.cache_expire_prefix("gpu/")
    a = b2.Corpus(b2.L2, 64).append(ya)
    b = b2.Corpus(b2.L2, 64).append(yb)
    da, ia = a.search(q, 5)
    S.cache_set_capacity(int(30000 * 64 * 4 * 1.5))      # room for b alone (or a alone), not both
    ha = S.cache_put("gpu/t/part_a/v1", a)
    assert a._h.value is None                            # the cache owns the device object now
    S.cache_release("gpu/t/part_a/v1")
    h, kind = S.cache_get("gpu/t/part_a/v1")
    assert h == ha and kind == S.CACHE_CORPUS
    d2, i2 = b2.Corpus.borrowed(h, b2.L2, 64).search(q, 5)   # search through the cached handle
    assert (i2 == ia).all() and np.array_equal(d2, da)
    S.cache_release("gpu/t/part_a/v1")
    S.cache_put("gpu/t/part_b/v1", b)                    # evicts part_a (LRU, unpinned): its HBM is freed
    with pytest.raises(S.CacheMiss):
        S.cache_get("gpu/t/part_a/v1")
    st = S.cache_stats()
    assert st["used"] >= 30000 * 64 * 4 and st["evictions"] >= 1
    S.cache_release("gpu/t/part_b/v1")
    assert S.cache_expire_prefix("gpu/") == 1
    S.cache_set_capacity(2 ** 63)


def test_search_on_an_empty_resident_corpus_and_thread_scratch_release():
    """A FLAT index created but not yet filled answers with empty slots on every path selection; the per-thread scratch
    of the one-shot entry points can be handed back."""
    q = np.random.default_rng(1).standard_normal((40, 32)).astype(F32)
    for dtype in (S.F32, S.BF16):
        c = b2.Corpus(b2.IP, 32, dtype=dtype)
        dis, ids = c.search(q, 5)
        assert (ids == -1).all()
        c.close()
    y = np.random.default_rng(2).standard_normal((1000, 32)).astype(F32)
    d1, i1 = b2.flat_knn(b2.L2, q[:3], y, 4)
    assert S.lib().b200_thread_release() == 0
    d2, i2 = b2.flat_knn(b2.L2, q[:3], y, 4)       # scratch is rebuilt on demand
    assert (i1 == i2).all() and np.array_equal(d1, d2)
