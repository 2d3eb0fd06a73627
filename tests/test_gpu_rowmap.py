"""Bitmap algebra and decoupled-part row-id remaps: bit-exact vs a numpy restatement of
VIWithDataPart.cpp:56-126 / VIUtils.cpp:479-497 (integer / byte work)."""
import numpy as np
import pytest

from myscaledb_b200 import search as S

pytestmark = pytest.mark.gpu


def test_bitmap_and_and_real_bitmap_and_remaps():
    rng = np.random.default_rng(0)
    for nbits in (1, 7, 64, 1000, 1_234_567):
        a = rng.integers(0, 256, (nbits + 7) // 8, dtype=np.uint8)
        b = rng.integers(0, 256, (nbits + 7) // 8, dtype=np.uint8)
        assert (S.bitmap_and(a, b, nbits) == (a & b)).all()
    # a merged (decoupled) part of 50k rows built from 3 old parts
    n_new, total_vec, own = 50_000, 20_000, 1
    src = rng.integers(0, 3, n_new).astype(np.uint8)
    inv_ids = np.zeros(n_new, np.uint64)
    for s in range(3):
        m = src == s
        inv_ids[m] = rng.permutation(total_vec)[:m.sum()] if m.sum() <= total_vec else rng.integers(0, total_vec, m.sum())
    filt = rng.random(n_new) < 0.3
    got = S.real_bitmap(np.packbits(filt, bitorder="little"), n_new, inv_ids, src, own, total_vec)
    exp = np.zeros(total_vec, bool)
    exp[inv_ids[filt & (src == own)].astype(np.int64)] = True          # VIUtils.cpp:488-494
    assert (np.unpackbits(got, bitorder="little")[:total_vec].astype(bool) == exp).all()
    # transferToNewRowIds
    row_ids_map = rng.permutation(60_000)[:total_vec].astype(np.uint64)
    labels = rng.integers(-1, total_vec, 300).astype(np.int64)
    out = S.remap_labels(row_ids_map, labels)
    assert (out == np.where(labels == -1, -1, row_ids_map[np.maximum(labels, 0)].astype(np.int64))).all()
    # TransferToOldRowIds
    new_ids = rng.integers(-1, n_new + 5, 200).astype(np.int64)
    new_dis = rng.random(200).astype(np.float32)
    o_i, o_d = S.transfer_to_old_row_ids(new_ids, new_dis, inv_ids, src, own)
    keep = [(int(inv_ids[i]), float(d)) for i, d in zip(new_ids, new_dis) if i != -1 and i < n_new and src[i] == own]
    assert list(zip(o_i.tolist(), [float(x) for x in o_d])) == keep


def test_filter_bitmaps_built_on_device_feed_the_device_searches():
    """the PREWHERE result (surviving _part_offset values) and the lightweight-delete mask become a DenseBitmap IN HBM
    (bit-exact vs numpy packbits), are intersected there, and drive b200_corpus_search_device / b200_index_search_device without any
    per-call bitmap upload; results equal the host-bitmap path."""
    import ctypes as C
    torch = pytest.importorskip("torch")
    from myscaledb_b200._lib import lib
    import myscaledb_b200 as b2
    import oracle as orc
    from myscaledb_b200.search import _check
    rng = np.random.default_rng(9)
    n, d = 50_007, 64
    y = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal((5, d)).astype(np.float32)
    keep = np.sort(rng.choice(n, size=n // 3, replace=False)).astype(np.uint64)
    row_exists = (rng.random(n) < 0.9).astype(np.uint8)
    nbytes = ((n + 7) // 8 + 3) // 4 * 4
    d_off = torch.tensor(keep.astype(np.int64)).cuda()
    d_re = torch.tensor(row_exists).cuda()
    b_f = torch.zeros(nbytes, dtype=torch.uint8, device="cuda"); b_d = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    b_and = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    _check(lib().b200_bitmap_from_offsets_device(C.c_void_p(d_off.data_ptr()), C.c_int64(len(keep)), C.c_int64(n), C.c_void_p(b_f.data_ptr()), C.c_void_p(s)))
    _check(lib().b200_bitmap_from_row_exists_device(C.c_void_p(d_re.data_ptr()), C.c_int64(n), C.c_void_p(b_d.data_ptr()), C.c_void_p(s)))
    _check(lib().b200_bitmap_and_device(C.c_void_p(b_f.data_ptr()), C.c_void_p(b_d.data_ptr()), C.c_int64(n), C.c_void_p(b_and.data_ptr()), C.c_void_p(s)))
    torch.cuda.synchronize()
    mask_f = np.zeros(n, bool); mask_f[keep.astype(np.int64)] = True
    want = orc.pack_bits(mask_f & (row_exists != 0))
    got = b_and.cpu().numpy()[:len(want)]
    assert np.array_equal(got, want)
    c = b2.Corpus(b2.L2, d).append(y)
    dh, ih = c.search(q, 10, alive_bits=want)
    tq = torch.tensor(q).cuda(); od = torch.empty((5, 10), device="cuda"); oi = torch.empty((5, 10), dtype=torch.int64, device="cuda")
    c.search_device(tq.data_ptr(), 5, 10, od.data_ptr(), oi.data_ptr(), alive_ptr=b_and.data_ptr(), stream=s)
    torch.cuda.synchronize()
    assert np.array_equal(oi.cpu().numpy(), ih) and np.allclose(od.cpu().numpy(), dh)
    c.close()
