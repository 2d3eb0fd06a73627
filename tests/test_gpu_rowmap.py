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
