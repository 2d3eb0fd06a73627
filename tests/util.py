"""Shared helpers for the parity tests."""
import numpy as np


def to_bf16_values(a):
    """Round fp32 -> bf16 (round-to-nearest-even) and return the values as fp32."""
    a = np.ascontiguousarray(a, np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(a.shape)


def exact_distance(metric, xq, yrow):
    """fp64 distance of one query / one row under the reference's definitions."""
    xq = xq.astype(np.float64)
    yrow = yrow.astype(np.float64)
    if metric == 0:
        return float(((xq - yrow) ** 2).sum())
    if metric == 1:
        return float((xq * yrow).sum())
    nx, ny = np.sqrt((xq * xq).sum()), np.sqrt((yrow * yrow).sum())
    xn = xq / nx if nx * nx >= np.finfo(np.float32).eps else xq
    yn = yrow / ny if ny * ny >= np.finfo(np.float32).eps else yrow
    return float(1.0 - (xn * yn).sum())


def check_topk(metric, x, y, dis_g, ids_g, dis_o, ids_o, rtol=1e-4, atol=1e-5, min_exact=0.999):
    """The drop-in contract (BASELINE.json north_star): row ids bit-exact, distances within
    1e-4 relative.  Where an id differs, it must be a near-tie swap: the GPU's row has a true
    (fp64) distance within tolerance of the oracle's distance at that rank."""
    assert dis_g.shape == dis_o.shape and ids_g.shape == ids_o.shape
    valid = ids_o >= 0
    assert ((ids_g >= 0) == valid).all(), "filled / unfilled slots differ"
    scale = np.maximum(np.abs(dis_o[valid]), 1.0)
    err = np.abs(dis_g[valid] - dis_o[valid]) / scale
    assert err.size == 0 or err.max() <= rtol + atol, f"distance mismatch: max rel err {err.max():.3e}"
    same = ids_g == ids_o
    frac = same[valid].mean() if valid.any() else 1.0
    bad = np.argwhere(valid & ~same)
    for q, j in bad:
        t = exact_distance(metric, x[q], y[ids_g[q, j]])
        ref = float(dis_o[q, j])
        assert abs(t - ref) <= (rtol + atol) * max(abs(ref), 1.0), (
            f"query {q} rank {j}: gpu id {ids_g[q, j]} (true {t}) vs oracle id {ids_o[q, j]} ({ref})")
    assert frac >= min_exact, f"only {frac:.5f} of ids identical"
    return frac
