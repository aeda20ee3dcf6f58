"""The conservative bound behind the half-precision box rounds of k_repulse_pairs<V, true> (wembed_b200/csrc/kernels.cuh), restated
in numpy with real float16 arithmetic: a child box may only be pruned if NO point of the fp32 box can be within the interaction
radius of the query.  This pins the derivation (directed rounding of the box, rounding displacement of the query, margin for
the half-precision sum); the kernel itself is checked on the GPU by tests/test_gpu_edge_cases.py (both box formats must agree
bit for bit)."""
import numpy as np
import pytest

F16 = np.float16


def _round_down_f16(a):
    with np.errstate(over="ignore"):
        h = a.astype(F16)
        up = h.astype(np.float64) > a
        h[up] = np.nextafter(h[up], F16(-np.inf))       # +inf -> 65504, like __float2half_rd
    return h


def _round_up_f16(a):
    with np.errstate(over="ignore"):
        h = a.astype(F16)
        dn = h.astype(np.float64) < a
        h[dn] = np.nextafter(h[dn], F16(np.inf))
    return h


def half_box_test(lo, hi, q, centre, iw_q, bound, L=1.0):
    """The box-round test, operation for operation: returns True if the child is KEPT.  lo/hi/q: (cases, d) float32."""
    d = lo.shape[1]
    hv = (d + 7) // 8                                         # 16-byte chunks of 8 halves
    margin_root = np.float32(1.0 + (4 * hv + 6) * 4.9e-4 + 1.0e-6)
    prune_l = np.float32(np.sqrt(np.float32(L * L) * np.float32(1.0 + 1e-5)))
    lo_h = _round_down_f16((lo.astype(np.float64) - centre))   # __fsub_rd + __float2half_rd: never above the exact difference
    hi_h = _round_up_f16((hi.astype(np.float64) - centre))
    qc = (q - centre.astype(np.float32)).astype(np.float32)    # fp32 subtraction, round to nearest
    q_h = qc.astype(F16)
    with np.errstate(over="ignore", invalid="ignore"):
        delta = np.sqrt(((qc - q_h.astype(np.float32)) ** 2).sum(1, dtype=np.float32)) * np.float32(1.001) * margin_root
        delta = np.where(delta >= 0, delta, np.inf).astype(np.float32)
        g1 = (lo_h - q_h).astype(F16)
        g2 = (q_h - hi_h).astype(F16)
        e = np.fmax(np.fmax(g1, g2), F16(0))                   # hmax2 drops NaN operands
        acc = np.zeros((len(lo), 2), F16)
        for k in range(d):                                     # alternating accumulators; product and sum rounded separately (worse than FMA)
            acc[:, k & 1] = ((e[:, k] * e[:, k]).astype(F16) + acc[:, k & 1]).astype(F16)
        total = (acc[:, 0] + acc[:, 1]).astype(F16).astype(np.float32)
        factor = (prune_l * margin_root * np.float32(1.000001) * np.nextafter((1.0 / iw_q).astype(np.float32), np.float32(np.inf)))
        inv_bound = np.nextafter((1.0 / bound).astype(np.float32), np.float32(np.inf))
        thr = (factor * inv_bound + delta).astype(np.float32)
        lim = thr * thr
        return (total <= lim) | (lim >= 6.0e4)


def exact_within(lo, hi, q, iw_q, bound, L=1.0):
    gap = np.maximum(0.0, np.maximum(lo.astype(np.float64) - q, q.astype(np.float64) - hi))
    return np.sqrt((gap * gap).sum(1)) * iw_q * bound <= L


@pytest.mark.parametrize("d", [1, 2, 4, 8, 11, 16, 32])
@pytest.mark.parametrize("scale", [0.3, 2.0, 40.0, 3000.0, 1.0e5])
def test_half_boxes_never_prune_a_box_within_reach(d, scale):
    rng = np.random.default_rng(d * 1000 + int(scale))
    cases = 20000
    centre = rng.normal(0, scale, d)
    mid = (centre + rng.normal(0, scale, (cases, d)))
    half = np.abs(rng.normal(0, 0.7, (cases, d)))
    lo, hi = (mid - half).astype(np.float32), (mid + half).astype(np.float32)
    iw_q = rng.uniform(0.3, 1.6, cases)
    bound = rng.uniform(0.3, 1.6, cases)
    radius = 1.0 / (iw_q * bound)
    # queries at 0.2 .. 1.5 radii from the box surface, a third of them EXACTLY at the threshold (the adversarial case)
    direction = rng.normal(0, 1, (cases, d))
    direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    t = rng.uniform(0.2, 1.5, cases)
    t[::3] = 1.0
    corner = np.where(direction > 0, hi, lo).astype(np.float64)
    q = (corner + direction * (t * radius)[:, None]).astype(np.float32)
    kept = half_box_test(lo, hi, q, centre, iw_q, bound)
    must = exact_within(lo, hi, q, iw_q, bound)
    assert must.sum() > cases // 10
    assert not (must & ~kept).any(), f"{(must & ~kept).sum()} boxes within reach were pruned"
    if scale <= 2.0 and d >= 4:                                # and the filter still prunes: most boxes beyond 1.3 radii go
        far = t > 1.3
        assert (~kept[far]).mean() > 0.5


def test_half_boxes_beyond_the_half_range():
    """Coordinates that round to +-inf in half precision: the query's displacement becomes inf and every box is kept for it."""
    lo = np.array([[1.0e5, 0.0], [0.0, 0.0], [-7.0e4, 1.0]], np.float32)
    hi = lo + np.float32(1.0)
    q = np.array([[1.0e5, 0.5], [7.0e4, 0.0], [-7.0e4, 1.5]], np.float32)
    kept = half_box_test(lo, hi, q, np.zeros(2), np.ones(3), np.ones(3))
    assert kept[0] and kept[2]                                 # query inside / next to the box
    assert kept[1]                                             # not representable: cannot be decided here, must be kept
