"""The bitmap pre-filter of the many-query box scan (csrc/box.cu, k_box_filter) restated in numpy: the set of
queries a row keeps after the AND of its 20 table entries must contain every query whose exact box test the row
passes (the filter may only ever drop pairs that cannot match).  CPU only: this is the exactness argument of the
kernel's header, checked on random, clustered, out-of-range, NaN and degenerate inputs."""
import numpy as np

BINS = 32


def _bin(x, scale, off):
    """sig_bin: floor(fma(x, scale, off)) clamped to [0, BINS), NaN -> 0 (float32 arithmetic, one rounding)."""
    with np.errstate(invalid="ignore", over="ignore"):
        t = (np.asarray(x, np.float64) * float(scale) + float(off)).astype(np.float32)
        b = np.floor(t)
    b = np.where(np.isnan(b), 0.0, b)
    return np.clip(b, 0, BINS - 1).astype(np.int64)


def _tables(lo, hi):
    """Per dim: bin geometry over the span of the non-empty intervals, and per (dim, bin) the query set."""
    Q, D = lo.shape
    some = lo <= hi
    gmin = np.where(some, lo, np.inf).min(0)
    gmax = np.where(some, hi, -np.inf).max(0)
    scale = np.zeros(D, np.float32)
    ok = (gmin < gmax) & np.isfinite(gmax - gmin)
    scale[ok] = (np.float32(BINS) / (gmax[ok] - gmin[ok]).astype(np.float32))
    gmin = np.where(gmin <= gmax, gmin, 0.0).astype(np.float32)
    off = (-gmin * scale).astype(np.float32)
    tab = np.zeros((D, BINS), np.uint64)
    for d in range(D):
        b0, b1 = _bin(lo[:, d], scale[d], off[d]), _bin(hi[:, d], scale[d], off[d])
        for q in range(Q):
            if some[q, d]:
                for b in range(b0[q], b1[q] + 1):
                    tab[d, b] |= np.uint64(1) << np.uint64(q)
    return scale, off, tab


def _survivors(rows, scale, off, tab):
    surv = np.full(len(rows), ~np.uint64(0))
    for d in range(rows.shape[1]):
        surv &= tab[d, _bin(rows[:, d], scale[d], off[d])]
    return surv


def _check(rows, lo, hi):
    scale, off, tab = _tables(lo, hi)
    surv = _survivors(rows, scale, off, tab)
    kept = 0
    for q in range(len(lo)):
        with np.errstate(invalid="ignore"):
            match = ((rows >= lo[q]) & (rows <= hi[q])).all(1)
        has = (surv >> np.uint64(q)) & np.uint64(1)
        assert not (match & (has == 0)).any(), f"query {q}: the filter dropped a matching row"
        kept += int(has.sum())
    return kept / (len(rows) * len(lo))


def test_filter_never_drops_a_matching_pair():
    rng = np.random.default_rng(3)
    D, Q, M = 20, 64, 4000
    w = rng.uniform(0.1, 50, D).astype(np.float32)
    centers = (rng.normal(0, 1, (300, D)) * w).astype(np.float32)
    rows = (centers[rng.integers(0, 300, M)] + rng.normal(0, 0.3, (M, D)) * w).astype(np.float32)
    qc = centers[:Q]
    lo, hi = (qc - w).astype(np.float32), (qc + w).astype(np.float32)
    frac = _check(rows, lo, hi)
    assert frac < 0.05                                  # and it is a filter: few pairs are left for the exact test
    # rows exactly on the faces, far outside the queries' range, NaN and infinite coordinates
    edge = np.vstack([lo[:10], hi[:10], rows[:5] + 1e6, rows[:5] - 1e6])
    edge[3, 4] = np.nan
    edge[7, 0] = np.inf
    edge[8, 19] = -np.inf
    _check(edge.astype(np.float32), lo, hi)


def test_degenerate_geometry():
    rng = np.random.default_rng(4)
    D = 20
    rows = rng.normal(0, 1, (500, D)).astype(np.float32)
    # all queries identical (zero-width range in no dim, but one shared interval), one empty interval (NaN query),
    # a dim where every interval is a single point, a huge range next to a tiny one
    base = rows[0]
    lo = np.repeat((base - 0.5)[None], 9, 0).astype(np.float32)
    hi = np.repeat((base + 0.5)[None], 9, 0).astype(np.float32)
    lo[3, 2], hi[3, 2] = np.inf, -np.inf                # empty interval: query 3 matches nothing
    lo[:, 5] = hi[:, 5] = base[5]                       # point intervals: scale 0, everything in bin 0
    lo[8, 7], hi[8, 7] = -1e30, 1e30
    frac = _check(np.vstack([rows, base[None]]), lo, hi)
    scale, off, tab = _tables(lo, hi)
    surv = _survivors(base[None].astype(np.float32), scale, off, tab)
    assert (surv[0] >> np.uint64(3)) & np.uint64(1) == 0          # the empty query survives nowhere
    assert (surv[0] >> np.uint64(0)) & np.uint64(1) == 1          # the row that sits at the centre keeps its queries
    assert 0 <= frac <= 1
