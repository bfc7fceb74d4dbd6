"""GPU parity tests proper: every call goes through the C ABI (ctypes) and is compared
with the CPU oracle on the same seeded inputs.  Bars: row ids / partition ids / counts
bit-exact; distances bit-exact too (the kernels reproduce the oracle's rounding order),
which is stricter than the 1e-4 relative tolerance north_star allows."""
import numpy as np
import pytest

import oracle
from lancedb_b200 import _native
from tests.util import queries, random_index

pytestmark = pytest.mark.gpu


def _check_search(ix, q, k, nprobes, **kw):
    gpu = _native.GpuIvfPq(ix)
    orc = oracle.OracleIndex.from_data(ix)
    gi, gd, gc = gpu.search(q, k=k, nprobes=nprobes, **kw)
    oi, od, oc = orc.search(q, k=k, nprobes=nprobes, nthreads=8, **kw)
    gpu.close()
    assert np.array_equal(gc, oc), f"counts differ: {np.nonzero(gc != oc)[0][:10]}"
    bad = np.nonzero((gi != oi).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} queries with different row ids, first {bad[:5]}: {gi[bad[0]]} vs {oi[bad[0]]}"
    assert np.array_equal(gd.view(np.uint32), od.view(np.uint32)), "distances not bit-identical"
    return gi, gd, gc


@pytest.mark.parametrize("metric", ["l2", "cosine", "dot"])
@pytest.mark.parametrize("dim,nlist", [(32, 7), (768, 64), (100, 33)])
def test_coarse_partitions(metric, dim, nlist):
    rng = np.random.default_rng(1)
    m = 4 if dim % 4 == 0 else 1
    ix = random_index(rng, dim=dim, nlist=nlist, m=m if (dim // m) in (1, 2, 4, 8, 16, 32) else dim // 4,
                      metric=metric, n=500)
    q = queries(rng, 37, dim)
    gpu = _native.GpuIvfPq(ix)
    orc = oracle.OracleIndex.from_data(ix)
    nprobes = min(5, nlist)
    parts, dists = gpu.debug_coarse(q, nprobes)
    for i in range(q.shape[0]):
        qn = oracle.normalize(q[i]) if metric == "cosine" else q[i]
        op, od, _ = orc.find_partitions(qn, nprobes)
        assert np.array_equal(parts[i], op)
        assert np.array_equal(dists[i].view(np.uint32), od.view(np.uint32))
    gpu.close()


@pytest.mark.parametrize("metric", ["l2", "cosine", "dot"])
@pytest.mark.parametrize("dim,m", [(768, 96), (64, 4), (80, 10), (32, 8), (64, 2), (24, 24)])
def test_partition_distances_bit_exact(metric, dim, m):
    """LUT build + code scan (K2+K3) against oracle: dsub 8/16/8/4/32/1, chunk padding (m=10)."""
    rng = np.random.default_rng(2)
    sizes = [0, 1, 31, 33, 977, 2048, 2049, 5000]
    ix = random_index(rng, dim=dim, nlist=len(sizes), m=m, metric=metric, sizes=sizes)
    q = queries(rng, 1, dim)[0]
    gpu = _native.GpuIvfPq(ix)
    orc = oracle.OracleIndex.from_data(ix)
    for p, n in enumerate(sizes):
        if n == 0:
            continue
        g = gpu.debug_partition_distances(q, p, n)
        o = orc.partition_distances(q, p)
        assert np.array_equal(g.view(np.uint32), o.view(np.uint32)), (
            f"partition {p} (n={n}): {np.nonzero(g != o)[0][:8]} {g[:4]} {o[:4]}")
    gpu.close()


@pytest.mark.parametrize("metric", ["l2", "cosine", "dot"])
def test_search_small(metric):
    rng = np.random.default_rng(3)
    ix = random_index(rng, dim=64, nlist=16, m=8, metric=metric, n=4000)
    _check_search(ix, queries(rng, 50, 64), k=10, nprobes=4)


@pytest.mark.parametrize("B", [1, 7, 8, 9, 129])
def test_search_batch_shapes(B):
    rng = np.random.default_rng(4)
    ix = random_index(rng, dim=128, nlist=20, m=16, n=6000)
    _check_search(ix, queries(rng, B, 128), k=10, nprobes=5)


def test_search_ragged_partitions():
    """empty, 1-row and multi-tile partitions; nprobes == nlist; k above the candidate count"""
    rng = np.random.default_rng(5)
    sizes = [0, 0, 1, 2, 40, 3000, 0, 4500, 17, 900]
    ix = random_index(rng, dim=64, nlist=len(sizes), m=8, sizes=sizes)
    q = queries(rng, 33, 64)
    _check_search(ix, q, k=10, nprobes=len(sizes))
    _check_search(ix, q, k=100, nprobes=3)
    _check_search(ix, q, k=1, nprobes=1)
    tiny = random_index(rng, dim=64, nlist=4, m=8, sizes=[1, 0, 2, 0])
    gi, gd, gc = _check_search(tiny, q, k=10, nprobes=4)
    assert (gc == 3).all() and (gi[:, 3:] == np.iinfo(np.uint64).max).all() and np.isinf(gd[:, 3:]).all()


def test_search_duplicate_codes_tiebreak():
    """identical PQ codes => identical distances; order must be by row id"""
    rng = np.random.default_rng(6)
    ix = random_index(rng, dim=32, nlist=3, m=4, sizes=[300, 200, 100])
    ct = ix.codes_t.copy()
    for p in range(3):            # make every row of a partition share one code word
        a, b = int(ix.part_offsets[p]), int(ix.part_offsets[p + 1])
        blk = ct[a * 4:b * 4].reshape(4, b - a)
        blk[:] = blk[:, :1]
    ix.codes_t = ct
    gi, gd, gc = _check_search(ix, queries(rng, 9, 32), k=25, nprobes=3)
    assert (np.diff(gd, axis=1) >= 0).all()


def test_search_distance_range_and_large_k():
    rng = np.random.default_rng(7)
    ix = random_index(rng, dim=64, nlist=8, m=8, n=3000)
    q = queries(rng, 20, 64)
    gi, gd, gc = _check_search(ix, q, k=50, nprobes=4)
    lo, hi = float(gd[0, 5]), float(gd[0, 30])
    _check_search(ix, q, k=50, nprobes=4, lower=lo, upper=hi)
    _check_search(ix, q, k=50, nprobes=4, upper=lo)
    _check_search(ix, q, k=50, nprobes=4, lower=hi)
    _check_search(ix, q, k=1024, nprobes=8)


@pytest.mark.parametrize("metric", ["l2", "cosine", "dot"])
def test_search_refine(metric):
    rng = np.random.default_rng(8)
    ix = random_index(rng, dim=48, nlist=6, m=6, metric=metric, n=2000, with_vectors=True)
    _check_search(ix, queries(rng, 17, 48), k=10, nprobes=3, refine_factor=5)


@pytest.mark.parametrize("metric", ["l2", "cosine", "dot"])
def test_search_filter_verify_with_ties_and_fixup(metric):
    """The default path (scan3.cu lower bounds from 16-bit per-query tables -> shortlist -> proof -> exact
    re-score, tables.cu) must return exactly what the oracle returns, including when every distance of a
    partition ties, which makes the proof fail and sends those queries through the exact fix-up pass."""
    rng = np.random.default_rng(21)
    ix = random_index(rng, dim=64, nlist=24, m=8, metric=metric, n=9000)
    _check_search(ix, queries(rng, 40, 64), k=10, nprobes=6)
    _check_search(ix, queries(rng, 40, 64), k=40, nprobes=6)         # kp = 112: block selector shortlist
    ct = ix.codes_t.copy()                       # identical codes -> every distance ties -> fix-up pass
    a, b = int(ix.part_offsets[3]), int(ix.part_offsets[4])
    blk = ct[a * 8:b * 8].reshape(8, b - a); blk[:] = blk[:, :1]
    ix.codes_t = ct
    _check_search(ix, queries(rng, 12, 64), k=10, nprobes=24)


def test_search_config2_shape_subset():
    """BASELINE config 2's geometry (d=768, m=96, nprobes=20, k=10) at 1/8 of its rows."""
    rng = np.random.default_rng(9)
    ix = random_index(rng, dim=768, nlist=128, m=96, n=125_000)
    _check_search(ix, queries(rng, 256, 768), k=10, nprobes=20)


def test_search_k100_cosine_config3_shape():
    rng = np.random.default_rng(10)
    ix = random_index(rng, dim=768, nlist=64, m=96, metric="cosine", n=150_000)
    _check_search(ix, queries(rng, 96, 768), k=100, nprobes=50)


@pytest.mark.parametrize("metric", ["l2", "cosine", "dot"])
@pytest.mark.parametrize("n,dim", [(1000, 128), (5000, 40), (257, 1536)])
def test_flat_parity(metric, n, dim):
    rng = np.random.default_rng(11)
    v = queries(rng, n, dim)
    v[n // 2] = v[3]                               # an exact duplicate row -> tie
    rid = rng.permutation(n).astype(np.uint64)
    q = queries(rng, 19, dim)
    q[0] = v[3]
    fl = _native.GpuFlat(v, row_ids=rid)
    gi, gd, gc = fl.search(q, k=10, metric=metric)
    oi, od, oc = oracle.flat_search(v, q, k=10, metric=metric, row_ids=rid, nthreads=8)
    fl.close()
    assert np.array_equal(gc, oc) and np.array_equal(gi, oi)
    assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))


def test_flat_reference_doctest_pins():
    """python/python/lancedb/table.py:3595-3603 and query.py:1563-1571 through the C ABI"""
    fl = _native.GpuFlat(np.array([[0.1, 2.3, 4.5], [0.5, 3.4, 1.3], [0.3, 6.2, 2.6]], np.float32))
    ids, dist, cnt = fl.search([[0.4, 1.4, 2.4]], k=3)   # the doctest filters row 0 out (original_width > 1000)
    assert list(ids[0]) == [1, 0, 2]
    assert [f"{d:.6f}" for d in dist[0]] == ["5.220000", f"{oracle.l2([0.4, 1.4, 2.4], [0.1, 2.3, 4.5]):.6f}", "23.089996"]
    fl.close()
    fl = _native.GpuFlat(np.array([[1.1, 1.2], [0.5, 1.3], [0.4, 0.4], [0.4, 0.4]], np.float32))
    ids, dist, cnt = fl.search([[0.4, 0.4]], k=3, metric="cosine")
    assert list(ids[0]) == [2, 3, 0] and [f"{d:.6f}" for d in dist[0]] == ["0.000000", "0.000000", "0.000944"]
    fl.close()


def test_error_contract():
    rng = np.random.default_rng(12)
    ix = random_index(rng, dim=32, nlist=4, m=4, n=100)
    gpu = _native.GpuIvfPq(ix)
    q = queries(rng, 2, 32)
    with pytest.raises(ValueError, match="minimum_nprobes must be greater than 0"):
        gpu.search(q, nprobes=0)
    with pytest.raises(ValueError, match="limit"):
        gpu.search(q, k=0)
    with pytest.raises(ValueError, match="refine_factor needs the raw vectors"):
        gpu.search(q, refine_factor=2)
    gpu.close()
    bad = random_index(rng, dim=30, nlist=2, m=2, n=10)      # dsub = 15
    with pytest.raises(ValueError, match="sub-vector length"):
        _native.GpuIvfPq(bad)


@pytest.mark.parametrize("metric,dim,m", [("l2", 768, 96), ("dot", 64, 8), ("cosine", 80, 10), ("l2", 64, 2),
                                          ("l2", 24, 24), ("dot", 32, 8), ("cosine", 1536, 96)])
def test_filter_and_exact_scans_agree_with_the_oracle(metric, dim, m, monkeypatch):
    """LGPU_EXACT_SCAN=1 sends every query through the exact kernel (scan2.cu); the default runs the filter
    kernel (scan3.cu) + verify.  Both must reproduce the oracle bit for bit: many tiles per CTA, mixed group
    sizes 1..8, partitions above one 1536- / 3072-row tile, dsub 8/8/8/32/1/4/16, m not a multiple of 8."""
    rng = np.random.default_rng(21)
    sizes = [0, 5, 1500, 1537, 3100, 40, 977, 200, 128, 129, 64, 1, 700, 0, 1024, 333, 6200]
    ix = random_index(rng, dim=dim, nlist=len(sizes), m=m, metric=metric, sizes=sizes)
    q = queries(rng, 77, dim)
    out = {}
    for exact in ("0", "1"):
        monkeypatch.setenv("LGPU_EXACT_SCAN", exact)
        out[exact] = _check_search(ix, q, k=10, nprobes=6)
    for a, b in zip(out["0"], out["1"]):
        assert np.array_equal(a, b)


def test_filter_scan_scaled_and_degenerate_tables():
    """Quantiser edge cases: huge / tiny magnitudes (the step follows the per-query range), a codebook with a
    constant sub-space (range 0 in that sub-space), and an all-equal codebook (step 0: every row ties)."""
    rng = np.random.default_rng(23)
    for scale in (1e-6, 1.0, 3e4):
        ix = random_index(rng, dim=64, nlist=10, m=8, n=5000, scale=scale)
        _check_search(ix, queries(rng, 30, 64, scale=scale), k=10, nprobes=4)
    ix = random_index(rng, dim=64, nlist=10, m=8, n=5000)
    cb = ix.codebook.copy(); cb[3] = cb[3, :1]; ix.codebook = cb            # sub-space 3: all codewords equal
    _check_search(ix, queries(rng, 30, 64), k=10, nprobes=4)
    cb = ix.codebook.copy(); cb[:] = cb[:, :1]; ix.codebook = cb            # every sub-space constant
    _check_search(ix, queries(rng, 9, 64), k=10, nprobes=4)


def test_streaming_scan_many_tiles_per_cta():
    """More tiles than CTAs (each persistent CTA streams tens of tiles through the table ring) with
    group sizes from 1 to 8 and both table-build mappings alternating."""
    rng = np.random.default_rng(22)
    nlist = 600
    sizes = rng.integers(1, 400, size=nlist)
    ix = random_index(rng, dim=32, nlist=nlist, m=4, sizes=sizes)
    _check_search(ix, queries(rng, 700, 32), k=5, nprobes=9)


# ---------------------------------------------------------------- prefilter (row-id allow-list)
@pytest.mark.parametrize("metric", ["l2", "cosine", "dot"])
@pytest.mark.parametrize("frac", [0.5, 0.02])
def test_prefiltered_ivf_search(metric, frac):
    """lgpu_search_filtered against the oracle with the same bitmap: dense and selective allow-lists,
    ids beyond allow_bits excluded, k larger than some queries' allowed candidates."""
    rng = np.random.default_rng(31)
    ix = random_index(rng, dim=64, nlist=24, m=8, metric=metric, n=9000)
    n = int(ix.part_offsets[-1])
    nbits = n - 37                                      # the last 37 row ids fall outside the bitmap
    mask = rng.random(nbits) < frac
    bm = _native.mask_bitmap(mask)
    assert np.array_equal(bm, oracle.allow_bitmap(np.nonzero(mask)[0], nbits))
    gi, gd, gc = _check_search(ix, queries(rng, 41, 64), k=25, nprobes=6, allow=bm, allow_bits=nbits)
    for i in range(41):
        got = gi[i, :gc[i]].astype(np.int64)
        assert (got < nbits).all() and mask[got].all()
    if frac < 0.1:
        assert (gc < 25).any()                          # selective filter: fewer than k allowed rows probed


def test_prefiltered_ivf_with_refine_and_range():
    rng = np.random.default_rng(32)
    ix = random_index(rng, dim=32, nlist=10, m=4, n=5000, with_vectors=True)
    mask = rng.random(5000) < 0.3
    bm = _native.mask_bitmap(mask)
    _check_search(ix, queries(rng, 19, 32), k=8, nprobes=5, refine_factor=3, allow=bm, allow_bits=5000)
    _check_search(ix, queries(rng, 19, 32), k=8, nprobes=5, lower=1.0, upper=40.0, allow=bm, allow_bits=5000)
    empty = np.zeros_like(bm)
    gi, gd, gc = _check_search(ix, queries(rng, 5, 32), k=8, nprobes=5, allow=empty, allow_bits=5000)
    assert (gc == 0).all() and (gi == np.iinfo(np.uint64).max).all()


@pytest.mark.parametrize("metric", ["l2", "cosine", "dot"])
def test_prefiltered_flat_search(metric):
    rng = np.random.default_rng(33)
    n, dim = 6000, 48                                   # n >= 4096: the unfiltered call would take the tensor-core path
    v = rng.standard_normal((n, dim)).astype(np.float32)
    ids = rng.permutation(n).astype(np.uint64)
    q = rng.standard_normal((33, dim)).astype(np.float32)
    mask = rng.random(n) < 0.1
    bm = _native.mask_bitmap(mask)
    fl = _native.GpuFlat(v, ids)
    gi, gd, gc = fl.search(q, k=12, metric=metric, allow=bm, allow_bits=n)
    fl.close()
    oi, od, oc = oracle.flat_search(v, q, k=12, metric=metric, row_ids=ids, allow=bm, allow_bits=n, nthreads=8)
    assert np.array_equal(gc, oc) and np.array_equal(gi, oi)
    assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))
    assert mask[gi[gc > 0][:, 0].astype(np.int64)].all()


@pytest.mark.parametrize("metric", ["l2", "dot"])
def test_streaming_scan_many_tiles_dsub8(metric):
    """The dsub == 8 build path (cp.async codebook staging, both table-build mappings) with ~7 tiles per
    persistent CTA and three table chunks per tile (m = 20: the last chunk is half padding)."""
    rng = np.random.default_rng(23)
    nlist = 300
    sizes = rng.integers(1, 260, size=nlist)
    ix = random_index(rng, dim=160, nlist=nlist, m=20, metric=metric, sizes=sizes)
    _check_search(ix, queries(rng, 500, 160), k=10, nprobes=12)


@pytest.mark.parametrize("metric", ["l2", "cosine"])
def test_maximum_nprobes_widening_under_a_selective_prefilter(metric):
    """rust/lancedb/src/query.rs:1250-1275: "the excess partitions will only be searched if the initial search
    does not return enough results ... useful when there is a narrow filter".  With a 1 % allow-list and
    minimum_nprobes = 2 most queries find fewer than k rows; those (and only those) must be answered from their
    maximum_nprobes nearest partitions, exactly as the oracle's restatement does."""
    rng = np.random.default_rng(41)
    n = 20000
    ix = random_index(rng, dim=64, nlist=32, m=8, metric=metric, n=n, shuffle_ids=False)
    q = queries(rng, 60, 64)
    keep = rng.random(n) < 0.01
    bm = oracle.allow_bitmap(np.nonzero(keep)[0], n)
    gpu = _native.GpuIvfPq(ix)
    orc = oracle.OracleIndex.from_data(ix)
    narrow = gpu.search(q, k=10, nprobes=2, allow=bm, allow_bits=n)
    wide = gpu.search(q, k=10, nprobes=2, max_nprobes=24, allow=bm, allow_bits=n)
    o_narrow = orc.search(q, k=10, nprobes=2, allow=bm, allow_bits=n)
    o_wide = orc.search(q, k=10, nprobes=2, max_nprobes=24, allow=bm, allow_bits=n)
    o_24 = orc.search(q, k=10, nprobes=24, allow=bm, allow_bits=n)
    gpu.close()
    for got, want in ((narrow, o_narrow), (wide, o_wide)):
        assert np.array_equal(got[2], want[2]) and np.array_equal(got[0], want[0])
        assert np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32))
    short = o_narrow[2] < 10
    assert short.any() and (~short).any()                       # the case has both kinds of queries
    assert np.array_equal(o_wide[0][short], o_24[0][short])      # widened queries = a 24-probe search
    assert np.array_equal(o_wide[0][~short], o_narrow[0][~short])  # the others are untouched
    # without a filter maximum_nprobes changes nothing
    gpu = _native.GpuIvfPq(ix)
    a = gpu.search(q, k=10, nprobes=2); b = gpu.search(q, k=10, nprobes=2, max_nprobes=24)
    gpu.close()
    assert np.array_equal(a[0], b[0])


@pytest.mark.parametrize("metric", ["l2", "cosine"])
def test_candidate_lists_hold_for_large_k(metric):
    """k = 100 over ~60 tiles per query of i.i.d. data: tile-local thresholds stop at the k-th smallest of ONE tile,
    so every tile would append ~k rows (6000 per query); the scanners tighten tau from the query's own list (scan3.cu)
    so the 2048-entry lists hold -- (almost) no query falls back to the exact kernels -- and the result stays
    bit-identical either way.  The order in which tiles publish thresholds is a race, so the counters get a margin;
    the ids and distance bits do not.  Few queries over many tiles (the last case) are routed to the dense mode by
    the host (api.cu: expected concurrent tiles per query x k against the capacity)."""
    rng = np.random.default_rng(61)
    ix = random_index(rng, dim=64, nlist=40, m=8, metric=metric, n=300000)
    q = queries(rng, 256, 64)
    _native.set_profiling(True)
    try:
        _check_search(ix, q, k=100, nprobes=12)
        st = _native.last_filter_stats()
        assert st["queries"] == 256 and st["flagged_queries"] <= 26, st   # <= 10 % through the exact fix-up (a race: 1..10 seen)
        assert 0 < st["candidates"] <= 256 * 1400, st          # candidate mode ran, well inside the capacity
        _check_search(ix, q, k=128, nprobes=12)
        assert _native.last_filter_stats()["flagged_queries"] <= 51   # 16 x k = capacity: <= 20 % (6..10 seen)
        _check_search(ix, q[:9], k=40, nprobes=40)               # every partition probed by 9 queries: dense mode
        assert _native.last_filter_stats()["flagged_queries"] <= 1
    finally:
        _native.set_profiling(False)


@pytest.mark.parametrize("mode", ["cand_overflow", "dense", "dense_above_32"])
def test_filter_scan_modes_and_candidate_overflow(mode, monkeypatch):
    """The filter scan's candidate mode (thresholds inside the scanners, per-query candidate lists) against its
    dense mode (LGPU_DENSE_FILTER=1: one lower bound per row + shortlist select) and against a candidate capacity
    of 32 (LGPU_CAND_CAP), which overflows for most queries and sends them through the exact fix-up pass."""
    if mode == "dense":
        monkeypatch.setenv("LGPU_DENSE_FILTER", "1")
    elif mode == "dense_above_32":
        monkeypatch.setenv("LGPU_CAND_KMAX", "32")               # k = 100 below: dense lower bounds + proven prefix
    else:
        monkeypatch.setenv("LGPU_CAND_CAP", "32")
    rng = np.random.default_rng(51)
    ix = random_index(rng, dim=64, nlist=40, m=8, n=60000)
    _check_search(ix, queries(rng, 90, 64), k=10, nprobes=12)
    _check_search(ix, queries(rng, 90, 64), k=32, nprobes=12)
    _check_search(ix, queries(rng, 40, 64), k=100, nprobes=12)       # 2048-entry candidate lists, block selector
    ixv = random_index(rng, dim=48, nlist=6, m=6, metric="cosine", n=4000, with_vectors=True)
    _check_search(ixv, queries(rng, 17, 48), k=5, nprobes=3, refine_factor=4)
