"""Pins the CPU oracle against every numeric known-answer the reference's own tests
hold for this path (SURVEY.md section 4 "Numeric pins"), and cross-checks the C oracle
against the independent NumPy mirror."""
import numpy as np
import pytest

import oracle
from oracle import oracle_np
from lancedb_b200.index import train_ivf_pq


def test_l2_doctest_pin():
    # python/python/lancedb/table.py:3595-3603: q=[0.4,1.4,2.4] -> 5.220000, 23.089996
    q = [0.4, 1.4, 2.4]
    assert f"{oracle.l2(q, [0.5, 3.4, 1.3]):.6f}" == "5.220000"
    assert f"{oracle.l2(q, [0.3, 6.2, 2.6]):.6f}" == "23.089996"
    assert oracle.l2(q, [0.1, 2.3, 4.5]) > 5.22


def test_cosine_doctest_pin():
    # python/python/lancedb/query.py:1563-1571
    assert f"{oracle.cosine([0.4, 0.4], [0.4, 0.4]):.6f}" == "0.000000"
    assert f"{oracle.cosine([0.4, 0.4], [1.1, 1.2]):.6f}" == "0.000944"


def test_cosine_matches_numpy_formula():
    # python/python/tests/test_query.py:993-1014,1045-1046 (abs=1e-6, in [0,1])
    rng = np.random.default_rng(0)
    for x, y in [([4, 8], [1, 2]), ([4, 8], [3, 4])] + [
        (rng.random(16), rng.random(16)) for _ in range(20)]:
        x = np.asarray(x, np.float64); y = np.asarray(y, np.float64)
        want = 1 - np.dot(x, y) / (np.linalg.norm(x) * np.linalg.norm(y))
        got = oracle.cosine(x, y)
        assert got == pytest.approx(want, abs=1e-6)
        assert -1e-6 <= got <= 1 + 1e-6


def test_exact_match_is_zero():
    # python/python/tests/test_db.py:198-199, test_table.py:2975-2978
    v = np.random.default_rng(1).standard_normal(128).astype(np.float32)
    assert oracle.l2(v, v) == 0.0


def test_flat_order_and_tiebreak():
    # test_query.py:562-570 ([0,0] -> id 1 then id 2); tie-break (_distance, _rowid)
    # pinned by the plan doctest python/python/lancedb/query.py:1364-1370
    vec = np.array([[1, 2], [3, 4]], np.float32)
    ids, dist, cnt = oracle.flat_search(vec, [[0, 0]], k=10, row_ids=[1, 2])
    assert cnt[0] == 2 and list(ids[0, :2]) == [1, 2] and list(dist[0, :2]) == [5.0, 25.0]
    dup = np.array([[1, 1], [2, 2], [1, 1], [1, 1]], np.float32)
    ids, dist, cnt = oracle.flat_search(dup, [[1, 1]], k=2, row_ids=[9, 3, 7, 5])
    assert list(ids[0]) == [5, 7] and list(dist[0]) == [0.0, 0.0]


def test_distance_range_semantics():
    # [lower, upper): python/python/tests/test_query.py:655-676
    vec = np.array([[1, 2], [3, 4]], np.float32)
    lo, hi = 5.0, 25.0
    assert oracle.flat_search(vec, [[0, 0]], upper=lo)[2][0] == 0
    ids, dist, cnt = oracle.flat_search(vec, [[0, 0]], lower=hi)
    assert cnt[0] == 1 and dist[0, 0] == hi
    ids, dist, cnt = oracle.flat_search(vec, [[0, 0]], upper=hi)
    assert cnt[0] == 1 and dist[0, 0] == lo
    assert oracle.flat_search(vec, [[0, 0]], lower=lo)[2][0] == 2


@pytest.mark.parametrize("d", [3, 8, 16, 17, 40, 128])
def test_c_vs_numpy_mirror_distances(d):
    rng = np.random.default_rng(d)
    for _ in range(5):
        x = rng.standard_normal(d).astype(np.float32)
        y = rng.standard_normal(d).astype(np.float32)
        assert oracle.l2(x, y) == oracle_np.l2(x, y)
        assert oracle.dot(x, y) == oracle_np.dot(x, y)
        assert oracle.cosine(x, y) == oracle_np.cosine(x, y)
        assert np.array_equal(oracle.normalize(x), oracle_np.normalize(x))


@pytest.mark.parametrize("dsub", [4, 8, 16])
def test_subvec_tree_order(dsub):
    rng = np.random.default_rng(7)
    x = rng.standard_normal(dsub).astype(np.float32)
    cb = rng.standard_normal((256, dsub)).astype(np.float32)
    want = oracle_np.l2_subvec_batch(x, cb)
    got = np.array([oracle.l2_subvec(x, c) for c in cb], np.float32)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("metric", ["l2", "cosine", "dot"])
def test_ivfpq_c_vs_numpy_mirror(metric):
    rng = np.random.default_rng(42)
    x = rng.standard_normal((600, 32)).astype(np.float32)
    ix = train_ivf_pq(x, num_partitions=6, num_sub_vectors=4, distance_type=metric, max_iterations=4)
    oix = oracle.OracleIndex.from_data(ix)
    q = np.random.default_rng(43).standard_normal((5, 32)).astype(np.float32)
    ids, dist, cnt = oix.search(q, k=7, nprobes=3)
    for i in range(5):
        wi, wd = oracle_np.ivfpq_search_one(ix, q[i], 7, 3)
        assert cnt[i] == len(wi)
        assert np.array_equal(ids[i, :cnt[i]], wi)
        assert np.array_equal(dist[i, :cnt[i]], wd)
    # multi-threaded run is identical
    ids2, dist2, cnt2 = oix.search(q, k=7, nprobes=3, nthreads=3)
    assert np.array_equal(ids, ids2) and np.array_equal(dist, dist2) and np.array_equal(cnt, cnt2)


@pytest.mark.parametrize("metric", ["l2", "cosine", "dot"])
def test_prefilter_and_build_passes_c_vs_numpy_mirror(metric):
    """The C oracle's prefilter, IVF assignment and PQ encoding against the independent NumPy statement."""
    rng = np.random.default_rng(44)
    x = rng.standard_normal((500, 32)).astype(np.float32)
    ix = train_ivf_pq(x, num_partitions=5, num_sub_vectors=4, distance_type=metric, max_iterations=4)
    oix = oracle.OracleIndex.from_data(ix)
    q = rng.standard_normal((4, 32)).astype(np.float32)
    allowed = sorted(int(a) for a in rng.choice(500, 60, replace=False))
    bm = oracle.allow_bitmap(allowed, 500)
    ids, dist, cnt = oix.search(q, k=9, nprobes=3, allow=bm, allow_bits=500)
    for i in range(4):
        wi, wd = oracle_np.ivfpq_search_one(ix, q[i], 9, 3, allowed=set(allowed))
        assert cnt[i] == len(wi)
        assert np.array_equal(ids[i, :cnt[i]], wi) and np.array_equal(dist[i, :cnt[i]], wd)
    v = rng.standard_normal((12, 32)).astype(np.float32)
    parts = oix.ivf_assign(v)
    assert np.array_equal(parts, oracle_np.ivf_assign(ix, v))
    assert np.array_equal(oix.pq_encode(v, parts), oracle_np.pq_encode(ix, v, parts))


def test_ivfpq_multivector_relational():
    # test_query.py:791-850: same query twice gives same per-query distances
    # (the "2x" multivector sum is applied above the ANN path and is out of scope)
    rng = np.random.default_rng(5)
    x = rng.standard_normal((256, 8)).astype(np.float32)
    ix = train_ivf_pq(x, num_partitions=1, num_sub_vectors=2, distance_type="cosine", max_iterations=4)
    oix = oracle.OracleIndex.from_data(ix)
    q = rng.standard_normal(8).astype(np.float32)
    ids, dist, cnt = oix.search(np.stack([q, q]), k=10, nprobes=1)
    assert np.array_equal(ids[0], ids[1]) and np.array_equal(dist[0], dist[1])
    assert (dist[0] >= 0).all() and (dist[0] <= 2).all()


def test_golden_reference_pins_file():
    """tests/golden/reference_pins.json: the reference's own known-answers as data (with their source
    locations); the oracle reproduces every one of them."""
    import json, os
    pins = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_pins.json")))
    p = pins["l2_doctest"]
    ids, dist, cnt = oracle.flat_search(np.array(p["vectors"], np.float32), np.array(p["query"], np.float32), k=3)
    caps = ["bar", "foo", "test"]                      # row order in the doctest's table
    assert [caps[i] for i in ids[0]] == p["captions_in_order"]
    assert [f"{d:.6f}" for d in dist[0]] == p["distances_6dp"]
    p = pins["cosine_doctest"]
    ids, dist, cnt = oracle.flat_search(np.array(p["vectors"], np.float32), np.array(p["query"], np.float32), k=3,
                                        metric="cosine")
    assert [p["b"][i] for i in ids[0]] == p["b_in_order"]       # (_distance, _rowid) tie-break: 6 before 10
    assert [f"{d:.6f}" for d in dist[0]] == p["distances_6dp"]
    v = np.array([[1.0, 2.0], [3.0, 4.0]], np.float32)
    assert oracle.flat_search(v, v[0], k=1)[1][0, 0] == pins["exact_match"]["distance"]


@pytest.mark.parametrize("metric", ["l2", "cosine", "dot"])
def test_golden_ivfpq_fixture_matches_oracle(metric):
    """The committed oracle outputs (tests/golden/ivfpq_small.npz) are what oracle.c computes today: freezes the
    restatement between rounds.  The GPU twin of this test is tests/test_gpu_zz_fullsize.py::test_golden_ivfpq_fixture."""
    from tests.util import load_golden, same_result
    ix, q, cases, flat = load_golden(metric)
    orc = oracle.OracleIndex.from_data(ix)
    for name, (kw, want) in cases.items():
        assert same_result(orc.search(q, **kw), want), name
    assert same_result(oracle.flat_search(ix.vectors, q, k=7, metric=metric, row_ids=ix.row_ids), flat)


def test_ivfpq_distance_range_relational():
    """python/python/tests/test_query.py:723-787 (the ANN half; appended un-indexed rows are out of scope): on an IVF_PQ
    index with one partition and two sub-vectors, [lower, upper) filters the ANN distances it reports."""
    rng = np.random.default_rng(7)
    x = rng.random((256, 2)).astype(np.float32)
    oix = oracle.OracleIndex.from_data(train_ivf_pq(x, num_partitions=1, num_sub_vectors=2, max_iterations=4))
    q = np.zeros((1, 2), np.float32)
    ids, dist, cnt = oix.search(q, k=10, nprobes=1)
    assert cnt[0] == 10 and (np.diff(dist[0]) >= 0).all()
    lo, hi = float(dist[0, 0]), float(dist[0, 9])
    assert oix.search(q, k=10, nprobes=1, upper=lo)[2][0] == 0
    d = oix.search(q, k=10, nprobes=1, lower=hi)
    assert d[2][0] >= 1 and (d[1][0, :d[2][0]] >= hi).all()
    d = oix.search(q, k=10, nprobes=1, upper=hi)
    assert (d[1][0, :d[2][0]] < hi).all()
    d = oix.search(q, k=10, nprobes=1, lower=lo)
    assert d[2][0] == 10 and (d[1][0] >= lo).all()
