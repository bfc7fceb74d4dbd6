"""BASELINE.json configs[1] at full size (1M x 768, nlist 1024, m 96, nprobes 20, k 10, batch 1024) on a
synthetic (untrained) index: size-independent properties of the result -- sorted by (_distance, _rowid),
k unique rows per query, idempotent, independent of the batch a query travels in, equal to the merge of a
2-way partition-sharded search -- plus a bit-exact oracle spot-check on a few queries.  Also the committed
golden fixture (tests/golden/ivfpq_small.npz) against the CUDA path.  (File name sorts last on purpose: the
small parity tests run first.)"""
import numpy as np
import pytest

import oracle
from tests.util import queries, random_index

pytestmark = pytest.mark.gpu

C2 = dict(n=1_000_000, dim=768, nlist=1024, m=96, nprobes=20, k=10, batch=1024)


def check_properties(search, ix, q, k, nprobes, spot=6):
    """`search(index_data, queries) -> (ids, dist, cnt)`; raises AssertionError on the first broken property."""
    ids, dist, cnt = search(ix, q)
    B = q.shape[0]
    assert ids.shape == (B, k) and dist.shape == (B, k) and cnt.shape == (B,)
    assert (cnt == k).all(), "every query probes far more than k rows here"
    assert np.isfinite(dist).all()
    # sorted by (_distance ASC, _rowid ASC) and duplicate-free
    d0, d1 = dist[:, :-1], dist[:, 1:]
    assert (d0 <= d1).all()
    tie = d0 == d1
    assert (ids[:, :-1][tie] < ids[:, 1:][tie]).all()
    assert all(len(set(row.tolist())) == k for row in ids[:64])
    # idempotent
    ids2, dist2, cnt2 = search(ix, q)
    assert np.array_equal(ids, ids2) and np.array_equal(dist.view(np.uint32), dist2.view(np.uint32))
    # a query's result does not depend on the batch it is in (tiles group queries by partition)
    sub = np.concatenate([np.arange(0, B, 17), [B - 1]])
    ids3, dist3, _ = search(ix, q[sub])
    assert np.array_equal(ids3, ids[sub]) and np.array_equal(dist3.view(np.uint32), dist[sub].view(np.uint32))
    # partition-sharded search (each shard holds half of the partitions) merges to the same top-k
    merged_i, merged_d = [], []
    parts = [search(ix.shard(r, 2), q[:128]) for r in range(2)]
    for qi in range(128):
        cand = sorted((float(parts[r][1][qi, j]), int(parts[r][0][qi, j])) for r in range(2) for j in range(int(parts[r][2][qi])))
        merged_d.append([c[0] for c in cand[:k]]); merged_i.append([c[1] for c in cand[:k]])
    assert np.array_equal(np.array(merged_i, np.uint64), ids[:128])
    assert np.array_equal(np.array(merged_d, np.float32).view(np.uint32), dist[:128].view(np.uint32))
    # oracle spot-check, bit-exact
    orc = oracle.OracleIndex.from_data(ix)
    pick = np.linspace(0, B - 1, spot).astype(int)
    oi, od, oc = orc.search(q[pick], k=k, nprobes=nprobes, nthreads=min(spot, 8))
    assert np.array_equal(oi, ids[pick]) and np.array_equal(od.view(np.uint32), dist[pick].view(np.uint32))
    assert np.array_equal(oc, cnt[pick])


@pytest.mark.parametrize("metric", ["l2", "cosine", "dot"])
def test_golden_ivfpq_fixture(metric):
    """The CUDA path reproduces the committed fixture (tests/golden/ivfpq_small.npz) byte for byte."""
    from lancedb_b200 import _native
    from tests.util import load_golden, same_result
    ix, q, cases, flat = load_golden(metric)
    gpu = _native.GpuIvfPq(ix)
    for name, (kw, want) in cases.items():
        assert same_result(gpu.search(q, **kw), want), name
    gpu.close()
    fl = _native.GpuFlat(ix.vectors, ix.row_ids)
    assert same_result(fl.search(q, k=7, metric=metric), flat)
    fl.close()


def test_config2_full_size_properties():
    from lancedb_b200 import _native
    rng = np.random.default_rng(77)
    ix = random_index(rng, dim=C2["dim"], nlist=C2["nlist"], m=C2["m"], n=C2["n"], shuffle_ids=False)
    q = queries(rng, C2["batch"], C2["dim"])
    handles = {}

    def search(data, qq):
        key = id(data)
        if key not in handles:                       # keep `data` referenced: a freed shard's id() could be reused
            handles[key] = (data, _native.GpuIvfPq(data))
        return handles[key][1].search(qq, k=C2["k"], nprobes=C2["nprobes"])

    try:
        check_properties(search, ix, q, C2["k"], C2["nprobes"])
    finally:
        for _, h in handles.values():
            h.close()
