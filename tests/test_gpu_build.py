"""Index-build passes through the C ABI (lgpu_ivf_assign / lgpu_pq_encode, csrc/build.cu) against the oracle:
bit-exact partition ids and PQ codes, and an index built with them searched against the oracle."""
import numpy as np
import pytest

import oracle
from lancedb_b200 import _native
from lancedb_b200.index import train_ivf_pq
from tests.util import queries, random_index

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("metric", ["l2", "cosine", "dot"])
@pytest.mark.parametrize("dim,m,nlist", [(64, 8, 17), (48, 3, 5), (32, 32, 40), (768, 96, 64)])
def test_assign_and_encode_match_oracle(metric, dim, m, nlist):
    rng = np.random.default_rng(41)
    ix = random_index(rng, dim=dim, nlist=nlist, m=m, metric=metric, n=200)
    orc = oracle.OracleIndex.from_data(ix)
    n = 1000 if dim < 768 else 300
    v = queries(rng, n, dim)
    v[3] = ix.centroids[min(2, nlist - 1)]                      # a row sitting exactly on a centroid
    parts = _native.ivf_assign(ix.centroids, v, metric)
    assert np.array_equal(parts, orc.ivf_assign(v))
    codes = _native.pq_encode(ix.centroids, ix.codebook, v, parts, metric)
    assert codes.shape == (n, m) and codes.dtype == np.uint8
    assert np.array_equal(codes, orc.pq_encode(v, parts))


def test_encode_rejects_bad_input():
    rng = np.random.default_rng(42)
    ix = random_index(rng, dim=32, nlist=4, m=4, n=50)
    v = queries(rng, 5, 32)
    with pytest.raises(ValueError):
        _native.pq_encode(ix.centroids, ix.codebook, v, np.full(5, 9, np.uint32))        # partition out of range
    assert _native.ivf_assign(ix.centroids, v[:0]).shape == (0,)


@pytest.mark.parametrize("metric", ["l2", "cosine"])
def test_index_built_with_native_passes(metric):
    """train_ivf_pq(native_passes=True): every row's own vector probes its partition first, its codes are the
    oracle's encoding for the trained centroids/codebook, and the index searches bit-identically."""
    rng = np.random.default_rng(43)
    n, dim = 3000, 32
    v = (rng.standard_normal((n, 8)) @ rng.standard_normal((8, dim))).astype(np.float32)
    v += 0.05 * rng.standard_normal((n, dim)).astype(np.float32)
    ix = train_ivf_pq(v, num_partitions=12, num_sub_vectors=4, distance_type=metric, max_iterations=8,
                      keep_vectors=True, device="cuda", native_passes=True)
    orc = oracle.OracleIndex.from_data(ix)
    part_of_row = np.repeat(np.arange(ix.nlist), np.diff(ix.part_offsets.astype(np.int64))).astype(np.uint32)
    assert np.array_equal(orc.ivf_assign(ix.vectors), part_of_row)
    want = orc.pq_encode(ix.vectors, part_of_row)
    for p in range(ix.nlist):
        a, b = int(ix.part_offsets[p]), int(ix.part_offsets[p + 1])
        got = ix.codes_t[a * ix.m:b * ix.m].reshape(ix.m, b - a).T
        assert np.array_equal(got, want[a:b])
    gpu = _native.GpuIvfPq(ix)
    q = v[rng.choice(n, 40, replace=False)]
    gi, gd, gc = gpu.search(q, k=5, nprobes=4)
    oi, od, oc = orc.search(q, k=5, nprobes=4, nthreads=4)
    gpu.close()
    assert np.array_equal(gi, oi) and np.array_equal(gd.view(np.uint32), od.view(np.uint32)) and np.array_equal(gc, oc)
