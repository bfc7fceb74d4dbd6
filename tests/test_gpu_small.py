"""The low-latency path for tiny batches (lancedb_b200/csrc/small.cu: one CTA per (query, probed partition), exact
f32 table in shared memory) against the oracle: same bit-identical contract as the batched kernels, over every
request feature the path serves (metrics, sub-vector lengths, k above one warp's selector, distance range,
refine, prefilter + maximum_nprobes widening).  B = 1 is the reference's own calling pattern
(rust/lancedb/src/query.rs:1005-1011: one plan per query vector)."""
import numpy as np
import pytest

import oracle
from lancedb_b200 import _native
from tests.util import queries, random_index

pytestmark = [pytest.mark.gpu, pytest.mark.small_path]


def _same(ix, q, k, nprobes, **kw):
    gpu = _native.GpuIvfPq(ix)
    orc = oracle.OracleIndex.from_data(ix)
    g = gpu.search(q, k=k, nprobes=nprobes, **kw)
    o = orc.search(q, k=k, nprobes=nprobes, nthreads=8, **kw)
    gpu.close()
    assert np.array_equal(g[2], o[2])
    assert np.array_equal(g[0], o[0])
    assert np.array_equal(g[1].view(np.uint32), o[1].view(np.uint32))
    return g


@pytest.mark.parametrize("metric", ["l2", "cosine", "dot"])
@pytest.mark.parametrize("dim,m", [(768, 96), (64, 8), (80, 10), (64, 2), (24, 24), (32, 8), (1536, 96)])
def test_small_batches_match_the_oracle(metric, dim, m):
    rng = np.random.default_rng(71)
    sizes = [0, 5, 1500, 1537, 3100, 40, 977, 200, 128, 129, 64, 1, 700, 0, 1024, 333]
    ix = random_index(rng, dim=dim, nlist=len(sizes), m=m, metric=metric, sizes=sizes)
    for B in (1, 2, 16):
        _same(ix, queries(rng, B, dim), k=10, nprobes=6)
    _same(ix, queries(rng, 1, dim), k=10, nprobes=len(sizes))           # every partition, empty ones included


def test_small_batch_request_features():
    rng = np.random.default_rng(72)
    ix = random_index(rng, dim=64, nlist=24, m=8, n=30000, with_vectors=True)
    q = queries(rng, 3, 64)
    _same(ix, q, k=1, nprobes=5)
    _same(ix, q, k=100, nprobes=5)                                       # block selector
    _same(ix, q, k=1024, nprobes=2)                                      # more than some probes hold
    g = _same(ix, q, k=20, nprobes=5)
    lo, hi = float(np.median(g[1][0])), float(g[1][0].max())
    _same(ix, q, k=20, nprobes=5, lower=lo, upper=hi)                    # distance_range
    _same(ix, q, k=7, nprobes=5, refine_factor=5)                        # refine on the raw vectors
    ixc = random_index(rng, dim=48, nlist=6, m=6, metric="cosine", n=4000, with_vectors=True)
    _same(ixc, queries(rng, 2, 48), k=5, nprobes=3, refine_factor=4)


def test_small_batch_prefilter_and_widening():
    rng = np.random.default_rng(73)
    ix = random_index(rng, dim=32, nlist=24, m=8, n=24000)
    q = queries(rng, 4, 32)
    bm = oracle.allow_bitmap(np.sort(rng.choice(24000, 240, replace=False)), 24000)          # 1 % selective
    gpu = _native.GpuIvfPq(ix)
    orc = oracle.OracleIndex.from_data(ix)
    for kw in (dict(nprobes=2), dict(nprobes=2, max_nprobes=24)):
        g = gpu.search(q, k=10, allow=bm, allow_bits=24000, **kw)
        o = orc.search(q, k=10, allow=bm, allow_bits=24000, nthreads=4, **kw)
        assert np.array_equal(g[2], o[2]) and np.array_equal(g[0], o[0])
        assert np.array_equal(g[1].view(np.uint32), o[1].view(np.uint32))
    gpu.close()


def test_the_small_path_is_the_one_that_ran():
    """4 launches (coarse distances, select, small_scan, select) instead of the batched pipeline's ~25."""
    rng = np.random.default_rng(74)
    ix = random_index(rng, dim=64, nlist=32, m=8, n=20000)
    gpu = _native.GpuIvfPq(ix)
    q = queries(rng, 1, 64)
    gpu.search(q, k=10, nprobes=8)
    n0 = _native.kernel_launch_count()
    gpu.search(q, k=10, nprobes=8)
    n1 = _native.kernel_launch_count()
    gpu.close()
    assert 1 <= n1 - n0 <= 6, n1 - n0
