"""tcgen05 GEMM shortlist: numerics of the GEMM itself (tolerance: it is bf16), and bit-exact
parity of the paths that use it as a candidate generator (flat L2, IVF coarse step) -- the
exact re-score decides, so ids/distances must still equal the oracle's bit for bit."""
import numpy as np
import pytest
import torch

import oracle
from lancedb_b200 import _native
from tests.util import queries, random_index

pytestmark = pytest.mark.gpu


def _bf16(a):
    return torch.from_numpy(a).to(torch.bfloat16).to(torch.float32).numpy()


@pytest.mark.parametrize("B,N,d", [(128, 256, 64), (1, 1, 8), (130, 300, 72), (1000, 5000, 768), (16, 70000, 128)])
def test_gemm_numerics(B, N, d):
    rng = np.random.default_rng(B + N + d)
    q = queries(rng, B, d); x = queries(rng, N, d)
    got = _native.debug_gemm(q, x)
    qb, xb = _bf16(q).astype(np.float64), _bf16(x).astype(np.float64)
    want = (x.astype(np.float64) ** 2).sum(1)[None, :] - 2.0 * qb @ xb.T
    scale = np.abs(want).max() + 1.0
    err = np.abs(got - want).max()
    assert err <= 2e-5 * scale * max(1.0, d / 256), f"max abs err {err} (scale {scale})"


@pytest.mark.parametrize("n,dim,B,k", [(20000, 128, 64, 10), (6000, 1536, 33, 10), (9000, 64, 200, 100)])
def test_flat_tensorcore_path_bit_exact(n, dim, B, k):
    rng = np.random.default_rng(n)
    v = queries(rng, n, dim)
    v[n // 2] = v[3]; v[n // 3] = v[3]                    # exact duplicates -> ties by row id
    rid = rng.permutation(n).astype(np.uint64)
    q = queries(rng, B, dim)
    q[0] = v[3]
    fl = _native.GpuFlat(v, row_ids=rid)
    gi, gd, gc = fl.search(q, k=k, metric="l2")
    oi, od, oc = oracle.flat_search(v, q, k=k, metric="l2", row_ids=rid, nthreads=8)
    fl.close()
    assert np.array_equal(gc, oc) and np.array_equal(gi, oi)
    assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))


def test_flat_tensorcore_filtered_epilogue_large_n():
    """N >= 256k takes the sampled-threshold + filtering-epilogue path (no dense score matrix)"""
    rng = np.random.default_rng(31)
    n, dim = 300_000, 64
    v = queries(rng, n, dim)
    v[200_000] = v[7]; v[299_999] = v[7]                  # ties across the sampled and unsampled parts
    q = queries(rng, 24, dim)
    q[0] = v[7]
    fl = _native.GpuFlat(v)
    for k in (10, 40):
        gi, gd, gc = fl.search(q, k=k)
        oi, od, oc = oracle.flat_search(v, q, k=k, nthreads=8)
        assert np.array_equal(gc, oc) and np.array_equal(gi, oi)
        assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))
    fl.close()


def test_coarse_filtered_path_large_nlist(monkeypatch):
    """nlist >= 4096: sampled threshold + filtering epilogue picks the probes; result must equal the oracle"""
    monkeypatch.setenv("LGPU_FORCE_TC_COARSE", "1")
    rng = np.random.default_rng(17)
    sizes = np.full(4200, 3, np.int64); sizes[::7] = 0
    ix = random_index(rng, dim=64, nlist=4200, m=8, sizes=sizes)
    q = queries(rng, 40, 64)
    gpu = _native.GpuIvfPq(ix)
    orc = oracle.OracleIndex.from_data(ix)
    for nprobes in (20, 50):
        gi, gd, gc = gpu.search(q, k=10, nprobes=nprobes)
        oi, od, oc = orc.search(q, k=10, nprobes=nprobes, nthreads=8)
        assert np.array_equal(gc, oc) and np.array_equal(gi, oi)
        assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))
    gpu.close()

@pytest.mark.parametrize("metric", ["l2", "cosine"])
def test_coarse_sampled_list_path(metric, monkeypatch):
    """Many lists (default from nlist 8192, here forced at 4200): dense scores of every 8th centroid give a per-query
    bound, the full GEMM's filtering epilogue appends (column, score) of the columns under it to a list, and the
    finishing kernel works on the list.  Probe sets -- hence results -- must equal the oracle's, including for queries
    that sit ON a centroid (distance 0, ties in the band) and with clustered centroids (a loose sample bound)."""
    monkeypatch.setenv("LGPU_FORCE_TC_COARSE", "1")
    monkeypatch.setenv("LGPU_COARSE_LIST_MIN", "1024")
    rng = np.random.default_rng(19)
    sizes = np.full(4200, 3, np.int64); sizes[::5] = 0
    ix = random_index(rng, dim=64, nlist=4200, m=8, metric=metric, sizes=sizes)
    # the first 600 centroids in one tight cluster (consecutive ids, like a hierarchical trainer leaves them)
    ix.centroids[:600] = ix.centroids[0] + 0.01 * rng.standard_normal((600, 64)).astype(np.float32)
    if metric == "cosine":
        ix.centroids /= np.linalg.norm(ix.centroids, axis=1, keepdims=True)
    q = queries(rng, 48, 64)
    q[:6] = ix.centroids[[5, 100, 599, 700, 2000, 4199]]
    q[6:9] = ix.centroids[3] * np.float32(1.0001)
    gpu = _native.GpuIvfPq(ix)
    orc = oracle.OracleIndex.from_data(ix)
    for nprobes in (1, 20, 50):
        gi, gd, gc = gpu.search(q, k=10, nprobes=nprobes)
        oi, od, oc = orc.search(q, k=10, nprobes=nprobes, nthreads=8)
        assert np.array_equal(gc, oc) and np.array_equal(gi, oi), nprobes
        assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))
    gpu.close()

def test_coarse_sampled_list_path_default_switch():
    """nlist 8192 with B x nlist >= 1M takes the sampled-bound + list-epilogue coarse step by default (no environment
    switch): probe sets, hence ids and distance bits, equal the oracle's; empty partitions interleaved."""
    rng = np.random.default_rng(29)
    sizes = np.full(8192, 2, np.int64); sizes[::3] = 0
    ix = random_index(rng, dim=64, nlist=8192, m=8, sizes=sizes)
    q = queries(rng, 160, 64)
    q[:4] = ix.centroids[[0, 8, 4097, 8191]]                     # on a sampled / unsampled centroid
    gpu = _native.GpuIvfPq(ix)
    gi, gd, gc = gpu.search(q, k=10, nprobes=20)
    oi, od, oc = oracle.OracleIndex.from_data(ix).search(q, k=10, nprobes=20, nthreads=8)
    gpu.close()
    assert np.array_equal(gc, oc) and np.array_equal(gi, oi)
    assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))


def test_coarse_default_path_c2_shape():
    """B x nlist >= 1M with nlist >= 1024 (BASELINE config 2's coarse shape) takes the tensor-core
    shortlist by default; probes and final results must equal the oracle."""
    rng = np.random.default_rng(23)
    sizes = rng.integers(0, 9, 1024).astype(np.int64)
    ix = random_index(rng, dim=64, nlist=1024, m=8, sizes=sizes)
    q = queries(rng, 1100, 64)
    gpu = _native.GpuIvfPq(ix)
    gi, gd, gc = gpu.search(q, k=10, nprobes=20)
    oi, od, oc = oracle.OracleIndex.from_data(ix).search(q, k=10, nprobes=20, nthreads=8)
    gpu.close()
    assert np.array_equal(gc, oc) and np.array_equal(gi, oi)
    assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))


def test_flat_tensorcore_clustered_fallback():
    """near-duplicate rows make the error band overflow the shortlist -> exact fix-up path"""
    rng = np.random.default_rng(5)
    base = queries(rng, 1, 64)
    v = (base + 1e-3 * rng.standard_normal((8192, 64))).astype(np.float32)
    q = (base + 1e-3 * rng.standard_normal((16, 64))).astype(np.float32)
    fl = _native.GpuFlat(v)
    gi, gd, gc = fl.search(q, k=10)
    oi, od, oc = oracle.flat_search(v, q, k=10, nthreads=8)
    fl.close()
    assert np.array_equal(gi, oi) and np.array_equal(gd.view(np.uint32), od.view(np.uint32))


@pytest.mark.parametrize("metric", ["l2", "cosine"])
@pytest.mark.parametrize("nprobes", [20, 37])
def test_coarse_tensorcore_path_bit_exact(metric, nprobes, monkeypatch):
    monkeypatch.setenv("LGPU_FORCE_TC_COARSE", "1")
    rng = np.random.default_rng(11)
    ix = random_index(rng, dim=128, nlist=700, m=16, metric=metric, n=30000)
    q = queries(rng, 70, 128)
    gpu = _native.GpuIvfPq(ix)
    orc = oracle.OracleIndex.from_data(ix)
    parts, dists = gpu.debug_coarse(q, nprobes)              # exact reference path
    gi, gd, gc = gpu.search(q, k=10, nprobes=nprobes)        # tensor-core shortlist path inside
    oi, od, oc = orc.search(q, k=10, nprobes=nprobes, nthreads=8)
    gpu.close()
    for i in range(q.shape[0]):
        qn = oracle.normalize(q[i]) if metric == "cosine" else q[i]
        op, _, _ = orc.find_partitions(qn, nprobes)
        assert np.array_equal(parts[i], op)
    assert np.array_equal(gc, oc) and np.array_equal(gi, oi)
    assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))
