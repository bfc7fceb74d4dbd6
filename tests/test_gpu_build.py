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


def test_kmeans_training_kernels_match_a_torch_lloyd_run():
    """lgpu_kmeans_train (csrc/kmeans.cu): from the same initial centres and the same number of iterations the
    GPU Lloyd run must end within 1 % of the inertia of a plain torch Lloyd run; one iteration
    from the same centres must give the same centres up to f32 rounding of the means."""
    import torch
    from lancedb_b200 import _native
    rng = np.random.default_rng(61)
    cen = rng.standard_normal((24, 64)).astype(np.float32) * 3
    x = (cen[rng.integers(0, 24, 30000)] + rng.standard_normal((30000, 64))).astype(np.float32)
    for k in (16, 300):                                   # 300: the tensor-core assignment (k >= 256)
        init = x[rng.choice(len(x), k, replace=False)].copy()

        def lloyd(c, iters):
            xt, ct = torch.from_numpy(x), torch.from_numpy(c.copy())
            for _ in range(iters):
                a = torch.cdist(xt.double(), ct.double()).argmin(1)      # exact nearest centre, as the kernels assign
                for j in range(k):
                    sel = a == j
                    if sel.any():
                        ct[j] = xt[sel].double().mean(0).float()
            d = torch.cdist(xt.double(), ct.double()).min(1).values
            return ct.numpy(), float((d.double() ** 2).sum())
        one_ref, _ = lloyd(init, 1)
        one_gpu = _native.kmeans_train(x, init, 1)
        assert np.allclose(one_gpu, one_ref, rtol=1e-4, atol=1e-4)
        ref_c, ref_inertia = lloyd(init, 12)
        gpu_c, gpu_inertia = _native.kmeans_train(x, init, 12, want_inertia=True)
        assert abs(gpu_inertia - ref_inertia) <= 0.01 * ref_inertia, (k, gpu_inertia, ref_inertia)


def test_pq_training_kernel_reduces_quantisation_error():
    from lancedb_b200 import _native
    rng = np.random.default_rng(62)
    x = rng.standard_normal((20000, 32)).astype(np.float32)
    m, dsub = 4, 8
    init = np.stack([x[rng.choice(len(x), 256, replace=False)].reshape(256, m, dsub)[:, i, :] for i in range(m)])

    def qerr(cb):
        xs = x.reshape(-1, m, dsub)
        return sum(float(((xs[:, i, None, :] - cb[i][None, :, :]) ** 2).sum(2).min(1).sum()) for i in range(m))
    e0 = qerr(init)
    cb1 = _native.pq_train(x, init, 1)
    cb10 = _native.pq_train(x, init, 10)
    assert qerr(cb1) < e0 and qerr(cb10) < qerr(cb1) * 0.999
    # one Lloyd step from `init`, restated with numpy
    xs = x.reshape(-1, m, dsub)
    for i in range(m):
        a = ((xs[:, i, None, :] - init[i][None, :, :]) ** 2).sum(2).argmin(1)
        for c in range(0, 256, 37):
            if (a == c).any():
                assert np.allclose(cb1[i, c], xs[a == c, i].astype(np.float64).mean(0), rtol=1e-4, atol=1e-5)


def test_create_index_with_accelerator_trains_on_the_gpu_and_matches_the_oracle():
    import lancedb_b200 as lancedb
    rng = np.random.default_rng(63)
    x = rng.standard_normal((8000, 64)).astype(np.float32)
    t = lancedb.connect("memory://").create_table("v", {"vector": x, "id": np.arange(8000)})
    n0 = _native_launches()
    t.create_index(metric="l2", num_partitions=16, num_sub_vectors=8, max_iterations=5, accelerator="cuda")
    assert _native_launches() > n0
    q = rng.standard_normal((4, 64)).astype(np.float32)
    orc = oracle.OracleIndex.from_data(t._index_data["vector"])
    oi, od, oc = orc.search(q, k=10, nprobes=6)
    for i in range(4):
        out = t.search(q[i]).nprobes(6).limit(10).with_row_id(True).to_arrow()
        assert out["_rowid"].to_pylist() == [int(v) for v in oi[i, :10]]


def _native_launches():
    from lancedb_b200 import _native
    return _native.kernel_launch_count()
