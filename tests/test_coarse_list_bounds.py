"""Host restatement of the coarse step at many lists (api.cu: sampled bound -> list epilogue -> finishing kernel on the
list; gemm.cu, dist.cu) checked on the CPU: with S[x] = |c_x|^2 - 2 bf16(q).bf16(c_x) (f32 accumulation of bf16
products) and E_q = 2^-7 (1 + 2^-8) |q| cmax + 4 d 2^-24 (|q| + cmax)^2,

  (1) |S[x] + |q|^2 - d*(q, x)| <= E_q for every centroid (the band the kernels rely on), d* the oracle's distance;
  (2) thr1 = (k-th smallest S over every 8th centroid) + 2 E_q admits every true probe into the list;
  (3) thr2 = (k-th smallest S of the LIST, found to within E_q / 4 from above) + 2 E_q still does, and the k-th
      smallest of the list equals the k-th smallest of the whole row;
  (4) the re-scored set is small: a few times k, not nlist / 8.

Random centroids, centroids in tight clusters with consecutive ids (what a hierarchical trainer leaves), queries on a
centroid, and scaled data.  The GPU tests (tests/test_gpu_tensorcore.py) check the kernels' results; this pins the
algebra they implement."""
import numpy as np
import pytest

import oracle
from tests.util import queries, random_index

F = np.float32
STRIDE = 8


def _bf16(x):
    """round-to-nearest-even f32 -> bf16 -> f32 (gemm.cu to_bf16_norm_kernel)"""
    u = np.ascontiguousarray(x, F).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(F).reshape(np.shape(x))


def _scores(q, C):
    qb, Cb = _bf16(q).astype(np.float64), _bf16(C).astype(np.float64)
    dot = (Cb @ qb).astype(F)                               # products of bf16 are exact in f32; the sum is f32
    cn2 = (C.astype(np.float64) ** 2).sum(1).astype(F)
    return (cn2 - F(2) * dot).astype(F)


def _check(ix, q, k):
    C = ix.centroids
    orc = oracle.OracleIndex.from_data(ix)
    qn = oracle.normalize(q) if ix.metric == "cosine" else q.astype(F)
    dstar = orc.find_partitions(qn, ix.nlist)[2].astype(np.float64)      # exact distance to every centroid
    S = _scores(qn, C)
    qn2 = float((qn.astype(np.float64) ** 2).sum())
    cmax = float(np.sqrt((C.astype(np.float64) ** 2).sum(1).max())) * 1.0001
    E = 0.0078125 * 1.00390625 * np.sqrt(qn2) * cmax + 4.0 * ix.dim * 5.9604645e-8 * (np.sqrt(qn2) + cmax) ** 2
    assert np.abs(S.astype(np.float64) + qn2 - dstar).max() <= E                      # (1)
    truth = np.lexsort((np.arange(ix.nlist), dstar))[:k]
    sample = S[::STRIDE][: ix.nlist // STRIDE]
    thr1 = np.sort(sample)[k - 1] + 2 * E
    in_list = S <= thr1
    assert in_list[truth].all()                                                        # (2)
    lst = np.sort(S[in_list])
    assert lst[k - 1] == np.sort(S)[k - 1]
    hi = lst[k - 1] + 0.25 * E                                                         # bisection stops within E / 4
    cand = in_list & (S <= hi + 2 * E)
    assert cand[truth].all()                                                           # (3)
    return int(in_list.sum()), int(cand.sum())


@pytest.mark.parametrize("metric", ["l2", "cosine"])
@pytest.mark.parametrize("scale", [1.0, 1e-2, 40.0])
def test_sampled_bound_and_list_threshold_hold_the_true_probes(metric, scale):
    rng = np.random.default_rng(41)
    nlist, dim, k = 2048, 96, 20
    ix = random_index(rng, dim=dim, nlist=nlist, m=8, metric=metric, sizes=np.ones(nlist, np.int64), scale=scale)
    ix.centroids[:300] = ix.centroids[0] + F(0.01 * scale) * rng.standard_normal((300, dim)).astype(F)   # one tight blob
    if metric == "cosine":
        ix.centroids /= np.linalg.norm(ix.centroids, axis=1, keepdims=True)
    qs = queries(rng, 12, dim, scale=scale)
    qs[0] = ix.centroids[5]; qs[1] = ix.centroids[1000]; qs[2] = ix.centroids[7] * F(1.001)
    listed, rescored = zip(*[_check(ix, q, k) for q in qs])
    # the list is what the sample bound costs (about k * STRIDE + band), the re-scored set what the band costs
    assert max(rescored) <= max(listed)
    assert np.median(rescored[3:]) <= 12 * k                                            # (4) random queries
