"""IVF_PQ index container and trainer.

`IvfPqIndexData` is the plain-array form of a Lance IVF_PQ index: exactly the arrays
the reference's search path consumes after `prewarm_index`
(rust/lancedb/src/table.rs:3283-3286) -- IVF centroids, PQ codebook, per-partition
transposed PQ codes and row ids.  The same arrays are handed to the C-ABI
(`lgpu_index_open`, include/lancedb_b200.h) and to the CPU oracle.

`train_ivf_pq` mirrors the *parameters* of `Index::IvfPq`
(rust/lancedb/src/index/vector.rs:266-319, rust/lancedb/src/table/create_index.rs:68-102,
283-303): num_partitions, num_sub_vectors (default dim/16, else dim/8, else 1),
num_bits = 8, sample_rate = 256, max_iterations = 50, distance_type.  Training itself
lives in the un-vendored lance crate; this is a plain k-means / PQ trainer written with
torch ops (CPU or CUDA), or, with `native_passes` (the `accelerator="cuda"` build), the library's own
training kernels (csrc/kmeans.cu) -- index *quality* is not on the hot path, and both the CUDA
path and the oracle consume the identical arrays it produces.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

METRICS = ("l2", "cosine", "dot")


def suggested_num_sub_vectors(dim: int) -> int:
    """rust/lancedb/src/index/vector.rs:306-319"""
    if dim % 16 == 0:
        return dim // 16
    if dim % 8 == 0:
        return dim // 8
    return 1


def suggested_num_partitions(num_rows: int, target_partition_size: int = 8192) -> int:
    """Default IVF sizing: 16384 rows => 2 partitions
    (rust/lancedb/src/table/create_index.rs:734-795)."""
    return max(1, num_rows // target_partition_size)


@dataclass
class IvfPqIndexData:
    dim: int
    nlist: int
    m: int
    metric: str
    centroids: np.ndarray      # f32 [nlist, dim]
    codebook: np.ndarray       # f32 [m, 256, dim/m]
    part_offsets: np.ndarray   # u64 [nlist+1]
    codes_t: np.ndarray        # u8 flat; partition p at [off[p]*m, off[p+1]*m) as [m][n_p]
    row_ids: np.ndarray        # u64 [n] in partition order
    vectors: Optional[np.ndarray] = None   # f32 [n, dim] partition order (refine), optional

    @property
    def nrows(self) -> int:
        return int(self.row_ids.size)

    @property
    def dsub(self) -> int:
        return self.dim // self.m

    def validate(self) -> None:
        assert self.metric in METRICS
        assert self.dim % self.m == 0
        assert self.centroids.shape == (self.nlist, self.dim) and self.centroids.dtype == np.float32
        assert self.codebook.shape == (self.m, 256, self.dsub) and self.codebook.dtype == np.float32
        assert self.part_offsets.shape == (self.nlist + 1,) and self.part_offsets.dtype == np.uint64
        assert int(self.part_offsets[-1]) == self.nrows
        assert self.codes_t.dtype == np.uint8 and self.codes_t.size == self.nrows * self.m
        assert self.row_ids.dtype == np.uint64

    def partition_codes(self, p: int) -> np.ndarray:
        a, b = int(self.part_offsets[p]), int(self.part_offsets[p + 1])
        return self.codes_t[a * self.m:b * self.m].reshape(self.m, b - a)

    def shard(self, rank: int, world: int) -> "IvfPqIndexData":
        """Partition-sharded view for multi-GPU search: centroids and
        codebook replicated, each partition's codes/row ids owned by exactly one rank
        (greedy size-balanced); non-owned partitions become empty."""
        sizes = np.diff(self.part_offsets.astype(np.int64))
        owner = assign_partitions(sizes, world)
        keep = owner == rank
        new_sizes = np.where(keep, sizes, 0)
        new_off = np.zeros(self.nlist + 1, np.uint64)
        new_off[1:] = np.cumsum(new_sizes)
        codes, rids, vecs = [], [], []
        for p in np.nonzero(keep)[0]:
            a, b = int(self.part_offsets[p]), int(self.part_offsets[p + 1])
            codes.append(self.codes_t[a * self.m:b * self.m])
            rids.append(self.row_ids[a:b])
            if self.vectors is not None:
                vecs.append(self.vectors[a:b])
        cat = lambda xs, dt, shape: (np.concatenate(xs) if xs else np.zeros(shape, dt))
        return IvfPqIndexData(
            self.dim, self.nlist, self.m, self.metric, self.centroids, self.codebook, new_off,
            cat(codes, np.uint8, (0,)), cat(rids, np.uint64, (0,)),
            cat(vecs, np.float32, (0, self.dim)) if self.vectors is not None else None)


def assign_partitions(sizes: np.ndarray, world: int) -> np.ndarray:
    """Greedy size-balanced bin packing of partitions onto ranks (largest first)."""
    owner = np.zeros(len(sizes), np.int64)
    load = np.zeros(world, np.int64)
    for p in np.argsort(-sizes, kind="stable"):
        r = int(np.argmin(load))
        owner[p] = r
        load[r] += int(sizes[p])
    return owner


# --------------------------------------------------------------------------------------
def _kmeans(x, k, iters, gen, chunk=1 << 16):
    """Lloyd k-means with torch ops; x [n, d] float32 (any device)."""
    import torch
    n = x.shape[0]
    perm = torch.randperm(n, generator=gen, device="cpu")[:k].to(x.device)
    c = x[perm].clone()
    if c.shape[0] < k:      # fewer points than centroids: pad with jittered copies
        extra = x[torch.randint(0, n, (k - c.shape[0],), generator=gen, device="cpu").to(x.device)]
        c = torch.cat([c, extra + 1e-3 * torch.randn(extra.shape, generator=gen).to(x.device)])
    assign = torch.empty(n, dtype=torch.long, device=x.device)
    for _ in range(max(1, iters)):
        cn = (c * c).sum(1)
        for s in range(0, n, chunk):
            xs = x[s:s + chunk]
            assign[s:s + chunk] = (cn[None, :] - 2.0 * (xs @ c.T)).argmin(1)
        sums = torch.zeros_like(c).index_add_(0, assign, x)
        cnt = torch.bincount(assign, minlength=k).to(x.dtype)
        empty = cnt == 0
        c = torch.where(empty[:, None], c, sums / cnt.clamp(min=1)[:, None])
        if empty.any():     # re-seed empty clusters from random points
            ne = int(empty.sum())
            idx = torch.randint(0, n, (ne,), generator=gen, device="cpu").to(x.device)
            c[empty] = x[idx]
    return c


def _init_rows(x, k, gen):
    """k initial centres: distinct random rows (jittered copies when there are fewer rows than centres)."""
    import torch
    n = x.shape[0]
    c = x[torch.randperm(n, generator=gen, device="cpu")[:k].to(x.device)].clone()
    if c.shape[0] < k:
        extra = x[torch.randint(0, n, (k - c.shape[0],), generator=gen, device="cpu").to(x.device)]
        c = torch.cat([c, extra + 1e-3 * torch.randn(extra.shape, generator=gen).to(x.device)])
    return c


def _assign(x, c, chunk=1 << 16, metric="l2"):
    """Partition of every row, ranked like the search's coarse step: L2 for l2/cosine, 1 - x.c for dot."""
    import torch
    n = x.shape[0]
    out = torch.empty(n, dtype=torch.long, device=x.device)
    cn = (c * c).sum(1)
    for s in range(0, n, chunk):
        xc = x[s:s + chunk] @ c.T
        out[s:s + chunk] = (-xc).argmin(1) if metric == "dot" else (cn[None, :] - 2.0 * xc).argmin(1)
    return out


def _batched_kmeans(x, k, iters, gen):
    """x [m, ns, dsub] -> centroids [m, k, dsub], all sub-spaces at once."""
    import torch
    m, ns, _ = x.shape
    idx = torch.stack([torch.randperm(ns, generator=gen, device="cpu")[:k] for _ in range(m)]).to(x.device)
    if idx.shape[1] < k:
        pad = torch.randint(0, ns, (m, k - idx.shape[1]), generator=gen, device="cpu").to(x.device)
        idx = torch.cat([idx, pad], 1)
    c = torch.gather(x, 1, idx[:, :, None].expand(-1, -1, x.shape[2])).clone()
    for _ in range(max(1, iters)):
        cn = (c * c).sum(2)                                       # [m, k]
        a = (cn[:, None, :] - 2.0 * torch.bmm(x, c.transpose(1, 2))).argmin(2)   # [m, ns]
        sums = torch.zeros_like(c).scatter_add_(1, a[:, :, None].expand(-1, -1, x.shape[2]), x)
        cnt = torch.zeros(m, k, device=x.device, dtype=x.dtype).scatter_add_(
            1, a, torch.ones_like(a, dtype=x.dtype))
        c = torch.where((cnt == 0)[:, :, None], c, sums / cnt.clamp(min=1)[:, :, None])
    return c


def train_ivf_pq(vectors, *, num_partitions: Optional[int] = None, num_sub_vectors: Optional[int] = None,
                 distance_type: str = "l2", sample_rate: int = 256, max_iterations: int = 50,
                 row_ids: Optional[np.ndarray] = None, keep_vectors: bool = False, seed: int = 45,
                 device: Optional[str] = None, encode_chunk: int = 1 << 16,
                 native_passes: bool = False) -> IvfPqIndexData:
    """Train IVF centroids + residual PQ codebooks and encode every row.

    vectors: [n, dim] float32 (numpy or torch).  Returns the plain-array index.
    native_passes: run the two passes over every row (IVF assignment, PQ encoding) through the C ABI
    (`lgpu_ivf_assign` / `lgpu_pq_encode`, csrc/build.cu) -- the search kernels' own arithmetic, so a row
    always lands in the partition its own vector probes first -- instead of torch's GEMM-form argmin.
    With native_passes the k-means training loops run in the library too (`lgpu_kmeans_train` / `lgpu_pq_train`,
    csrc/kmeans.cu); otherwise they are torch ops.
    """
    import torch
    metric = distance_type.lower()
    if metric not in METRICS:
        raise ValueError(f"unknown distance_type {distance_type!r}")
    x = torch.as_tensor(vectors, dtype=torch.float32)
    if device is not None:
        x = x.to(device)
    if x.device.type == "cpu" and torch.get_num_threads() > 16:
        # many-core hosts with a small CPU quota: a 100+ thread OpenMP team spins instead of working
        prev = torch.get_num_threads()
        torch.set_num_threads(16)
        try:
            return train_ivf_pq(x, num_partitions=num_partitions, num_sub_vectors=num_sub_vectors,
                                distance_type=distance_type, sample_rate=sample_rate, max_iterations=max_iterations,
                                row_ids=row_ids, keep_vectors=keep_vectors, seed=seed, device=None,
                                encode_chunk=encode_chunk, native_passes=native_passes)
        finally:
            torch.set_num_threads(prev)
    n, dim = x.shape
    nlist = int(num_partitions or suggested_num_partitions(n))
    m = int(num_sub_vectors or suggested_num_sub_vectors(dim))
    if dim % m:
        raise ValueError(f"num_sub_vectors {m} does not divide dimension {dim}")
    dsub = dim // m
    gen = torch.Generator(device="cpu").manual_seed(seed)
    raw = x
    if metric == "cosine":       # index stores normalised vectors; search is L2 on them
        x = x / x.norm(dim=1, keepdim=True).clamp(min=1e-30)

    ns = min(n, sample_rate * nlist)
    samp = x[torch.randperm(n, generator=gen, device="cpu")[:ns].to(x.device)] if ns < n else x
    dev_index = x.device.index or 0 if x.device.type == "cuda" else 0
    if native_passes:
        # accelerator path: the Lloyd loops run in the library's own kernels (csrc/kmeans.cu through
        # lgpu_kmeans_train): no torch op inside the loop, only the random initial sample is drawn here
        from . import _native
        samp_np = samp.detach().cpu().numpy()
        init = _init_rows(samp, nlist, gen).cpu().numpy()
        centroids = torch.as_tensor(_native.kmeans_train(samp_np, init, max_iterations, dev_index), device=x.device)
    else:
        centroids = _kmeans(samp, nlist, max_iterations, gen)
    if native_passes:
        raw_np = raw.detach().cpu().numpy()
        assign = torch.as_tensor(_native.ivf_assign(centroids.cpu().numpy(), raw_np, metric, dev_index).astype(np.int64),
                                 device=x.device)
    else:
        assign = _assign(x, centroids, metric=metric)

    # PQ codebooks: residuals for l2/cosine, raw vectors for dot
    nps = min(n, max(256, sample_rate) * 256)
    pidx = torch.randperm(n, generator=gen, device="cpu")[:nps].to(x.device)
    ps = x[pidx] - centroids[assign[pidx]] if metric != "dot" else x[pidx]
    if native_passes:
        init_cb = torch.stack([ps[torch.randperm(nps, generator=gen, device="cpu")[:256].to(x.device) if nps >= 256
                                  else torch.randint(0, nps, (256,), generator=gen, device="cpu").to(x.device)]
                               .reshape(256, m, dsub)[:, i, :] for i in range(m)])              # [m, 256, dsub]
        codebook = torch.as_tensor(_native.pq_train(ps.detach().cpu().numpy(), init_cb.cpu().numpy(), max_iterations,
                                                    dev_index), device=x.device)
    else:
        ps3 = ps.reshape(nps, m, dsub).transpose(0, 1).contiguous()      # [m, nps, dsub]
        codebook = _batched_kmeans(ps3, 256, max_iterations, gen)        # [m, 256, dsub]

    cbn = (codebook * codebook).sum(2)                                     # [m, 256]
    codes = torch.empty((n, m), dtype=torch.uint8, device=x.device)
    if native_passes:
        codes = torch.as_tensor(_native.pq_encode(centroids.cpu().numpy(), codebook.cpu().numpy(), raw_np,
                                                  assign.cpu().numpy().astype(np.uint32), metric, dev_index),
                                device=x.device)
    for s in range(0, n if not native_passes else 0, encode_chunk):
        xs = x[s:s + encode_chunk]
        r = xs - centroids[assign[s:s + encode_chunk]] if metric != "dot" else xs
        r = r.reshape(-1, m, dsub).transpose(0, 1)                         # [m, c, dsub]
        d = cbn[:, None, :] - 2.0 * torch.bmm(r, codebook.transpose(1, 2))
        codes[s:s + encode_chunk] = d.argmin(2).transpose(0, 1).to(torch.uint8)

    order = torch.argsort(assign, stable=True)                            # ascending row id per partition
    sizes = torch.bincount(assign, minlength=nlist).cpu().numpy().astype(np.int64)
    part_offsets = np.zeros(nlist + 1, np.uint64)
    part_offsets[1:] = np.cumsum(sizes)
    codes_sorted = codes[order].cpu().numpy()                              # [n, m] partition order
    codes_t = np.empty(n * m, np.uint8)
    for p in range(nlist):
        a, b = int(part_offsets[p]), int(part_offsets[p + 1])
        if b > a:
            codes_t[a * m:b * m] = codes_sorted[a:b].T.reshape(-1)
    order_np = order.cpu().numpy()
    rid = np.arange(n, dtype=np.uint64) if row_ids is None else np.asarray(row_ids, np.uint64)
    data = IvfPqIndexData(
        dim=dim, nlist=nlist, m=m, metric=metric,
        centroids=centroids.cpu().numpy().astype(np.float32),
        codebook=codebook.cpu().numpy().astype(np.float32),
        part_offsets=part_offsets, codes_t=codes_t, row_ids=rid[order_np],
        vectors=(raw[order].cpu().numpy().astype(np.float32) if keep_vectors else None))
    data.validate()
    return data
