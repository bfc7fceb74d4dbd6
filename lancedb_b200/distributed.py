"""Partition-sharded IVF_PQ search across GPUs: one process per GPU.

IVF partitions are the shardable unit: centroids and the PQ codebook are replicated, each
partition's codes + row ids live on exactly one rank (`IvfPqIndexData.shard`).  Every rank
gets the full query batch, runs the coarse step redundantly (so all ranks agree on the
probe set without communicating), scans only the probed partitions it owns, and produces a
local top-k.  The single exchange step is one all-gather of `[B, k]` (row id u64, distance
f32) per rank over NCCL/NVLink, consumed directly by the merge kernel
(`lgpu_merge_topk_device`), which re-selects the global top-k by (_distance, _rowid).
The reference has no equivalent (LanceDB OSS is single-process); it is only
needed where the index exceeds one GPU's HBM -- otherwise replicas with the batch split are
faster (zero communication), which is what bench.py measures by default.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _native
from .index import IvfPqIndexData


def gather_shape(world: int, B: int, k: int):
    """Layout of the gathered candidate lists: [world][B][k] (rank-major), which is what
    lgpu_merge_topk_device expects (inner = k, outer stride = B*k)."""
    return (world, B, k)


class ShardedIvfPq:
    def __init__(self, data: IvfPqIndexData, *, group=None, device: Optional[int] = None):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (one process per GPU)")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.cuda.current_device() if device is None else device
        self.dim = data.dim
        self.local = _native.GpuIvfPq(data.shard(self.rank, self.world), device=self.device, with_vectors=False)

    def close(self):
        self.local.close()

    def search_device(self, d_q, k: int = 10, nprobes: int = 20, lower=None, upper=None):
        """d_q: [B, dim] float32 CUDA tensor (identical on every rank).  Returns CUDA tensors
        (ids int64 holding the u64 row ids, dist float32, count int32), identical on every rank."""
        import torch
        import torch.distributed as dist
        B = d_q.shape[0]
        dev = d_q.device
        stream = torch.cuda.current_stream().cuda_stream
        ids = torch.empty(B, k, dtype=torch.int64, device=dev)
        dst = torch.empty(B, k, dtype=torch.float32, device=dev)
        cnt = torch.empty(B, dtype=torch.int32, device=dev)
        p = _native.make_params(k=k, nprobes=nprobes, lower=lower, upper=upper)
        self.local.search_device(d_q.data_ptr(), B, p, ids.data_ptr(), dst.data_ptr(), cnt.data_ptr(), stream)
        g_ids = torch.empty(gather_shape(self.world, B, k), dtype=torch.int64, device=dev)
        g_dst = torch.empty(gather_shape(self.world, B, k), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(g_ids.view(-1, k), ids, group=self.group)
        dist.all_gather_into_tensor(g_dst.view(-1, k), dst, group=self.group)
        _native.merge_topk_device(self.device, self.world, B, k, g_ids.data_ptr(), g_dst.data_ptr(),
                                  ids.data_ptr(), dst.data_ptr(), cnt.data_ptr(), stream)
        return ids, dst, cnt

    def search(self, queries, k: int = 10, nprobes: int = 20, lower=None, upper=None):
        import torch
        q = torch.as_tensor(np.ascontiguousarray(queries, np.float32).reshape(-1, self.dim)).cuda(self.device)
        ids, dst, cnt = self.search_device(q, k, nprobes, lower, upper)
        torch.cuda.synchronize()
        return (ids.cpu().numpy().view(np.uint64), dst.cpu().numpy(), cnt.cpu().numpy().view(np.uint32))
