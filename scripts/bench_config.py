#!/usr/bin/env python
"""Throughput + oracle spot-check of the IVF_PQ path on an arbitrary configuration with a synthetic
(untrained, uniform-partition) index -- used for BASELINE.json configs[2] (10M x 768, nlist 4096,
nprobes 50, k 100, cosine, batch 4096) and for a configs[4]-shaped per-GPU shard.  Index contents are
random (parity and throughput do not depend on index quality); prints one JSON line."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lancedb_b200 import _native
from lancedb_b200.index import IvfPqIndexData

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--nlist", type=int, default=4096)
ap.add_argument("--m", type=int, default=96)
ap.add_argument("--nprobes", type=int, default=50)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--metric", default="cosine")
ap.add_argument("--owned", type=float, default=1.0, help="fraction of partitions held by this GPU (shard emulation)")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--check", type=int, default=8)
a = ap.parse_args()
rng = np.random.default_rng(7)
dsub = a.dim // a.m
base = a.n // a.nlist
sizes = rng.integers(int(base * 0.7), int(base * 1.3) + 1, a.nlist).astype(np.int64)
if a.owned < 1.0:
    sizes[rng.random(a.nlist) >= a.owned] = 0
n = int(sizes.sum())
cent = rng.standard_normal((a.nlist, a.dim), dtype=np.float32)
if a.metric == "cosine":
    cent /= np.linalg.norm(cent, axis=1, keepdims=True)
cb = (rng.standard_normal((a.m, 256, dsub), dtype=np.float32) * 0.3).astype(np.float32)
off = np.zeros(a.nlist + 1, np.uint64); off[1:] = np.cumsum(sizes)
codes = rng.integers(0, 256, size=n * a.m, dtype=np.uint8)
ids = np.arange(n, dtype=np.uint64)
ix = IvfPqIndexData(a.dim, a.nlist, a.m, a.metric, cent, cb, off, codes, ids, None)
t0 = time.time()
gpu = _native.GpuIvfPq(ix, with_vectors=False)
open_s = time.time() - t0
g = torch.Generator().manual_seed(3)
q = torch.randn(2, a.batch, a.dim, generator=g)
dq = q.cuda()
oi = torch.empty(a.batch, a.k, dtype=torch.int64, device="cuda"); od = torch.empty(a.batch, a.k, device="cuda")
oc = torch.empty(a.batch, dtype=torch.int32, device="cuda")
p = _native.make_params(k=a.k, nprobes=a.nprobes)
st = torch.cuda.current_stream().cuda_stream
for i in range(2):
    gpu.search_device(dq[i % 2].data_ptr(), a.batch, p, oi.data_ptr(), od.data_ptr(), oc.data_ptr(), st)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
for i in range(a.steps):
    ev[i][0].record()
    gpu.search_device(dq[i % 2].data_ptr(), a.batch, p, oi.data_ptr(), od.data_ptr(), oc.data_ptr(), st)
    ev[i][1].record()
torch.cuda.synchronize()
ms = float(np.mean([s.elapsed_time(e) for s, e in ev]))
_native.set_profiling(True)
gpu.search_device(dq[(a.steps - 1) % 2].data_ptr(), a.batch, p, oi.data_ptr(), od.data_ptr(), oc.data_ptr(), st)
stage = _native.last_stage_ms(); code_bytes = _native.last_scanned_code_bytes(); fstats = _native.last_filter_stats()
_native.set_profiling(False)
out = {"config": vars(a), "rows": n, "index_open_s": open_s, "ms_per_batch": ms, "qps": a.batch / (ms / 1e3),
       "stage_ms": stage, "scan_algorithmic_GBps": code_bytes / (stage["scan"] / 1e3) / 1e9,
       "index_device_bytes": gpu.device_bytes(), "filter_stats": fstats}
if a.check:
    import oracle
    gi = oi.cpu().numpy().view(np.uint64); gd = od.cpu().numpy(); gc = oc.cpu().numpy()
    last = (a.steps - 1) % 2
    ri, rd, rc = oracle.OracleIndex.from_data(ix).search(q[last, :a.check].numpy(), k=a.k, nprobes=a.nprobes,
                                                         nthreads=os.cpu_count())
    out["oracle_check"] = bool(np.array_equal(gi[:a.check], ri) and np.array_equal(gd[:a.check].view(np.uint32), rd.view(np.uint32))
                               and np.array_equal(gc[:a.check].view(np.uint32), rc))
print(json.dumps(out))
