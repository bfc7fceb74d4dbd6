#!/usr/bin/env python
"""Flat brute-force L2 (BASELINE.json configs[3]: 1M x 1536 f32, batch 1024, k 10) through the
tensor-core shortlist + exact re-score path; prints one JSON line.  Not the headline bench
(bench.py is); used to measure the GEMM kernel against the bf16 tensor roofline."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lancedb_b200 import _native

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--dim", type=int, default=1536)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--check", type=int, default=16, help="queries verified against the CPU oracle")
a = ap.parse_args()
g = torch.Generator().manual_seed(42)
x = torch.randn(a.n, a.dim, generator=g).numpy()
q = torch.randn(4, a.batch, a.dim, generator=g)
fl = _native.GpuFlat(x)
dq = q.cuda()
ids = torch.empty(a.batch, a.k, dtype=torch.int64, device="cuda"); dist = torch.empty(a.batch, a.k, device="cuda")
cnt = torch.empty(a.batch, dtype=torch.int32, device="cuda")
p = _native.make_params(k=a.k)
st = torch.cuda.current_stream().cuda_stream
for i in range(2):
    fl.search_device("l2", dq[i % 4].data_ptr(), a.batch, p, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), st)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
for i in range(a.steps):
    ev[i][0].record()
    fl.search_device("l2", dq[i % 4].data_ptr(), a.batch, p, ids.data_ptr(), dist.data_ptr(), cnt.data_ptr(), st)
    ev[i][1].record()
torch.cuda.synchronize()
ms = float(np.mean([s.elapsed_time(e) for s, e in ev]))
out = {"metric": "flat L2 queries/sec", "value": a.batch / (ms / 1e3), "ms_per_step": ms,
       "config": f"{a.n}x{a.dim} f32 flat L2, batch {a.batch}, k {a.k}",
       "gemm_tflops_equiv": 2.0 * a.batch * a.n * a.dim / (ms / 1e3) / 1e12}
if a.check:
    import oracle
    gi = ids.cpu().numpy().view(np.uint64); gd = dist.cpu().numpy()
    last = (a.steps - 1) % 4
    oi, od, oc = oracle.flat_search(x, q[last, :a.check].numpy(), k=a.k, nthreads=os.cpu_count())
    out["oracle_check"] = bool(np.array_equal(gi[:a.check], oi) and np.array_equal(gd[:a.check].view(np.uint32), od.view(np.uint32)))
print(json.dumps(out))
