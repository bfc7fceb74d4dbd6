"""Repro: B=1 cosine search with refine_factor through the table mirror (tests/test_gpu_api.py)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lancedb_b200 as lancedb

metric = sys.argv[1] if len(sys.argv) > 1 else "cosine"
rng = np.random.default_rng(0)
x = rng.standard_normal((6000, 64)).astype(np.float32)
db = lancedb.connect("memory://")
t = db.create_table("v", {"vector": x, "id": np.arange(6000)})
t.create_index(metric=metric, num_partitions=16, num_sub_vectors=8, max_iterations=4, accelerator="cuda")
q = rng.standard_normal((5, 64)).astype(np.float32)
for i in range(5):
    out = t.search(q[i]).distance_type(metric).nprobes(4).limit(10).offset(2).with_row_id(True).to_arrow()
print("plain ok", flush=True)
out = t.search(q[0]).distance_type(metric).nprobes(4).refine_factor(3).limit(5).to_arrow()
print("refine ok", out["_distance"].to_pylist(), flush=True)
