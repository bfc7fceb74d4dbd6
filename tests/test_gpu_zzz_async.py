"""The async surface (lancedb_b200/aio.py) on the GPU: coroutines run the synchronous builder in worker threads, so
`asyncio.gather` over many single-vector queries exercises the library's re-entrancy (per-call stream + workspace)
the way the reference's tokio workers would (python/src/runtime.rs:113-151).  Results must equal the synchronous
builder's, which the parity suite ties to the oracle.  (Sorted last on purpose: written without a GPU at hand.)"""
import asyncio

import numpy as np
import pytest

import lancedb_b200 as lancedb
from lancedb_b200 import aio

pytestmark = pytest.mark.gpu


def test_async_queries_equal_sync_queries_and_overlap():
    rng = np.random.default_rng(91)
    x = rng.standard_normal((8000, 32)).astype(np.float32)
    db = lancedb.connect("memory://")
    t = db.create_table("v", {"vector": x, "id": np.arange(8000), "b": np.arange(8000) % 5})
    t.create_index(metric="l2", num_partitions=16, num_sub_vectors=4, max_iterations=4, accelerator="cuda")
    at = aio.AsyncTable(t)
    q = rng.standard_normal((24, 32)).astype(np.float32)

    async def main():
        outs = await asyncio.gather(*[at.vector_search(v).nprobes(6).limit(7).with_row_id().to_arrow() for v in q])
        for v, o in zip(q, outs):
            assert o.equals(t.search(v).nprobes(6).limit(7).with_row_id(True).to_arrow())
        multi = await at.query().nearest_to([q[0], q[1]]).nprobes(6).refine_factor(2).limit(5).to_arrow()
        assert multi.equals(t.search(q[:2]).nprobes(6).refine_factor(2).limit(5).to_arrow())
        flt = await at.query().where("b = 1").nearest_to(q[2]).nprobes(16).limit(6).to_list()
        assert [r["id"] for r in flt] == t.search(q[2]).where("b = 1").nprobes(16).limit(6).to_arrow()["id"].to_pylist()
        flat = await at.vector_search(q[3]).bypass_vector_index().limit(4).to_list()
        assert [r["id"] for r in flat] == t.search(q[3]).bypass_vector_index().limit(4).to_arrow()["id"].to_pylist()
    asyncio.run(main())
