"""Partition-sharded search over NCCL (needs >= 2 GPUs; skipped otherwise): every rank must end
up with exactly the single-GPU / oracle result."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _rank_main(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import oracle
    from lancedb_b200.distributed import ShardedIvfPq
    from tests.util import queries, random_index
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}",
                            device_id=torch.device(f"cuda:{rank}"))
    rng = np.random.default_rng(3)
    ix = random_index(rng, dim=64, nlist=40, m=8, n=20000)
    q = queries(rng, 77, 64)
    sh = ShardedIvfPq(ix, device=rank)
    oi, od, oc = oracle.OracleIndex.from_data(ix).search(q, k=10, nprobes=9, nthreads=4)
    # host-buffer collective (lgpu_search_sharded): in-library ncclAllGather of 16-byte records + merge
    ids, dst, cnt = sh.search(q, k=10, nprobes=9)
    assert np.array_equal(cnt, oc) and np.array_equal(ids, oi)
    assert np.array_equal(dst.view(np.uint32), od.view(np.uint32))
    # device-buffer collective (lgpu_search_sharded_device) on torch's current stream
    d_ids, d_dst, d_cnt = sh.search_device(torch.from_numpy(q).cuda(rank), k=10, nprobes=9)
    torch.cuda.synchronize()
    assert np.array_equal(d_ids.cpu().numpy().view(np.uint64), oi) and np.array_equal(d_cnt.cpu().numpy().view(np.uint32), oc)
    assert np.array_equal(d_dst.cpu().numpy().view(np.uint32), od.view(np.uint32))
    sh.close()
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")


def test_sharded_search_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_rank_main, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
