#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on its config 2 (the 1xB200 IVF_PQ case):
ANN queries/sec, 1M x 768 f32, IVF_PQ nlist=1024, PQ m=96x8bit, nprobes=20, k=10, batch=1024.

A "step" = one pass of the hot path over one batch of 1024 synthetic queries.
  value     : whole-job QPS with queries/results resident in HBM (lgpu_search_device),
              timed per step with CUDA events on the launching stream, L2 flushed
              (untimed) between steps;
  e2e       : the same metric through the host-buffer C-ABI call (lgpu_search) with pinned
              host buffers, H2D of the queries and D2H of the results inside the timed region;
  roofline  : algorithmic PQ-code bytes of the batch / the scan kernel's measured duration
              (CUDA events recorded around the kernel by the library) vs the measured HBM peak;
  cpu_baseline : the CPU oracle (a port of the reference's lance path) on the host cores.
N > 1 (torchrun): independent replicas, one batch per rank per step, no data-path
collective ("scaling": "weak"); `--parallelism sharded` times the partition-sharded path
with one NCCL all-gather of per-rank top-k + merge instead.
`--impl reference` times the CPU oracle alone (the reference's Rust path cannot be built
here: no cargo, lance un-vendored), rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]
    "c2": dict(n=1_000_000, dim=768, nlist=1024, m=96, nprobes=20, k=10, batch=1024, metric="l2"),
    # small variant for local CPU checks of the harness itself
    "tiny": dict(n=20_000, dim=64, nlist=32, m=8, nprobes=4, k=10, batch=64, metric="l2"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


LATENT_RANK = 32
DATA_DESC = ("synthetic float32: rank-32 Gaussian latent z@A (A fixed, seed 44) + N(0, 0.05^2) noise; "
             "base seed 42, queries seed 43")


def synth_vectors(cfg, n, seed, device):
    """Synthetic float32[dim] vectors with a low intrinsic dimension, like real embeddings:
    x = z A + 0.05 eps, z ~ N(0, I_32), A a fixed 32 x dim matrix.  Pure i.i.d. N(0,1) in 768-d
    (SURVEY.md 8d's first variant) has no neighbourhood structure at all: k-means on it
    degenerates (partition sizes std/mean 2.3, the probed partitions hold 5x the nominal
    nprobes*N/nlist rows, recall@10 ~ 0.04), so the workload would no longer be BASELINE.md's
    1.875 MB of codes per query.  With this generator partitions are balanced (std/mean ~0.2)
    and recall is meaningful."""
    import torch
    ga = torch.Generator(device="cpu").manual_seed(44)
    A = (torch.randn(LATENT_RANK, cfg["dim"], generator=ga) / LATENT_RANK ** 0.5).to(device)
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = torch.empty(n, cfg["dim"], dtype=torch.float32, device=device)
    chunk = 1 << 17
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        z = torch.randn(e - s, LATENT_RANK, generator=g).to(device)
        eps = torch.randn(e - s, cfg["dim"], generator=g).to(device)
        out[s:e] = z @ A + 0.05 * eps
    return out


def index_cache_path(cfg, tag):
    key = "_".join(f"{k}{cfg[k]}" for k in ("n", "dim", "nlist", "m", "metric"))
    return f"/tmp/lancedb_b200_bench_v2_{tag}_{key}.npz"


def get_index(cfg, tag, device):
    """Train (torch ops; setup, untimed) or load the synthetic index; also returns exact
    top-k ground truth for a few queries."""
    from lancedb_b200.index import IvfPqIndexData, train_ivf_pq
    import torch
    path = index_cache_path(cfg, tag)
    if os.path.exists(path):
        z = np.load(path)
        ix = IvfPqIndexData(int(z["dim"]), int(z["nlist"]), int(z["m"]), str(z["metric"]), z["centroids"],
                            z["codebook"], z["part_offsets"], z["codes_t"], z["row_ids"], None)
        return ix, z["gt_queries"], z["gt_ids"]
    t0 = time.time()
    x = synth_vectors(cfg, cfg["n"], 42, device)
    ix = train_ivf_pq(x, num_partitions=cfg["nlist"], num_sub_vectors=cfg["m"], distance_type=cfg["metric"],
                      max_iterations=12, sample_rate=64, device=device)
    gq = synth_vectors(cfg, 128, 4343, device)
    xs = x / x.norm(dim=1, keepdim=True) if cfg["metric"] == "cosine" else x
    qs = gq / gq.norm(dim=1, keepdim=True) if cfg["metric"] == "cosine" else gq
    d = (xs * xs).sum(1)[None, :] - 2.0 * qs @ xs.T
    gt = d.topk(cfg["k"], largest=False).indices.cpu().numpy().astype(np.uint64)
    gqn = gq.cpu().numpy()
    del x, xs, d
    tmp = path + f".{os.getpid()}.tmp.npz"
    np.savez(tmp, dim=ix.dim, nlist=ix.nlist, m=ix.m, metric=ix.metric, centroids=ix.centroids,
             codebook=ix.codebook, part_offsets=ix.part_offsets, codes_t=ix.codes_t, row_ids=ix.row_ids,
             gt_queries=gqn, gt_ids=gt)
    os.replace(tmp, path)
    log(f"[bench] index built in {time.time() - t0:.1f}s -> {path}")
    return ix, gqn, gt


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic():
    """dram__bytes_read+write per launch of the scan kernel from the committed ncu --set full capture
    (profiles/): a profiler number, so it is read from the profile, never measured in this run."""
    try:
        cands = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_scan_traffic.json"))
        with open(os.path.join(ROOT, "profiles", cands[-1])) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except Exception:
        return None


def cpu_baseline(cfg, ix, queries, seconds=12.0):
    """The oracle (port of the lance CPU path) on all host cores, bounded sample."""
    import oracle
    cores = os.cpu_count() or 1
    orc = oracle.OracleIndex.from_data(ix)
    probe = queries[:max(cores, 8)]
    t0 = time.perf_counter()
    orc.search(probe, k=cfg["k"], nprobes=cfg["nprobes"], nthreads=cores)
    per_q = (time.perf_counter() - t0) / len(probe)
    n = int(min(len(queries), max(cores * 4, seconds / max(per_q, 1e-6))))
    t0 = time.perf_counter()
    orc.search(queries[:n], k=cfg["k"], nprobes=cfg["nprobes"], nthreads=cores)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"{n} queries of the same workload, {dt:.1f}s, oracle/oracle.c with {cores} threads"}


def run_reference(args, cfg):
    """--impl reference: the CPU oracle alone, rank 0 only."""
    rank, local, world = dist_env()
    if rank != 0:
        return
    import oracle
    import torch
    device = f"cuda:{local}" if torch.cuda.is_available() else "cpu"
    ix, gq, gt = get_index(cfg, args.workload, device)
    B = cfg["batch"]
    cores = os.cpu_count() or 1
    orc = oracle.OracleIndex.from_data(ix)
    nb = 4
    q = synth_vectors(cfg, B * nb, 43, "cpu").numpy().reshape(nb, B, cfg["dim"])
    for i in range(args.warmup):
        orc.search(q[i % nb][: max(cores, B // 8)], k=cfg["k"], nprobes=cfg["nprobes"], nthreads=cores)
    t0 = time.perf_counter()
    for i in range(args.steps):
        orc.search(q[i % nb], k=cfg["k"], nprobes=cfg["nprobes"], nthreads=cores)
    dt = time.perf_counter() - t0
    qps = B * args.steps / dt
    sample = f"{B} queries per step (the full batch), oracle/oracle.c, {cores} threads"
    _emit({
        "impl": "reference", "metric": "ANN queries/sec (IVF_PQ)", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": DATA_DESC,
        "config": workload_config(cfg, args, 1, "cpu"),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def workload_config(cfg, args, world, par):
    return {"workload": f"{cfg['n']}x{cfg['dim']} f32, IVF_PQ nlist={cfg['nlist']} m={cfg['m']}x8bit, "
                        f"nprobes={cfg['nprobes']}, k={cfg['k']}, batch={cfg['batch']}, {cfg['metric']}",
            "baseline_config": "BASELINE.json configs[1]" if args.workload == "c2" else args.workload,
            "batch_per_gpu": cfg["batch"], "global_batch": cfg["batch"] * (world if par == "replicas" else 1),
            "parallelism": par if world > 1 else "single",
            "l2_flush": "512 MiB write between steps (untimed); each step uses a different query batch"}


_REAL_STDOUT = None


def _quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries underneath (NCCL prints its version banner to
    fd 1 when NCCL_DEBUG is set in the environment) write there too, so fd 1 points at stderr until
    _emit() prints the line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(obj):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(obj), flush=True)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--parallelism", default="replicas", choices=["replicas", "sharded"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = WORKLOADS[args.workload]
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        return run_reference(args, cfg)

    import torch
    from lancedb_b200 import _native
    rank, local, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (lancedb_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(device))
    par = args.parallelism if world > 1 else "replicas"

    # ---- setup (untimed): index, queries ----
    if rank == 0:
        ix, gq, gt = get_index(cfg, args.workload, device)
    if world > 1:
        dist.barrier()
    if rank != 0:
        ix, gq, gt = get_index(cfg, args.workload, device)
    torch.cuda.empty_cache()
    full_ix = ix
    if par == "sharded":
        ix = ix.shard(rank, world)
    gpu = _native.GpuIvfPq(ix, device=local, with_vectors=False)
    B, k, dim = cfg["batch"], cfg["k"], cfg["dim"]
    nb = 8
    qseed = 43 if par == "sharded" else 43 + 1000 * rank
    q_host = torch.empty(nb, B, dim, dtype=torch.float32).pin_memory()
    q_host.copy_(synth_vectors(cfg, nb * B, qseed, "cpu").reshape(nb, B, dim))
    d_q = q_host.to(device)
    d_ids = torch.empty(B, k, dtype=torch.int64, device=device)     # u64 payload
    d_dist = torch.empty(B, k, dtype=torch.float32, device=device)
    d_cnt = torch.empty(B, dtype=torch.int32, device=device)
    h_ids = torch.empty(B, k, dtype=torch.int64).pin_memory()
    h_dist = torch.empty(B, k, dtype=torch.float32).pin_memory()
    h_cnt = torch.empty(B, dtype=torch.int32).pin_memory()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=device)
    p = _native.make_params(k=k, nprobes=cfg["nprobes"])
    stream = torch.cuda.current_stream().cuda_stream
    if par == "sharded":
        g_ids = torch.empty(world, B, k, dtype=torch.int64, device=device)
        g_dist = torch.empty(world, B, k, dtype=torch.float32, device=device)
        m_ids = torch.empty_like(d_ids); m_dist = torch.empty_like(d_dist); m_cnt = torch.empty_like(d_cnt)

    def step_device(i):
        gpu.search_device(d_q[i % nb].data_ptr(), B, p, d_ids.data_ptr(), d_dist.data_ptr(), d_cnt.data_ptr(), stream)
        if par == "sharded":
            dist.all_gather_into_tensor(g_ids.view(-1, k), d_ids)
            dist.all_gather_into_tensor(g_dist.view(-1, k), d_dist)
            _native.merge_topk_device(local, world, B, k, g_ids.data_ptr(), g_dist.data_ptr(), m_ids.data_ptr(),
                                      m_dist.data_ptr(), m_cnt.data_ptr(), stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness gate before timing: recall vs exact, and the oracle on a few queries ----
    recall = None
    if par == "replicas":
        gi, gd, gc = gpu.search(gq, k=k, nprobes=cfg["nprobes"])
        recall = float(np.mean([len(set(gi[i].tolist()) & set(gt[i].tolist())) / k for i in range(len(gq))]))

    sampler = ClockSampler(local)      # samples clocks / throttle reasons through regions (1) and (2)
    sampler.start()
    t_wait = time.time()
    while not sampler.rows and time.time() - t_wait < 3.0:      # rank-local work only: no collectives here
        gpu.search_device(d_q[0].data_ptr(), B, p, d_ids.data_ptr(), d_dist.data_ptr(), d_cnt.data_ptr(), stream)
        torch.cuda.synchronize()
    for i in range(args.warmup):
        step_device(i)
    barrier()
    # ---- (1) device-resident timed region ----
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record()
        step_device(i)
        ev[i][1].record()
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    units = B * args.steps * (world if par == "replicas" else 1)
    value = units / (total_ms / 1e3)

    # ---- (2) end to end through the host-buffer C-ABI call (pinned host memory) ----
    qn = q_host.numpy(); hi = h_ids.numpy().view(np.uint64); hd = h_dist.numpy(); hc = h_cnt.numpy().view(np.uint32)
    e2e_val = None
    if par == "replicas":
        for i in range(2):
            gpu.search_into(qn[i % nb], p, hi, hd, hc)
        barrier()
        e2e_s = 0.0
        for i in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            gpu.search_into(qn[i % nb], p, hi, hd, hc)      # H2D + kernels + D2H + sync inside
            e2e_s += time.perf_counter() - t0
        e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e_val = units / float(e2e_t.item())
    clocks = sampler.stop()

    # ---- (3) per-kernel times (library CUDA events on the launching stream) -> roofline ----
    _native.set_profiling(True)
    stage = {}
    scan_ms, code_bytes = [], 0
    for i in range(args.steps):
        flush.zero_()
        gpu.search_device(d_q[i % nb].data_ptr(), B, p, d_ids.data_ptr(), d_dist.data_ptr(), d_cnt.data_ptr(), stream)
        s = _native.last_stage_ms()
        for kk, v in s.items():
            stage[kk] = stage.get(kk, 0.0) + v / args.steps
        scan_ms.append(s["scan"])
        code_bytes = _native.last_scanned_code_bytes()
    _native.set_profiling(False)
    peak, peak_src = measured_peak_gbs()
    scan_avg = float(np.mean(scan_ms))
    achieved = code_bytes / (scan_avg / 1e3) / 1e9

    if rank == 0:
        out = {
            "metric": "ANN queries/sec (IVF_PQ)", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak" if par == "replicas" else "strong", "vs_baseline": None,
            "dtype": "f32", "data": DATA_DESC,
            "config": workload_config(cfg, args, world, par),
            "clocks": clocks,
            # to_bf16, gemm (sample), select, threshold, gemm (filtered), overflow flags, pair distance, select,
            # dist_matrix + select fix-ups, 3 group kernels, scan, select(top-k)
            "gpu_launches": args.steps * 16,
            "recall_at_k": recall,
            "stage_ms": stage,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": profiled_traffic(), "kernel": "scan2_kernel<8,false> (fused PQ table build + code scan, streaming)",
                         "kernel_ms": scan_avg,
                         "algorithmic_bytes_per_launch": code_bytes, "peak_source": peak_src,
                         "compulsory_bytes_per_launch": int(full_ix.codes_t.size)},
        }
        if e2e_val is not None:
            out["e2e"] = {"value": e2e_val, "unit": "queries/s", "h2d_bytes_per_step": B * dim * 4,
                          "d2h_bytes_per_step": B * k * 12 + B * 4}
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(cfg, full_ix, qn.reshape(-1, dim))
        _emit(out)
    gpu.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
