#!/usr/bin/env python
"""Markdown block for DESIGN.md section 7 from a bench.py JSON line (and optionally the N=2 line):
    python scripts/design_numbers.py profiles/r02_bench_c2.json [profiles/r02_bench_c2_n2.json]"""
import json, sys

j = json.load(open(sys.argv[1]))
st, rf, e2e, cpu = j["stage_ms"], j["roofline"], j["e2e"], j["cpu_baseline"]
print(f"1 x B200, `python bench.py` (N = 1, {j['steps']} steps, SM clock {j['clocks']['sm_mhz']:.0f} MHz, no throttle reasons: "
      f"{j['clocks']['reasons']}), config 2 (1M x 768, nlist 1024, m 96, nprobes 20, k 10, batch 1024):\n")
print("| quantity | value |\n|---|---|")
print(f"| `value` (device-resident) | {j['value'] / 1e6:.3f} M QPS, {j['ms_per_step']:.3f} ms per batch |")
print(f"| `e2e` (`lgpu_search`, host buffers, H2D + D2H inside) | {e2e['value'] / 1e6:.3f} M QPS; two async calls in flight {e2e['pipelined']['value'] / 1e6:.3f} M QPS |")
print(f"| stage ms (profiled batch) | coarse+probes {st['coarse'] + st['select_probes']:.3f}, regroup {st['group']:.3f}, scan {st['scan']:.3f}, finalize {st['topk']:.3f} |")
print(f"| `roofline` (scan kernel, algorithmic code bytes) | {rf['achieved']:.0f} of {rf['peak']:.0f} GB/s = **{rf['frac']:.3f}**; whole step {rf['whole_step_frac']:.3f}; DRAM traffic of the launch {rf['traffic'] / 1e6:.0f} MB vs {rf['algorithmic_bytes_per_launch'] / 1e6:.0f} MB algorithmic |")
print(f"| filter statistics | {j['filter_stats']['candidates'] / j['filter_stats']['queries']:.0f} candidates and {j['filter_stats']['rescored'] / j['filter_stats']['queries']:.1f} exact re-scores per query, {j['filter_stats']['flagged_queries']} queries through the exact fix-up |")
print(f"| gate | GPU == oracle (plain and refine x10) on {j['gate']['queries']} queries: {j['gate']['gpu_equals_oracle_plain']} / {j['gate']['gpu_equals_oracle_refine10']}; recall@10 {j['gate']['recall_at_k_gpu']:.3f} (refine x10: {j['gate']['recall_at_k_refine10_gpu']:.3f}), identical for the CPU arm |")
print(f"| `cpu_baseline` | {cpu['value']:.0f} QPS, oracle port on {cpu['cores']} threads ({cpu['cpu_model']}, cgroup quota {cpu['cgroup_quota_cpus']}) |")
lat = j.get("latency", {})
if lat:
    print(f"| latency B = 1 (host call) | IVF_PQ p50 {lat['ivf_pq_b1']['p50_us']:.0f} us / p99 {lat['ivf_pq_b1']['p99_us']:.0f} us; flat C1 (100k x 128) p50 {lat['c1_flat_b1']['p50_us']:.0f} us (CPU port {lat['c1_flat_b1']['cpu_port_p50_us']:.0f} us) |")
for w in j.get("extra_workloads", []):
    r = w["roofline"]
    extra = f", stage ms scan {w['stage_ms']['scan']:.2f} / finalize {w['stage_ms']['topk']:.2f} / coarse {w['stage_ms']['select_probes']:.2f}" if w.get("stage_ms") else ""
    print(f"| {w['name']} ({w['config']}) | {w['qps'] / 1e3:.0f} k QPS, {w['ms_per_batch']:.2f} ms per batch, roofline frac {r['frac']:.3f} ({r['bound']}){extra}; oracle check {w.get('oracle_check', w.get('gpu_equals_oracle_plain'))} |")
if len(sys.argv) > 2:
    n = json.load(open(sys.argv[2]))
    s, c = n["sharded"], n["c5"]
    print(f"\n{n['n_gpus']} x B200 (`torchrun --nproc-per-node {n['n_gpus']} bench.py --gpus {n['n_gpus']}`): replicas {n['value'] / 1e6:.3f} M QPS "
          f"({n['ms_per_step']:.3f} ms per batch per rank).  Partition-sharded config 2 on one shared batch: {s['qps'] / 1e6:.3f} M QPS, "
          f"{s['ms']:.3f} ms = local search {s['local_search_ms']:.3f} (coarse {s['coarse_ms']:.3f}, scan {s['scan_ms']:.3f}) + one "
          f"`ncclAllGather` of {s['allgather_bytes_per_rank']} B per rank {s['allgather_ms']:.3f} + merge {s['merge_ms']:.3f}; "
          f"parity with the single-GPU result on every rank: {s['parity']}, oracle check: {s['oracle_check']}.  Config-5-shaped shards "
          f"({c['rows_per_gpu']} rows per GPU, nlist 16384, batch 8192): {c['qps'] / 1e6:.3f} M QPS, {c['ms']:.2f} ms (coarse {c['coarse_ms']:.2f}, "
          f"scan {c['scan_ms']:.2f}, finalize {c['topk_ms']:.2f}, all-gather {c['allgather_ms']:.3f}, merge {c['merge_ms']:.3f}); oracle check {c['oracle_check']}.")
