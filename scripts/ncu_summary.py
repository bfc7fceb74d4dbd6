#!/usr/bin/env python
"""Summarise the ncu --set full captures that scripts/gpu_profile.sh leaves in gpurun_out/ into the tracked
profiles/ directory: per-launch headline metrics + warp-stall mix for the scan kernel and the tcgen05 GEMM,
and the scan kernel's DRAM traffic (bench.py reads it for roofline.traffic).  Usage:
    python scripts/ncu_summary.py r01"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def stall_mix(d):
    tot, parts = 0.0, []
    for k, (v, _) in d.items():
        if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued"):
            try:
                parts.append((float(v), k[len("smsp__pcsamp_warps_issue_stalled_"):]))
                tot += float(v)
            except ValueError:
                pass
    parts.sort(reverse=True)
    return ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in parts[:8]) if tot else "n/a"


def main():
    """python scripts/ncu_summary.py <tag> [label=report.ncu-rep ...]; the first report is the scan kernel (its DRAM
    traffic goes to profiles/<tag>_scan_traffic.json, which bench.py reads for roofline.traffic)."""
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    reps = [a.split("=", 1) for a in sys.argv[2:]] or [["scan", "gpurun_out/r02_scan3.ncu-rep"], ["finalize", "gpurun_out/r02_fin.ncu-rep"]]
    lines = ["ncu --set full --clock-control none --import-source on, 1 x B200 (under gpurun); per-launch values of the",
             "steady-state launch (bench.py C2 workload: 1M x 768, nlist 1024, m 96, nprobes 20, k 10, B 1024) unless the",
             "label says otherwise.  Sources: " + ", ".join(r for _, r in reps) + " (scratch, not tracked).", ""]
    first = True
    for label, rep in reps:
        rep = rep if os.path.isabs(rep) else os.path.join(ROOT, rep)
        if not os.path.exists(rep):
            continue
        d = raw_page(rep)
        kern = d.get("Kernel Name", ("?", ""))[0]
        lines.append(f"[{label}] kernel: {kern}")
        for m in METRICS + ["l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
                            "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]:
            if m in d:
                lines.append(f"  {m:75s} {d[m][0]} {d[m][1]}")
        lines.append(f"  stall reasons: {stall_mix(d)}")
        lines.append("")
        if first:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tr = sum(float(d[k][0]) * scale.get(d[k][1], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            with open(os.path.join(ROOT, "profiles", f"{tag}_scan_traffic.json"), "w") as f:
                json.dump({"kernel": kern, "dram_bytes_per_launch": tr,
                           "source": f"profiles/{tag}_ncu_summary.txt (ncu --set full, 1 launch)"}, f)
                f.write("\n")
            first = False
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.txt"), "w") as f:
        f.write("\n".join(lines))
    print("\n".join(lines))


if __name__ == "__main__":
    main()
