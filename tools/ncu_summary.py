#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU needed) into a small text file for profiles/.
usage: tools/ncu_summary.py gpurun_out/adds.ncu-rep profiles/adds_rNN_summary.txt [launches.csv]"""
import csv, io, subprocess, sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]

def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full --clock-control none summary of {rep}", ""]
    for r in rows[2:]:
        d = {h: (u, v) for h, u, v in zip(hdr, units, r)}
        lines.append(f"kernel: {d.get('Kernel Name', ('', '?'))[1]}  grid {d.get('Grid Size', ('', '?'))[1]} block {d.get('Block Size', ('', '?'))[1]}")
        for k in KEYS:
            if k in d:
                lines.append(f"  {k:75s} {d[k][1]:>18s} {d[k][0]}")
        stalls = sorted(((float(v[1]), k) for k, v in d.items()
                         if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")), reverse=True)
        lines.append("  warp stall reasons (warps per issue-active cycle):")
        for v, k in stalls[:8]:
            lines.append(f"    {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:.3f}")
        lines.append("")
    if len(sys.argv) > 3:
        agg = defaultdict(lambda: [0, 0.0])
        rows = [r for r in csv.reader(open(sys.argv[3])) if len(r) > 5]
        h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
        for r in rows[1:]:
            try:
                agg[r[ki][:70]][1] += float(r[vi].replace(",", "")); agg[r[ki][:70]][0] += 1
            except ValueError:
                pass
        tot = sum(v[1] for v in agg.values())
        lines.append(f"# launch list ({sys.argv[3]}): gpu__time_duration.sum per kernel (cold-cache, serialised: compare shares)")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            lines.append(f"  {k:70s} launches={v[0]:4d} total_ms={v[1] / 1e6:10.3f} share={v[1] / tot * 100:6.2f}%")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))

if __name__ == "__main__":
    main()
