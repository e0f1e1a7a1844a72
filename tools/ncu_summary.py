#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep, one kernel launch) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof_jac.ncu-rep profiles/r01_k_jacobian.txt

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU) and keeps the
counters DESIGN.md argues from: duration, FP64 pipe activity, issue rate, stall reasons,
occupancy limiters, instruction-cache hit rate, DRAM traffic, local-memory traffic.
"""
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_issued.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_average_branch_targets_threads_uniform.pct", "sm__icc_request_hit_rate.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append("kernel: %s   grid %s block %s" % (d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
        for k in KEEP:
            if k in d:
                lines.append("  %-70s %s %s" % (k, d[k], u[k]))
        stalls = sorted(((float(d[k]), k) for k in d if k.startswith(STALL) and k.endswith("_per_issue_active.ratio")
                         and d[k] not in ("", "n/a")), reverse=True)
        lines.append("  stall reasons (warps stalled per issue-active cycle), top 8:")
        for v, k in stalls[:8]:
            lines.append("    %-40s %.3f" % (k[len(STALL):-len("_per_issue_active.ratio")], v))
    # machine-readable DRAM traffic of the profiled launch, read by bench.py (roofline.traffic)
    import json, os, re
    tj = os.path.join(os.path.dirname(os.path.abspath(out)), "traffic.json")
    db = json.load(open(tj)) if os.path.exists(tj) else {}
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        name = re.match(r"\w+", d.get("Kernel Name", "?")).group(0)
        try:
            tot = sum(float(d[k]) * mult[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        except (KeyError, ValueError):
            continue
        db[name] = {"dram_bytes_per_launch": tot, "grid": d.get("Grid Size"), "source": os.path.basename(out)}
    json.dump(db, open(tj, "w"), indent=1, sort_keys=True)
    open(out, "w").write("source: %s (ncu --set full --clock-control none, one launch inside bench.py)\n" % rep
                         + "\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
