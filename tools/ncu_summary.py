#!/usr/bin/env python
"""tools/ncu_summary.py — turns an .ncu-rep brought back in gpurun_out/ into the small, tracked summaries under
profiles/: one CSV row per profiled launch with the metrics the roofline discussion uses, and (optionally) a JSON
with the per-launch DRAM traffic of the two hot kernels that bench.py reads for `roofline.traffic`.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_full.csv [profiles/traffic_r01.json]
  python tools/ncu_summary.py --launches gpurun_out/launches.csv profiles/r01_launches.csv
"""
import csv
import io
import json
import subprocess
import sys

KEEP = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_tensor_subpipe_dmma.sum",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value) * scale[unit]


def full(rep, out_csv, traffic_json=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    if traffic_json:
        name, rd, wr, dur = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        acc = {}
        for r in rows[2:]:
            for key in ("reduce", "apply", "combine"):
                if key + "_kernel" in r[name]:
                    acc.setdefault(key, []).append(to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr]))
        out = {"source": rep, "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none"}
        for key, v in acc.items():
            out[key + "_kernel_dram_bytes_per_launch"] = sum(v) / len(v)
        json.dump(out, open(traffic_json, "w"), indent=1)


def launches(src, dst):
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    k, v, g = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["launch", "kernel", "grid", "gpu__time_duration.sum [ns]"])
        for i, r in enumerate(rows[1:]):
            w.writerow([i, r[k], r[g], r[v]])


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
