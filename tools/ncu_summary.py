#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_decode_full.txt [kernel-key-for-traffic.json]"""
import csv
import json
import os
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    key = sys.argv[3] if len(sys.argv) > 3 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# summary of {os.path.basename(rep)} (ncu --set full --clock-control none; cold-cache, serialised replays)"]
    traffic = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines.append(f"\n== launch id {r[0]}: {name}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"  {w:90s} {r[i]} {units[i]}")
        i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        t = to_bytes(r[i_r], units[i_r]) + to_bytes(r[i_w], units[i_w])
        traffic.append(t)
        lines.append(f"  traffic = dram read + write = {t / 1e9:.4f} GB per launch")
    open(out, "w").write("\n".join(lines) + "\n")
    if key and traffic:
        p = os.path.join(os.path.dirname(out), "traffic.json")
        d = json.load(open(p)) if os.path.exists(p) else {}
        d[key] = round(sum(traffic) / len(traffic))
        json.dump(d, open(p, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
