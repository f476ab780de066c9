#!/usr/bin/env python
"""Summary of one kernel launch of an ncu --set full report as JSON (the files under profiles/ncu_*.json).
usage: python tools/ncu_summary.py report.ncu-rep "command line the capture was taken with" > profiles/ncu_<name>.json"""
import csv, json, subprocess, sys

KEYS = {
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "gpu__time_duration.sum": "duration",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefronts_pct",
    "launch__block_size": "block", "launch__grid_size": "grid", "launch__cluster_size": "cluster_size",
    "launch__occupancy_limit_registers": "occ_limit_reg_blocks", "launch__occupancy_limit_shared_mem": "occ_limit_smem_blocks",
    "launch__registers_per_thread": "registers_per_thread", "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active": "dmma_subpipe_pct",
    "sm__warps_active.avg.per_cycle_active": "warps_active_per_cycle",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_sb",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_sb",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__warps_eligible.avg.per_cycle_active": "eligible_warps_per_scheduler",
}

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
out = []
for vals in rows[2:]:
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            d[KEYS[h]] = (v + " " + u).strip()
        if h == "Kernel Name":
            d["kernel"] = v
    if len(sys.argv) > 2:
        d["command"] = sys.argv[2]
    out.append(d)
print(json.dumps(out[0] if len(out) == 1 else out, indent=1))
