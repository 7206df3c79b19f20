#!/usr/bin/env python
"""Text summary of an `ncu --set full` report: one line per captured launch with the counters DESIGN.md cites.

    python tools/summarize_ncu_full.py gpurun_out/prof_conv.ncu-rep profiles/r1_ncu_full_conv_summary.txt
"""
import csv
import subprocess
import sys

WANT = [
    ("kernel", "Kernel Name"), ("grid", "Grid Size"), ("us", "gpu__time_duration.sum"),
    ("tensor_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("issue_per_cycle", "smsp__issue_active.avg.per_cycle_active"),
    ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("smem_lsu_wavefront_pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("regs", "launch__registers_per_thread"), ("dyn_smem_KB", "launch__shared_mem_per_block_dynamic"),
    ("stall_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall_short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(n, hdr.index(h)) for n, h in WANT if h in hdr]
    with open(out, "w") as f:
        f.write("# %s  (ncu --set full --clock-control none; per-launch, cold cache, serialised; units as ncu reports them)\n" % rep)
        f.write(" | ".join("%s[%s]" % (n, units[i]) if units[i] else n for n, i in idx) + "\n")
        for r in rows[2:]:
            f.write(" | ".join(r[i][:60] for _, i in idx) + "\n")
    print(open(out).read())


if __name__ == "__main__":
    main(*sys.argv[1:3])
