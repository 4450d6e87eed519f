#!/usr/bin/env python
"""Condense an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of per-kernel numbers
DESIGN.md and bench.py quote.  Usage: python tools/ncu_summary.py gpurun_out/X.ncu-rep > profiles/X.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name"))
        for k in KEYS:
            if k in d:
                print("  %-70s %s %s" % (k, d[k], u.get(k, "")))
        stalls = {k: float(v) for k, v in d.items()
                  if "issue_stalled" in k and k.endswith("_per_issue_active.ratio") and v}
        for k, v in sorted(stalls.items(), key=lambda x: -x[1])[:6]:
            print("  stall %-64s %.3f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace(
                "_per_issue_active.ratio", ""), v))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
