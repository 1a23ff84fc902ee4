#!/usr/bin/env python3
"""Text summary of an `ncu --set full` report for profiles/ (run here, no GPU needed): ncu_summary.py rep.ncu-rep [header line]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "sm__cycles_elapsed.avg.per_second"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    if len(sys.argv) > 2:
        print(sys.argv[2])
    stalls = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
    for n, r in enumerate(rows[2:]):
        print(f"--- launch {n}: {r[hdr.index('Kernel Name')]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        for k in KEYS:
            if k in hdr:
                print(f"  {k} = {r[hdr.index(k)]} {units[hdr.index(k)]}")
        top = sorted(((float(r[hdr.index(h)] or 0), h.split("stalled_")[1].split("_per")[0]) for h in stalls), reverse=True)[:6]
        print("  stalls per issue: " + ", ".join(f"{n} {v:.2f}" for v, n in top))
        try:
            # ncu picks a unit per metric AND per report (Mbyte for one, Gbyte for the other): normalise to MB
            scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            rd, wr = float(r[ir]) * scale[units[ir]], float(r[iw]) * scale[units[iw]]
            print(f"  => DRAM traffic {rd + wr:.1f} MB per launch ({rd:.1f} read + {wr:.1f} written)")
        except (ValueError, IndexError, KeyError):
            pass


if __name__ == "__main__":
    main()
