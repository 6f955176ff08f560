#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the counters DESIGN.md / bench.py quote, per profiled kernel:
duration, DRAM bytes and throughput, sectors per request (loads / stores), L1 / L2 hit rates, pipe utilisation,
occupancy, and the warp stall reasons (cycles a warp waits per issued instruction, by reason).
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [out.txt]"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warps_eligible.avg.per_cycle_active",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"]).decode()
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for d in data:
        name = d[hdr.index("Kernel Name")]
        out.append(f"== {name}")
        val = {}
        stalls = []
        for i, h in enumerate(hdr):
            if h in WANT:
                out.append(f"{h} [{units[i]}] = {d[i]}")
                val[h] = d[i]
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(d[i].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass

        def num(k):
            try:
                return float(val[k].replace(",", ""))
            except (KeyError, ValueError):
                return None
        for op in ("ld", "st"):
            s, r = num(f"l1tex__t_sectors_pipe_lsu_mem_global_op_{op}.sum"), num(f"l1tex__t_requests_pipe_lsu_mem_global_op_{op}.sum")
            if s is not None and r:
                out.append(f"derived: sectors per request, global {op} = {s / r:.2f}")
        out.append("warp stall reasons (warp-cycles stalled per issued instruction; 'selected' = issuing):")
        for v, n in sorted(stalls, reverse=True):
            if v >= 0.005:
                out.append(f"  {n:28s} {v:7.3f}")
    txt = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main()
