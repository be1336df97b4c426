#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here with `ncu -i ... --page raw --csv`) into a small
markdown + JSON pair under profiles/.   python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_sense_n1024"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__warps_active.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]

UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kernels = []
    for vals in rows[2:]:
        d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
        kernels.append(d)
    md = ["# ncu --set full summary: %s" % rep, "", note, ""]
    js = []
    for d in kernels:
        name = d.get("Kernel Name", ("?", ""))[0]
        md += ["## %s" % name, "", "| metric | value | unit |", "|---|---|---|"]
        rec = {"kernel": name}
        for k in KEYS:
            if k in d:
                md.append("| %s | %s | %s |" % (k, d[k][0], d[k][1]))
                try:
                    rec[k] = float(d[k][0].replace(",", ""))
                    rec[k + ".unit"] = d[k][1]
                except ValueError:
                    rec[k] = d[k][0]
        rd = rec.get("dram__bytes_read.sum", 0) * UNIT_SCALE.get(rec.get("dram__bytes_read.sum.unit", "byte"), 1)
        wr = rec.get("dram__bytes_write.sum", 0) * UNIT_SCALE.get(rec.get("dram__bytes_write.sum.unit", "byte"), 1)
        rec["dram_bytes_per_launch"] = rd + wr
        md += ["", "dram traffic per launch = read + write = %.6g bytes" % (rd + wr), ""]
        js.append(rec)
    open(out + ".md", "w").write("\n".join(md) + "\n")
    json.dump(js, open(out + ".json", "w"), indent=1)
    print("\n".join(md))


if __name__ == "__main__":
    main()
