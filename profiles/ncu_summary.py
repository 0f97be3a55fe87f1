#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers quoted in DESIGN.md / profiles/*.md.
usage: python profiles/ncu_summary.py gpurun_out/foo.ncu-rep"""
import csv
import subprocess
import sys


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "smsp__sass_average_branch_targets_threads_uniform.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_pipe_xu.sum",
            "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
            "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_cbu.sum",
            "sm__inst_executed_pipe_adu.sum", "sm__inst_executed_pipe_uniform.sum"]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:90])
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:75s} {r[i]:>16s} {units[i]}")
        st = []
        for i, h in enumerate(hdr):
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                try:
                    st.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1.0
        print("  stall samples:", ", ".join(f"{n}={v / tot:.3f}" for v, n in sorted(st, reverse=True)[:8]))


if __name__ == "__main__":
    main(sys.argv[1])
