"""Summarise one kernel of an .ncu-rep (ncu --set full) as `metric = value` lines: the numbers profiles/*.txt quote.
usage: python tools/ncu_summary.py report.ncu-rep "header line" ... > profiles/rN_<kernel>_ncu.txt"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def main():
    rep = sys.argv[1]
    for h in sys.argv[2:]:
        print("# " + h)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    for h, u, v in sorted(zip(hdr, units, vals)):
        if h in WANT or ("pcsamp_warps_issue_stalled" in h and "not_issued" not in h):
            print(f"{h} [{u}] = {v}")


if __name__ == "__main__":
    main()
