"""Summarises an ncu report (ncu -i X.ncu-rep --page raw --csv) into the handful of metrics DESIGN.md cites.
usage: python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/rNN_name.txt"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for k, row in enumerate(rows[2:]):
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"== launch {k}: {name[:90]}")
        for i, h in enumerate(hdr):
            if h in WANT:
                print(f"  {h:70s} {row[i]:>18s} {units[i]}")
            elif "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    if float(row[i]) >= 0.1:
                        print(f"  stall {h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''):64s} {row[i]:>18s} per issued instr")
                except ValueError:
                    pass


if __name__ == "__main__":
    main(sys.argv[1])
