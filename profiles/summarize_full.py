"""Condense an `ncu --set full` report into one CSV row per captured launch with the columns the roofline
discussion needs (duration, DRAM bytes, DRAM / L2 / L1 / SM throughput %, occupancy, registers, tensor pipe).

    python profiles/summarize_full.py gpurun_out/<tag>_full.ncu-rep > profiles/<tag>_ncu_full_summary.csv
"""
import csv
import io
import subprocess
import sys

COLS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    idx = [head.index(c) for c in COLS if c in head]
    w = csv.writer(sys.stdout)
    w.writerow([head[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in body:
        w.writerow([r[i] for i in idx])


if __name__ == "__main__":
    main(sys.argv[1])
