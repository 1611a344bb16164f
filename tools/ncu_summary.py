"""Summarise an .ncu-rep: per launch, the metrics this project reads (duration, DRAM bytes, hit rates, issue
utilisation, active threads per instruction).  Usage: python tools/ncu_summary.py report.ncu-rep [--csv out.csv]"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    cols = [i for i, n in enumerate(h) if n in WANT]
    table = [[h[i] + (" [%s]" % units[i] if units[i] else "") for i in cols]] + [[r[i] for i in cols] for r in rows[2:]]
    if "--csv" in sys.argv:
        with open(sys.argv[sys.argv.index("--csv") + 1], "w", newline="") as f:
            csv.writer(f).writerows(table)
    for r in table[1:]:
        print("----")
        for n, v in zip(table[0], r):
            print("  %-75s %s" % (n, v))


if __name__ == "__main__":
    main()
