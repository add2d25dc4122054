"""Summarise an `ncu --page raw --csv` dump: one block of key metrics per kernel launch."""
import csv
import sys

WANT = [
    "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
    stalls = [(h, i) for i, h in enumerate(hdr)
              if h.startswith(STALL) and h.endswith("_per_issue_active.ratio")]
    for r in rows[2:]:
        print("-" * 100)
        for w, i in cols:
            print("%-60s %s %s" % (w, r[i], units[i]))
        top = sorted(((float(r[i].replace(",", "") or 0), h) for h, i in stalls),
                     reverse=True)[:6]
        for v, h in top:
            print("  stall %-50s %.2f warps/issue" %
                  (h[len(STALL):-len("_per_issue_active.ratio")], v))


if __name__ == "__main__":
    main(sys.argv[1])
