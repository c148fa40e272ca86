"""Markdown table of the metrics DESIGN.md / profiles/*.md quote from an `ncu --page raw --csv` dump.
usage: ncu -i prof.ncu-rep --page raw --csv > raw.csv; python tools/ncu_summary.py raw.csv"""
import csv
import re
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_sector_hit_rate.pct",
]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio$")


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    units = rows[1] if len(rows) > 1 and not rows[1][0].strip().isdigit() else None
    data = [r for r in rows[1:] if r and r[0].strip().isdigit()]
    kcol = hdr.index("Kernel Name")
    names = [re.sub(r"\(.*", "", r[kcol]).replace("void ", "") for r in data]
    print("| metric | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    for m in METRICS:
        if m not in hdr:
            continue
        c = hdr.index(m)
        u = f" [{units[c]}]" if units and units[c] else ""
        print(f"| {m}{u} | " + " | ".join(r[c] for r in data) + " |")
    print()
    for n, r in zip(names, data):
        st = []
        for c, h in enumerate(hdr):
            mm = STALL.search(h)
            if mm and r[c]:
                try:
                    st.append((float(r[c].replace(",", "")), mm.group(1)))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print(f"* `{n}`: " + ", ".join(f"{k} {v:.2f}" for v, k in st[:5]))


if __name__ == "__main__":
    main(sys.argv[1])
