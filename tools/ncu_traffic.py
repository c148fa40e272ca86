"""profiles/traffic.json from `ncu --page raw --csv` dumps of tools/prof_band.py (one dump per band).
usage: python tools/ncu_traffic.py OUT.json WORKLOAD SOURCE_NOTE band0=raw_b0.csv band7=raw_b7.csv
Per kernel and band: launch time, measured DRAM bytes (read + write), FMA-pipe / issue / warps-active percentages,
shared-memory wavefronts and bank conflicts, executed warp instructions — the measured side of the roofs bench.py
reports (`roofline.traffic`, `roofline.binding_roofs`)."""
import csv
import json
import re
import sys

KEEP = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "smsp__inst_executed.sum": "warp_instructions",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid", "launch__block_size": "block",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3,
        "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}


def parse(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    units = rows[1]
    data = [r for r in rows[1:] if r and r[0].strip().isdigit()]
    kcol = hdr.index("Kernel Name")
    out = {}
    for r in data:
        name = re.sub(r"[<(].*", "", r[kcol]).replace("void ", "").strip()
        # logical pass names: the pair-engine kernels (rows2.cuh / cols2.cuh) keep the keys of the passes they replace;
        # k_cols2 serves both directions, in launch order forward then inverse
        if name == "k_cols2":
            name = "k_cols_fwd" if "k_cols_fwd" not in out else "k_cols_inv"
            kern = "k_cols2"
        else:
            kern = name
            name = {"k_rows2_fwd": "k_rows_fwd", "k_rows2_inv": "k_rows_inv"}.get(name, name)
        ent = {"kernel": kern}
        for m, key in KEEP.items():
            if m not in hdr:
                continue
            c = hdr.index(m)
            try:
                v = float(r[c].replace(",", ""))
            except ValueError:
                continue
            ent[key] = v * UNIT.get(units[c].strip(), 1.0)
        if "dram_read" in ent and "dram_write" in ent:
            ent["dram_bytes"] = ent["dram_read"] + ent["dram_write"]
        ent["time_ms"] = ent.pop("time", None)
        out.setdefault(name, ent)  # first captured launch of every kernel
    return out


def main():
    out_path, workload, note = sys.argv[1:4]
    try:
        doc = json.load(open(out_path))
    except Exception:
        doc = {}
    w = doc.setdefault(workload, {})
    w["source"] = note
    for arg in sys.argv[4:]:
        band, path = arg.split("=")
        w[band] = parse(path)
    # keys bench.py's `roofline.traffic` reads: DRAM bytes per launch of band 0
    if "band0" in w:
        for k, v in w["band0"].items():
            if "dram_bytes" in v:
                w[k] = int(v["dram_bytes"])
    json.dump(doc, open(out_path, "w"), indent=1, sort_keys=True)
    for band in sorted(k for k in w if k.startswith("band")):
        tot = sum(v.get("dram_bytes", 0) for v in w[band].values())
        ms = sum(v.get("time_ms", 0) or 0 for v in w[band].values())
        print(f"{workload} {band}: {tot / 1e9:.2f} GB DRAM traffic in {ms:.2f} ms of kernels ({tot / ms / 1e6:.0f} GB/s)")


if __name__ == "__main__":
    main()
