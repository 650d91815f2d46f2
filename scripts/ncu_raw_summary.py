#!/usr/bin/env python
"""Pick the judged metrics out of `ncu -i X.ncu-rep --page raw --csv` dumps and print a markdown table.
usage: ncu_raw_summary.py label=file.raw.csv [label=file.raw.csv ...]"""
import csv
import sys

WANT = [
    ("kernel", "Kernel Name"),
    ("duration", "gpu__time_duration.sum"),
    ("SM clock", "sm__cycles_elapsed.avg.per_second"),
    ("grid x block", "launch__grid_size"),
    ("registers/thread", "launch__registers_per_thread"),
    ("dynamic smem / block", "launch__shared_mem_per_block_dynamic"),
    ("SM active cycles (avg)", "sm__cycles_active.avg"),
    ("elapsed cycles (gpc max)", "gpc__cycles_elapsed.max"),
    ("tensor pipe hmma cycles active, realtime (avg/SM)", "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg"),
    ("tensor memory (TMEM/UMMA operand) pipe active % of elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("DRAM read", "dram__bytes_read.sum"),
    ("DRAM write", "dram__bytes_write.sum"),
    ("DRAM bytes/s", "dram__bytes.sum.per_second"),
    ("L2 throughput %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 hit rate %", "lts__t_sector_hit_rate.pct"),
    ("L1/TEX (smem) throughput %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps active % of peak", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue slots busy %", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
]


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, v, u in zip(hdr, vals, units)}


def main():
    items = [a.split("=", 1) for a in sys.argv[1:]]
    data = [(lab, load(p)) for lab, p in items]
    print("| metric | " + " | ".join(l for l, _ in data) + " |")
    print("|---|" + "---|" * len(data))
    for name, key in WANT:
        cells = []
        for _, d in data:
            cand = [k for k in d if k == key] or [k for k in d if k.startswith(key)]
            if not cand:
                cells.append("n/a")
                continue
            v, u = d[cand[0]]
            if key == "Kernel Name":
                v = v.replace("void nnb::<unnamed>::", "").split("(nnb")[0].split("(const")[0]
                v = "`" + v[:70] + "`"
            if key == "launch__grid_size":
                v = v + " x " + d.get("launch__block_size", ("?", ""))[0]
            cells.append((v + " " + u).strip())
        print(f"| {name} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
