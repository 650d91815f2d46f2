#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table
(markdown on stdout).  usage: summarize_launches.py launches.csv [title]"""
import csv
import io
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    with open(path, errors="replace") as f:
        lines = f.readlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(io.StringIO("".join(lines[start:]))))
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        tot[name] += us
        cnt[name] += 1
    total = sum(tot.values())
    ours = sum(v for k, v in tot.items() if k.startswith("nnb"))
    print(f"# {title}\n")
    print(f"Kernels: {sum(cnt.values())}; sum of durations {total:.1f} us; nnb:: kernels {ours:.1f} us ({100*ours/total:.1f}%), "
          f"torch array back-end {total-ours:.1f} us ({100*(total-ours)/total:.1f}%)\n")
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:45]:
        print(f"| `{k[:90]}` | {cnt[k]} | {v:.1f} | {100*v/total:.1f}% |")


if __name__ == "__main__":
    main()
