#!/usr/bin/env python
"""ncu CSV of scripts/gemm_classes_once.py (+ its stdout) -> profiles/rN_gemm_classes_dram.{json,md}: DRAM bytes per launch of
every nn.Linear GEMM class of the GPT-small step, and the per-step weighted mean that bench.py reports as `roofline.traffic`.
usage: gemm_classes_json.py classes.log dram.csv out_prefix"""
import csv
import json
import re
import sys


def main():
    log, path, out = sys.argv[1:4]
    classes = []
    for line in open(log):
        m = re.match(r"class (.+) (fwd|dgrad|wgrad) M=(\d+) K=(\d+) N=(\d+) per_step=(\d+) kernels_per_gemm=(\d+)", line)
        if m:
            classes.append(dict(layer=m.group(1), form=m.group(2), M=int(m.group(3)), K=int(m.group(4)), N=int(m.group(5)),
                                per_step=int(m.group(6)), kernels_per_gemm=int(m.group(7))))
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    iid, ik, im, iu, iv = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3,
             "msecond": 1e3}
    launches, order = {}, []
    for r in rows[hi + 1:]:
        if len(r) <= iv or not r[iid].strip().isdigit():
            continue
        k = int(r[iid])
        if k not in launches:
            launches[k] = {"kernel": re.sub(r"\(.*", "", r[ik]).replace("void nnb::<unnamed>::", "")}
            order.append(k)
        launches[k][r[im]] = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
    per = len(order) // max(len(classes), 1)
    assert per * len(classes) == len(order), (len(order), len(classes))
    tot_w, tot_alg, n_gemm = 0.0, 0.0, 0
    for ci, c in enumerate(classes):
        ls = [launches[k] for k in order[ci * per:(ci + 1) * per]][1:]  # drop the warm-up execution
        c["kernel"] = ls[-1]["kernel"]
        c["dram_read"] = sum(l["dram__bytes_read.sum"] for l in ls) / len(ls)
        c["dram_write"] = sum(l["dram__bytes_write.sum"] for l in ls) / len(ls)
        c["ncu_us"] = sum(l["gpu__time_duration.sum"] for l in ls) / len(ls)
        M, K, N = c["M"], c["K"], c["N"]  # bf16 operand planes in, fp32 result out
        c["algorithmic_bytes"] = {"fwd": 2 * (M * K + N * K) + 4 * M * N,      # X, W -> O
                                  "dgrad": 2 * (M * N + N * K) + 4 * M * K,    # dZ, W -> dX
                                  "wgrad": 2 * (M * N + M * K) + 4 * N * K}[c["form"]]  # dZ, X -> dW
        tot_w += (c["dram_read"] + c["dram_write"]) * c["per_step"]
        tot_alg += c["algorithmic_bytes"] * c["per_step"]
        n_gemm += c["per_step"]
    d = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "
                   "regex:gemm_tcgen05 python scripts/gemm_classes_once.py (B200; default cache control = L2 flushed before every "
                   "kernel, so outputs that fit the 126 MB L2 are not yet written back when the kernel ends)",
         "traffic_per_launch_weighted": tot_w / n_gemm, "algorithmic_bytes_per_launch_weighted": tot_alg / n_gemm,
         "gemms_per_step": n_gemm, "classes": classes}
    json.dump(d, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write("# DRAM traffic per launch of the nn.Linear GEMM classes of the GPT-small step (ncu)\n\n" + d["source"] + "\n\n")
        f.write(f"Per-step weighted mean over {n_gemm} GEMM launches: measured {tot_w / n_gemm / 1e6:.2f} MB, algorithmic (bf16 operands "
                f"in, fp32 result out) {tot_alg / n_gemm / 1e6:.2f} MB per launch.\n\n")
        f.write("| layer | form | M x K x N | per step | kernel | DRAM read MB | DRAM write MB | algorithmic MB | ncu us |\n|---|---|---|---|---|---|---|---|---|\n")
        for c in classes:
            f.write(f"| {c['layer']} | {c['form']} | {c['M']} x {c['K']} x {c['N']} | {c['per_step']} | `{c['kernel']}` | {c['dram_read'] / 1e6:.2f} | "
                    f"{c['dram_write'] / 1e6:.2f} | {c['algorithmic_bytes'] / 1e6:.2f} | {c['ncu_us']:.1f} |\n")


if __name__ == "__main__":
    main()
