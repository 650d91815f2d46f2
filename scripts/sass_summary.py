#!/usr/bin/env python
"""Per-kernel SASS evidence for the Blackwell execution model (cuobjdump -sass of the shipped libneunet_b200.so):
counts of tcgen05 MMA (UTCHMMA, .2CTA pairs), TMEM loads (LDTM), TMA loads/stores (UTMALDG / UTMASTG), tensor-map
prefetch (UTMAPF / UTMACCTL), programmatic dependent launch (ACQBULK / PREEXIT) and mbarrier ops (SYNCS) per kernel.
usage: python scripts/sass_summary.py > profiles/sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "numpy-nn-model_b200", "lib", "libneunet_b200.so")
PATS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "ACQBULK", "PREEXIT", "FFMA", "HMMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(anonymous namespace\)::", "", cur)
            cur = re.sub(r"\(.*", "", cur)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for p in PATS:
            if p == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    counts[cur][p] += 1
            elif op.startswith(p):
                counts[cur][p] += 1
    print("# SASS summary of libneunet_b200.so (sm_100a) -- `python scripts/sass_summary.py`\n")
    print("Instruction counts per kernel from `cuobjdump -sass`. UTCHMMA = tcgen05.mma (`.2CTA` = cta_group::2 pairs), LDTM = "
          "tcgen05.ld (TMEM -> registers), UTMALDG / UTMASTG = TMA tensor loads / stores (cp.async.bulk.tensor), UTCBAR = "
          "tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents "
          "(programmatic dependent launch). HMMA (mma.sync; here HMMA.1688.F32.TF32) appears only in the fused attention kernels, whose 64 x 64 x 64 per-head products are too small for a 128-row tcgen05 tile; every other contraction on the tensor path is tcgen05.\n")
    print("| kernel | " + " | ".join(PATS) + " |")
    print("|---|" + "---|" * len(PATS))
    tot = collections.Counter()
    for k, c in counts.items():
        if not any(c[p] for p in PATS):
            continue
        tot.update(c)
        print(f"| `{k[:110]}` | " + " | ".join(str(c[p]) if c[p] else "" for p in PATS) + " |")
    print("| **total** | " + " | ".join(str(tot[p]) for p in PATS) + " |")


if __name__ == "__main__":
    sys.exit(main())
