#!/bin/bash
mkdir -p gpurun_out/c13
for w in conv mlp; do
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c13/launches_$w.csv python scripts/profile_step.py --workload $w > gpurun_out/c13/ncu_$w.log 2>&1; echo "ncu $w rc=$?"
python scripts/summarize_launches.py gpurun_out/c13/launches_$w.csv > gpurun_out/c13/launches_$w.md 2>&1; head -40 gpurun_out/c13/launches_$w.md
done
