#!/bin/bash
mkdir -p gpurun_out/c14
PT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 60 --timeout-method=thread -m gpu"
timeout 300 $PT tests > gpurun_out/c14/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/c14/pytest.log
for w in conv mlp; do
  timeout 200 python bench.py --workload $w --steps 30 --warmup 3 --watchdog 180 > gpurun_out/c14/bench_$w.json 2> gpurun_out/c14/bench_$w.err; echo "bench $w rc=$?"; head -c 260 gpurun_out/c14/bench_$w.json; echo
done
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c14/launches_conv.csv python scripts/profile_step.py --workload conv > gpurun_out/c14/ncu_conv.log 2>&1
python scripts/summarize_launches.py gpurun_out/c14/launches_conv.csv > gpurun_out/c14/launches_conv.md 2>&1; head -14 gpurun_out/c14/launches_conv.md
