#!/bin/bash
OUT=gpurun_out/${1:-ddpm}
mkdir -p $OUT
PT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 60 --timeout-method=thread -m gpu"
timeout 300 $PT tests > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest.log
timeout 200 python bench.py --steps 30 --warmup 5 --no-also --no-x3 --workload ddpm > $OUT/bench_ddpm.json 2> $OUT/bench_ddpm.err; echo "bench rc=$?"; head -c 260 $OUT/bench_ddpm.json; echo
timeout 120 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_ddpm.csv python scripts/profile_step.py --workload ddpm > $OUT/ncu_ddpm.log 2>&1
python scripts/summarize_launches.py $OUT/launches_ddpm.csv > $OUT/launches_ddpm.md 2>&1; head -40 $OUT/launches_ddpm.md
