#!/bin/bash
mkdir -p gpurun_out/c10
PT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 60 --timeout-method=thread -m gpu"
timeout 300 $PT tests > gpurun_out/c10/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/c10/pytest.log
timeout 200 python bench.py --steps 30 --warmup 3 --no-also --no-x3 --watchdog 180 > gpurun_out/c10/bench_gpt.json 2> gpurun_out/c10/bench_gpt.err; echo "bench rc=$?"; head -c 300 gpurun_out/c10/bench_gpt.json; echo
timeout 240 python bench.py --workload ddpm --steps 20 --warmup 3 --no-x3 --watchdog 220 > gpurun_out/c10/bench_ddpm.json 2> gpurun_out/c10/bench_ddpm.err; echo "bench ddpm rc=$?"; head -c 300 gpurun_out/c10/bench_ddpm.json; echo
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c10/launches_gpt.csv python scripts/profile_step.py --workload gpt > gpurun_out/c10/ncu_gpt.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/c10/launches_gpt.csv > gpurun_out/c10/launches_gpt.md 2>&1; head -32 gpurun_out/c10/launches_gpt.md
