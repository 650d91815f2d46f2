#!/bin/bash
mkdir -p gpurun_out/attn3
PT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 60 --timeout-method=thread -m gpu"
timeout 300 $PT tests > gpurun_out/attn3/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/attn3/pytest.log
timeout 200 python bench.py --steps 30 --warmup 5 --no-also > gpurun_out/attn3/bench_gpt.json 2> gpurun_out/attn3/bench_gpt.err; echo "bench rc=$?"; head -c 300 gpurun_out/attn3/bench_gpt.json; echo
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/attn3/launches_gpt.csv python scripts/profile_step.py --workload gpt > gpurun_out/attn3/ncu_gpt.log 2>&1
python scripts/summarize_launches.py gpurun_out/attn3/launches_gpt.csv > gpurun_out/attn3/launches_gpt.md 2>&1; head -34 gpurun_out/attn3/launches_gpt.md
