#!/bin/bash
mkdir -p gpurun_out/attn2
PT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 60 --timeout-method=thread -m gpu"
timeout 300 $PT tests > gpurun_out/attn2/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/attn2/pytest.log
timeout 200 python bench.py --steps 30 --warmup 5 --no-also > gpurun_out/attn2/bench_gpt.json 2> gpurun_out/attn2/bench_gpt.err; echo "bench rc=$?"; head -c 300 gpurun_out/attn2/bench_gpt.json; echo
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/attn2/launches_gpt.csv python scripts/profile_step.py --workload gpt > gpurun_out/attn2/ncu_gpt.log 2>&1
python scripts/summarize_launches.py gpurun_out/attn2/launches_gpt.csv > gpurun_out/attn2/launches_gpt.md 2>&1; head -30 gpurun_out/attn2/launches_gpt.md
timeout 200 ncu --set full --import-source on --clock-control none -f -o gpurun_out/attn2/attn_mma --profile-from-start off \
  -k regex:"attn_fwd_mma_kernel|attn_bwd_mma_kernel" -c 2 python scripts/profile_step.py --workload gpt --warmup 2 > gpurun_out/attn2/ncu_full.log 2>&1; echo "ncu full rc=$?"
du -sh gpurun_out/attn2
