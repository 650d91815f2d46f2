#!/bin/bash
# round 2, 1-GPU call 4: fixed tests, A/B of the fused step, launch lists, the other workloads
mkdir -p gpurun_out/c4
PT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 60 --timeout-method=thread -m gpu"
timeout 300 $PT tests > gpurun_out/c4/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/c4/pytest.log
timeout 200 python bench.py --steps 30 --warmup 3 --no-also --no-x3 --watchdog 180 > gpurun_out/c4/bench_gpt.json 2> gpurun_out/c4/bench_gpt.err; echo "bench rc=$?"; head -c 300 gpurun_out/c4/bench_gpt.json; echo
NEUNET_B200_FUSE=0 timeout 200 python bench.py --steps 30 --warmup 3 --no-also --no-x3 --watchdog 180 > gpurun_out/c4/bench_gpt_nofuse.json 2> gpurun_out/c4/bench_gpt_nofuse.err; echo "bench nofuse rc=$?"; head -c 300 gpurun_out/c4/bench_gpt_nofuse.json; echo
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c4/launches_gpt.csv python scripts/profile_step.py --workload gpt > gpurun_out/c4/ncu_gpt.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/c4/launches_gpt.csv > gpurun_out/c4/launches_gpt.md 2>&1; head -45 gpurun_out/c4/launches_gpt.md
for w in ddpm conv mlp; do
  timeout 240 python bench.py --workload $w --steps 20 --warmup 3 --no-x3 --watchdog 220 > gpurun_out/c4/bench_$w.json 2> gpurun_out/c4/bench_$w.err; echo "bench $w rc=$?"; head -c 250 gpurun_out/c4/bench_$w.json; echo; tail -2 gpurun_out/c4/bench_$w.err
done
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c4/launches_ddpm.csv python scripts/profile_step.py --workload ddpm > gpurun_out/c4/ncu_ddpm.log 2>&1; echo "ncu ddpm rc=$?"
python scripts/summarize_launches.py gpurun_out/c4/launches_ddpm.csv > gpurun_out/c4/launches_ddpm.md 2>&1; head -40 gpurun_out/c4/launches_ddpm.md
