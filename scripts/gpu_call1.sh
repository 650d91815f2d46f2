set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload mlp > gpurun_out/bench_mlp.json 2> gpurun_out/bench_mlp.err; tail -c 600 gpurun_out/bench_mlp.err
timeout 600 python bench.py --workload gpt --steps 20 --warmup 3 > gpurun_out/bench_gpt.json 2> gpurun_out/bench_gpt.err; tail -c 600 gpurun_out/bench_gpt.err
cat gpurun_out/bench_mlp.json gpurun_out/bench_gpt.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_gpt.csv python scripts/profile_step.py --workload gpt > gpurun_out/prof_gpt.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_gpt.csv "GPT step launch list" > gpurun_out/launches_gpt.md; head -30 gpurun_out/launches_gpt.md
timeout 300 python bench.py --workload gpt --steps 3 --warmup 3 --impl reference > gpurun_out/bench_gpt_ref.json 2>&1
