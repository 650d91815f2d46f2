mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python bench.py --steps 20 --warmup 3 > gpurun_out/last_bench_gpt.json 2> gpurun_out/last_bench_gpt.err; tail -c 200 gpurun_out/last_bench_gpt.err; cut -c1-180 gpurun_out/last_bench_gpt.json
