set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_gpt_simt.json 2> gpurun_out/bench_gpt_simt.err; tail -c 300 gpurun_out/bench_gpt_simt.err
cut -c1-200 gpurun_out/bench_gpt_simt.json
NNB_MATMUL_NO_SIMT=1 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_gpt_nosimt.json 2> gpurun_out/bench_gpt_nosimt.err; tail -c 300 gpurun_out/bench_gpt_nosimt.err
cut -c1-200 gpurun_out/bench_gpt_nosimt.json
