set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_gpt_pdl.json 2> gpurun_out/bench_gpt_pdl.err; tail -c 400 gpurun_out/bench_gpt_pdl.err
cut -c1-200 gpurun_out/bench_gpt_pdl.json; python -c "import json;d=json.loads(open('gpurun_out/bench_gpt_pdl.json').read().splitlines()[-1]);print(d['config'])"
NNB_PDL=0 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_gpt_nopdl.json 2> gpurun_out/bench_gpt_nopdl.err; tail -c 300 gpurun_out/bench_gpt_nopdl.err
cut -c1-200 gpurun_out/bench_gpt_nopdl.json
