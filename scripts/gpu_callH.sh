set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_gpt_h.json 2> gpurun_out/bench_gpt_h.err; tail -c 300 gpurun_out/bench_gpt_h.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_gpt_h.json').read().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline']['frac']);[print(c['layer'],c['form'],c['us']) for c in d['roofline']['classes']]"
