#!/bin/bash
# round 2, 1-GPU call: new fused kernels first, then the whole GPU suite, a bench line and an ncu launch list
mkdir -p gpurun_out/c2
timeout 600 python -m pytest tests/test_fused_gpu.py tests/test_conv_implicit_gpu.py -q -m gpu --tb=short > gpurun_out/c2/fused.log 2>&1; echo "fused rc=$?"; tail -25 gpurun_out/c2/fused.log
timeout 900 python -m pytest tests -q -m gpu --tb=short > gpurun_out/c2/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/c2/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-also --no-x3 > gpurun_out/c2/bench_gpt.json 2> gpurun_out/c2/bench_gpt.err; echo "bench rc=$?"; head -c 1200 gpurun_out/c2/bench_gpt.json; tail -5 gpurun_out/c2/bench_gpt.err
NEUNET_B200_FUSE=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-also --no-x3 > gpurun_out/c2/bench_gpt_nofuse.json 2> gpurun_out/c2/bench_gpt_nofuse.err; echo "bench nofuse rc=$?"; head -c 400 gpurun_out/c2/bench_gpt_nofuse.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c2/launches_gpt.csv python scripts/profile_step.py --workload gpt > gpurun_out/c2/ncu.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/c2/launches_gpt.csv > gpurun_out/c2/launches_gpt.md 2>&1; head -40 gpurun_out/c2/launches_gpt.md
timeout 900 python bench.py --steps 40 --warmup 3 > gpurun_out/c2/bench_full.json 2> gpurun_out/c2/bench_full.err; echo "bench full rc=$?"; tail -3 gpurun_out/c2/bench_full.err; python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/c2/bench_full.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['config']['precision_modes'])
    for k,v in d['config']['also'].items(): print(k, {a:v.get(a) for a in ('value','ms_per_step','error')}, (v.get('roofline') or {}).get('frac'), (v.get('roofline') or {}).get('error'))
    print('roofline', d['roofline'].get('frac'), d['roofline'].get('bf16x3'), d['roofline'].get('error'))
    for r in d['roofline'].get('ladder',[]): print(r)
except Exception as e:
    print('parse failed', e)
P
