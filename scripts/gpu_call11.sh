#!/bin/bash
mkdir -p gpurun_out/c11
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/c11/bench_default.json 2> gpurun_out/c11/bench_default.err; echo "default bench rc=$?"; tail -4 gpurun_out/c11/bench_default.err
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/c11/bench_default.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['config']['precision_modes'])
    for k,v in d['config']['also'].items(): print(k, {a:v.get(a) for a in ('value','ms_per_step','error')}, (v.get('roofline') or {}).get('frac'), (v.get('roofline') or {}).get('error'), (v.get('cpu_baseline') or {}).get('value'))
    print('roofline', d['roofline'].get('frac'), d['roofline'].get('bf16x3'), d['roofline'].get('error'))
    for r in d['roofline'].get('ladder',[]): print(r)
    print(d['cpu_baseline'])
except Exception as e:
    print('parse failed', e)
P
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/c11/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/c11/memcheck.log
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/c11/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/c11/racecheck.log
