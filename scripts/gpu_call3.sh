#!/bin/bash
# round 2, 1-GPU call: every step bounded (per-test 60 s watchdog kills the run at the first hung kernel)
mkdir -p gpurun_out/c3
PT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 60 --timeout-method=thread -m gpu"
timeout 200 $PT -v tests/test_fused_gpu.py > gpurun_out/c3/fused.log 2>&1; echo "fused rc=$?"; grep -E "PASS|FAIL|ERROR|passed|failed|Timeout" gpurun_out/c3/fused.log | tail -30
timeout 300 $PT -v tests/test_conv_implicit_gpu.py > gpurun_out/c3/conv.log 2>&1; echo "conv rc=$?"; grep -E "PASS|FAIL|ERROR|passed|failed|Timeout" gpurun_out/c3/conv.log | tail -40
timeout 300 $PT tests --deselect tests/test_fused_gpu.py --deselect tests/test_conv_implicit_gpu.py --ignore tests/test_fused_gpu.py --ignore tests/test_conv_implicit_gpu.py > gpurun_out/c3/rest.log 2>&1; echo "rest rc=$?"; tail -30 gpurun_out/c3/rest.log
timeout 240 python bench.py --steps 20 --warmup 3 --no-also --no-x3 --watchdog 200 > gpurun_out/c3/bench_gpt.json 2> gpurun_out/c3/bench_gpt.err; echo "bench rc=$?"; head -c 700 gpurun_out/c3/bench_gpt.json; tail -3 gpurun_out/c3/bench_gpt.err
