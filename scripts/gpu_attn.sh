#!/bin/bash
mkdir -p gpurun_out/attn
PT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 60 --timeout-method=thread -m gpu"
timeout 200 $PT tests/test_fused_gpu.py > gpurun_out/attn/pytest_fused.log 2>&1; echo "fused rc=$?"; tail -15 gpurun_out/attn/pytest_fused.log
timeout 100 python scripts/time_attention.py 2>&1 | tail -3
NEUNET_B200_ATTN_SIMT=1 timeout 100 python scripts/time_attention.py 2>&1 | tail -3
timeout 300 $PT tests > gpurun_out/attn/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/attn/pytest.log
timeout 200 python bench.py --steps 30 --warmup 5 --no-also > gpurun_out/attn/bench_gpt.json 2> gpurun_out/attn/bench_gpt.err; echo "bench rc=$?"; head -c 300 gpurun_out/attn/bench_gpt.json; echo
