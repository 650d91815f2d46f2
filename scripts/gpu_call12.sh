#!/bin/bash
mkdir -p gpurun_out/c12
PT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 60 --timeout-method=thread -m gpu"
timeout 300 $PT tests > gpurun_out/c12/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/c12/pytest.log
timeout 200 python bench.py --steps 30 --warmup 3 --no-also --no-x3 --watchdog 180 > gpurun_out/c12/bench_gpt.json 2> gpurun_out/c12/bench_gpt.err; echo "bench rc=$?"; head -c 300 gpurun_out/c12/bench_gpt.json; echo
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/c12/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/c12/memcheck.log
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/c12/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/c12/racecheck.log
