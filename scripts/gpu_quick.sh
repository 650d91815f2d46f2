#!/bin/bash
# shortest useful check: GPU tests + one short GPT bench line
OUT=gpurun_out/${1:-quick}
mkdir -p $OUT
timeout 120 python -m pytest -q --tb=short -p no:cacheprovider --timeout 60 --timeout-method=thread -m gpu tests > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest.log
timeout 100 python bench.py --steps 30 --warmup 3 --no-also --no-x3 > $OUT/bench_gpt.json 2> $OUT/bench_gpt.err; echo "bench rc=$?"; head -c 260 $OUT/bench_gpt.json; echo
