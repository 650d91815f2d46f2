#!/bin/bash
# compute-sanitizer over the small-shape pass of every kernel family (scripts/sanitize_small.py)
OUT=gpurun_out/${1:-san}
mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  timeout 280 compute-sanitizer --tool $tool python scripts/sanitize_small.py > $OUT/$tool.log 2>&1; echo "$tool rc=$?"; tail -3 $OUT/$tool.log
done
