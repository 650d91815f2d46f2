#!/bin/bash
# One-GPU validation pass: GPU tests, the default bench line, ncu launch lists of every workload and bounded
# `--set full` captures of the non-GEMM kernels. Everything lands in gpurun_out/$1 (default: val).
OUT=gpurun_out/${1:-val}
mkdir -p $OUT
PT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 60 --timeout-method=thread -m gpu"
timeout 300 $PT tests > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest.log
timeout 400 python bench.py --steps 30 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "default bench rc=$?"; head -c 260 $OUT/bench_default.json; echo
for wl in gpt ddpm conv mlp; do
  timeout 120 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$wl.csv python scripts/profile_step.py --workload $wl > $OUT/ncu_$wl.log 2>&1
  python scripts/summarize_launches.py $OUT/launches_$wl.csv > $OUT/launches_$wl.md 2>&1; head -4 $OUT/launches_$wl.md | tail -2
done
timeout 200 ncu --set full --import-source on --clock-control none -f -o $OUT/attn_mma --profile-from-start off \
  -k regex:"attn_fwd_mma_kernel|attn_bwd_mma_kernel" -c 10 python scripts/profile_step.py --workload gpt --warmup 2 > $OUT/ncu_full_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 200 ncu --set full --clock-control none -f -o $OUT/gpt_elem --profile-from-start off \
  -k regex:"dropout_kernel<1, 1>|rmsnorm_bwd_fused|rmsnorm_fwd_fused" -c 3 python scripts/profile_step.py --workload gpt --warmup 2 > $OUT/ncu_full_elem.log 2>&1; echo "ncu elem rc=$?"
du -sh $OUT
