#!/bin/bash
mkdir -p gpurun_out/c16
PT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 60 --timeout-method=thread -m gpu"
timeout 300 $PT tests > gpurun_out/c16/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c16/pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/c16/bench_default.json 2> gpurun_out/c16/bench_default.err; echo "default bench rc=$?"; head -c 250 gpurun_out/c16/bench_default.json; echo
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c16/launches_gpt.csv python scripts/profile_step.py --workload gpt > gpurun_out/c16/ncu_gpt.log 2>&1
python scripts/summarize_launches.py gpurun_out/c16/launches_gpt.csv > gpurun_out/c16/launches_gpt.md 2>&1; head -8 gpurun_out/c16/launches_gpt.md
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c16/launches_ddpm.csv python scripts/profile_step.py --workload ddpm > gpurun_out/c16/ncu_ddpm.log 2>&1
python scripts/summarize_launches.py gpurun_out/c16/launches_ddpm.csv > gpurun_out/c16/launches_ddpm.md 2>&1; head -8 gpurun_out/c16/launches_ddpm.md
# one --set full capture per kernel class, ONE launch each (bounded: ~10 s per kernel)
timeout 240 ncu --set full --clock-control none -f -o gpurun_out/c16/gpt_k --profile-from-start off \
  -k regex:"attn_bwd_kernel|rmsnorm_bwd_fused|adamw_multi|dropout_kernel|ce_backward_staged" -c 5 python scripts/profile_step.py --workload gpt --warmup 2 > gpurun_out/c16/ncu_full_gpt.log 2>&1; echo "ncu full gpt rc=$?"
timeout 240 ncu --set full --clock-control none -f -o gpurun_out/c16/ddpm_k --profile-from-start off \
  -k regex:"bn_reduce_kernel|bn_apply_kernel|nchw_to_nhwc" -c 3 python scripts/profile_step.py --workload ddpm --warmup 2 > gpurun_out/c16/ncu_full_ddpm.log 2>&1; echo "ncu full ddpm rc=$?"
timeout 240 ncu --set full --clock-control none -f -o gpurun_out/c16/ddpm_gemm --profile-from-start off \
  -k regex:"gemm_tcgen05_kernel<256, 0, 0, 2>" -c 2 python scripts/profile_step.py --workload ddpm --warmup 2 > gpurun_out/c16/ncu_full_ddpm_gemm.log 2>&1; echo "ncu full ddpm gemm rc=$?"
du -sh gpurun_out/c16
