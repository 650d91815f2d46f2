#!/bin/bash
mkdir -p gpurun_out/c15
timeout 400 ncu --set full --import-source on --clock-control none -f -o gpurun_out/c15/gpt_kernels \
  -k regex:"attn_bwd|attn_fwd|rmsnorm_fwd_fused|rmsnorm_bwd_fused|dropout_kernel|adamw_multi|ce_backward_staged|ce_rows|stage_rows|embedding_scatter" \
  --launch-skip 0 -c 60 python scripts/profile_step.py --workload gpt --warmup 2 > gpurun_out/c15/ncu_gpt.log 2>&1; echo "ncu gpt rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none -f -o gpurun_out/c15/ddpm_kernels \
  -k regex:"gemm_tcgen05|bn_reduce|bn_apply|nchw_to_nhwc|stage_weight_klc|convT_interleave|permute_dw" \
  --launch-skip 0 -c 40 python scripts/profile_step.py --workload ddpm --warmup 2 > gpurun_out/c15/ncu_ddpm.log 2>&1; echo "ncu ddpm rc=$?"
ls -la gpurun_out/c15
