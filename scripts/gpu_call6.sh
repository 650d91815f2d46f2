#!/bin/bash
mkdir -p gpurun_out/c6
PT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 90 --timeout-method=thread -m gpu"
for v in default "NNB_CONV_IMPLICIT=0" "NEUNET_B200_FUSE=0"; do
  echo "== ddpm full-size, $v"; env $( [ "$v" = default ] || echo $v ) timeout 120 $PT tests/test_ddpm_fullsize_gpu.py -k unet 2>&1 | grep -E "AssertionError|passed|failed" | cut -c1-700
done
timeout 400 ncu --set full --import-source on -k regex:attn_ -c 2 --clock-control none -f -o gpurun_out/c6/attn python scripts/profile_step.py --workload gpt --warmup 1 > gpurun_out/c6/ncu_attn.log 2>&1; echo "ncu attn rc=$?"; ls -la gpurun_out/c6/
