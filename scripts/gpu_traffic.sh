#!/bin/bash
# DRAM bytes per launch of the Linear GEMM classes (source of bench.py's roofline.traffic)
mkdir -p gpurun_out/traffic
timeout 100 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:gemm_tcgen05 --csv --log-file gpurun_out/traffic/dram.csv python scripts/gemm_classes_once.py > gpurun_out/traffic/classes.log 2>&1
echo "rc=$?"; grep -c "^class" gpurun_out/traffic/classes.log; tail -2 gpurun_out/traffic/classes.log
