#!/bin/bash
# 8-GPU check of bench.py (round 2, call 1): heartbeat logs per rank, bounded by `timeout`; fallbacks only if the first fails.
mkdir -p gpurun_out/n8
export BENCH_HB_DIR=gpurun_out/n8 NCCL_DEBUG=WARN NEUNET_B200_FUSE=0
run() { # name, extra args
  name=$1; shift
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 8 --steps 20 --warmup 3 --watchdog 150 "$@" > gpurun_out/n8/$name.json 2> gpurun_out/n8/$name.err
  rc=$?; echo "$name rc=$rc"; tail -c 600 gpurun_out/n8/$name.json; return $rc
}
run full || { run nooverlap --no-overlap; run nograph --no-graph; }
timeout 200 python -m pytest tests/test_multigpu.py -x -q -m gpu 2>&1 | tail -5
nvidia-smi --query-gpu=index,utilization.gpu,memory.used --format=csv | head -10
