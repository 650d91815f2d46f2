#!/bin/bash
# final-code 8-GPU run of the default workload (GPT-small data parallel)
mkdir -p gpurun_out/n8
export BENCH_HB_DIR=gpurun_out/n8 NCCL_DEBUG=WARN
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 8 --steps 30 --warmup 3 --watchdog 170 --no-x3 > gpurun_out/n8/gpt8.json 2> gpurun_out/n8/gpt8.err
echo "gpt8 rc=$?"; head -c 330 gpurun_out/n8/gpt8.json; echo
