#!/bin/bash
# 4-GPU run of the default workload (GPT-small data parallel, chunked all-reduce + chunked optimizer)
mkdir -p gpurun_out/n4
export BENCH_HB_DIR=gpurun_out/n4
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 \
   bench.py --gpus 4 --steps 30 --warmup 3 --watchdog 130 --no-x3 > gpurun_out/n4/gpt4.json 2> gpurun_out/n4/gpt4.err
echo "gpt4 rc=$?"; grep '^{' gpurun_out/n4/gpt4.json | head -c 330; echo
