#!/bin/bash
mkdir -p gpurun_out/n2
export BENCH_HB_DIR=gpurun_out/n2 NCCL_DEBUG=WARN
run() { name=$1; np=$2; shift 2
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $np --steps 30 --warmup 3 --watchdog 170 --no-x3 "$@" > gpurun_out/n2/$name.json 2> gpurun_out/n2/$name.err
  rc=$?; echo "$name rc=$rc"; head -c 330 gpurun_out/n2/$name.json; echo; [ $rc -ne 0 ] && tail -5 gpurun_out/n2/$name.err; return 0
}
run gpt2 2
run ddpm2 2 --workload ddpm
timeout 150 python -m pytest tests/test_multigpu.py -x -q -m gpu -p no:cacheprovider --timeout 100 --timeout-method=thread 2>&1 | tail -3
