#!/bin/bash
# 8-GPU check (round 2): GPT-small and DDPM data-parallel benches, the 2-rank NCCL parity tests (torch.distributed and nnb_comm)
mkdir -p gpurun_out/n8b
export BENCH_HB_DIR=gpurun_out/n8b NCCL_DEBUG=WARN
run() { # name, nproc, extra args
  name=$1; np=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $np --steps 30 --warmup 3 --watchdog 200 --no-x3 "$@" > gpurun_out/n8b/$name.json 2> gpurun_out/n8b/$name.err
  rc=$?; echo "$name rc=$rc"; head -c 420 gpurun_out/n8b/$name.json; echo; [ $rc -ne 0 ] && tail -5 gpurun_out/n8b/$name.err; return $rc
}
run gpt8 8 || run gpt8_nooverlap 8 --no-overlap
run ddpm8 8 --workload ddpm
run gpt4 4
timeout 200 python -m pytest tests/test_multigpu.py -x -q -m gpu -p no:cacheprovider --timeout 120 --timeout-method=thread 2>&1 | tail -5
