#!/bin/bash
# two-GPU check: NCCL parity tests (torch.distributed and nnb_comm, chunked optimizer step) + the data-parallel bench lines
mkdir -p gpurun_out/n2
export BENCH_HB_DIR=gpurun_out/n2
timeout 200 python -m pytest -q --tb=short -p no:cacheprovider --timeout 150 --timeout-method=thread -m gpu tests/test_multigpu.py > gpurun_out/n2/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/n2/pytest.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus 2 --steps 30 --warmup 3 --watchdog 170 --no-x3 > gpurun_out/n2/gpt2.json 2> gpurun_out/n2/gpt2.err
echo "gpt2 rc=$?"; grep '^{' gpurun_out/n2/gpt2.json | head -c 330; echo
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
   bench.py --gpus 2 --steps 30 --warmup 3 --watchdog 170 --no-x3 --workload ddpm > gpurun_out/n2/ddpm2.json 2> gpurun_out/n2/ddpm2.err
echo "ddpm2 rc=$?"; grep '^{' gpurun_out/n2/ddpm2.json | head -c 330; echo
