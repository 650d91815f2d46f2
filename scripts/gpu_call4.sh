set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_gpt_n1.json 2> gpurun_out/bench_gpt_n1.err; tail -c 1000 gpurun_out/bench_gpt_n1.err
cut -c1-700 gpurun_out/bench_gpt_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_gpt_n2.json 2> gpurun_out/bench_gpt_n2.err; tail -c 1500 gpurun_out/bench_gpt_n2.err
cut -c1-900 gpurun_out/bench_gpt_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-overlap > gpurun_out/bench_gpt_n2_noov.json 2> gpurun_out/bench_gpt_n2_noov.err; tail -c 600 gpurun_out/bench_gpt_n2_noov.err
cut -c1-900 gpurun_out/bench_gpt_n2_noov.json
timeout 300 python bench.py --workload mlp > gpurun_out/bench_mlp.json 2> gpurun_out/bench_mlp.err; cut -c1-500 gpurun_out/bench_mlp.json
timeout 600 python scripts/conv_ladder.py > gpurun_out/conv_ladder.md 2> gpurun_out/conv_ladder.err; cat gpurun_out/conv_ladder.md; tail -5 gpurun_out/conv_ladder.err
