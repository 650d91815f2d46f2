set -x
mkdir -p gpurun_out
( time timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_gpt_n2.json 2> gpurun_out/bench_gpt_n2.err ) 2>&1 | tail -4; echo "rc=$?"
tail -c 700 gpurun_out/bench_gpt_n2.err
cut -c1-250 gpurun_out/bench_gpt_n2.json
( time timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err ) 2>&1 | tail -4
cut -c1-250 gpurun_out/bench_ref_n2.json; tail -c 300 gpurun_out/bench_ref_n2.err
