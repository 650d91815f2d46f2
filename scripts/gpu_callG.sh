set -x
mkdir -p gpurun_out
( time timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_gpt_n4.json 2> gpurun_out/bench_gpt_n4.err ) 2>&1 | tail -4
tail -c 600 gpurun_out/bench_gpt_n4.err
cut -c1-250 gpurun_out/bench_gpt_n4.json
python -c "import json;d=json.loads(open('gpurun_out/bench_gpt_n4.json').read().splitlines()[-1]);print(d['config']['grad_allreduce'], d['config']['graph_error'], d['e2e'], d['clocks'])"
