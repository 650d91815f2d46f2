set -x
mkdir -p gpurun_out
run() { tag=$1; shift; ( timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 4 --steps 20 --warmup 3 "$@" > gpurun_out/bench_gpt_n4_$tag.json 2> gpurun_out/bench_gpt_n4_$tag.err ); python -c "import json;d=json.loads(open('gpurun_out/bench_gpt_n4_$tag.json').read().splitlines()[-1]);print('VARIANT $tag', d['value'], d['ms_per_step'], d['config']['graph_error'], d['config']['gemm_sms'])" || tail -c 600 gpurun_out/bench_gpt_n4_$tag.err; }
PORT=29541 run direct
PORT=29542 NCCL_MAX_CTAS=16 run sms132 --gemm-sms 132
