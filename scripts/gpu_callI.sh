set -x
mkdir -p gpurun_out
for v in 4 8 16; do
( NCCL_MAX_CTAS=$v timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2953$((v%10)) bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_gpt_n4_cta$v.json 2> gpurun_out/bench_gpt_n4_cta$v.err )
python -c "import json;d=json.loads(open('gpurun_out/bench_gpt_n4_cta$v.json').read().splitlines()[-1]);print('NCCL_MAX_CTAS=$v', d['value'], d['ms_per_step'], d['config']['graph_error'])" || tail -c 500 gpurun_out/bench_gpt_n4_cta$v.err
done
