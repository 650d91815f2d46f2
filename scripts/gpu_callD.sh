set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_gpt_n1.json 2> gpurun_out/bench_gpt_n1.err; tail -c 600 gpurun_out/bench_gpt_n1.err
cut -c1-300 gpurun_out/bench_gpt_n1.json
cd numpy-nn-model_b200/csrc/build
for shape in "4096 512 512" "4096 512 2048" "4096 2048 512"; do
  for sp in 0 1 2 4; do for bn in 0 64 128; do
    echo "--- wgrad shape $shape splits=$sp bn=$bn"; NNB_GEMM_SPLITS=$sp NNB_GEMM_BN=$bn NNB_GEMM_VERBOSE=1 timeout -s KILL 60 ./test_gemm gbench $shape 0 20 2>&1 | grep -E "wgrad|majors=11" | sort -u | cut -c1-150
  done; done
done > ../../../gpurun_out/wgrad_splits.log 2>&1
cd ../../..
cat gpurun_out/wgrad_splits.log | grep -E "^---|gemm wgrad" | paste - - | cut -c1-200
timeout 200 python scripts/conv_ladder.py 2>&1 | head -5
