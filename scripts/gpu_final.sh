set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/final_bench_gpt.json 2> gpurun_out/final_bench_gpt.err; tail -c 300 gpurun_out/final_bench_gpt.err; cut -c1-220 gpurun_out/final_bench_gpt.json
timeout 300 python bench.py --workload mlp > gpurun_out/final_bench_mlp.json 2> gpurun_out/final_bench_mlp.err; tail -c 300 gpurun_out/final_bench_mlp.err; cut -c1-220 gpurun_out/final_bench_mlp.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2>&1; cut -c1-220 gpurun_out/final_bench_ref.json
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_gpt_v3.csv python scripts/profile_step.py --workload gpt > gpurun_out/prof_gpt.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_gpt_v3.csv "Round 1: ncu launch list of ONE eager GPT-small training step (final code state of the round)" > gpurun_out/launches_gpt_v3.md; head -12 gpurun_out/launches_gpt_v3.md
cd numpy-nn-model_b200/csrc/build && timeout -s KILL 300 ./test_gemm 2>&1 | grep -E "FAIL|correctness|gemm |timed out|error" > ../../../gpurun_out/test_gemm_final.log; cd ../../..
for shape in "4096 1024 4096" "16384 512 2048" "8192 4096 4096"; do timeout -s KILL 60 numpy-nn-model_b200/csrc/build/test_gemm gbench $shape 0 20 2>&1 | grep "gemm " ; done > gpurun_out/test_gemm_gbench_final.log
cat gpurun_out/test_gemm_final.log gpurun_out/test_gemm_gbench_final.log | cut -c1-200
