set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python scripts/conv_ladder.py > gpurun_out/conv_ladder_v2.md 2> gpurun_out/conv_ladder.err; cat gpurun_out/conv_ladder_v2.md; tail -5 gpurun_out/conv_ladder.err
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_gpt_n1.json 2> gpurun_out/bench_gpt_n1.err; tail -c 600 gpurun_out/bench_gpt_n1.err
cut -c1-300 gpurun_out/bench_gpt_n1.json
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_tcgen05_kernel<\(int\)256, \(bool\)0, \(bool\)0" -c 1 -o gpurun_out/ncu_gemm_fc1_fwd python scripts/profile_step.py --workload gpt > gpurun_out/ncu1.log 2>&1; tail -3 gpurun_out/ncu1.log
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_tcgen05_kernel<\(int\)128, \(bool\)0, \(bool\)0" -s 4 -c 1 -o gpurun_out/ncu_gemm_512_fwd python scripts/profile_step.py --workload gpt > gpurun_out/ncu3.log 2>&1; tail -3 gpurun_out/ncu3.log
for f in ncu_gemm_fc1_fwd ncu_gemm_512_fwd; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
ls -la gpurun_out | head -40
