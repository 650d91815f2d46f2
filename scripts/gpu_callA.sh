set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_gpt_n1.json 2> gpurun_out/bench_gpt_n1.err; tail -c 600 gpurun_out/bench_gpt_n1.err
cut -c1-300 gpurun_out/bench_gpt_n1.json
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_gpt_v2.csv python scripts/profile_step.py --workload gpt > gpurun_out/prof_gpt.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_gpt_v2.csv "GPT step launch list v2" > gpurun_out/launches_gpt_v2.md; head -32 gpurun_out/launches_gpt_v2.md
timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tcgen05 --csv --log-file gpurun_out/gemm_classes_dram.csv python scripts/gemm_classes_once.py > gpurun_out/gemm_classes_once.log 2>&1; tail -13 gpurun_out/gemm_classes_once.log
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_tcgen05_kernel<256, 0, 0, 2>" -c 1 -o gpurun_out/ncu_gemm_fc1_fwd python scripts/profile_step.py --workload gpt > gpurun_out/ncu1.log 2>&1; tail -3 gpurun_out/ncu1.log
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:stage_rows_kernel" -s 30 -c 1 -o gpurun_out/ncu_stage_rows python scripts/profile_step.py --workload gpt > gpurun_out/ncu2.log 2>&1; tail -3 gpurun_out/ncu2.log
for f in ncu_gemm_fc1_fwd ncu_stage_rows; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
ls -la gpurun_out
