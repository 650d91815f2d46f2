# GPU-only pacing (graph replay) of the GPT-small Linear shapes at M = B*T = 4096, per tile config.
cd numpy-nn-model_b200/csrc/build
run() { timeout -s KILL 60 ./test_gemm gbench $1 $2 $3 0 20 2>&1 | grep -E "gemm |nnb gemm|error|FAIL" ; }
for shape in "4096 512 512" "4096 512 2048" "4096 2048 512" "4096 512 15000" "16384 512 2048"; do
  echo "=== shape $shape (model choice)"; NNB_GEMM_VERBOSE=1 run $shape | sort -u
  for cfg in "64 1" "128 1" "256 1" "128 2" "256 2"; do set -- $cfg
    echo "--- BN=$1 CG=$2"; NNB_GEMM_BN=$1 NNB_GEMM_CG=$2 run $shape
  done
done
echo "=== epilogue split, 4096 512 2048 BN=256 CG=2: debug 0 / 1 (no stores) / 2 (no tmem loads, no stores)"
for d in 0 1 2; do NNB_GEMM_DEBUG=$d NNB_GEMM_BN=256 NNB_GEMM_CG=2 run 4096 512 2048 | grep fwd; done
for d in 0 1 2; do NNB_GEMM_DEBUG=$d NNB_GEMM_BN=128 NNB_GEMM_CG=1 run 4096 512 2048 | grep fwd; done
echo "=== host-paced (bench) vs graph-paced, 4096 512 512"
timeout -s KILL 60 ./test_gemm bench 4096 512 512 0 50 | grep "gemm "
