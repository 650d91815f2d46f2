cd numpy-nn-model_b200/csrc/build
timeout -s KILL 300 ./test_gemm 2>&1 | grep -E "FAIL|correctness|gemm |timed out|error"
for cg in 1 2; do echo "== CG=$cg"; for shape in "4096 1024 4096" "4096 4096 4096" "8192 8192 8192"; do NNB_GEMM_CG=$cg NNB_GEMM_BN=256 timeout -s KILL 120 ./test_gemm bench $shape 0 10 2>&1 | grep "gemm fwd"; done; done
