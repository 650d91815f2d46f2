/*
 * neunet_b200.h -- C-ABI of libneunet_b200.so: the B200-native (sm_100a) dense forward/backward
 * hot path of AkiRusProd/numpy-nn-model ("neunet").
 *
 * This is the boundary a maintainer of the reference binds with ctypes.
 * It follows the conventions of the reference's own native plug-ins (neunet/nn/experimental):
 *   - plain `extern "C"` functions, raw DEVICE pointers, C-contiguous fp32 unless stated,
 *     the caller allocates every input, output and scratch buffer
 *     (reference: experimental/utils.py:64-85, experimental/linear/linear.py:126-150,184-212);
 *   - outputs are overwritten, never accumulated (accumulation stays in Tensor.apply_grad,
 *     neunet/autograd.py:85-93);
 *   - trailing `cudaStream_t` like cudaLinearSwishForward / FusedAdamWStep.
 * Deliberate fixes over the reference ABI: every function returns an int status instead of
 * printf+exit (linear_cublaslt_no_manual_mem.cu:91-94), dims are int64_t (the reference declares
 * `int` in C and c_size_t in Python, experimental/linear/linear.py:38-65), and all symbols carry a
 * unique `nnb_` prefix (the reference's .so files export clashing names under RTLD_GLOBAL).
 *
 * Numerics: tensor-core contractions take bf16 operands with fp32 accumulation.
 *   NNB_PREC_BF16    one bf16 product            (~2e-3 rel. vs fp32; throughput mode)
 *   NNB_PREC_BF16X3  x = hi + lo split, hi*hi + hi*lo + lo*hi  (~1e-5 rel.; parity mode)
 * Everything that is not a contraction (epilogues, normalisation, optimizer) is fp32.
 */
#ifndef NEUNET_B200_H
#define NEUNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

/* status codes */
#define NNB_OK 0
#define NNB_ERR_INVALID 1     /* bad argument (null pointer, non-positive dim, misaligned buffer) */
#define NNB_ERR_CUDA 2        /* a CUDA runtime/driver call failed */
#define NNB_ERR_WORKSPACE 3   /* workspace too small; call the matching *_workspace_bytes() */
#define NNB_ERR_UNSUPPORTED 4 /* shape/option outside what the kernels implement (never a CPU fallback) */

#define NNB_PREC_BF16 0
#define NNB_PREC_BF16X3 1

#define NNB_ACT_NONE 0
#define NNB_ACT_SWISH 1 /* x * sigmoid(beta * x), neunet/nn/activations.py:221-233 */

#define NNB_OPT_ADAM_L2 0 /* grad += wd * p          (neunet/optim.py:24-25) */
#define NNB_OPT_ADAMW 1   /* p -= lr * wd * p first  (neunet/optim.py:59-60) */

/* ---- library ------------------------------------------------------------------------------ */
const char* nnb_last_error(void);
int nnb_version(void);
/* Fails with NNB_ERR_UNSUPPORTED unless the current device is compute capability 10.x. */
int nnb_device_check(int* sm_count, int* cc_major, int* cc_minor);
/* Number of kernels this library has launched since the last reset (bench.py's `gpu_launches`). */
/* Programmatic dependent launch for the library's hot-path kernels (on by default, NNB_PDL=0 in the
 * environment turns it off): returns the previous setting. Results are identical either way. */
int nnb_set_pdl(int on);
/* Size persistent grids (GEMM) for at most `sms` SMs (0 = all): leaves room for NCCL's CTAs during
 * data-parallel training. Returns the previous budget. NNB_SM_BUDGET in the environment sets the default. */
int nnb_set_sm_budget(int sms);
uint64_t nnb_launch_count(void);
void nnb_launch_count_reset(void);

/* ---- nn.Linear ------------------------------------------------------------------------------
 * Replaces cudaLinearModuleForward / cudaLinearModuleBackward
 * (neunet/nn/experimental/linear/linear_cublaslt_no_manual_mem.cu:114-140, 142-184) and the fused
 * cudaLinearSwishForward / cudaLinearSwishBackward
 * (neunet/nn/experimental/linear_swish/linear_swish_cutlass_evt_full.cu:558-663, 680-818).
 * Semantics are those of neunet/nn/layers/linear.py:17-24,48-58:
 *   O[M,N] = act(X[M,K] . W[N,K]^T + bias[N])         Z (optional) receives the pre-activation
 *   dX[M,K] = dZ . W      dW[N,K] = dZ^T . X      db[N] = sum_rows dZ,
 *   with dZ = dO (act NONE) or dO * swish'(Z) (act SWISH; Z must be the saved pre-activation).
 * bias, Z, dX, db may be NULL (dX NULL skips the dgrad GEMM).
 * W_staged: NULL, or a buffer filled by nnb_stage_weight() for the CURRENT contents of W
 * (lets the caller convert weights to bf16 once per optimizer step instead of once per call).
 * X_staged_out (forward): NULL, or a 256-byte aligned buffer of nnb_weight_staged_bytes(M, K, prec)
 * bytes that receives the bf16 planes of X; hand it back as X_staged to nnb_linear_backward to skip
 * re-converting X there (X itself may then be NULL). nnb_stage_weight() converts ANY row-major
 * fp32 matrix, not only weights; nnb_linear_forward_staged() runs the GEMM on two staged operands.
 */
size_t nnb_linear_workspace_bytes(int64_t M, int64_t K, int64_t N, int prec, int backward);
int nnb_linear_forward(const float* X, const float* W, const float* bias, float* O, float* Z,
                       int64_t M, int64_t K, int64_t N, int act, float beta, int prec,
                       const void* W_staged, void* X_staged_out, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream);
int nnb_linear_forward_staged(const void* X_staged, const void* W_staged, const float* bias, float* O,
                              float* Z, int64_t M, int64_t K, int64_t N, int act, float beta,
                              int prec, void* workspace, size_t workspace_bytes,
                              cudaStream_t stream);
int nnb_linear_backward(const float* X, const float* W, const float* Z, const float* dO,
                        float* dX, float* dW, float* db, int64_t M, int64_t K, int64_t N, int act,
                        float beta, int prec, const void* W_staged, const void* X_staged,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* nnb_linear_backward on an upstream gradient that still has to pass an nn.Dropout (neunet/nn/layers/dropout.py:39-46:
 * dx = grad * mask): dO is multiplied by the dropout mask of the Philox ticket (drop_p, seed, call_id, epoch | *epoch_dev;
 * flat index over the contiguous [M, N] matrix, same bits as nnb_dropout) inside the staging pass that converts dO to
 * bf16 planes, before swish' and the bias column sums -- the stand-alone mask pass over dO (one read + one write of
 * M*N floats) disappears. N % 4 == 0, otherwise NNB_ERR_UNSUPPORTED (run nnb_dropout, then nnb_linear_backward). */
int nnb_linear_backward_dropped(const float* X, const float* W, const float* Z, const float* dO,
                                float* dX, float* dW, float* db, int64_t M, int64_t K, int64_t N, int act,
                                float beta, int prec, const void* W_staged, const void* X_staged,
                                void* workspace, size_t workspace_bytes, float drop_p, uint64_t seed, uint32_t call_id,
                                uint64_t epoch, const uint64_t* epoch_dev, cudaStream_t stream);
size_t nnb_weight_staged_bytes(int64_t rows, int64_t cols, int prec);
int nnb_stage_weight(const float* W, int64_t rows, int64_t cols, int prec, void* dst,
                     cudaStream_t stream);

/* ---- Tensor.matmul --------------------------------------------------------------------------
 * The reference has no native matmul entry point (neunet/autograd.py:192-230 calls xp.matmul);
 * this one is defined in the same style. Operands are 4-D strided fp32 views
 * A[b0,b1,M,K], B[b0,b1,K,N] given by ELEMENT strides {s_b0, s_b1, s_row, s_col} (0 = broadcast,
 * transposed views allowed); C / dA / dB are C-contiguous [b0,b1,M,N] / [b0,b1,M,K] / [b0,b1,K,N].
 *   forward : C  = alpha * A . B
 *   backward: dA = alpha * G . B^T   (autograd.py:209)     dB = alpha * A^T . G   (autograd.py:211)
 * dA or dB may be NULL. Un-broadcasting a gradient stays in Tensor._reverse_broadcast.
 * A_staged_out / B_staged_out (forward): NULL, or 256-byte aligned buffers of
 * nnb_matmul_staged_bytes(b0, b1, M, K, prec) / (b0, b1, K, N, prec) bytes that receive the bf16 planes
 * of the operands; handed back as A_staged / B_staged to nnb_matmul_backward (same strides, same
 * prec) they save re-converting A and B there.
 */
size_t nnb_matmul_staged_bytes(int64_t b0, int64_t b1, int64_t rows, int64_t cols, int prec);
/* Opt-in fp32 CUDA-core path for small batched products (>= 16 batch elements with M, K, N <= 128, e.g. the
 * 64 x 64 x 64 attention products of examples/gpt.ipynb): exact fp32, one launch, no staging, no workspace;
 * prec and the staged-plane arguments are ignored there. Off by default (the tcgen05 path measured faster
 * inside the GPT step); nnb_matmul_set_small_path(1) turns it on process-wide and returns the previous
 * setting. nnb_matmul_uses_tensor_cores() tells which path a call with these sizes takes (1 = tcgen05). */
int nnb_matmul_set_small_path(int on);
int nnb_matmul_uses_tensor_cores(int64_t b0, int64_t b1, int64_t M, int64_t K, int64_t N);
size_t nnb_matmul_workspace_bytes(int64_t b0, int64_t b1, int64_t M, int64_t K, int64_t N,
                                  int prec, int backward);
int nnb_matmul_forward(const float* A, const int64_t a_strides[4], const float* B,
                       const int64_t b_strides[4], float* C, int64_t b0, int64_t b1, int64_t M,
                       int64_t K, int64_t N, float alpha, int prec, void* A_staged_out,
                       void* B_staged_out, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream);
int nnb_matmul_backward(const float* A, const int64_t a_strides[4], const float* B,
                        const int64_t b_strides[4], const float* G, float* dA, float* dB,
                        int64_t b0, int64_t b1, int64_t M, int64_t K, int64_t N, float alpha,
                        int prec, const void* A_staged, const void* B_staged, void* workspace,
                        size_t workspace_bytes, cudaStream_t stream);

/* ---- nn.Conv2d ------------------------------------------------------------------------------
 * NCHW cross-correlation exactly as neunet/nn/layers/conv2d.py:297-355 (forward) and 16-117
 * (backward): X[B,Cin,H,W], Wt[Cout,Cin,kh,kw], bias[Cout] or NULL, O[B,Cout,Ho,Wo] with
 * Ho = (H + pad[0] + pad[1] - dil[0]*(kh-1) - 1) / stride[0] + 1 (conv2d.py:246-260);
 * pad = {top, bottom, left, right} (the 4-tuple Conv2d.build() produces, conv2d.py:237-243).
 * The reference has no native conv entry point. dX or db may be NULL.
 */
typedef struct nnb_conv2d_desc {
    int64_t B, Cin, H, W, Cout, kh, kw;
    int32_t stride[2];
    int32_t pad[4];
    int32_t dil[2];
} nnb_conv2d_desc;
int nnb_conv2d_out_shape(const nnb_conv2d_desc* d, int64_t* Ho, int64_t* Wo);
size_t nnb_conv2d_workspace_bytes(const nnb_conv2d_desc* d, int prec, int backward);
int nnb_conv2d_forward(const nnb_conv2d_desc* d, const float* X, const float* Wt, const float* bias,
                       float* O, int prec, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream);
int nnb_conv2d_backward(const nnb_conv2d_desc* d, const float* X, const float* Wt, const float* dO,
                        float* dX, float* dW, float* db, int prec, void* workspace,
                        size_t workspace_bytes, cudaStream_t stream);
/* Same contractions with the channels-last bf16 planes of X handled by the caller (round 2: "keep forward planes
 * for backward"). Layers whose channel counts are multiples of 8 convert X once to planes [B][H][W][Cin]
 * (nnb_conv2d_planes_bytes(desc, prec) bytes, 256-byte aligned; 0 = this geometry does not use planes):
 *   forward_ex : X_planes != NULL -> read them (X is not converted again); else X_planes_out != NULL -> the converted
 *                planes are written there and stay valid for backward_ex; both NULL = workspace scratch.
 *   backward_ex: X_planes != NULL -> wgrad reads them instead of converting X again.
 * With Cin % 64 == 0 (forward, wgrad) / Cout % 64 == 0 and stride 1 (dgrad) and a power-of-two style position grid the
 * GEMM reads its operand STRAIGHT from the planes through 4-D TMA boxes (implicit GEMM: no `col` matrix; the conv
 * padding is the copy engine's out-of-bounds zero fill, the conv stride its traversal stride);
 * NNB_CONV_IMPLICIT=0 in the environment forces the materialised-`col` path. */
size_t nnb_conv2d_planes_bytes(const nnb_conv2d_desc* desc, int prec);
int nnb_conv2d_forward_ex(const nnb_conv2d_desc* desc, const float* X, const float* W, const float* bias,
                          float* O, int prec, const void* X_planes, void* X_planes_out, void* workspace,
                          size_t workspace_bytes, cudaStream_t stream);
int nnb_conv2d_backward_ex(const nnb_conv2d_desc* desc, const float* X, const float* W, const float* dO,
                           float* dX, float* dW, float* db, int prec, const void* X_planes, void* workspace,
                           size_t workspace_bytes, cudaStream_t stream);

/* ---- epilogue ops as standalone kernels (same maths as the fused forms) ----------------------
 * Swish: cudaSwishForward/Backward (experimental/activations/swish/swish.cu:14-80),
 *        semantics neunet/nn/activations.py:208-233.
 * Softmax over the middle axis of a [outer, n, inner] view: cudaSoftmaxForward/Backward
 *        (experimental/activations/softmax/softmax.cu:144-151, 229-237), activations.py:437-459.
 * RMSNorm: RMSNormForward/Backward (experimental/rmsnorm/rmsnorm.cu:116-140, 282-308),
 *        semantics neunet/nn/layers/rmsnorm.py:39-94. dw/db are full column sums over rows.
 */
/* Fused forms (north_star: "fused with the bias/Swish/RMSNorm/Softmax epilogues"):
 * nnb_rmsnorm_forward_fused: S = X (+ dropout(A) when A != NULL, written to S), Y = rmsnorm(S) * w (+ b), X_std,
 *        and optionally the bf16 planes of Y for the next nnb_linear_forward_staged (RMSNorm as the Linear's
 *        prologue: the GEMM operand is produced by the norm itself). cols <= 1024, cols % 4 == 0 (8 with planes),
 *        else NNB_ERR_UNSUPPORTED (callers then use the plain entry points).
 * nnb_rmsnorm_backward_acc: like nnb_rmsnorm_backward, dX = dX_add + d/dX when dX_add != NULL (the residual
 *        branch's gradient is accumulated in the same pass instead of a separate add, autograd.py:85-93). */
int nnb_rmsnorm_forward_fused(const float* X, const float* A, float p, uint64_t seed, uint32_t call_id,
                              uint64_t epoch, const uint64_t* epoch_dev, float* S, const float* w, const float* b,
                              float* Y, float* X_std, void* Y_staged_out, int prec, int64_t rows, int64_t cols,
                              float eps, cudaStream_t stream);
int nnb_rmsnorm_backward_acc(const float* gY, const float* X, const float* w, const float* X_std,
                             const float* X_norm, const float* dX_add, float* dX, float* dw, float* db,
                             int64_t rows, int64_t cols, void* workspace, size_t workspace_bytes,
                             cudaStream_t stream);
int nnb_swish_forward(const float* x, float* y, int64_t n, float beta, cudaStream_t stream);
int nnb_swish_backward(const float* x, const float* grad, float* dx, int64_t n, float beta,
                       cudaStream_t stream);
int nnb_softmax_forward(const float* x, float* y, int64_t outer, int64_t n, int64_t inner,
                        cudaStream_t stream);
int nnb_softmax_backward(const float* y, const float* grad, float* dx, int64_t outer, int64_t n,
                         int64_t inner, cudaStream_t stream);
int nnb_rmsnorm_forward(const float* X, const float* w, const float* b, float* Y, float* X_std,
                        float* X_norm, int64_t rows, int64_t cols, float eps, cudaStream_t stream);
size_t nnb_rmsnorm_workspace_bytes(int64_t rows, int64_t cols);
int nnb_rmsnorm_backward(const float* gY, const float* X, const float* w, const float* X_std,
                         const float* X_norm, float* dX, float* dw, float* db, int64_t rows,
                         int64_t cols, void* workspace, size_t workspace_bytes,
                         cudaStream_t stream);

/* ---- measurement (bench.py roofline) -----------------------------------------------------------------
 * GPU-paced microseconds of ONE tcgen05 GEMM launch in an nn.Linear form (0 fwd X.W^T [+bias],
 * 1 dgrad dO.W, 2 wgrad dO^T.X) on staged bf16 operands: `sets` operand/output sets (choose them to
 * exceed the 126 MB L2 in total) x `rounds`, captured into a CUDA graph and timed by CUDA events on
 * a private stream created for the measurement (ordered after the work already queued on `stream`). us_per_launch includes the split-K finishing kernel when one is used
 * (launches_per_gemm = 2). with_bias: bit 0 = add bias, bit 1 = Swish epilogue + Z side output
 * (form 0 only), bit 2 = NNB_PREC_BF16X3 operands (hi + lo planes, three products). Diagnostics only. */
int nnb_probe_linear_gemm(int64_t M, int64_t K, int64_t N, int form, int with_bias, int sets,
                          int rounds, float* us_per_launch, int* launches_per_gemm,
                          cudaStream_t stream);

/* ---- lm_head + CrossEntropy (row N3 of SURVEY.md 8f) ------------------------------------------------
 * When the logits of a CrossEntropyLoss are the output of an nn.Linear (the GPT example's fc_out: 4096 x 15000 fp32 =
 * 246 MB), the loss backward does not materialise dlogits in fp32: nnb_cross_entropy_backward_staged writes them
 * straight as the bf16 operand planes (nnb_weight_staged_bytes(rows, C, prec) bytes) plus the bias gradient db (column
 * sums, deterministic two-pass), and nnb_linear_backward_staged runs dgrad / wgrad on those planes -- no fp32 dlogits
 * write + read and no staging pass (the reference's native kernel overwrites the logits in place instead,
 * experimental/losses/cross_entropy_loss/cross_entropy.cu:195-212). Labels as in nnb_cross_entropy_forward. */
size_t nnb_cross_entropy_staged_workspace_bytes(int64_t rows, int64_t C);
int nnb_cross_entropy_backward_staged(const float* logits, const int32_t* targets, const float* lse,
                                      const float* inv_denom, const float* upstream, int64_t rows, int64_t C,
                                      int64_t ignore_index, void* dZ_staged_out, int prec, float* db, void* workspace,
                                      size_t workspace_bytes, cudaStream_t stream);
int nnb_linear_backward_staged(const float* X, const float* W, const void* dO_staged, float* dX, float* dW,
                               int64_t M, int64_t K, int64_t N, int prec, const void* W_staged, const void* X_staged,
                               void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* ---- gradient all-reduce over NCCL (the reference has no distributed code) -----
 * The ONE collective of the data-parallel step: an in-place sum all-reduce of the flat fp32 gradient bucket over
 * NVLink / NVSwitch. NCCL is resolved at run time from the libnccl.so.2 already in the process (or on the loader path):
 * the library does not link it, and without NCCL these return NNB_ERR_UNSUPPORTED. One communicator per rank / GPU
 * (the device current at nnb_comm_init). Rank 0 calls nnb_comm_unique_id and hands the 128 bytes to the other ranks
 * out of band (file, MPI, torch.distributed store ...). The averaging 1 / world is not a pass of its own: it is the
 * grad_scale of nnb_adamw_step. nnb_comm_allreduce_sum is asynchronous on `stream` and may be captured in a CUDA graph. */
typedef struct nnb_comm nnb_comm;
int nnb_comm_available(int* nccl_version);
int nnb_comm_unique_id(void* id_out_128_bytes);
int nnb_comm_init(nnb_comm** comm, const void* unique_id_128_bytes, int world, int rank);
int nnb_comm_allreduce_sum(nnb_comm* comm, float* buf, int64_t n_elems, cudaStream_t stream);
int nnb_comm_destroy(nnb_comm* comm);

/* ---- native ConvTranspose2d (row N1 of SURVEY.md 8f) -----------------------------------------------
 * neunet/nn/layers/convtranspose2d.py:115-387: weights (out, in, kh, kw), NOT flipped; the layer is a stride-1
 * correlation over the zero-stuffed, (k-1)-padded, padding-cropped input (:165-181, 321), so the reference multiplies
 * s0*s1 - 1 zeros for every real tap. Here: gather form with real taps only -- the output is split into s0*s1 parity
 * classes, each an implicit GEMM over its own tap subset read from the channels-last planes of X; dX is ONE strided
 * implicit GEMM over dO, dW one implicit wgrad GEMM. desc: B, Cin, H, W describe the INPUT, Cout / kh / kw / stride /
 * pad / dil the layer. Supported natively for Cin % 64 == Cout % 64 == 0, dilation 1, output_padding 0, output size
 * divisible by the stride (nnb_conv_transpose2d_supported); otherwise NNB_ERR_UNSUPPORTED and the caller uses the
 * zero-stuffed formulation over nnb_conv2d_* (same results). X_planes_out / X_planes: channels-last bf16 planes of
 * X kept from forward for wgrad (nnb_conv_transpose2d_planes_bytes). */
int nnb_conv_transpose2d_supported(const nnb_conv2d_desc* desc, int out_pad0, int out_pad1);
int nnb_conv_transpose2d_out_shape(const nnb_conv2d_desc* desc, int out_pad0, int out_pad1, int64_t* Ho, int64_t* Wo);
size_t nnb_conv_transpose2d_planes_bytes(const nnb_conv2d_desc* desc, int prec);
size_t nnb_conv_transpose2d_workspace_bytes(const nnb_conv2d_desc* desc, int out_pad0, int out_pad1, int prec,
                                            int backward);
int nnb_conv_transpose2d_forward(const nnb_conv2d_desc* desc, int out_pad0, int out_pad1, const float* X,
                                 const float* W, const float* bias, float* O, int prec, void* X_planes_out,
                                 void* workspace, size_t workspace_bytes, cudaStream_t stream);
int nnb_conv_transpose2d_backward(const nnb_conv2d_desc* desc, int out_pad0, int out_pad1, const float* X,
                                  const float* W, const float* dO, float* dX, float* dW, float* db, int prec,
                                  const void* X_planes, void* workspace, size_t workspace_bytes,
                                  cudaStream_t stream);

/* ---- Fused LeakyReLU + BatchNorm2d over NCHW fp32 (row N2 of SURVEY.md 8f) -----------------------
 * a = leaky_relu(x, alpha) (neunet/nn/activations.py:73-93; alpha = 1 gives plain BatchNorm2d),
 * y = (a - mean_c) * inv_std_c * w_c + b_c with batch statistics over (B, H, W), ddof 0
 * (neunet/nn/layers/batchnorm2d.py:57-115; backward :11-55). The reference runs these as ~12 array passes per
 * conv -> LeakyReLU -> BatchNorm2d group of the DDPM ResBlock (examples/ddpm.ipynb cell 5 l.41-55).
 *   forward : nnb_bn_stats -> sums[2C] fp64 {sum a, sum a^2}; (SyncBN: the caller all-reduces sums and passes the
 *             global count) -> nnb_bn_finalize: mean, inv_std = 1/sqrt(var + eps), and, when non-NULL,
 *             running = momentum * running + (1 - momentum) * stat (the reference's convention, :87-88)
 *             -> nnb_bn_apply. w / b may be NULL (affine = False).
 *   backward: nnb_bn_backward_stats -> sums[2C] {sum g, sum g * xhat} -> nnb_bn_backward_apply:
 *             dx = lrelu'(x) * inv_std * (w g - mean(w g) - xhat * mean(w g xhat)), dw = sum g xhat, db = sum g
 *             (dx / dw / db nullable, so the parameter gradients can be taken from LOCAL sums before a SyncBN
 *             all-reduce of sums). Only x, mean, inv_std are kept from forward.
 * workspace: nnb_bn_workspace_bytes(B, C) bytes for the per-channel partial sums (deterministic order). */
size_t nnb_bn_workspace_bytes(int64_t B, int64_t C);
int nnb_bn_stats(const float* x, int64_t B, int64_t C, int64_t HW, float alpha, double* sums, void* workspace,
                 size_t workspace_bytes, cudaStream_t stream);
int nnb_bn_finalize(const double* sums, double count, int64_t C, float eps, float momentum, float* mean,
                    float* inv_std, float* running_mean, float* running_var, cudaStream_t stream);
int nnb_bn_apply(const float* x, const float* mean, const float* inv_std, const float* w, const float* b,
                 int64_t B, int64_t C, int64_t HW, float alpha, float* y, cudaStream_t stream);
int nnb_bn_backward_stats(const float* x, const float* grad, const float* mean, const float* inv_std, int64_t B,
                          int64_t C, int64_t HW, float alpha, double* sums, void* workspace, size_t workspace_bytes,
                          cudaStream_t stream);
int nnb_bn_backward_apply(const float* x, const float* grad, const float* mean, const float* inv_std, const float* w,
                          const double* sums, double count, int64_t B, int64_t C, int64_t HW, float alpha,
                          float* dx, float* dw, float* db, cudaStream_t stream);

/* ---- nn.Embedding (row N4 of SURVEY.md 8f) ----------------------------------------------------------
 * neunet/nn/layers/embedding.py:61-75: forward = weight[ids]; backward through Tensor.__getitem__ ASSIGNS
 * (neunet/autograd.py:909-910), so with duplicate ids the LAST occurrence wins and the other rows' gradients are dropped.
 * nnb_embedding_backward reproduces that deterministically (atomicMax of the position per row, then the winner copies).
 * ids: int32 or int64 device array of n entries; negative ids wrap like NumPy; ids outside [-V, V) (IndexError in the
 * reference) yield NaN rows forward and are ignored backward. dW [V, D] is fully overwritten (zeros for untouched rows).
 * workspace: nnb_embedding_workspace_bytes(V). */
int nnb_embedding_forward(const float* W, const void* ids, int ids_are_int64, int64_t n, int64_t V, int64_t D, float* out,
                          cudaStream_t stream);
size_t nnb_embedding_workspace_bytes(int64_t V);
int nnb_embedding_backward(const void* ids, int ids_are_int64, const float* grad, int64_t n, int64_t V, int64_t D, float* dW,
                           void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* ---- Fused attention for short sequences (row N4 of SURVEY.md 8f) -------------------------------
 * examples/gpt.ipynb cell 2 l.25-40 as ONE kernel per direction:
 *   scores = q . kT / scale; scores = where(mask, fill, scores); p = softmax(scores, -1);
 *   attn = dropout(p); out = attn . v
 * replacing two Tensor.matmul (neunet/autograd.py:192-230), a division, where (autograd.py:658-684),
 * nn.Softmax (activations.py:437-459) and nn.Dropout (layers/dropout.py:17-46). The per-head products (<= 64 x 64 x 64)
 * run on the tensor cores as warp-level TF32 MMAs with fp32 accumulation, straight from fp32 shared-memory tiles:
 * prec = NNB_PREC_BF16 -> one TF32 product (~5e-4 rel. per contraction), NNB_PREC_BF16X3 -> hi/lo split, three
 * products (~1e-6 rel., fp32-grade; parity mode). Softmax, masking, dropout and all sums are fp32.
 * q: logical (B,H,Tq,D), kT: logical (B,H,D,Tk), v: logical (B,H,Tk,D), dO: logical (B,H,Tq,D), each with four
 * element strides (any layout). mask (nullable) is described by mask_kind: 1 = float tensor, masked where != 0;
 * 2 = int32 tensor, masked where == mask_cmp; 3 = float tensor, masked where == mask_cmp; strides over
 * (B,H,Tq,Tk), 0 on broadcast axes. attn: (B,H,Tq,Tk) contiguous (post-dropout, what the example returns).
 * out / dQ: (B,Tq,H,D) contiguous; dK / dV: (B,Tk,H,D) contiguous -- the memory order of the example's
 * reshape/transpose views, so no copy is needed either side. out_row_pitch (backward; 0 = H*D) is the distance in floats
 * between consecutive (b, t) rows of dQ / dK / dV: 3*H*D lets the three be the column blocks dq | dk | dv of ONE
 * [B*T, 3*H*D] matrix, which is the upstream gradient of a fused q/k/v projection as it stands. out_staged (nullable): bf16 planes [B*Tq][H*D].
 * Dropout: p in [0,1), mask = Philox(seed, call_id, epoch | *epoch_dev) over the flat attn index, regenerated in
 * backward (nothing O(T^2) is saved). Limits: Tq, Tk, D <= 64, Tk % 4 == D % 4 == 0 (nnb_attention_supported),
 * otherwise NNB_ERR_UNSUPPORTED and the caller runs the un-fused ops. */
int nnb_attention_supported(int64_t Tq, int64_t Tk, int64_t D);
int nnb_attention_forward(const float* Q, const int64_t q_strides[4], const float* KT, const int64_t kt_strides[4],
                          const float* V, const int64_t v_strides[4], const void* mask, int mask_kind, float mask_cmp,
                          const int64_t mask_strides[4], float fill, float scale, float p, uint64_t seed,
                          uint32_t call_id, uint64_t epoch, const uint64_t* epoch_dev, float* attn, float* out,
                          void* out_staged, int prec, int64_t B, int64_t H, int64_t Tq, int64_t Tk, int64_t D,
                          cudaStream_t stream);
int nnb_attention_backward(const float* Q, const int64_t q_strides[4], const float* KT, const int64_t kt_strides[4],
                           const float* V, const int64_t v_strides[4], const void* mask, int mask_kind, float mask_cmp,
                           const int64_t mask_strides[4], float fill, float scale, float p, uint64_t seed,
                           uint32_t call_id, uint64_t epoch, const uint64_t* epoch_dev, const float* dO,
                           const int64_t do_strides[4], float* dQ, float* dK, float* dV, int64_t out_row_pitch,
                           int prec, int64_t B, int64_t H, int64_t Tq, int64_t Tk, int64_t D, cudaStream_t stream);

/* ---- Dropout with a device RNG (row N4 of SURVEY.md 8f) ----------------------------------------
 * neunet/nn/layers/dropout.py:17-46: y = x * mask, mask ~ Bernoulli(1-p) / (1-p); backward is the
 * same call on the upstream gradient. The mask is not materialised: it is Philox4x32-10 of
 * (seed, call_id, epoch, element index), so the same (seed, call_id, epoch) regenerates it.
 * epoch_dev, when non-NULL, is a device-resident uint64 that overrides `epoch` (CUDA-graph replays:
 * the graph bakes the pointer; nnb_rng_advance increments the value between replays).
 */
int nnb_dropout(const float* x, float* y, int64_t n, float p, uint64_t seed, uint32_t call_id,
                uint64_t epoch, const uint64_t* epoch_dev, cudaStream_t stream);
int nnb_rng_advance(uint64_t* epoch_dev, cudaStream_t stream);
/* y = (residual +) dropout(x) over a [rows, cols] matrix -- the example's `x = x + self.dropout(a)`
 * (examples/gpt.ipynb cell 5 l.12-13) in one pass -- and, when Y_staged_out is non-NULL, also the bf16
 * planes of y (nnb_weight_staged_bytes(rows, cols, prec) bytes, 256-byte aligned, cols % 8 == 0) that a
 * following nnb_linear_forward_staged consumes, so the consumer needs no staging pass.
 * residual may be NULL. Backward of the dropout branch is nnb_dropout on the upstream gradient. */
int nnb_dropout_fused(const float* x, const float* residual, float* y, int64_t rows, int64_t cols, float p,
                      uint64_t seed, uint32_t call_id, uint64_t epoch, const uint64_t* epoch_dev,
                      void* Y_staged_out, int prec, cudaStream_t stream);
/* y = dropout(swish(z, beta)) over a [rows, cols] matrix, same mask / planes contract as nnb_dropout_fused: the
 * feed-forward block of examples/gpt.ipynb cell 4 l.9-13 (fc_2(dropout(swish(fc_1(x))))) = nn.Swish
 * (neunet/nn/activations.py:208-233) + nn.Dropout (layers/dropout.py:17-46) in one pass over the pre-activation z that
 * the fc_1 GEMM wrote; the Swish output itself is never materialised (backward needs z and the Philox ticket only). */
int nnb_swish_dropout_fused(const float* z, float beta, float* y, int64_t rows, int64_t cols, float p, uint64_t seed,
                            uint32_t call_id, uint64_t epoch, const uint64_t* epoch_dev, void* Y_staged_out, int prec,
                            cudaStream_t stream);

/* ---- fused CrossEntropyLoss (row N3 of SURVEY.md 8f, first half) --------------------------------
 * LogSoftmax(axis=1) + NLLLoss with unit class weights (neunet/nn/losses.py:59-126); native analogue
 * in the reference: cudaCrossEntropyForwardBackward (experimental/losses/cross_entropy_loss/
 * cross_entropy.cu:249-292, which overwrites the logits in place; here dlogits is a separate buffer
 * and the upstream gradient is applied in the same pass).
 * logits[rows, C] fp32, targets[rows] int32; rows whose target == ignore_index contribute 0.
 * reduction: 0 none (row_loss is the result), 1 mean over kept rows, 2 sum.
 * forward writes row_loss[rows], lse[rows] (log-sum-exp, kept for backward) and, for mean/sum, the
 * scalar loss_out and inv_denom (1/kept-rows or 1). backward: dlogits = (softmax - onehot) * keep *
 * inv_denom * upstream, upstream being one device scalar or (upstream_per_row) a [rows] vector.
 */
int nnb_cross_entropy_forward(const float* logits, const int32_t* targets, int64_t rows, int64_t C,
                              int64_t ignore_index, int reduction, float* row_loss, float* lse,
                              float* loss_out, float* inv_denom, cudaStream_t stream);
int nnb_cross_entropy_backward(const float* logits, const int32_t* targets, const float* lse,
                               const float* inv_denom, const float* upstream, int upstream_per_row,
                               int64_t rows, int64_t C, int64_t ignore_index, float* dlogits,
                               cudaStream_t stream);

/* ---- multi-tensor Adam / AdamW ---------------------------------------------------------------
 * Replaces the per-tensor Python loops of neunet/optim.py:17-33 (Adam) and 52-69 (AdamW) and the
 * reference's CreateFusedOptimizer / FusedAdamWStep / DestroyFusedOptimizer
 * (experimental/optim/fused_adamw/fused_adamw_multitensor.cu:306-340).
 * nnb_adamw_create uploads the pointer table ONCE (the reference re-uploads it every step,
 * fused_adamw_multitensor.cu:288-291); p/g/m/v are arrays of n device pointers, sizes in elements.
 * Tensors whose g[i] is NULL are skipped, like `if param.grad is None: continue` (optim.py:21-22).
 * step: 1-based step count t (or 0, see nnb_adamw_set_step). Hyper-parameters are doubles: 1-beta, 1-beta^t and lr*wd are formed in
 * double on the host and rounded once to fp32, exactly where the reference forms them as Python
 * floats before NumPy casts them to the array dtype. grad_scale multiplies every gradient first (1/world_size after
 * a sum all-reduce; 1.0 otherwise).
 */
typedef struct nnb_adamw nnb_adamw;
int nnb_adamw_create(nnb_adamw** out, int n, float* const* p, const float* const* g,
                     float* const* m, float* const* v, const int64_t* sizes, cudaStream_t stream);
int nnb_adamw_set_grads(nnb_adamw* opt, const float* const* g, cudaStream_t stream);
int nnb_adamw_step(nnb_adamw* opt, double lr, double beta1, double beta2, double eps,
                   double weight_decay, int64_t step, int mode, float grad_scale,
                   cudaStream_t stream);
/* The same update for tensors [first_tensor, first_tensor + n_tensors) only: data-parallel training steps each gradient
 * chunk as soon as its all-reduce has landed, so the optimizer of the early chunks runs underneath the collective of
 * the last ones (neunet/distributed.py: GradBucket.all_reduce_and_step). One optimizer step = ranges that cover every
 * tensor once, all with the same `step`; with step = 0 exactly one of them passes advance_counter = 1 (the first). */
int nnb_adamw_step_range(nnb_adamw* opt, int first_tensor, int n_tensors, double lr, double beta1, double beta2, double eps,
                         double weight_decay, int64_t step, int mode, float grad_scale, int advance_counter,
                         cudaStream_t stream);
/* CUDA-graph support: pass step = 0 to nnb_adamw_step to use (and first advance) a step counter
 * kept in device memory, so replaying a captured step keeps the bias corrections moving;
 * nnb_adamw_set_step() seeds that counter (number of steps already taken). */
int nnb_adamw_set_step(nnb_adamw* opt, int64_t step, cudaStream_t stream);
/* Fused weight staging: for every tensor i with staged[i] != NULL the step kernel also writes the bf16
 * planes of the UPDATED weight, viewed as [size_i / cols[i], cols[i]], into staged[i] (a buffer of
 * nnb_weight_staged_bytes(rows, cols, prec) bytes, the layout nnb_stage_weight() produces), so the next
 * nnb_linear_forward / backward can take it as W_staged without a conversion pass. staged = NULL turns
 * the feature off. Not callable during stream capture (it uploads tables synchronously). */
int nnb_adamw_set_staging(nnb_adamw* opt, void* const* staged, const int64_t* cols, int prec,
                          cudaStream_t stream);
int nnb_adamw_destroy(nnb_adamw* opt);

#ifdef __cplusplus
}
#endif
#endif /* NEUNET_B200_H */
