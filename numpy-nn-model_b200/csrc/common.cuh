// Shared host/device declarations for libneunet_b200 (internal; the public C-ABI is include/neunet_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "../../include/neunet_b200.h"

namespace nnb {

// ---- error plumbing -------------------------------------------------------------------------
// The reference's native modules printf+exit() on failure (linear_cublaslt_no_manual_mem.cu:91-94);
// here every entry point returns an int status and keeps the message for nnb_last_error().
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define NNB_CUDA_OK(expr)                                                                  \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            return ::nnb::fail(NNB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                               cudaGetErrorString(_e), __FILE__, __LINE__);                \
    } while (0)

#define NNB_REQUIRE(cond, ...)                                         \
    do {                                                               \
        if (!(cond)) return ::nnb::fail(NNB_ERR_INVALID, __VA_ARGS__); \
    } while (0)

// ---- NVTX ranges around the C-ABI entry points (the reference wraps its optimizer step in
// nvtx ranges for profiling, scripts/profile_adam.py:22-43). Header-only NVTX v3: without a profiler attached a range is
// one relaxed load and a branch. NNB_NVTX=0 in the environment removes even that.
bool nvtx_enabled();
void nvtx_push(const char* name);
void nvtx_pop();
struct NvtxRange {
    bool on;
    explicit NvtxRange(const char* name) : on(nvtx_enabled()) { if (on) nvtx_push(name); }
    ~NvtxRange() { if (on) nvtx_pop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
#define NNB_RANGE(name) ::nnb::NvtxRange _nnb_range_(name)

// Kernel-launch accounting (bench.py reports `gpu_launches`).
void count_launch(int n = 1);

int num_sms();            // SMs the library sizes its grids for (device count, or the budget below)
int set_sm_budget(int n);  // 0 = all SMs

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------
// A training step is a chain of several hundred short dependent kernels; with the
// programmatic-stream-serialization launch attribute the next kernel's CTAs are scheduled while the current
// kernel drains (it calls pdl_trigger() at its start), run their prologue, and block in pdl_wait() until the
// predecessor has completed and flushed. Kernels launched this way must not touch global memory before
// pdl_wait(). nnb_set_pdl(0) / NNB_PDL=0 turns the attribute off (the device calls are then no-ops).
bool pdl_enabled();
int set_pdl(int on);

#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ---- staged bf16 operands ---------------------------------------------------------------------
// The fp32 tensors the reference API hands us are converted once per use ("staged") into bf16
// planes that TMA can tile: `hi` = bf16(x) and, in BF16X3 mode, `lo` = bf16(x - hi).
// A staged matrix is row-major [batch][rows][ld] with ld = round_up(cols, 8) so the row pitch is a
// multiple of 16 bytes (TMA requirement), and every plane starts 256-byte aligned.
struct Staged {
    const __nv_bfloat16* hi = nullptr;
    const __nv_bfloat16* lo = nullptr;  // null in BF16 mode
    int64_t rows = 0, cols = 0, ld = 0;
    int64_t batch = 1;
    int64_t batch_stride = 0;  // elements
};

static inline int64_t staged_ld(int64_t cols) { return round_up(cols, 8); }
static inline size_t staged_plane_bytes(int64_t batch, int64_t rows, int64_t cols) {
    return (size_t)round_up(batch * rows * staged_ld(cols) * 2, 256);
}

// 4-D strided fp32 view -> staged bf16 planes (optionally transposing the last two dims).
// view dims: [b0][b1][rows][cols] with element strides s_b0, s_b1, s_r, s_c.
struct View4 {
    const float* ptr;
    int64_t b0, b1, rows, cols;
    int64_t s_b0, s_b1, s_r, s_c;
};
static inline View4 view2d(const float* p, int64_t rows, int64_t cols, int64_t ld) {
    return View4{p, 1, 1, rows, cols, 0, 0, ld, 1};
}

// Elementwise transform fused into staging.
enum StageOp {
    STAGE_COPY = 0,
    STAGE_SWISH_BWD = 1,  // out = src * swish'(aux)  (aux = saved pre-activation z, same shape/strides)
};

size_t stage_colsum_scratch_bytes(int64_t cols);

// dst planes must hold staged_plane_bytes() each. If `colsum` is non-null it receives
// sum over (batch, rows) of the *transformed fp32* values per column (bias gradient), length cols.
// `drop` (nullable): the source is first multiplied by the dropout mask of that Philox ticket (x * keep / (1 - p), flat
// index = row * cols + col, so the view must be one contiguous [rows, cols] matrix with cols % 4 == 0): the backward of
// nn.Dropout folded into the staging of the upstream gradient of the nn.Linear below it.
struct DropArgs;
int stage_operand(const View4& src, bool transpose, int prec, __nv_bfloat16* dst_hi,
                  __nv_bfloat16* dst_lo, int op, const float* aux, float beta, float* colsum,
                  float* colsum_scratch, cudaStream_t stream, Staged* out, const DropArgs* drop = nullptr);

// ---- GEMM ---------------------------------------------------------------------------------------
// D[b][m][n] = epilogue( alpha * sum_k A[b][m][k] * B[b][n][k] )
// A "K-major" operand is staged as [rows = M or N][cols = K]; an "MN-major" operand is staged as
// [rows = K][cols = M or N] (i.e. the untransposed matrix when the reduction runs over its rows).
struct GemmOperand {
    Staged st;
    bool mn_major = false;
};

struct GemmEpilogue {
    float alpha = 1.0f;
    const float* bias = nullptr;  // [N], added per output column
    float* Z = nullptr;           // optional side store of the pre-activation (same layout as D)
    int act = NNB_ACT_NONE;
    float beta = 1.0f;
    __nv_bfloat16* D16 = nullptr;  // optional bf16 shadow of D (ld = ldd16)
    int64_t ldd16 = 0;
};

// Implicit-GEMM convolution: the B operand is not a staged matrix but bf16 channels-last planes
// src[b][y][x][c] (C % 64 == 0) read through 4-D TMA boxes; a GEMM "position" is m = (b * P + p) * Q + q and
// position (p, q) with tap (k, l) reads pixel (p * sp0 - off0 + k * dk0, q * sp1 - off1 + l * dk1), zero outside.
//   mode 1 (forward / stride-1 dgrad): B[n = position][k = (tap, c)]   (K-major;  GEMM N = positions, K = taps * C)
//   mode 2 (wgrad)                   : B[k = position][n = (tap, c)]   (MN-major; GEMM K = positions, N = taps * C)
struct ConvOperand {
    int mode = 0;
    const __nv_bfloat16* hi = nullptr;
    const __nv_bfloat16* lo = nullptr;
    int64_t B = 0, C = 0, Hs = 0, Ws = 0;  // source planes
    int64_t P = 0, Q = 0;                  // position grid per image
    int kh = 1, kw = 1;
    int sp0 = 1, sp1 = 1, off0 = 0, off1 = 0, dk0 = 1, dk1 = 1;
};

struct GemmProblem {
    int64_t M = 0, N = 0, K = 0, batch = 1;
    GemmOperand A, B;
    ConvOperand conv;  // conv.mode != 0: B comes from channels-last planes (B.st is ignored)
    float* D = nullptr;
    int64_t ldd = 0, batch_stride_d = 0;
    bool reduce_batch = false;  // sum over the batch dimension into ONE output (wgrad-style)
    // optional output column map: addr = (col / col_group) * group_stride + row * ldd + col % col_group
    // (lets a [Cout, B*Ho*Wo] GEMM write NCHW directly); 0 = plain row-major
    int64_t col_group = 0, group_stride = 0;
    bool bias_per_row = false;
    GemmEpilogue epi;
    // split-K scratch (fp32); required when the heuristic picks splits > 1
    float* splitk_ws = nullptr;
    size_t splitk_ws_bytes = 0;
    int force_bn = 0;      // 0 = heuristic
    int force_splits = 0;  // 0 = heuristic
    int force_cg = 0;      // 0 = heuristic, 1 / 2 = CTAs per tile (tcgen05 cta_group)
    unsigned long long* clk_out = nullptr;  // optional device buffer [2]: {SM cycles, ns} of CTA 0 (clock probe)
};

size_t gemm_splitk_ws_bytes(int64_t M, int64_t N, int64_t K, int64_t batch);
int gemm(const GemmProblem& p, cudaStream_t stream);
// true when conv-mode boxes exist for this position grid (power-of-two style shapes; see gemm.cu: conv_rect)
bool gemm_conv_supported(int mode, int64_t C, int64_t P, int64_t Q, int sp0, int sp1);

}  // namespace nnb
