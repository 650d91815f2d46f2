// Tensor.matmul forward / backward (neunet/autograd.py:192-230) for batched, strided operands.
// GPT attention calls it on transposed views: q.k^T with (B,h,T,d) x (B,h,d,T) and attn.v
// (examples/gpt.ipynb cell 2 l.25-36). A view whose last dim is not contiguous but whose
// second-to-last is gets staged in its memory order and handed to the GEMM with the other operand
// major, so no transposing copy is ever made.
#include "common.cuh"
#include "workspace.cuh"

namespace nnb {
namespace {

int planes(int prec) { return prec == NNB_PREC_BF16X3 ? 2 : 1; }

struct Op {
    GemmOperand g;
};

// Stage a logical [b0,b1,R,C] view. `want_rows_are_reduction` tells how the GEMM will use it:
// the GEMM needs to know whether the staged ROWS or the staged COLS run along the reduction.
// Returns the operand with mn_major set accordingly.
//   reduction_is_cols = true : logical matrix is [out_dim, reduction]
//   reduction_is_cols = false: logical matrix is [reduction, out_dim]
// `keep`: optional caller-owned buffer of nnb_matmul_staged_bytes(); the planes are written there
// instead of the workspace (forward, so backward can reuse them) or, with `have` set, are taken from
// it as already converted (backward) and nothing is launched.
int stage_view(Bump& ws, const float* ptr, const int64_t st[4], int64_t b0, int64_t b1, int64_t R,
               int64_t C, bool reduction_is_cols, int prec, cudaStream_t stream, GemmOperand* out,
               void* keep = nullptr, bool have = false) {
    View4 v;
    v.ptr = ptr;
    const bool bcast = (st[0] == 0 || b0 == 1) && (st[1] == 0 || b1 == 1);
    v.b0 = bcast ? 1 : b0;
    v.b1 = bcast ? 1 : b1;
    v.s_b0 = st[0];
    v.s_b1 = st[1];
    bool transposed = false;
    if (st[3] == 1 || C == 1 || st[2] != 1) {
        v.rows = R; v.cols = C; v.s_r = st[2]; v.s_c = st[3];
    } else {  // st[2] == 1: memory order is [C][R]; stage that and flip the major
        v.rows = C; v.cols = R; v.s_r = st[3]; v.s_c = 1;
        transposed = true;
    }
    const int64_t batch = v.b0 * v.b1;
    const size_t pb = staged_plane_bytes(batch, v.rows, v.cols);
    __nv_bfloat16 *hi, *lo = nullptr;
    if (keep != nullptr) {
        NNB_REQUIRE((reinterpret_cast<uintptr_t>(keep) & 255) == 0, "matmul: staged buffer must be 256-byte aligned");
        hi = static_cast<__nv_bfloat16*>(keep);
        if (prec == NNB_PREC_BF16X3) lo = reinterpret_cast<__nv_bfloat16*>(static_cast<char*>(keep) + pb);
    } else {
        hi = static_cast<__nv_bfloat16*>(ws.take(pb));
        if (prec == NNB_PREC_BF16X3) lo = static_cast<__nv_bfloat16*>(ws.take(pb));
        if (!ws.ok()) return fail(NNB_ERR_WORKSPACE, "matmul: workspace too small (need >= %zu bytes)", ws.off);
    }
    if (keep != nullptr && have) {
        Staged& s = out->st;
        s.hi = hi; s.lo = lo; s.rows = v.rows; s.cols = v.cols; s.ld = staged_ld(v.cols);
        s.batch = batch; s.batch_stride = v.rows * s.ld;
    } else {
        int rc = stage_operand(v, false, prec, hi, lo, STAGE_COPY, nullptr, 0.f, nullptr, nullptr, stream, &out->st);
        if (rc) return rc;
    }
    // staged rows = R (cols = C) unless transposed. K-major <=> staged cols are the reduction.
    const bool staged_cols_are_reduction = transposed ? !reduction_is_cols : reduction_is_cols;
    out->mn_major = !staged_cols_are_reduction;
    return NNB_OK;
}

const int64_t* contig_strides(int64_t b1, int64_t R, int64_t C, int64_t out[4]) {
    out[3] = 1; out[2] = C; out[1] = R * C; out[0] = b1 * R * C;
    return out;
}

}  // namespace
}  // namespace nnb

using namespace nnb;

extern "C" {

size_t nnb_matmul_workspace_bytes(int64_t b0, int64_t b1, int64_t M, int64_t K, int64_t N,
                                  int prec, int backward) {
    if (b0 <= 0 || b1 <= 0 || M <= 0 || K <= 0 || N <= 0) return 0;
    const int64_t b = b0 * b1;
    const size_t p = planes(prec);
    // staged_plane_bytes depends on which dim is padded; take the larger orientation
    auto pl = [&](int64_t r, int64_t c) { return std::max(staged_plane_bytes(b, r, c), staged_plane_bytes(b, c, r)); };
    size_t bytes = 4096 + p * (pl(M, K) + pl(K, N));
    if (backward) {
        bytes += p * pl(M, N);
        bytes += std::max(gemm_splitk_ws_bytes(M, K, N, b), gemm_splitk_ws_bytes(K, N, M, b));
    } else {
        bytes += gemm_splitk_ws_bytes(M, N, K, b);
    }
    return bytes;
}

size_t nnb_matmul_staged_bytes(int64_t b0, int64_t b1, int64_t rows, int64_t cols, int prec) {
    if (b0 <= 0 || b1 <= 0 || rows <= 0 || cols <= 0) return 0;
    const int64_t b = b0 * b1;
    return (size_t)planes(prec) * std::max(staged_plane_bytes(b, rows, cols), staged_plane_bytes(b, cols, rows));
}

int nnb_matmul_forward(const float* A, const int64_t a_strides[4], const float* B,
                       const int64_t b_strides[4], float* C, int64_t b0, int64_t b1, int64_t M,
                       int64_t K, int64_t N, float alpha, int prec, void* A_staged_out,
                       void* B_staged_out, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream) {
    NNB_REQUIRE(A && B && C && a_strides && b_strides, "nnb_matmul_forward: null pointer");
    NNB_REQUIRE(b0 > 0 && b1 > 0 && M > 0 && K > 0 && N > 0, "nnb_matmul_forward: non-positive dimension");
    NNB_REQUIRE(prec == NNB_PREC_BF16 || prec == NNB_PREC_BF16X3, "nnb_matmul_forward: bad prec");
    Bump ws(workspace, workspace_bytes);
    GemmProblem g;
    g.M = M; g.N = N; g.K = K; g.batch = b0 * b1;
    int rc = stage_view(ws, A, a_strides, b0, b1, M, K, /*reduction_is_cols=*/true, prec, stream, &g.A, A_staged_out);
    if (rc) return rc;
    rc = stage_view(ws, B, b_strides, b0, b1, K, N, /*reduction_is_cols=*/false, prec, stream, &g.B, B_staged_out);
    if (rc) return rc;
    g.D = C; g.ldd = N; g.batch_stride_d = M * N;
    g.epi.alpha = alpha;
    g.splitk_ws_bytes = ws.remaining();
    g.splitk_ws = static_cast<float*>(ws.take(g.splitk_ws_bytes));
    return gemm(g, stream);
}

int nnb_matmul_backward(const float* A, const int64_t a_strides[4], const float* B,
                        const int64_t b_strides[4], const float* G, float* dA, float* dB,
                        int64_t b0, int64_t b1, int64_t M, int64_t K, int64_t N, float alpha,
                        int prec, const void* A_staged, const void* B_staged, void* workspace,
                        size_t workspace_bytes, cudaStream_t stream) {
    NNB_REQUIRE(A && B && G && a_strides && b_strides, "nnb_matmul_backward: null pointer");
    NNB_REQUIRE(b0 > 0 && b1 > 0 && M > 0 && K > 0 && N > 0, "nnb_matmul_backward: non-positive dimension");
    NNB_REQUIRE(prec == NNB_PREC_BF16 || prec == NNB_PREC_BF16X3, "nnb_matmul_backward: bad prec");
    Bump ws(workspace, workspace_bytes);
    const int64_t batch = b0 * b1;
    int64_t gst[4];
    contig_strides(b1, M, N, gst);
    GemmOperand g_red_cols, g_red_rows;
    // G is used as [M, N(reduction)] for dA and as [M(reduction), N] for dB: same staged planes.
    int rc = stage_view(ws, G, gst, b0, b1, M, N, true, prec, stream, &g_red_cols);
    if (rc) return rc;
    g_red_rows = g_red_cols;
    g_red_rows.mn_major = true;
    GemmOperand a_op, b_op;
    if (dA) {  // dA[M,K] = alpha * G[M,N] . B[K,N]^T   (autograd.py:209)
        rc = stage_view(ws, B, b_strides, b0, b1, K, N, /*reduction_is_cols=*/true, prec, stream, &b_op,
                        const_cast<void*>(B_staged), B_staged != nullptr);
        if (rc) return rc;
    }
    if (dB) {  // dB[K,N] = alpha * A[M,K]^T . G[M,N]   (autograd.py:211)
        rc = stage_view(ws, A, a_strides, b0, b1, M, K, /*reduction_is_cols=*/false, prec, stream, &a_op,
                        const_cast<void*>(A_staged), A_staged != nullptr);
        if (rc) return rc;
    }
    const size_t sk_bytes = ws.remaining();
    float* sk = static_cast<float*>(ws.take(sk_bytes));
    if (dA) {
        GemmProblem g;
        g.M = M; g.N = K; g.K = N; g.batch = batch;
        g.A = g_red_cols; g.B = b_op;
        g.D = dA; g.ldd = K; g.batch_stride_d = M * K;
        g.epi.alpha = alpha;
        g.splitk_ws = sk; g.splitk_ws_bytes = sk_bytes;
        rc = gemm(g, stream);
        if (rc) return rc;
    }
    if (dB) {
        GemmProblem g;
        g.M = K; g.N = N; g.K = M; g.batch = batch;
        g.A = a_op; g.B = g_red_rows;
        g.D = dB; g.ldd = N; g.batch_stride_d = K * N;
        g.epi.alpha = alpha;
        g.splitk_ws = sk; g.splitk_ws_bytes = sk_bytes;
        rc = gemm(g, stream);
        if (rc) return rc;
    }
    return NNB_OK;
}

}  // extern "C"
