// Tensor.matmul forward / backward (neunet/autograd.py:192-230) for batched, strided operands.
// GPT attention calls it on transposed views: q.k^T with (B,h,T,d) x (B,h,d,T) and attn.v
// (examples/gpt.ipynb cell 2 l.25-36). A view whose last dim is not contiguous but whose
// second-to-last is gets staged in its memory order and handed to the GEMM with the other operand
// major, so no transposing copy is ever made.
#include <cstdlib>

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

// ---- small batched products on the CUDA cores (fp32 FMA) --------------------------------------------
// OPT-IN alternative for GPT attention at the notebook's sizes (512 independent 64 x 64 x 64 products per
// call, gpt cell 2 l.25-36), where a 128-row tensor tile is half padding and every operand needs its own
// bf16 staging launch: one CTA computes one 64 x 64 output tile of one batch element straight from the
// strided fp32 views (either operand may be a transposed view), so a call is ONE launch and the result is
// fp32-exact. C[b] = alpha * A[b] . B[b], A: [M,K] strides (ar, ac), B: [K,N] strides (br, bc).
struct SmallBmm {
    const float* A; const float* B; float* C;
    long long a_s0, a_s1, a_r, a_c, b_s0, b_s1, b_r, b_c;
    int b1, M, K, N;
    float alpha;
};

__global__ void __launch_bounds__(128) small_bmm_kernel(const SmallBmm p) {
    // 128 threads, 4 x 8 outputs each: per k step 3 LDS.128 feed 32 FMAs, so the FMA pipe (not the
    // shared-memory port) is the limit
    constexpr int TM = 64, TN = 64, TK = 32;
    __shared__ __align__(16) float As[TK][TM + 4];
    __shared__ __align__(16) float Bs[TK][TN + 4];
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;  // tx: 8 column groups of 8, ty: 16 row groups of 4
    const int bz = blockIdx.z, i0 = bz / p.b1, i1 = bz - i0 * p.b1;
    const float* A = p.A + i0 * p.a_s0 + i1 * p.a_s1;
    const float* B = p.B + i0 * p.b_s0 + i1 * p.b_s1;
    float* C = p.C + (long long)bz * p.M * p.N;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const bool a_k_fast = p.a_c == 1 || p.a_r != 1;   // consecutive threads walk the contiguous dimension
    const bool b_n_fast = p.b_c == 1 || p.b_r != 1;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < p.K; k0 += TK) {
#pragma unroll 4
        for (int i = 0; i < (TM * TK) / 128; ++i) {
            const int e = tid + i * 128;
            const int r = a_k_fast ? e / TK : e % TM, kk = a_k_fast ? e % TK : e / TM;
            As[kk][r] = (m0 + r < p.M && k0 + kk < p.K) ? A[(long long)(m0 + r) * p.a_r + (long long)(k0 + kk) * p.a_c] : 0.f;
            const int c = b_n_fast ? e % TN : e / TK, kb = b_n_fast ? e / TN : e % TK;
            Bs[kb][c] = (n0 + c < p.N && k0 + kb < p.K) ? B[(long long)(k0 + kb) * p.b_r + (long long)(n0 + c) * p.b_c] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < TK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8 + 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = m0 + ty * 4 + i;
        if (r >= p.M) continue;
        float* crow = C + (long long)r * p.N;
        const int c0 = n0 + tx * 8;
        if (c0 + 7 < p.N && ((reinterpret_cast<uintptr_t>(crow + c0) & 15) == 0)) {
            *reinterpret_cast<float4*>(crow + c0) = make_float4(p.alpha * acc[i][0], p.alpha * acc[i][1], p.alpha * acc[i][2], p.alpha * acc[i][3]);
            *reinterpret_cast<float4*>(crow + c0 + 4) = make_float4(p.alpha * acc[i][4], p.alpha * acc[i][5], p.alpha * acc[i][6], p.alpha * acc[i][7]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (c0 + j < p.N) crow[c0 + j] = p.alpha * acc[i][j];
        }
    }
}

// Off by default: measured inside the GPT-small step the tcgen05 path (2 staging launches + a half-padded
// 128-row tile) still wins, 5.92 vs 6.36 ms per step. nnb_matmul_set_small_path(1) or
// NNB_MATMUL_SIMT=1 selects the fp32 kernel (exact fp32 products, one launch per call).
int g_small_path = [] { const char* e = getenv("NNB_MATMUL_SIMT"); return e ? atoi(e) : 0; }();

bool small_bmm_ok(int64_t batch, int64_t M, int64_t K, int64_t N) {
    return g_small_path && batch >= 16 && M <= 128 && N <= 128 && K <= 128;
}

// strides {s0, s1, row, col} of the logical [b0,b1,R,C] operand
int small_bmm(const float* A, const int64_t as[4], const float* B, const int64_t bs[4], float* C, int64_t b0,
              int64_t b1, int64_t M, int64_t K, int64_t N, float alpha, cudaStream_t stream) {
    NNB_REQUIRE(b0 * b1 <= 65535, "matmul: more than 65535 batch elements on the small-product path");
    SmallBmm p{A, B, C, as[0], as[1], as[2], as[3], bs[0], bs[1], bs[2], bs[3], (int)b1, (int)M, (int)K, (int)N, alpha};
    dim3 grid((unsigned)ceil_div(N, 64), (unsigned)ceil_div(M, 64), (unsigned)(b0 * b1));
    small_bmm_kernel<<<grid, 128, 0, stream>>>(p);
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
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

int nnb_matmul_set_small_path(int on) {
    const int prev = g_small_path;
    g_small_path = on ? 1 : 0;
    return prev;
}

int nnb_matmul_uses_tensor_cores(int64_t b0, int64_t b1, int64_t M, int64_t K, int64_t N) {
    return (small_bmm_ok(b0 * b1, M, K, N) && b0 * b1 <= 65535) ? 0 : 1;
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
    NNB_RANGE("nnb_matmul_forward");
    NNB_REQUIRE(A && B && C && a_strides && b_strides, "nnb_matmul_forward: null pointer");
    NNB_REQUIRE(b0 > 0 && b1 > 0 && M > 0 && K > 0 && N > 0, "nnb_matmul_forward: non-positive dimension");
    NNB_REQUIRE(prec == NNB_PREC_BF16 || prec == NNB_PREC_BF16X3, "nnb_matmul_forward: bad prec");
    if (small_bmm_ok(b0 * b1, M, K, N) && b0 * b1 <= 65535)
        return small_bmm(A, a_strides, B, b_strides, C, b0, b1, M, K, N, alpha, stream);
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
    NNB_RANGE("nnb_matmul_backward");
    NNB_REQUIRE(A && B && G && a_strides && b_strides, "nnb_matmul_backward: null pointer");
    NNB_REQUIRE(b0 > 0 && b1 > 0 && M > 0 && K > 0 && N > 0, "nnb_matmul_backward: non-positive dimension");
    NNB_REQUIRE(prec == NNB_PREC_BF16 || prec == NNB_PREC_BF16X3, "nnb_matmul_backward: bad prec");
    const int64_t batch = b0 * b1;
    int64_t gst[4];
    contig_strides(b1, M, N, gst);
    if (small_bmm_ok(batch, M, K, N) && batch <= 65535) {
        if (dA) {  // dA[M,K] = alpha * G[M,N] . (B[K,N])^T: B^T is the same memory with row/col strides swapped
            const int64_t bt[4] = {b_strides[0], b_strides[1], b_strides[3], b_strides[2]};
            int rc = small_bmm(G, gst, B, bt, dA, b0, b1, M, N, K, alpha, stream);
            if (rc) return rc;
        }
        if (dB) {  // dB[K,N] = alpha * (A[M,K])^T . G[M,N]
            const int64_t at[4] = {a_strides[0], a_strides[1], a_strides[3], a_strides[2]};
            int rc = small_bmm(A, at, G, gst, dB, b0, b1, K, M, N, alpha, stream);
            if (rc) return rc;
        }
        return NNB_OK;
    }
    Bump ws(workspace, workspace_bytes);
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
