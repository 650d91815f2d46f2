// nn.Linear forward / backward on the tcgen05 GEMM.
// Reference semantics: neunet/nn/layers/linear.py:17-24 (backward), 48-58 (forward);
// fused Swish: neunet/nn/experimental/linear_swish/linear_swish_cutlass_evt_full.cu:558-818.
#include "common.cuh"
#include "philox.cuh"
#include "workspace.cuh"

namespace nnb {

static int planes(int prec) { return prec == NNB_PREC_BF16X3 ? 2 : 1; }

static Staged weight_view(const void* blob, int64_t rows, int64_t cols, int prec) {
    Staged s;
    const uint8_t* b = static_cast<const uint8_t*>(blob);
    s.hi = reinterpret_cast<const __nv_bfloat16*>(b);
    s.lo = prec == NNB_PREC_BF16X3
               ? reinterpret_cast<const __nv_bfloat16*>(b + staged_plane_bytes(1, rows, cols))
               : nullptr;
    s.rows = rows;
    s.cols = cols;
    s.ld = staged_ld(cols);
    s.batch = 1;
    s.batch_stride = rows * s.ld;
    return s;
}

static int stage_into(Bump& ws, const View4& v, int prec, int op, const float* aux, float beta,
                      float* colsum, cudaStream_t stream, Staged* out, const DropArgs* drop = nullptr) {
    const int64_t batch = v.b0 * v.b1;
    auto* hi = static_cast<__nv_bfloat16*>(ws.take(staged_plane_bytes(batch, v.rows, v.cols)));
    __nv_bfloat16* lo = nullptr;
    if (prec == NNB_PREC_BF16X3)
        lo = static_cast<__nv_bfloat16*>(ws.take(staged_plane_bytes(batch, v.rows, v.cols)));
    float* scratch = nullptr;
    if (colsum) scratch = static_cast<float*>(ws.take(stage_colsum_scratch_bytes(v.cols)));
    if (!ws.ok()) return fail(NNB_ERR_WORKSPACE, "workspace too small (need >= %zu bytes)", ws.off);
    return stage_operand(v, false, prec, hi, lo, op, aux, beta, colsum, scratch, stream, out, drop);
}

}  // namespace nnb

using namespace nnb;

extern "C" {

size_t nnb_weight_staged_bytes(int64_t rows, int64_t cols, int prec) {
    if (rows <= 0 || cols <= 0) return 0;
    return planes(prec) * staged_plane_bytes(1, rows, cols);
}

int nnb_stage_weight(const float* W, int64_t rows, int64_t cols, int prec, void* dst,
                     cudaStream_t stream) {
    NNB_RANGE("nnb_stage_weight");
    NNB_REQUIRE(W && dst, "nnb_stage_weight: null pointer");
    NNB_REQUIRE(rows > 0 && cols > 0, "nnb_stage_weight: non-positive dimension");
    NNB_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 255) == 0, "nnb_stage_weight: dst must be 256-byte aligned");
    uint8_t* b = static_cast<uint8_t*>(dst);
    auto* hi = reinterpret_cast<__nv_bfloat16*>(b);
    auto* lo = prec == NNB_PREC_BF16X3
                   ? reinterpret_cast<__nv_bfloat16*>(b + staged_plane_bytes(1, rows, cols))
                   : nullptr;
    return stage_operand(view2d(W, rows, cols, cols), false, prec, hi, lo, STAGE_COPY, nullptr, 0.f,
                         nullptr, nullptr, stream, nullptr);
}

size_t nnb_linear_workspace_bytes(int64_t M, int64_t K, int64_t N, int prec, int backward) {
    if (M <= 0 || K <= 0 || N <= 0) return 0;
    const size_t p = planes(prec);
    size_t b = 4096;
    b += p * staged_plane_bytes(1, M, K);  // X
    b += p * staged_plane_bytes(1, N, K);  // W (when the caller passes no W_staged)
    if (!backward) {
        b += gemm_splitk_ws_bytes(M, N, K, 1);
    } else {
        b += p * staged_plane_bytes(1, M, N);  // dZ
        b += stage_colsum_scratch_bytes(N);
        b += std::max(gemm_splitk_ws_bytes(M, K, N, 1), gemm_splitk_ws_bytes(N, K, M, 1));
    }
    return b;
}

int nnb_linear_forward(const float* X, const float* W, const float* bias, float* O, float* Z,
                       int64_t M, int64_t K, int64_t N, int act, float beta, int prec,
                       const void* W_staged, void* X_staged_out, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream) {
    NNB_RANGE("nnb_linear_forward");
    NNB_REQUIRE(X && W && O, "nnb_linear_forward: null X/W/O");
    NNB_REQUIRE(M > 0 && K > 0 && N > 0, "nnb_linear_forward: non-positive dimension");
    NNB_REQUIRE(prec == NNB_PREC_BF16 || prec == NNB_PREC_BF16X3, "nnb_linear_forward: bad prec");
    NNB_REQUIRE(act == NNB_ACT_NONE || act == NNB_ACT_SWISH, "nnb_linear_forward: bad act");
    Bump ws(workspace, workspace_bytes);
    GemmProblem g;
    g.M = M; g.N = N; g.K = K;
    int rc;
    if (X_staged_out) {  // stage X into the caller's buffer so the backward pass can reuse it
        NNB_REQUIRE((reinterpret_cast<uintptr_t>(X_staged_out) & 255) == 0, "nnb_linear_forward: X_staged_out must be 256-byte aligned");
        rc = nnb_stage_weight(X, M, K, prec, X_staged_out, stream);
        if (rc) return rc;
        g.A.st = weight_view(X_staged_out, M, K, prec);
    } else {
        rc = stage_into(ws, view2d(X, M, K, K), prec, STAGE_COPY, nullptr, 0.f, nullptr, stream, &g.A.st);
        if (rc) return rc;
    }
    if (W_staged) {
        g.B.st = weight_view(W_staged, N, K, prec);
    } else {
        rc = stage_into(ws, view2d(W, N, K, K), prec, STAGE_COPY, nullptr, 0.f, nullptr, stream, &g.B.st);
        if (rc) return rc;
    }
    g.D = O; g.ldd = N;
    g.epi.bias = bias; g.epi.Z = Z; g.epi.act = act; g.epi.beta = beta;
    g.splitk_ws_bytes = ws.remaining();
    g.splitk_ws = static_cast<float*>(ws.take(g.splitk_ws_bytes));
    return gemm(g, stream);
}

int nnb_linear_forward_staged(const void* X_staged, const void* W_staged, const float* bias, float* O,
                              float* Z, int64_t M, int64_t K, int64_t N, int act, float beta,
                              int prec, void* workspace, size_t workspace_bytes,
                              cudaStream_t stream) {
    NNB_RANGE("nnb_linear_forward_staged");
    NNB_REQUIRE(X_staged && W_staged && O, "nnb_linear_forward_staged: null pointer");
    NNB_REQUIRE(M > 0 && K > 0 && N > 0, "nnb_linear_forward_staged: non-positive dimension");
    NNB_REQUIRE(prec == NNB_PREC_BF16 || prec == NNB_PREC_BF16X3, "nnb_linear_forward_staged: bad prec");
    Bump ws(workspace, workspace_bytes);
    GemmProblem g;
    g.M = M; g.N = N; g.K = K;
    g.A.st = weight_view(X_staged, M, K, prec);
    g.B.st = weight_view(W_staged, N, K, prec);
    g.D = O; g.ldd = N;
    g.epi.bias = bias; g.epi.Z = Z; g.epi.act = act; g.epi.beta = beta;
    g.splitk_ws_bytes = ws.remaining();
    g.splitk_ws = static_cast<float*>(ws.take(g.splitk_ws_bytes));
    return gemm(g, stream);
}

static int linear_backward_impl(const float* X, const float* W, const float* Z, const float* dO,
                                float* dX, float* dW, float* db, int64_t M, int64_t K, int64_t N, int act,
                                float beta, int prec, const void* W_staged, const void* X_staged,
                                void* workspace, size_t workspace_bytes, cudaStream_t stream, const nnb::DropArgs* drop) {
    NNB_REQUIRE((X || X_staged) && W && dO && dW, "nnb_linear_backward: null X/W/dO/dW");
    NNB_REQUIRE(M > 0 && K > 0 && N > 0, "nnb_linear_backward: non-positive dimension");
    NNB_REQUIRE(prec == NNB_PREC_BF16 || prec == NNB_PREC_BF16X3, "nnb_linear_backward: bad prec");
    NNB_REQUIRE(act == NNB_ACT_NONE || (act == NNB_ACT_SWISH && Z), "nnb_linear_backward: Swish needs Z");
    Bump ws(workspace, workspace_bytes);
    Staged gs, xs, wsd;
    // dZ = dO (* swish'(Z)); db = column sums of dZ, fused into the same pass over dO
    int rc = stage_into(ws, view2d(dO, M, N, N), prec, act == NNB_ACT_SWISH ? STAGE_SWISH_BWD : STAGE_COPY,
                        Z, beta, db, stream, &gs, drop);
    if (rc) return rc;
    if (X_staged) {
        xs = weight_view(X_staged, M, K, prec);
    } else {
        rc = stage_into(ws, view2d(X, M, K, K), prec, STAGE_COPY, nullptr, 0.f, nullptr, stream, &xs);
        if (rc) return rc;
    }
    if (dX) {
        if (W_staged) {
            wsd = weight_view(W_staged, N, K, prec);
        } else {
            rc = stage_into(ws, view2d(W, N, K, K), prec, STAGE_COPY, nullptr, 0.f, nullptr, stream, &wsd);
            if (rc) return rc;
        }
    }
    const size_t sk_bytes = ws.remaining();
    float* sk = static_cast<float*>(ws.take(sk_bytes));
    if (dX) {
        // dX[M,K] = dZ[M,N] . W[N,K]: reduction over N. dZ is K-major; W (rows = N) is MN-major.
        GemmProblem g;
        g.M = M; g.N = K; g.K = N;
        g.A.st = gs; g.A.mn_major = false;
        g.B.st = wsd; g.B.mn_major = true;
        g.D = dX; g.ldd = K;
        g.splitk_ws = sk; g.splitk_ws_bytes = sk_bytes;
        rc = gemm(g, stream);
        if (rc) return rc;
    }
    {
        // dW[N,K] = dZ^T[N,M] . X[M,K]: reduction over M; both operands are MN-major as stored.
        GemmProblem g;
        g.M = N; g.N = K; g.K = M;
        g.A.st = gs; g.A.mn_major = true;
        g.B.st = xs; g.B.mn_major = true;
        g.D = dW; g.ldd = K;
        g.splitk_ws = sk; g.splitk_ws_bytes = sk_bytes;
        rc = gemm(g, stream);
        if (rc) return rc;
    }
    return NNB_OK;
}

int nnb_linear_backward(const float* X, const float* W, const float* Z, const float* dO,
                        float* dX, float* dW, float* db, int64_t M, int64_t K, int64_t N, int act,
                        float beta, int prec, const void* W_staged, const void* X_staged,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    NNB_RANGE("nnb_linear_backward");
    return linear_backward_impl(X, W, Z, dO, dX, dW, db, M, K, N, act, beta, prec, W_staged, X_staged, workspace,
                                workspace_bytes, stream, nullptr);
}

int nnb_linear_backward_dropped(const float* X, const float* W, const float* Z, const float* dO,
                                float* dX, float* dW, float* db, int64_t M, int64_t K, int64_t N, int act,
                                float beta, int prec, const void* W_staged, const void* X_staged,
                                void* workspace, size_t workspace_bytes, float drop_p, uint64_t seed, uint32_t call_id,
                                uint64_t epoch, const uint64_t* epoch_dev, cudaStream_t stream) {
    NNB_RANGE("nnb_linear_backward_dropped");
    NNB_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "nnb_linear_backward_dropped: p must be in [0, 1)");
    if (N % 4 != 0) return fail(NNB_ERR_UNSUPPORTED, "nnb_linear_backward_dropped: needs N %% 4 == 0");
    const nnb::DropArgs d = nnb::make_drop_args(drop_p, seed, call_id, epoch, epoch_dev);
    return linear_backward_impl(X, W, Z, dO, dX, dW, db, M, K, N, act, beta, prec, W_staged, X_staged, workspace,
                                workspace_bytes, stream, &d);
}

int nnb_linear_backward_staged(const float* X, const float* W, const void* dO_staged, float* dX, float* dW,
                               int64_t M, int64_t K, int64_t N, int prec, const void* W_staged, const void* X_staged,
                               void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    NNB_RANGE("nnb_linear_backward_staged");
    NNB_REQUIRE((X || X_staged) && (W || W_staged || !dX) && dO_staged && dW, "nnb_linear_backward_staged: null pointer");
    NNB_REQUIRE(M > 0 && K > 0 && N > 0, "nnb_linear_backward_staged: non-positive dimension");
    NNB_REQUIRE(prec == NNB_PREC_BF16 || prec == NNB_PREC_BF16X3, "nnb_linear_backward_staged: bad prec");
    Bump ws(workspace, workspace_bytes);
    Staged gs = weight_view(dO_staged, M, N, prec), xs, wsd;
    int rc;
    if (X_staged) {
        xs = weight_view(X_staged, M, K, prec);
    } else {
        rc = stage_into(ws, view2d(X, M, K, K), prec, STAGE_COPY, nullptr, 0.f, nullptr, stream, &xs);
        if (rc) return rc;
    }
    if (dX) {
        if (W_staged) {
            wsd = weight_view(W_staged, N, K, prec);
        } else {
            rc = stage_into(ws, view2d(W, N, K, K), prec, STAGE_COPY, nullptr, 0.f, nullptr, stream, &wsd);
            if (rc) return rc;
        }
    }
    const size_t sk_bytes = ws.remaining();
    float* sk = static_cast<float*>(ws.take(sk_bytes));
    if (dX) {
        GemmProblem g;
        g.M = M; g.N = K; g.K = N;
        g.A.st = gs; g.A.mn_major = false;
        g.B.st = wsd; g.B.mn_major = true;
        g.D = dX; g.ldd = K;
        g.splitk_ws = sk; g.splitk_ws_bytes = sk_bytes;
        rc = gemm(g, stream);
        if (rc) return rc;
    }
    GemmProblem g;
    g.M = N; g.N = K; g.K = M;
    g.A.st = gs; g.A.mn_major = true;
    g.B.st = xs; g.B.mn_major = true;
    g.D = dW; g.ldd = K;
    g.splitk_ws = sk; g.splitk_ws_bytes = sk_bytes;
    return gemm(g, stream);
}

}  // extern "C"
