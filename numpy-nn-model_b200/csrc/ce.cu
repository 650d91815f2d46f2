// Fused CrossEntropyLoss = LogSoftmax(axis=1) + NLLLoss (neunet/nn/losses.py:59-126,
// neunet/nn/activations.py:462-491) for 2-D logits and unit class weights; the reference's
// native analogue is cudaCrossEntropyForwardBackward (experimental/losses/cross_entropy_loss/
// cross_entropy.cu:18-292). Forward reads the logits once (online max/sum per row) and keeps the
// per-row log-sum-exp; backward writes dlogits = (softmax - onehot) * keep / denom * upstream in one
// pass. The GPT example's 15 000-wide logits are the largest activation of the model, so the
// ~40 element-wise launches of the composed form are replaced by three kernels.
#include <cfloat>

#include "common.cuh"

namespace nnb {
namespace {

__device__ __forceinline__ void online_merge(float& m, float& s, float m2, float s2) {
    const float mm = fmaxf(m, m2);
    s = s * expf(m - mm) + s2 * expf(m2 - mm);
    m = mm;
}

// one warp (C <= 2048) or one block per row
template <int THREADS>
__global__ void __launch_bounds__(256)
ce_rows_kernel(const float* __restrict__ logits, const int* __restrict__ targets, long long rows, int C,
               int ignore_index, float* __restrict__ row_loss, float* __restrict__ lse_out) {
    constexpr bool WARP = THREADS == 32;
    pdl_trigger();
    pdl_wait();
    const long long row = WARP ? ((long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)) : blockIdx.x;
    const int lane = WARP ? (threadIdx.x & 31) : threadIdx.x;
    const int stride = WARP ? 32 : THREADS;
    if (row >= rows) return;
    const float* x = logits + row * C;
    float m = -FLT_MAX, s = 0.f;
    if (((reinterpret_cast<uintptr_t>(x) & 15) == 0) && C >= 4) {
        // 16-byte loads; the running max moves rarely, so one rescale per float4 instead of per element
        const int c4 = C >> 2;
        for (int i = lane; i < c4; i += stride) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
            const float vm = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
            if (vm > m) { s *= expf(m - vm); m = vm; }
            s += expf(v.x - m) + expf(v.y - m) + expf(v.z - m) + expf(v.w - m);
        }
        for (int i = (c4 << 2) + lane; i < C; i += stride) {
            const float v = x[i];
            if (v > m) { s *= expf(m - v); m = v; }
            s += expf(v - m);
        }
    } else {
        for (int i = lane; i < C; i += stride) {
            const float v = x[i];
            if (v > m) { s *= expf(m - v); m = v; }
            s += expf(v - m);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        online_merge(m, s, m2, s2);
    }
    if (!WARP) {
        __shared__ float sm[THREADS / 32], ss[THREADS / 32];
        if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = m; ss[threadIdx.x >> 5] = s; }
        __syncthreads();
        if (threadIdx.x < 32) {
            m = threadIdx.x < THREADS / 32 ? sm[threadIdx.x] : -FLT_MAX;
            s = threadIdx.x < THREADS / 32 ? ss[threadIdx.x] : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
                online_merge(m, s, m2, s2);
            }
        }
    }
    if (lane == 0) {
        const float lse = m + logf(s);
        lse_out[row] = lse;
        // targets are range-checked: negative labels wrap like NumPy's x[arange, t]; anything else outside [0, C)
        // (the reference raises IndexError, losses.py:107) poisons the row with NaN instead of reading out of bounds
        int t = targets[row];
        const bool ignored = t == ignore_index;  // decided on the raw label, like `y != ignore_index` (losses.py:104)
        if (!ignored && t < 0 && t >= -C) t += C;
        row_loss[row] = ignored ? 0.f : ((t >= 0 && t < C) ? (lse - x[t]) : __int_as_float(0x7fc00000));   // -log_softmax[target]
    }
}

// fixed-order single-block finish: loss = sum(row_loss) (/ count of kept rows), inv_denom for backward
__global__ void __launch_bounds__(1024)
ce_finish_kernel(const float* __restrict__ row_loss, const int* __restrict__ targets, long long rows,
                 int ignore_index, int reduction, float* __restrict__ loss_out, float* __restrict__ inv_denom) {
    __shared__ float ssum[1024];
    __shared__ float scnt[1024];
    pdl_trigger();
    pdl_wait();
    float s = 0.f, c = 0.f;
    for (long long i = threadIdx.x; i < rows; i += 1024) {
        s += row_loss[i];
        c += (targets[i] != ignore_index) ? 1.f : 0.f;
    }
    ssum[threadIdx.x] = s;
    scnt[threadIdx.x] = c;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { ssum[threadIdx.x] += ssum[threadIdx.x + o]; scnt[threadIdx.x] += scnt[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float inv = (reduction == 1) ? 1.0f / scnt[0] : 1.0f;
        *inv_denom = inv;
        *loss_out = ssum[0] * inv;
    }
}

// grid = (column chunks, rows): the per-row scalars (target, lse, upstream) are read once per block and
// the row is streamed with 16-byte accesses (4 B read + 4 B written per logit)
__global__ void __launch_bounds__(256)
ce_backward_kernel(const float* __restrict__ logits, const int* __restrict__ targets,
                   const float* __restrict__ lse, const float* __restrict__ inv_denom,
                   const float* __restrict__ upstream, int upstream_per_row, long long rows, int C,
                   int ignore_index, float* __restrict__ dlogits, int vec) {
    pdl_trigger();
    pdl_wait();
    const long long r = blockIdx.y;
    int t = targets[r];
    const bool ignored = t == ignore_index;
    if (!ignored && t < 0 && t >= -C) t += C;  // NumPy wrap, as in the forward kernel
    const float* x = logits + r * C;
    float* d = dlogits + r * C;
    const int c0 = blockIdx.x * (256 * 4 * 4);
    const int c1 = min(c0 + 256 * 4 * 4, C);
    if (ignored) {
        for (int c = c0 + threadIdx.x; c < c1; c += 256) d[c] = 0.f;
        return;
    }
    const float l = lse[r];
    float sc = (inv_denom ? *inv_denom : 1.0f) * (upstream_per_row ? upstream[r] : upstream[0]);
    if (t < 0 || t >= C) sc = __int_as_float(0x7fc00000);  // out-of-range label: NaN row, matching the poisoned loss
    if (vec) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int c = c0 + (it * 256 + threadIdx.x) * 4;
            if (c + 3 < c1) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(x + c));
                float4 g;
                g.x = (expf(v.x - l) - (c == t ? 1.f : 0.f)) * sc;
                g.y = (expf(v.y - l) - (c + 1 == t ? 1.f : 0.f)) * sc;
                g.z = (expf(v.z - l) - (c + 2 == t ? 1.f : 0.f)) * sc;
                g.w = (expf(v.w - l) - (c + 3 == t ? 1.f : 0.f)) * sc;
                *reinterpret_cast<float4*>(d + c) = g;
            } else {
                for (int j = c; j < c1; ++j) d[j] = (expf(x[j] - l) - (j == t ? 1.f : 0.f)) * sc;
            }
        }
    } else {
        for (int c = c0 + threadIdx.x; c < c1; c += 256) d[c] = (expf(x[c] - l) - (c == t ? 1.f : 0.f)) * sc;
    }
}


// lm_head + CrossEntropy (row N3 of SURVEY.md 8f): dlogits = (softmax - onehot) * keep / denom * upstream written
// DIRECTLY as the bf16 operand planes (hi [+ lo]) that the Linear's dgrad / wgrad GEMMs read, plus per-block column partial
// sums for the bias gradient -- the fp32 dlogits (246 MB at the GPT sizes) are never materialised and the Linear backward
// needs no staging pass. Block = 64 rows x 256 columns; thread = two adjacent columns, walking the rows.
constexpr int CE_ROWS = 64;
template <bool X3>
__global__ void __launch_bounds__(128)
ce_backward_staged_kernel(const float* __restrict__ logits, const int* __restrict__ targets, const float* __restrict__ lse,
                          const float* __restrict__ inv_denom, const float* __restrict__ upstream, long long rows, int C,
                          int ignore_index, long long ld, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                          float* __restrict__ partial) {
    __shared__ int s_t[CE_ROWS];
    __shared__ float s_l[CE_ROWS];
    pdl_trigger();
    pdl_wait();
    const long long r0 = (long long)blockIdx.y * CE_ROWS;
    const int nr = (int)min((long long)CE_ROWS, rows - r0);
    const float sc0 = (inv_denom ? *inv_denom : 1.0f) * upstream[0];
    for (int i = threadIdx.x; i < CE_ROWS; i += 128) {
        int t = -2;  // -2: ignored row (all zeros); -1: out-of-range label (NaN row)
        if (i < nr) {
            t = targets[r0 + i];
            if (t == ignore_index) t = -2;
            else { if (t < 0 && t >= -C) t += C; if (t < 0 || t >= C) t = -1; }
            s_l[i] = lse[r0 + i];
        }
        s_t[i] = t;
    }
    __syncthreads();
    const int c = (blockIdx.x * 128 + threadIdx.x) * 2;
    if (c >= C) return;
    const bool two = c + 1 < C;
    float cs0 = 0.f, cs1 = 0.f;
    for (int i = 0; i < nr; ++i) {
        const int t = s_t[i];
        float g0 = 0.f, g1 = 0.f;
        if (t != -2) {
            const float sc = t == -1 ? __int_as_float(0x7fc00000) : sc0;
            const float* x = logits + (r0 + i) * C + c;
            g0 = (expf(x[0] - s_l[i]) - (c == t ? 1.f : 0.f)) * sc;
            if (two) g1 = (expf(x[1] - s_l[i]) - (c + 1 == t ? 1.f : 0.f)) * sc;
        }
        cs0 += g0;
        cs1 += g1;
        const __nv_bfloat16 h0 = __float2bfloat16_rn(g0), h1 = __float2bfloat16_rn(g1);
        const long long o = (r0 + i) * ld + c;  // ld and c are even: 4-byte aligned pair (pad column written as 0)
        *reinterpret_cast<__nv_bfloat162*>(hi + o) = __nv_bfloat162(h0, h1);
        if (X3)
            *reinterpret_cast<__nv_bfloat162*>(lo + o) = __nv_bfloat162(__float2bfloat16_rn(g0 - __bfloat162float(h0)),
                                                                         __float2bfloat16_rn(g1 - __bfloat162float(h1)));
    }
    float* prow = partial + (long long)blockIdx.y * C;
    prow[c] = cs0;
    if (two) prow[c + 1] = cs1;
}

// db[c] = sum over row blocks (fixed order)
__global__ void ce_colsum_kernel(const float* __restrict__ partial, int nblocks, int C, float* __restrict__ db) {
    pdl_trigger();
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int b = 0; b < nblocks; ++b) s += partial[(long long)b * C + c];
    db[c] = s;
}

}  // namespace
}  // namespace nnb

using namespace nnb;

extern "C" {

int nnb_cross_entropy_forward(const float* logits, const int32_t* targets, int64_t rows, int64_t C,
                              int64_t ignore_index, int reduction, float* row_loss, float* lse,
                              float* loss_out, float* inv_denom, cudaStream_t stream) {
    NNB_RANGE("nnb_cross_entropy_forward");
    NNB_REQUIRE(logits && targets && row_loss && lse, "nnb_cross_entropy_forward: null pointer");
    NNB_REQUIRE(rows > 0 && C > 0 && C < (1ll << 31), "nnb_cross_entropy_forward: bad shape");
    NNB_REQUIRE(reduction >= 0 && reduction <= 2, "nnb_cross_entropy_forward: bad reduction");
    NNB_REQUIRE(reduction == 0 || (loss_out && inv_denom), "nnb_cross_entropy_forward: reduction needs loss_out/inv_denom");
    if (C <= 2048) {
        const int wpb = 8;
        NNB_CUDA_OK(launch_pdl(ce_rows_kernel<32>, dim3((unsigned)ceil_div(rows, wpb)), dim3(32 * wpb), 0, stream, logits,
                               (const int*)targets, (long long)rows, (int)C, (int)ignore_index, row_loss, lse));
    } else {
        NNB_CUDA_OK(launch_pdl(ce_rows_kernel<256>, dim3((unsigned)rows), dim3(256), 0, stream, logits, (const int*)targets,
                               (long long)rows, (int)C, (int)ignore_index, row_loss, lse));
    }
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    if (reduction != 0) {
        NNB_CUDA_OK(launch_pdl(ce_finish_kernel, dim3(1), dim3(1024), 0, stream, (const float*)row_loss, (const int*)targets,
                               (long long)rows, (int)ignore_index, reduction, loss_out, inv_denom));
        count_launch();
        NNB_CUDA_OK(cudaGetLastError());
    }
    return NNB_OK;
}

int nnb_cross_entropy_backward(const float* logits, const int32_t* targets, const float* lse,
                               const float* inv_denom, const float* upstream, int upstream_per_row,
                               int64_t rows, int64_t C, int64_t ignore_index, float* dlogits,
                               cudaStream_t stream) {
    NNB_RANGE("nnb_cross_entropy_backward");
    NNB_REQUIRE(logits && targets && lse && upstream && dlogits, "nnb_cross_entropy_backward: null pointer");
    NNB_REQUIRE(rows > 0 && C > 0 && C < (1ll << 31), "nnb_cross_entropy_backward: bad shape");
    const int vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(dlogits)) & 15) == 0;
    // blockIdx.y is limited to 65535: fold larger row counts through several launches
    for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
        const int64_t nr = std::min<int64_t>(65535, rows - r0);
        dim3 grid((unsigned)ceil_div(C, 256 * 16), (unsigned)nr);
        NNB_CUDA_OK(launch_pdl(ce_backward_kernel, grid, dim3(256), 0, stream, logits + r0 * C, (const int*)(targets + r0),
                               lse + r0, inv_denom, upstream_per_row ? upstream + r0 : upstream, upstream_per_row,
                               (long long)nr, (int)C, (int)ignore_index, dlogits + r0 * C, vec));
        count_launch();
    }
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

size_t nnb_cross_entropy_staged_workspace_bytes(int64_t rows, int64_t C) {
    if (rows <= 0 || C <= 0) return 0;
    return (size_t)round_up(ceil_div(rows, CE_ROWS) * C * 4, 256) + 256;
}

int nnb_cross_entropy_backward_staged(const float* logits, const int32_t* targets, const float* lse,
                                      const float* inv_denom, const float* upstream, int64_t rows, int64_t C,
                                      int64_t ignore_index, void* dZ_staged_out, int prec, float* db, void* workspace,
                                      size_t workspace_bytes, cudaStream_t stream) {
    NNB_RANGE("nnb_cross_entropy_backward_staged");
    NNB_REQUIRE(logits && targets && lse && upstream && dZ_staged_out, "nnb_cross_entropy_backward_staged: null pointer");
    NNB_REQUIRE(rows > 0 && C > 0 && C < (1ll << 31), "nnb_cross_entropy_backward_staged: bad shape");
    NNB_REQUIRE(prec == NNB_PREC_BF16 || prec == NNB_PREC_BF16X3, "nnb_cross_entropy_backward_staged: bad prec");
    NNB_REQUIRE((reinterpret_cast<uintptr_t>(dZ_staged_out) & 255) == 0, "nnb_cross_entropy_backward_staged: dZ_staged_out must be 256-byte aligned");
    NNB_REQUIRE(workspace && workspace_bytes >= nnb_cross_entropy_staged_workspace_bytes(rows, C), "nnb_cross_entropy_backward_staged: workspace too small");
    float* partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    auto* hi = static_cast<__nv_bfloat16*>(dZ_staged_out);
    auto* lo = prec == NNB_PREC_BF16X3
                   ? reinterpret_cast<__nv_bfloat16*>(static_cast<uint8_t*>(dZ_staged_out) + staged_plane_bytes(1, rows, C))
                   : nullptr;
    const int nrb = (int)ceil_div(rows, CE_ROWS);
    NNB_REQUIRE(nrb <= 65535, "nnb_cross_entropy_backward_staged: too many rows");
    dim3 grid((unsigned)ceil_div(C, 256), (unsigned)nrb);
    if (lo) NNB_CUDA_OK(launch_pdl(ce_backward_staged_kernel<true>, grid, dim3(128), 0, stream, logits, (const int*)targets, lse, inv_denom,
                                   upstream, (long long)rows, (int)C, (int)ignore_index, (long long)staged_ld(C), hi, lo, partial));
    else NNB_CUDA_OK(launch_pdl(ce_backward_staged_kernel<false>, grid, dim3(128), 0, stream, logits, (const int*)targets, lse, inv_denom,
                                upstream, (long long)rows, (int)C, (int)ignore_index, (long long)staged_ld(C), hi, lo, partial));
    count_launch();
    if (db != nullptr) {
        NNB_CUDA_OK(launch_pdl(ce_colsum_kernel, dim3((unsigned)ceil_div(C, 256)), dim3(256), 0, stream, (const float*)partial, nrb, (int)C, db));
        count_launch();
    }
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

}  // extern "C"
