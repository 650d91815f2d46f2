// Stand-alone (un-fused) forms of the epilogue ops: Swish, Softmax(axis), RMSNorm.
// All are HBM-bound streaming kernels: 16-byte accesses where alignment allows, warp-shuffle
// reductions, grids sized in multiples of the SM count.
// Reference semantics: neunet/nn/activations.py:208-233 (Swish), 437-459 (Softmax),
// neunet/nn/layers/rmsnorm.py:39-94 (RMSNorm).
#include <algorithm>
#include <cfloat>

#include "common.cuh"
#include "philox.cuh"
#include "workspace.cuh"

namespace nnb {
namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float swish_f(float x, float beta) { return x * sigmoidf_(beta * x); }
__device__ __forceinline__ float swish_df(float x, float beta) {
    const float s = sigmoidf_(beta * x);
    const float f = x * s;
    return beta * f + s * (1.0f - beta * f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------- Swish
template <bool BWD>
__global__ void swish_kernel(const float* __restrict__ x, const float* __restrict__ g,
                             float* __restrict__ y, long long n, float beta, int vec) {
    pdl_trigger();
    pdl_wait();
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (vec) {
        const long long n4 = n >> 2;
        for (long long i = tid; i < n4; i += stride) {
            const float4 a = reinterpret_cast<const float4*>(x)[i];
            float4 r;
            if (BWD) {
                const float4 gg = reinterpret_cast<const float4*>(g)[i];
                r.x = gg.x * swish_df(a.x, beta); r.y = gg.y * swish_df(a.y, beta);
                r.z = gg.z * swish_df(a.z, beta); r.w = gg.w * swish_df(a.w, beta);
            } else {
                r.x = swish_f(a.x, beta); r.y = swish_f(a.y, beta);
                r.z = swish_f(a.z, beta); r.w = swish_f(a.w, beta);
            }
            reinterpret_cast<float4*>(y)[i] = r;
        }
        for (long long i = (n4 << 2) + tid; i < n; i += stride)
            y[i] = BWD ? g[i] * swish_df(x[i], beta) : swish_f(x[i], beta);
    } else {
        for (long long i = tid; i < n; i += stride)
            y[i] = BWD ? g[i] * swish_df(x[i], beta) : swish_f(x[i], beta);
    }
}

// ---------------------------------------------------------------- Softmax
// inner == 1: one warp per row (online max/sum in one pass, second pass writes).
template <bool BWD>
__global__ void softmax_rows_kernel(const float* __restrict__ a, const float* __restrict__ g,
                                    float* __restrict__ out, long long rows, int n) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const float* ar = a + row * n;
    float* orow = out + row * n;
    if (!BWD) {
        float m = -FLT_MAX, s = 0.f;
        for (int i = lane; i < n; i += 32) {
            const float v = ar[i];
            if (v > m) { s *= expf(m - v); m = v; }
            s += expf(v - m);
        }
        const float mm = warp_max(m);
        s *= expf(m - mm);
        s = warp_sum(s);
        for (int i = lane; i < n; i += 32) orow[i] = expf(ar[i] - mm) / s;
    } else {
        const float* gr = g + row * n;
        float d = 0.f;
        for (int i = lane; i < n; i += 32) d += gr[i] * ar[i];
        d = warp_sum(d);
        for (int i = lane; i < n; i += 32) orow[i] = (gr[i] - d) * ar[i];
    }
}

// inner > 1: one thread per (outer, inner) column, striding over the softmax axis (coalesced on inner).
template <bool BWD>
__global__ void softmax_strided_kernel(const float* __restrict__ a, const float* __restrict__ g,
                                       float* __restrict__ out, long long outer, int n,
                                       long long inner) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= outer * inner) return;
    const long long o = idx / inner, in = idx - o * inner;
    const long long base = o * n * inner + in;
    if (!BWD) {
        float m = -FLT_MAX;
        for (int i = 0; i < n; ++i) m = fmaxf(m, a[base + i * inner]);
        float s = 0.f;
        for (int i = 0; i < n; ++i) s += expf(a[base + i * inner] - m);
        for (int i = 0; i < n; ++i) out[base + i * inner] = expf(a[base + i * inner] - m) / s;
    } else {
        float d = 0.f;
        for (int i = 0; i < n; ++i) d += g[base + i * inner] * a[base + i * inner];
        for (int i = 0; i < n; ++i) out[base + i * inner] = (g[base + i * inner] - d) * a[base + i * inner];
    }
}

// ---------------------------------------------------------------- RMSNorm
// one warp per row
__global__ void rmsnorm_fwd_kernel(const float* __restrict__ X, const float* __restrict__ w,
                                   const float* __restrict__ b, float* __restrict__ Y,
                                   float* __restrict__ Xstd, float* __restrict__ Xnorm,
                                   long long rows, int cols, float eps) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const float* x = X + row * cols;
    float ss = 0.f;
    for (int i = lane; i < cols; i += 32) ss += x[i] * x[i];
    ss = warp_sum(ss);
    const float std = sqrtf(ss / (float)cols + eps);
    if (lane == 0 && Xstd) Xstd[row] = std;
    for (int i = lane; i < cols; i += 32) {
        const float xn = x[i] / std;
        if (Xnorm) Xnorm[row * cols + i] = xn;
        float y = xn * w[i];
        if (b) y += b[i];
        Y[row * cols + i] = y;
    }
}

// Fused forward for cols <= 1024, cols % 4 == 0 (one warp per row, the row lives in registers):
//   S = X (+ dropout(A))        -- optional residual prologue: `x = x + dropout(a)` then `norm(x)` (gpt cell 5)
//   Y = S / sqrt(mean(S^2) + eps) * w (+ b), X_std, and optionally the bf16 planes of Y for the next Linear
// so the residual add, the dropout, the norm and the operand staging of the following GEMM are ONE pass.
template <int NJ4>
__global__ void __launch_bounds__(256) rmsnorm_fwd_fused_kernel(
    const float* __restrict__ X, const float* __restrict__ A, const DropArgs d, float* __restrict__ S,
    const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ Y, float* __restrict__ Xstd,
    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long rows, int cols, float eps) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const uint64_t epoch = A != nullptr ? drop_epoch(d) : 0;
    float4 x[NJ4];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NJ4; ++j) {
        const int c = (lane + 32 * j) * 4;
        if (c < cols) {
            x[j] = *reinterpret_cast<const float4*>(X + row * cols + c);
            if (A != nullptr) {
                float4 a = *reinterpret_cast<const float4*>(A + row * cols + c);
                uint32_t r[4];
                drop_words(d, epoch, (row * cols + c) >> 2, r);
                a.x = r[0] >= d.thresh ? a.x * d.scale : 0.f;
                a.y = r[1] >= d.thresh ? a.y * d.scale : 0.f;
                a.z = r[2] >= d.thresh ? a.z * d.scale : 0.f;
                a.w = r[3] >= d.thresh ? a.w * d.scale : 0.f;
                x[j].x += a.x; x[j].y += a.y; x[j].z += a.z; x[j].w += a.w;
                *reinterpret_cast<float4*>(S + row * cols + c) = x[j];
            }
            ss += x[j].x * x[j].x + x[j].y * x[j].y + x[j].z * x[j].z + x[j].w * x[j].w;
        } else {
            x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    ss = warp_sum(ss);
    const float std = sqrtf(ss / (float)cols + eps);
    if (lane == 0 && Xstd) Xstd[row] = std;
    const float inv_std = 1.0f / std;  // one IEEE division per row; x * (1 / std) is within 1 ulp of x / std
#pragma unroll
    for (int j = 0; j < NJ4; ++j) {
        const int c = (lane + 32 * j) * 4;
        if (c >= cols) continue;
        const float4 wv = *reinterpret_cast<const float4*>(w + c);
        float4 y;
        y.x = x[j].x * inv_std * wv.x; y.y = x[j].y * inv_std * wv.y; y.z = x[j].z * inv_std * wv.z; y.w = x[j].w * inv_std * wv.w;
        if (b != nullptr) {
            const float4 bv = *reinterpret_cast<const float4*>(b + c);
            y.x += bv.x; y.y += bv.y; y.z += bv.z; y.w += bv.w;
        }
        *reinterpret_cast<float4*>(Y + row * cols + c) = y;
        if (hi != nullptr) {
            const float yv[4] = {y.x, y.y, y.z, y.w};
            __nv_bfloat16 h[4], l[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                h[k] = __float2bfloat16_rn(yv[k]);
                l[k] = __float2bfloat16_rn(yv[k] - __bfloat162float(h[k]));
            }
            *reinterpret_cast<uint2*>(hi + row * cols + c) = *reinterpret_cast<const uint2*>(h);
            if (lo != nullptr) *reinterpret_cast<uint2*>(lo + row * cols + c) = *reinterpret_cast<const uint2*>(l);
        }
    }
}

__global__ void rmsnorm_bwd_dx_kernel(const float* __restrict__ gY, const float* __restrict__ X,
                                      const float* __restrict__ w, const float* __restrict__ Xstd,
                                      float* __restrict__ dX, long long rows, int cols) {
    const int lane = threadIdx.x & 31;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const float* x = X + row * cols;
    const float* g = gY + row * cols;
    const float std = Xstd[row];
    float dot = 0.f;
    for (int i = lane; i < cols; i += 32) dot += w[i] * g[i] * x[i] / std;
    dot = warp_sum(dot);
    const float c = dot / (float)cols;
    const float inv2 = 1.0f / (std * std);
    // (dXhat * std - x * sum(dXhat * x / std) / N) / std^2   (rmsnorm.py:51)
    for (int i = lane; i < cols; i += 32) dX[row * cols + i] = (w[i] * g[i] * std - x[i] * c) * inv2;
}

// column partial sums of gY * (X / std) and gY over a chunk of rows; thread = column
__global__ void rmsnorm_bwd_cols_kernel(const float* __restrict__ gY, const float* __restrict__ X,
                                        const float* __restrict__ Xstd, float* __restrict__ pw,
                                        float* __restrict__ pb, long long rows, int cols,
                                        int rows_per_block) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(r0 + rows_per_block, rows);
    float sw = 0.f, sb = 0.f;
    for (long long r = r0; r < r1; ++r) {
        const float g = gY[r * cols + c];
        sw += g * (X[r * cols + c] / Xstd[r]);
        sb += g;
    }
    pw[(long long)blockIdx.y * cols + c] = sw;
    if (pb) pb[(long long)blockIdx.y * cols + c] = sb;
}

// Fused RMSNorm backward for cols <= 1024, cols % 4 == 0: one pass over gY and X produces dX AND the
// per-block column partials of dw = sum_rows gY * X/std and db = sum_rows gY (rmsnorm.py:51-59).
// A warp owns a row at a time (float4 per lane, NJ4 of them), keeps its column partials in registers
// across its rows, and the 8 warps of the block are combined through shared memory at the end.
template <int NJ4>
__global__ void __launch_bounds__(256) rmsnorm_bwd_fused_kernel(
    const float* __restrict__ gY, const float* __restrict__ X, const float* __restrict__ w,
    const float* __restrict__ Xstd, float* __restrict__ dX, float* __restrict__ pw,
    float* __restrict__ pb, long long rows, int cols, int rows_per_block, const float* __restrict__ dXadd) {
    __shared__ float red[8][NJ4 * 128 + 4];
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = min(r0 + rows_per_block, rows);
    float4 wv[NJ4], aw[NJ4], ab[NJ4];
#pragma unroll
    for (int j = 0; j < NJ4; ++j) {
        const int c = (lane + 32 * j) * 4;
        wv[j] = c < cols ? *reinterpret_cast<const float4*>(w + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        aw[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        ab[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (long long r = r0 + warp; r < r1; r += 8) {
        float4 x[NJ4], g[NJ4];
#pragma unroll
        for (int j = 0; j < NJ4; ++j) {
            const int c = (lane + 32 * j) * 4;
            if (c < cols) {
                x[j] = *reinterpret_cast<const float4*>(X + r * cols + c);
                g[j] = *reinterpret_cast<const float4*>(gY + r * cols + c);
            } else {
                x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                g[j] = x[j];
            }
        }
        const float std = Xstd[r];
        const float inv_std = 1.0f / std;  // one division per row instead of eight per 16 bytes
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < NJ4; ++j) {
            dot += wv[j].x * g[j].x * (x[j].x * inv_std) + wv[j].y * g[j].y * (x[j].y * inv_std) +
                   wv[j].z * g[j].z * (x[j].z * inv_std) + wv[j].w * g[j].w * (x[j].w * inv_std);
        }
        dot = warp_sum(dot);
        const float cc = dot / (float)cols;
        const float inv2 = inv_std * inv_std;
#pragma unroll
        for (int j = 0; j < NJ4; ++j) {
            const int c = (lane + 32 * j) * 4;
            if (c < cols) {
                float4 d;
                d.x = (wv[j].x * g[j].x * std - x[j].x * cc) * inv2;
                d.y = (wv[j].y * g[j].y * std - x[j].y * cc) * inv2;
                d.z = (wv[j].z * g[j].z * std - x[j].z * cc) * inv2;
                d.w = (wv[j].w * g[j].w * std - x[j].w * cc) * inv2;
                if (dXadd != nullptr) {  // accumulate onto the gradient X already holds (residual branch)
                    const float4 e = *reinterpret_cast<const float4*>(dXadd + r * cols + c);
                    d.x += e.x; d.y += e.y; d.z += e.z; d.w += e.w;
                }
                *reinterpret_cast<float4*>(dX + r * cols + c) = d;
            }
            aw[j].x += g[j].x * (x[j].x * inv_std); aw[j].y += g[j].y * (x[j].y * inv_std);
            aw[j].z += g[j].z * (x[j].z * inv_std); aw[j].w += g[j].w * (x[j].w * inv_std);
            ab[j].x += g[j].x; ab[j].y += g[j].y; ab[j].z += g[j].z; ab[j].w += g[j].w;
        }
    }
    if (pw == nullptr) return;
    for (int pass = 0; pass < 2; ++pass) {
        float* dst = pass == 0 ? pw : pb;
        if (dst == nullptr) break;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < NJ4; ++j)
            *reinterpret_cast<float4*>(&red[warp][(lane + 32 * j) * 4]) = pass == 0 ? aw[j] : ab[j];
        __syncthreads();
        for (int c = threadIdx.x; c < cols; c += 256) {
            float t = 0.f;
#pragma unroll
            for (int y = 0; y < 8; ++y) t += red[y][c];
            dst[(long long)blockIdx.x * cols + c] = t;
        }
    }
}

// dw[c] = sum_p pw[p][c] (and db from pb): 32 columns x 32 partial-lanes per block, fixed order
__global__ void __launch_bounds__(1024) colsum2_reduce_kernel(const float* __restrict__ pw, const float* __restrict__ pb,
                                                              int nparts, int cols, float* __restrict__ dw,
                                                              float* __restrict__ db) {
    __shared__ float red[2][32][33];
    pdl_trigger();
    pdl_wait();
    const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    float sw = 0.f, sb = 0.f;
    if (c < cols) {
#pragma unroll 4
        for (int p = py; p < nparts; p += 32) {
            sw += __ldg(pw + (long long)p * cols + c);
            if (db) sb += __ldg(pb + (long long)p * cols + c);
        }
    }
    red[0][py][cx] = sw;
    red[1][py][cx] = sb;
    __syncthreads();
    if (py == 0 && c < cols) {
        float tw = 0.f, tb = 0.f;
#pragma unroll
        for (int y = 0; y < 32; ++y) { tw += red[0][y][cx]; tb += red[1][y][cx]; }
        dw[c] = tw;
        if (db) db[c] = tb;
    }
}

int ew_grid(long long n_threads_needed, int threads) {
    const long long blocks = ceil_div(n_threads_needed, threads);
    return (int)std::max<long long>(1, std::min<long long>(blocks, (long long)num_sms() * 16));
}

// partial rows of the dw / db column sums. 512: with 4096 rows every warp of the fused backward owns ONE row (28 warps per
// SM instead of 14 -- the kernel is latency-bound, ncu: 39 % of peak warps on the forward twin) and the finishing pass
// still reads only 2 x 512 x cols floats.
constexpr int RMS_MAX_PARTS = 512;

}  // namespace
}  // namespace nnb

using namespace nnb;

extern "C" {

int nnb_swish_forward(const float* x, float* y, int64_t n, float beta, cudaStream_t stream) {
    NNB_RANGE("nnb_swish_forward");
    NNB_REQUIRE(x && y && n > 0, "nnb_swish_forward: bad arguments");
    const int vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    NNB_CUDA_OK(launch_pdl(swish_kernel<false>, dim3(ew_grid(vec ? (n + 3) / 4 : n, 256)), dim3(256), 0, stream, x,
                           (const float*)nullptr, y, (long long)n, beta, vec));
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

int nnb_swish_backward(const float* x, const float* grad, float* dx, int64_t n, float beta,
                       cudaStream_t stream) {
    NNB_RANGE("nnb_swish_backward");
    NNB_REQUIRE(x && grad && dx && n > 0, "nnb_swish_backward: bad arguments");
    const int vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(grad) |
                      reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
    NNB_CUDA_OK(launch_pdl(swish_kernel<true>, dim3(ew_grid(vec ? (n + 3) / 4 : n, 256)), dim3(256), 0, stream, x, grad, dx,
                           (long long)n, beta, vec));
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

int nnb_softmax_forward(const float* x, float* y, int64_t outer, int64_t n, int64_t inner,
                        cudaStream_t stream) {
    NNB_RANGE("nnb_softmax_forward");
    NNB_REQUIRE(x && y, "nnb_softmax_forward: null pointer");
    NNB_REQUIRE(outer > 0 && n > 0 && inner > 0 && n < (1ll << 31), "nnb_softmax_forward: bad shape");
    if (inner == 1) {
        const long long blocks = ceil_div(outer * 32, 256);
        NNB_CUDA_OK(launch_pdl(softmax_rows_kernel<false>, dim3((unsigned)blocks), dim3(256), 0, stream, x,
                               (const float*)nullptr, y, (long long)outer, (int)n));
    } else {
        const long long blocks = ceil_div(outer * inner, 256);
        softmax_strided_kernel<false><<<(unsigned)blocks, 256, 0, stream>>>(x, nullptr, y, outer, (int)n, inner);
    }
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

int nnb_softmax_backward(const float* y, const float* grad, float* dx, int64_t outer, int64_t n,
                         int64_t inner, cudaStream_t stream) {
    NNB_RANGE("nnb_softmax_backward");
    NNB_REQUIRE(y && grad && dx, "nnb_softmax_backward: null pointer");
    NNB_REQUIRE(outer > 0 && n > 0 && inner > 0 && n < (1ll << 31), "nnb_softmax_backward: bad shape");
    if (inner == 1) {
        const long long blocks = ceil_div(outer * 32, 256);
        NNB_CUDA_OK(launch_pdl(softmax_rows_kernel<true>, dim3((unsigned)blocks), dim3(256), 0, stream, y, grad, dx,
                               (long long)outer, (int)n));
    } else {
        const long long blocks = ceil_div(outer * inner, 256);
        softmax_strided_kernel<true><<<(unsigned)blocks, 256, 0, stream>>>(y, grad, dx, outer, (int)n, inner);
    }
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

int nnb_rmsnorm_forward(const float* X, const float* w, const float* b, float* Y, float* X_std,
                        float* X_norm, int64_t rows, int64_t cols, float eps, cudaStream_t stream) {
    NNB_RANGE("nnb_rmsnorm_forward");
    NNB_REQUIRE(X && w && Y, "nnb_rmsnorm_forward: null pointer");
    NNB_REQUIRE(rows > 0 && cols > 0 && cols < (1ll << 31), "nnb_rmsnorm_forward: bad shape");
    const long long blocks = ceil_div(rows * 32, 256);
    NNB_CUDA_OK(launch_pdl(rmsnorm_fwd_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, X, w, b, Y, X_std, X_norm,
                           (long long)rows, (int)cols, eps));
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

int nnb_rmsnorm_forward_fused(const float* X, const float* A, float p, uint64_t seed, uint32_t call_id,
                              uint64_t epoch, const uint64_t* epoch_dev, float* S, const float* w, const float* b,
                              float* Y, float* X_std, void* Y_staged_out, int prec, int64_t rows, int64_t cols,
                              float eps, cudaStream_t stream) {
    NNB_RANGE("nnb_rmsnorm_forward_fused");
    NNB_REQUIRE(X && w && Y, "nnb_rmsnorm_forward_fused: null pointer");
    NNB_REQUIRE(rows > 0 && cols > 0, "nnb_rmsnorm_forward_fused: bad shape");
    NNB_REQUIRE(A == nullptr || S != nullptr, "nnb_rmsnorm_forward_fused: the residual prologue needs the sum output S");
    NNB_REQUIRE(A == nullptr || (p >= 0.f && p < 1.f), "nnb_rmsnorm_forward_fused: p must be in [0, 1)");
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!(cols <= 1024 && cols % 4 == 0 && al16(X) && al16(w) && al16(Y) && (!A || (al16(A) && al16(S))) && (!b || al16(b)) &&
          (Y_staged_out == nullptr || cols % 8 == 0)))
        return fail(NNB_ERR_UNSUPPORTED, "nnb_rmsnorm_forward_fused: needs cols <= 1024, cols %% 4 == 0 (8 with planes), 16-byte aligned rows");
    __nv_bfloat16 *hi = nullptr, *lo = nullptr;
    if (Y_staged_out != nullptr) {
        NNB_REQUIRE((reinterpret_cast<uintptr_t>(Y_staged_out) & 255) == 0, "nnb_rmsnorm_forward_fused: Y_staged_out must be 256-byte aligned");
        NNB_REQUIRE(prec == NNB_PREC_BF16 || prec == NNB_PREC_BF16X3, "nnb_rmsnorm_forward_fused: bad prec");
        hi = static_cast<__nv_bfloat16*>(Y_staged_out);
        if (prec == NNB_PREC_BF16X3)
            lo = reinterpret_cast<__nv_bfloat16*>(static_cast<uint8_t*>(Y_staged_out) + staged_plane_bytes(1, rows, cols));
    }
    const DropArgs d = make_drop_args(A ? p : 0.f, seed, call_id, epoch, epoch_dev);
    const dim3 grid((unsigned)ceil_div(rows * 32, 256)), block(256);
    const long long rows_ll = rows;
    const int cols_i = (int)cols, nj4 = (int)ceil_div(cols, 128);
    if (nj4 <= 1) NNB_CUDA_OK(launch_pdl(rmsnorm_fwd_fused_kernel<1>, grid, block, 0, stream, X, A, d, S, w, b, Y, X_std, hi, lo, rows_ll, cols_i, eps));
    else if (nj4 <= 2) NNB_CUDA_OK(launch_pdl(rmsnorm_fwd_fused_kernel<2>, grid, block, 0, stream, X, A, d, S, w, b, Y, X_std, hi, lo, rows_ll, cols_i, eps));
    else if (nj4 <= 4) NNB_CUDA_OK(launch_pdl(rmsnorm_fwd_fused_kernel<4>, grid, block, 0, stream, X, A, d, S, w, b, Y, X_std, hi, lo, rows_ll, cols_i, eps));
    else NNB_CUDA_OK(launch_pdl(rmsnorm_fwd_fused_kernel<8>, grid, block, 0, stream, X, A, d, S, w, b, Y, X_std, hi, lo, rows_ll, cols_i, eps));
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

size_t nnb_rmsnorm_workspace_bytes(int64_t rows, int64_t cols) {
    (void)rows;
    return (size_t)(2 * RMS_MAX_PARTS * cols * 4 + 512);
}

int nnb_rmsnorm_backward(const float* gY, const float* X, const float* w, const float* X_std,
                         const float* X_norm, float* dX, float* dw, float* db, int64_t rows,
                         int64_t cols, void* workspace, size_t workspace_bytes,
                         cudaStream_t stream) {
    return nnb_rmsnorm_backward_acc(gY, X, w, X_std, X_norm, nullptr, dX, dw, db, rows, cols, workspace, workspace_bytes, stream);
}

int nnb_rmsnorm_backward_acc(const float* gY, const float* X, const float* w, const float* X_std,
                             const float* X_norm, const float* dX_add, float* dX, float* dw, float* db,
                             int64_t rows, int64_t cols, void* workspace, size_t workspace_bytes,
                             cudaStream_t stream) {
    NNB_RANGE("nnb_rmsnorm_backward_acc");
    (void)X_norm;  // recomputed as X / X_std: cheaper than reading a second [rows, cols] array
    NNB_REQUIRE(gY && X && w && X_std && dX, "nnb_rmsnorm_backward: null pointer");
    NNB_REQUIRE(rows > 0 && cols > 0 && cols < (1ll << 31), "nnb_rmsnorm_backward: bad shape");
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const bool fused = cols <= 1024 && (cols % 4) == 0 && al16(gY) && al16(X) && al16(w) && al16(dX) && (!dX_add || al16(dX_add));
    NNB_REQUIRE(fused || dX_add == nullptr, "nnb_rmsnorm_backward_acc: dX_add needs the fused path (cols <= 1024, cols %% 4 == 0)");
    float *pw = nullptr, *pb = nullptr;
    if (dw) {
        Bump ws(workspace, workspace_bytes);
        pw = static_cast<float*>(ws.take((size_t)RMS_MAX_PARTS * cols * 4));
        pb = db ? static_cast<float*>(ws.take((size_t)RMS_MAX_PARTS * cols * 4)) : nullptr;
        if (!ws.ok()) return fail(NNB_ERR_WORKSPACE, "nnb_rmsnorm_backward: workspace too small");
    }
    if (fused) {
        // >= 8 rows per block (one per warp), <= RMS_MAX_PARTS partial rows for the finishing pass
        long long parts = std::max<long long>(1, std::min<long long>(RMS_MAX_PARTS, ceil_div(rows, 8)));
        const int rpb = (int)ceil_div(rows, parts);
        parts = ceil_div(rows, rpb);
        const int nj4 = (int)ceil_div(cols, 128);
        const dim3 fg((unsigned)parts), fb(256);
        const long long rows_ll = rows;
        const int cols_i = (int)cols;
        if (nj4 <= 1) NNB_CUDA_OK(launch_pdl(rmsnorm_bwd_fused_kernel<1>, fg, fb, 0, stream, gY, X, w, X_std, dX, pw, pb, rows_ll, cols_i, rpb, dX_add));
        else if (nj4 <= 2) NNB_CUDA_OK(launch_pdl(rmsnorm_bwd_fused_kernel<2>, fg, fb, 0, stream, gY, X, w, X_std, dX, pw, pb, rows_ll, cols_i, rpb, dX_add));
        else if (nj4 <= 4) NNB_CUDA_OK(launch_pdl(rmsnorm_bwd_fused_kernel<4>, fg, fb, 0, stream, gY, X, w, X_std, dX, pw, pb, rows_ll, cols_i, rpb, dX_add));
        else NNB_CUDA_OK(launch_pdl(rmsnorm_bwd_fused_kernel<8>, fg, fb, 0, stream, gY, X, w, X_std, dX, pw, pb, rows_ll, cols_i, rpb, dX_add));
        count_launch();
        NNB_CUDA_OK(cudaGetLastError());
        if (dw) {
            NNB_CUDA_OK(launch_pdl(colsum2_reduce_kernel, dim3((unsigned)ceil_div(cols, 32)), dim3(1024), 0, stream,
                                   (const float*)pw, (const float*)pb, (int)parts, (int)cols, dw, db));
            count_launch();
            NNB_CUDA_OK(cudaGetLastError());
        }
        return NNB_OK;
    }
    const long long blocks = ceil_div(rows * 32, 256);
    rmsnorm_bwd_dx_kernel<<<(unsigned)blocks, 256, 0, stream>>>(gY, X, w, X_std, dX, rows, (int)cols);
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    if (dw) {
        const int gx = (int)ceil_div(cols, 128);
        long long parts = std::max<long long>(1, std::min<long long>(RMS_MAX_PARTS, (long long)num_sms() * 4 / gx));
        parts = std::min<long long>(parts, rows);
        const int rpb = (int)ceil_div(rows, parts);
        parts = ceil_div(rows, rpb);
        rmsnorm_bwd_cols_kernel<<<dim3(gx, (unsigned)parts), 128, 0, stream>>>(gY, X, X_std, pw, pb, rows, (int)cols, rpb);
        NNB_CUDA_OK(launch_pdl(colsum2_reduce_kernel, dim3((unsigned)ceil_div(cols, 32)), dim3(1024), 0, stream,
                               (const float*)pw, (const float*)pb, (int)parts, (int)cols, dw, db));
        count_launch(2);
        NNB_CUDA_OK(cudaGetLastError());
    }
    return NNB_OK;
}

}  // extern "C"
