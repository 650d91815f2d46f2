// Fused LeakyReLU + BatchNorm2d (row N2 of SURVEY.md 8f). Every DDPM ResBlock is conv -> LeakyReLU -> BatchNorm2d
// twice (examples/ddpm.ipynb cell 5 l.41-55); as separate array ops that is ~12 element-wise / reduction passes
// per block. Here, over NCHW fp32:
//   a = leaky_relu(x, alpha)                       (neunet/nn/activations.py:73-93; alpha = 1 -> plain BatchNorm2d)
//   y = (a - mean_c) / sqrt(var_c + eps) * w_c + b_c   with batch statistics over (B, H, W), ddof = 0
//                                                  (neunet/nn/layers/batchnorm2d.py:57-115)
// forward  = nnb_bn_stats (per-channel sum / sum of squares of a, fp64 accumulators) -> [optional NCCL all-reduce of
//            the 2*C doubles: SyncBN] -> nnb_bn_finalize (mean, 1/std, running stats with the reference's
//            momentum convention, batchnorm2d.py:87-88) -> nnb_bn_apply
// backward = nnb_bn_backward_stats (sum g, sum g*xhat) -> [optional all-reduce] -> nnb_bn_backward_apply
//            dx = lrelu'(x) * inv * (w*g - mean(w*g) - xhat * mean(w*g*xhat));  dw = sum g*xhat;  db = sum g
// Only x, mean and 1/std are saved for backward: a and xhat are recomputed (HBM-bound: one read per pass).
#include <algorithm>

#include "common.cuh"

namespace nnb {
namespace {

__device__ __forceinline__ float lrelu(float x, float alpha) { return x <= 0.f ? alpha * x : x; }

constexpr int BN_CHUNKS = 32;  // partial-sum rows per channel (fixed order -> deterministic)

// grid = (C, chunks): block (c, j) reduces images b = j, j + chunks, ... of channel c.
// MODE 0: {sum a, sum a^2};  MODE 1: {sum g, sum g * xhat}
template <int MODE>
__global__ void __launch_bounds__(256) bn_reduce_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                        const float* __restrict__ mean, const float* __restrict__ inv,
                                                        int B, int C, int HW, float alpha, double* __restrict__ partial) {
    pdl_trigger();
    pdl_wait();
    const int c = blockIdx.x;
    const float mu = MODE == 1 ? mean[c] : 0.f, is = MODE == 1 ? inv[c] : 0.f;
    double s0 = 0.0, s1 = 0.0;
    const bool vec = (HW % 4) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g)) & 15) == 0;
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const float* xp = x + ((long long)b * C + c) * HW;
        const float* gp = MODE == 1 ? g + ((long long)b * C + c) * HW : nullptr;
        float t0 = 0.f, t1 = 0.f;  // fp32 within one image plane per thread (<= HW / 256 terms), fp64 across
        if (vec) {
            for (int i = threadIdx.x * 4; i < HW; i += 1024) {
                const float4 v = *reinterpret_cast<const float4*>(xp + i);
                const float a[4] = {lrelu(v.x, alpha), lrelu(v.y, alpha), lrelu(v.z, alpha), lrelu(v.w, alpha)};
                if (MODE == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { t0 += a[j]; t1 += a[j] * a[j]; }
                } else {
                    const float4 gv = *reinterpret_cast<const float4*>(gp + i);
                    const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) { t0 += gg[j]; t1 += gg[j] * ((a[j] - mu) * is); }
                }
            }
        } else {
            for (int i = threadIdx.x; i < HW; i += 256) {
                const float a = lrelu(xp[i], alpha);
                if (MODE == 0) { t0 += a; t1 += a * a; }
                else { t0 += gp[i]; t1 += gp[i] * ((a - mu) * is); }
            }
        }
        s0 += (double)t0;
        s1 += (double)t1;
    }
    __shared__ double r0[256], r1[256];
    r0[threadIdx.x] = s0;
    r1[threadIdx.x] = s1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { r0[threadIdx.x] += r0[threadIdx.x + o]; r1[threadIdx.x] += r1[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[((long long)blockIdx.y * C + c) * 2] = r0[0];
        partial[((long long)blockIdx.y * C + c) * 2 + 1] = r1[0];
    }
}

// sums[c] = {sum_j partial[j][c][0], sum_j partial[j][c][1]}  (fixed order)
__global__ void bn_fold_kernel(const double* __restrict__ partial, int chunks, int C, double* __restrict__ sums) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * C) return;
    double s = 0.0;
    for (int j = 0; j < chunks; ++j) s += partial[(long long)j * 2 * C + i];
    sums[i] = s;
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, int C, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ inv, float* __restrict__ run_mean,
                                   float* __restrict__ run_var) {
    pdl_trigger();
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = sums[2 * c] / count;
    double v = sums[2 * c + 1] / count - m * m;  // fp64: no cancellation problem at fp32 data precision
    if (v < 0.0) v = 0.0;
    mean[c] = (float)m;
    inv[c] = 1.0f / sqrtf((float)v + eps);
    if (run_mean != nullptr) {
        // reference convention (batchnorm2d.py:87-88): momentum weights the OLD value
        run_mean[c] = momentum * run_mean[c] + (1.0f - momentum) * (float)m;
        run_var[c] = momentum * run_var[c] + (1.0f - momentum) * (float)v;
    }
}

// MODE 0: y = (lrelu(x) - mean) * inv * w + b
// MODE 1: dx = lrelu'(x) * inv * (w * g - s_g * w / n - xhat * s_gx * w / n)
template <int MODE>
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                       const float* __restrict__ mean, const float* __restrict__ inv,
                                                       const float* __restrict__ w, const float* __restrict__ b,
                                                       const double* __restrict__ sums, double count, int C, int HW,
                                                       long long planes, float alpha, float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const bool vec = (HW % 4) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                                        reinterpret_cast<uintptr_t>(g)) & 15) == 0;
    for (long long pl = blockIdx.x; pl < planes; pl += gridDim.x) {  // one (b, c) image plane at a time
        const int c = (int)(pl % C);
        const float mu = mean[c], is = inv[c], wc = w ? w[c] : 1.f, bc = b ? b[c] : 0.f;
        float k1 = 0.f, k2 = 0.f;
        if (MODE == 1) {
            k1 = (float)(sums[2 * c] / count) * wc;       // mean(w * g)
            k2 = (float)(sums[2 * c + 1] / count) * wc;   // mean(w * g * xhat)
        }
        const float* xp = x + pl * HW;
        const float* gp = MODE == 1 ? g + pl * HW : nullptr;
        float* op = out + pl * HW;
        if (vec) {
            for (int i = threadIdx.x * 4; i < HW; i += blockDim.x * 4) {
                const float4 v = *reinterpret_cast<const float4*>(xp + i);
                const float xv[4] = {v.x, v.y, v.z, v.w};
                float r[4];
                if (MODE == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) r[j] = (lrelu(xv[j], alpha) - mu) * is * wc + bc;
                } else {
                    const float4 gv = *reinterpret_cast<const float4*>(gp + i);
                    const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float xh = (lrelu(xv[j], alpha) - mu) * is;
                        r[j] = (xv[j] <= 0.f ? alpha : 1.f) * is * (wc * gg[j] - k1 - xh * k2);
                    }
                }
                *reinterpret_cast<float4*>(op + i) = make_float4(r[0], r[1], r[2], r[3]);
            }
        } else {
            for (int i = threadIdx.x; i < HW; i += blockDim.x) {
                if (MODE == 0) {
                    op[i] = (lrelu(xp[i], alpha) - mu) * is * wc + bc;
                } else {
                    const float xh = (lrelu(xp[i], alpha) - mu) * is;
                    op[i] = (xp[i] <= 0.f ? alpha : 1.f) * is * (wc * gp[i] - k1 - xh * k2);
                }
            }
        }
    }
}

__global__ void bn_param_grads_kernel(const double* __restrict__ sums, int C, float* __restrict__ dw, float* __restrict__ db) {
    pdl_trigger();
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (db) db[c] = (float)sums[2 * c];
    if (dw) dw[c] = (float)sums[2 * c + 1];
}

int check_shape(int64_t B, int64_t C, int64_t HW, const char* who) {
    NNB_REQUIRE(B > 0 && C > 0 && HW > 0, "%s: non-positive dimension", who);
    NNB_REQUIRE(B < (1ll << 31) && C < 65536 && HW < (1ll << 31), "%s: dimension too large", who);
    return NNB_OK;
}

int chunks_for(int64_t B) { return (int)std::min<int64_t>(BN_CHUNKS, B); }

template <int MODE>
int run_reduce(const float* x, const float* g, const float* mean, const float* inv, int64_t B, int64_t C, int64_t HW,
               float alpha, double* sums, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    const int chunks = chunks_for(B);
    NNB_REQUIRE(workspace != nullptr && workspace_bytes >= (size_t)chunks * C * 2 * sizeof(double) + 256,
                "nnb_bn: workspace too small (nnb_bn_workspace_bytes)");
    double* partial = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    NNB_CUDA_OK(launch_pdl(bn_reduce_kernel<MODE>, dim3((unsigned)C, (unsigned)chunks), dim3(256), 0, stream, x, g, mean, inv,
                           (int)B, (int)C, (int)HW, alpha, partial));
    NNB_CUDA_OK(launch_pdl(bn_fold_kernel, dim3((unsigned)ceil_div(2 * C, 256)), dim3(256), 0, stream, (const double*)partial,
                           chunks, (int)C, sums));
    count_launch(2);
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

}  // namespace
}  // namespace nnb

using namespace nnb;

extern "C" {

size_t nnb_bn_workspace_bytes(int64_t B, int64_t C) {
    if (B <= 0 || C <= 0) return 0;
    return (size_t)BN_CHUNKS * C * 2 * sizeof(double) + 512;
}

int nnb_bn_stats(const float* x, int64_t B, int64_t C, int64_t HW, float alpha, double* sums, void* workspace,
                 size_t workspace_bytes, cudaStream_t stream) {
    NNB_RANGE("nnb_bn_stats");
    NNB_REQUIRE(x && sums, "nnb_bn_stats: null pointer");
    int rc = check_shape(B, C, HW, "nnb_bn_stats");
    if (rc) return rc;
    return run_reduce<0>(x, nullptr, nullptr, nullptr, B, C, HW, alpha, sums, workspace, workspace_bytes, stream);
}

int nnb_bn_finalize(const double* sums, double count, int64_t C, float eps, float momentum, float* mean,
                    float* inv_std, float* running_mean, float* running_var, cudaStream_t stream) {
    NNB_RANGE("nnb_bn_finalize");
    NNB_REQUIRE(sums && mean && inv_std, "nnb_bn_finalize: null pointer");
    NNB_REQUIRE(C > 0 && count > 0, "nnb_bn_finalize: bad size");
    NNB_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "nnb_bn_finalize: running stats come in pairs");
    NNB_CUDA_OK(launch_pdl(bn_finalize_kernel, dim3((unsigned)ceil_div(C, 256)), dim3(256), 0, stream, sums, count, (int)C, eps,
                           momentum, mean, inv_std, running_mean, running_var));
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

int nnb_bn_apply(const float* x, const float* mean, const float* inv_std, const float* w, const float* b,
                 int64_t B, int64_t C, int64_t HW, float alpha, float* y, cudaStream_t stream) {
    NNB_RANGE("nnb_bn_apply");
    NNB_REQUIRE(x && mean && inv_std && y, "nnb_bn_apply: null pointer");
    int rc = check_shape(B, C, HW, "nnb_bn_apply");
    if (rc) return rc;
    const long long planes = (long long)B * C;
    const int blocks = (int)std::min<long long>(planes, (long long)num_sms() * 16);
    const int threads = HW >= 1024 ? 256 : (HW >= 256 ? 64 : 32);
    NNB_CUDA_OK(launch_pdl(bn_apply_kernel<0>, dim3(blocks), dim3(threads), 0, stream, x, (const float*)nullptr, mean, inv_std, w, b,
                           (const double*)nullptr, 1.0, (int)C, (int)HW, planes, alpha, y));
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

int nnb_bn_backward_stats(const float* x, const float* grad, const float* mean, const float* inv_std, int64_t B,
                          int64_t C, int64_t HW, float alpha, double* sums, void* workspace, size_t workspace_bytes,
                          cudaStream_t stream) {
    NNB_RANGE("nnb_bn_backward_stats");
    NNB_REQUIRE(x && grad && mean && inv_std && sums, "nnb_bn_backward_stats: null pointer");
    int rc = check_shape(B, C, HW, "nnb_bn_backward_stats");
    if (rc) return rc;
    return run_reduce<1>(x, grad, mean, inv_std, B, C, HW, alpha, sums, workspace, workspace_bytes, stream);
}

int nnb_bn_backward_apply(const float* x, const float* grad, const float* mean, const float* inv_std, const float* w,
                          const double* sums, double count, int64_t B, int64_t C, int64_t HW, float alpha,
                          float* dx, float* dw, float* db, cudaStream_t stream) {
    NNB_RANGE("nnb_bn_backward_apply");
    NNB_REQUIRE(x && grad && mean && inv_std && sums, "nnb_bn_backward_apply: null pointer");
    NNB_REQUIRE(count > 0, "nnb_bn_backward_apply: bad count");
    int rc = check_shape(B, C, HW, "nnb_bn_backward_apply");
    if (rc) return rc;
    if (dx != nullptr) {
        const long long planes = (long long)B * C;
        const int blocks = (int)std::min<long long>(planes, (long long)num_sms() * 16);
        const int threads = HW >= 1024 ? 256 : (HW >= 256 ? 64 : 32);
        NNB_CUDA_OK(launch_pdl(bn_apply_kernel<1>, dim3(blocks), dim3(threads), 0, stream, x, grad, mean, inv_std, w,
                               (const float*)nullptr, sums, count, (int)C, (int)HW, planes, alpha, dx));
        count_launch();
    }
    if (dw != nullptr || db != nullptr) {
        NNB_CUDA_OK(launch_pdl(bn_param_grads_kernel, dim3((unsigned)ceil_div(C, 256)), dim3(256), 0, stream, sums, (int)C, dw, db));
        count_launch();
    }
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

}  // extern "C"
