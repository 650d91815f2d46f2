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

// grid = (C, chunks): block (c, j) reduces images b = j, j + chunks, ... of channel c. The block's threads walk the
// (image, pixel) pairs of those planes as ONE flat index, so a 4 x 4 plane of a deep UNet level keeps four lanes busy per
// image and 64 images fill the block -- a block per plane left 252 of 256 threads idle there (ncu launch list, round 2:
// 91 us for the 4 MB of a C = 1024 level against 15 us for the 33 MB of the C = 128 level). `chunks` shrinks with the
// plane so every block still has ~2 K elements (chunks_for).
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
    const int nimg = (B - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y;  // images of this block
    const long long total = (long long)nimg * HW;
    const long long img_stride = (long long)gridDim.y * C * HW, base = ((long long)blockIdx.y * C + c) * HW;
    if (vec) {
        for (long long e = (long long)threadIdx.x * 4; e < total; e += 1024) {
            const long long k = e / HW;
            const long long off = base + k * img_stride + (e - k * HW);
            const float4 v = *reinterpret_cast<const float4*>(x + off);
            const float a[4] = {lrelu(v.x, alpha), lrelu(v.y, alpha), lrelu(v.z, alpha), lrelu(v.w, alpha)};
            float t0 = 0.f, t1 = 0.f;  // fp32 over the four elements, fp64 across
            if (MODE == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { t0 += a[j]; t1 += a[j] * a[j]; }
            } else {
                const float4 gv = *reinterpret_cast<const float4*>(g + off);
                const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) { t0 += gg[j]; t1 += gg[j] * ((a[j] - mu) * is); }
            }
            s0 += (double)t0;
            s1 += (double)t1;
        }
    } else {
        for (long long e = threadIdx.x; e < total; e += 256) {
            const long long k = e / HW;
            const long long off = base + k * img_stride + (e - k * HW);
            const float a = lrelu(x[off], alpha);
            if (MODE == 0) { s0 += (double)a; s1 += (double)(a * a); }
            else { s0 += (double)g[off]; s1 += (double)(g[off] * ((a - mu) * is)); }
        }
    }
    // fixed-order block reduction: shuffles inside the warp, then the eight warp totals in warp order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_down_sync(0xffffffffu, s0, o);
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
    }
    __shared__ double r0[8], r1[8];
    if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = s0; r1[threadIdx.x >> 5] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { a0 += r0[w]; a1 += r1[w]; }
        partial[((long long)blockIdx.y * C + c) * 2] = a0;
        partial[((long long)blockIdx.y * C + c) * 2 + 1] = a1;
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
                                                       long long planes, float alpha, float* __restrict__ out, int lpp) {
    pdl_trigger();
    pdl_wait();
    const bool vec = (HW % 4) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                                        reinterpret_cast<uintptr_t>(g)) & 15) == 0;
    // `lpp` lanes (a power of two chosen by the host: ~ HW / 4, at most the block) share one (b, c) image plane, so a block
    // works on blockDim / lpp planes at a time: the 4 x 4 and 8 x 8 planes of the deep UNet levels fill the block too
    const int sub = threadIdx.x / lpp, lane = threadIdx.x - sub * lpp, ppb = blockDim.x / lpp;
    const double inv_count = 1.0 / count;
    for (long long pl = (long long)blockIdx.x * ppb + sub; pl < planes; pl += (long long)gridDim.x * ppb) {
        const int c = (int)(pl % C);
        const float mu = mean[c], is = inv[c], wc = w ? w[c] : 1.f, bc = b ? b[c] : 0.f;
        float k1 = 0.f, k2 = 0.f;
        if (MODE == 1) {
            k1 = (float)(sums[2 * c] * inv_count) * wc;       // mean(w * g)
            k2 = (float)(sums[2 * c + 1] * inv_count) * wc;   // mean(w * g * xhat)
        }
        const float* xp = x + pl * HW;
        const float* gp = MODE == 1 ? g + pl * HW : nullptr;
        float* op = out + pl * HW;
        if (vec) {
            for (int i = lane * 4; i < HW; i += lpp * 4) {
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
            for (int i = lane; i < HW; i += lpp) {
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

// partial rows per channel: every block gets >= ~2 K elements, at most BN_CHUNKS (and at most B) rows
int chunks_for(int64_t B, int64_t HW) {
    return (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(BN_CHUNKS, B), B * HW / 2048));
}

// lanes that share one image plane in bn_apply_kernel (power of two, <= 256) and the grid that goes with them
int apply_lanes(int64_t HW) {
    const int64_t want = (HW % 4 == 0) ? HW / 4 : HW;
    int lpp = 1;
    while (lpp < 256 && lpp < want) lpp <<= 1;
    return lpp;
}
int apply_blocks(long long planes, int lpp) {
    const long long ppb = 256 / lpp;
    return (int)std::max<long long>(1, std::min<long long>((planes + ppb - 1) / ppb, (long long)num_sms() * 8));
}

template <int MODE>
int run_reduce(const float* x, const float* g, const float* mean, const float* inv, int64_t B, int64_t C, int64_t HW,
               float alpha, double* sums, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    const int chunks = chunks_for(B, HW);
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
    const int lpp = apply_lanes(HW);
    NNB_CUDA_OK(launch_pdl(bn_apply_kernel<0>, dim3(apply_blocks(planes, lpp)), dim3(256), 0, stream, x, (const float*)nullptr, mean,
                           inv_std, w, b, (const double*)nullptr, 1.0, (int)C, (int)HW, planes, alpha, y, lpp));
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
        const int lpp = apply_lanes(HW);
        NNB_CUDA_OK(launch_pdl(bn_apply_kernel<1>, dim3(apply_blocks(planes, lpp)), dim3(256), 0, stream, x, grad, mean, inv_std, w,
                               (const float*)nullptr, sums, count, (int)C, (int)HW, planes, alpha, dx, lpp));
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
