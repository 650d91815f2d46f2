// Dropout with a counter-based device RNG (row N4 of SURVEY.md 8f: "Dropout device RNG").
//
// Reference semantics (neunet/nn/layers/dropout.py:17-46): mask ~ Bernoulli(1 - p) / (1 - p),
// y = x * mask, dx = grad * mask with the SAME mask. The reference draws the mask with
// xp.random.binomial and keeps it as a full fp32 array; here the mask is never stored: it is a pure
// function of (seed, call_id, epoch, element index) through Philox4x32-10, so forward and backward
// regenerate identical bits and the op is one HBM pass each way (4 B read + 4 B written per element).
//
// CUDA-graph replays must not repeat masks, so the epoch may live in device memory (`epoch_dev`):
// the graph bakes the pointer, nnb_rng_advance() bumps the value between replays.
#include "common.cuh"
#include "philox.cuh"

namespace nnb {
namespace {

__device__ __forceinline__ void split2(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// Swish (neunet/nn/activations.py:208-211) with the fast exponential / reciprocal (~2 ulp): with four sigmoids and one
// Philox block per 16 bytes this pass is issue-bound, not HBM-bound
__device__ __forceinline__ float swish_of(float z, float beta) { return z * __fdividef(1.0f, 1.0f + __expf(-beta * z)); }

// y = (residual +) dropout(x) -- x = swish(z) when SWISH (the FFN of examples/gpt.ipynb cell 4: fc_2(dropout(swish(fc_1 x)))
// reads the pre-activation once and never materialises the activation); optionally also the bf16 planes of y (hi [+ lo])
// for the next Linear, so the consumer needs no staging pass. keep iff word >= p * 2^32  (P[keep] = 1 - p to within 2^-32).
template <bool PLANES, bool SWISH = false>
__global__ void __launch_bounds__(256) dropout_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                      float* __restrict__ y, long long n, const DropArgs d,
                                                      __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                      int vec, float beta = 1.0f) {
    pdl_trigger();
    pdl_wait();
    const uint64_t epoch = drop_epoch(d);
    const long long nvec = (n + 3) >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
         i += (long long)gridDim.x * blockDim.x) {
        uint32_t c[4];
        drop_words(d, epoch, i, c);
        const long long e = i << 2;
        if (vec && e + 3 < n) {
            float4 v = *reinterpret_cast<const float4*>(x + e);
            if (SWISH) { v.x = swish_of(v.x, beta); v.y = swish_of(v.y, beta); v.z = swish_of(v.z, beta); v.w = swish_of(v.w, beta); }
            v.x = c[0] >= d.thresh ? v.x * d.scale : 0.f;
            v.y = c[1] >= d.thresh ? v.y * d.scale : 0.f;
            v.z = c[2] >= d.thresh ? v.z * d.scale : 0.f;
            v.w = c[3] >= d.thresh ? v.w * d.scale : 0.f;
            if (res != nullptr) {
                const float4 r = *reinterpret_cast<const float4*>(res + e);
                v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
            }
            *reinterpret_cast<float4*>(y + e) = v;
            if (PLANES) {
                const float xv[4] = {v.x, v.y, v.z, v.w};
                __nv_bfloat16 h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) split2(xv[j], h[j], l[j]);
                *reinterpret_cast<uint2*>(hi + e) = *reinterpret_cast<const uint2*>(h);
                if (lo != nullptr) *reinterpret_cast<uint2*>(lo + e) = *reinterpret_cast<const uint2*>(l);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (e + j < n) {
                    const float xv = SWISH ? swish_of(x[e + j], beta) : x[e + j];
                    float v = c[j] >= d.thresh ? xv * d.scale : 0.f;
                    if (res != nullptr) v += res[e + j];
                    y[e + j] = v;
                    if (PLANES) {
                        __nv_bfloat16 h, l;
                        split2(v, h, l);
                        hi[e + j] = h;
                        if (lo != nullptr) lo[e + j] = l;
                    }
                }
        }
    }
}

__global__ void rng_advance_kernel(unsigned long long* epoch) { *epoch += 1ull; }

}  // namespace
}  // namespace nnb

using namespace nnb;

extern "C" {

int nnb_dropout_fused(const float* x, const float* residual, float* y, int64_t rows, int64_t cols, float p,
                      uint64_t seed, uint32_t call_id, uint64_t epoch, const uint64_t* epoch_dev,
                      void* Y_staged_out, int prec, cudaStream_t stream) {
    NNB_RANGE("nnb_dropout_fused");
    NNB_REQUIRE(x && y, "nnb_dropout: null pointer");
    NNB_REQUIRE(rows > 0 && cols > 0, "nnb_dropout: bad size");
    NNB_REQUIRE(p >= 0.f && p < 1.f, "nnb_dropout: p must be in [0, 1)");
    const int64_t n = rows * cols;
    const DropArgs d = make_drop_args(p, seed, call_id, epoch, epoch_dev);
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const int vec = al16(x) && al16(y) && (residual == nullptr || al16(residual));
    const long long nvec = (n + 3) / 4;
    const int blocks = (int)std::max<long long>(1, std::min<long long>(ceil_div(nvec, 256), (long long)num_sms() * 16));
    __nv_bfloat16 *hi = nullptr, *lo = nullptr;
    if (Y_staged_out != nullptr) {
        // planes are [rows][ld = round_up(cols, 8)]: only a pitch equal to cols lets the flat kernel write them
        NNB_REQUIRE(cols % 8 == 0, "nnb_dropout: staged output needs cols % 8 == 0");
        NNB_REQUIRE((reinterpret_cast<uintptr_t>(Y_staged_out) & 255) == 0, "nnb_dropout: Y_staged_out must be 256-byte aligned");
        NNB_REQUIRE(prec == NNB_PREC_BF16 || prec == NNB_PREC_BF16X3, "nnb_dropout: bad prec");
        hi = static_cast<__nv_bfloat16*>(Y_staged_out);
        if (prec == NNB_PREC_BF16X3)
            lo = reinterpret_cast<__nv_bfloat16*>(static_cast<uint8_t*>(Y_staged_out) + staged_plane_bytes(1, rows, cols));
        NNB_CUDA_OK(launch_pdl(dropout_kernel<true, false>, dim3(blocks), dim3(256), 0, stream, x, residual, y, (long long)n, d, hi, lo, vec, 1.0f));
    } else {
        NNB_CUDA_OK(launch_pdl(dropout_kernel<false, false>, dim3(blocks), dim3(256), 0, stream, x, residual, y, (long long)n, d, hi, lo, vec, 1.0f));
    }
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

int nnb_swish_dropout_fused(const float* z, float beta, float* y, int64_t rows, int64_t cols, float p, uint64_t seed,
                            uint32_t call_id, uint64_t epoch, const uint64_t* epoch_dev, void* Y_staged_out, int prec,
                            cudaStream_t stream) {
    NNB_RANGE("nnb_swish_dropout_fused");
    NNB_REQUIRE(z && y, "nnb_swish_dropout: null pointer");
    NNB_REQUIRE(rows > 0 && cols > 0, "nnb_swish_dropout: bad size");
    NNB_REQUIRE(p >= 0.f && p < 1.f, "nnb_swish_dropout: p must be in [0, 1)");
    const int64_t n = rows * cols;
    const DropArgs d = make_drop_args(p, seed, call_id, epoch, epoch_dev);
    const int vec = ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    const long long nvec = (n + 3) / 4;
    const int blocks = (int)std::max<long long>(1, std::min<long long>(ceil_div(nvec, 256), (long long)num_sms() * 16));
    __nv_bfloat16 *hi = nullptr, *lo = nullptr;
    const float* none = nullptr;
    if (Y_staged_out != nullptr) {
        NNB_REQUIRE(cols % 8 == 0, "nnb_swish_dropout: staged output needs cols % 8 == 0");
        NNB_REQUIRE((reinterpret_cast<uintptr_t>(Y_staged_out) & 255) == 0, "nnb_swish_dropout: Y_staged_out must be 256-byte aligned");
        NNB_REQUIRE(prec == NNB_PREC_BF16 || prec == NNB_PREC_BF16X3, "nnb_swish_dropout: bad prec");
        hi = static_cast<__nv_bfloat16*>(Y_staged_out);
        if (prec == NNB_PREC_BF16X3)
            lo = reinterpret_cast<__nv_bfloat16*>(static_cast<uint8_t*>(Y_staged_out) + staged_plane_bytes(1, rows, cols));
        NNB_CUDA_OK(launch_pdl(dropout_kernel<true, true>, dim3(blocks), dim3(256), 0, stream, z, none, y, (long long)n, d, hi, lo, vec, beta));
    } else {
        NNB_CUDA_OK(launch_pdl(dropout_kernel<false, true>, dim3(blocks), dim3(256), 0, stream, z, none, y, (long long)n, d, hi, lo, vec, beta));
    }
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

int nnb_dropout(const float* x, float* y, int64_t n, float p, uint64_t seed, uint32_t call_id,
                uint64_t epoch, const uint64_t* epoch_dev, cudaStream_t stream) {
    NNB_REQUIRE(n > 0, "nnb_dropout: bad size");
    return nnb_dropout_fused(x, nullptr, y, 1, n, p, seed, call_id, epoch, epoch_dev, nullptr, NNB_PREC_BF16, stream);
}

int nnb_rng_advance(uint64_t* epoch_dev, cudaStream_t stream) {
    NNB_REQUIRE(epoch_dev, "nnb_rng_advance: null pointer");
    rng_advance_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<unsigned long long*>(epoch_dev));
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

}  // extern "C"
