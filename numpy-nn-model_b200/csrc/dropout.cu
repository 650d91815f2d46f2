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

namespace nnb {
namespace {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// keep iff u32 >= p * 2^32  (P[keep] = 1 - p to within 2^-32)
__global__ void __launch_bounds__(256) dropout_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                      long long n, uint32_t thresh, float scale,
                                                      uint64_t seed, uint32_t call_id, uint64_t epoch_host,
                                                      const unsigned long long* __restrict__ epoch_dev, int vec) {
    pdl_trigger();
    pdl_wait();
    const uint64_t epoch = epoch_dev ? (uint64_t)*epoch_dev : epoch_host;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const long long nvec = (n + 3) >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
         i += (long long)gridDim.x * blockDim.x) {
        uint32_t c[4] = {(uint32_t)i, (uint32_t)((uint64_t)i >> 32) ^ (call_id * 0x9E3779B1u), (uint32_t)epoch,
                         (uint32_t)(epoch >> 32)};
        philox4x32_10(c, k0, k1);
        const long long e = i << 2;
        if (vec && e + 3 < n) {
            float4 v = *reinterpret_cast<const float4*>(x + e);
            v.x = c[0] >= thresh ? v.x * scale : 0.f;
            v.y = c[1] >= thresh ? v.y * scale : 0.f;
            v.z = c[2] >= thresh ? v.z * scale : 0.f;
            v.w = c[3] >= thresh ? v.w * scale : 0.f;
            *reinterpret_cast<float4*>(y + e) = v;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (e + j < n) y[e + j] = c[j] >= thresh ? x[e + j] * scale : 0.f;
        }
    }
}

__global__ void rng_advance_kernel(unsigned long long* epoch) { *epoch += 1ull; }

}  // namespace
}  // namespace nnb

using namespace nnb;

extern "C" {

int nnb_dropout(const float* x, float* y, int64_t n, float p, uint64_t seed, uint32_t call_id,
                uint64_t epoch, const uint64_t* epoch_dev, cudaStream_t stream) {
    NNB_REQUIRE(x && y, "nnb_dropout: null pointer");
    NNB_REQUIRE(n > 0, "nnb_dropout: bad size");
    NNB_REQUIRE(p >= 0.f && p < 1.f, "nnb_dropout: p must be in [0, 1)");
    const double t = (double)p * 4294967296.0;
    const uint32_t thresh = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
    const float scale = (float)(1.0 / (1.0 - (double)p));
    const int vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    const long long nvec = (n + 3) / 4;
    const int blocks = (int)std::max<long long>(1, std::min<long long>(ceil_div(nvec, 256), (long long)num_sms() * 16));
    NNB_CUDA_OK(launch_pdl(dropout_kernel, dim3(blocks), dim3(256), 0, stream, x, y, (long long)n, thresh, scale, seed,
                           call_id, epoch, reinterpret_cast<const unsigned long long*>(epoch_dev), vec));
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

int nnb_rng_advance(uint64_t* epoch_dev, cudaStream_t stream) {
    NNB_REQUIRE(epoch_dev, "nnb_rng_advance: null pointer");
    rng_advance_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<unsigned long long*>(epoch_dev));
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

}  // extern "C"
