// Philox4x32-10 counter-based RNG shared by every kernel that applies a dropout mask
// (dropout.cu, fused.cu, attention.cu). The mask of element e of a flat tensor is word (e & 3) of
// philox(counter = {e >> 2, call_id, epoch}, key = seed): any kernel that knows (seed, call_id, epoch)
// and the flat index regenerates the same bits, so masks are never stored
// (reference semantics: neunet/nn/layers/dropout.py:17-46 keeps the mask array instead).
#pragma once
#include <cstdint>

namespace nnb {

struct DropArgs {
    uint32_t thresh;  // keep iff word >= thresh  (thresh = p * 2^32)
    float scale;      // 1 / (1 - p)
    uint64_t seed;
    uint32_t call_id;
    uint64_t epoch_host;
    const unsigned long long* epoch_dev;  // overrides epoch_host when non-null (CUDA-graph replays)
};

static inline DropArgs make_drop_args(float p, uint64_t seed, uint32_t call_id, uint64_t epoch, const uint64_t* epoch_dev) {
    DropArgs d;
    const double t = (double)p * 4294967296.0;
    d.thresh = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
    d.scale = (float)(1.0 / (1.0 - (double)p));
    d.seed = seed;
    d.call_id = call_id;
    d.epoch_host = epoch;
    d.epoch_dev = reinterpret_cast<const unsigned long long*>(epoch_dev);
    return d;
}

#if defined(__CUDACC__)
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

// the four mask words of elements 4*group .. 4*group+3
__device__ __forceinline__ void drop_words(const DropArgs& d, uint64_t epoch, long long group, uint32_t (&c)[4]) {
    c[0] = (uint32_t)group;
    c[1] = (uint32_t)((uint64_t)group >> 32) ^ (d.call_id * 0x9E3779B1u);
    c[2] = (uint32_t)epoch;
    c[3] = (uint32_t)(epoch >> 32);
    philox4x32_10(c, (uint32_t)d.seed, (uint32_t)(d.seed >> 32));
}

__device__ __forceinline__ uint64_t drop_epoch(const DropArgs& d) {
    return d.epoch_dev ? (uint64_t)*d.epoch_dev : d.epoch_host;
}
#endif

}  // namespace nnb
