// nn.Embedding on the device (row N4 of SURVEY.md 8f: "Embedding gather + last-write-wins scatter kernel").
// Reference: neunet/nn/layers/embedding.py:61-75 -- forward is `weight[ids]` through Tensor.__getitem__, whose backward
// ASSIGNS instead of accumulating (`full = zeros_like(w); full[ids] = grad`, neunet/autograd.py:909-910): with
// duplicate token ids NumPy keeps the LAST occurrence. A device scatter with duplicates is a race, so the order is made
// explicit: pass 1 records, per vocabulary row, the largest position that refers to it (atomicMax on the position index),
// pass 2 lets exactly that position copy its gradient row. Deterministic, and one read of `grad` + one write per
// touched row instead of the ~10 array-library kernels the composed form needs.
#include "common.cuh"

namespace nnb {
namespace {

__device__ __forceinline__ long long load_id(const void* ids, int is64, long long i) {
    return is64 ? static_cast<const long long*>(ids)[i] : (long long)static_cast<const int*>(ids)[i];
}

// out[i][:] = W[ids[i]][:]; negative ids wrap like NumPy; ids outside [-V, V) (IndexError in the reference) give NaN rows
__global__ void __launch_bounds__(256) embedding_fwd_kernel(const float* __restrict__ W, const void* __restrict__ ids, int is64,
                                                            long long n, long long V, int D, float* __restrict__ out, int vec) {
    pdl_trigger();
    pdl_wait();
    const int per_row = vec ? D >> 2 : D;
    const long long total = n * per_row;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx / per_row;
        const int c = (int)(idx - i * per_row);
        long long t = load_id(ids, is64, i);
        if (t < 0) t += V;
        const bool ok = t >= 0 && t < V;
        if (vec) {
            const float nanv = __int_as_float(0x7fc00000);
            const float4 v = ok ? __ldg(reinterpret_cast<const float4*>(W + t * D) + c) : make_float4(nanv, nanv, nanv, nanv);
            reinterpret_cast<float4*>(out + i * D)[c] = v;
        } else {
            out[i * D + c] = ok ? W[t * D + c] : __int_as_float(0x7fc00000);
        }
    }
}

__global__ void embedding_last_kernel(const void* __restrict__ ids, int is64, long long n, long long V, int* __restrict__ last) {
    pdl_trigger();
    pdl_wait();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long t = load_id(ids, is64, i);
        if (t < 0) t += V;
        if (t >= 0 && t < V) atomicMax(&last[t], (int)i);
    }
}

// one warp per position: the winner of its row copies grad[i][:] -> dW[ids[i]][:]
__global__ void __launch_bounds__(256) embedding_scatter_kernel(const void* __restrict__ ids, int is64, const float* __restrict__ grad,
                                                                long long n, long long V, int D, const int* __restrict__ last,
                                                                float* __restrict__ dW) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += ((long long)gridDim.x * blockDim.x) >> 5) {
        long long t = load_id(ids, is64, i);
        if (t < 0) t += V;
        if (t < 0 || t >= V || last[t] != (int)i) continue;
        const float* g = grad + i * D;
        float* d = dW + t * D;
        if ((D & 3) == 0 && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
            for (int c = lane; c < (D >> 2); c += 32) reinterpret_cast<float4*>(d)[c] = reinterpret_cast<const float4*>(g)[c];
        } else {
            for (int c = lane; c < D; c += 32) d[c] = g[c];
        }
    }
}

}  // namespace
}  // namespace nnb

using namespace nnb;

extern "C" {

int nnb_embedding_forward(const float* W, const void* ids, int ids_are_int64, int64_t n, int64_t V, int64_t D, float* out,
                          cudaStream_t stream) {
    NNB_RANGE("nnb_embedding_forward");
    NNB_REQUIRE(W && ids && out, "nnb_embedding_forward: null pointer");
    NNB_REQUIRE(n > 0 && V > 0 && D > 0 && D < (1ll << 31), "nnb_embedding_forward: bad shape");
    const int vec = (D % 4) == 0 && ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const long long total = n * (vec ? D / 4 : D);
    const int blocks = (int)std::max<long long>(1, std::min<long long>(ceil_div(total, 256), (long long)num_sms() * 16));
    NNB_CUDA_OK(launch_pdl(embedding_fwd_kernel, dim3(blocks), dim3(256), 0, stream, W, ids, ids_are_int64, (long long)n, (long long)V,
                           (int)D, out, vec));
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

size_t nnb_embedding_workspace_bytes(int64_t V) { return V > 0 ? (size_t)round_up(V * 4, 256) + 256 : 0; }

int nnb_embedding_backward(const void* ids, int ids_are_int64, const float* grad, int64_t n, int64_t V, int64_t D, float* dW,
                           void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    NNB_RANGE("nnb_embedding_backward");
    NNB_REQUIRE(ids && grad && dW, "nnb_embedding_backward: null pointer");
    NNB_REQUIRE(n > 0 && n < (1ll << 31) && V > 0 && D > 0 && D < (1ll << 31), "nnb_embedding_backward: bad shape");
    NNB_REQUIRE(workspace && workspace_bytes >= nnb_embedding_workspace_bytes(V), "nnb_embedding_backward: workspace too small");
    int* last = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    NNB_CUDA_OK(cudaMemsetAsync(last, 0xFF, (size_t)V * 4, stream));          // -1: row not referenced
    NNB_CUDA_OK(cudaMemsetAsync(dW, 0, (size_t)V * D * 4, stream));           // rows nobody refers to get a zero gradient
    const int b1 = (int)std::max<long long>(1, std::min<long long>(ceil_div(n, 256), (long long)num_sms() * 8));
    NNB_CUDA_OK(launch_pdl(embedding_last_kernel, dim3(b1), dim3(256), 0, stream, ids, ids_are_int64, (long long)n, (long long)V, last));
    const int b2 = (int)std::max<long long>(1, std::min<long long>(ceil_div(n * 32, 256), (long long)num_sms() * 16));
    NNB_CUDA_OK(launch_pdl(embedding_scatter_kernel, dim3(b2), dim3(256), 0, stream, ids, ids_are_int64, grad, (long long)n, (long long)V,
                           (int)D, (const int*)last, dW));
    count_launch(2);
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

}  // extern "C"
