// Measurement entry point for bench.py's roofline: GPU-paced time of ONE tcgen05 GEMM launch in each
// of the three nn.Linear forms, on operand sets that rotate through more memory than the L2 holds.
// The launches are captured once into a CUDA graph (so no host tensor-map encoding or launch latency
// sits between kernels) and the graph replay is timed with CUDA events on the launching stream.
// Diagnostics only: no product path calls this.
#include <vector>

#include "common.cuh"

namespace nnb {
namespace {

__global__ void fill_bf16_kernel(__nv_bfloat16* p, long long n, uint32_t seed) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        p[i] = __float2bfloat16_rn((float)(h & 0xFFFF) * (2.0f / 65535.0f) - 1.0f);  // U(-1, 1)
    }
}

struct Bufs {
    std::vector<void*> all;
    ~Bufs() { for (void* p : all) cudaFree(p); }
    void* get(size_t bytes) {
        void* p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
        all.push_back(p);
        return p;
    }
};

Staged make_staged(const __nv_bfloat16* p, int64_t rows, int64_t cols) {
    Staged s;
    s.hi = p; s.lo = nullptr; s.rows = rows; s.cols = cols; s.ld = staged_ld(cols);
    s.batch = 1; s.batch_stride = rows * s.ld;
    return s;
}

}  // namespace
}  // namespace nnb

using namespace nnb;

extern "C" int nnb_probe_linear_gemm(int64_t M, int64_t K, int64_t N, int form, int with_bias, int sets,
                                     int rounds, float* us_per_launch, int* launches_per_gemm,
                                     cudaStream_t stream) {
    NNB_REQUIRE(M > 0 && K > 0 && N > 0 && form >= 0 && form <= 2, "nnb_probe_linear_gemm: bad arguments");
    NNB_REQUIRE(sets > 0 && sets <= 64 && rounds > 0 && us_per_launch, "nnb_probe_linear_gemm: bad arguments");
    cudaStreamCaptureStatus cs;
    NNB_CUDA_OK(cudaStreamIsCapturing(stream, &cs));
    NNB_REQUIRE(cs == cudaStreamCaptureStatusNone, "nnb_probe_linear_gemm: stream is capturing");
    // the legacy default stream cannot be captured: run on a private stream, ordered after the caller's work
    NNB_CUDA_OK(cudaStreamSynchronize(stream));
    struct OwnStream {
        cudaStream_t s = nullptr;
        ~OwnStream() { if (s) cudaStreamDestroy(s); }
    } own;
    NNB_CUDA_OK(cudaStreamCreateWithFlags(&own.s, cudaStreamNonBlocking));
    stream = own.s;
    Bufs bufs;
    // fwd: A = X[M,K], B = W[N,K] -> D[M,N];  dgrad: A = G[M,N], B = W[N,K] (MN-major) -> D[M,K];
    // wgrad: A = G[M,N] (MN-major), B = X[M,K] (MN-major) -> D[N,K]   (linear.py:19-22, 54)
    const int64_t ar = (form == 0) ? M : M, ac = (form == 0) ? K : N;
    const int64_t br = (form == 2) ? M : N, bc = K;
    const int64_t dr = (form == 2) ? N : M, dc = (form == 0) ? N : K;
    std::vector<GemmProblem> probs((size_t)sets);
    const size_t skb = gemm_splitk_ws_bytes(dr, dc, (form == 0) ? K : (form == 1 ? N : M), 1);
    float* sk = static_cast<float*>(bufs.get(skb));
    float* bias = static_cast<float*>(bufs.get((size_t)dc * 4 + 256));
    NNB_REQUIRE(sk && bias, "nnb_probe_linear_gemm: out of device memory");
    NNB_CUDA_OK(cudaMemsetAsync(bias, 0, (size_t)dc * 4, stream));
    for (int s = 0; s < sets; ++s) {
        auto* a = static_cast<__nv_bfloat16*>(bufs.get(staged_plane_bytes(1, ar, ac)));
        auto* b = static_cast<__nv_bfloat16*>(bufs.get(staged_plane_bytes(1, br, bc)));
        auto* d = static_cast<float*>(bufs.get((size_t)dr * dc * 4));
        NNB_REQUIRE(a && b && d, "nnb_probe_linear_gemm: out of device memory");
        fill_bf16_kernel<<<1024, 256, 0, stream>>>(a, (long long)(staged_plane_bytes(1, ar, ac) / 2), 17u + s);
        fill_bf16_kernel<<<1024, 256, 0, stream>>>(b, (long long)(staged_plane_bytes(1, br, bc) / 2), 101u + s);
        GemmProblem& g = probs[(size_t)s];
        g.A.st = make_staged(a, ar, ac);
        g.B.st = make_staged(b, br, bc);
        if (with_bias & 4) {  // bf16x3: hi/lo planes, three products per k-block
            auto* al = static_cast<__nv_bfloat16*>(bufs.get(staged_plane_bytes(1, ar, ac)));
            auto* bl = static_cast<__nv_bfloat16*>(bufs.get(staged_plane_bytes(1, br, bc)));
            NNB_REQUIRE(al && bl, "nnb_probe_linear_gemm: out of device memory");
            fill_bf16_kernel<<<1024, 256, 0, stream>>>(al, (long long)(staged_plane_bytes(1, ar, ac) / 2), 23u + s);
            fill_bf16_kernel<<<1024, 256, 0, stream>>>(bl, (long long)(staged_plane_bytes(1, br, bc) / 2), 131u + s);
            g.A.st.lo = al;
            g.B.st.lo = bl;
        }
        if (form == 0) {
            g.M = M; g.N = N; g.K = K;
            if (with_bias & 1) g.epi.bias = bias;
            if (with_bias & 2) {  // Swish epilogue with the pre-activation side output (LinearSwish forward)
                g.epi.Z = static_cast<float*>(bufs.get((size_t)dr * dc * 4));
                NNB_REQUIRE(g.epi.Z, "nnb_probe_linear_gemm: out of device memory");
                g.epi.act = NNB_ACT_SWISH;
            }
        }
        if (form == 1) { g.M = M; g.N = K; g.K = N; g.B.mn_major = true; }
        if (form == 2) { g.M = N; g.N = K; g.K = M; g.A.mn_major = true; g.B.mn_major = true; }
        g.ldd = dc;
        g.D = d;
        g.splitk_ws = sk;
        g.splitk_ws_bytes = skb;
    }
    NNB_CUDA_OK(cudaGetLastError());
    const uint64_t before = nnb_launch_count();
    int rc = gemm(probs[0], stream);  // warm-up + launch count of one GEMM (split-K adds a reduce kernel)
    if (rc) return rc;
    if (launches_per_gemm) *launches_per_gemm = (int)(nnb_launch_count() - before);
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    NNB_CUDA_OK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
    for (int r = 0; r < rounds && !rc; ++r)
        for (int s = 0; s < sets && !rc; ++s) rc = gemm(probs[(size_t)s], stream);
    cudaError_t ce = cudaStreamEndCapture(stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    NNB_CUDA_OK(ce);
    NNB_CUDA_OK(cudaGraphInstantiate(&exec, graph, 0));
    cudaEvent_t e0, e1;
    NNB_CUDA_OK(cudaEventCreate(&e0));
    NNB_CUDA_OK(cudaEventCreate(&e1));
    NNB_CUDA_OK(cudaGraphLaunch(exec, stream));
    NNB_CUDA_OK(cudaEventRecord(e0, stream));
    NNB_CUDA_OK(cudaGraphLaunch(exec, stream));
    NNB_CUDA_OK(cudaEventRecord(e1, stream));
    NNB_CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    NNB_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    *us_per_launch = ms * 1e3f / (float)(sets * rounds);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaGraphExecDestroy(exec);
    cudaGraphDestroy(graph);
    NNB_CUDA_OK(cudaStreamSynchronize(stream));
    return NNB_OK;
}
