// Multi-tensor Adam / AdamW: ONE launch updates every parameter tensor.
// Replaces the per-tensor Python loops of neunet/optim.py:17-33 / 52-69 (about ten full-size NumPy
// temporaries per tensor) and the reference's FusedAdamWStep
// (experimental/optim/fused_adamw/fused_adamw_multitensor.cu:108-303).
// HBM-bound: per element 16 B read (p, g, m, v) + 12 B written (p, m, v), 16-byte accesses.
// The pointer tables and the block -> (tensor, chunk) map live on the device and are uploaded once
// at create time; only the gradient pointer table is refreshed when gradients move.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

struct nnb_adamw {
    int n = 0;
    int nblocks = 0;
    float** d_p = nullptr;
    const float** d_g = nullptr;
    float** d_m = nullptr;
    float** d_v = nullptr;
    long long* d_sizes = nullptr;
    int* d_blk_tensor = nullptr;
    int* d_blk_chunk = nullptr;
    std::vector<const float*> h_g;  // pageable on purpose: the driver stages such async copies at call
                                    // time, so the next eager step may overwrite it immediately
    void** d_stage = nullptr;       // per tensor: bf16 staging destination (hi plane) or null
    long long* d_stage_cols = nullptr;  // per tensor: columns of the 2-D weight view
    long long* d_stage_lo = nullptr;    // per tensor: element offset of the lo plane (0 = BF16 mode)
    bool staging = false;
    long long* d_t = nullptr;     // device-resident step counter (CUDA-graph replays advance it)
    std::vector<const float**> snapshots;  // pinned pointer tables owned by captured graphs
    std::vector<int> blk_start;   // [n + 1]: first block of every tensor (nnb_adamw_step_range)
};

namespace nnb {
namespace {

constexpr int ADAM_THREADS = 512;
constexpr int ADAM_CHUNK = ADAM_THREADS * 4 * 4;  // 8192 elements per block: 4 x float4 per thread

struct AdamScalars {
    float lr, b1, b2, one_m_b1, one_m_b2, bc1, bc2, eps, lr_wd, wd, gscale;
    double beta1, beta2;  // for the device-side bias correction
    int mode;
};

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamScalars& s) {
    g *= s.gscale;
    if (s.mode == NNB_OPT_ADAMW) {
        if (s.lr_wd != 0.f) p -= s.lr_wd * p;          // optim.py:59-60
    } else {
        if (s.wd != 0.f) g = g + s.wd * p;             // optim.py:24-25
    }
    m = s.b1 * m + s.one_m_b1 * g;                      // optim.py:27 / 63
    v = s.b2 * v + s.one_m_b2 * (g * g);                // optim.py:28 / 64
    const float m_hat = m / s.bc1;
    const float v_hat = v / s.bc2;
    p -= s.lr * m_hat / (sqrtf(v_hat) + s.eps);         // optim.py:33 / 69
}

// bf16 (hi [+ lo]) copy of an updated weight element range into the staged plane the next forward GEMM
// reads through TMA (common.cuh: Staged, ld = round_up(cols, 8)), so weights are never re-converted.
__device__ __forceinline__ void emit_staged(__nv_bfloat16* hi, long long lo_off, long long cols, long long i,
                                            const float* x, int n) {
    const long long ld = (cols + 7) & ~7ll;
    const long long row = i / cols, col = i - row * cols;
    if (n == 4 && (cols & 3) == 0) {
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h[j] = __float2bfloat16_rn(x[j]);
            l[j] = __float2bfloat16_rn(x[j] - __bfloat162float(h[j]));
        }
        const long long o = row * ld + col;
        *reinterpret_cast<uint2*>(hi + o) = *reinterpret_cast<const uint2*>(h);
        if (lo_off) *reinterpret_cast<uint2*>(hi + lo_off + o) = *reinterpret_cast<const uint2*>(l);
        return;
    }
    long long r = row, c = col;
    for (int j = 0; j < n; ++j) {
        const __nv_bfloat16 h = __float2bfloat16_rn(x[j]);
        hi[r * ld + c] = h;
        if (lo_off) hi[lo_off + r * ld + c] = __float2bfloat16_rn(x[j] - __bfloat162float(h));
        if (++c == cols) { c = 0; ++r; }
    }
}

__global__ void __launch_bounds__(ADAM_THREADS)
adamw_multi_kernel(float* const* __restrict__ P, const float* const* __restrict__ G,
                   float* const* __restrict__ Mm, float* const* __restrict__ V,
                   const long long* __restrict__ sizes, const int* __restrict__ blk_tensor,
                   const int* __restrict__ blk_chunk, AdamScalars s,
                   const long long* __restrict__ d_t, void* const* __restrict__ stage,
                   const long long* __restrict__ stage_cols, const long long* __restrict__ stage_lo) {
    if (d_t != nullptr) {
        // graph mode: the step count lives on the device, so a replayed graph keeps advancing the
        // bias corrections 1 - beta^t (formed in double, rounded once, like the host path)
        __shared__ float bc[2];
        if (threadIdx.x == 0) {
            const double tt = (double)(*d_t);
            bc[0] = (float)(1.0 - pow(s.beta1, tt));
            bc[1] = (float)(1.0 - pow(s.beta2, tt));
        }
        __syncthreads();
        s.bc1 = bc[0];
        s.bc2 = bc[1];
    }
    const int t = blk_tensor[blockIdx.x];
    const float* g = G[t];
    if (g == nullptr) return;  // `if param.grad is None: continue` (optim.py:21-22)
    float* p = P[t];
    float* m = Mm[t];
    float* v = V[t];
    const long long n = sizes[t];
    __nv_bfloat16* st_hi = stage ? static_cast<__nv_bfloat16*>(stage[t]) : nullptr;
    const long long st_cols = st_hi ? stage_cols[t] : 1, st_lo = st_hi ? stage_lo[t] : 0;
    const long long base = (long long)blk_chunk[blockIdx.x] * ADAM_CHUNK;
    const long long end = min(base + ADAM_CHUNK, n);
    const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                           reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (aligned) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const long long i = base + ((long long)it * ADAM_THREADS + threadIdx.x) * 4;
            if (i + 3 < end) {
                float4 pp = *reinterpret_cast<float4*>(p + i);
                const float4 gg = *reinterpret_cast<const float4*>(g + i);
                float4 mm = *reinterpret_cast<float4*>(m + i);
                float4 vv = *reinterpret_cast<float4*>(v + i);
                adam_update(pp.x, gg.x, mm.x, vv.x, s);
                adam_update(pp.y, gg.y, mm.y, vv.y, s);
                adam_update(pp.z, gg.z, mm.z, vv.z, s);
                adam_update(pp.w, gg.w, mm.w, vv.w, s);
                *reinterpret_cast<float4*>(p + i) = pp;
                *reinterpret_cast<float4*>(m + i) = mm;
                *reinterpret_cast<float4*>(v + i) = vv;
                if (st_hi) {
                    const float x[4] = {pp.x, pp.y, pp.z, pp.w};
                    emit_staged(st_hi, st_lo, st_cols, i, x, 4);
                }
            } else {
                for (long long j = i; j < end; ++j) {
                    adam_update(p[j], g[j], m[j], v[j], s);
                    if (st_hi) emit_staged(st_hi, st_lo, st_cols, j, p + j, 1);
                }
            }
        }
    } else {
        for (long long j = base + threadIdx.x; j < end; j += ADAM_THREADS) {
            adam_update(p[j], g[j], m[j], v[j], s);
            if (st_hi) emit_staged(st_hi, st_lo, st_cols, j, p + j, 1);
        }
    }
}

__global__ void adam_advance_step_kernel(long long* d_t) { *d_t += 1; }

}  // namespace
}  // namespace nnb

using namespace nnb;

extern "C" {

int nnb_adamw_create(nnb_adamw** out, int n, float* const* p, const float* const* g,
                     float* const* m, float* const* v, const int64_t* sizes, cudaStream_t stream) {
    NNB_REQUIRE(out && p && m && v && sizes && n > 0, "nnb_adamw_create: bad arguments");
    std::vector<int> bt, bc;
    std::vector<long long> sz(n);
    for (int i = 0; i < n; ++i) {
        NNB_REQUIRE(sizes[i] > 0 && p[i] && m[i] && v[i], "nnb_adamw_create: tensor %d invalid", i);
        sz[i] = sizes[i];
        const long long chunks = ceil_div(sizes[i], ADAM_CHUNK);
        for (long long c = 0; c < chunks; ++c) {
            bt.push_back(i);
            bc.push_back((int)c);
        }
    }
    nnb_adamw* o = new nnb_adamw();
    o->n = n;
    o->nblocks = (int)bt.size();
    o->blk_start.assign(n + 1, (int)bt.size());
    for (int b = (int)bt.size() - 1; b >= 0; --b) o->blk_start[bt[b]] = b;
    o->h_g.assign(n, nullptr);
    if (g) for (int i = 0; i < n; ++i) o->h_g[i] = g[i];
    NNB_CUDA_OK(cudaMalloc(&o->d_p, n * sizeof(float*)));
    NNB_CUDA_OK(cudaMalloc(&o->d_g, n * sizeof(float*)));
    NNB_CUDA_OK(cudaMalloc(&o->d_m, n * sizeof(float*)));
    NNB_CUDA_OK(cudaMalloc(&o->d_v, n * sizeof(float*)));
    NNB_CUDA_OK(cudaMalloc(&o->d_sizes, n * sizeof(long long)));
    NNB_CUDA_OK(cudaMalloc(&o->d_stage, n * sizeof(void*)));
    NNB_CUDA_OK(cudaMalloc(&o->d_stage_cols, n * sizeof(long long)));
    NNB_CUDA_OK(cudaMalloc(&o->d_stage_lo, n * sizeof(long long)));
    NNB_CUDA_OK(cudaMemsetAsync(o->d_stage, 0, n * sizeof(void*), stream));
    NNB_CUDA_OK(cudaMalloc(&o->d_t, sizeof(long long)));
    NNB_CUDA_OK(cudaMemsetAsync(o->d_t, 0, sizeof(long long), stream));
    NNB_CUDA_OK(cudaMalloc(&o->d_blk_tensor, bt.size() * sizeof(int)));
    NNB_CUDA_OK(cudaMalloc(&o->d_blk_chunk, bc.size() * sizeof(int)));
    NNB_CUDA_OK(cudaMemcpyAsync(o->d_p, p, n * sizeof(float*), cudaMemcpyHostToDevice, stream));
    NNB_CUDA_OK(cudaMemcpyAsync(o->d_g, o->h_g.data(), n * sizeof(float*), cudaMemcpyHostToDevice, stream));
    NNB_CUDA_OK(cudaMemcpyAsync(o->d_m, m, n * sizeof(float*), cudaMemcpyHostToDevice, stream));
    NNB_CUDA_OK(cudaMemcpyAsync(o->d_v, v, n * sizeof(float*), cudaMemcpyHostToDevice, stream));
    NNB_CUDA_OK(cudaMemcpyAsync(o->d_sizes, sz.data(), n * sizeof(long long), cudaMemcpyHostToDevice, stream));
    NNB_CUDA_OK(cudaMemcpyAsync(o->d_blk_tensor, bt.data(), bt.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
    NNB_CUDA_OK(cudaMemcpyAsync(o->d_blk_chunk, bc.data(), bc.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
    NNB_CUDA_OK(cudaStreamSynchronize(stream));  // host vectors go out of scope
    *out = o;
    return NNB_OK;
}

int nnb_adamw_set_grads(nnb_adamw* opt, const float* const* g, cudaStream_t stream) {
    NNB_REQUIRE(opt && g, "nnb_adamw_set_grads: bad arguments");
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    NNB_CUDA_OK(cudaStreamIsCapturing(stream, &cap));
    if (cap == cudaStreamCaptureStatusActive) {
        // A captured copy re-reads its pinned source on every replay, so it gets a private
        // snapshot of the table: later eager steps must not be able to change what the graph uploads.
        const float** snap = nullptr;
        // allocation calls are illegal under global capture mode: switch this thread to relaxed
        // mode just for the pinned allocation
        cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
        NNB_CUDA_OK(cudaThreadExchangeStreamCaptureMode(&mode));
        const cudaError_t ae = cudaHostAlloc((void**)&snap, opt->n * sizeof(float*), cudaHostAllocDefault);
        NNB_CUDA_OK(cudaThreadExchangeStreamCaptureMode(&mode));
        NNB_CUDA_OK(ae);
        for (int i = 0; i < opt->n; ++i) snap[i] = g[i];
        opt->snapshots.push_back(snap);
        NNB_CUDA_OK(cudaMemcpyAsync(opt->d_g, snap, opt->n * sizeof(float*), cudaMemcpyHostToDevice, stream));
        return NNB_OK;
    }
    // eager: always upload -- a graph replay in between may have rewritten the device table
    for (int i = 0; i < opt->n; ++i) opt->h_g[i] = g[i];
    NNB_CUDA_OK(cudaMemcpyAsync(opt->d_g, opt->h_g.data(), opt->n * sizeof(float*), cudaMemcpyHostToDevice, stream));
    return NNB_OK;
}

static int adamw_launch(nnb_adamw* opt, int first, int count, double lr, double beta1, double beta2, double eps,
                        double weight_decay, int64_t step, int mode, float grad_scale, bool advance, cudaStream_t stream) {
    NNB_REQUIRE(opt, "nnb_adamw_step: null handle");
    NNB_REQUIRE(step >= 0, "nnb_adamw_step: step must be >= 1, or 0 to use the device-resident counter");
    NNB_REQUIRE(mode == NNB_OPT_ADAM_L2 || mode == NNB_OPT_ADAMW, "nnb_adamw_step: bad mode");
    NNB_REQUIRE(first >= 0 && count > 0 && first + count <= opt->n, "nnb_adamw_step_range: bad tensor range");
    AdamScalars s;
    // Scalars are formed in double exactly where the reference forms them in Python floats, then
    // rounded once to fp32 (NumPy casts a Python float operand to the array dtype).
    s.lr = (float)lr;
    s.b1 = (float)beta1;
    s.b2 = (float)beta2;
    s.one_m_b1 = (float)(1.0 - beta1);
    s.one_m_b2 = (float)(1.0 - beta2);
    s.bc1 = (float)(1.0 - std::pow(beta1, (double)std::max<int64_t>(step, 1)));
    s.bc2 = (float)(1.0 - std::pow(beta2, (double)std::max<int64_t>(step, 1)));
    s.beta1 = beta1;
    s.beta2 = beta2;
    s.eps = (float)eps;
    s.lr_wd = (float)(lr * weight_decay);
    s.wd = (float)weight_decay;
    s.gscale = grad_scale;
    s.mode = mode;
    if (step == 0 && advance) {
        adam_advance_step_kernel<<<1, 1, 0, stream>>>(opt->d_t);
        count_launch();
    }
    const int b0 = opt->blk_start[first], b1 = opt->blk_start[first + count];
    if (b1 > b0) {
        adamw_multi_kernel<<<b1 - b0, ADAM_THREADS, 0, stream>>>(opt->d_p, opt->d_g, opt->d_m, opt->d_v, opt->d_sizes,
                                                                 opt->d_blk_tensor + b0, opt->d_blk_chunk + b0, s,
                                                                 step == 0 ? opt->d_t : nullptr,
                                                                 opt->staging ? opt->d_stage : nullptr,
                                                                 opt->d_stage_cols, opt->d_stage_lo);
        count_launch();
    }
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

int nnb_adamw_step(nnb_adamw* opt, double lr, double beta1, double beta2, double eps,
                   double weight_decay, int64_t step, int mode, float grad_scale,
                   cudaStream_t stream) {
    NNB_RANGE("nnb_adamw_step");
    NNB_REQUIRE(opt, "nnb_adamw_step: null handle");
    return adamw_launch(opt, 0, opt->n, lr, beta1, beta2, eps, weight_decay, step, mode, grad_scale, true, stream);
}

int nnb_adamw_step_range(nnb_adamw* opt, int first_tensor, int n_tensors, double lr, double beta1, double beta2, double eps,
                         double weight_decay, int64_t step, int mode, float grad_scale, int advance_counter,
                         cudaStream_t stream) {
    NNB_RANGE("nnb_adamw_step_range");
    return adamw_launch(opt, first_tensor, n_tensors, lr, beta1, beta2, eps, weight_decay, step, mode, grad_scale,
                        advance_counter != 0, stream);
}

int nnb_adamw_set_staging(nnb_adamw* opt, void* const* staged, const int64_t* cols, int prec,
                          cudaStream_t stream) {
    NNB_REQUIRE(opt, "nnb_adamw_set_staging: null handle");
    NNB_REQUIRE(prec == NNB_PREC_BF16 || prec == NNB_PREC_BF16X3, "nnb_adamw_set_staging: bad prec");
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    NNB_CUDA_OK(cudaStreamIsCapturing(stream, &cap));
    NNB_REQUIRE(cap == cudaStreamCaptureStatusNone, "nnb_adamw_set_staging: not allowed during stream capture");
    if (staged == nullptr) {
        opt->staging = false;
        return NNB_OK;
    }
    NNB_REQUIRE(cols, "nnb_adamw_set_staging: null cols");
    std::vector<long long> sizes(opt->n), hc(opt->n), hl(opt->n);
    NNB_CUDA_OK(cudaMemcpyAsync(sizes.data(), opt->d_sizes, opt->n * sizeof(long long), cudaMemcpyDeviceToHost, stream));
    NNB_CUDA_OK(cudaStreamSynchronize(stream));
    bool any = false;
    for (int i = 0; i < opt->n; ++i) {
        hc[i] = 1;
        hl[i] = 0;
        if (staged[i] == nullptr) continue;
        NNB_REQUIRE(cols[i] > 0 && sizes[i] % cols[i] == 0, "nnb_adamw_set_staging: tensor %d: cols must divide its size", i);
        NNB_REQUIRE((reinterpret_cast<uintptr_t>(staged[i]) & 255) == 0, "nnb_adamw_set_staging: tensor %d: buffer not 256-byte aligned", i);
        hc[i] = cols[i];
        if (prec == NNB_PREC_BF16X3) hl[i] = (long long)(staged_plane_bytes(1, sizes[i] / cols[i], cols[i]) / 2);
        any = true;
    }
    NNB_CUDA_OK(cudaMemcpyAsync(opt->d_stage, staged, opt->n * sizeof(void*), cudaMemcpyHostToDevice, stream));
    NNB_CUDA_OK(cudaMemcpyAsync(opt->d_stage_cols, hc.data(), opt->n * sizeof(long long), cudaMemcpyHostToDevice, stream));
    NNB_CUDA_OK(cudaMemcpyAsync(opt->d_stage_lo, hl.data(), opt->n * sizeof(long long), cudaMemcpyHostToDevice, stream));
    NNB_CUDA_OK(cudaStreamSynchronize(stream));
    opt->staging = any;
    return NNB_OK;
}

int nnb_adamw_set_step(nnb_adamw* opt, int64_t step, cudaStream_t stream) {
    NNB_REQUIRE(opt && step >= 0, "nnb_adamw_set_step: bad arguments");
    const long long v = step;
    NNB_CUDA_OK(cudaMemcpyAsync(opt->d_t, &v, sizeof(v), cudaMemcpyHostToDevice, stream));
    NNB_CUDA_OK(cudaStreamSynchronize(stream));
    return NNB_OK;
}

int nnb_adamw_destroy(nnb_adamw* opt) {
    if (!opt) return NNB_OK;
    cudaFree(opt->d_p); cudaFree((void*)opt->d_g); cudaFree(opt->d_m); cudaFree(opt->d_v);
    cudaFree(opt->d_sizes); cudaFree(opt->d_blk_tensor); cudaFree(opt->d_blk_chunk);
    cudaFree(opt->d_stage); cudaFree(opt->d_stage_cols); cudaFree(opt->d_stage_lo);
    for (auto* sn : opt->snapshots) cudaFreeHost((void*)sn);
    cudaFree(opt->d_t);
    delete opt;
    return NNB_OK;
}

}  // extern "C"
