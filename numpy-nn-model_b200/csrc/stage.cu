// Operand staging: strided fp32 views -> bf16 (hi [+ lo]) planes that TMA can tile.
// HBM-bound streaming kernels: 16-byte loads, 8-byte packed bf16 stores, optional fused
// Swish-backward transform (dZ = dO * swish'(Z), neunet/nn/activations.py:212-216) and fused
// column sums (bias gradient, neunet/nn/layers/linear.py:23-24) so `dO` is read from HBM once.
#include <algorithm>

#include "common.cuh"
#include "philox.cuh"

namespace nnb {
namespace {

__device__ __forceinline__ float swish_grad(float z, float beta) {
    // d/dz [z * sigmoid(beta z)] = beta*f + sigmoid(beta z) * (1 - beta*f)
    const float s = __fdividef(1.0f, 1.0f + __expf(-beta * z));  // MUFU.EX2 + MUFU.RCP: this pass is issue-bound, not HBM-bound
    const float f = z * s;
    return beta * f + s * (1.0f - beta * f);
}

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

struct StageArgs {
    View4 v;
    const float* aux;
    float beta;
    __nv_bfloat16* hi;
    __nv_bfloat16* lo;
    long long ld, plane_stride;  // elements
    float* partial;              // [gridDim.z * gridDim.y][cols] column partial sums, or null
    int rows_per_block;
    int vec_ok;
    int tx;  // threads along columns (power of two <= 128)
    DropArgs drop;  // DROP: dropout mask applied to the source first (flat index row * cols + col)
};

// Block = 128 threads arranged as TX (columns, 4 elements each) x TY (rows): wide matrices use
// TX = 128, narrow ones (e.g. the 4096 x 10 logits gradient) fold rows into the block so no lane
// idles. blockIdx.y owns a chunk of rows, blockIdx.z = batch. Column partial sums (bias gradient)
// are reduced over TY in shared memory and written once per block.
template <bool X3, int OP, bool DROP = false>
__global__ void __launch_bounds__(512) stage_rows_kernel(const StageArgs a) {
    pdl_trigger();
    pdl_wait();  // launched with the PDL attribute: nothing above touches global memory
    const uint64_t epoch = DROP ? drop_epoch(a.drop) : 0;
    const int TX = a.tx, TY = (int)blockDim.x / a.tx;
    const int tx = threadIdx.x & (TX - 1), ty = threadIdx.x / TX;
    const long long c = ((long long)blockIdx.x * TX + tx) * 4;
    const int bz = blockIdx.z;
    const int i0 = bz / (int)a.v.b1, i1 = bz - i0 * (int)a.v.b1;
    const float* src = a.v.ptr + i0 * a.v.s_b0 + i1 * a.v.s_b1;
    const float* aux = (OP == STAGE_SWISH_BWD) ? a.aux + i0 * a.v.s_b0 + i1 * a.v.s_b1 : nullptr;
    const long long r0 = (long long)blockIdx.y * a.rows_per_block;
    const long long r1 = min(r0 + a.rows_per_block, (long long)a.v.rows);
    float cs[4] = {0.f, 0.f, 0.f, 0.f};
    if (OP == STAGE_COPY && c + 3 < a.v.cols && a.vec_ok) {
        // hot path (contiguous rows, plain conversion): four rows per trip with all loads issued before
        // the first store, so each thread keeps 64 B in flight instead of 16 B
        for (long long r = r0 + ty; r < r1; r += 4 * TY) {
            float4 t[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long rr = r + (long long)u * TY;
                t[u] = rr < r1 ? __ldg(reinterpret_cast<const float4*>(src + rr * a.v.s_r + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long rr = r + (long long)u * TY;
                if (rr >= r1) break;
                float x[4] = {t[u].x, t[u].y, t[u].z, t[u].w};
                if (DROP) {
                    uint32_t w[4];
                    drop_words(a.drop, epoch, (rr * a.v.cols + c) >> 2, w);
#pragma unroll
                    for (int j = 0; j < 4; ++j) x[j] = w[j] >= a.drop.thresh ? x[j] * a.drop.scale : 0.f;
                }
                __nv_bfloat16 h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    split_bf16(x[j], h[j], l[j]);
                    cs[j] += x[j];
                }
                const long long o = (long long)bz * a.plane_stride + rr * a.ld + c;
                *reinterpret_cast<uint2*>(a.hi + o) = *reinterpret_cast<const uint2*>(h);
                if (X3) *reinterpret_cast<uint2*>(a.lo + o) = *reinterpret_cast<const uint2*>(l);
            }
        }
    } else if (c < a.v.cols) {
        const bool full = c + 3 < a.v.cols;
        for (long long r = r0 + ty; r < r1; r += TY) {
            float x[4] = {0.f, 0.f, 0.f, 0.f};
            if (full && a.vec_ok) {
                const float4 t = *reinterpret_cast<const float4*>(src + r * a.v.s_r + c);
                x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
                if (DROP) {
                    uint32_t w[4];
                    drop_words(a.drop, epoch, (r * a.v.cols + c) >> 2, w);
#pragma unroll
                    for (int j = 0; j < 4; ++j) x[j] = w[j] >= a.drop.thresh ? x[j] * a.drop.scale : 0.f;
                }
                if (OP == STAGE_SWISH_BWD) {
                    const float4 z = *reinterpret_cast<const float4*>(aux + r * a.v.s_r + c);
                    x[0] *= swish_grad(z.x, a.beta); x[1] *= swish_grad(z.y, a.beta);
                    x[2] *= swish_grad(z.z, a.beta); x[3] *= swish_grad(z.w, a.beta);
                }
            } else {
                uint32_t w[4] = {0u, 0u, 0u, 0u};
                if (DROP) drop_words(a.drop, epoch, (r * a.v.cols + c) >> 2, w);  // cols % 4 == 0: c .. c + 3 share a group
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (c + j < a.v.cols) {
                        const long long off = r * a.v.s_r + (c + j) * a.v.s_c;
                        x[j] = src[off];
                        if (DROP) x[j] = w[j] >= a.drop.thresh ? x[j] * a.drop.scale : 0.f;
                        if (OP == STAGE_SWISH_BWD) x[j] *= swish_grad(aux[off], a.beta);
                    }
                }
            }
            __nv_bfloat16 h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                split_bf16(x[j], h[j], l[j]);
                cs[j] += x[j];
            }
            const long long o = (long long)bz * a.plane_stride + r * a.ld + c;
            if (full) {
                *reinterpret_cast<uint2*>(a.hi + o) = *reinterpret_cast<const uint2*>(h);
                if (X3) *reinterpret_cast<uint2*>(a.lo + o) = *reinterpret_cast<const uint2*>(l);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (c + j < a.v.cols) {
                        a.hi[o + j] = h[j];
                        if (X3) a.lo[o + j] = l[j];
                    }
                }
            }
        }
    }
    if (a.partial != nullptr) {
        __shared__ float red[512 * 4];
#pragma unroll
        for (int j = 0; j < 4; ++j) red[(ty * TX + tx) * 4 + j] = cs[j];
        __syncthreads();
        if (ty == 0 && c < a.v.cols) {
            float* prow = a.partial + ((long long)blockIdx.z * gridDim.y + blockIdx.y) * a.v.cols;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (c + j < a.v.cols) {
                    float sum = 0.f;
                    for (int y = 0; y < TY; ++y) sum += red[(y * TX + tx) * 4 + j];
                    prow[c + j] = sum;
                }
            }
        }
    }
}

// out[c] = sum_p partial[p][c]: 32 columns x 32 partial-lanes per block, fixed-order tree at the end
// (deterministic): <= 1024 partial rows are <= 32 independent loads per thread.
__global__ void __launch_bounds__(1024) colsum_reduce_kernel(const float* __restrict__ partial, int nparts,
                                                             long long cols, float* __restrict__ out) {
    __shared__ float red[32][33];
    pdl_trigger();
    pdl_wait();
    const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
    const long long c = (long long)blockIdx.x * 32 + cx;
    float s = 0.f;
    if (c < cols) {
#pragma unroll 8
        for (int p = py; p < nparts; p += 32) s += __ldg(partial + (long long)p * cols + c);
    }
    red[py][cx] = s;
    __syncthreads();
    if (py == 0 && c < cols) {
        float t = 0.f;
#pragma unroll
        for (int y = 0; y < 32; ++y) t += red[y][cx];
        out[c] = t;
    }
}

}  // namespace

size_t stage_colsum_scratch_bytes(int64_t cols) {
    return (size_t)round_up(cols * 4 * 1024, 256);  // <= 1024 row chunks per column
}

int stage_operand(const View4& src, bool transpose, int prec, __nv_bfloat16* dst_hi,
                  __nv_bfloat16* dst_lo, int op, const float* aux, float beta, float* colsum,
                  float* colsum_scratch, cudaStream_t stream, Staged* out, const DropArgs* drop) {
    NNB_REQUIRE(!transpose, "stage_operand: transposing stage not implemented (use MN-major operand)");
    NNB_REQUIRE(!drop || (src.b0 * src.b1 == 1 && src.s_c == 1 && src.s_r == src.cols && src.cols % 4 == 0),
                "stage_operand: a dropout mask needs one contiguous [rows, cols] matrix with cols % 4 == 0");
    NNB_REQUIRE(src.ptr && dst_hi, "stage_operand: null pointer");
    NNB_REQUIRE(src.rows > 0 && src.cols > 0 && src.b0 > 0 && src.b1 > 0, "stage_operand: empty view");
    const bool x3 = prec == NNB_PREC_BF16X3;
    NNB_REQUIRE(!x3 || dst_lo, "stage_operand: BF16X3 needs a lo plane");
    NNB_REQUIRE(op == STAGE_COPY || aux, "stage_operand: transform needs aux");
    NNB_REQUIRE(!colsum || colsum_scratch, "stage_operand: colsum needs scratch");
    const int64_t batch = src.b0 * src.b1;
    const int64_t ld = staged_ld(src.cols);

    StageArgs a;
    a.v = src;
    a.aux = aux;
    a.beta = beta;
    a.hi = dst_hi;
    a.lo = dst_lo;
    a.ld = ld;
    a.plane_stride = src.rows * ld;
    a.partial = nullptr;
    if (drop) a.drop = *drop; else a.drop = DropArgs{};
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    a.vec_ok = src.s_c == 1 && (src.s_r % 4) == 0 && (src.s_b0 % 4) == 0 && (src.s_b1 % 4) == 0 &&
               al16(src.ptr) && (aux == nullptr || al16(aux));

    // with column sums: 512-thread blocks, so ~256 row chunks (= partial rows to re-read) still put
    // >= 128 K threads in flight; plain conversion keeps small blocks and up to 1024 chunks
    const int threads = colsum ? 512 : 128;
    int tx = 128;
    while (tx > 1 && (tx / 2) * 4 >= src.cols) tx >>= 1;  // smallest power of two covering the columns
    a.tx = tx;
    const int ty = threads / tx;
    const int64_t gx = ceil_div(src.cols, (int64_t)tx * 4);
    const int sms = num_sms();
    int64_t want_y = std::max<int64_t>(1, (int64_t)sms * 16 / std::max<int64_t>(1, gx * batch));
    want_y = std::min<int64_t>(want_y, 1024);
    // with column sums every row chunk adds a partial row that colsum_reduce_kernel must re-read:
    // ~256 chunks keep all SMs busy while the finishing pass stays a few microseconds
    if (colsum) want_y = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want_y, 256), 1024 / batch));
    NNB_REQUIRE(!colsum || batch <= 1024, "stage_operand: colsum with batch > 1024");
    int64_t rpb = std::max<int64_t>(ty, ceil_div(src.rows, want_y));
    rpb = round_up(rpb, ty);
    const int64_t gy = ceil_div(src.rows, rpb);
    NNB_REQUIRE(batch <= 65535 && gy <= 65535, "stage_operand: grid too large");
    a.rows_per_block = (int)rpb;
    if (colsum) a.partial = colsum_scratch;
    dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)batch);
    const dim3 block(threads);
#define NNB_STAGE_LAUNCH(X3_, OP_, DROP_) NNB_CUDA_OK(launch_pdl(stage_rows_kernel<X3_, OP_, DROP_>, grid, block, 0, stream, a))
    const bool swish = op == STAGE_SWISH_BWD;
    if (drop) {
        if (x3) { if (swish) NNB_STAGE_LAUNCH(true, STAGE_SWISH_BWD, true); else NNB_STAGE_LAUNCH(true, STAGE_COPY, true); }
        else    { if (swish) NNB_STAGE_LAUNCH(false, STAGE_SWISH_BWD, true); else NNB_STAGE_LAUNCH(false, STAGE_COPY, true); }
    } else {
        if (x3) { if (swish) NNB_STAGE_LAUNCH(true, STAGE_SWISH_BWD, false); else NNB_STAGE_LAUNCH(true, STAGE_COPY, false); }
        else    { if (swish) NNB_STAGE_LAUNCH(false, STAGE_SWISH_BWD, false); else NNB_STAGE_LAUNCH(false, STAGE_COPY, false); }
    }
#undef NNB_STAGE_LAUNCH
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    if (colsum) {
        const int nparts = (int)(gy * batch);
        NNB_CUDA_OK(launch_pdl(colsum_reduce_kernel, dim3((unsigned)ceil_div(src.cols, 32)), dim3(1024), 0, stream,
                               (const float*)colsum_scratch, nparts, (long long)src.cols, colsum));
        count_launch();
        NNB_CUDA_OK(cudaGetLastError());
    }
    if (out) {
        out->hi = dst_hi;
        out->lo = x3 ? dst_lo : nullptr;
        out->rows = src.rows;
        out->cols = src.cols;
        out->ld = ld;
        out->batch = batch;
        out->batch_stride = src.rows * ld;
    }
    return NNB_OK;
}

}  // namespace nnb
