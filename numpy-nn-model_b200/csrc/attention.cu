// Fused attention for short sequences (row N4 of SURVEY.md 8f): one kernel does
//     scores = q . k^T / scale;  scores = where(mask, fill, scores);  p = softmax(scores, -1);
//     attn = dropout(p);  out = attn . v
// (examples/gpt.ipynb cell 2 l.25-40 calls these as separate Tensor ops: two Tensor.matmul
// (neunet/autograd.py:192-230), a division, where, nn.Softmax (activations.py:437-459) and nn.Dropout
// (layers/dropout.py:17-46)), and one kernel does the whole backward chain. The Python side
// (neunet/autograd.py here) recognises the chain on deferred tensors and calls these instead.
//
// One CTA owns one (batch, head): Tq, Tk, D <= 64, so Q, K, V, the score tile and the gradients of one head
// live in shared memory and the O(T^2) intermediates never travel to HBM (only `attn`, which the API
// returns, is written). The per-head products (64 x 64 x 64) are warp-level TF32 tensor-core MMAs with fp32
// accumulation (see "tensor-core products" below: one product in bf16 mode, the hi/lo three-product split in
// bf16x3 mode); masking, softmax, dropout and every sum are fp32. The dropout mask is regenerated
// from its Philox ticket in backward (philox.cuh), indexed like the stand-alone kernel over the contiguous
// (B, H, Tq, Tk) `attn` tensor. Outputs are written in (B, T, H, D) memory order -- the layout the
// surrounding reshape/transpose views of the example expect, so no `.contiguous()` copy is ever made --
// and `out` can also be emitted as bf16 planes [B*Tq][H*D] for the following nn.Linear.
#include <cfloat>

#include "common.cuh"
#include "philox.cuh"

namespace nnb {
namespace {

constexpr int T = 64;    // tile edge: Tq, Tk, D <= T
constexpr int LD = 68;   // padded pitch (floats): 16-byte aligned rows, column walks spread over banks

struct AttnView {
    const float* p;
    long long s[4];  // element strides of the logical 4-D tensor
};

struct AttnMask {
    const void* p;     // null: no mask
    int kind;          // 1: float tensor, masked where value != 0; 2: int32 tensor == cmp; 3: float tensor == cmp
    float cmp;
    long long s[4];    // strides over (B, H, Tq, Tk); 0 on broadcast axes
    float fill;
};

struct AttnArgs {
    AttnView q, kt, v;     // q: (B,H,Tq,D); kt: (B,H,D,Tk); v: (B,H,Tk,D)
    AttnMask m;
    float scale;           // scores are DIVIDED by it, like the example
    float inv_scale;       // 1 / scale (the tensor-core kernels multiply)
    int drop;              // 0: no dropout
    DropArgs d;
    int H, Tq, Tk, D;
};

__device__ __forceinline__ bool masked(const AttnMask& m, long long off) {
    if (m.kind == 2) return (float)static_cast<const int*>(m.p)[off] == m.cmp;
    const float x = static_cast<const float*>(m.p)[off];
    return m.kind == 1 ? x != 0.f : x == m.cmp;
}

// One 64 x 64 (zero padded) tile of a strided fp32 matrix, held in registers between the global loads and the shared
// stores: every thread issues its 4 x 16-byte loads back to back (and the caller issues ALL its tiles before the first
// store), so a CTA pays the HBM/L2 latency once instead of once per element -- with one CTA of 8 warps per SM a
// load -> store -> load chain was the whole kernel time (ncu, round 2: 247 us per backward launch before this).
// `inner` is the axis that is contiguous in global memory (stride 1): 16 threads x float4 cover it, rows = outer axis.
struct TileRegs {
    float4 v[4];
};

__device__ __forceinline__ bool tile_vec_ok(const float* p, long long sr, long long sc, int R, int C) {
    const bool col_fast = sc == 1;
    const long long so = col_fast ? sr : sc;
    const int inner = col_fast ? C : R;
    return (sc == 1 || sr == 1) && (so % 4) == 0 && (inner % 4) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0;
}

// logical matrix (rows R, cols C), strides (sr, sc); requires tile_vec_ok
__device__ __forceinline__ void tile_load(TileRegs& t, const float* p, long long sr, long long sc, int R, int C, int tid) {
    const bool col_fast = sc == 1;                 // inner axis = cols
    const long long so = col_fast ? sr : sc;       // stride of the outer axis
    const int n_in = col_fast ? C : R, n_out = col_fast ? R : C;
    const int i4 = (tid & 15) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int o = (tid >> 4) + 16 * j;
        t.v[j] = (o < n_out && i4 < n_in) ? __ldg(reinterpret_cast<const float4*>(p + o * so + i4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// store the registers as dst[r][c] (transposed = false) or dst[c][r] (transposed = true) of the LOGICAL matrix
__device__ __forceinline__ void tile_store(const TileRegs& t, float* dst, long long sc, bool transposed, int tid, int ld = LD) {
    const bool col_fast = sc == 1;
    const int i4 = (tid & 15) * 4;
    // memory-order (outer o, inner i): logical (r, c) = col_fast ? (o, i) : (i, o); wanted smem order rows = transposed ? c : r
    const bool inner_is_smem_col = (col_fast != transposed);  // the inner axis runs along smem columns -> float4 store
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int o = (tid >> 4) + 16 * j;
        if (inner_is_smem_col) {
            *reinterpret_cast<float4*>(dst + o * ld + i4) = t.v[j];
        } else {
            dst[(i4 + 0) * ld + o] = t.v[j].x;
            dst[(i4 + 1) * ld + o] = t.v[j].y;
            dst[(i4 + 2) * ld + o] = t.v[j].z;
            dst[(i4 + 3) * ld + o] = t.v[j].w;
        }
    }
}

// generic fallback (any strides / alignment): element-wise, zero padded
__device__ __forceinline__ void load_tile(float* dst, const float* p, long long sr, long long sc, int R, int C,
                                          bool transposed, int tid, int nthreads, int ld = LD) {
    const bool col_fast = sc == 1 || sr != 1;
    for (int e = tid; e < T * T; e += nthreads) {
        const int r = col_fast ? e / T : e % T, c = col_fast ? e % T : e / T;
        const float x = (r < R && c < C) ? p[r * sr + c * sc] : 0.f;
        dst[transposed ? c * ld + r : r * ld + c] = x;
    }
}

// Reductions over the 16 lanes that own one row. The two half-warps of a warp own DIFFERENT rows and may diverge
// (ragged Tq: one row live, the other padding), so each names only its own 16 lanes in the shuffle mask.
__device__ __forceinline__ unsigned half_mask() { return (threadIdx.x & 16) ? 0xffff0000u : 0x0000ffffu; }
__device__ __forceinline__ float half_warp_sum(float v) {
    const unsigned m = half_mask();
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o);
    return v;
}
__device__ __forceinline__ float half_warp_max(float v) {
    const unsigned m = half_mask();
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(m, v, o));
    return v;
}

// Row phase shared by forward and backward: S (scaled, masked scores, [q][key]) -> P = softmax rows, in place.
// A half-warp owns a row: lane l holds keys 4l..4l+3 (Tk % 4 == 0). Returns this lane's four probabilities.
__device__ __forceinline__ void softmax_row(float* Srow, int Tk, int l, float (&p)[4]) {
    const int k0 = 4 * l;
    float4 s = *reinterpret_cast<const float4*>(Srow + k0);
    float sv[4] = {s.x, s.y, s.z, s.w};
    float mx = -FLT_MAX;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (k0 + j < Tk) mx = fmaxf(mx, sv[j]);
    mx = half_warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        p[j] = (k0 + j < Tk) ? __expf(sv[j] - mx) : 0.f;  // argument <= 0: ex2.approx, ~2 ulp
        sum += p[j];
    }
    sum = half_warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] *= inv;
}

// ------------------------------------------------------------------------------------------ tensor-core products
// Every 64 x 64 x 64 product runs on the tensor cores: warp-level mma.sync m16n8k8 (TF32 operands, fp32 accumulate)
// straight from the fp32 shared-memory tiles. tcgen05 is the wrong tool at this size -- its M = 128 tile is half padding
// for one head and its operands must be staged in the UMMA shared-memory layout by TMA, i.e. from bf16 planes no
// producer has -- while the warp-level path takes fragments from the tiles as they lie. The first version of these
// kernels did the products as fp32 FMAs on the CUDA cores and was issue-bound (ncu, profiles/r2_attention_ncu.md: 46 % issue
// slots busy, 21.6 M warp instructions per forward launch, 92 us per backward launch); the MMA form halves the
// instruction stream (10.8 M) and runs the backward in 36 us. Precision follows the library's two modes: NNB_PREC_BF16 -> one TF32
// product (10-bit mantissa, tighter than the bf16 GEMMs around it), NNB_PREC_BF16X3 -> the same hi/lo split as the
// GEMMs, x = hi + lo with hi = tf32(x), lo = tf32(x - hi), three products (lo*hi + hi*lo + hi*hi, ~1e-6 relative: fp32-grade).
//
// Tiles are kept in their NATURAL memory order ([q][d], [key][d]) so no transposed copy is made; a fragment read is
// one LDS.32 per element. With a pitch of 68 floats the [row][k] reads of a warp (8 rows x 4 k) fall in 32 distinct
// banks; the [k][col] reads (4 k x 8 cols) are conflict-free at a pitch of 72 and 2-way conflicted at 68.
constexpr int LP = 72;  // pitch of the tiles that are only ever walked k-major

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// acc (16 x 32 block of C at rows m0.., cols n0..; acc[nt] = the m16n8 fragment of columns n0 + 8 nt..) += A . B over
// k in [0, K), K a multiple of 8 (tiles are zero padded). A(m, k) is A[m * lda + k] (A_MK) or A[k * lda + m];
// B(k, n) is B[n * ldb + k] (B_NK) or B[k * ldb + n].
// Fragment ownership (PTX ISA, m16n8k8 .tf32): g = lane / 4, t = lane % 4;
//   a0 (g, t)  a1 (g + 8, t)  a2 (g, t + 4)  a3 (g + 8, t + 4);  b0 (k = t, n = g)  b1 (k = t + 4, n = g);
//   c0 (g, 2t)  c1 (g, 2t + 1)  c2 (g + 8, 2t)  c3 (g + 8, 2t + 1).
template <bool A_MK, bool B_NK, bool X3>
__device__ __forceinline__ void warp_mma(float (&acc)[4][4], const float* __restrict__ A, int lda,
                                         const float* __restrict__ B, int ldb, int K, int m0, int n0, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const float* a_lo = A_MK ? A + (m0 + g) * lda + t : A + t * lda + m0 + g;   // (m0 + g, k = t)
    const int a_dm = A_MK ? 8 * lda : 8, a_dk = A_MK ? 4 : 4 * lda, a_step = A_MK ? 8 : 8 * lda;
    const float* b_lo = B_NK ? B + (n0 + g) * ldb + t : B + t * ldb + n0 + g;   // (k = t, n0 + g)
    const int b_dn = B_NK ? 8 * ldb : 8, b_dk = B_NK ? 4 : 4 * ldb, b_step = B_NK ? 8 : 8 * ldb;
#pragma unroll 2
    for (int k0 = 0; k0 < K; k0 += 8) {
        const float af[4] = {a_lo[0], a_lo[a_dm], a_lo[a_dk], a_lo[a_dm + a_dk]};
        uint32_t ah[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ah[i] = to_tf32(af[i]);
            al[i] = X3 ? to_tf32(af[i] - __uint_as_float(ah[i])) : 0u;
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const float bf[2] = {b_lo[nt * b_dn], b_lo[nt * b_dn + b_dk]};
            const uint32_t bh[2] = {to_tf32(bf[0]), to_tf32(bf[1])};
            if (X3) {
                const uint32_t bl[2] = {to_tf32(bf[0] - __uint_as_float(bh[0])), to_tf32(bf[1] - __uint_as_float(bh[1]))};
                mma_tf32(acc[nt], al, bh);
                mma_tf32(acc[nt], ah, bl);
            }
            mma_tf32(acc[nt], ah, bh);
        }
        a_lo += a_step;
        b_lo += b_step;
    }
}

__device__ __forceinline__ void zero_frag(float (&acc)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}

// The mask tile as bits in shared memory: thread tid owns row q = tid / 4, keys 16 (tid % 4) .. + 15 and returns them as one
// 16-bit word (bit j = key 16 (tid % 4) + j is masked). Called between the tile loads and the tile stores, so the mask
// requests travel with the q / k / v requests instead of stalling the score epilogue (ncu: 21 % of all stall samples of
// the first tensor-core version sat on the per-element mask load). All three mask kinds are 4-byte elements: contiguous,
// aligned rows go as 16-byte loads.
__device__ __forceinline__ uint32_t mask_word_load(const AttnArgs& a, int b, int h, int tid) {
    if (a.m.p == nullptr) return 0u;
    const int q = tid >> 2, k0 = (tid & 3) * 16;
    if (q >= a.Tq || k0 >= a.Tk) return 0u;
    const long long base = b * a.m.s[0] + h * a.m.s[1] + q * a.m.s[2];
    uint32_t bits = 0;
    const bool vec = a.m.s[3] == 1 && ((a.m.s[0] | a.m.s[1] | a.m.s[2]) & 3) == 0 && (reinterpret_cast<uintptr_t>(a.m.p) & 15) == 0 &&
                     (a.Tk & 3) == 0;
    if (vec) {
        const int4* src = reinterpret_cast<const int4*>(static_cast<const int*>(a.m.p) + base + k0);
        int4 w[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) w[c] = (k0 + 4 * c < a.Tk) ? __ldg(src + c) : make_int4(0, 0, 0, 0);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (k0 + 4 * c >= a.Tk) continue;
            const int e[4] = {w[c].x, w[c].y, w[c].z, w[c].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool m = a.m.kind == 2 ? (float)e[j] == a.m.cmp
                                             : (a.m.kind == 1 ? __int_as_float(e[j]) != 0.f : __int_as_float(e[j]) == a.m.cmp);
                if (m) bits |= 1u << (4 * c + j);
            }
        }
    } else {
#pragma unroll 4
        for (int j = 0; j < 16; ++j)
            if (k0 + j < a.Tk && masked(a.m, base + (k0 + j) * a.m.s[3])) bits |= 1u << j;
    }
    return bits;
}

// scores fragment -> S[q][key] (pitch ld): * 1/scale, mask
__device__ __forceinline__ void scores_frag_to_smem(const AttnArgs& a, const float (&acc)[4][4], const uint16_t* mask_sm, float* S,
                                                    int ld, int m0, int n0, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        const int q = m0 + g + 8 * hf;
        const uint32_t mrow = reinterpret_cast<const uint32_t*>(mask_sm)[q * 2 + (n0 >> 5)];  // keys n0 .. n0 + 31 of row q
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int key = n0 + nt * 8 + 2 * t;
            float o[2];
#pragma unroll
            for (int j = 0; j < 2; ++j)
                o[j] = (mrow >> (nt * 8 + 2 * t + j)) & 1u ? a.m.fill : acc[nt][2 * hf + j] * a.inv_scale;
            *reinterpret_cast<float2*>(S + q * ld + key) = make_float2(o[0], o[1]);
        }
    }
}

// fragment -> global matrix with rows `rows` x cols `cols` (row pitch in floats), 8-byte stores
__device__ __forceinline__ void frag_to_global(const float (&acc)[4][4], float* dst, long long pitch, int rows, int cols, int m0,
                                               int n0, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const int c = n0 + nt * 8 + 2 * t;
        if (c >= cols) continue;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int r = m0 + g + 8 * hf;
            if (r < rows) *reinterpret_cast<float2*>(dst + r * pitch + c) = make_float2(acc[nt][2 * hf], acc[nt][2 * hf + 1]);
        }
    }
}

template <bool X3>
__global__ void __launch_bounds__(256, 4) attn_fwd_mma_kernel(const AttnArgs a, float* __restrict__ attn, float* __restrict__ out,
                                                              __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
    extern __shared__ __align__(16) float sm[];
    float* Qn = sm;                 // [q][d]    pitch LD
    float* Kn = Qn + T * LD;        // [key][d]  pitch LD
    float* Vn = Kn + T * LD;        // [key][d]  pitch LP (read k-major by P . V)
    float* S = Qn;                  // [q][key]  scores -> probabilities, re-uses Q once the scores are in registers
    uint16_t* mask_sm = reinterpret_cast<uint16_t*>(Vn + T * LP);  // [q][4]: the mask tile as bits
    pdl_trigger();
    pdl_wait();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = (warp & 3) * 16, n0 = (warp >> 2) * 32;
    const int bh = blockIdx.x, b = bh / a.H, h = bh - b * a.H;
    {
        const float* qp = a.q.p + b * a.q.s[0] + h * a.q.s[1];
        const float* kp = a.kt.p + b * a.kt.s[0] + h * a.kt.s[1];
        const float* vp = a.v.p + b * a.v.s[0] + h * a.v.s[1];
        if (tile_vec_ok(qp, a.q.s[2], a.q.s[3], a.Tq, a.D) && tile_vec_ok(kp, a.kt.s[2], a.kt.s[3], a.D, a.Tk) &&
            tile_vec_ok(vp, a.v.s[2], a.v.s[3], a.Tk, a.D)) {
            TileRegs rq, rk, rv;
            tile_load(rq, qp, a.q.s[2], a.q.s[3], a.Tq, a.D, tid);
            tile_load(rk, kp, a.kt.s[2], a.kt.s[3], a.D, a.Tk, tid);
            tile_load(rv, vp, a.v.s[2], a.v.s[3], a.Tk, a.D, tid);
            const uint32_t mw = mask_word_load(a, b, h, tid);
            tile_store(rq, Qn, a.q.s[3], false, tid, LD);
            tile_store(rk, Kn, a.kt.s[3], true, tid, LD);   // kT is (d, key): stored transposed = [key][d]
            tile_store(rv, Vn, a.v.s[3], false, tid, LP);
            mask_sm[tid] = (uint16_t)mw;
        } else {
            mask_sm[tid] = (uint16_t)mask_word_load(a, b, h, tid);
            load_tile(Qn, qp, a.q.s[2], a.q.s[3], a.Tq, a.D, false, tid, 256, LD);
            load_tile(Kn, kp, a.kt.s[2], a.kt.s[3], a.D, a.Tk, true, tid, 256, LD);
            load_tile(Vn, vp, a.v.s[2], a.v.s[3], a.Tk, a.D, false, tid, 256, LP);
        }
    }
    __syncthreads();
    {
        float acc[4][4];
        zero_frag(acc);
        warp_mma<true, true, X3>(acc, Qn, LD, Kn, LD, (a.D + 7) & ~7, m0, n0, lane);
        __syncthreads();  // S re-uses Q's storage
        scores_frag_to_smem(a, acc, mask_sm, S, LD, m0, n0, lane);
    }
    __syncthreads();
    {
        const uint64_t epoch = a.drop ? drop_epoch(a.d) : 0;
        const int l = tid & 15;
        for (int q = tid >> 4; q < T; q += 16) {
            float p[4] = {0.f, 0.f, 0.f, 0.f};
            if (q < a.Tq) {
                softmax_row(S + q * LD, a.Tk, l, p);
                if (4 * l < a.Tk) {
                    const long long e = ((long long)bh * a.Tq + q) * a.Tk + 4 * l;
                    if (a.drop) {
                        uint32_t r[4];
                        drop_words(a.d, epoch, e >> 2, r);
#pragma unroll
                        for (int j = 0; j < 4; ++j) p[j] = r[j] >= a.d.thresh ? p[j] * a.d.scale : 0.f;
                    }
                    *reinterpret_cast<float4*>(attn + e) = make_float4(p[0], p[1], p[2], p[3]);
                }
            }
            *reinterpret_cast<float4*>(S + q * LD + 4 * l) = make_float4(p[0], p[1], p[2], p[3]);  // A of P . V, in place
        }
    }
    __syncthreads();
    if (n0 < a.D) {
        float acc[4][4];
        zero_frag(acc);
        warp_mma<true, false, X3>(acc, S, LD, Vn, LP, (a.Tk + 7) & ~7, m0, n0, lane);
        const long long hd = (long long)a.H * a.D;
        const long long base = (long long)b * a.Tq * hd + (long long)h * a.D;  // (B, Tq, H, D)
        frag_to_global(acc, out + base, hd, a.Tq, a.D, m0, n0, lane);
        if (out_hi != nullptr) {
            const int g = lane >> 2, t = lane & 3;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int c = n0 + nt * 8 + 2 * t;
                if (c >= a.D) continue;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int r = m0 + g + 8 * hf;
                    if (r >= a.Tq) continue;
                    const float x0 = acc[nt][2 * hf], x1 = acc[nt][2 * hf + 1];
                    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                    const long long off = base + r * hd + c;
                    *reinterpret_cast<__nv_bfloat162*>(out_hi + off) = __nv_bfloat162(h0, h1);
                    if (out_lo != nullptr)
                        *reinterpret_cast<__nv_bfloat162*>(out_lo + off) =
                            __nv_bfloat162(__float2bfloat16_rn(x0 - __bfloat162float(h0)), __float2bfloat16_rn(x1 - __bfloat162float(h1)));
                }
            }
        }
    }
}

template <bool X3>
__global__ void __launch_bounds__(256, 2) attn_bwd_mma_kernel(const AttnArgs a, const AttnView dO, float* __restrict__ dQ,
                                                              float* __restrict__ dK, float* __restrict__ dV, long long pitch) {
    extern __shared__ __align__(16) float sm[];
    float* Qn = sm;                  // [q][d]     A of S, B of dK
    float* Kn = Qn + T * LD;         // [key][d]   B of S, B of dQ
    float* Vn = Kn + T * LD;         // [key][d]   B of dPd
    float* dOn = Vn + T * LD;        // [q][d]     A of dPd, B of dV
    float* P = dOn + T * LD;         // [q][key]   pitch LP: S -> P -> P*mask/(1-p); A of dV (k-major)
    float* dS = Vn;                  // [q][key]   dPd -> dS; re-uses V after phase 1 (5 tiles = 88 KB: 2 CTAs per SM)
    uint16_t* mask_sm = reinterpret_cast<uint16_t*>(P + T * LP);  // [q][4]: the mask tile as bits
    pdl_trigger();
    pdl_wait();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = (warp & 3) * 16, n0 = (warp >> 2) * 32;
    const int bh = blockIdx.x, b = bh / a.H, h = bh - b * a.H;
    {
        const float* qp = a.q.p + b * a.q.s[0] + h * a.q.s[1];
        const float* kp = a.kt.p + b * a.kt.s[0] + h * a.kt.s[1];
        const float* vp = a.v.p + b * a.v.s[0] + h * a.v.s[1];
        const float* gp = dO.p + b * dO.s[0] + h * dO.s[1];
        if (tile_vec_ok(qp, a.q.s[2], a.q.s[3], a.Tq, a.D) && tile_vec_ok(kp, a.kt.s[2], a.kt.s[3], a.D, a.Tk) &&
            tile_vec_ok(vp, a.v.s[2], a.v.s[3], a.Tk, a.D) && tile_vec_ok(gp, dO.s[2], dO.s[3], a.Tq, a.D)) {
            TileRegs rq, rk, rv, rg;
            tile_load(rq, qp, a.q.s[2], a.q.s[3], a.Tq, a.D, tid);
            tile_load(rk, kp, a.kt.s[2], a.kt.s[3], a.D, a.Tk, tid);
            tile_load(rv, vp, a.v.s[2], a.v.s[3], a.Tk, a.D, tid);
            tile_load(rg, gp, dO.s[2], dO.s[3], a.Tq, a.D, tid);
            const uint32_t mw = mask_word_load(a, b, h, tid);
            tile_store(rq, Qn, a.q.s[3], false, tid, LD);
            tile_store(rk, Kn, a.kt.s[3], true, tid, LD);
            tile_store(rv, Vn, a.v.s[3], false, tid, LD);
            tile_store(rg, dOn, dO.s[3], false, tid, LD);
            mask_sm[tid] = (uint16_t)mw;
        } else {
            mask_sm[tid] = (uint16_t)mask_word_load(a, b, h, tid);
            load_tile(Qn, qp, a.q.s[2], a.q.s[3], a.Tq, a.D, false, tid, 256, LD);
            load_tile(Kn, kp, a.kt.s[2], a.kt.s[3], a.D, a.Tk, true, tid, 256, LD);
            load_tile(Vn, vp, a.v.s[2], a.v.s[3], a.Tk, a.D, false, tid, 256, LD);
            load_tile(dOn, gp, dO.s[2], dO.s[3], a.Tq, a.D, false, tid, 256, LD);
        }
    }
    __syncthreads();
    const int Dp = (a.D + 7) & ~7, Tqp = (a.Tq + 7) & ~7, Tkp = (a.Tk + 7) & ~7;
    // phase 1: S = Q . K^T and dPd = dO . V^T, every warp its 16 x 32 block of both
    {
        float accS[4][4], accG[4][4];
        zero_frag(accS);
        zero_frag(accG);
        warp_mma<true, true, X3>(accS, Qn, LD, Kn, LD, Dp, m0, n0, lane);
        warp_mma<true, true, X3>(accG, dOn, LD, Vn, LD, Dp, m0, n0, lane);
        __syncthreads();  // dS re-uses V's tile: all reads of phase 1 are done
        scores_frag_to_smem(a, accS, mask_sm, P, LP, m0, n0, lane);
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf)
                *reinterpret_cast<float2*>(dS + (m0 + g + 8 * hf) * LD + n0 + nt * 8 + 2 * t) =
                    make_float2(accG[nt][2 * hf], accG[nt][2 * hf + 1]);
    }
    __syncthreads();
    // phase 2: rows (a half-warp per row)
    {
        const uint64_t epoch = a.drop ? drop_epoch(a.d) : 0;
        const int l = tid & 15;
        for (int q = tid >> 4; q < T; q += 16) {
            float p[4] = {0.f, 0.f, 0.f, 0.f}, ds[4] = {0.f, 0.f, 0.f, 0.f}, pd[4] = {0.f, 0.f, 0.f, 0.f};
            if (q < a.Tq) {
                const unsigned mrow = (unsigned)mask_sm[q * 4 + (l >> 2)] >> ((l & 3) * 4);  // this lane's four mask bits
                softmax_row(P + q * LP, a.Tk, l, p);
                const float4 g4 = *reinterpret_cast<const float4*>(dS + q * LD + 4 * l);
                float dp[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) pd[j] = p[j];
                if (a.drop && 4 * l < a.Tk) {
                    uint32_t r[4];
                    drop_words(a.d, epoch, (((long long)bh * a.Tq + q) * a.Tk + 4 * l) >> 2, r);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool keep = r[j] >= a.d.thresh;
                        dp[j] = keep ? dp[j] * a.d.scale : 0.f;
                        pd[j] = keep ? p[j] * a.d.scale : 0.f;
                    }
                }
                float tsum = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) tsum += (4 * l + j < a.Tk) ? dp[j] * p[j] : 0.f;
                tsum = half_warp_sum(tsum);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int key = 4 * l + j;
                    const float v = (key < a.Tk) ? (dp[j] - tsum) * p[j] : 0.f;
                    ds[j] = (mrow >> j) & 1u ? 0.f : v * a.inv_scale;  // where(): masked scores receive no gradient
                }
            }
            *reinterpret_cast<float4*>(P + q * LP + 4 * l) = make_float4(pd[0], pd[1], pd[2], pd[3]);
            *reinterpret_cast<float4*>(dS + q * LD + 4 * l) = make_float4(ds[0], ds[1], ds[2], ds[3]);
        }
    }
    __syncthreads();
    // phase 3: dV = Pd^T . dO, dQ = dS . K, dK = dS^T . Q; every warp its 16 x 32 block of each
    if (n0 < a.D) {
        const long long base = (long long)h * a.D;  // + (b * T_out + r) * pitch: (B, T_out, H, D) memory order
        float acc[4][4];
        if (dV != nullptr && m0 < a.Tk) {
            zero_frag(acc);
            warp_mma<false, false, X3>(acc, P, LP, dOn, LD, Tqp, m0, n0, lane);
            frag_to_global(acc, dV + (long long)b * a.Tk * pitch + base, pitch, a.Tk, a.D, m0, n0, lane);
        }
        if (dQ != nullptr && m0 < a.Tq) {
            zero_frag(acc);
            warp_mma<true, false, X3>(acc, dS, LD, Kn, LD, Tkp, m0, n0, lane);
            frag_to_global(acc, dQ + (long long)b * a.Tq * pitch + base, pitch, a.Tq, a.D, m0, n0, lane);
        }
        if (dK != nullptr && m0 < a.Tk) {
            zero_frag(acc);
            warp_mma<false, false, X3>(acc, dS, LD, Qn, LD, Tqp, m0, n0, lane);
            frag_to_global(acc, dK + (long long)b * a.Tk * pitch + base, pitch, a.Tk, a.D, m0, n0, lane);
        }
    }
}

int fill_args(AttnArgs& a, const float* Q, const int64_t* qs, const float* KT, const int64_t* ks, const float* V,
              const int64_t* vs, const void* mask, int mask_kind, float mask_cmp, const int64_t* ms, float fill,
              float scale, float p, uint64_t seed, uint32_t call_id, uint64_t epoch, const uint64_t* epoch_dev,
              int64_t B, int64_t H, int64_t Tq, int64_t Tk, int64_t D) {
    NNB_REQUIRE(Q && KT && V && qs && ks && vs, "nnb_attention: null pointer");
    NNB_REQUIRE(B > 0 && H > 0 && Tq > 0 && Tk > 0 && D > 0, "nnb_attention: non-positive dimension");
    if (!(Tq <= T && Tk <= T && D <= T && Tk % 4 == 0 && D % 4 == 0))
        return fail(NNB_ERR_UNSUPPORTED, "nnb_attention: the fused kernel needs Tq, Tk, D <= 64 with Tk %% 4 == D %% 4 == 0");
    NNB_REQUIRE(B * H < (1ll << 31), "nnb_attention: too many heads");
    NNB_REQUIRE(mask == nullptr || (ms != nullptr && mask_kind >= 1 && mask_kind <= 3), "nnb_attention: bad mask description");
    NNB_REQUIRE(scale != 0.f, "nnb_attention: scale must be non-zero");
    NNB_REQUIRE(p >= 0.f && p < 1.f, "nnb_attention: p must be in [0, 1)");
    a.q.p = Q; a.kt.p = KT; a.v.p = V;
    for (int i = 0; i < 4; ++i) { a.q.s[i] = qs[i]; a.kt.s[i] = ks[i]; a.v.s[i] = vs[i]; a.m.s[i] = mask ? ms[i] : 0; }
    a.m.p = mask; a.m.kind = mask_kind; a.m.cmp = mask_cmp; a.m.fill = fill;
    a.scale = scale;
    a.inv_scale = (float)(1.0 / (double)scale);
    a.drop = p > 0.f ? 1 : 0;
    a.d = make_drop_args(p, seed, call_id, epoch, epoch_dev);
    a.H = (int)H; a.Tq = (int)Tq; a.Tk = (int)Tk; a.D = (int)D;
    return NNB_OK;
}

}  // namespace
}  // namespace nnb

using namespace nnb;

extern "C" {

int nnb_attention_supported(int64_t Tq, int64_t Tk, int64_t D) {
    return (Tq > 0 && Tk > 0 && D > 0 && Tq <= T && Tk <= T && D <= T && Tk % 4 == 0 && D % 4 == 0) ? 1 : 0;
}

int nnb_attention_forward(const float* Q, const int64_t q_strides[4], const float* KT, const int64_t kt_strides[4],
                          const float* V, const int64_t v_strides[4], const void* mask, int mask_kind, float mask_cmp,
                          const int64_t mask_strides[4], float fill, float scale, float p, uint64_t seed,
                          uint32_t call_id, uint64_t epoch, const uint64_t* epoch_dev, float* attn, float* out,
                          void* out_staged, int prec, int64_t B, int64_t H, int64_t Tq, int64_t Tk, int64_t D,
                          cudaStream_t stream) {
    NNB_RANGE("nnb_attention_forward");
    AttnArgs a;
    int rc = fill_args(a, Q, q_strides, KT, kt_strides, V, v_strides, mask, mask_kind, mask_cmp, mask_strides, fill, scale, p,
                       seed, call_id, epoch, epoch_dev, B, H, Tq, Tk, D);
    if (rc) return rc;
    NNB_REQUIRE(attn && out, "nnb_attention_forward: null output");
    NNB_REQUIRE(((reinterpret_cast<uintptr_t>(attn) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "nnb_attention_forward: outputs must be 16-byte aligned");
    __nv_bfloat16 *hi = nullptr, *lo = nullptr;
    if (out_staged != nullptr) {
        NNB_REQUIRE((H * D) % 8 == 0, "nnb_attention_forward: staged output needs H*D %% 8 == 0");
        NNB_REQUIRE((reinterpret_cast<uintptr_t>(out_staged) & 255) == 0, "nnb_attention_forward: out_staged must be 256-byte aligned");
        hi = static_cast<__nv_bfloat16*>(out_staged);
        if (prec == NNB_PREC_BF16X3)
            lo = reinterpret_cast<__nv_bfloat16*>(static_cast<uint8_t*>(out_staged) + staged_plane_bytes(1, B * Tq, H * D));
    }
    NNB_REQUIRE(prec == NNB_PREC_BF16 || prec == NNB_PREC_BF16X3, "nnb_attention_forward: bad prec");
    static bool configured = false;
    const size_t smem_mma = (size_t)(2 * T * LD + T * LP) * sizeof(float) + T * 4 * sizeof(uint16_t);
    if (!configured) {
        NNB_CUDA_OK(cudaFuncSetAttribute(attn_fwd_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma));
        NNB_CUDA_OK(cudaFuncSetAttribute(attn_fwd_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma));
        configured = true;
    }
    const dim3 grid((unsigned)(B * H));
    if (prec == NNB_PREC_BF16X3) NNB_CUDA_OK(launch_pdl(attn_fwd_mma_kernel<true>, grid, dim3(256), smem_mma, stream, a, attn, out, hi, lo));
    else NNB_CUDA_OK(launch_pdl(attn_fwd_mma_kernel<false>, grid, dim3(256), smem_mma, stream, a, attn, out, hi, lo));
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

int nnb_attention_backward(const float* Q, const int64_t q_strides[4], const float* KT, const int64_t kt_strides[4],
                           const float* V, const int64_t v_strides[4], const void* mask, int mask_kind, float mask_cmp,
                           const int64_t mask_strides[4], float fill, float scale, float p, uint64_t seed,
                           uint32_t call_id, uint64_t epoch, const uint64_t* epoch_dev, const float* dO,
                           const int64_t do_strides[4], float* dQ, float* dK, float* dV, int64_t out_row_pitch,
                           int prec, int64_t B, int64_t H, int64_t Tq, int64_t Tk, int64_t D, cudaStream_t stream) {
    NNB_RANGE("nnb_attention_backward");
    AttnArgs a;
    int rc = fill_args(a, Q, q_strides, KT, kt_strides, V, v_strides, mask, mask_kind, mask_cmp, mask_strides, fill, scale, p,
                       seed, call_id, epoch, epoch_dev, B, H, Tq, Tk, D);
    if (rc) return rc;
    NNB_REQUIRE(dO && do_strides, "nnb_attention_backward: null dO");
    NNB_REQUIRE(((reinterpret_cast<uintptr_t>(dQ) | reinterpret_cast<uintptr_t>(dK) | reinterpret_cast<uintptr_t>(dV)) & 15) == 0,
                "nnb_attention_backward: outputs must be 16-byte aligned");
    AttnView g;
    g.p = dO;
    for (int i = 0; i < 4; ++i) g.s[i] = do_strides[i];
    NNB_REQUIRE(prec == NNB_PREC_BF16 || prec == NNB_PREC_BF16X3, "nnb_attention_backward: bad prec");
    NNB_REQUIRE(out_row_pitch == 0 || (out_row_pitch >= H * D && out_row_pitch % 4 == 0), "nnb_attention_backward: bad out_row_pitch");
    static bool configured = false;
    const size_t smem_mma = (size_t)(4 * T * LD + T * LP) * sizeof(float) + T * 4 * sizeof(uint16_t);
    if (!configured) {
        NNB_CUDA_OK(cudaFuncSetAttribute(attn_bwd_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma));
        NNB_CUDA_OK(cudaFuncSetAttribute(attn_bwd_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma));
        configured = true;
    }
    const dim3 grid((unsigned)(B * H));
    const long long pitch = (long long)(out_row_pitch ? out_row_pitch : H * D);
    if (prec == NNB_PREC_BF16X3) NNB_CUDA_OK(launch_pdl(attn_bwd_mma_kernel<true>, grid, dim3(256), smem_mma, stream, a, g, dQ, dK, dV, pitch));
    else NNB_CUDA_OK(launch_pdl(attn_bwd_mma_kernel<false>, grid, dim3(256), smem_mma, stream, a, g, dQ, dK, dV, pitch));
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

}  // extern "C"
