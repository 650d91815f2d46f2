// nn.Conv2d forward / backward (NCHW), semantics of neunet/nn/layers/conv2d.py:16-117, 297-355.
//
// Large-channel layers run as GEMMs on the tcgen05 kernel:
//   fwd  : O[o, (b,ho,wo)]  = W[o, (i,k,l)] . col(X)[(b,ho,wo), (i,k,l)]^T      + bias[o]
//   wgrad: dW[o, (i,k,l)]   = g'[o, (b,ho,wo)] . col(X)[(b,ho,wo), (i,k,l)]     (reduction over b,ho,wo)
//   dgrad: dX[i, (b,y,x)]   = Wr[i, (o,k,l)] . colT(g)[(b,y,x), (o,k,l)]^T      (Wr = rot180, in/out swapped;
//          colT gathers g through the stride -- the reference's zero-stuffed, (k-1)-padded grad, conv2d.py:35-96)
// with one generic gather kernel producing the bf16 `col` planes and the GEMM epilogue writing NCHW
// directly through its column-group index map. Tiny-channel layers (the MNIST classifier's 1->8 and
// 8->16, the DDPM 128->3 output) use direct fp32 CUDA-core kernels: a 128-wide tensor tile would
// be >90 % padding there.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "workspace.cuh"

namespace nnb {
namespace {

int planes(int prec) { return prec == NNB_PREC_BF16X3 ? 2 : 1; }
int g_implicit = [] { const char* e = getenv("NNB_CONV_IMPLICIT"); return e ? atoi(e) : 1; }();  // 0: always materialise `col`

struct Geo {
    int B, Cin, H, W, Cout, kh, kw, Ho, Wo;
    int s0, s1, pt, pl, d0, d1;
};

// Generic gather ("im2col"):
//   col[(b, p, q)][(c, k, l)] = src[b, c, yy, xx]  with  ny = p*sp0 - off0 + k*dk0,  yy = ny / up0
//   (0 unless ny % up0 == 0 and 0 <= yy < Hs; same along x).
// forward / wgrad: src = X,  (p,q) = (ho,wo), sp = stride, off = pad, dk = dilation, up = 1
// dgrad          : src = dO, (p,q) = (y,x),   sp = 1, off = dil*(k-1) - pad, dk = dilation, up = stride
//                  (taps are then the rot180-flipped ones, matched by the flipped weight staging)
struct GatherArgs {
    const float* src;
    int C, Hs, Ws;          // source channels / spatial size
    int P, Q;               // output positions per image
    int kh, kw;
    int sp0, sp1, off0, off1, dk0, dk1, up0, up1;
    long long M;            // B * P * Q rows
    int Kc;                 // C * kh * kw cols
    long long ld;
    __nv_bfloat16* hi;
    __nv_bfloat16* lo;
};

template <bool X3>
__global__ void __launch_bounds__(256) gather_cols_kernel(const GatherArgs a) {
    // thread -> one (row m, 4 consecutive kidx); kidx fastest so stores coalesce
    const int kq = (a.Kc + 3) >> 2;
    const long long total = a.M * kq;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long m = idx / kq;
        const int k4 = (int)(idx - m * kq) * 4;
        const int q = (int)(m % a.Q);
        const long long t = m / a.Q;
        const int p = (int)(t % a.P);
        const int b = (int)(t / a.P);
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kidx = k4 + j;
            float v = 0.f;
            if (kidx < a.Kc) {
                const int lw = kidx % a.kw;
                const int t2 = kidx / a.kw;
                const int kk = t2 % a.kh;
                const int c = t2 / a.kh;
                const int ny = p * a.sp0 - a.off0 + kk * a.dk0;
                const int nx = q * a.sp1 - a.off1 + lw * a.dk1;
                if (ny >= 0 && nx >= 0 && (ny % a.up0) == 0 && (nx % a.up1) == 0) {
                    const int yy = ny / a.up0, xx = nx / a.up1;
                    if (yy < a.Hs && xx < a.Ws)
                        v = a.src[(((long long)b * a.C + c) * a.Hs + yy) * a.Ws + xx];
                }
            }
            h[j] = __float2bfloat16_rn(v);
            l[j] = __float2bfloat16_rn(v - __bfloat162float(h[j]));
        }
        const long long o = m * a.ld + k4;  // ld % 8 == 0 and k4 % 4 == 0 -> 8-byte aligned
        *reinterpret_cast<uint2*>(a.hi + o) = *reinterpret_cast<const uint2*>(h);
        if (X3) *reinterpret_cast<uint2*>(a.lo + o) = *reinterpret_cast<const uint2*>(l);
    }
}

// Wr[i][(o, k', l')] = W[o][i][kh-1-k'][kw-1-l']   (rot180 + in/out swap, conv2d.py:91)
template <bool X3>
__global__ void stage_weight_dgrad_kernel(const float* __restrict__ W, int Cout, int Cin, int kh,
                                          int kw, long long ld, __nv_bfloat16* hi,
                                          __nv_bfloat16* lo) {
    const int khw = kh * kw;
    const long long total = (long long)Cin * Cout * khw;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int kl = (int)(idx % khw);
        const long long t = idx / khw;
        const int o = (int)(t % Cout);
        const int i = (int)(t / Cout);
        const float v = W[((long long)o * Cin + i) * khw + (khw - 1 - kl)];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const long long dst = (long long)i * ld + (long long)o * khw + kl;
        hi[dst] = h;
        if (X3) lo[dst] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

// g'[o][(b, hw)] = dO[b][o][hw]
template <bool X3>
__global__ void stage_nchw_to_c_bhw_kernel(const float* __restrict__ g, int B, int C, int HW,
                                           long long ld, __nv_bfloat16* hi, __nv_bfloat16* lo) {
    const long long total = (long long)B * C * HW;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int hw = (int)(idx % HW);
        const long long t = idx / HW;
        const int c = (int)(t % C);
        const int b = (int)(t / C);
        const float v = g[idx];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const long long dst = (long long)c * ld + (long long)b * HW + hw;
        hi[dst] = h;
        if (X3) lo[dst] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

// ---------------------------------------------------------------- channels-last fast path (C % 8 == 0)
// The generic gather above pays ~15 integer divisions per ELEMENT and reads NCHW with a 9-way scatter.
// When both channel counts are multiples of 8 the activations are first converted ONCE to bf16
// channels-last planes ([B*H*W][C], tiled transpose), and `col` is built in (k, l, c) order, so every
// (output position, tap) is one contiguous run of C bf16 copied with 16-byte accesses. The weights are
// staged in the same (k, l, c) order; wgrad's [Cout][(k,l,c)] result is permuted back to the reference's
// [Cout][Cin][kh][kw] by a small kernel.

// src fp32 [B][C][HW] -> dst bf16 [B][HW][C]   (32 x 32 tiles through shared memory)
// Cp >= C is the channel pitch of the planes (C rounded up to 8 so every pixel is a whole number of 16-byte groups);
// the pad channels are written as zeros.
template <bool X3>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ src, int C, int Cp, int HW,
                                                           __nv_bfloat16* __restrict__ hi,
                                                           __nv_bfloat16* __restrict__ lo) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const float* s = src + (long long)b * C * HW;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, hw = hw0 + tx;
        tile[r][tx] = (c < C && hw < HW) ? s[(long long)c * HW + hw] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int hw = hw0 + r, c = c0 + tx;
        if (hw < HW && c < Cp) {
            const float v = c < C ? tile[tx][r] : 0.f;
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            const long long o = ((long long)b * HW + hw) * Cp + c;
            hi[o] = h;
            if (X3) lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
    }
}

// col[(b,p,q)][(k,l,c)] = nhwc[b][yy][xx][c] (same tap geometry as GatherArgs); one thread = 8 channels
struct GatherNhwcArgs {
    const __nv_bfloat16* hi_src;
    const __nv_bfloat16* lo_src;
    int C, Hs, Ws, P, Q, kh, kw;
    int sp0, sp1, off0, off1, dk0, dk1, up0, up1;
    long long M;
    long long ld;
    __nv_bfloat16* hi;
    __nv_bfloat16* lo;
};

template <bool X3>
__global__ void __launch_bounds__(256) gather_nhwc_kernel(const GatherNhwcArgs a) {
    const int c8 = a.C >> 3;
    const int taps = a.kh * a.kw;
    const long long per_row = (long long)taps * c8;
    const long long total = a.M * per_row;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long m = idx / per_row;
        const int rem = (int)(idx - m * per_row);
        const int t = rem / c8;
        const int j = rem - t * c8;
        const int kk = t / a.kw, lw = t - kk * a.kw;
        const int q = (int)(m % a.Q);
        const long long t2 = m / a.Q;
        const int p = (int)(t2 % a.P);
        const int b = (int)(t2 / a.P);
        const int ny = p * a.sp0 - a.off0 + kk * a.dk0;
        const int nx = q * a.sp1 - a.off1 + lw * a.dk1;
        uint4 vh = make_uint4(0u, 0u, 0u, 0u), vl = vh;
        if (ny >= 0 && nx >= 0 && (ny % a.up0) == 0 && (nx % a.up1) == 0) {
            const int yy = ny / a.up0, xx = nx / a.up1;
            if (yy < a.Hs && xx < a.Ws) {
                const long long so = (((long long)b * a.Hs + yy) * a.Ws + xx) * a.C + j * 8;
                vh = *reinterpret_cast<const uint4*>(a.hi_src + so);
                if (X3) vl = *reinterpret_cast<const uint4*>(a.lo_src + so);
            }
        }
        const long long o = m * a.ld + (long long)t * a.C + j * 8;
        *reinterpret_cast<uint4*>(a.hi + o) = vh;
        if (X3) *reinterpret_cast<uint4*>(a.lo + o) = vl;
    }
}

// Wk[o][(k,l,c)] = W[o][c][k][l]; with `flip` (dgrad): Wr[i][(k',l',o)] = W[o][i][kh-1-k'][kw-1-l'].
// One thread per (staged row r, channel c): it reads its khw taps (adjacent threads read adjacent 4*khw-byte runs
// when the channel is W's inner index, i.e. without flip) and writes one bf16 per tap with c fastest (coalesced).
template <bool X3>
__global__ void stage_weight_klc_kernel(const float* __restrict__ W, int Cout, int Cin, int khw, int flip,
                                        long long ld, __nv_bfloat16* hi, __nv_bfloat16* lo) {
    const int Cc = flip ? Cout : Cin;   // channel (fastest) dimension of the staged rows
    const int Ccp = (Cc + 7) & ~7;      // ... padded to the channel pitch of the activation planes (zeros)
    const int R = flip ? Cin : Cout;    // staged rows
    const long long total = (long long)R * Ccp;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % Ccp);
        const int r = (int)(idx / Ccp);
        const int cs = c < Cc ? c : 0;
        const float* src = flip ? W + ((long long)cs * Cin + r) * khw : W + ((long long)r * Cin + cs) * khw;
        for (int tap = 0; tap < khw; ++tap) {
            const float v = c < Cc ? (flip ? src[khw - 1 - tap] : src[tap]) : 0.f;
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            const long long dst = (long long)r * ld + (long long)tap * Ccp + c;
            hi[dst] = h;
            if (X3) lo[dst] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
    }
}

// dW[o][c][tap] = T[o][(tap, c)]: one thread per (o, c) -- reads coalesced over c, writes its khw contiguous taps
__global__ void permute_dw_kernel(const float* __restrict__ T, int Cout, int Cin, int khw, float* __restrict__ dW) {
    const long long total = (long long)Cout * Cin;
    const int Cinp = (Cin + 7) & ~7;  // channel pitch of T's (tap, c) columns
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % Cin);
        const int o = (int)(idx / Cin);
        for (int tap = 0; tap < khw; ++tap) dW[idx * khw + tap] = T[((long long)o * khw + tap) * Cinp + c];
    }
}

// db[o] = sum_{b,h,w} dO[b,o,h,w]   (conv2d.py:94): grid = (channels, chunks of the batch); partial sums are
// added in a fixed order by direct_wgrad_finish_kernel (deterministic)
constexpr int CHANNEL_SUM_CHUNKS = 64;

__global__ void __launch_bounds__(256) channel_sum_kernel(const float* __restrict__ g, int B, int C,
                                                          int HW, float* __restrict__ partial) {
    const int c = blockIdx.x;
    const int bper = (B + gridDim.y - 1) / gridDim.y;
    const int b0 = blockIdx.y * bper, b1 = min(b0 + bper, B);
    float s = 0.f;
    // the chunk's (image, pixel) pairs are walked as ONE flat range so small feature maps (4 x 4) still use every lane
    const long long n = (long long)(b1 - b0) * HW;
    if ((HW % 4) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        for (long long e = (long long)threadIdx.x * 4; e < n; e += 1024) {
            const int b = b0 + (int)(e / HW);
            const int i = (int)(e - (long long)(b - b0) * HW);
            const float4 v = *reinterpret_cast<const float4*>(g + ((long long)b * C + c) * HW + i);
            s += (v.x + v.y) + (v.z + v.w);
        }
    } else {
        for (long long e = threadIdx.x; e < n; e += 256) {
            const int b = b0 + (int)(e / HW);
            const int i = (int)(e - (long long)(b - b0) * HW);
            s += g[((long long)b * C + c) * HW + i];
        }
    }
    __shared__ float red[256];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(long long)blockIdx.y * C + c] = red[0];
}

// ---------------------------------------------------------------- direct fp32 kernels (tiny channel counts)
__global__ void __launch_bounds__(256)
direct_fwd_kernel(const Geo g, const float* __restrict__ X, const float* __restrict__ Wt,
                  const float* __restrict__ bias, float* __restrict__ O) {
    const long long total = (long long)g.B * g.Cout * g.Ho * g.Wo;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int wo = (int)(idx % g.Wo);
        long long t = idx / g.Wo;
        const int ho = (int)(t % g.Ho); t /= g.Ho;
        const int o = (int)(t % g.Cout);
        const int b = (int)(t / g.Cout);
        float acc = 0.f;
        for (int i = 0; i < g.Cin; ++i) {
            const float* xp = X + ((long long)b * g.Cin + i) * g.H * g.W;
            const float* wp = Wt + ((long long)o * g.Cin + i) * g.kh * g.kw;
            for (int k = 0; k < g.kh; ++k) {
                const int y = ho * g.s0 - g.pt + k * g.d0;
                if (y < 0 || y >= g.H) continue;
                for (int l = 0; l < g.kw; ++l) {
                    const int x = wo * g.s1 - g.pl + l * g.d1;
                    if (x < 0 || x >= g.W) continue;
                    acc += xp[y * g.W + x] * wp[k * g.kw + l];
                }
            }
        }
        if (bias) acc += bias[o];
        O[idx] = acc;
    }
}

__global__ void __launch_bounds__(256)
direct_dgrad_kernel(const Geo g, const float* __restrict__ dO, const float* __restrict__ Wt,
                    float* __restrict__ dX) {
    const long long total = (long long)g.B * g.Cin * g.H * g.W;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % g.W);
        long long t = idx / g.W;
        const int y = (int)(t % g.H); t /= g.H;
        const int i = (int)(t % g.Cin);
        const int b = (int)(t / g.Cin);
        float acc = 0.f;
        for (int k = 0; k < g.kh; ++k) {
            const int ny = y + g.pt - k * g.d0;
            if (ny < 0 || ny % g.s0) continue;
            const int ho = ny / g.s0;
            if (ho >= g.Ho) continue;
            for (int l = 0; l < g.kw; ++l) {
                const int nx = x + g.pl - l * g.d1;
                if (nx < 0 || nx % g.s1) continue;
                const int wo = nx / g.s1;
                if (wo >= g.Wo) continue;
                for (int o = 0; o < g.Cout; ++o)
                    acc += dO[(((long long)b * g.Cout + o) * g.Ho + ho) * g.Wo + wo] *
                           Wt[(((long long)o * g.Cin + i) * g.kh + k) * g.kw + l];
            }
        }
        dX[idx] = acc;
    }
}

// grid = (weight elements (o, i, k, l), chunks of the (b, ho, wo) reduction): every block writes one partial
// sum, direct_wgrad_finish_kernel adds the chunks in a fixed order (deterministic, no atomics)
constexpr int DIRECT_WGRAD_CHUNKS = 64;

__global__ void __launch_bounds__(256)
direct_wgrad_kernel(const Geo g, const float* __restrict__ X, const float* __restrict__ dO,
                    float* __restrict__ partial) {
    int idx = blockIdx.x;
    const int l = idx % g.kw; idx /= g.kw;
    const int k = idx % g.kh; idx /= g.kh;
    const int i = idx % g.Cin;
    const int o = idx / g.Cin;
    const int HWo = g.Ho * g.Wo;
    const long long n = (long long)g.B * HWo;
    const long long per = (n + gridDim.y - 1) / gridDim.y;
    const long long e0 = (long long)blockIdx.y * per, e1 = min(e0 + per, n);
    float s = 0.f;
    for (long long e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        const int b = (int)(e / HWo);
        const int r = (int)(e - (long long)b * HWo);
        const int ho = r / g.Wo, wo = r - ho * g.Wo;
        const int y = ho * g.s0 - g.pt + k * g.d0;
        const int x = wo * g.s1 - g.pl + l * g.d1;
        if (y >= 0 && y < g.H && x >= 0 && x < g.W)
            s += X[(((long long)b * g.Cin + i) * g.H + y) * g.W + x] *
                 dO[((long long)b * g.Cout + o) * HWo + r];
    }
    __shared__ float red[256];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(long long)blockIdx.y * gridDim.x + blockIdx.x] = red[0];
}

__global__ void direct_wgrad_finish_kernel(const float* __restrict__ partial, int nw, int chunks,
                                           float* __restrict__ dW) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nw) return;
    float s = 0.f;
    for (int c = 0; c < chunks; ++c) s += partial[(long long)c * nw + w];
    dW[w] = s;
}

// ---------------------------------------------------------------- native ConvTranspose2d (row N1 of SURVEY.md 8f)
// The reference defines ConvTranspose2d as a stride-1 correlation of the UN-flipped (out,in,kh,kw) kernel over the
// zero-stuffed, (k-1)-padded, padding-cropped input (neunet/nn/layers/convtranspose2d.py:165-181, 321), so at
// stride s only 1 / (s0*s1) of the multiplied taps are non-zero. Gather form with real taps only: output pixels are
// split by parity class (y % s0, x % s1); inside a class every pixel uses the same tap subset
//   k = k0 + j*s  with  k0 = (-(py + pad - (kh-1))) mod s,   source row u + j + e,  e = (py + pad - (kh-1) + k0) / s
// i.e. a stride-1 correlation of X with an (nk0 x nk1)-tap kernel -> one implicit GEMM per class.

// Wc[o][(j0, j1, c)] = Wt[o][c][k0y + j0*s0][k0x + j1*s1], Wt = the (out, in, kh, kw) kernel of the TRANSPOSED conv.
// swap = 0: Wt is W itself (nn.ConvTranspose2d forward); swap = 1: Wt[o][c][k][l] = W[c][o][kh-1-k][kw-1-l] with
// W stored (Cin_t = c rows, ...) i.e. the flipped, in/out-swapped kernel of a strided nn.Conv2d whose dgrad this is.
template <bool X3>
__global__ void stage_weight_class_kernel(const float* __restrict__ W, int Cout, int Cin, int kh, int kw, int k0y, int k0x,
                                          int s0, int s1, int nk0, int nk1, int swap, long long ld, __nv_bfloat16* hi,
                                          __nv_bfloat16* lo) {
    const long long total = (long long)Cout * nk0 * nk1 * Cin;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % Cin);
        long long t = idx / Cin;
        const int j1 = (int)(t % nk1); t /= nk1;
        const int j0 = (int)(t % nk0);
        const int o = (int)(t / nk0);
        const int k = k0y + j0 * s0, l = k0x + j1 * s1;
        const float v = swap ? W[(((long long)c * Cout + o) * kh + (kh - 1 - k)) * kw + (kw - 1 - l)]
                             : W[(((long long)o * Cin + c) * kh + k) * kw + l];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const long long dst = (long long)o * ld + ((long long)j0 * nk1 + j1) * Cin + c;
        hi[dst] = h;
        if (X3) lo[dst] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

// out[b][o][s0*u + py][s1*v + px] = T[(py, px)][o][(b*P + u)*Q + v] + bias[o]
__global__ void __launch_bounds__(256) convT_interleave_kernel(const float* __restrict__ T, const float* __restrict__ bias,
                                                               int B, int Cout, int P, int Q, int s0, int s1,
                                                               float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int Ho = P * s0, Wo = Q * s1;
    const long long N = (long long)B * P * Q;
    const long long total = (long long)B * Cout * Ho * Wo;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % Wo);
        long long t = idx / Wo;
        const int y = (int)(t % Ho); t /= Ho;
        const int o = (int)(t % Cout);
        const int b = (int)(t / Cout);
        const int cls = (y % s0) * s1 + (x % s1);
        float v = T[((long long)cls * Cout + o) * N + ((long long)b * P + y / s0) * Q + x / s1];
        if (bias != nullptr) v += bias[o];
        out[idx] = v;
    }
}

// The same for stride (s0, 2) with Wo % 4 == 0 and < 2^31 elements (every transposed conv of the DDPM UNet): one thread
// per FOUR adjacent output pixels = two adjacent class positions of the px = 0 and px = 1 planes -> two 8-byte loads, one
// 16-byte store, 32-bit index arithmetic (the generic kernel spends its time in six 64-bit divisions per element).
__global__ void __launch_bounds__(256) convT_interleave_s2_kernel(const float* __restrict__ T, const float* __restrict__ bias,
                                                                  unsigned B, unsigned Cout, unsigned P, unsigned Q, unsigned s0,
                                                                  float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const unsigned Ho = P * s0, W4 = Q / 2;  // Wo / 4 quads per output row
    const unsigned N = B * P * Q;
    const unsigned total4 = B * Cout * Ho * W4;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total4; idx += gridDim.x * blockDim.x) {
        const unsigned xq = idx % W4;
        unsigned t = idx / W4;
        const unsigned y = t % Ho; t /= Ho;
        const unsigned o = t % Cout;
        const unsigned b = t / Cout;
        const unsigned py = y % s0, u = y / s0;
        const size_t n = (size_t)(b * P + u) * Q + 2 * xq;                     // class position of the quad's first pixel
        const size_t plane0 = ((size_t)(py * 2) * Cout + o) * N + n;           // px = 0 plane; px = 1 is Cout * N further
        const float2 c0 = *reinterpret_cast<const float2*>(T + plane0);
        const float2 c1 = *reinterpret_cast<const float2*>(T + plane0 + (size_t)Cout * N);
        const float bo = bias != nullptr ? bias[o] : 0.f;
        *reinterpret_cast<float4*>(out + (size_t)idx * 4) = make_float4(c0.x + bo, c1.x + bo, c0.y + bo, c1.y + bo);
    }
}

// dW[o][c][kh-1-k'][kw-1-l'] = T[c][((k'*kw + l')*Cout + o)]: one thread per (c, o), reads coalesced over o
__global__ void permute_dw_convT_kernel(const float* __restrict__ T, int Cout, int Cin, int kh, int kw, float* __restrict__ dW) {
    const int khw = kh * kw;
    const long long total = (long long)Cout * Cin;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(idx % Cout);
        const int c = (int)(idx / Cout);
        float* dst = dW + ((long long)o * Cin + c) * khw;
        for (int tap = 0; tap < khw; ++tap) dst[tap] = T[((long long)c * khw + (khw - 1 - tap)) * Cout + o];
    }
}

struct TGeo {          // transposed-conv geometry: X (B,Cin,H,W) -> O (B,Cout,Ho,Wo)
    Geo g;             // g.H/W = input, g.Ho/Wo = OUTPUT of the transposed conv, g.pt/pl = padding, s, d, kh, kw
    int P, Q;          // per-class position grid = Ho / s0, Wo / s1
};

int make_tgeo(const nnb_conv2d_desc* d, int op0, int op1, TGeo* t) {
    NNB_REQUIRE(d, "conv_transpose2d: null descriptor");
    NNB_REQUIRE(d->B > 0 && d->Cin > 0 && d->H > 0 && d->W > 0 && d->Cout > 0 && d->kh > 0 && d->kw > 0,
                "conv_transpose2d: non-positive dimension");
    NNB_REQUIRE(d->stride[0] > 0 && d->stride[1] > 0 && d->dil[0] > 0 && d->dil[1] > 0, "conv_transpose2d: bad stride/dilation");
    Geo& g = t->g;
    g.B = (int)d->B; g.Cin = (int)d->Cin; g.H = (int)d->H; g.W = (int)d->W; g.Cout = (int)d->Cout;
    g.kh = (int)d->kh; g.kw = (int)d->kw; g.s0 = d->stride[0]; g.s1 = d->stride[1];
    g.pt = d->pad[0]; g.pl = d->pad[2]; g.d0 = d->dil[0]; g.d1 = d->dil[1];
    const int dk0 = g.d0 * (g.kh - 1) + 1, dk1 = g.d1 * (g.kw - 1) + 1;
    // convtranspose2d.py:165-181: stuffed size s*H - (s-1) + output_padding, padded by (dk-1) each side, cropped by the padding,
    // then a stride-1 valid correlation
    const long long ho = (long long)g.s0 * g.H - (g.s0 - 1) + op0 + 2 * (dk0 - 1) - d->pad[0] - d->pad[1] - (dk0 - 1);
    const long long wo = (long long)g.s1 * g.W - (g.s1 - 1) + op1 + 2 * (dk1 - 1) - d->pad[2] - d->pad[3] - (dk1 - 1);
    NNB_REQUIRE(ho > 0 && wo > 0, "conv_transpose2d: empty output");
    g.Ho = (int)ho; g.Wo = (int)wo;
    t->P = g.Ho / g.s0; t->Q = g.Wo / g.s1;
    return NNB_OK;
}

bool tconv_native(const TGeo& t, int op0, int op1) {
    const Geo& g = t.g;
    return g_implicit && op0 == 0 && op1 == 0 && g.d0 == 1 && g.d1 == 1 && (g.Cin % 64) == 0 && (g.Cout % 64) == 0 &&
           (g.Ho % g.s0) == 0 && (g.Wo % g.s1) == 0 && g.kh >= g.s0 && g.kw >= g.s1 &&
           gemm_conv_supported(1, g.Cin, t.P, t.Q, 1, 1) &&            // forward classes (positions = class grid)
           gemm_conv_supported(1, g.Cout, g.H, g.W, g.s0, g.s1) &&     // dgrad (positions = input pixels, strided reads of dO)
           gemm_conv_supported(2, g.Cout, g.H, g.W, g.s0, g.s1);       // wgrad
}

// taps of parity class `par` along one axis
void class_taps(int par, int pad, int k, int s, int* k0, int* nk, int* e) {
    int r = (par + pad - (k - 1)) % s;
    if (r < 0) r += s;
    *k0 = (s - r) % s;
    *nk = (k - *k0 + s - 1) / s;
    *e = (par + pad - (k - 1) + *k0) / s;  // exact
}

// ---------------------------------------------------------------- tiny-channel layers, one image per block in shared memory
// The direct kernels above read every operand from global memory per multiply; for the MNIST-classifier layers
// (1 -> 8 and 8 -> 16 channels, 28 x 28 / 14 x 14) a whole zero-padded image (all input channels), its output (or output
// gradient) and the kernel fit in shared memory, so a block stages them once and every multiply reads shared memory.
// wgrad keeps a thread's weight elements in registers ACROSS the images the block walks, so the partial-sum volume is
// (blocks x weights), not (images x weights). Deterministic: fixed image order per block, fixed-order finish.
struct TileGeo {
    int Hp, Wp;       // padded image extent held in shared memory (top/left pad + what the taps reach)
    int imgs_per_block;
};

__device__ __forceinline__ void load_padded_image(float* Xs, const float* __restrict__ X, const Geo& g, int Hp, int Wp, int b) {
    for (int e = threadIdx.x; e < g.Cin * Hp * Wp; e += blockDim.x) {
        const int x = e % Wp, y = (e / Wp) % Hp, i = e / (Wp * Hp);
        const int yy = y - g.pt, xx = x - g.pl;
        Xs[e] = (yy >= 0 && yy < g.H && xx >= 0 && xx < g.W) ? X[(((long long)b * g.Cin + i) * g.H + yy) * g.W + xx] : 0.f;
    }
}

__global__ void __launch_bounds__(256)
tile_fwd_kernel(const Geo g, const TileGeo t, const float* __restrict__ X, const float* __restrict__ Wt,
                const float* __restrict__ bias, float* __restrict__ O) {
    extern __shared__ __align__(16) float smf[];
    float* Ws = smf;                                   // [Cout][Cin][kh][kw]
    float* Xs = Ws + g.Cout * g.Cin * g.kh * g.kw;     // [Cin][Hp][Wp]
    const int nw = g.Cout * g.Cin * g.kh * g.kw, HWo = g.Ho * g.Wo, khw = g.kh * g.kw;
    for (int e = threadIdx.x; e < nw; e += blockDim.x) Ws[e] = Wt[e];
    for (int b = blockIdx.x; b < g.B; b += gridDim.x) {
        __syncthreads();
        load_padded_image(Xs, X, g, t.Hp, t.Wp, b);
        __syncthreads();
        for (int e = threadIdx.x; e < g.Cout * HWo; e += blockDim.x) {
            const int pos = e % HWo, o = e / HWo;
            const int ho = pos / g.Wo, wo = pos - ho * g.Wo;
            float acc = bias ? bias[o] : 0.f;
            const float* wp = Ws + o * g.Cin * khw;
            for (int i = 0; i < g.Cin; ++i) {
                const float* xp = Xs + (i * t.Hp + ho * g.s0) * t.Wp + wo * g.s1;
                for (int k = 0; k < g.kh; ++k)
                    for (int l = 0; l < g.kw; ++l) acc = fmaf(xp[k * g.d0 * t.Wp + l * g.d1], wp[(i * g.kh + k) * g.kw + l], acc);
            }
            O[((long long)b * g.Cout + o) * HWo + pos] = acc;
        }
    }
}

// partial[block][w] = sum over the block's images and all positions of X[.., ho*s + k*d, wo*s + l*d] * dO[.., o, ho, wo]
__global__ void __launch_bounds__(256)
tile_wgrad_kernel(const Geo g, const TileGeo t, const float* __restrict__ X, const float* __restrict__ dO,
                  float* __restrict__ partial) {
    extern __shared__ __align__(16) float smf[];
    float* Gs = smf;                     // [Cout][Ho*Wo]
    float* Xs = Gs + g.Cout * g.Ho * g.Wo;  // [Cin][Hp][Wp]
    const int nw = g.Cout * g.Cin * g.kh * g.kw, HWo = g.Ho * g.Wo;
    // a thread owns weights w = tid, tid + 256, ... (<= 4 of them: nw <= 31 * 256 / ... checked on the host)
    constexpr int MAXW = 32;
    float acc[MAXW];
#pragma unroll
    for (int j = 0; j < MAXW; ++j) acc[j] = 0.f;
    // when there are fewer weights than threads, L threads share a weight and split the positions
    const int L = nw < 256 ? 256 / nw : 1;
    const int lanes = nw < 256 ? nw * L : 256;
    for (int b = blockIdx.x; b < g.B; b += gridDim.x) {
        __syncthreads();
        for (int e = threadIdx.x; e < g.Cout * HWo; e += blockDim.x) Gs[e] = dO[(long long)b * g.Cout * HWo + e];
        load_padded_image(Xs, X, g, t.Hp, t.Wp, b);
        __syncthreads();
        if ((int)threadIdx.x < lanes) {
            const int sub = nw < 256 ? (int)threadIdx.x / nw : 0;
            int j = 0;
            for (int w = nw < 256 ? (int)threadIdx.x % nw : (int)threadIdx.x; w < nw; w += 256, ++j) {
                const int l = w % g.kw, k = (w / g.kw) % g.kh, i = (w / (g.kw * g.kh)) % g.Cin, o = w / (g.kw * g.kh * g.Cin);
                const float* gp = Gs + o * HWo;
                const float* xp = Xs + (i * t.Hp + k * g.d0) * t.Wp + l * g.d1;
                float s = 0.f;
                for (int pos = sub; pos < HWo; pos += L) {
                    const int ho = pos / g.Wo, wo = pos - ho * g.Wo;
                    s = fmaf(xp[ho * g.s0 * t.Wp + wo * g.s1], gp[pos], s);
                }
                acc[j] += s;
            }
        }
    }
    // partial row of this block: [L sub-lanes][nw] (the finish kernel adds blocks and sub-lanes in a fixed order)
    float* prow = partial + (long long)blockIdx.x * L * nw;
    if ((int)threadIdx.x < lanes) {
        const int sub = nw < 256 ? (int)threadIdx.x / nw : 0;
        int j = 0;
        for (int w = nw < 256 ? (int)threadIdx.x % nw : (int)threadIdx.x; w < nw; w += 256, ++j) prow[(long long)sub * nw + w] = acc[j];
    }
}

__global__ void __launch_bounds__(256)
tile_dgrad_kernel(const Geo g, const float* __restrict__ dO, const float* __restrict__ Wt, float* __restrict__ dX) {
    extern __shared__ __align__(16) float smf[];
    float* Ws = smf;                                   // [Cout][Cin][kh][kw]
    float* Gs = Ws + g.Cout * g.Cin * g.kh * g.kw;     // [Cout][Ho*Wo]
    const int nw = g.Cout * g.Cin * g.kh * g.kw, HWo = g.Ho * g.Wo, HW = g.H * g.W, khw = g.kh * g.kw;
    for (int e = threadIdx.x; e < nw; e += blockDim.x) Ws[e] = Wt[e];
    for (int b = blockIdx.x; b < g.B; b += gridDim.x) {
        __syncthreads();
        for (int e = threadIdx.x; e < g.Cout * HWo; e += blockDim.x) Gs[e] = dO[(long long)b * g.Cout * HWo + e];
        __syncthreads();
        for (int e = threadIdx.x; e < g.Cin * HW; e += blockDim.x) {
            const int x = e % g.W, y = (e / g.W) % g.H, i = e / HW;
            float acc = 0.f;
            for (int k = 0; k < g.kh; ++k) {
                const int ny = y + g.pt - k * g.d0;
                if (ny < 0 || ny % g.s0) continue;
                const int ho = ny / g.s0;
                if (ho >= g.Ho) continue;
                for (int l = 0; l < g.kw; ++l) {
                    const int nx = x + g.pl - l * g.d1;
                    if (nx < 0 || nx % g.s1) continue;
                    const int wo = nx / g.s1;
                    if (wo >= g.Wo) continue;
                    const float* gp = Gs + ho * g.Wo + wo;
                    const float* wp = Ws + (i * g.kh + k) * g.kw + l;
                    for (int o = 0; o < g.Cout; ++o) acc = fmaf(gp[o * HWo], wp[o * g.Cin * khw], acc);
                }
            }
            dX[(long long)b * g.Cin * HW + e] = acc;
        }
    }
}

constexpr size_t TILE_SMEM_LIMIT = 160 * 1024;
constexpr int TILE_MAX_BLOCKS = 296;  // two per SM: each block walks B / blocks images

bool tile_geo(const Geo& g, TileGeo* t) {
    t->Hp = g.pt + std::max(g.H, (g.Ho - 1) * g.s0 + (g.kh - 1) * g.d0 + 1);
    t->Wp = g.pl + std::max(g.W, (g.Wo - 1) * g.s1 + (g.kw - 1) * g.d1 + 1);
    const size_t nw = (size_t)g.Cout * g.Cin * g.kh * g.kw;
    const size_t img = (size_t)g.Cin * t->Hp * t->Wp, out = (size_t)g.Cout * g.Ho * g.Wo;
    t->imgs_per_block = 1;
    return nw <= 32 * 256 && (std::max(nw, out) + std::max(img, out)) * 4 <= TILE_SMEM_LIMIT;
}

int tile_blocks(const Geo& g) { return std::min(g.B, TILE_MAX_BLOCKS); }
int tile_wgrad_lanes(const Geo& g) { const int nw = g.Cout * g.Cin * g.kh * g.kw; return nw < 256 ? 256 / nw : 1; }

template <typename K>
int tile_smem_attr(K kernel, size_t bytes) {
    NNB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_SMEM_LIMIT));
    (void)bytes;
    return NNB_OK;
}

int grid_for(long long n, int threads) {
    return (int)std::max<long long>(1, std::min<long long>(ceil_div(n, threads), (long long)num_sms() * 32));
}

int make_geo(const nnb_conv2d_desc* d, Geo* g) {
    NNB_REQUIRE(d, "conv2d: null descriptor");
    NNB_REQUIRE(d->B > 0 && d->Cin > 0 && d->H > 0 && d->W > 0 && d->Cout > 0 && d->kh > 0 && d->kw > 0,
                "conv2d: non-positive dimension");
    NNB_REQUIRE(d->stride[0] > 0 && d->stride[1] > 0 && d->dil[0] > 0 && d->dil[1] > 0, "conv2d: bad stride/dilation");
    NNB_REQUIRE(d->pad[0] >= 0 && d->pad[1] >= 0 && d->pad[2] >= 0 && d->pad[3] >= 0, "conv2d: negative padding");
    const int64_t eh = d->H + d->pad[0] + d->pad[1] - d->dil[0] * (d->kh - 1) - 1;
    const int64_t ew = d->W + d->pad[2] + d->pad[3] - d->dil[1] * (d->kw - 1) - 1;
    NNB_REQUIRE(eh >= 0 && ew >= 0, "conv2d: kernel larger than padded input");
    g->B = (int)d->B; g->Cin = (int)d->Cin; g->H = (int)d->H; g->W = (int)d->W; g->Cout = (int)d->Cout;
    g->kh = (int)d->kh; g->kw = (int)d->kw;
    g->Ho = (int)(eh / d->stride[0] + 1);  // conv2d.py:246-260
    g->Wo = (int)(ew / d->stride[1] + 1);
    g->s0 = d->stride[0]; g->s1 = d->stride[1];
    g->pt = d->pad[0]; g->pl = d->pad[2];
    g->d0 = d->dil[0]; g->d1 = d->dil[1];
    NNB_REQUIRE((int64_t)g->B * g->Ho * g->Wo < (1ll << 31) && (int64_t)g->B * g->H * g->W < (1ll << 31),
                "conv2d: too many output positions");
    return NNB_OK;
}

// Direct fp32 CUDA-core kernels only when BOTH the output channels and the reduction are tiny (MNIST classifier: 1 -> 8,
// 8 -> 16): a 128-row tensor tile would be > 90 % padding there. A small Cout over a long reduction (DDPM's 128 -> 3
// output layer, K = 1152) still belongs on the tensor cores: measured 2.2 ms per step in the direct kernels (round 2).
bool use_direct(const Geo& g) { return g.Cout < 32 && (int64_t)g.Cin * g.kh * g.kw <= 256; }

int run_gather(const GatherArgs& a, bool x3, cudaStream_t stream) {
    const long long total = a.M * ((a.Kc + 3) / 4);
    if (x3) gather_cols_kernel<true><<<grid_for(total, 256), 256, 0, stream>>>(a);
    else gather_cols_kernel<false><<<grid_for(total, 256), 256, 0, stream>>>(a);
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

struct Planes {
    __nv_bfloat16* hi;
    __nv_bfloat16* lo;
};
Planes take_planes(Bump& ws, int64_t rows, int64_t cols, int prec) {
    Planes p;
    p.hi = static_cast<__nv_bfloat16*>(ws.take(staged_plane_bytes(1, rows, cols)));
    p.lo = prec == NNB_PREC_BF16X3 ? static_cast<__nv_bfloat16*>(ws.take(staged_plane_bytes(1, rows, cols))) : nullptr;
    return p;
}
Staged as_staged(const Planes& p, int64_t rows, int64_t cols) {
    Staged s;
    s.hi = p.hi; s.lo = p.lo; s.rows = rows; s.cols = cols; s.ld = staged_ld(cols);
    s.batch = 1; s.batch_stride = rows * s.ld;
    return s;
}

// Channels-last planes for every tensor-core layer: channel counts that are not multiples of 8 (the 3-channel image of
// the DDPM input / output layers) get a zero-padded pitch; NNB_CONV_NHWC=0 selects the generic NCHW gather instead.
int g_nhwc = [] { const char* e = getenv("NNB_CONV_NHWC"); return e ? atoi(e) : 1; }();
bool use_nhwc(const Geo& g) { return !use_direct(g) && (g_nhwc || ((g.Cin % 8) == 0 && (g.Cout % 8) == 0)); }
int cpad(int c) { return (c + 7) & ~7; }

// Implicit GEMM (no materialised `col`): the GEMM's B operand is read straight from the channels-last planes through
// 4-D TMA boxes (gemm.cu, ConvOperand). Needs 64-channel k-blocks and a position grid that tiles into pixel boxes.
bool implicit_fwd(const Geo& g) {
    return g_implicit && use_nhwc(g) && (g.Cin % 64) == 0 && gemm_conv_supported(1, g.Cin, g.Ho, g.Wo, g.s0, g.s1) &&
           gemm_conv_supported(2, g.Cin, g.Ho, g.Wo, g.s0, g.s1);
}
bool implicit_dgrad(const Geo& g) {  // stride-1 layers: dX is a plain correlation of dO with the flipped kernel
    return g_implicit && use_nhwc(g) && (g.Cout % 64) == 0 && g.s0 == 1 && g.s1 == 1 &&
           gemm_conv_supported(1, g.Cout, g.H, g.W, 1, 1);
}

ConvOperand conv_operand(int mode, const Planes& src, int64_t B, int64_t C, int64_t Hs, int64_t Ws, int64_t P, int64_t Q,
                         int kh, int kw, int sp0, int sp1, int off0, int off1, int dk0, int dk1) {
    ConvOperand c;
    c.mode = mode; c.hi = src.hi; c.lo = src.lo;
    c.B = B; c.C = C; c.Hs = Hs; c.Ws = Ws; c.P = P; c.Q = Q; c.kh = kh; c.kw = kw;
    c.sp0 = sp0; c.sp1 = sp1; c.off0 = off0; c.off1 = off1; c.dk0 = dk0; c.dk1 = dk1;
    return c;
}

size_t nhwc_bytes(int64_t positions, int64_t C) { return (size_t)round_up(positions * round_up(C, 8) * 2, 256); }
Planes planes_at(void* buf, int64_t positions, int64_t C, int prec) {
    Planes p;
    p.hi = static_cast<__nv_bfloat16*>(buf);
    p.lo = prec == NNB_PREC_BF16X3 ? reinterpret_cast<__nv_bfloat16*>(static_cast<uint8_t*>(buf) + nhwc_bytes(positions, C)) : nullptr;
    return p;
}

Planes take_nhwc(Bump& ws, int64_t positions, int64_t C, int prec) {
    Planes p;
    const size_t b = (size_t)round_up(positions * round_up(C, 8) * 2, 256);
    p.hi = static_cast<__nv_bfloat16*>(ws.take(b));
    p.lo = prec == NNB_PREC_BF16X3 ? static_cast<__nv_bfloat16*>(ws.take(b)) : nullptr;
    return p;
}

int run_to_nhwc(const float* src, int B, int C, int HW, const Planes& dst, bool x3, cudaStream_t stream) {
    const int Cp = (int)round_up(C, 8);
    dim3 grid((unsigned)ceil_div(HW, 32), (unsigned)ceil_div(Cp, 32), (unsigned)B);
    NNB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "conv2d: too many channels / images for the layout pass");
    if (x3) nchw_to_nhwc_kernel<true><<<grid, 256, 0, stream>>>(src, C, Cp, HW, dst.hi, dst.lo);
    else nchw_to_nhwc_kernel<false><<<grid, 256, 0, stream>>>(src, C, Cp, HW, dst.hi, dst.lo);
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

int run_gather_nhwc(const GatherNhwcArgs& a, bool x3, cudaStream_t stream) {
    const long long total = a.M * a.kh * a.kw * (a.C / 8);
    if (x3) gather_nhwc_kernel<true><<<grid_for(total, 256), 256, 0, stream>>>(a);
    else gather_nhwc_kernel<false><<<grid_for(total, 256), 256, 0, stream>>>(a);
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

int run_stage_weight_klc(const float* W, const Geo& g, bool flip, const Planes& dst, int64_t ld, bool x3,
                         cudaStream_t stream) {
    const long long total = (long long)round_up(g.Cout, 8) * round_up(g.Cin, 8);
    if (x3) stage_weight_klc_kernel<true><<<grid_for(total, 256), 256, 0, stream>>>(W, g.Cout, g.Cin, g.kh * g.kw, flip ? 1 : 0, ld, dst.hi, dst.lo);
    else stage_weight_klc_kernel<false><<<grid_for(total, 256), 256, 0, stream>>>(W, g.Cout, g.Cin, g.kh * g.kw, flip ? 1 : 0, ld, dst.hi, dst.lo);
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}


// Gather-form transposed convolution over ready channels-last planes `xh` of the input (B, Cin, H, W): s0*s1 parity-class
// implicit GEMMs into T, then the interleave (+ bias) into O (B, Cout, Ho, Wo). Shared by nn.ConvTranspose2d forward
// (swap = 0) and by the dgrad of a strided nn.Conv2d (swap = 1: "input" = dO, weights flipped and in/out swapped).
int tconv_classes(const TGeo& t, const Planes& xh, const float* Wt, int swap, const float* bias, float* O, int prec, Bump ws,
                  cudaStream_t stream) {
    const Geo& g = t.g;
    const bool x3 = prec == NNB_PREC_BF16X3;
    const int64_t N = (int64_t)g.B * t.P * t.Q;
    float* T = static_cast<float*>(ws.take((size_t)round_up((int64_t)g.s0 * g.s1 * g.Cout * N * 4, 256)));
    if (!ws.ok()) return fail(NNB_ERR_WORKSPACE, "transposed conv: workspace too small (need >= %zu)", ws.off);
    for (int py = 0; py < g.s0; ++py) {
        for (int px = 0; px < g.s1; ++px) {
            int k0y, nk0, ey, k0x, nk1, ex;
            class_taps(py, g.pt, g.kh, g.s0, &k0y, &nk0, &ey);
            class_taps(px, g.pl, g.kw, g.s1, &k0x, &nk1, &ex);
            const int64_t Kc = (int64_t)nk0 * nk1 * g.Cin;
            Bump cws = ws;  // per-class scratch is reused: the stream serialises the classes
            Planes wc = take_planes(cws, g.Cout, Kc, prec);
            if (!cws.ok()) return fail(NNB_ERR_WORKSPACE, "transposed conv: workspace too small (need >= %zu)", cws.off);
            const long long total = (long long)g.Cout * Kc;
            if (x3) stage_weight_class_kernel<true><<<grid_for(total, 256), 256, 0, stream>>>(Wt, g.Cout, g.Cin, g.kh, g.kw, k0y, k0x, g.s0, g.s1, nk0, nk1, swap, staged_ld(Kc), wc.hi, wc.lo);
            else stage_weight_class_kernel<false><<<grid_for(total, 256), 256, 0, stream>>>(Wt, g.Cout, g.Cin, g.kh, g.kw, k0y, k0x, g.s0, g.s1, nk0, nk1, swap, staged_ld(Kc), wc.hi, wc.lo);
            count_launch();
            NNB_CUDA_OK(cudaGetLastError());
            GemmProblem p;
            p.M = g.Cout; p.N = N; p.K = Kc;
            p.A.st = as_staged(wc, g.Cout, Kc);
            p.conv = conv_operand(1, xh, g.B, g.Cin, g.H, g.W, t.P, t.Q, nk0, nk1, 1, 1, -ey, -ex, 1, 1);
            p.D = T + ((int64_t)(py * g.s1 + px) * g.Cout) * N; p.ldd = N;
            p.splitk_ws_bytes = cws.remaining();
            p.splitk_ws = static_cast<float*>(cws.take(p.splitk_ws_bytes));
            int rc = gemm(p, stream);
            if (rc) return rc;
        }
    }
    const long long total = (long long)g.B * g.Cout * g.Ho * g.Wo;
    const long long npos = (long long)g.B * t.P * t.Q;
    if (g.s1 == 2 && (t.Q % 2) == 0 && total < (1ll << 31) && (long long)g.s0 * 2 * g.Cout * npos < (1ll << 40) &&
        ((reinterpret_cast<uintptr_t>(T) | reinterpret_cast<uintptr_t>(O)) & 15) == 0) {
        NNB_CUDA_OK(launch_pdl(convT_interleave_s2_kernel, dim3(grid_for(total / 4, 256)), dim3(256), 0, stream, (const float*)T, bias,
                               (unsigned)g.B, (unsigned)g.Cout, (unsigned)t.P, (unsigned)t.Q, (unsigned)g.s0, O));
    } else {
        NNB_CUDA_OK(launch_pdl(convT_interleave_kernel, dim3(grid_for(total, 256)), dim3(256), 0, stream, (const float*)T, bias, g.B,
                               g.Cout, t.P, t.Q, g.s0, g.s1, O));
    }
    count_launch();
    NNB_CUDA_OK(cudaGetLastError());
    return NNB_OK;
}

// The dgrad of a strided nn.Conv2d IS a transposed convolution of dO (conv2d.py:35-96 builds it by zero-stuffing): same
// stride and padding, kernel flipped with in/out swapped. Geometry of that transposed conv, or false when the class form
// does not apply (then the zero-stuffed gather is used).
bool strided_dgrad_as_tconv(const Geo& g, TGeo* t) {
    if (!g_implicit || (g.s0 == 1 && g.s1 == 1) || g.d0 != 1 || g.d1 != 1) return false;
    if ((g.Cout % 64) != 0 || g.kh < g.s0 || g.kw < g.s1) return false;
    Geo& q = t->g;
    q = g;
    q.Cin = g.Cout; q.Cout = g.Cin; q.H = g.Ho; q.W = g.Wo;  // the "input" is dO, the "output" is dX
    q.Ho = g.s0 * g.Ho - (g.s0 - 1) + (g.kh - 1) - 2 * g.pt;  // convtranspose2d.py:165-181 with output_padding 0
    q.Wo = g.s1 * g.Wo - (g.s1 - 1) + (g.kw - 1) - 2 * g.pl;
    if (q.Ho != g.H || q.Wo != g.W) return false;             // asymmetric padding / ragged sizes: not a plain transposed conv
    if ((q.Ho % g.s0) != 0 || (q.Wo % g.s1) != 0) return false;
    t->P = q.Ho / g.s0; t->Q = q.Wo / g.s1;
    return gemm_conv_supported(1, q.Cin, t->P, t->Q, 1, 1);
}

}  // namespace
}  // namespace nnb

using namespace nnb;

extern "C" {

int nnb_conv2d_out_shape(const nnb_conv2d_desc* d, int64_t* Ho, int64_t* Wo) {
    Geo g{};
    int rc = make_geo(d, &g);
    if (rc) return rc;
    if (Ho) *Ho = g.Ho;
    if (Wo) *Wo = g.Wo;
    return NNB_OK;
}

size_t nnb_conv2d_workspace_bytes(const nnb_conv2d_desc* d, int prec, int backward) {
    Geo g{};
    if (make_geo(d, &g)) return 0;
    const size_t dbp = (size_t)round_up((int64_t)CHANNEL_SUM_CHUNKS * g.Cout * 4, 256) + 256;  // db partial sums
    if (use_direct(g))
        return 512 + dbp + (size_t)std::max(DIRECT_WGRAD_CHUNKS, TILE_MAX_BLOCKS * tile_wgrad_lanes(g)) * g.Cout * g.Cin * g.kh * g.kw * 4;
    const size_t p = planes(prec);
    const int cip = use_nhwc(g) ? cpad(g.Cin) : g.Cin, cop = use_nhwc(g) ? cpad(g.Cout) : g.Cout;
    const int64_t M = (int64_t)g.B * g.Ho * g.Wo, Kc = (int64_t)cip * g.kh * g.kw;
    const int64_t Mx = (int64_t)g.B * g.H * g.W, Kg = (int64_t)cop * g.kh * g.kw;
    size_t b = 8192 + dbp;
    if (use_nhwc(g)) {  // channels-last planes of X and dO, fp32 [Cout][(k,l,c)] wgrad result
        b += p * (size_t)round_up((int64_t)g.B * g.H * g.W * cip * 2, 256);
        if (backward) {
            b += p * (size_t)round_up(M * cop * 2, 256);
            b += (size_t)round_up((int64_t)g.Cout * Kc * 4, 256);
        }
    }
    if (!implicit_fwd(g)) b += p * staged_plane_bytes(1, M, Kc);  // col(X)
    if (!backward) {
        b += p * staged_plane_bytes(1, g.Cout, Kc);  // W
        b += gemm_splitk_ws_bytes(g.Cout, M, Kc, 1);
    } else {
        if (!use_nhwc(g)) b += p * staged_plane_bytes(1, g.Cout, M);   // g'
        if (!implicit_dgrad(g)) b += p * staged_plane_bytes(1, Mx, Kg);      // colT(g)
        b += p * staged_plane_bytes(1, g.Cin, Kg);   // Wr
        b += std::max(gemm_splitk_ws_bytes(g.Cout, Kc, M, 1), gemm_splitk_ws_bytes(g.Cin, Mx, Kg, 1));
    }
    return b;
}

size_t nnb_conv2d_planes_bytes(const nnb_conv2d_desc* d, int prec) {
    Geo g{};
    if (make_geo(d, &g) || !use_nhwc(g)) return 0;
    return planes(prec) * nhwc_bytes((int64_t)g.B * g.H * g.W, g.Cin);
}

int nnb_conv2d_forward(const nnb_conv2d_desc* d, const float* X, const float* Wt, const float* bias,
                       float* O, int prec, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream) {
    return nnb_conv2d_forward_ex(d, X, Wt, bias, O, prec, nullptr, nullptr, workspace, workspace_bytes, stream);
}

int nnb_conv2d_forward_ex(const nnb_conv2d_desc* d, const float* X, const float* Wt, const float* bias,
                          float* O, int prec, const void* X_planes, void* X_planes_out, void* workspace,
                          size_t workspace_bytes, cudaStream_t stream) {
    NNB_RANGE("nnb_conv2d_forward_ex");
    NNB_REQUIRE(X && Wt && O, "nnb_conv2d_forward: null pointer");
    NNB_REQUIRE(prec == NNB_PREC_BF16 || prec == NNB_PREC_BF16X3, "nnb_conv2d_forward: bad prec");
    Geo g{};
    int rc = make_geo(d, &g);
    if (rc) return rc;
    if (use_direct(g)) {
        TileGeo tg{};
        if (tile_geo(g, &tg)) {
            static bool cfg = false;
            if (!cfg) { int rc2 = tile_smem_attr(tile_fwd_kernel, 0); if (rc2) return rc2; cfg = true; }
            const size_t smem = ((size_t)g.Cout * g.Cin * g.kh * g.kw + (size_t)g.Cin * tg.Hp * tg.Wp) * 4;
            tile_fwd_kernel<<<tile_blocks(g), 256, smem, stream>>>(g, tg, X, Wt, bias, O);
        } else {
            const long long total = (long long)g.B * g.Cout * g.Ho * g.Wo;
            direct_fwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>(g, X, Wt, bias, O);
        }
        count_launch();
        NNB_CUDA_OK(cudaGetLastError());
        return NNB_OK;
    }
    const bool x3 = prec == NNB_PREC_BF16X3;
    const int cip = use_nhwc(g) ? cpad(g.Cin) : g.Cin;  // channel pitch of the planes / of col's (k, l, c) columns
    const int64_t HWo = (int64_t)g.Ho * g.Wo, M = g.B * HWo, Kc = (int64_t)cip * g.kh * g.kw;
    Bump ws(workspace, workspace_bytes);
    const bool implicit = implicit_fwd(g);
    Planes col{nullptr, nullptr};
    if (!implicit) col = take_planes(ws, M, Kc, prec);
    Planes wp = take_planes(ws, g.Cout, Kc, prec);
    if (!ws.ok()) return fail(NNB_ERR_WORKSPACE, "nnb_conv2d_forward: workspace too small (need >= %zu)", ws.off);
    Staged wst;
    if (use_nhwc(g)) {
        // channels-last bf16 planes of X: taken from the caller (a producer emitted them), written into the caller's
        // buffer (kept for the backward pass), or scratch
        const int64_t xpos = (int64_t)g.B * g.H * g.W;
        Planes xh;
        if (X_planes != nullptr) {
            xh = planes_at(const_cast<void*>(X_planes), xpos, g.Cin, prec);
        } else {
            if (X_planes_out != nullptr) {
                NNB_REQUIRE((reinterpret_cast<uintptr_t>(X_planes_out) & 255) == 0, "nnb_conv2d_forward: X_planes_out must be 256-byte aligned");
                xh = planes_at(X_planes_out, xpos, g.Cin, prec);
            } else {
                xh = take_nhwc(ws, xpos, g.Cin, prec);
                if (!ws.ok()) return fail(NNB_ERR_WORKSPACE, "nnb_conv2d_forward: workspace too small (need >= %zu)", ws.off);
            }
            rc = run_to_nhwc(X, g.B, g.Cin, g.H * g.W, xh, x3, stream);
            if (rc) return rc;
        }
        if (implicit) {
            rc = run_stage_weight_klc(Wt, g, false, wp, staged_ld(Kc), x3, stream);
            if (rc) return rc;
            GemmProblem p;
            p.M = g.Cout; p.N = M; p.K = Kc;
            p.A.st = as_staged(wp, g.Cout, Kc);
            p.conv = conv_operand(1, xh, g.B, g.Cin, g.H, g.W, g.Ho, g.Wo, g.kh, g.kw, g.s0, g.s1, g.pt, g.pl, g.d0, g.d1);
            p.D = O; p.ldd = HWo;
            p.col_group = HWo; p.group_stride = (int64_t)g.Cout * HWo;
            p.epi.bias = bias; p.bias_per_row = true;
            p.splitk_ws_bytes = ws.remaining();
            p.splitk_ws = static_cast<float*>(ws.take(p.splitk_ws_bytes));
            return gemm(p, stream);
        }
        GatherNhwcArgs a{xh.hi, xh.lo, cip, g.H, g.W, g.Ho, g.Wo, g.kh, g.kw, g.s0, g.s1, g.pt, g.pl, g.d0, g.d1, 1, 1,
                         M, staged_ld(Kc), col.hi, col.lo};
        rc = run_gather_nhwc(a, x3, stream);
        if (rc) return rc;
        rc = run_stage_weight_klc(Wt, g, false, wp, staged_ld(Kc), x3, stream);
        if (rc) return rc;
        wst = as_staged(wp, g.Cout, Kc);
    } else {
        GatherArgs a{X, g.Cin, g.H, g.W, g.Ho, g.Wo, g.kh, g.kw, g.s0, g.s1, g.pt, g.pl, g.d0, g.d1, 1, 1,
                     M, (int)Kc, staged_ld(Kc), col.hi, col.lo};
        rc = run_gather(a, x3, stream);
        if (rc) return rc;
        rc = stage_operand(view2d(Wt, g.Cout, Kc, Kc), false, prec, wp.hi, wp.lo, STAGE_COPY, nullptr, 0.f,
                           nullptr, nullptr, stream, &wst);
        if (rc) return rc;
    }
    GemmProblem p;
    p.M = g.Cout; p.N = M; p.K = Kc;
    p.A.st = wst;
    p.B.st = as_staged(col, M, Kc);
    p.D = O; p.ldd = HWo;
    p.col_group = HWo; p.group_stride = (int64_t)g.Cout * HWo;
    p.epi.bias = bias; p.bias_per_row = true;
    p.splitk_ws_bytes = ws.remaining();
    p.splitk_ws = static_cast<float*>(ws.take(p.splitk_ws_bytes));
    return gemm(p, stream);
}

int nnb_conv2d_backward(const nnb_conv2d_desc* d, const float* X, const float* Wt, const float* dO,
                        float* dX, float* dW, float* db, int prec, void* workspace,
                        size_t workspace_bytes, cudaStream_t stream) {
    return nnb_conv2d_backward_ex(d, X, Wt, dO, dX, dW, db, prec, nullptr, workspace, workspace_bytes, stream);
}

int nnb_conv2d_backward_ex(const nnb_conv2d_desc* d, const float* X, const float* Wt, const float* dO,
                           float* dX, float* dW, float* db, int prec, const void* X_planes, void* workspace,
                           size_t workspace_bytes, cudaStream_t stream) {
    NNB_RANGE("nnb_conv2d_backward_ex");
    NNB_REQUIRE(X && Wt && dO && dW, "nnb_conv2d_backward: null pointer");
    NNB_REQUIRE(prec == NNB_PREC_BF16 || prec == NNB_PREC_BF16X3, "nnb_conv2d_backward: bad prec");
    Geo g{};
    int rc = make_geo(d, &g);
    if (rc) return rc;
    const int64_t HWo = (int64_t)g.Ho * g.Wo;
    Bump ws(workspace, workspace_bytes);
    if (db) {
        // >= ~4096 elements per block: tiny feature maps otherwise drown in block-scheduling overhead
        const int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(CHANNEL_SUM_CHUNKS, g.B), ((int64_t)g.B * HWo + 4095) / 4096));
        float* part = static_cast<float*>(ws.take((size_t)CHANNEL_SUM_CHUNKS * g.Cout * 4));
        if (!ws.ok()) return fail(NNB_ERR_WORKSPACE, "nnb_conv2d_backward: workspace too small (need >= %zu)", ws.off);
        channel_sum_kernel<<<dim3((unsigned)g.Cout, (unsigned)chunks), 256, 0, stream>>>(dO, g.B, g.Cout, (int)HWo, part);
        direct_wgrad_finish_kernel<<<(unsigned)ceil_div(g.Cout, 256), 256, 0, stream>>>(part, g.Cout, chunks, db);
        count_launch(2);
        NNB_CUDA_OK(cudaGetLastError());
    }
    if (use_direct(g)) {
        const int nw = g.Cout * g.Cin * g.kh * g.kw;
        TileGeo tg{};
        const bool tiled = tile_geo(g, &tg);
        Bump& dws = ws;
        if (tiled) {
            static bool cfg = false;
            if (!cfg) {
                int rc2 = tile_smem_attr(tile_wgrad_kernel, 0);
                if (!rc2) rc2 = tile_smem_attr(tile_dgrad_kernel, 0);
                if (rc2) return rc2;
                cfg = true;
            }
            const int blocks = tile_blocks(g), rows = blocks * tile_wgrad_lanes(g);
            float* partial = static_cast<float*>(dws.take((size_t)rows * nw * 4));
            if (!dws.ok()) return fail(NNB_ERR_WORKSPACE, "nnb_conv2d_backward: workspace too small (need >= %zu)", dws.off);
            const size_t smem = ((size_t)g.Cout * HWo + (size_t)g.Cin * tg.Hp * tg.Wp) * 4;
            tile_wgrad_kernel<<<blocks, 256, smem, stream>>>(g, tg, X, dO, partial);
            direct_wgrad_finish_kernel<<<(unsigned)ceil_div(nw, 256), 256, 0, stream>>>(partial, nw, rows, dW);
            count_launch(2);
            if (dX) {
                const size_t smem2 = ((size_t)nw + (size_t)g.Cout * HWo) * 4;
                tile_dgrad_kernel<<<blocks, 256, smem2, stream>>>(g, dO, Wt, dX);
                count_launch();
            }
            NNB_CUDA_OK(cudaGetLastError());
            return NNB_OK;
        }
        const int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(DIRECT_WGRAD_CHUNKS, (g.B * HWo + 4095) / 4096));
        float* partial = static_cast<float*>(dws.take((size_t)chunks * nw * 4));
        if (!dws.ok()) return fail(NNB_ERR_WORKSPACE, "nnb_conv2d_backward: workspace too small (need >= %zu)", dws.off);
        direct_wgrad_kernel<<<dim3((unsigned)nw, (unsigned)chunks), 256, 0, stream>>>(g, X, dO, partial);
        direct_wgrad_finish_kernel<<<(unsigned)ceil_div(nw, 256), 256, 0, stream>>>(partial, nw, chunks, dW);
        count_launch(2);
        if (dX) {
            const long long total = (long long)g.B * g.Cin * g.H * g.W;
            direct_dgrad_kernel<<<grid_for(total, 256), 256, 0, stream>>>(g, dO, Wt, dX);
            count_launch();
        }
        NNB_CUDA_OK(cudaGetLastError());
        return NNB_OK;
    }
    const bool x3 = prec == NNB_PREC_BF16X3;
    const int cip = use_nhwc(g) ? cpad(g.Cin) : g.Cin, cop = use_nhwc(g) ? cpad(g.Cout) : g.Cout;
    const int64_t M = g.B * HWo, Kc = (int64_t)cip * g.kh * g.kw;
    const int64_t Mx = (int64_t)g.B * g.H * g.W, Kg = (int64_t)cop * g.kh * g.kw;
    const bool imp_w = implicit_fwd(g), imp_d = dX != nullptr && implicit_dgrad(g);
    Planes col{nullptr, nullptr};
    if (!imp_w) col = take_planes(ws, M, Kc, prec);
    Planes gp{nullptr, nullptr};
    if (!use_nhwc(g)) gp = take_planes(ws, g.Cout, M, prec);
    Planes colg{nullptr, nullptr}, wr{nullptr, nullptr};
    TGeo tg{};
    // stride > 1: dX as a transposed convolution of dO in parity classes (real taps only) instead of the zero-stuffed gather
    const bool cls_d = dX != nullptr && !imp_d && use_nhwc(g) && strided_dgrad_as_tconv(g, &tg);
    if (dX && !cls_d) {
        if (!imp_d) colg = take_planes(ws, Mx, Kg, prec);
        wr = take_planes(ws, g.Cin, Kg, prec);
    }
    if (!ws.ok()) return fail(NNB_ERR_WORKSPACE, "nnb_conv2d_backward: workspace too small (need >= %zu)", ws.off);
    if (use_nhwc(g)) {
        // channels-last path: X and dO become bf16 [positions][C] planes once (X's are reused from forward when the
        // caller kept them); dO's planes are wgrad's A operand as they are (MN-major: rows = reduction) and the
        // source of dgrad
        Planes xh;
        if (X_planes != nullptr) xh = planes_at(const_cast<void*>(X_planes), (int64_t)g.B * g.H * g.W, g.Cin, prec);
        else xh = take_nhwc(ws, (int64_t)g.B * g.H * g.W, g.Cin, prec);
        Planes gh = take_nhwc(ws, M, g.Cout, prec);
        float* dwt = static_cast<float*>(ws.take((size_t)round_up((int64_t)g.Cout * Kc * 4, 256)));
        if (!ws.ok()) return fail(NNB_ERR_WORKSPACE, "nnb_conv2d_backward: workspace too small (need >= %zu)", ws.off);
        const size_t skb = ws.remaining();
        float* skp = static_cast<float*>(ws.take(skb));
        if (X_planes == nullptr) {
            rc = run_to_nhwc(X, g.B, g.Cin, g.H * g.W, xh, x3, stream);
            if (rc) return rc;
        }
        rc = run_to_nhwc(dO, g.B, g.Cout, (int)HWo, gh, x3, stream);
        if (rc) return rc;
        if (!imp_w) {
            GatherNhwcArgs a{xh.hi, xh.lo, cip, g.H, g.W, g.Ho, g.Wo, g.kh, g.kw, g.s0, g.s1, g.pt, g.pl, g.d0, g.d1, 1, 1,
                             M, staged_ld(Kc), col.hi, col.lo};
            rc = run_gather_nhwc(a, x3, stream);
            if (rc) return rc;
        }
        {   // wgrad: T[o][(k,l,c)] = sum_m gh[m][o] * col[m][(k,l,c)], then T -> dW[o][c][k][l]
            GemmProblem p;
            p.M = g.Cout; p.N = Kc; p.K = M;
            p.A.st = as_staged(gh, M, g.Cout); p.A.mn_major = true;
            if (imp_w) {
                p.B.mn_major = true;
                p.conv = conv_operand(2, xh, g.B, g.Cin, g.H, g.W, g.Ho, g.Wo, g.kh, g.kw, g.s0, g.s1, g.pt, g.pl, g.d0, g.d1);
            } else {
                p.B.st = as_staged(col, M, Kc); p.B.mn_major = true;
            }
            p.D = dwt; p.ldd = Kc;
            p.splitk_ws = skp; p.splitk_ws_bytes = skb;
            rc = gemm(p, stream);
            if (rc) return rc;
            const long long total = (long long)g.Cout * g.Cin;
            permute_dw_kernel<<<grid_for(total, 256), 256, 0, stream>>>(dwt, g.Cout, g.Cin, g.kh * g.kw, dW);
            count_launch();
            NNB_CUDA_OK(cudaGetLastError());
        }
        if (cls_d) {
            rc = tconv_classes(tg, gh, Wt, 1, nullptr, dX, prec, Bump(skp, skb), stream);
            if (rc) return rc;
        } else if (dX) {
            if (!imp_d) {
                GatherNhwcArgs ga{gh.hi, gh.lo, cop, g.Ho, g.Wo, g.H, g.W, g.kh, g.kw, 1, 1,
                                  g.d0 * (g.kh - 1) - g.pt, g.d1 * (g.kw - 1) - g.pl, g.d0, g.d1, g.s0, g.s1,
                                  Mx, staged_ld(Kg), colg.hi, colg.lo};
                rc = run_gather_nhwc(ga, x3, stream);
                if (rc) return rc;
            }
            rc = run_stage_weight_klc(Wt, g, true, wr, staged_ld(Kg), x3, stream);
            if (rc) return rc;
            const int64_t HW = (int64_t)g.H * g.W;
            GemmProblem p;
            p.M = g.Cin; p.N = Mx; p.K = Kg;
            p.A.st = as_staged(wr, g.Cin, Kg);
            if (imp_d)  // positions = pixels of dX, source = dO planes, flipped taps (the weights are staged flipped)
                p.conv = conv_operand(1, gh, g.B, g.Cout, g.Ho, g.Wo, g.H, g.W, g.kh, g.kw, 1, 1,
                                      g.d0 * (g.kh - 1) - g.pt, g.d1 * (g.kw - 1) - g.pl, g.d0, g.d1);
            else
                p.B.st = as_staged(colg, Mx, Kg);
            p.D = dX; p.ldd = HW;
            p.col_group = HW; p.group_stride = (int64_t)g.Cin * HW;
            p.splitk_ws = skp; p.splitk_ws_bytes = skb;
            rc = gemm(p, stream);
            if (rc) return rc;
        }
        return NNB_OK;
    }
    const size_t sk_bytes = ws.remaining();
    float* sk = static_cast<float*>(ws.take(sk_bytes));

    // ---- wgrad: dW[o, (i,k,l)] = sum_{(b,ho,wo)} g'[o, (b,ho,wo)] * col[(b,ho,wo), (i,k,l)]   (conv2d.py:93)
    GatherArgs a{X, g.Cin, g.H, g.W, g.Ho, g.Wo, g.kh, g.kw, g.s0, g.s1, g.pt, g.pl, g.d0, g.d1, 1, 1,
                 M, (int)Kc, staged_ld(Kc), col.hi, col.lo};
    rc = run_gather(a, x3, stream);
    if (rc) return rc;
    {
        const long long total = (long long)g.B * g.Cout * HWo;
        if (x3) stage_nchw_to_c_bhw_kernel<true><<<grid_for(total, 256), 256, 0, stream>>>(dO, g.B, g.Cout, (int)HWo, staged_ld(M), gp.hi, gp.lo);
        else stage_nchw_to_c_bhw_kernel<false><<<grid_for(total, 256), 256, 0, stream>>>(dO, g.B, g.Cout, (int)HWo, staged_ld(M), gp.hi, gp.lo);
        count_launch();
        NNB_CUDA_OK(cudaGetLastError());
        GemmProblem p;
        p.M = g.Cout; p.N = Kc; p.K = M;
        p.A.st = as_staged(gp, g.Cout, M);               // [Cout][M]: K-major
        p.B.st = as_staged(col, M, Kc); p.B.mn_major = true;  // [M(reduction)][Kc]
        p.D = dW; p.ldd = Kc;
        p.splitk_ws = sk; p.splitk_ws_bytes = sk_bytes;
        rc = gemm(p, stream);
        if (rc) return rc;
    }
    // ---- dgrad: dX[i, (b,y,x)] = sum_{(o,k',l')} Wr[i,(o,k',l')] * colT[(b,y,x),(o,k',l')]   (conv2d.py:35-106)
    if (dX) {
        GatherArgs ga{dO, g.Cout, g.Ho, g.Wo, g.H, g.W, g.kh, g.kw, 1, 1,
                      g.d0 * (g.kh - 1) - g.pt, g.d1 * (g.kw - 1) - g.pl, g.d0, g.d1, g.s0, g.s1,
                      Mx, (int)Kg, staged_ld(Kg), colg.hi, colg.lo};
        rc = run_gather(ga, x3, stream);
        if (rc) return rc;
        const long long total = (long long)g.Cin * Kg;
        if (x3) stage_weight_dgrad_kernel<true><<<grid_for(total, 256), 256, 0, stream>>>(Wt, g.Cout, g.Cin, g.kh, g.kw, staged_ld(Kg), wr.hi, wr.lo);
        else stage_weight_dgrad_kernel<false><<<grid_for(total, 256), 256, 0, stream>>>(Wt, g.Cout, g.Cin, g.kh, g.kw, staged_ld(Kg), wr.hi, wr.lo);
        count_launch();
        NNB_CUDA_OK(cudaGetLastError());
        const int64_t HW = (int64_t)g.H * g.W;
        GemmProblem p;
        p.M = g.Cin; p.N = Mx; p.K = Kg;
        p.A.st = as_staged(wr, g.Cin, Kg);
        p.B.st = as_staged(colg, Mx, Kg);
        p.D = dX; p.ldd = HW;
        p.col_group = HW; p.group_stride = (int64_t)g.Cin * HW;
        p.splitk_ws = sk; p.splitk_ws_bytes = sk_bytes;
        rc = gemm(p, stream);
        if (rc) return rc;
    }
    return NNB_OK;
}

int nnb_conv_transpose2d_supported(const nnb_conv2d_desc* d, int out_pad0, int out_pad1) {
    TGeo t{};
    if (make_tgeo(d, out_pad0, out_pad1, &t)) return 0;
    return tconv_native(t, out_pad0, out_pad1) ? 1 : 0;
}

int nnb_conv_transpose2d_out_shape(const nnb_conv2d_desc* d, int out_pad0, int out_pad1, int64_t* Ho, int64_t* Wo) {
    TGeo t{};
    int rc = make_tgeo(d, out_pad0, out_pad1, &t);
    if (rc) return rc;
    if (Ho) *Ho = t.g.Ho;
    if (Wo) *Wo = t.g.Wo;
    return NNB_OK;
}

size_t nnb_conv_transpose2d_planes_bytes(const nnb_conv2d_desc* d, int prec) {
    if (!d || d->B <= 0 || d->Cin <= 0 || d->H <= 0 || d->W <= 0) return 0;
    return planes(prec) * nhwc_bytes(d->B * d->H * d->W, d->Cin);
}

size_t nnb_conv_transpose2d_workspace_bytes(const nnb_conv2d_desc* d, int out_pad0, int out_pad1, int prec, int backward) {
    TGeo t{};
    if (make_tgeo(d, out_pad0, out_pad1, &t)) return 0;
    const Geo& g = t.g;
    const size_t p = planes(prec);
    const int64_t xpos = (int64_t)g.B * g.H * g.W, opos = (int64_t)g.B * g.Ho * g.Wo;
    const int64_t Kfull = (int64_t)g.kh * g.kw;
    size_t b = 16384 + (size_t)round_up((int64_t)CHANNEL_SUM_CHUNKS * g.Cout * 4, 256);
    b += p * nhwc_bytes(xpos, g.Cin);
    if (!backward) {
        b += (size_t)round_up(opos * g.Cout * 4, 256);                          // class outputs
        b += p * staged_plane_bytes(1, g.Cout, Kfull * g.Cin);                  // class weights (all classes together <= full kernel)
        b += 4 * 256 * p;
        b += gemm_splitk_ws_bytes(g.Cout, xpos, Kfull * g.Cin, 1);
    } else {
        b += p * nhwc_bytes(opos, g.Cout);                                       // dO planes
        b += p * staged_plane_bytes(1, g.Cin, Kfull * g.Cout);                   // flipped weights
        b += (size_t)round_up((int64_t)g.Cin * Kfull * g.Cout * 4, 256);         // T
        b += std::max(gemm_splitk_ws_bytes(g.Cin, xpos, Kfull * g.Cout, 1), gemm_splitk_ws_bytes(g.Cin, Kfull * g.Cout, xpos, 1));
    }
    return b;
}

int nnb_conv_transpose2d_forward(const nnb_conv2d_desc* d, int out_pad0, int out_pad1, const float* X, const float* Wt,
                                 const float* bias, float* O, int prec, void* X_planes_out, void* workspace,
                                 size_t workspace_bytes, cudaStream_t stream) {
    NNB_RANGE("nnb_conv_transpose2d_forward");
    NNB_REQUIRE(X && Wt && O, "nnb_conv_transpose2d_forward: null pointer");
    NNB_REQUIRE(prec == NNB_PREC_BF16 || prec == NNB_PREC_BF16X3, "nnb_conv_transpose2d_forward: bad prec");
    TGeo t{};
    int rc = make_tgeo(d, out_pad0, out_pad1, &t);
    if (rc) return rc;
    if (!tconv_native(t, out_pad0, out_pad1))
        return fail(NNB_ERR_UNSUPPORTED, "nnb_conv_transpose2d_forward: geometry outside the native gather form "
                                          "(needs Cin, Cout %% 64 == 0, dilation 1, no output padding, Ho %% stride == 0)");
    const Geo& g = t.g;
    const bool x3 = prec == NNB_PREC_BF16X3;
    const int64_t xpos = (int64_t)g.B * g.H * g.W;
    Bump ws(workspace, workspace_bytes);
    Planes xh;
    if (X_planes_out != nullptr) {
        NNB_REQUIRE((reinterpret_cast<uintptr_t>(X_planes_out) & 255) == 0, "nnb_conv_transpose2d_forward: X_planes_out must be 256-byte aligned");
        xh = planes_at(X_planes_out, xpos, g.Cin, prec);
    } else {
        xh = take_nhwc(ws, xpos, g.Cin, prec);
    }
    if (!ws.ok()) return fail(NNB_ERR_WORKSPACE, "nnb_conv_transpose2d_forward: workspace too small (need >= %zu)", ws.off);
    rc = run_to_nhwc(X, g.B, g.Cin, g.H * g.W, xh, x3, stream);
    if (rc) return rc;
    return tconv_classes(t, xh, Wt, 0, bias, O, prec, ws, stream);
}

int nnb_conv_transpose2d_backward(const nnb_conv2d_desc* d, int out_pad0, int out_pad1, const float* X, const float* Wt,
                                  const float* dO, float* dX, float* dW, float* db, int prec, const void* X_planes,
                                  void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    NNB_RANGE("nnb_conv_transpose2d_backward");
    NNB_REQUIRE(X && Wt && dO && dW, "nnb_conv_transpose2d_backward: null pointer");
    NNB_REQUIRE(prec == NNB_PREC_BF16 || prec == NNB_PREC_BF16X3, "nnb_conv_transpose2d_backward: bad prec");
    TGeo t{};
    int rc = make_tgeo(d, out_pad0, out_pad1, &t);
    if (rc) return rc;
    if (!tconv_native(t, out_pad0, out_pad1))
        return fail(NNB_ERR_UNSUPPORTED, "nnb_conv_transpose2d_backward: geometry outside the native gather form");
    const Geo& g = t.g;
    const bool x3 = prec == NNB_PREC_BF16X3;
    const int64_t xpos = (int64_t)g.B * g.H * g.W, opos = (int64_t)g.B * g.Ho * g.Wo, HWo = (int64_t)g.Ho * g.Wo;
    const int64_t Kg = (int64_t)g.kh * g.kw * g.Cout;
    Bump ws(workspace, workspace_bytes);
    if (db) {
        // >= ~4096 elements per block: tiny feature maps otherwise drown in block-scheduling overhead
        const int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(CHANNEL_SUM_CHUNKS, g.B), ((int64_t)g.B * HWo + 4095) / 4096));
        float* part = static_cast<float*>(ws.take((size_t)CHANNEL_SUM_CHUNKS * g.Cout * 4));
        if (!ws.ok()) return fail(NNB_ERR_WORKSPACE, "nnb_conv_transpose2d_backward: workspace too small (need >= %zu)", ws.off);
        channel_sum_kernel<<<dim3((unsigned)g.Cout, (unsigned)chunks), 256, 0, stream>>>(dO, g.B, g.Cout, (int)HWo, part);
        direct_wgrad_finish_kernel<<<(unsigned)ceil_div(g.Cout, 256), 256, 0, stream>>>(part, g.Cout, chunks, db);
        count_launch(2);
        NNB_CUDA_OK(cudaGetLastError());
    }
    Planes xh;
    if (X_planes != nullptr) xh = planes_at(const_cast<void*>(X_planes), xpos, g.Cin, prec);
    else xh = take_nhwc(ws, xpos, g.Cin, prec);
    Planes gh = take_nhwc(ws, opos, g.Cout, prec);
    Planes wf = take_planes(ws, g.Cin, Kg, prec);
    float* T = static_cast<float*>(ws.take((size_t)round_up((int64_t)g.Cin * Kg * 4, 256)));
    if (!ws.ok()) return fail(NNB_ERR_WORKSPACE, "nnb_conv_transpose2d_backward: workspace too small (need >= %zu)", ws.off);
    const size_t skb = ws.remaining();
    float* skp = static_cast<float*>(ws.take(skb));
    if (X_planes == nullptr) {
        rc = run_to_nhwc(X, g.B, g.Cin, g.H * g.W, xh, x3, stream);
        if (rc) return rc;
    }
    rc = run_to_nhwc(dO, g.B, g.Cout, (int)HWo, gh, x3, stream);
    if (rc) return rc;
    // dO read through the stride: pixel (u, v) of X with tap (k', l') meets dO[s*u - pad + k'][s*v - pad + l']
    const ConvOperand gsrc1 = conv_operand(1, gh, g.B, g.Cout, g.Ho, g.Wo, g.H, g.W, g.kh, g.kw, g.s0, g.s1, g.pt, g.pl, 1, 1);
    if (dX) {
        // dX[c][(b,u,v)] = sum_{(k',l',o)} Wf[c][(k',l',o)] * dO[...][o],  Wf[c][(k',l',o)] = W[o][c][kh-1-k'][kw-1-l']
        rc = run_stage_weight_klc(Wt, g, true, wf, staged_ld(Kg), x3, stream);
        if (rc) return rc;
        const int64_t HW = (int64_t)g.H * g.W;
        GemmProblem p;
        p.M = g.Cin; p.N = xpos; p.K = Kg;
        p.A.st = as_staged(wf, g.Cin, Kg);
        p.conv = gsrc1;
        p.D = dX; p.ldd = HW;
        p.col_group = HW; p.group_stride = (int64_t)g.Cin * HW;
        p.splitk_ws = skp; p.splitk_ws_bytes = skb;
        rc = gemm(p, stream);
        if (rc) return rc;
    }
    {
        // T[c][(k',l',o)] = sum_{(b,u,v)} X[(b,u,v)][c] * dO[(b, s*u - pad + k', s*v - pad + l')][o]
        GemmProblem p;
        p.M = g.Cin; p.N = Kg; p.K = xpos;
        p.A.st = as_staged(xh, xpos, g.Cin); p.A.mn_major = true;
        p.B.mn_major = true;
        p.conv = gsrc1;
        p.conv.mode = 2;
        p.D = T; p.ldd = Kg;
        p.splitk_ws = skp; p.splitk_ws_bytes = skb;
        rc = gemm(p, stream);
        if (rc) return rc;
        const long long total = (long long)g.Cout * g.Cin;
        permute_dw_convT_kernel<<<grid_for(total, 256), 256, 0, stream>>>(T, g.Cout, g.Cin, g.kh, g.kw, dW);
        count_launch();
        NNB_CUDA_OK(cudaGetLastError());
    }
    return NNB_OK;
}

}  // extern "C"
