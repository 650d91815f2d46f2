// tcgen05 GEMM for sm_100a: D = epilogue(alpha * sum_seg A_seg . B_seg^T)
//
//   * operands are staged bf16 planes (common.cuh: Staged); fp32 accumulation in TMEM
//   * TMA (cp.async.bulk.tensor, 128B swizzle) feeds a multi-stage smem ring,
//     one thread issues tcgen05.mma (UMMA 128 x BN x 16), accumulators are double-buffered in TMEM
//     so the epilogue of tile i overlaps the MMAs of tile i+1 (persistent CTAs, one per SM)
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
//     warps 4..7 = epilogue (tcgen05.ld -> smem transpose -> coalesced global stores)
//   * both operand majors are supported through the UMMA descriptors, so the three Linear GEMMs
//     (fwd X.W^T, dgrad g.W, wgrad g^T.X; neunet/nn/layers/linear.py:19-22,54) and the matmul
//     backward forms (neunet/autograd.py:209-211) read the SAME staged buffers, no transposes
//   * BF16X3 precision = three (A,B) segment pairs accumulated into one TMEM tile
//   * split-K / reduce-over-batch for wgrad-shaped problems (tiny output, huge reduction)
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace nnb {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle-128B atom row
constexpr int UMMA_K = 16;
constexpr int MAX_SEG = 3;
constexpr int NUM_THREADS = 256;
constexpr int EPI_SCRATCH_BYTES = 4 * 8192;  // per epilogue warp: two 4 KB TMA-store tiles (or the 32x33 transpose scratch)
constexpr int SMEM_BUDGET = 200 * 1024;  // operand ring budget

struct GemmMaps {
    CUtensorMap a[MAX_SEG];
    CUtensorMap b[MAX_SEG];
    CUtensorMap d;  // fp32 output, box 32 x 32, 128B swizzle (TMA-store epilogue)
    CUtensorMap z;  // fp32 pre-activation side output
};

struct KArgs {
    int M, N;
    int kblocks;       // ceil(K / BK)
    int out_batches;   // independent output matrices
    int red_batches;   // batches summed into one output (reduce_batch mode), else 1
    int a_bcast, b_bcast;
    int nseg;
    int splits, iters_per_split;  // split over the (red_batches * kblocks) iteration space
    int num_m, num_n;
    float* D;
    long long ldd, batch_stride_d;
    int col_group;           // output column index map: addr = (col / cg) * gs + row * ldd + col % cg
    long long group_stride;
    float alpha;
    const float* bias;
    int bias_per_row;
    float* Z;
    int act;
    float beta;
    __nv_bfloat16* D16;
    long long ldd16;
    float* ws;  // split-K partials [split][out_batch][M][N]
    int tma_store;  // 1: epilogue writes D/Z through TMA (needs 16-byte aligned rows)
    int bias_vec;   // 1: bias pointer 16-byte aligned (float4 loads)
    unsigned long long* clk_out;  // optional: CTA 0 writes {SM cycles, ns} of its lifetime (clock probe)
    // implicit-GEMM convolution (common.cuh: ConvOperand); conv_mode 0 = plain B operand
    int conv_mode;
    int cv_P, cv_Q;             // position grid per image
    int cv_bw, cv_bh;           // box extent in (q, p); the image extent follows from the rows per box
    int cv_sp0, cv_sp1, cv_off0, cv_off1, cv_dk0, cv_dk1, cv_kw;
    int cv_cblocks;             // C / 64
    int stages;     // operand ring depth actually used (<= Cfg::STAGES; experiments only)
    int debug;      // profiling experiments only (NNB_GEMM_DEBUG): 1 = epilogue skips the stores, 2 = also skips the TMEM loads
};

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
    // UMMA shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version field = 1.
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}

__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn, int m) {
    // kind::f16 instruction descriptor: D=f32, A=B=bf16, M = 128 (cta_group::1) or 256 (cta_group::2).
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) |
           ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// Grouped tile order: walk GROUP_M row-blocks under each column-block before moving on, so the
// CTAs of one wave share a [GROUP_M*128 x K] slab of A and a few B slabs that stay L2-resident.
constexpr int GROUP_M = 16;
__device__ __forceinline__ void tile_coords(int r, int num_m, int num_n, int& m_blk, int& n_blk) {
    const int per_group = GROUP_M * num_n;
    const int g = r / per_group;
    const int first_m = g * GROUP_M;
    const int gm = min(num_m - first_m, GROUP_M);
    const int in_g = r - g * per_group;
    n_blk = in_g / gm;
    m_blk = first_m + (in_g - n_blk * gm);
}

// 4-D box origin (c, x, y, b) of the rows starting at position m0 for tap `tap`, channel block cb.
__device__ __forceinline__ void conv_coords(const KArgs& p, int m0, int tap, int cb, int& c0, int& x0, int& y0, int& b0) {
    const int q0 = m0 % p.cv_Q;
    const int t = m0 / p.cv_Q;
    const int p0 = t % p.cv_P;
    b0 = t / p.cv_P;
    const int kk = tap / p.cv_kw, ll = tap - kk * p.cv_kw;
    c0 = cb * 64;
    x0 = q0 * p.cv_sp1 - p.cv_off1 + ll * p.cv_dk1;
    y0 = p0 * p.cv_sp0 - p.cv_off0 + kk * p.cv_dk0;
}

__device__ __forceinline__ float apply_act(float x, int act, float beta) {
    if (act == NNB_ACT_SWISH) return x / (1.0f + __expf(-beta * x));
    return x;
}

// CG = CTAs per tile (tcgen05 cta_group): with CG == 2 a CTA pair computes a 256 x BN tile, each CTA
// holding its own 128 rows of A and HALF of the B tile; the UMMA reads both halves, so operand fill
// and shared-memory reads per SM drop from (16 + BN/8) KB to (16 + BN/16) KB per k-block.
template <int BN, int CG>
struct Cfg {
    static constexpr int BNL = BN / CG;  // B rows (or columns, MN-major) staged by this CTA
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BNL * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (SMEM_BUDGET / STAGE_BYTES);
    static constexpr int TMEM_COLS = (2 * BN) < 32 ? 32 : 2 * BN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_SCRATCH_BYTES + 256 + 1024;
};

template <int BN, bool A_MN, bool B_MN, int CG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ GemmMaps maps, const KArgs p) {
    using C = Cfg<BN, CG>;
    constexpr int BNL = C::BNL;
    const uint32_t cta_rank = (CG == 2) ? ptx::cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    constexpr int STAGES = C::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * C::A_BYTES;
    float* epi_scratch = reinterpret_cast<float*>(smem + STAGES * C::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES + EPI_SCRATCH_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    unsigned long long clk0 = 0, ns0 = 0;
    if (p.clk_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
        clk0 = clock64();
        ns0 = ptx::globaltimer_ns();
    }

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.nseg; ++s) {
            ptx::tma_prefetch_desc(&maps.a[s]);
            ptx::tma_prefetch_desc(&maps.b[s]);
        }
        if (p.tma_store) {
            ptx::tma_prefetch_desc(&maps.d);
            if (p.Z != nullptr) ptx::tma_prefetch_desc(&maps.z);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            ptx::mbar_init(&full_bar[i], CG);  // one arrival per CTA of the pair (leader's copy is used)
            ptx::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tmem_full[i], 1);
            ptx::mbar_init(&tmem_empty[i], 128 * CG);  // epilogue threads of every CTA of the pair
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        if (CG == 2) ptx::tmem_alloc_2sm<C::TMEM_COLS>(tmem_slot); else ptx::tmem_alloc<C::TMEM_COLS>(tmem_slot);
    }
    ptx::tc_fence_before();
    if (CG == 2) ptx::cluster_sync(); else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // PDL: everything above (tensor-map prefetch, barrier init, TMEM allocation) overlapped the previous
    // kernel's tail; operands, bias and outputs may only be touched once it has completed
    pdl_trigger();
    pdl_wait();

    const int tiles_per_batch = p.num_m * p.num_n;
    const int tiles = tiles_per_batch * p.out_batches;
    const int work = tiles * p.splits;
    const int total_iters = p.red_batches * p.kblocks;
    const int worker = blockIdx.x / CG, num_workers = gridDim.x / CG;  // a worker = one CTA or one CTA pair
    const int nstages = (p.stages > 0 && p.stages < STAGES) ? p.stages : STAGES;

    if (warp == 0) {
        // ===================== TMA producer =====================
        // elect.sync (not `lane == 0`): ptxas then knows the region is single-threaded and keeps the
        // TMA / barrier operands in uniform registers instead of wrapping each issue in a waterfall loop
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = worker; w < work; w += num_workers) {
                const int split = w / tiles;
                const int t = w - split * tiles;
                const int ob = t / tiles_per_batch;
                const int r = t - ob * tiles_per_batch;
                int m_blk, n_blk;
                tile_coords(r, p.num_m, p.num_n, m_blk, n_blk);
                const int m0 = m_blk * (BM * CG) + (int)cta_rank * BM;
                const int n0 = n_blk * BN + (int)cta_rank * BNL;  // this CTA's share of the B tile
                const int it0 = split * p.iters_per_split;
                const int it1 = min(it0 + p.iters_per_split, total_iters);
                for (int it = it0; it < it1; ++it) {
                    const int rb = it / p.kblocks;
                    const int k0 = (it - rb * p.kblocks) * BK;
                    const int bidx = (p.red_batches > 1) ? rb : ob;
                    const int ba = p.a_bcast ? 0 : bidx;
                    const int bb = p.b_bcast ? 0 : bidx;
                    for (int s = 0; s < p.nseg; ++s) {
                        ptx::mbar_wait(&empty_bar[stage], phase ^ 1, 1);
                        // all transaction bytes of the pair are counted on the LEADER's barrier
                        if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES * CG);
                        else ptx::mbar_arrive_cluster(&full_bar[stage], 0);
                        uint8_t* sa = smem_a + stage * C::A_BYTES;
                        uint8_t* sb = smem_b + stage * C::B_BYTES;
                        uint64_t* fb = &full_bar[stage];
                        if (!A_MN) {
                            ptx::tma_load_3d_cg<CG>(sa, &maps.a[s], fb, k0, m0, ba);
                        } else {
#pragma unroll
                            for (int i = 0; i < BM / 64; ++i)
                                ptx::tma_load_3d_cg<CG>(sa + i * (BK * 128), &maps.a[s], fb, m0 + i * 64, k0, ba);
                        }
                        if (p.conv_mode != 0) {
                            int c0, x0, y0, b0;
                            if (!B_MN) {
                                // rows = BNL consecutive positions, k-block = (tap, 64 channels): ONE box
                                const int kb = k0 / BK, tap = kb / p.cv_cblocks;
                                conv_coords(p, n0, tap, kb - tap * p.cv_cblocks, c0, x0, y0, b0);
                                ptx::tma_load_4d_cg<CG>(sb, &maps.b[s], fb, c0, x0, y0, b0);
                            } else {
                                // k-block = 64 consecutive positions; every 64-column atom is (tap, 64 channels)
#pragma unroll
                                for (int i = 0; i < BNL / 64; ++i) {
                                    const int cbi = (n0 + i * 64) / 64, tap = cbi / p.cv_cblocks;
                                    conv_coords(p, k0, tap, cbi - tap * p.cv_cblocks, c0, x0, y0, b0);
                                    ptx::tma_load_4d_cg<CG>(sb + i * (BK * 128), &maps.b[s], fb, c0, x0, y0, b0);
                                }
                            }
                        } else if (!B_MN) {
                            ptx::tma_load_3d_cg<CG>(sb, &maps.b[s], fb, k0, n0, bb);
                        } else {
#pragma unroll
                            for (int i = 0; i < BNL / 64; ++i)
                                ptx::tma_load_3d_cg<CG>(sb + i * (BK * 128), &maps.b[s], fb, n0 + i * 64, k0, bb);
                        }
                        if (++stage == nstages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1 && leader) {
        // ===================== MMA issuer (leader CTA only) =====================
        constexpr uint32_t idesc = make_idesc(BN, A_MN, B_MN, BM * CG);
        // K-major : rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused.
        // MN-major: atoms of [64 mn x 8 k] = 1024 B; SBO = stride between 8-k groups (1024 B),
        //           LBO = stride between 64-wide mn atoms (BK * 128 B).
        constexpr uint32_t A_LBO = A_MN ? BK * 128 : 0, B_LBO = B_MN ? BK * 128 : 0;
        constexpr uint32_t A_KSTEP = A_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
        constexpr uint32_t B_KSTEP = B_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
        int stage = 0;
        uint32_t phase = 0;
        int local_iter = 0;
        for (int w = worker; w < work; w += num_workers, ++local_iter) {
            const int split = w / tiles;
            const int it0 = split * p.iters_per_split;
            const int it1 = min(it0 + p.iters_per_split, total_iters);
            const int n_pipe = (it1 - it0) * p.nseg;
            const int acc = local_iter & 1;
            const uint32_t acc_phase = (local_iter >> 1) & 1;
            ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 2);
            ptx::tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BN;
            for (int i = 0; i < n_pipe; ++i) {
                ptx::mbar_wait(&full_bar[stage], phase, 3);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint64_t da =
                        make_smem_desc(ptx::smem_u32(smem_a + stage * C::A_BYTES), A_LBO, 1024);
                    const uint64_t db =
                        make_smem_desc(ptx::smem_u32(smem_b + stage * C::B_BYTES), B_LBO, 1024);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        if (CG == 2)
                            ptx::umma_f16_ss_2sm(tmem_d, da + (uint64_t)(k * A_KSTEP), db + (uint64_t)(k * B_KSTEP),
                                                 idesc, (i > 0 || k > 0) ? 1u : 0u);
                        else
                            ptx::umma_f16_ss(tmem_d, da + (uint64_t)(k * A_KSTEP), db + (uint64_t)(k * B_KSTEP),
                                             idesc, (i > 0 || k > 0) ? 1u : 0u);
                    }
                    if (CG == 2) {  // completion is signalled to the same barrier in BOTH CTAs
                        ptx::umma_commit_2sm(&empty_bar[stage], 0b11);
                        if (i == n_pipe - 1) ptx::umma_commit_2sm(&tmem_full[acc], 0b11);
                    } else {
                        ptx::umma_commit(&empty_bar[stage]);
                        if (i == n_pipe - 1) ptx::umma_commit(&tmem_full[acc]);
                    }
                }
                __syncwarp();
                if (++stage == nstages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp - 4;  // TMEM lane quarter this warp may read
        float* scratch = epi_scratch + q * (8192 / 4);
        uint8_t* tile0 = reinterpret_cast<uint8_t*>(scratch);  // two 1024B-aligned 4 KB tiles
        int store_parity = 0;
        int local_iter = 0;
        for (int w = worker; w < work; w += num_workers, ++local_iter) {
            const int split = w / tiles;
            const int t = w - split * tiles;
            const int ob = t / tiles_per_batch;
            const int r = t - ob * tiles_per_batch;
            int m_blk, n_blk;
            tile_coords(r, p.num_m, p.num_n, m_blk, n_blk);
            const int m0 = m_blk * (BM * CG) + (int)cta_rank * BM;
            const int n0 = n_blk * BN;
            const int acc = local_iter & 1;
            const uint32_t acc_phase = (local_iter >> 1) & 1;
            // per-column bias: lane l holds bias[col0 + l] of the NEXT chunk (one coalesced 128 B load,
            // issued a whole chunk ahead -- the first one before the accumulator wait) and the 32 values
            // are broadcast with shuffles, so no global-load latency sits between tcgen05.ld and the store
            const bool col_bias = p.bias != nullptr && !p.bias_per_row;
            float bias_next = 0.f;
            if (col_bias && n0 + lane < p.N) bias_next = __ldg(p.bias + n0 + lane);
            ptx::mbar_wait(&tmem_full[acc], acc_phase, 4);
            ptx::tc_fence_after();
            const int row_base = m0 + q * 32;
            const bool rows_live = row_base < p.M;
            constexpr int NCHUNK = BN / 32;
#pragma unroll 1
            for (int c = 0; c < NCHUNK; ++c) {
                const int col0 = n0 + c * 32;
                const float bias_cur = bias_next;
                if (col_bias && c + 1 < NCHUNK && col0 + 32 + lane < p.N) bias_next = __ldg(p.bias + col0 + 32 + lane);
                else bias_next = 0.f;
                if (col0 >= p.N || !rows_live) {
                    if (c == NCHUNK - 1) {
                        ptx::tc_fence_before();
                        if (CG == 2) ptx::mbar_arrive_cluster(&tmem_empty[acc], 0); else ptx::mbar_arrive(&tmem_empty[acc]);
                    }
                    continue;
                }
                uint32_t v[32];
                if (p.debug < 2) {
                    ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + c * 32, v);
                    ptx::tmem_ld_wait();
                }
                if (c == NCHUNK - 1) {
                    // accumulator fully drained into registers: hand the TMEM buffer back
                    ptx::tc_fence_before();
                    if (CG == 2) ptx::mbar_arrive_cluster(&tmem_empty[acc], 0); else ptx::mbar_arrive(&tmem_empty[acc]);
                }
                if (p.debug >= 1) continue;
                if (p.tma_store) {
                    // ---- fast path: registers -> swizzled smem tile -> TMA store (coalescing, tail
                    // clipping and the fp32 writes are done by the copy engine, not by LSU traffic).
                    // Split-K partials take the same path into the workspace ([split][batch][M][N], raw accumulators;
                    // alpha / bias / activation are applied by the finishing kernel)
                    const bool partial = p.splits > 1;
                    float x[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = partial ? __uint_as_float(v[j]) : p.alpha * __uint_as_float(v[j]);
                    if (!partial && p.bias != nullptr) {
                        if (p.bias_per_row) {
                            const int row = row_base + lane;
                            const float rb = row < p.M ? p.bias[row] : 0.f;
#pragma unroll
                            for (int j = 0; j < 32; ++j) x[j] += rb;
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) x[j] += __shfl_sync(0xffffffffu, bias_cur, j);  // 0 past N
                        }
                    }
                    const bool has_z = !partial && p.Z != nullptr;
                    // with a Z side-output each chunk uses both tiles (wait for all reads);
                    // otherwise the two tiles double-buffer the D stores
                    uint8_t* tile_d = has_z ? tile0 : tile0 + store_parity * 4096;
                    uint8_t* tile_z = tile0 + 4096;
                    if (lane == 0) {
                        if (has_z) ptx::tma_store_wait_read<0>(); else ptx::tma_store_wait_read<1>();
                    }
                    __syncwarp();
                    const uint32_t sw = (uint32_t)(lane & 7);
                    if (has_z) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<float4*>(tile_z + lane * 128 + ((j ^ sw) << 4)) =
                                make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
                    }
                    if (!partial && p.act != NNB_ACT_NONE) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] = apply_act(x[j], p.act, p.beta);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(tile_d + lane * 128 + ((j ^ sw) << 4)) =
                            make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    if (ptx::elect_one()) {
                        if (partial) {
                            ptx::tma_store_3d(&maps.d, tile_d, col0, row_base, split * p.out_batches + ob);
                        } else if (p.col_group > 0) {  // grouped columns: (column in group, row, group)
                            const int grp = col0 / p.col_group;
                            ptx::tma_store_3d(&maps.d, tile_d, col0 - grp * p.col_group, row_base, grp);
                        } else {
                            ptx::tma_store_3d(&maps.d, tile_d, col0, row_base, ob);
                            if (has_z) ptx::tma_store_3d(&maps.z, tile_z, col0, row_base, ob);
                        }
                        ptx::tma_store_commit();
                    }
                    store_parity ^= 1;
                    continue;
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) scratch[lane * 33 + j] = __uint_as_float(v[j]);
                __syncwarp();
                const int col = col0 + lane;
                const bool col_ok = col < p.N;
                if (p.splits > 1) {
                    float* wsp = p.ws + ((long long)(split * p.out_batches + ob) * p.M) * p.N;
#pragma unroll 4
                    for (int rr = 0; rr < 32; ++rr) {
                        const int row = row_base + rr;
                        if (row < p.M && col_ok)
                            wsp[(long long)row * p.N + col] = scratch[rr * 33 + lane];
                    }
                } else {
                    long long coff;
                    if (p.col_group > 0) {
                        const int g = col / p.col_group;
                        coff = (long long)g * p.group_stride + (col - g * p.col_group);
                    } else {
                        coff = col;
                    }
                    coff += (long long)ob * p.batch_stride_d;
                    const float cb = bias_cur;  // bias[col0 + lane] (0 past N or without a column bias)
#pragma unroll 4
                    for (int rr = 0; rr < 32; ++rr) {
                        const int row = row_base + rr;
                        if (row < p.M && col_ok) {
                            float x = p.alpha * scratch[rr * 33 + lane];
                            if (p.bias != nullptr) x += p.bias_per_row ? p.bias[row] : cb;
                            const long long off = (long long)row * p.ldd + coff;
                            if (p.Z != nullptr) p.Z[off] = x;
                            x = apply_act(x, p.act, p.beta);
                            p.D[off] = x;
                            if (p.D16 != nullptr)
                                p.D16[(long long)row * p.ldd16 + col] = __float2bfloat16(x);
                        }
                    }
                }
                __syncwarp();
            }
        }
        if (lane == 0) ptx::tma_store_wait_read<0>();  // smem tiles must outlive the bulk stores
    }

    ptx::tc_fence_before();
    if (CG == 2) ptx::cluster_sync(); else __syncthreads();
    if (p.clk_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
        p.clk_out[0] = clock64() - clk0;
        p.clk_out[1] = ptx::globaltimer_ns() - ns0;
    }
    if (warp == 2) {
        if (CG == 2) ptx::tmem_dealloc_2sm<C::TMEM_COLS>(tmem_base); else ptx::tmem_dealloc<C::TMEM_COLS>(tmem_base);
    }
}

// Sum split-K partials and apply the epilogue.
__global__ void splitk_reduce_kernel(const KArgs p) {
    pdl_trigger();
    pdl_wait();
    const long long per = (long long)p.M * p.N;
    const long long total = per * p.out_batches;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int s = 0; s < p.splits; ++s) acc += p.ws[(long long)s * total + i];
        const int ob = (int)(i / per);
        const long long rcol = i - (long long)ob * per;
        const int row = (int)(rcol / p.N);
        const int col = (int)(rcol - (long long)row * p.N);
        long long coff;
        if (p.col_group > 0) {
            const int g = col / p.col_group;
            coff = (long long)g * p.group_stride + (col - g * p.col_group);
        } else {
            coff = col;
        }
        coff += (long long)ob * p.batch_stride_d;
        float x = p.alpha * acc;
        if (p.bias != nullptr) x += p.bias_per_row ? p.bias[row] : p.bias[col];
        const long long off = (long long)row * p.ldd + coff;
        if (p.Z != nullptr) p.Z[off] = x;
        x = apply_act(x, p.act, p.beta);
        p.D[off] = x;
        if (p.D16 != nullptr) p.D16[(long long)row * p.ldd16 + col] = __float2bfloat16(x);
    }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// inner = contiguous dimension. box = {64 x box_rows x 1}.
int encode_map(CUtensorMap* m, const __nv_bfloat16* ptr, int64_t inner, int64_t rows, int64_t ld,
               int64_t batch, int64_t batch_stride, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(NNB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t bs = (batch > 1) ? (cuuint64_t)batch_stride * 2 : (cuuint64_t)rows * ld * 2;
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, bs};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (strides[0] & 15) || (strides[1] & 15))
        return fail(NNB_ERR_INVALID, "staged operand not 16-byte aligned for TMA");
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)ptr, dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(NNB_ERR_CUDA,
                    "cuTensorMapEncodeTiled failed (%d) inner=%lld rows=%lld ld=%lld batch=%lld",
                    (int)r, (long long)inner, (long long)rows, (long long)ld, (long long)batch);
    return NNB_OK;
}

// fp32 output map for the TMA-store epilogue: box 32 cols x 32 rows x 1 batch.
int encode_out_map(CUtensorMap* m, const float* ptr, int64_t cols, int64_t rows, int64_t ld,
                   int64_t batch, int64_t batch_stride) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(NNB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t bs = (batch > 1) ? (cuuint64_t)batch_stride * 4 : (cuuint64_t)rows * ld * 4;
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, bs};
    cuuint32_t box[3] = {32, 32, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)ptr, dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NNB_ERR_CUDA, "cuTensorMapEncodeTiled (output) failed (%d)", (int)r);
    return NNB_OK;
}

// R consecutive positions (m aligned to R) of a P x Q grid as one box (bw x bh x bb) of (q, p, image)
bool conv_rect(int64_t R, int64_t P, int64_t Q, int sp0, int sp1, int* bw, int* bh, int* bb) {
    int64_t w = std::min<int64_t>(Q, R);
    if (Q % w != 0 || R % w != 0) return false;
    int64_t h = std::min<int64_t>(P, R / w);
    if (w < Q && h != 1) return false;        // a partial row cannot continue on the next line
    if (P % h != 0 || (R / w) % h != 0) return false;
    int64_t b = R / (w * h);
    if (h < P && b != 1) return false;        // a partial image cannot continue in the next image
    if (w * sp1 > 256 || h * sp0 > 256 || b > 256) return false;  // TMA box limits
    *bw = (int)w; *bh = (int)h; *bb = (int)b;
    return true;
}

// NHWC planes [B][Hs][Ws][C] as a 4-D tensor (c, x, y, b); box = 64 channels x (bw, bh, bb) pixels with the conv
// stride as TMA traversal stride (boxDim = N * stride loads N elements, cuda.h cuTensorMapEncodeTiled)
int encode_conv_map(CUtensorMap* m, const __nv_bfloat16* ptr, const ConvOperand& cv, int bw, int bh, int bb) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(NNB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {(cuuint64_t)cv.C, (cuuint64_t)cv.Ws, (cuuint64_t)cv.Hs, (cuuint64_t)cv.B};
    cuuint64_t strides[3] = {(cuuint64_t)cv.C * 2, (cuuint64_t)cv.Ws * cv.C * 2, (cuuint64_t)cv.Hs * cv.Ws * cv.C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(bw * cv.sp1), (cuuint32_t)(bh * cv.sp0), (cuuint32_t)bb};
    cuuint32_t es[4] = {1, (cuuint32_t)cv.sp1, (cuuint32_t)cv.sp0, 1};
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (strides[0] & 15))
        return fail(NNB_ERR_INVALID, "conv planes not 16-byte aligned for TMA");
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)ptr, dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(NNB_ERR_CUDA, "cuTensorMapEncodeTiled (conv planes) failed (%d): C=%lld W=%lld H=%lld B=%lld box=%dx%dx%d stride %dx%d",
                    (int)r, (long long)cv.C, (long long)cv.Ws, (long long)cv.Hs, (long long)cv.B, bw, bh, bb, cv.sp0, cv.sp1);
    return NNB_OK;
}

template <int BN, bool A_MN, bool B_MN, int CG>
int launch(const GemmMaps& maps, const KArgs& ka, int grid, cudaStream_t stream) {
    using C = Cfg<BN, CG>;
    auto kfn = gemm_tcgen05_kernel<BN, A_MN, B_MN, CG>;
    static bool configured = false;  // per instantiation
    if (!configured) {
        NNB_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (CG == 2) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = CG;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    NNB_CUDA_OK(cudaLaunchKernelEx(&cfg, kfn, maps, ka));
    count_launch();
    return NNB_OK;
}

template <int BN, int CG>
int launch_major(bool a_mn, bool b_mn, const GemmMaps& maps, const KArgs& ka, int grid,
                 cudaStream_t stream) {
    if (!a_mn && !b_mn) return launch<BN, false, false, CG>(maps, ka, grid, stream);
    if (a_mn && !b_mn) return launch<BN, true, false, CG>(maps, ka, grid, stream);
    if constexpr (BN / CG >= 64) {
        if (!a_mn && b_mn) return launch<BN, false, true, CG>(maps, ka, grid, stream);
        return launch<BN, true, true, CG>(maps, ka, grid, stream);
    } else {
        return fail(NNB_ERR_UNSUPPORTED, "MN-major B needs >= 64 columns per CTA");
    }
}

}  // namespace

bool gemm_conv_supported(int mode, int64_t C, int64_t P, int64_t Q, int sp0, int sp1) {
    if (C <= 0 || C % 64 != 0 || P <= 0 || Q <= 0) return false;
    int a, b, c;
    if (mode == 2) return conv_rect(64, P, Q, sp0, sp1, &a, &b, &c);
    // mode 1: some tile width must be tileable (the search in gemm() skips the others)
    for (int r : {32, 64, 128, 256})
        if (conv_rect(r, P, Q, sp0, sp1, &a, &b, &c)) return true;
    return false;
}

size_t gemm_splitk_ws_bytes(int64_t M, int64_t N, int64_t K, int64_t batch) {
    // worst case the heuristic may pick: <= 2 * SMs partial tiles of 128 x 256 worth of output
    (void)K;
    const int64_t max_splits = 64;
    int64_t per = M * N * batch * 4;
    int64_t cap = (int64_t)2 * 160 * 128 * 256 * 4 + per;  // splits are chosen so partials stay below this
    return (size_t)std::min(per * max_splits, cap);
}

int gemm(const GemmProblem& g, cudaStream_t stream) {
    NNB_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0 && g.batch > 0, "gemm: non-positive dimension");
    NNB_REQUIRE(g.D != nullptr, "gemm: null output");
    NNB_REQUIRE(g.M < (1ll << 31) && g.N < (1ll << 31) && g.K < (1ll << 31), "gemm: dim too large");
    const bool conv = g.conv.mode != 0;
    if (conv) {
        NNB_REQUIRE(g.conv.mode == 1 || g.conv.mode == 2, "gemm: bad conv mode");
        NNB_REQUIRE(g.conv.hi && g.conv.C > 0 && g.conv.C % 64 == 0, "gemm: conv planes need C %% 64 == 0");
        NNB_REQUIRE(g.batch == 1 && !g.reduce_batch, "gemm: conv mode is unbatched");
        NNB_REQUIRE((g.conv.mode == 2) == g.B.mn_major, "gemm: conv mode / B major mismatch");
        const int64_t pos = g.conv.B * g.conv.P * g.conv.Q, kc = (int64_t)g.conv.kh * g.conv.kw * g.conv.C;
        NNB_REQUIRE(g.conv.mode == 1 ? (g.N == pos && g.K == kc) : (g.K == pos && g.N == kc), "gemm: conv shape mismatch");
        NNB_REQUIRE(pos < (1ll << 31), "gemm: too many conv positions");
    }
    const bool x3 = g.A.st.lo != nullptr && (conv ? g.conv.lo != nullptr : g.B.st.lo != nullptr);
    const bool a_mn = g.A.mn_major, b_mn = g.B.mn_major;
    const bool reduce_batch = g.reduce_batch;
    const int64_t out_batches = reduce_batch ? 1 : g.batch;
    const int64_t red_batches = reduce_batch ? g.batch : 1;
    const int sms = num_sms();

    // ---- tile width and split-K factor from a small cycle model (constants from the measured
    // B200/B300 pacing: UMMA 128xNx16 = max(N/2, 40) cycles, ~100 B/cycle/SM operand fill,
    // TMA-store epilogue ~160 cycles per 32-column chunk, HBM ~3400 B/cycle chip-wide)
    const int64_t total_iters_all = ceil_div(g.K, BK) * red_batches;
    const int nseg_h = x3 ? 3 : 1;
    const bool tma_ok_h = (g.col_group == 0 || ((g.col_group % 32) == 0 && (g.N % g.col_group) == 0 && g.epi.Z == nullptr &&
                                                (g.group_stride % 4) == 0)) &&
                          (g.ldd % 4) == 0 && (reinterpret_cast<uintptr_t>(g.D) & 15) == 0;
    int bn = 0, cg = 1;
    int64_t splits = 1;
    {
        static const int env_cg = [] { const char* e = getenv("NNB_GEMM_CG"); return e ? atoi(e) : 0; }();
        static const int verbose = [] { const char* e = getenv("NNB_GEMM_VERBOSE"); return e ? atoi(e) : 0; }();
        static const int env_bn = [] { const char* e = getenv("NNB_GEMM_BN"); return e ? atoi(e) : 0; }();
        static const int env_splits = [] { const char* e = getenv("NNB_GEMM_SPLITS"); return e ? atoi(e) : 0; }();
        const int force_splits = g.force_splits ? g.force_splits : env_splits;
        int force_cg = g.force_cg ? g.force_cg : env_cg;
        int force_bn = g.force_bn ? g.force_bn : env_bn;
        if (force_cg == 2 && g.M <= BM) force_cg = 1;  // a single row-block cannot use a CTA pair
      retry_search:
        const int cands[4] = {256, 128, 64, 32};
        const int scand[12] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64};
        double best = 1e300;
        for (int cgi = 1; cgi <= 2; ++cgi) {
            if (force_cg && cgi != force_cg) continue;
            if (cgi == 2 && (g.M <= BM || sms < 2)) continue;  // a pair needs two row-blocks to be useful
            for (int ci = 0; ci < 4; ++ci) {
                const int c = cands[ci];
                if (force_bn && c != force_bn) continue;
                if (cgi == 2 && c < 128) continue;
                if (b_mn && c / cgi < 64) continue;
                if (g.conv.mode == 1) {  // the CTA's B rows must form one pixel box
                    int t0, t1, t2;
                    if (!conv_rect(c / cgi, g.conv.P, g.conv.Q, g.conv.sp0, g.conv.sp1, &t0, &t1, &t2)) continue;
                }
                if (!force_bn && c > 32 && c / 2 >= g.N && !(b_mn && c / cgi == 64)) continue;  // half the width covers N
                const int64_t t = ceil_div(g.M, BM * cgi) * ceil_div(g.N, c) * out_batches;
                for (int si = 0; si < 12; ++si) {
                    const int64_t sp = scand[si];
                    if (force_splits && sp != force_splits) continue;
                    if (sp > 1 && (sp > total_iters_all || total_iters_all / sp < 2)) continue;
                    if (sp > 1) {
                        const size_t need = (size_t)sp * out_batches * g.M * g.N * 4;
                        if (g.splitk_ws == nullptr || g.splitk_ws_bytes < need) continue;
                    }
                    const double waves = (double)ceil_div(t * sp, sms / cgi);
                    const double iters = (double)ceil_div(total_iters_all, sp) * nseg_h;
                    const double mma = iters * 4.0 * std::max(c / 2.0, 40.0);
                    // shared-memory traffic per k-block: TMA fill + UMMA operand reads of the local tiles
                    const double fill = iters * 2.0 * (16384.0 + (c / cgi) * 128.0) / 128.0;
                    const bool fast_epi = sp == 1 ? tma_ok_h : ((g.N % 4) == 0);  // split-K partials also leave through TMA stores
                    const double epi = (c / 32) * (fast_epi ? 160.0 : 1300.0) + 300.0;
                    double cyc = waves * std::max(std::max(mma, fill), epi) + 700.0 + 3000.0 + (cgi == 2 ? 2000.0 : 0.0);
                    if (sp > 1) {
                        const double out_bytes = (double)g.M * g.N * out_batches * 4.0;
                        cyc += (2.0 * sp + 1.0) * out_bytes / 3400.0 + 7000.0;  // partial traffic + the dependent finishing launch (measured: profiles/r1_linear_ladder.md, wgrad sweep)
                    }
                    if (cyc < best) { best = cyc; bn = c; splits = sp; cg = cgi; }
                }
            }
        }
        if (verbose)
            fprintf(stderr, "[nnb gemm] M=%lld N=%lld K=%lld batch=%lld majors=%d%d x3=%d -> BN=%d CG=%d splits=%lld (model %.0f cycles)\n",
                    (long long)g.M, (long long)g.N, (long long)g.K, (long long)g.batch, (int)a_mn, (int)b_mn, (int)x3, bn, cg,
                    (long long)splits, best);
        if (bn == 0 && (env_cg || env_bn) && !g.force_cg && !g.force_bn && force_cg + force_bn != 0) {
            force_cg = 0;  // environment overrides are best-effort: fall back to the model's choice
            force_bn = 0;
            goto retry_search;
        }
        if (bn == 0) {
            if (g.force_splits > 1)
                return fail(NNB_ERR_WORKSPACE, "gemm: split-K workspace too small for %d splits", g.force_splits);
            return fail(NNB_ERR_UNSUPPORTED, "gemm: no tile configuration for this problem");
        }
    }
    NNB_REQUIRE(bn == 32 || bn == 64 || bn == 128 || bn == 256, "gemm: bad BN %d", bn);
    NNB_REQUIRE(!(b_mn && bn / cg < 64), "gemm: MN-major B needs >= 64 columns per CTA");

    const int64_t num_m = ceil_div(g.M, BM * cg), num_n = ceil_div(g.N, bn);
    const int64_t tiles = num_m * num_n * out_batches;
    const int64_t kblocks = ceil_div(g.K, BK);
    const int64_t total_iters = kblocks * red_batches;
    int64_t ips = ceil_div(total_iters, splits);
    splits = ceil_div(total_iters, ips);

    // ---- tensor maps
    GemmMaps maps;
    std::memset(&maps, 0, sizeof(maps));
    const int nseg = x3 ? 3 : 1;
    // segment order: small terms first (lo*hi, hi*lo), then hi*hi
    const __nv_bfloat16* a_planes[3];
    const __nv_bfloat16* b_planes[3];
    const __nv_bfloat16* b_hi = conv ? g.conv.hi : g.B.st.hi;
    const __nv_bfloat16* b_lo = conv ? g.conv.lo : g.B.st.lo;
    if (x3) {
        a_planes[0] = g.A.st.lo; b_planes[0] = b_hi;
        a_planes[1] = g.A.st.hi; b_planes[1] = b_lo;
        a_planes[2] = g.A.st.hi; b_planes[2] = b_hi;
    } else {
        a_planes[0] = g.A.st.hi; b_planes[0] = b_hi;
    }
    int cv_bw = 0, cv_bh = 0, cv_bb = 0;
    if (conv && !conv_rect(g.conv.mode == 1 ? bn / cg : 64, g.conv.P, g.conv.Q, g.conv.sp0, g.conv.sp1, &cv_bw, &cv_bh, &cv_bb))
        return fail(NNB_ERR_UNSUPPORTED, "gemm: conv position grid %lldx%lld cannot be tiled by pixel boxes", (long long)g.conv.P, (long long)g.conv.Q);
    const bool a_bcast = g.batch > 1 && g.A.st.batch <= 1;
    const bool b_bcast = g.batch > 1 && g.B.st.batch <= 1;
    for (int s = 0; s < nseg; ++s) {
        int rc;
        // K-major: staged [rows = M][cols = K] -> inner K, box rows = BM.
        // MN-major: staged [rows = K][cols = M] -> inner M, box = 64 x BK.
        if (!a_mn) {
            NNB_REQUIRE(g.A.st.rows == g.M && g.A.st.cols == g.K, "gemm: A staged shape mismatch");
            rc = encode_map(&maps.a[s], a_planes[s], g.K, g.M, g.A.st.ld, a_bcast ? 1 : g.batch,
                            g.A.st.batch_stride, BM);
        } else {
            NNB_REQUIRE(g.A.st.rows == g.K && g.A.st.cols == g.M, "gemm: A staged shape mismatch (MN)");
            rc = encode_map(&maps.a[s], a_planes[s], g.M, g.K, g.A.st.ld, a_bcast ? 1 : g.batch,
                            g.A.st.batch_stride, BK);
        }
        if (rc) return rc;
        if (conv) {
            rc = encode_conv_map(&maps.b[s], b_planes[s], g.conv, cv_bw, cv_bh, cv_bb);
        } else if (!b_mn) {
            NNB_REQUIRE(g.B.st.rows == g.N && g.B.st.cols == g.K, "gemm: B staged shape mismatch");
            rc = encode_map(&maps.b[s], b_planes[s], g.K, g.N, g.B.st.ld, b_bcast ? 1 : g.batch,
                            g.B.st.batch_stride, bn / cg);
        } else {
            NNB_REQUIRE(g.B.st.rows == g.K && g.B.st.cols == g.N, "gemm: B staged shape mismatch (MN)");
            rc = encode_map(&maps.b[s], b_planes[s], g.N, g.K, g.B.st.ld, b_bcast ? 1 : g.batch,
                            g.B.st.batch_stride, BK);
        }
        if (rc) return rc;
    }

    KArgs ka;
    std::memset(&ka, 0, sizeof(ka));
    ka.M = (int)g.M; ka.N = (int)g.N;
    ka.kblocks = (int)kblocks;
    ka.out_batches = (int)out_batches;
    ka.red_batches = (int)red_batches;
    ka.a_bcast = a_bcast; ka.b_bcast = b_bcast;
    ka.nseg = nseg;
    ka.splits = (int)splits; ka.iters_per_split = (int)ips;
    ka.num_m = (int)num_m; ka.num_n = (int)num_n;
    ka.D = g.D; ka.ldd = g.ldd; ka.batch_stride_d = reduce_batch ? 0 : g.batch_stride_d;
    ka.col_group = 0; ka.group_stride = 0;
    ka.alpha = g.epi.alpha; ka.bias = g.epi.bias; ka.bias_per_row = 0;
    ka.Z = g.epi.Z; ka.act = g.epi.act; ka.beta = g.epi.beta;
    ka.D16 = g.epi.D16; ka.ldd16 = g.epi.ldd16;
    ka.ws = g.splitk_ws;
    if (g.col_group > 0) { ka.col_group = (int)g.col_group; ka.group_stride = g.group_stride; }
    ka.bias_per_row = g.bias_per_row ? 1 : 0;
    {
        static const int dbg = [] { const char* e = getenv("NNB_GEMM_DEBUG"); return e ? atoi(e) : 0; }();
        ka.debug = dbg;
        static const int st = [] { const char* e = getenv("NNB_GEMM_STAGES"); return e ? atoi(e) : 0; }();
        ka.stages = st;
        ka.clk_out = g.clk_out;
    }
    ka.bias_vec = g.epi.bias && (reinterpret_cast<uintptr_t>(g.epi.bias) & 15) == 0;
    if (conv) {
        ka.conv_mode = g.conv.mode;
        ka.cv_P = (int)g.conv.P; ka.cv_Q = (int)g.conv.Q;
        ka.cv_bw = cv_bw; ka.cv_bh = cv_bh;
        ka.cv_sp0 = g.conv.sp0; ka.cv_sp1 = g.conv.sp1; ka.cv_off0 = g.conv.off0; ka.cv_off1 = g.conv.off1;
        ka.cv_dk0 = g.conv.dk0; ka.cv_dk1 = g.conv.dk1; ka.cv_kw = g.conv.kw;
        ka.cv_cblocks = (int)(g.conv.C / 64);
    }
    {
        auto ok16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
        const int64_t bsd = reduce_batch ? 0 : g.batch_stride_d;
        bool tma = splits == 1 && g.col_group == 0 && g.epi.D16 == nullptr && ok16(g.D) &&
                   (g.ldd % 4) == 0 && (out_batches == 1 || (bsd % 4) == 0) &&
                   (g.epi.Z == nullptr || ok16(g.epi.Z));
        // column-group outputs (conv: D[o][(b, hw)] -> NCHW) go through TMA too when a 32-column chunk never
        // straddles two groups: the group index becomes the third tensor-map coordinate
        const bool tma_grouped = splits == 1 && g.col_group > 0 && (g.col_group % 32) == 0 &&
                                 (g.N % g.col_group) == 0 && out_batches == 1 && g.epi.D16 == nullptr &&
                                 g.epi.Z == nullptr && ok16(g.D) && (g.ldd % 4) == 0 && (g.group_stride % 4) == 0;
        if (tma) {
            int rc2 = encode_out_map(&maps.d, g.D, g.N, g.M, g.ldd, out_batches, bsd);
            if (!rc2 && g.epi.Z) rc2 = encode_out_map(&maps.z, g.epi.Z, g.N, g.M, g.ldd, out_batches, bsd);
            if (rc2) return rc2;
        } else if (tma_grouped) {
            int rc2 = encode_out_map(&maps.d, g.D, g.col_group, g.M, g.ldd, g.N / g.col_group, g.group_stride);
            if (rc2) return rc2;
        }
        bool tma_partial = false;
        if (splits > 1 && (g.N % 4) == 0 && ok16(g.splitk_ws)) {
            // split-K partials through the TMA-store epilogue too: the map covers the workspace as [split * batch][M][N]
            int rc2 = encode_out_map(&maps.d, g.splitk_ws, g.N, g.M, g.N, splits * out_batches, g.M * g.N);
            if (rc2) return rc2;
            tma_partial = true;
        }
        ka.tma_store = (tma || tma_grouped || tma_partial) ? 1 : 0;
    }

    const int64_t work = tiles * splits;
    const int grid = (int)std::min<int64_t>(work, sms / cg) * cg;
    int rc;
    if (cg == 2) {
        if (bn == 128) rc = launch_major<128, 2>(a_mn, b_mn, maps, ka, grid, stream);
        else rc = launch_major<256, 2>(a_mn, b_mn, maps, ka, grid, stream);
    } else {
        switch (bn) {
            case 32: rc = launch_major<32, 1>(a_mn, b_mn, maps, ka, grid, stream); break;
            case 64: rc = launch_major<64, 1>(a_mn, b_mn, maps, ka, grid, stream); break;
            case 128: rc = launch_major<128, 1>(a_mn, b_mn, maps, ka, grid, stream); break;
            default: rc = launch_major<256, 1>(a_mn, b_mn, maps, ka, grid, stream); break;
        }
    }
    if (rc) return rc;
    if (splits > 1) {
        const int64_t total = g.M * g.N * out_batches;
        const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)sms * 8);
        NNB_CUDA_OK(launch_pdl(splitk_reduce_kernel, dim3(blocks), dim3(256), 0, stream, ka));
        count_launch();
        NNB_CUDA_OK(cudaGetLastError());
    }
    return NNB_OK;
}

}  // namespace nnb
