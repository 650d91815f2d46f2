// Stand-alone GPU self-test + micro-benchmark for the Linear path of libneunet_b200.so.
// Run on a B200 (gpurun): build/test_gemm [quick]
// Checks forward / dgrad / wgrad / db against a double-precision host reference on sampled
// entries, for both precisions, and times the bare tcgen05 GEMM on a shape ladder.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "common.cuh"

using namespace nnb;

#define CK(x)                                                                       \
    do {                                                                            \
        cudaError_t e = (x);                                                        \
        if (e != cudaSuccess) {                                                     \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(2);                                                                \
        }                                                                           \
    } while (0)
#define NK(x)                                                          \
    do {                                                               \
        int rc = (x);                                                  \
        if (rc) {                                                      \
            printf("nnb error %d: %s (%s:%d)\n", rc, nnb_last_error(), __FILE__, __LINE__); \
            exit(3);                                                   \
        }                                                              \
    } while (0)

static std::vector<float> rnd(size_t n, float lo, float hi, unsigned seed) {
    std::mt19937 g(seed);
    std::uniform_real_distribution<float> d(lo, hi);
    std::vector<float> v(n);
    for (auto& x : v) x = d(g);
    return v;
}
static float* dev(const std::vector<float>& h) {
    float* d;
    CK(cudaMalloc(&d, h.size() * 4 + 16));
    CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    return d;
}
static std::vector<float> host(const float* d, size_t n) {
    std::vector<float> h(n);
    CK(cudaMemcpy(h.data(), d, n * 4, cudaMemcpyDeviceToHost));
    return h;
}
static double swish(double x, double b) { return x / (1.0 + std::exp(-b * x)); }
static double swish_g(double z, double b) {
    double s = 1.0 / (1.0 + std::exp(-b * z)), f = z * s;
    return b * f + s * (1 - b * f);
}

static int g_fail = 0;

struct Err {
    double max_abs = 0, max_ref = 0;
    void add(double got, double ref) {
        max_abs = std::max(max_abs, std::fabs(got - ref));
        max_ref = std::max(max_ref, std::fabs(ref));
    }
    double rel() const { return max_abs / std::max(max_ref, 1e-30); }
};

static void report(const char* what, const Err& e, double tol) {
    const bool ok = e.rel() <= tol && std::isfinite(e.rel());
    printf("    %-6s max|err|=%.3e  max|ref|=%.3e  rel=%.3e  tol=%.1e  %s\n", what, e.max_abs,
           e.max_ref, e.rel(), tol, ok ? "ok" : "FAIL");
    if (!ok) g_fail++;
}

static void test_linear(int64_t M, int64_t K, int64_t N, int prec, int act, unsigned seed) {
    const float beta = 1.5f;
    printf("linear M=%lld K=%lld N=%lld prec=%s act=%s\n", (long long)M, (long long)K, (long long)N,
           prec ? "bf16x3" : "bf16", act ? "swish" : "none");
    auto X = rnd(M * K, -1, 1, seed), W = rnd(N * K, -1.f / std::sqrt((float)K), 1.f / std::sqrt((float)K), seed + 1),
         b = rnd(N, -0.5f, 0.5f, seed + 2), G = rnd(M * N, -1, 1, seed + 3);
    float *dX = dev(X), *dW = dev(W), *db = dev(b), *dG = dev(G);
    float *dO, *dZ, *dgX, *dgW, *dgb;
    CK(cudaMalloc(&dO, M * N * 4)); CK(cudaMalloc(&dZ, M * N * 4));
    CK(cudaMalloc(&dgX, M * K * 4)); CK(cudaMalloc(&dgW, N * K * 4)); CK(cudaMalloc(&dgb, N * 4));
    CK(cudaMemset(dO, 0xFF, M * N * 4)); CK(cudaMemset(dgX, 0xFF, M * K * 4)); CK(cudaMemset(dgW, 0xFF, N * K * 4));
    size_t wsb = std::max(nnb_linear_workspace_bytes(M, K, N, prec, 0), nnb_linear_workspace_bytes(M, K, N, prec, 1));
    void* ws; CK(cudaMalloc(&ws, wsb));
    void* wst; CK(cudaMalloc(&wst, nnb_weight_staged_bytes(N, K, prec)));
    NK(nnb_stage_weight(dW, N, K, prec, wst, 0));
    NK(nnb_linear_forward(dX, dW, db, dO, act ? dZ : nullptr, M, K, N, act, beta, prec, wst, nullptr, ws, wsb, 0));
    NK(nnb_linear_backward(dX, dW, act ? dZ : nullptr, dG, dgX, dgW, dgb, M, K, N, act, beta, prec, nullptr, nullptr, ws, wsb, 0));
    CK(cudaDeviceSynchronize());
    auto O = host(dO, M * N), gX = host(dgX, M * K), gW = host(dgW, N * K), gb = host(dgb, N);
    std::vector<float> Z;
    if (act) Z = host(dZ, M * N);

    const double tol = prec ? 2e-5 : 6e-3;
    std::mt19937 g(seed + 9);
    const int samples = 3000;
    Err ef, ez, ex, ew, eb;
    // forward samples
    for (int s = 0; s < samples; ++s) {
        int64_t m = g() % M, n = g() % N;
        if (s < 8) { m = (s & 1) ? M - 1 : 0; n = (s & 2) ? N - 1 : 0; }
        double a = b[n];
        for (int64_t k = 0; k < K; ++k) a += (double)X[m * K + k] * W[n * K + k];
        if (act) { ez.add(Z[m * N + n], a); a = swish(a, beta); }
        ef.add(O[m * N + n], a);
    }
    // dZ on host (full) only when needed lazily per sample
    auto dz = [&](int64_t m, int64_t n) -> double {
        double gg = G[m * N + n];
        if (act) gg *= swish_g(Z[m * N + n], beta);  // uses the device Z (already validated above)
        return gg;
    };
    for (int s = 0; s < samples; ++s) {
        int64_t m = g() % M, k = g() % K;
        if (s < 8) { m = (s & 1) ? M - 1 : 0; k = (s & 2) ? K - 1 : 0; }
        double a = 0;
        for (int64_t n = 0; n < N; ++n) a += dz(m, n) * W[n * K + k];
        ex.add(gX[m * K + k], a);
    }
    const int wsamples = (M > 2048) ? 600 : samples;
    for (int s = 0; s < wsamples; ++s) {
        int64_t n = g() % N, k = g() % K;
        if (s < 8) { n = (s & 1) ? N - 1 : 0; k = (s & 2) ? K - 1 : 0; }
        double a = 0;
        for (int64_t m = 0; m < M; ++m) a += dz(m, n) * X[m * K + k];
        ew.add(gW[n * K + k], a);
    }
    for (int64_t n = 0; n < std::min<int64_t>(N, 64); ++n) {
        double a = 0;
        for (int64_t m = 0; m < M; ++m) a += dz(m, n);
        eb.add(gb[n], a);
    }
    report("fwd", ef, tol);
    if (act) report("Z", ez, tol);
    report("dX", ex, tol);
    report("dW", ew, tol);
    report("db", eb, 1e-5);
    cudaFree(dX); cudaFree(dW); cudaFree(db); cudaFree(dG); cudaFree(dO); cudaFree(dZ);
    cudaFree(dgX); cudaFree(dgW); cudaFree(dgb); cudaFree(ws); cudaFree(wst);
}

static bool g_graph = false;
// Time the bare GEMM (operands already staged) for the three Linear forms.
static void bench_gemm(int64_t M, int64_t K, int64_t N, int prec, int iters) {
    auto X = rnd(M * K, -1, 1, 1), W = rnd(N * K, -1, 1, 2), G = rnd(M * N, -1, 1, 3);
    float *dX = dev(X), *dW = dev(W), *dG = dev(G);
    const int np = prec ? 2 : 1;
    auto alloc_st = [&](const float* src, int64_t r, int64_t c) {
        Staged s;
        __nv_bfloat16 *hi, *lo = nullptr;
        CK(cudaMalloc(&hi, staged_plane_bytes(1, r, c)));
        if (np == 2) CK(cudaMalloc(&lo, staged_plane_bytes(1, r, c)));
        NK(stage_operand(view2d(src, r, c, c), false, prec, hi, lo, STAGE_COPY, nullptr, 0, nullptr, nullptr, 0, &s));
        return s;
    };
    Staged xs = alloc_st(dX, M, K), wsd = alloc_st(dW, N, K), gs = alloc_st(dG, M, N);
    float* out; CK(cudaMalloc(&out, std::max({M * N, M * K, N * K}) * 4));
    size_t skb = 256 << 20; float* sk; CK(cudaMalloc(&sk, skb));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    unsigned long long* clk; CK(cudaMalloc(&clk, 16));
    const char* names[3] = {"fwd  X.W^T", "dgrad g.W ", "wgrad g^T.X"};
    for (int form = 0; form < 3; ++form) {
        GemmProblem g;
        if (form == 0) { g.M = M; g.N = N; g.K = K; g.A.st = xs; g.B.st = wsd; g.ldd = N; }
        if (form == 1) { g.M = M; g.N = K; g.K = N; g.A.st = gs; g.B.st = wsd; g.B.mn_major = true; g.ldd = K; }
        if (form == 2) { g.M = N; g.N = K; g.K = M; g.A.st = gs; g.A.mn_major = true; g.B.st = xs; g.B.mn_major = true; g.ldd = K; }
        g.D = out; g.splitk_ws = sk; g.splitk_ws_bytes = skb; g.clk_out = clk;
        for (int i = 0; i < 3; ++i) NK(gemm(g, 0));
        float ms;
        if (g_graph) {
            // GPU-only pacing: `iters` launches captured once into a CUDA graph (no host tensor-map
            // encoding or launch latency between kernels), replayed 3 times, last replay timed
            cudaStream_t st; CK(cudaStreamCreate(&st));
            cudaGraph_t gr; cudaGraphExec_t ge;
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            for (int i = 0; i < iters; ++i) NK(gemm(g, st));
            CK(cudaStreamEndCapture(st, &gr));
            CK(cudaGraphInstantiate(&ge, gr, 0));
            for (int r = 0; r < 2; ++r) CK(cudaGraphLaunch(ge, st));
            CK(cudaEventRecord(e0, st));
            CK(cudaGraphLaunch(ge, st));
            CK(cudaEventRecord(e1, st));
            CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ms, e0, e1);
            cudaGraphExecDestroy(ge); cudaGraphDestroy(gr); cudaStreamDestroy(st);
        } else {
            CK(cudaEventRecord(e0));
            for (int i = 0; i < iters; ++i) NK(gemm(g, 0));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ms, e0, e1);
        }
        const double t = ms / iters * 1e-3;
        unsigned long long hclk[2];
        CK(cudaMemcpy(hclk, clk, 16, cudaMemcpyDeviceToHost));
        const double mhz = hclk[1] ? (double)hclk[0] / (double)hclk[1] * 1e3 : 0.0;
        const double peak = 148.0 * 8192.0 * mhz * 1e6 * 1e-12;  // TFLOP/s at the observed SM clock
        printf("  gemm %s M=%lld K=%lld N=%lld %s: %.1f us  %.1f TFLOP/s (algorithmic 2MKN)  [SM clock %.0f MHz -> tensor peak %.0f TFLOP/s, %.0f%%]\n", names[form],
               (long long)M, (long long)K, (long long)N, prec ? "bf16x3" : "bf16", t * 1e6, 2.0 * M * K * N / t * 1e-12,
               mhz, peak, peak > 0 ? 100.0 * (2.0 * M * K * N / t * 1e-12) / peak : 0.0);
    }
    cudaFree(dX); cudaFree(dW); cudaFree(dG); cudaFree(out); cudaFree(sk);
    cudaFree((void*)xs.hi); cudaFree((void*)wsd.hi); cudaFree((void*)gs.hi);
    if (np == 2) { cudaFree((void*)xs.lo); cudaFree((void*)wsd.lo); cudaFree((void*)gs.lo); }
}

int main(int argc, char** argv) {
    if (argc >= 6 && !strcmp(argv[1], "gbench")) { g_graph = true; argv[1] = (char*)"bench"; }  // graph-replayed pacing
    if (argc >= 6 && !strcmp(argv[1], "bench")) {  // test_gemm bench M K N prec [iters]
        int sms0, maj0, min0;
        NK(nnb_device_check(&sms0, &maj0, &min0));
        bench_gemm(atoll(argv[2]), atoll(argv[3]), atoll(argv[4]), atoi(argv[5]), argc > 6 ? atoi(argv[6]) : 20);
        return 0;
    }
    const bool quick = argc > 1 && !strcmp(argv[1], "quick");
    int sms, maj, min;
    NK(nnb_device_check(&sms, &maj, &min));
    printf("device: sm_%d%d, %d SMs\n", maj, min, sms);
    struct S { int64_t M, K, N; };
    std::vector<S> shapes = {{128, 256, 512}, {256, 64, 64}, {300, 200, 130}, {4096, 784, 128}, {4096, 128, 10}, {1000, 72, 24}};
    if (!quick) { shapes.push_back({2048, 1024, 4096}); shapes.push_back({777, 513, 1031}); }
    unsigned seed = 100;
    for (auto s : shapes)
        for (int prec = 0; prec < 2; ++prec)
            for (int act = 0; act < 2; ++act) {
                if (act && s.M > 1000 && prec == 0) continue;
                test_linear(s.M, s.K, s.N, prec, act, seed += 10);
            }
    printf("== correctness: %s (%d failing checks)\n", g_fail ? "FAILED" : "PASSED", g_fail);
    if (!quick) {
        bench_gemm(4096, 1024, 4096, 0, 20);
        bench_gemm(8192, 8192, 8192, 0, 10);
        bench_gemm(16384, 512, 2048, 0, 20);
        bench_gemm(4096, 784, 128, 0, 50);
        bench_gemm(4096, 1024, 4096, 1, 10);
    }
    return g_fail ? 1 : 0;
}
