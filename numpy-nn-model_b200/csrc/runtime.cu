// Library-wide plumbing: error string, launch counter, device queries.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

namespace nnb {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static std::atomic<int> g_pdl{[] { const char* e = getenv("NNB_PDL"); return e ? atoi(e) : 1; }()};
bool pdl_enabled() { return g_pdl.load(std::memory_order_relaxed) != 0; }
int set_pdl(int on) { return g_pdl.exchange(on ? 1 : 0); }

static const bool g_nvtx = [] { const char* e = getenv("NNB_NVTX"); return e ? atoi(e) != 0 : true; }();
bool nvtx_enabled() { return g_nvtx; }
void nvtx_push(const char* name) { nvtxRangePushA(name); }
void nvtx_pop() { nvtxRangePop(); }

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// SM budget for persistent grids: data-parallel training wants a few SMs left for NCCL's CTAs, which cannot
// share an SM with a GEMM CTA (230 KB of dynamic shared memory) and would otherwise only run between GEMMs.
static std::atomic<int> g_sm_budget{[] { const char* e = getenv("NNB_SM_BUDGET"); return e ? atoi(e) : 0; }()};
int set_sm_budget(int n) { return g_sm_budget.exchange(n > 0 ? n : 0); }

static int device_sms() {
    static int sms[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (sms[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
            v = 148;
        sms[dev] = v;
    }
    return sms[dev];
}

int num_sms() {
    const int all = device_sms(), b = g_sm_budget.load(std::memory_order_relaxed);
    return (b > 0 && b < all) ? std::max(b, 2) : all;
}

}  // namespace nnb

extern "C" {

const char* nnb_last_error(void) { return nnb::g_err; }

int nnb_version(void) { return 100; }

int nnb_device_check(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0, maj = 0, min = 0, sms = 0;
    NNB_CUDA_OK(cudaGetDevice(&dev));
    NNB_CUDA_OK(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
    NNB_CUDA_OK(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
    NNB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    if (maj != 10)
        return nnb::fail(NNB_ERR_UNSUPPORTED,
                         "libneunet_b200 is built for sm_100a only; device %d is sm_%d%d", dev, maj,
                         min);
    return NNB_OK;
}

int nnb_set_pdl(int on) { return nnb::set_pdl(on); }
int nnb_set_sm_budget(int sms) { return nnb::set_sm_budget(sms); }

uint64_t nnb_launch_count(void) { return nnb::g_launches.load(); }
void nnb_launch_count_reset(void) { nnb::g_launches.store(0); }

}  // extern "C"
