// Gradient all-reduce inside the C-ABI (nnb_comm_init / allreduce / destroy).
//
// The reference has no distributed code; the north-star adds ONE collective -- a sum all-reduce of the flat fp32
// gradient bucket over NCCL / NVLink -- and a maintainer who binds libneunet_b200.so from the reference (ctypes, no
// torch) must get it from the same library. NCCL is not linked: the entry points are resolved at run time from the
// libnccl.so.2 already loaded in the process (torch's bundled copy when torch is present) or found by the dynamic
// loader, so `ldd libneunet_b200.so` stays free of NCCL and a machine without it gets NNB_ERR_UNSUPPORTED, never a
// load failure. One communicator = one rank = one GPU (the device current at nnb_comm_init).
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

namespace nnb {
namespace {

// the few NCCL declarations used (nccl.h is not required at build time)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;      // ncclSuccess == 0
constexpr int kNcclFloat32 = 7;  // ncclDataType_t: ncclFloat32
constexpr int kNcclSum = 0;      // ncclRedOp_t: ncclSum

struct Nccl {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    bool ok = false;
};

Nccl& nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {  // prefer the copy that is already mapped (torch's): one NCCL per process
            n.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
            if (n.handle) break;
        }
        if (!n.handle)
            for (const char* nm : names) {
                n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
                if (n.handle) break;
            }
        if (!n.handle) return;
        n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(dlsym(n.handle, "ncclGetUniqueId"));
        n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(dlsym(n.handle, "ncclCommInitRank"));
        n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(dlsym(n.handle, "ncclAllReduce"));
        n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(dlsym(n.handle, "ncclCommDestroy"));
        n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(dlsym(n.handle, "ncclGetErrorString"));
        n.GetVersion = reinterpret_cast<decltype(n.GetVersion)>(dlsym(n.handle, "ncclGetVersion"));
        n.ok = n.GetUniqueId && n.CommInitRank && n.AllReduce && n.CommDestroy;
    });
    return n;
}

int nccl_fail(const char* what, ncclResult_t r) {
    Nccl& n = nccl();
    return fail(NNB_ERR_CUDA, "%s failed: NCCL error %d (%s)", what, (int)r, n.GetErrorString ? n.GetErrorString(r) : "?");
}

}  // namespace
}  // namespace nnb

struct nnb_comm {
    nnb::ncclComm_t comm;
    int world, rank, device;
};

using namespace nnb;

extern "C" {

int nnb_comm_available(int* nccl_version) {
    Nccl& n = nccl();
    if (!n.ok) return 0;
    if (nccl_version) {
        *nccl_version = 0;
        if (n.GetVersion) n.GetVersion(nccl_version);
    }
    return 1;
}

int nnb_comm_unique_id(void* id_out) {
    NNB_REQUIRE(id_out, "nnb_comm_unique_id: null pointer");
    Nccl& n = nccl();
    if (!n.ok) return fail(NNB_ERR_UNSUPPORTED, "nnb_comm: libnccl.so.2 not found in this process or on the loader path");
    ncclUniqueId id;
    ncclResult_t r = n.GetUniqueId(&id);
    if (r != 0) return nccl_fail("ncclGetUniqueId", r);
    memcpy(id_out, id.internal, sizeof(id.internal));
    return NNB_OK;
}

int nnb_comm_init(nnb_comm** comm, const void* unique_id, int world, int rank) {
    NNB_REQUIRE(comm && unique_id, "nnb_comm_init: null pointer");
    NNB_REQUIRE(world > 0 && rank >= 0 && rank < world, "nnb_comm_init: bad world / rank");
    Nccl& n = nccl();
    if (!n.ok) return fail(NNB_ERR_UNSUPPORTED, "nnb_comm: libnccl.so.2 not found in this process or on the loader path");
    ncclUniqueId id;
    memcpy(id.internal, unique_id, sizeof(id.internal));
    nnb_comm* c = new nnb_comm{nullptr, world, rank, 0};
    cudaError_t ce = cudaGetDevice(&c->device);
    if (ce != cudaSuccess) { delete c; return fail(NNB_ERR_CUDA, "cudaGetDevice failed: %s", cudaGetErrorString(ce)); }
    ncclResult_t r = n.CommInitRank(&c->comm, world, id, rank);
    if (r != 0) { delete c; return nccl_fail("ncclCommInitRank", r); }
    *comm = c;
    return NNB_OK;
}

int nnb_comm_allreduce_sum(nnb_comm* comm, float* buf, int64_t n_elems, cudaStream_t stream) {
    NNB_RANGE("nnb_comm_allreduce_sum");
    NNB_REQUIRE(comm && comm->comm, "nnb_comm_allreduce_sum: null communicator");
    NNB_REQUIRE(buf && n_elems > 0, "nnb_comm_allreduce_sum: bad buffer");
    ncclResult_t r = nccl().AllReduce(buf, buf, (size_t)n_elems, kNcclFloat32, kNcclSum, comm->comm, stream);
    if (r != 0) return nccl_fail("ncclAllReduce", r);
    count_launch();
    return NNB_OK;
}

int nnb_comm_destroy(nnb_comm* comm) {
    if (!comm) return NNB_OK;
    ncclResult_t r = 0;
    if (comm->comm) r = nccl().CommDestroy(comm->comm);
    delete comm;
    if (r != 0) return nccl_fail("ncclCommDestroy", r);
    return NNB_OK;
}

}  // extern "C"
