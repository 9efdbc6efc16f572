#include "comm.cuh"

#include <dlfcn.h>
#include <stdlib.h>
#include <nccl.h>

#include <mutex>

namespace pamnet {
namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // the copy torch.distributed loaded, if any
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy && api.GetErrorString;
    });
    return &api;
}

constexpr int kMaxDev = 16;
struct Comm {
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    int world = 1;
    bool enabled = false;
};
std::mutex g_mu;
Comm g_comm[kMaxDev];

Comm* current() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return nullptr;
    return &g_comm[dev];
}

#define PAMNET_NCCL(expr)                                                                                    \
    do {                                                                                                     \
        ncclResult_t _r = (expr);                                                                            \
        if (_r != ncclSuccess) {                                                                             \
            set_error("%s failed: %s", #expr, nccl_api()->GetErrorString(_r));                               \
            return -1;                                                                                       \
        }                                                                                                    \
    } while (0)

}  // namespace

int comm_unique_id(void* out128) {
    NcclApi* a = nccl_api();
    PAMNET_CHECK_ARG(a->ok, "NCCL is not available (libnccl.so.2 could not be loaded)");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    PAMNET_NCCL(a->GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return 0;
}

int comm_init(const void* id128, int rank, int world) {
    NcclApi* a = nccl_api();
    PAMNET_CHECK_ARG(a->ok, "NCCL is not available (libnccl.so.2 could not be loaded)");
    PAMNET_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "comm_init: rank %d of %d", rank, world);
    std::lock_guard<std::mutex> lk(g_mu);
    Comm* c = current();
    PAMNET_CHECK_ARG(c != nullptr, "comm_init: bad device");
    if (c->comm) { a->CommDestroy(c->comm); c->comm = nullptr; }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    PAMNET_NCCL(a->CommInitRank(&c->comm, world, id, rank));
    if (!c->stream) {
        // default: lowest priority -- the collective fills SMs the layer loop leaves idle instead of competing with it
        // (PAMNET_COMM_PRIO=high for the opposite)
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        const char* e = getenv("PAMNET_COMM_PRIO");
        PAMNET_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, (e && strcmp(e, "high") == 0) ? hi : lo));
    }
    c->world = world;
    c->enabled = true;
    return 0;
}

int comm_enable(int on) {
    std::lock_guard<std::mutex> lk(g_mu);
    Comm* c = current();
    if (c && c->comm) c->enabled = on != 0;
    return 0;
}

int comm_destroy() {
    std::lock_guard<std::mutex> lk(g_mu);
    Comm* c = current();
    if (c && c->comm) {
        nccl_api()->CommDestroy(c->comm);
        c->comm = nullptr;
        c->enabled = false;
    }
    return 0;
}

bool comm_active() {
    std::lock_guard<std::mutex> lk(g_mu);
    Comm* c = current();
    return c && c->comm && c->enabled && c->world > 1;
}

cudaStream_t comm_stream() {
    std::lock_guard<std::mutex> lk(g_mu);
    Comm* c = current();
    return c ? c->stream : nullptr;
}

int comm_allreduce_avg(float* buf, int64_t count) {
    if (count <= 0) return 0;
    ncclComm_t comm;
    cudaStream_t st;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        Comm* c = current();
        PAMNET_CHECK_ARG(c && c->comm, "comm_allreduce: no communicator");
        comm = c->comm; st = c->stream;
    }
    PAMNET_NCCL(nccl_api()->AllReduce(buf, buf, (size_t)count, ncclFloat32, ncclAvg, comm, st));
    count_launch();
    return 0;
}

}  // namespace pamnet
