// nccl_group.cpp -- single-process, multi-device NCCL communicator group, bound with dlopen.
//
// The reference has no NCCL anywhere; its direct evaluators are confined to one MPI rank
// (src/core/system.cpp:618-623), so the multi-GPU form of the drop-in is ONE process driving P devices:
// ncclCommInitAll + one grouped ncclAllGather of source strengths per evaluation.
#include "skb_internal.hpp"

#include <dlfcn.h>
#include <vector>

namespace skb {

namespace {
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t; // ncclSuccess == 0
enum { kNcclFloat64 = 8 }; // ncclDataType_t::ncclFloat64 / ncclDouble (nccl.h)

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

int load_api(NcclApi &api) {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle)
            break;
    }
    if (!api.handle)
        return set_error(SKB_ERR_NCCL, "multi-GPU context needs NCCL: dlopen(libnccl.so.2) failed: %s", dlerror());
#define LOAD(field, sym)                                                                                              \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym));                                        \
    if (!api.field)                                                                                                   \
        return set_error(SKB_ERR_NCCL, "NCCL symbol %s not found", sym);
    LOAD(CommInitAll, "ncclCommInitAll")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(AllGather, "ncclAllGather")
    LOAD(ReduceScatter, "ncclReduceScatter")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    return SKB_OK;
}

struct NcclGroup {
    NcclApi api;
    std::vector<ncclComm_t> comms;
    std::vector<int> devs;
};
} // namespace

int nccl_group_create(size_t n, const std::function<int(int)> &device_of, void **out) {
    NcclGroup *g = new NcclGroup;
    int rc = load_api(g->api);
    if (rc != SKB_OK) {
        delete g;
        return rc;
    }
    g->devs.resize(n);
    for (size_t i = 0; i < n; ++i)
        g->devs[i] = device_of((int)i);
    g->comms.resize(n, nullptr);
    ncclResult_t r = g->api.CommInitAll(g->comms.data(), (int)n, g->devs.data());
    if (r != 0) {
        rc = set_error(SKB_ERR_NCCL, "ncclCommInitAll(%zu) failed: %s", n, g->api.GetErrorString(r));
        delete g;
        return rc;
    }
    *out = g;
    return SKB_OK;
}

// in-place all-gather: rank g's contribution already sits at bufs[g] + g*count
int nccl_group_allgather_inplace(void *group, void *const *bufs, size_t count, const cudaStream_t *streams) {
    NcclGroup *g = static_cast<NcclGroup *>(group);
    if (!g)
        return set_error(SKB_ERR_STATE, "NCCL group missing");
    ncclResult_t r = g->api.GroupStart();
    for (size_t i = 0; r == 0 && i < g->comms.size(); ++i) {
        cudaSetDevice(g->devs[i]);
        double *base = static_cast<double *>(bufs[i]);
        r = g->api.AllGather(base + i * count, base, count, kNcclFloat64, g->comms[i], streams[i]);
    }
    ncclResult_t r2 = g->api.GroupEnd();
    if (r == 0)
        r = r2;
    if (r != 0)
        return set_error(SKB_ERR_NCCL, "ncclAllGather failed: %s", g->api.GetErrorString(r));
    return SKB_OK;
}

int nccl_group_reduce_scatter(void *group, void *const *send, void *const *recv, size_t count,
                              const cudaStream_t *streams) {
    NcclGroup *g = static_cast<NcclGroup *>(group);
    if (!g)
        return set_error(SKB_ERR_STATE, "NCCL group missing");
    enum { kNcclSum = 0 }; // ncclRedOp_t::ncclSum
    ncclResult_t r = g->api.GroupStart();
    for (size_t i = 0; r == 0 && i < g->comms.size(); ++i) {
        cudaSetDevice(g->devs[i]);
        r = g->api.ReduceScatter(send[i], recv[i], count, kNcclFloat64, kNcclSum, g->comms[i], streams[i]);
    }
    ncclResult_t r2 = g->api.GroupEnd();
    if (r == 0)
        r = r2;
    if (r != 0)
        return set_error(SKB_ERR_NCCL, "ncclReduceScatter failed: %s", g->api.GetErrorString(r));
    return SKB_OK;
}

void nccl_group_destroy(void *group) {
    NcclGroup *g = static_cast<NcclGroup *>(group);
    if (!g)
        return;
    for (size_t i = 0; i < g->comms.size(); ++i)
        if (g->comms[i]) {
            cudaSetDevice(g->devs[i]);
            g->api.CommDestroy(g->comms[i]);
        }
    delete g;
}

} // namespace skb
