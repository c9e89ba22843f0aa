// The one collective on the path: the gradient all-reduce of data-parallel training, inside the C ABI.
// One NCCL communicator per device behind an opaque bdetr_comm; libnccl is bound at run time with dlopen (the copy the
// host process already carries is preferred), so libbdetr.so keeps no link-time dependency and still loads on a
// GPU-less build host.  ncclAllReduce is stream-ordered and CUDA-graph capturable, like every other entry point.
// Reference: tf.distribute.MirroredStrategy (parameters.py:74) all-reduces the replica gradients with TF-internal NCCL.
#include <dlfcn.h>
#include <mutex>
#include <nccl.h>
#include "common.cuh"

namespace bdetr {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    const char *(*GetErrorString)(ncclResult_t);
    ncclResult_t (*GetVersion)(int *);
    bool ok = false;
};

static NcclApi *nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // already in the process (e.g. next to torch)
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        auto sym = [&](const char *n) { return dlsym(h, n); };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.Broadcast && api.CommDestroy && api.GetErrorString;
    });
    return api.ok ? &api : nullptr;
}

}  // namespace bdetr

struct bdetr_comm {
    ncclComm_t comm;
    int rank, world, device;
};

using namespace bdetr;
#define API extern "C" __attribute__((visibility("default")))
#define BDETR_NCCL(call)                                                                     \
    do {                                                                                     \
        ncclResult_t r__ = (call);                                                           \
        if (r__ != ncclSuccess) {                                                            \
            set_error("%s failed: %s", #call, api->GetErrorString(r__));                     \
            return BDETR_E_NCCL;                                                             \
        }                                                                                    \
    } while (0)

API int bdetr_comm_unique_id(void *id128)
{
    BDETR_REQUIRE(id128, BDETR_E_NULL, "null pointer");
    NcclApi *api = nccl_api();
    BDETR_REQUIRE(api, BDETR_E_NCCL, "libnccl.so.2 could not be loaded");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    BDETR_NCCL(api->GetUniqueId(reinterpret_cast<ncclUniqueId *>(id128)));
    return BDETR_OK;
}

API int bdetr_comm_init(bdetr_comm **comm, int rank, int world, const void *id128)
{
    BDETR_REQUIRE(comm && id128, BDETR_E_NULL, "null pointer");
    BDETR_REQUIRE(world >= 1 && rank >= 0 && rank < world, BDETR_E_BAD_SHAPE, "bad rank / world size");
    NcclApi *api = nccl_api();
    BDETR_REQUIRE(api, BDETR_E_NCCL, "libnccl.so.2 could not be loaded");
    bdetr_comm *c = new bdetr_comm();
    c->rank = rank; c->world = world;
    BDETR_CUDA(cudaGetDevice(&c->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclResult_t r = api->CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        set_error("ncclCommInitRank failed: %s", api->GetErrorString(r));
        delete c;
        return BDETR_E_NCCL;
    }
    *comm = c;
    return BDETR_OK;
}

API int bdetr_comm_destroy(bdetr_comm *comm)
{
    if (!comm) return BDETR_OK;
    NcclApi *api = nccl_api();
    BDETR_REQUIRE(api, BDETR_E_NCCL, "libnccl.so.2 could not be loaded");
    ncclResult_t r = api->CommDestroy(comm->comm);
    delete comm;
    if (r != ncclSuccess) { set_error("ncclCommDestroy failed: %s", api->GetErrorString(r)); return BDETR_E_NCCL; }
    return BDETR_OK;
}

API int bdetr_comm_info(const bdetr_comm *comm, int *rank, int *world, int *nccl_version)
{
    BDETR_REQUIRE(comm, BDETR_E_NULL, "null communicator");
    if (rank) *rank = comm->rank;
    if (world) *world = comm->world;
    if (nccl_version) { NcclApi *api = nccl_api(); *nccl_version = 0; if (api && api->GetVersion) api->GetVersion(nccl_version); }
    return BDETR_OK;
}

API int bdetr_allreduce(bdetr_comm *comm, float *buf, size_t count, void *stream)
{
    BDETR_REQUIRE(comm && buf, BDETR_E_NULL, "null pointer");
    if (count == 0 || comm->world == 1) return BDETR_OK;
    NcclApi *api = nccl_api();
    BDETR_REQUIRE(api, BDETR_E_NCCL, "libnccl.so.2 could not be loaded");
    BDETR_NCCL(api->AllReduce(buf, buf, count, ncclFloat32, ncclSum, comm->comm, as_stream(stream)));
    count_launch();
    return BDETR_OK;
}

API int bdetr_broadcast(bdetr_comm *comm, float *buf, size_t count, int root, void *stream)
{
    BDETR_REQUIRE(comm && buf, BDETR_E_NULL, "null pointer");
    if (count == 0 || comm->world == 1) return BDETR_OK;
    NcclApi *api = nccl_api();
    BDETR_REQUIRE(api, BDETR_E_NCCL, "libnccl.so.2 could not be loaded");
    BDETR_NCCL(api->Broadcast(buf, buf, count, ncclFloat32, root, comm->comm, as_stream(stream)));
    return BDETR_OK;
}
