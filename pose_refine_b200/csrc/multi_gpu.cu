// multi_gpu.cu -- sharding a batch of pose hypotheses over the GPUs of one node, through the C ABI.
//
// The reference has no multi-GPU code at all; its README (README.md:15) only says "call it from many
// host threads".  The path shards by hypothesis with no collective in the data path (hypotheses are independent, the
// scene and the mesh are read-only), so what a host needs is: a shard plan, the scene depth image on every GPU
// (ncclBroadcast, once per scene) and the 72-byte results of every shard on every GPU (ncclAllGather, once per
// batch, off the compute stream).  One process or thread per GPU.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): the library has no link-time dependency on it, a process that
// already loaded an NCCL (e.g. the one bundled with PyTorch) shares it, and single-GPU users need none.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>
#include <mutex>
#include <new>

struct pr_comm {
    ncclComm_t comm = nullptr;
    bool owned = false;
    int nranks = 1, rank = 0;
    cudaStream_t side = nullptr;      // the gather runs here, off the caller's compute stream
    cudaEvent_t ready = nullptr;      // recorded on the caller's stream: results of the batch are complete
    cudaEvent_t gathered = nullptr;   // recorded on the side stream: the all-gather is complete
};

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    bool ok = false;
};

const NcclApi* nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
        api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(dlsym(h, "ncclBroadcast"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Broadcast && api.AllGather;
    });
    return api.ok ? &api : nullptr;
}

inline int nccl_status(ncclResult_t r) { return r == ncclSuccess ? PR_OK : PR_ERR_COMM; }

int finish_comm(pr_comm* c) {
    cudaError_t e = cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->gathered, cudaEventDisableTiming);
    return e == cudaSuccess ? PR_OK : (int)e;
}

}  // namespace

extern "C" {

int pr_device_count(int* count) {
    if (!count) return PR_ERR_INVALID_ARGUMENT;
    PR_CUDA_TRY(cudaGetDeviceCount(count));
    return PR_OK;
}
int pr_set_device(int device) {
    PR_CUDA_TRY(cudaSetDevice(device));
    return PR_OK;
}

int pr_shard_plan(size_t n_hyp, int nranks, int rank, size_t* begin, size_t* count) {
    if (nranks <= 0 || rank < 0 || rank >= nranks || !begin || !count) return PR_ERR_INVALID_ARGUMENT;
    // contiguous shards whose sizes differ by at most one; the first n_hyp % nranks ranks hold the extra hypothesis
    const size_t base = n_hyp / (size_t)nranks, rem = n_hyp % (size_t)nranks, r = (size_t)rank;
    *begin = r * base + (r < rem ? r : rem);
    *count = base + (r < rem ? 1 : 0);
    return PR_OK;
}

int pr_nccl_unique_id(void* id128) {
    if (!id128) return PR_ERR_INVALID_ARGUMENT;
    const NcclApi* api = nccl();
    if (!api) return PR_ERR_COMM;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId");
    ncclUniqueId id;
    const int rc = nccl_status(api->GetUniqueId(&id));
    if (rc == PR_OK) memcpy(id128, &id, 128);
    return rc;
}

int pr_comm_create(pr_comm** out, const void* id128, int nranks, int rank) {
    if (!out || !id128 || nranks <= 0 || rank < 0 || rank >= nranks) return PR_ERR_INVALID_ARGUMENT;
    const NcclApi* api = nccl();
    if (!api) return PR_ERR_COMM;
    pr_comm* c = new (std::nothrow) pr_comm();
    if (!c) return PR_ERR_INVALID_ARGUMENT;
    c->nranks = nranks; c->rank = rank; c->owned = true;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    int rc = nccl_status(api->CommInitRank(&c->comm, nranks, id, rank));
    if (rc == PR_OK) rc = finish_comm(c);
    if (rc != PR_OK) { pr_comm_destroy(c); return rc; }
    *out = c;
    return PR_OK;
}

int pr_comm_adopt(pr_comm** out, void* nccl_comm, int nranks, int rank) {
    if (!out || !nccl_comm || nranks <= 0 || rank < 0 || rank >= nranks) return PR_ERR_INVALID_ARGUMENT;
    if (!nccl()) return PR_ERR_COMM;
    pr_comm* c = new (std::nothrow) pr_comm();
    if (!c) return PR_ERR_INVALID_ARGUMENT;
    c->comm = static_cast<ncclComm_t>(nccl_comm); c->nranks = nranks; c->rank = rank; c->owned = false;
    const int rc = finish_comm(c);
    if (rc != PR_OK) { pr_comm_destroy(c); return rc; }
    *out = c;
    return PR_OK;
}

void pr_comm_destroy(pr_comm* c) {
    if (!c) return;
    if (c->side) { cudaStreamSynchronize(c->side); cudaStreamDestroy(c->side); }
    if (c->ready) cudaEventDestroy(c->ready);
    if (c->gathered) cudaEventDestroy(c->gathered);
    if (c->comm && c->owned) { const NcclApi* api = nccl(); if (api) api->CommDestroy(c->comm); }
    delete c;
}

int pr_broadcast_scene(pr_comm* c, void* buf_dev, size_t bytes, int root, pr_stream_t stream) {
    if (!c || !buf_dev || root < 0 || root >= c->nranks) return PR_ERR_INVALID_ARGUMENT;
    if (bytes == 0 || c->nranks == 1) return PR_OK;
    return nccl_status(nccl()->Broadcast(buf_dev, buf_dev, bytes, ncclUint8, root, c->comm, prb::as_stream(stream)));
}

int pr_gather_results(pr_comm* c, const pr_registration_result* local_dev, size_t n_per_rank, pr_registration_result* all_dev,
                      pr_stream_t stream) {
    if (!c || !local_dev || !all_dev) return PR_ERR_INVALID_ARGUMENT;
    if (n_per_rank == 0) return PR_OK;
    // the gather starts once everything queued on `stream` so far (the batch that produced local_dev) is complete,
    // and runs on the communicator's own stream: the caller's next batch overlaps it
    PR_CUDA_TRY(cudaEventRecord(c->ready, prb::as_stream(stream)));
    PR_CUDA_TRY(cudaStreamWaitEvent(c->side, c->ready, 0));
    const size_t bytes = n_per_rank * sizeof(pr_registration_result);
    int rc = PR_OK;
    if (c->nranks == 1) {
        if ((const void*)local_dev != (const void*)all_dev)
            PR_CUDA_TRY(cudaMemcpyAsync(all_dev, local_dev, bytes, cudaMemcpyDeviceToDevice, c->side));
    } else {
        rc = nccl_status(nccl()->AllGather(local_dev, all_dev, bytes, ncclUint8, c->comm, c->side));
    }
    if (rc != PR_OK) return rc;
    PR_CUDA_TRY(cudaEventRecord(c->gathered, c->side));
    return PR_OK;
}

int pr_gather_wait(pr_comm* c, pr_stream_t stream, int host_sync) {
    if (!c) return PR_ERR_INVALID_ARGUMENT;
    if (host_sync) { PR_CUDA_TRY(cudaEventSynchronize(c->gathered)); return PR_OK; }
    PR_CUDA_TRY(cudaStreamWaitEvent(prb::as_stream(stream), c->gathered, 0));
    return PR_OK;
}

}  // extern "C"
