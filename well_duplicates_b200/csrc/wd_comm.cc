// Multi-GPU entry points of the C ABI: one NCCL communicator per context, one
// ncclAllReduce(int64, sum) of the published counter rows over NVLink.
//
// Replaces the reference's cross-lane "reduction" -- one OS process per lane
// and a `tail` over their output files (Snakefile.count_dups:146-160) -- for
// callers that do not bring torch.distributed: rank 0 makes the unique id, the
// caller carries its 128 bytes to the other ranks, every rank joins.
//
// libnccl is loaded on first use (dlopen "libnccl.so.2": the copy a host
// process such as PyTorch already mapped is reused, otherwise the system one),
// so the library loads -- and every single-GPU entry point works -- on a box
// without NCCL.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <mutex>

#include "wd_common.cuh"

namespace wd {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    int (*GetVersion)(int *) = nullptr;
};

static NcclApi g_nccl;
static std::once_flag g_nccl_once;
static char g_nccl_why[256] = "not tried";          // why libnccl could not be used

static const NcclApi *nccl() {
    std::call_once(g_nccl_once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (h == nullptr) {
            const char *why = dlerror();
            snprintf(g_nccl_why, sizeof(g_nccl_why), "%s", why ? why : "dlopen failed");
            return;
        }
        NcclApi a;
        a.handle = h;
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
        a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
        if (a.GetUniqueId && a.CommInitRank && a.AllReduce && a.CommDestroy && a.GetErrorString) g_nccl = a;
        else snprintf(g_nccl_why, sizeof(g_nccl_why), "libnccl lacks one of the five entry points used");
    });
    return g_nccl.handle ? &g_nccl : nullptr;
}

#define WD_NCCL(api, call)                                                                          \
    do {                                                                                            \
        ncclResult_t r__ = (call);                                                                  \
        if (r__ != ncclSuccess) {                                                                   \
            wd::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, (api)->GetErrorString(r__)); \
            return WD_E_CUDA;                                                                       \
        }                                                                                           \
    } while (0)

int comm_destroy(wd_ctx *ctx) {
    if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
    ctx->comm_pending[0] = ctx->comm_pending[1] = false;
    if (ctx->comm != nullptr) {
        const NcclApi *api = nccl();
        if (api) api->CommDestroy(static_cast<ncclComm_t>(ctx->comm));
        ctx->comm = nullptr;
    }
    ctx->comm_rank = 0;
    ctx->comm_ranks = 1;
    return WD_OK;
}

}  // namespace wd

using namespace wd;

extern "C" {

int wd_comm_unique_id(void *id128) {
    if (id128 == nullptr) WD_FAIL(WD_E_ARG, "wd_comm_unique_id: null output");
    const NcclApi *api = nccl();
    if (api == nullptr) WD_FAIL(WD_E_CUDA, "wd_comm: libnccl.so.2 cannot be used (%s)", g_nccl_why);
    static_assert(sizeof(ncclUniqueId) == WD_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    WD_NCCL(api, api->GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return WD_OK;
}

int wd_comm_init(wd_ctx *ctx, const void *id128, int rank, int nranks) {
    if (ctx == nullptr || id128 == nullptr) WD_FAIL(WD_E_ARG, "wd_comm_init: null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) WD_FAIL(WD_E_ARG, "wd_comm_init: rank %d of %d", rank, nranks);
    const NcclApi *api = nccl();
    if (api == nullptr) WD_FAIL(WD_E_CUDA, "wd_comm: libnccl.so.2 cannot be used (%s)", g_nccl_why);
    WD_CUDA(cudaSetDevice(ctx->device));
    comm_destroy(ctx);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    WD_NCCL(api, api->CommInitRank(&comm, nranks, id, rank));
    ctx->comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_ranks = nranks;
    return WD_OK;
}

int wd_comm_destroy(wd_ctx *ctx) {
    if (ctx == nullptr) return WD_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    return comm_destroy(ctx);
}

int wd_allreduce_i64(wd_ctx *ctx, void *buf, size_t n) {
    if (ctx == nullptr) WD_FAIL(WD_E_ARG, "wd_allreduce_i64: null context");
    const bool published = buf == nullptr;
    if (published) {
        if (ctx->publish_n == 0) WD_FAIL(WD_E_ARG, "wd_allreduce_i64: nothing has been published (wd_publish_counters)");
        n = ctx->publish_n;
        buf = ctx->publish.as<unsigned long long>() + (size_t)ctx->publish_cur * n;
    }
    if (ctx->comm == nullptr) {
        if (ctx->comm_ranks == 1) return WD_OK;              // a single rank: the sum is the buffer
        WD_FAIL(WD_E_ARG, "wd_allreduce_i64: call wd_comm_init first");
    }
    const NcclApi *api = nccl();
    WD_CUDA(cudaSetDevice(ctx->device));
    if (!published) {
        WD_NCCL(api, api->AllReduce(buf, buf, n, ncclInt64, ncclSum, static_cast<ncclComm_t>(ctx->comm), ctx->stream));
        return WD_OK;
    }
    // The published rows are reduced on a stream of their own, behind the kernel that wrote them: the next
    // wd_count on the context's stream does not wait for the collective (its rows go to the other buffer).
    if (ctx->comm_stream == nullptr) {
        int lo = 0, hi = 0;
        WD_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        WD_CUDA(cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, hi));
    }
    const int cur = ctx->publish_cur;
    WD_CUDA(cudaEventRecord(ctx->pub_ready, ctx->stream));
    WD_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->pub_ready, 0));
    WD_NCCL(api, api->AllReduce(buf, buf, n, ncclInt64, ncclSum, static_cast<ncclComm_t>(ctx->comm), ctx->comm_stream));
    WD_CUDA(cudaEventRecord(ctx->comm_done[cur], ctx->comm_stream));
    ctx->comm_pending[cur] = true;
    return WD_OK;
}

int wd_comm_join(wd_ctx *ctx) {
    if (ctx == nullptr) WD_FAIL(WD_E_ARG, "wd_comm_join: null context");
    WD_CUDA(cudaSetDevice(ctx->device));
    for (int b = 0; b < 2; ++b)
        if (ctx->comm_pending[b]) WD_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->comm_done[b], 0));
    return WD_OK;
}

int wd_published_fetch(wd_ctx *ctx, int64_t *rows, size_t n_int64) {
    if (ctx == nullptr || rows == nullptr) WD_FAIL(WD_E_ARG, "wd_published_fetch: null argument");
    if (ctx->publish_n == 0 || n_int64 != ctx->publish_n)
        WD_FAIL(WD_E_ARG, "wd_published_fetch: %zu values published, caller asks for %zu", ctx->publish_n, n_int64);
    WD_CUDA(cudaSetDevice(ctx->device));
    const int cur = ctx->publish_cur;
    if (ctx->comm_pending[cur]) WD_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->comm_done[cur], 0));
    WD_CUDA(cudaMemcpyAsync(rows, ctx->publish.as<unsigned long long>() + (size_t)cur * n_int64, n_int64 * 8,
                            cudaMemcpyDeviceToHost, ctx->stream));
    WD_CUDA(cudaStreamSynchronize(ctx->stream));
    return WD_OK;
}

}  // extern "C"
