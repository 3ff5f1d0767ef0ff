// nccl_bcast.cu -- the one collective on the hot path: broadcast of the packed reference
// planes from rank 0 to every GPU of the box over NVLink 5 / NVSwitch. Records and windows shard
// independently, so nothing else crosses devices.
//
// libnccl.so.2 is dlopen()ed on first use (override with PAVGPU_NCCL_LIB) so that libpavgpu.so loads
// on boxes without NCCL and so that a process that already holds torch's bundled NCCL reuses it.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace {

struct NcclUniqueId { char internal[128]; };
typedef void *NcclComm;
typedef int NcclResult;  // 0 == ncclSuccess
constexpr int NCCL_UINT8 = 1;

struct NcclApi {
    void *handle = nullptr;
    NcclResult (*GetUniqueId)(NcclUniqueId *) = nullptr;
    NcclResult (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    NcclResult (*Broadcast)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    NcclResult (*GroupStart)() = nullptr;
    NcclResult (*GroupEnd)() = nullptr;
    NcclResult (*CommDestroy)(NcclComm) = nullptr;
    const char *(*GetErrorString)(NcclResult) = nullptr;
};

NcclApi g_nccl;

int load_nccl()
{
    if (g_nccl.handle) return PAVGPU_OK;
    const char *names[] = {getenv("PAVGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { pav_set_error("cannot dlopen libnccl.so.2: %s", dlerror()); return PAVGPU_ERR_NCCL; }
#define LOAD(field, sym)                                                              \
    *(void **)(&g_nccl.field) = dlsym(h, sym);                                        \
    if (!g_nccl.field) { pav_set_error("libnccl: missing symbol %s", sym); dlclose(h); return PAVGPU_ERR_NCCL; }
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(Broadcast, "ncclBroadcast")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    g_nccl.handle = h;
    return PAVGPU_OK;
}

#define NCCL_TRY(expr)                                                                       \
    do {                                                                                     \
        NcclResult r_ = (expr);                                                              \
        if (r_ != 0) {                                                                       \
            pav_set_error("NCCL error %s at %s:%d (%s)", g_nccl.GetErrorString(r_), __FILE__, __LINE__, #expr); \
            return PAVGPU_ERR_NCCL;                                                          \
        }                                                                                    \
    } while (0)

// Communicators are kept for the life of the process, one per (device, rank, n_ranks): ncclCommInitRank costs seconds, the
// broadcast it serves milliseconds (r02, 2 x B200: 2.3-5 s of a 4.6 s distributed call were communicator set-up).
std::mutex g_comm_mu;
std::map<std::tuple<int, int, int>, NcclComm> g_comms;

}  // namespace

extern "C" __attribute__((visibility("default"))) int pavgpu_nccl_comm_cached(pavgpu_ctx *ctx, int32_t rank, int32_t n_ranks)
{
    if (!ctx) return 0;
    std::lock_guard<std::mutex> lk(g_comm_mu);
    return g_comms.count(std::make_tuple(ctx->device, (int)rank, (int)n_ranks)) ? 1 : 0;
}

extern "C" __attribute__((visibility("default"))) void pavgpu_nccl_comm_release_all(void)
{
    std::lock_guard<std::mutex> lk(g_comm_mu);
    for (auto &kv : g_comms)
        if (kv.second && g_nccl.CommDestroy) g_nccl.CommDestroy(kv.second);
    g_comms.clear();
}

extern "C" __attribute__((visibility("default"))) int pavgpu_nccl_unique_id(uint8_t id_out[128])
{
    if (!id_out) { pav_set_error("nccl_unique_id: NULL"); return PAVGPU_ERR_ARG; }
    int rc = load_nccl();
    if (rc) return rc;
    NcclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id_out, id.internal, 128);
    return PAVGPU_OK;
}

extern "C" __attribute__((visibility("default"))) int pavgpu_seqstore_broadcast(pavgpu_ctx *ctx, pavgpu_seqstore *store, const uint8_t id_in[128],
                                                                                  int32_t rank, int32_t n_ranks, float *ms_out)
{
    if (!ctx || !store || rank < 0 || rank >= n_ranks) { pav_set_error("seqstore_broadcast: bad argument"); return PAVGPU_ERR_ARG; }
    if (ms_out) *ms_out = 0.f;
    if (n_ranks == 1) return PAVGPU_OK;
    int rc = load_nccl();
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    // id_in == NULL: the communicator of an earlier call of this process with the same (device, rank, n_ranks); else a new one from id_in
    // (it replaces a cached one) -- every rank must make the same choice, pavgpu_nccl_comm_cached() tells which one is possible
    const auto key = std::make_tuple(ctx->device, (int)rank, (int)n_ranks);
    NcclComm comm = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_comm_mu);
        auto it = g_comms.find(key);
        if (!id_in) {
            if (it == g_comms.end()) { pav_set_error("seqstore_broadcast: no communicator cached for rank %d of %d and no unique id given", rank, n_ranks); return PAVGPU_ERR_ARG; }
            comm = it->second;
        } else {
            if (it != g_comms.end()) { g_nccl.CommDestroy(it->second); g_comms.erase(it); }
            NcclUniqueId id;
            memcpy(id.internal, id_in, 128);
            NCCL_TRY(g_nccl.CommInitRank(&comm, n_ranks, id, rank));
            g_comms[key] = comm;
        }
    }
    const bool fresh = id_in != nullptr;
    rc = [&]() -> int {
        if (fresh) {   // warm the communicator (channel setup) on a few bytes so the timed transfer is the transfer
            NCCL_TRY(g_nccl.Broadcast(store->d_nmask, store->d_nmask, 4, NCCL_UINT8, 0, comm, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        }
        CUDA_TRY(cudaEventRecord(ctx->ev[0], ctx->stream));
        NCCL_TRY(g_nccl.GroupStart());
        NCCL_TRY(g_nccl.Broadcast(store->d_pack2, store->d_pack2, store->pack2_bytes, NCCL_UINT8, 0, comm, ctx->stream));
        NCCL_TRY(g_nccl.Broadcast(store->d_nmask, store->d_nmask, store->nmask_bytes, NCCL_UINT8, 0, comm, ctx->stream));
        NCCL_TRY(g_nccl.GroupEnd());
        CUDA_TRY(cudaEventRecord(ctx->ev[1], ctx->stream));
        if (rank != 0) { int nrc = pav_build_nsum(store); if (nrc) return nrc; }   // receivers derive the N summary from the mask plane they got
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (ms_out) *ms_out = ev_ms(ctx->ev[0], ctx->ev[1]);
        return PAVGPU_OK;
    }();
    if (rc != PAVGPU_OK) {   // a communicator that failed is not reused
        std::lock_guard<std::mutex> lk(g_comm_mu);
        g_nccl.CommDestroy(comm);
        g_comms.erase(key);
    }
    return rc;
}
