// K6: the hot path's only collective -- SUM all-reduce of the packed metric accumulators (integer PCK / curve counts
// and float64 EPE / loss sums) over NCCL (NVLink 5 / NVSwitch), stream-ordered right after K5 and graph-capturable.
//
// The reference has no live counterpart (src/utils/comm.py is never initialised); the buffers are what its
// AverageMeters hold (src/utils/metric_logger.py:7-28, src/utils/method.py:241-250) and what pred_eval counts
// (src/utils/argparser.py:347-378).  Integer counts are order-independent, so N ranks == 1 rank bit for bit.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: inside a torch process that is the copy torch already loaded,
// in a plain C host the system one), so the library loads and every other entry point works on hosts without NCCL.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "lh_common.cuh"

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    bool ok = false;
};

const NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        bool all = true;
        auto sym = [&](const char* name) { void* f = dlsym(h, name); all = all && f != nullptr; return f; };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.CommCount = reinterpret_cast<decltype(api.CommCount)>(sym("ncclCommCount"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.ok = all;
    });
    return api;
}

}  // namespace

static_assert(LH_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "lh_comm id size follows ncclUniqueId");

extern "C" int lh_comm_unique_id(void* id_out) {
    if (!id_out) return LH_ENULL;
    const NcclApi& n = nccl();
    if (!n.ok) return LH_ENCCL;
    ncclUniqueId id;
    if (n.GetUniqueId(&id) != ncclSuccess) return LH_ENCCL;
    memcpy(id_out, &id, sizeof(id));
    return LH_OK;
}

extern "C" int lh_comm_init(const void* id, int world, int rank, int device, void** comm_out) {
    if (!id || !comm_out) return LH_ENULL;
    if (world < 1 || rank < 0 || rank >= world) return LH_EINVAL;
    const NcclApi& n = nccl();
    if (!n.ok) return LH_ENCCL;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t c = nullptr;
    if (n.CommInitRank(&c, world, uid, rank) != ncclSuccess) return LH_ENCCL;
    *comm_out = c;
    return LH_OK;
}

extern "C" int lh_comm_destroy(void* comm) {
    if (!comm) return LH_OK;
    const NcclApi& n = nccl();
    if (!n.ok) return LH_ENCCL;
    return n.CommDestroy(reinterpret_cast<ncclComm_t>(comm)) == ncclSuccess ? LH_OK : LH_ENCCL;
}

// One grouped NCCL operation (a single fused launch on `stream`) over n_bufs device buffers, in place.
extern "C" int lh_metrics_allreduce_multi(void* comm, int n_bufs, void* const* bufs, const size_t* elems, const int* is_f64,
                                          lh_stream_t stream) {
    if (n_bufs < 0) return LH_EINVAL;
    if (n_bufs == 0) return LH_OK;
    if (!comm || !bufs || !elems || !is_f64) return LH_ENULL;
    const NcclApi& n = nccl();
    if (!n.ok) return LH_ENCCL;
    ncclComm_t c = reinterpret_cast<ncclComm_t>(comm);
    for (int i = 0; i < n_bufs; ++i) if (!bufs[i] && elems[i]) return LH_ENULL;
    bool fail = n.GroupStart() != ncclSuccess;
    for (int i = 0; i < n_bufs && !fail; ++i) {
        if (!elems[i]) continue;
        fail = n.AllReduce(bufs[i], bufs[i], elems[i], is_f64[i] ? ncclFloat64 : ncclUint64, ncclSum, c,
                           (cudaStream_t)stream) != ncclSuccess;
    }
    fail = (n.GroupEnd() != ncclSuccess) || fail;
    return fail ? LH_ENCCL : LH_OK;
}

extern "C" int lh_metrics_allreduce(void* comm, unsigned long long* counts, int n_counts, double* sums, int n_sums,
                                    lh_stream_t stream) {
    if (n_counts < 0 || n_sums < 0) return LH_EINVAL;
    void* bufs[2] = {counts, sums};
    const size_t elems[2] = {(size_t)n_counts, (size_t)n_sums};
    const int f64[2] = {0, 1};
    return lh_metrics_allreduce_multi(comm, 2, bufs, elems, f64, stream);
}
