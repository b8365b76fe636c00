// bt_comm.cu — the multi-GPU communicator of the sharded preprocessor (C ABI: bt_comm_*).
//
// One process per GPU; the ranks of one box exchange the preprocessor's sub-root tiles and border strips over
// NVLink / NVSwitch.  Two transports behind one handle:
//   * NCCL (ncclAllGather, grouped) — resolved with dlopen at bt_comm_create, so single-GPU hosts never load it and a
//     torch process keeps the libnccl.so.2 it has already mapped (same SONAME -> same object);
//   * peer memory — every rank publishes the CUDA IPC handle of its pyramid arena and exchange buffers, the handles are
//     all-gathered once through NCCL, and from then on kernels dereference peer pointers directly (ld.global over NVLink):
//     the stitch kernel reads its neighbours' border strips out of the owning GPU's HBM with no pack / collective / unpack.
// The reference is single-GPU (preprocess/mod.rs:143-217); this file has no counterpart there.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types and enums only: every function is resolved at run time

#include <cstdio>
#include <cstring>
#include <vector>

#include "bt_comm.h"
#include "bt_runtime.h"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi g_nccl;
char g_nccl_error[256] = "";

bool load_nccl() {
    if (g_nccl.handle) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { std::snprintf(g_nccl_error, sizeof(g_nccl_error), "dlopen(libnccl.so.2): %s", dlerror()); return false; }
    NcclApi a;
    a.handle = h;
#define BT_SYM(field, name) \
    *reinterpret_cast<void**>(&a.field) = dlsym(h, name); \
    if (!a.field) { std::snprintf(g_nccl_error, sizeof(g_nccl_error), "libnccl: symbol %s missing", name); return false; }
    BT_SYM(GetUniqueId, "ncclGetUniqueId")
    BT_SYM(CommInitRank, "ncclCommInitRank")
    BT_SYM(CommDestroy, "ncclCommDestroy")
    BT_SYM(AllGather, "ncclAllGather")
    BT_SYM(AllReduce, "ncclAllReduce")
    BT_SYM(GroupStart, "ncclGroupStart")
    BT_SYM(GroupEnd, "ncclGroupEnd")
    BT_SYM(GetErrorString, "ncclGetErrorString")
    BT_SYM(GetVersion, "ncclGetVersion")
#undef BT_SYM
    g_nccl = a;
    return true;
}

}  // namespace

#define BT_NCCL(comm_, expr)                                                                                    \
    do {                                                                                                        \
        ncclResult_t r_ = (expr);                                                                               \
        if (r_ != ncclSuccess) {                                                                                \
            (comm_)->ctx->set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(r_)); \
            return BT_ERR_NCCL;                                                                                 \
        }                                                                                                       \
    } while (0)

extern "C" {

bt_status bt_comm_unique_id(void* out, size_t bytes) {
    if (!out || bytes < BT_COMM_UNIQUE_ID_BYTES) return BT_ERR_INVALID_ARG;
    static_assert(sizeof(ncclUniqueId) == BT_COMM_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
    if (!load_nccl()) return BT_ERR_NCCL;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return BT_ERR_NCCL;
    std::memcpy(out, &id, sizeof(id));
    return BT_OK;
}

const char* bt_comm_load_error(void) { return g_nccl_error; }

bt_status bt_comm_create(bt_context* ctx, uint32_t rank, uint32_t world, const void* unique_id, size_t id_bytes, bt_comm** out) {
    if (!ctx || !out || !unique_id || id_bytes < BT_COMM_UNIQUE_ID_BYTES || world == 0 || rank >= world || world > BT_COMM_MAX_RANKS) return BT_ERR_INVALID_ARG;
    *out = nullptr;
    if (!load_nccl()) { ctx->set_error("bt_comm_create: %s", g_nccl_error); return BT_ERR_NCCL; }
    BT_CUDA(ctx, cudaSetDevice(ctx->device));
    bt_comm* c = new bt_comm();
    c->ctx = ctx;
    c->rank = rank;
    c->world = world;
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof(id));
    ncclComm_t nc = nullptr;
    ncclResult_t r = g_nccl.CommInitRank(&nc, (int)world, id, (int)rank);
    if (r != ncclSuccess) { ctx->set_error("bt_comm_create: ncclCommInitRank: %s", g_nccl.GetErrorString(r)); delete c; return BT_ERR_NCCL; }
    c->nccl = nc;
    g_nccl.GetVersion(&c->nccl_version);
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_scratch, 4096);
    if (e == cudaSuccess) e = cudaMemset(c->d_scratch, 0, 4096);
    if (e != cudaSuccess) { ctx->set_error("bt_comm_create: %s", cudaGetErrorString(e)); bt_comm_destroy(c); return BT_ERR_CUDA; }
    *out = c;
    return BT_OK;
}

void bt_comm_destroy(bt_comm* c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    cudaDeviceSynchronize();
    for (auto& m : c->peer_maps)
        for (uint32_t r = 0; r < c->world; ++r)
            if (r != c->rank && m.ptr[r]) cudaIpcCloseMemHandle(m.ptr[r]);
    if (c->nccl) g_nccl.CommDestroy((ncclComm_t)c->nccl);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->ready) cudaEventDestroy(c->ready);
    if (c->done) cudaEventDestroy(c->done);
    cudaFree(c->d_scratch);
    delete c;
}

bt_status bt_comm_info(const bt_comm* c, uint32_t* rank, uint32_t* world, int32_t* nccl_version, uint32_t* peer_access) {
    if (!c) return BT_ERR_INVALID_ARG;
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    if (nccl_version) *nccl_version = c->nccl_version;
    if (peer_access) *peer_access = c->peer_ok ? 1u : 0u;
    return BT_OK;
}

// stream-ordered barrier over all ranks (a one-word all-reduce): everything enqueued on `stream` before it, on every
// rank, has completed when work enqueued after it starts
bt_status bt_comm_barrier(bt_comm* c, void* stream) {
    if (!c) return BT_ERR_INVALID_ARG;
    BT_CUDA(c->ctx, cudaSetDevice(c->ctx->device));
    BT_NCCL(c, g_nccl.AllReduce(c->d_scratch, c->d_scratch, 1, ncclUint32, ncclSum, (ncclComm_t)c->nccl, (cudaStream_t)stream));
    return BT_OK;
}

bt_status bt_comm_all_reduce_sum_u64(bt_comm* c, uint64_t* values, uint32_t count, void* stream) {
    if (!c || !values || count == 0 || count > 64) return BT_ERR_INVALID_ARG;   // 512 B of the scratch page
    BT_CUDA(c->ctx, cudaSetDevice(c->ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    BT_CUDA(c->ctx, cudaMemcpyAsync(c->d_scratch + 256, values, count * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    BT_NCCL(c, g_nccl.AllReduce(c->d_scratch + 256, c->d_scratch + 256, count, ncclUint64, ncclSum, (ncclComm_t)c->nccl, s));
    BT_CUDA(c->ctx, cudaMemcpyAsync(values, c->d_scratch + 256, count * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    BT_CUDA(c->ctx, cudaStreamSynchronize(s));
    return BT_OK;
}

}  // extern "C"

// ---- internal interface used by bt_preprocess.cu ------------------------------------------------------------------------------
bt_status bt_comm_all_gather_group(bt_comm* c, const bt_comm_segment* segs, uint32_t n, cudaStream_t stream) {
    if (n == 0) return BT_OK;
    BT_NCCL(c, g_nccl.GroupStart());
    for (uint32_t i = 0; i < n; ++i) {
        ncclResult_t r = g_nccl.AllGather(segs[i].send, segs[i].recv, segs[i].bytes_per_rank, ncclUint8, (ncclComm_t)c->nccl, stream);
        if (r != ncclSuccess) { g_nccl.GroupEnd(); c->ctx->set_error("ncclAllGather: %s", g_nccl.GetErrorString(r)); return BT_ERR_NCCL; }
    }
    BT_NCCL(c, g_nccl.GroupEnd());
    return BT_OK;
}

// Publishes `local` (a cudaMalloc allocation base) to every rank and maps every peer's allocation into this process.
// Collective: all ranks call it in the same order.  out[r] = pointer usable in kernels on THIS device (out[rank] = local).
bt_status bt_comm_map_peers(bt_comm* c, void* local, void** out, cudaStream_t stream) {
    bt_context* ctx = c->ctx;
    for (const auto& m : c->peer_maps)
        if (m.ptr[c->rank] == local) { for (uint32_t r = 0; r < c->world; ++r) out[r] = m.ptr[r]; return BT_OK; }
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof(mine));
    cudaError_t e = cudaIpcGetMemHandle(&mine, local);
    uint32_t ok = e == cudaSuccess ? 1u : 0u;
    if (e != cudaSuccess) (void)cudaGetLastError();
    // slot layout in d_scratch: [1024 .. 1024 + world * 72): {handle (64 B), ok (4 B), pad}
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    unsigned char rec[72];
    std::memset(rec, 0, sizeof(rec));
    std::memcpy(rec, &mine, 64);
    std::memcpy(rec + 64, &ok, 4);
    unsigned char* d_send = c->d_scratch + 1024;
    unsigned char* d_recv = c->d_scratch + 1024 + 72;
    static_assert(1024 + 72 + BT_COMM_MAX_RANKS * 72 <= 4096, "scratch size");
    BT_CUDA(ctx, cudaMemcpyAsync(d_send, rec, sizeof(rec), cudaMemcpyHostToDevice, stream));
    BT_NCCL(c, g_nccl.AllGather(d_send, d_recv, sizeof(rec), ncclUint8, (ncclComm_t)c->nccl, stream));
    std::vector<unsigned char> all((size_t)c->world * 72);
    BT_CUDA(ctx, cudaMemcpyAsync(all.data(), d_recv, all.size(), cudaMemcpyDeviceToHost, stream));
    BT_CUDA(ctx, cudaStreamSynchronize(stream));
    bool all_ok = true;
    for (uint32_t r = 0; r < c->world; ++r) { uint32_t k; std::memcpy(&k, &all[(size_t)r * 72 + 64], 4); all_ok = all_ok && k; }
    bt_comm::PeerMap m;
    if (all_ok) {
        for (uint32_t r = 0; r < c->world && all_ok; ++r) {
            if (r == c->rank) { m.ptr[r] = local; continue; }
            cudaIpcMemHandle_t h;
            std::memcpy(&h, &all[(size_t)r * 72], 64);
            void* p = nullptr;
            e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { (void)cudaGetLastError(); all_ok = false; ctx->set_error("cudaIpcOpenMemHandle(rank %u): %s", r, cudaGetErrorString(e)); }
            m.ptr[r] = p;
        }
    }
    // every rank must agree on whether peer access is usable: one more tiny reduction (min)
    uint32_t flag = all_ok ? 1u : 0u;
    BT_CUDA(ctx, cudaMemcpyAsync(c->d_scratch, &flag, 4, cudaMemcpyHostToDevice, stream));
    BT_NCCL(c, g_nccl.AllReduce(c->d_scratch, c->d_scratch, 1, ncclUint32, ncclMin, (ncclComm_t)c->nccl, stream));
    BT_CUDA(ctx, cudaMemcpyAsync(&flag, c->d_scratch, 4, cudaMemcpyDeviceToHost, stream));
    BT_CUDA(ctx, cudaStreamSynchronize(stream));
    uint32_t zero = 0;
    BT_CUDA(ctx, cudaMemcpyAsync(c->d_scratch, &zero, 4, cudaMemcpyHostToDevice, stream));   // the barrier word stays 0
    if (!flag) {
        for (uint32_t r = 0; r < c->world; ++r) if (r != c->rank && m.ptr[r]) cudaIpcCloseMemHandle(m.ptr[r]);
        c->peer_ok = false;
        return BT_ERR_NOT_FOUND;   // peer mapping unavailable on this box: the caller falls back to the NCCL transport
    }
    c->peer_ok = true;
    c->peer_maps.push_back(m);
    for (uint32_t r = 0; r < c->world; ++r) out[r] = m.ptr[r];
    return BT_OK;
}
