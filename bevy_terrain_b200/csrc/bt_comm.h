// bt_comm.h — internal definition of the multi-GPU communicator handle (bt_comm.cu) shared with bt_preprocess.cu.
#pragma once

#include <cuda_runtime.h>

#include <vector>

#include "../../include/bevy_terrain_b200.h"

struct bt_comm {
    bt_context* ctx = nullptr;
    uint32_t rank = 0, world = 1;
    void* nccl = nullptr;            // ncclComm_t
    int nccl_version = 0;
    cudaStream_t stream = nullptr;   // the exchange runs here, beside the caller's compute stream
    cudaEvent_t ready = nullptr, done = nullptr;
    unsigned char* d_scratch = nullptr;   // 4 KiB: barrier word, small reductions, IPC handle exchange
    bool peer_ok = false;            // peer (CUDA IPC over NVLink) mappings are usable on this box
    struct PeerMap { void* ptr[BT_COMM_MAX_RANKS] = {}; };
    std::vector<PeerMap> peer_maps;  // one per published allocation
};

struct bt_comm_segment { const void* send; void* recv; size_t bytes_per_rank; };

// grouped all-gathers (one NCCL group = one fused launch) on `stream`
bt_status bt_comm_all_gather_group(bt_comm* c, const bt_comm_segment* segs, uint32_t n, cudaStream_t stream);
// collective: publish a local cudaMalloc allocation and map every peer's (BT_ERR_NOT_FOUND = peer access unavailable)
bt_status bt_comm_map_peers(bt_comm* c, void* local, void** out, cudaStream_t stream);
