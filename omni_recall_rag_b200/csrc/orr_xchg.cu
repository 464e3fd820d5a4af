// orr_xchg.cu — exchange buffers of the fused all-gather + merge (multi-GPU).
//
// The reference is a single process with no collective; the row-sharded B200 design needs exactly one
// exchange step per query: k x 24 B of exact local hits per GPU.  Instead of an NCCL all-gather followed by a
// merge kernel, every rank pushes its list straight into its peers' HBM over NVLink/NVSwitch and merges what
// the peers pushed to it, in ONE one-CTA kernel (orr_xchg_merge_kernel, orr_rescore.cu).  This file owns the
// buffers: cudaMalloc'ed per rank, shared with the other ranks either through CUDA IPC handles (one process
// per GPU: torch.distributed carries the 64-byte handles) or directly (one host process driving all GPUs,
// the .NET deployment: cudaDeviceEnablePeerAccess).
#include <cstring>

#include "orr_internal.h"

struct orr_xchg {
    int device = 0, world = 1, rank = 0, kmax = 1;
    size_t slot_bytes = 0, bytes = 0;
    uint8_t* local = nullptr;
    uint8_t* peers[ORR_XCHG_MAX_WORLD] = {};
    bool ipc_opened[ORR_XCHG_MAX_WORLD] = {};
    uint32_t seq = 0;
    double timeout_s = 5.0;
    std::mutex mu;
};

extern "C" {

int orr_xchg_create(int32_t device, int32_t world, int32_t rank, int32_t max_top_k, orr_xchg** out) {
    if (!out || world < 1 || world > ORR_XCHG_MAX_WORLD || rank < 0 || rank >= world || max_top_k < 1 ||
        (int64_t)world * max_top_k > ORR_SORT_MAX) {
        orr_set_error("orr_xchg_create: bad argument (world %d, rank %d, max_top_k %d)", world, rank, max_top_k);
        return ORR_E_INVALID;
    }
    *out = nullptr;
    ORR_CUDA_OK(cudaSetDevice(device));
    std::unique_ptr<orr_xchg> x(new orr_xchg());
    x->device = device; x->world = world; x->rank = rank; x->kmax = max_top_k;
    // per source rank: max_top_k hits (6 words each) + 2 status words, every word an 8-byte {data, seq} LL store
    const size_t raw = (size_t)world * ((size_t)max_top_k * 6 + 2) * 8;
    x->slot_bytes = (raw + 127) / 128 * 128;
    x->bytes = x->slot_bytes * ORR_XCHG_SLOTS;
    ORR_CUDA_OK(cudaMalloc(&x->local, x->bytes));
    ORR_CUDA_OK(cudaMemset(x->local, 0, x->bytes));
    ORR_CUDA_OK(cudaDeviceSynchronize());          // zeroed before any peer can learn the address
    x->peers[rank] = x->local;
    *out = x.release();
    return ORR_OK;
}

void orr_xchg_destroy(orr_xchg* x) {
    if (!x) return;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < x->world; ++p)
        if (x->ipc_opened[p] && x->peers[p]) cudaIpcCloseMemHandle(x->peers[p]);
    cudaFree(x->local);
    delete x;
}

int orr_xchg_get_handle(orr_xchg* x, void* handle_out) {
    if (!x || !handle_out) { orr_set_error("orr_xchg_get_handle: NULL argument"); return ORR_E_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == ORR_XCHG_HANDLE_BYTES, "IPC handle size");
    ORR_CUDA_OK(cudaSetDevice(x->device));
    cudaIpcMemHandle_t h;
    ORR_CUDA_OK(cudaIpcGetMemHandle(&h, x->local));
    memcpy(handle_out, &h, sizeof h);
    return ORR_OK;
}

int orr_xchg_open_peer(orr_xchg* x, int32_t peer_rank, const void* handle) {
    if (!x || !handle || peer_rank < 0 || peer_rank >= x->world) { orr_set_error("orr_xchg_open_peer: bad argument"); return ORR_E_INVALID; }
    if (peer_rank == x->rank) return ORR_OK;
    ORR_CUDA_OK(cudaSetDevice(x->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void* p = nullptr;
    ORR_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    x->peers[peer_rank] = (uint8_t*)p;
    x->ipc_opened[peer_rank] = true;
    return ORR_OK;
}

int orr_xchg_attach_peer(orr_xchg* x, int32_t peer_rank, orr_xchg* peer) {
    if (!x || !peer || peer_rank < 0 || peer_rank >= x->world || peer->rank != peer_rank || peer->world != x->world ||
        peer->kmax != x->kmax) {
        orr_set_error("orr_xchg_attach_peer: bad argument");
        return ORR_E_INVALID;
    }
    if (peer->device != x->device) {
        ORR_CUDA_OK(cudaSetDevice(x->device));
        int can = 0;
        ORR_CUDA_OK(cudaDeviceCanAccessPeer(&can, x->device, peer->device));
        if (!can) { orr_set_error("device %d cannot access device %d", x->device, peer->device); return ORR_E_UNSUPPORTED; }
        cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ORR_CUDA_OK(e);
        cudaGetLastError();
    }
    x->peers[peer_rank] = peer->local;
    return ORR_OK;
}

// After a time-out (a rank issued fewer searches than its peers, or died) the ranks' sequence numbers are out of step and
// every later exchange would time out too.  Resynchronisation is an agreement made OUT OF BAND: every rank drains its
// stream, the ranks agree on a number larger than any sequence number in use (e.g. an all-reduce MAX of
// orr_xchg_sequence() plus ORR_XCHG_SLOTS), every rank calls orr_xchg_resync with it, and a barrier follows before the
// next exchange.  Stale words in the slots carry older tags and can never match.
int orr_xchg_resync(orr_xchg* x, uint32_t next_seq_base) {
    if (!x) { orr_set_error("orr_xchg_resync: NULL argument"); return ORR_E_INVALID; }
    std::lock_guard<std::mutex> g(x->mu);
    x->seq = next_seq_base;
    return ORR_OK;
}

uint32_t orr_xchg_sequence(orr_xchg* x) {
    if (!x) return 0;
    std::lock_guard<std::mutex> g(x->mu);
    return x->seq;
}

// How long the exchange kernel spins for a missing peer before it gives up and flags ORR_STATUS_XCHG_TIMEOUT
// (default 5 s; the GPU is occupied by one CTA meanwhile).
int orr_xchg_set_timeout_ms(orr_xchg* x, double ms) {
    if (!x || !(ms > 0.0) || ms > 600000.0) { orr_set_error("orr_xchg_set_timeout_ms: bad argument"); return ORR_E_INVALID; }
    std::lock_guard<std::mutex> g(x->mu);
    x->timeout_s = ms / 1000.0;
    return ORR_OK;
}

int orr_xchg_allgather_merge(orr_xchg* x, const orr_hit* hits_dev, const int32_t* status_dev, int32_t top_k,
                             orr_hit* out_dev, int32_t* out_status_dev, void* cuda_stream) {
    if (!x || !hits_dev || !status_dev || !out_dev || !out_status_dev) { orr_set_error("orr_xchg_allgather_merge: NULL argument"); return ORR_E_INVALID; }
    const int k = top_k < 1 ? 1 : top_k;
    if (k > x->kmax) { orr_set_error("orr_xchg_allgather_merge: top_k %d > max_top_k %d", k, x->kmax); return ORR_E_INVALID; }
    for (int p = 0; p < x->world; ++p)
        if (!x->peers[p]) { orr_set_error("orr_xchg_allgather_merge: peer %d was never opened/attached", p); return ORR_E_INVALID; }
    std::lock_guard<std::mutex> g(x->mu);
    ORR_CUDA_OK(cudaSetDevice(x->device));
    OrrXchgArgs a{};
    a.src_hits = hits_dev; a.src_status = status_dev;
    for (int p = 0; p < x->world; ++p) a.peer_base[p] = x->peers[p];
    a.slot_bytes = x->slot_bytes;
    a.world = x->world; a.rank = x->rank; a.kmax = x->kmax; a.top_k = top_k;
    a.seq = ++x->seq;
    if (a.seq == 0) a.seq = ++x->seq;              // 0 is the cleared state
    a.slot = (int)(a.seq % ORR_XCHG_SLOTS);
    a.timeout_ns = (unsigned long long)(x->timeout_s * 1e9);
    a.out = out_dev; a.out_status = out_status_dev;
    return orr_launch_xchg_merge(a, (cudaStream_t)cuda_stream);
}

}  // extern "C"
