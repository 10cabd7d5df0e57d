// Host side of the peer-memory exchange (peer.cuh): one PeerBlock per rank, shared with the other processes of the node
// through CUDA IPC, and the stand-alone all-reduce launch.  The evaluation pass fuses the same exchange into its last
// kernel (metrics.cu, polcue_eval_pass_peer_f32).
#include <cstring>

#include "peer.cuh"
#include "polcue_host.h"

static_assert(POLCUE_PEER_HANDLE_BYTES == sizeof(cudaIpcMemHandle_t), "the IPC handle travels as an opaque byte string");
static_assert(POLCUE_PEER_MAX_VALUES == polcue::kPeerMaxValues && POLCUE_PEER_MAX_RANKS == polcue::kPeerMaxRanks, "header constants");

namespace polcue {
namespace {

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_kernel(const __grid_constant__ PeerParams pp, const double* __restrict__ in,
                                                                      int n, double* __restrict__ out) {
    const int k = threadIdx.x;
    const double total = peer_allreduce_cta(pp, k < n ? in[k] : 0.0, n);
    if (k < n) out[k] = total;
}

}  // namespace

bool peer_params(const polcue_peer* peer, PeerParams& pp) {
    if (!peer || !peer->connected) return false;
    pp.world = peer->world;
    pp.rank = peer->rank;
    for (int r = 0; r < kPeerMaxRanks; ++r) pp.block[r] = r < peer->world ? static_cast<PeerBlock*>(peer->block[r]) : nullptr;
    return true;
}

}  // namespace polcue

using namespace polcue;

extern "C" {

int polcue_peer_create(int world, int rank, polcue_peer** out, void* ipc_handle) {
    if (!out) return POLCUE_EINVAL;
    *out = nullptr;
    if (world < 1 || world > kPeerMaxRanks || rank < 0 || rank >= world || (world > 1 && !ipc_handle)) return POLCUE_EINVAL;
    auto* peer = new polcue_peer();
    peer->world = world;
    peer->rank = rank;
    cudaError_t e = cudaGetDevice(&peer->device);
    if (e == cudaSuccess) e = cudaMalloc(&peer->block[rank], sizeof(PeerBlock));
    if (e == cudaSuccess) e = cudaMemset(peer->block[rank], 0, sizeof(PeerBlock));
    if (e == cudaSuccess && ipc_handle) {
        cudaIpcMemHandle_t h;
        e = cudaIpcGetMemHandle(&h, peer->block[rank]);
        if (e == cudaSuccess) std::memcpy(ipc_handle, &h, sizeof(h));
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();      // the zeroed block is in memory before any peer can reach it
    if (e != cudaSuccess) {
        if (peer->block[rank]) cudaFree(peer->block[rank]);
        delete peer;
        return e == cudaErrorMemoryAllocation ? POLCUE_ENOMEM : (int)e;
    }
    peer->connected = world == 1;
    *out = peer;
    return POLCUE_OK;
}

int polcue_peer_connect(polcue_peer* peer, const void* ipc_handles) {
    if (!peer || (!ipc_handles && peer->world > 1)) return POLCUE_EINVAL;
    if (peer->connected) return POLCUE_OK;
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev != peer->device) return POLCUE_EINVAL;
    for (int r = 0; r < peer->world; ++r) {
        if (r == peer->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const char*>(ipc_handles) + (size_t)r * sizeof(h), sizeof(h));
        e = cudaIpcOpenMemHandle(&peer->block[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int q = 0; q < r; ++q)
                if (q != peer->rank && peer->block[q]) {
                    cudaIpcCloseMemHandle(peer->block[q]);
                    peer->block[q] = nullptr;
                }
            cudaGetLastError();
            return (int)e;
        }
    }
    peer->connected = true;
    return POLCUE_OK;
}

int polcue_peer_allreduce_f64(polcue_peer* peer, const double* in, int n, double* out, polcue_stream_t stream) {
    PeerParams pp;
    if (!peer_params(peer, pp) || !in || !out || n < 1 || n > kPeerMaxValues) return POLCUE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 7) return POLCUE_EINVAL;
    peer_allreduce_kernel<<<1, kPeerThreads, 0, (cudaStream_t)stream>>>(pp, in, n, out);
    return launch_status();
}

int polcue_peer_status(polcue_peer* peer, unsigned long long* calls, unsigned long long* failed_call) {
    if (!peer || !peer->block[peer->rank]) return POLCUE_EINVAL;
    unsigned long long host[2] = {0, 0};
    const cudaError_t e = cudaMemcpy(host, peer->block[peer->rank], sizeof(host), cudaMemcpyDeviceToHost);   // synchronises
    if (e != cudaSuccess) return (int)e;
    if (calls) *calls = host[0];
    if (failed_call) *failed_call = host[1];
    return POLCUE_OK;
}

int polcue_peer_destroy(polcue_peer* peer) {
    if (!peer) return POLCUE_OK;
    cudaError_t first = cudaSuccess;
    for (int r = 0; r < peer->world; ++r) {
        if (!peer->block[r]) continue;
        const cudaError_t e = r == peer->rank ? cudaFree(peer->block[r]) : cudaIpcCloseMemHandle(peer->block[r]);
        if (e != cudaSuccess && first == cudaSuccess) first = e;
    }
    delete peer;
    return (int)first;
}

}  // extern "C"
