// Host-side internals shared by the translation units of libpolcue.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <vector>

#include "../../include/polcue.h"

struct polcue_lut {
    double n = 0.0;
    int cells[3] = {0, 0, 0};      // cell count per table (diffuse, spec1, spec2)
    int offset[3] = {0, 0, 0};     // first cell of each table inside the blob (float4 units)
    float scale[3] = {0, 0, 0};    // cells per unit g
    std::vector<float4> blob;      // host copy of the device blob
    std::vector<double> kx[3], ky[3];  // sorted knots as scipy's interp1d holds them
    // Steep end segments (|slope| > 512, e.g. -10797 for the second specular branch at n = 1.8, where a knot lands
    // within 1.5e-7 of the peak): float32 cannot hold rho or theta to 1e-3 rad there, so queries beyond the
    // second-to-last knot are evaluated in float64 on the device (polcue::steep_sincos).
    int steep_mask = 0;                 // bit t: table t has a steep end segment
    double steep_x[3] = {0, 0, 0};      // second-to-last knot (the anchor scipy's _call_linear uses)
    double steep_y[3] = {0, 0, 0};
    double steep_slope[3] = {0, 0, 0};
    float4* d_blob = nullptr;      // device copy (null for host-only builds)
    int trig_mufu = 1;             // zenith-angle sincos of launches that use this handle: 1 = MUFU sin/cos (3.6e-7 abs), 0 = polynomial (1.4e-7)
    int device = -1;
    size_t bytes() const { return blob.size() * sizeof(float4); }
};

struct polcue_peer {
    int world = 1, rank = 0, device = -1;
    bool connected = false;
    void* block[POLCUE_PEER_MAX_RANKS] = {};   // block[r]: rank r's PeerBlock (peer.cuh) as mapped here; block[rank] is this rank's own
};

namespace polcue {

extern std::atomic<unsigned long long> g_launches;

struct PeerParams;                                            // peer.cuh
bool peer_params(const polcue_peer* peer, PeerParams& pp);    // peer.cu: false unless the peer is connected

inline int launch_status() {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

struct DeviceInfo {
    int sms = 0;
    int smem_optin = 0;
};
const DeviceInfo& device_info();  // of the current device (cached per device)

// fused.cu: XOLP + normals from planes laid out B x 4 x H x W (used by the loader front end, resize.cu)
int fused_planes_strided(const uint8_t* planes, int B, int H, int W, const polcue_lut* lut, float* iun, float* xolp,
                         float* normals, cudaStream_t stream, float* xolp_norm = nullptr, float norm_mean = 0.0f,
                         float norm_std = 1.0f);

// stats.cu: canonical float64 fold of per-tile float records (polcue_device.cuh, "Deterministic sums").  The workspace
// holds a ticket word (first 8 bytes, zero before the first use; the kernel re-zeroes it), the segment partials and the
// tile records [n_tiles][stride]; out receives n_values (<= stride, <= 32) doubles.
size_t fold_workspace_bytes(uint32_t n_tiles, int stride);
float* fold_records(void* workspace, uint32_t n_tiles);
int launch_fold_tiles(void* workspace, uint32_t n_tiles, int stride, int n_values, double* out, cudaStream_t stream);

// magic numbers for polcue::FastDiv (exact for numerators < 2^31)
inline void make_fastdiv(uint32_t d, uint32_t& mul, uint32_t& shift) {
    if (d <= 1) {
        mul = 0;
        shift = 0;
        return;
    }
    uint32_t s = 0;
    while ((1ull << s) < d) ++s;  // s = ceil(log2 d) >= 1
    mul = (uint32_t)(((1ull << (31 + s)) / d) + 1);
    shift = s - 1;
}

}  // namespace polcue
