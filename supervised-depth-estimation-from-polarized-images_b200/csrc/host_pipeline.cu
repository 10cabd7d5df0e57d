// Host-buffer entry points of the fused pipeline: the calls a user of the Python/C API makes when the
// mosaics (or the loader's stored images) live in host memory.  Chunks flow through a 3-slot device ring on
// three streams so the host->device copy of chunk i+1, the kernels on chunk i and the device->host copy of
// chunk i-1 overlap (PCIe is full duplex).  bench.py's `e2e` figure times polcue_fused_mosaic_u8_host.
//
// Rings (device buffers, streams, events, the resize plan of the front end) are cached per
// (device, geometry, output set) in a small LRU.  The cache map is guarded by one short-lived mutex; each ring
// has its OWN mutex held for the duration of a call, so calls for different devices or geometries run
// concurrently and two threads asking for the same ring serialise on it (include/polcue.h states this).
// Every CUDA return value is checked; after a failure the ring's streams are drained before returning so no
// copy is still writing into the caller's buffers.
//
// polcue_host_alloc: pinned host memory placed on the NUMA node of the GPU that will DMA into it (the
// device->host stream of 44 B per pixel is what bounds the host entry point; on a two-socket 8-GPU box a
// buffer on the far socket crosses the inter-socket link).
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <mutex>

#include "polcue_host.h"

namespace {

constexpr int kSlots = 3;
constexpr size_t kMaxRings = 4;           // LRU capacity per ring kind

#define PC_CHECK(expr)                          \
    do {                                        \
        const cudaError_t e_ = (expr);          \
        if (e_ != cudaSuccess && err == cudaSuccess) err = e_; \
    } while (0)

int as_code(cudaError_t e) { return e == cudaSuccess ? POLCUE_OK : (e == cudaErrorMemoryAllocation ? POLCUE_ENOMEM : (int)e); }

// Streams and events shared by both ring kinds.
struct Lanes {
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    cudaEvent_t in_done[kSlots] = {}, run_done[kSlots] = {}, out_done[kSlots] = {};

    cudaError_t create() {
        cudaError_t err = cudaSuccess;
        PC_CHECK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
        PC_CHECK(cudaStreamCreateWithFlags(&s_run, cudaStreamNonBlocking));
        PC_CHECK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
        for (int i = 0; i < kSlots; ++i) {
            PC_CHECK(cudaEventCreateWithFlags(&in_done[i], cudaEventDisableTiming));
            PC_CHECK(cudaEventCreateWithFlags(&run_done[i], cudaEventDisableTiming));
            PC_CHECK(cudaEventCreateWithFlags(&out_done[i], cudaEventDisableTiming));
        }
        return err;
    }
    void destroy() {
        for (int i = 0; i < kSlots; ++i) {
            if (in_done[i]) cudaEventDestroy(in_done[i]);
            if (run_done[i]) cudaEventDestroy(run_done[i]);
            if (out_done[i]) cudaEventDestroy(out_done[i]);
            in_done[i] = run_done[i] = out_done[i] = nullptr;
        }
        if (s_in) cudaStreamDestroy(s_in);
        if (s_run) cudaStreamDestroy(s_run);
        if (s_out) cudaStreamDestroy(s_out);
        s_in = s_run = s_out = nullptr;
    }
    // Wait for everything queued so far; returns the first error seen.
    cudaError_t drain() {
        cudaError_t err = cudaSuccess;
        PC_CHECK(cudaStreamSynchronize(s_out));
        PC_CHECK(cudaStreamSynchronize(s_run));
        PC_CHECK(cudaStreamSynchronize(s_in));
        return err;
    }
};

struct MosaicKey {
    int device, chunk, H, W, iun, normals;
    bool operator<(const MosaicKey& o) const {
        return std::memcmp(this, &o, sizeof(MosaicKey)) < 0;
    }
};

struct MosaicRing {
    std::mutex busy;
    Lanes lanes;
    unsigned char* d_mosaic[kSlots] = {};
    float* d_iun[kSlots] = {};
    float* d_xolp[kSlots] = {};
    float* d_normals[kSlots] = {};

    int create(const MosaicKey& k) {
        const size_t px = (size_t)(k.H / 2) * (k.W / 2);
        cudaError_t err = lanes.create();
        for (int i = 0; i < kSlots && err == cudaSuccess; ++i) {
            PC_CHECK(cudaMalloc(&d_mosaic[i], (size_t)k.chunk * k.H * k.W));
            PC_CHECK(cudaMalloc(&d_xolp[i], k.chunk * 2 * px * sizeof(float)));
            if (k.iun) PC_CHECK(cudaMalloc(&d_iun[i], k.chunk * px * sizeof(float)));
            if (k.normals) PC_CHECK(cudaMalloc(&d_normals[i], k.chunk * 9 * px * sizeof(float)));
        }
        return as_code(err);
    }
    ~MosaicRing() {
        for (int i = 0; i < kSlots; ++i) {
            cudaFree(d_mosaic[i]);
            cudaFree(d_iun[i]);
            cudaFree(d_xolp[i]);
            cudaFree(d_normals[i]);
        }
        lanes.destroy();
    }
};

struct FrontKey {
    int device, chunk, in_h, in_w, out_h, out_w, normals, xnorm, planes;
    bool operator<(const FrontKey& o) const { return std::memcmp(this, &o, sizeof(FrontKey)) < 0; }
};

// The same three-stream ring for the loader front end: four full-resolution image stacks (+ the flip flags of the
// chunk) in; resized planes, XOLP, normalised XOLP and normals out.  The resize plan (device weights) lives with it.
struct FrontRing {
    std::mutex busy;
    Lanes lanes;
    polcue_resize_plan* plan = nullptr;
    unsigned char* d_in[kSlots] = {};        // 4 x chunk x in_h x in_w
    unsigned char* d_flip[kSlots] = {};      // chunk flags
    unsigned char* d_ws[kSlots] = {};
    unsigned char* d_planes[kSlots] = {};
    float* d_xolp[kSlots] = {};
    float* d_xnorm[kSlots] = {};
    float* d_normals[kSlots] = {};

    int create(const FrontKey& k) {
        const int rc = polcue_resize_plan_create(k.in_h, k.in_w, k.out_h, k.out_w, &plan);
        if (rc != POLCUE_OK) return rc;
        const size_t opx = (size_t)k.out_h * k.out_w;
        cudaError_t err = lanes.create();
        for (int i = 0; i < kSlots && err == cudaSuccess; ++i) {
            PC_CHECK(cudaMalloc(&d_in[i], (size_t)4 * k.chunk * k.in_h * k.in_w));
            PC_CHECK(cudaMalloc(&d_flip[i], (size_t)k.chunk));
            PC_CHECK(cudaMalloc(&d_ws[i], polcue_resize_workspace_bytes(plan, 4 * k.chunk)));
            PC_CHECK(cudaMalloc(&d_planes[i], (size_t)k.chunk * 4 * opx));
            PC_CHECK(cudaMalloc(&d_xolp[i], k.chunk * 2 * opx * sizeof(float)));
            if (k.xnorm) PC_CHECK(cudaMalloc(&d_xnorm[i], k.chunk * 2 * opx * sizeof(float)));
            if (k.normals) PC_CHECK(cudaMalloc(&d_normals[i], k.chunk * 9 * opx * sizeof(float)));
        }
        return as_code(err);
    }
    ~FrontRing() {
        for (int i = 0; i < kSlots; ++i) {
            cudaFree(d_in[i]); cudaFree(d_flip[i]); cudaFree(d_ws[i]); cudaFree(d_planes[i]);
            cudaFree(d_xolp[i]); cudaFree(d_xnorm[i]); cudaFree(d_normals[i]);
        }
        lanes.destroy();
        if (plan) polcue_resize_plan_destroy(plan);
    }
};

// Least-recently-used cache of rings.  `acquire` returns a shared_ptr so an entry evicted by another thread stays
// alive until its current user is done; device memory of an evicted ring is released by the last owner.
template <typename Key, typename Ring>
class RingCache {
public:
    int acquire(const Key& key, std::shared_ptr<Ring>& out) {
        std::lock_guard<std::mutex> guard(mu_);
        for (auto it = order_.begin(); it != order_.end(); ++it) {
            if (!(it->first < key) && !(key < it->first)) {
                order_.splice(order_.begin(), order_, it);       // most recently used first
                out = it->second;
                return POLCUE_OK;
            }
        }
        while (order_.size() >= kMaxRings) order_.pop_back();    // drop before allocating: the device may be nearly full
        auto ring = std::make_shared<Ring>();
        const int rc = ring->create(key);
        if (rc != POLCUE_OK) return rc;
        order_.emplace_front(key, ring);
        out = ring;
        return POLCUE_OK;
    }

private:
    std::mutex mu_;
    std::list<std::pair<Key, std::shared_ptr<Ring>>> order_;
};

RingCache<MosaicKey, MosaicRing> g_mosaic_rings;
RingCache<FrontKey, FrontRing> g_front_rings;

// ---- NUMA-local pinned memory ------------------------------------------------------------------------------------
struct HostBlock {
    size_t bytes;
    bool mapped;       // true: mmap + cudaHostRegister, false: cudaHostAlloc
};
std::mutex g_host_mutex;
std::map<void*, HostBlock> g_host_blocks;

int gpu_numa_node(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), device) != cudaSuccess) return -1;
    for (char* c = bus; *c; ++c) *c = (char)std::tolower((unsigned char)*c);
    char path[96];
    std::snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = std::fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (std::fscanf(f, "%d", &node) != 1) node = -1;
    std::fclose(f);
    return node;
}

}  // namespace

extern "C" {

int polcue_host_numa_node(int device) {
    int dev = device;
    if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) return -1;
    return gpu_numa_node(dev);
}

int polcue_host_alloc_on(void** ptr, size_t bytes, int device) {
    if (!ptr || bytes == 0) return POLCUE_EINVAL;
    *ptr = nullptr;
    int dev = device;
    if (dev < 0) {
        const cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return (int)e;
    }
    const int node = gpu_numa_node(dev);
    const bool bind = node >= 0 && node < 1024;
    const char* mode = std::getenv("POLCUE_HOST_ALLOC");       // A/B probing only: "mapped" = huge-page mapping + cudaHostRegister
    if (mode && std::strcmp(mode, "mapped") == 0) {
        // Anonymous mapping backed by transparent huge pages, bound to the GPU's node, populated, then registered with the
        // driver.  Measured on B200 / PCIe 5 (profiles/pcie_probe_r02*.json, bench e2e): device->host copies into such a
        // registered mapping reach 39 GB/s where driver-allocated pinned memory reaches 57 GB/s, so this is NOT the default.
        const size_t page = 2u << 20;
        const size_t len = (bytes + page - 1) / page * page;
        void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p != MAP_FAILED) {
            madvise(p, len, MADV_HUGEPAGE);
            if (bind) {
                unsigned long mask[16] = {0};
                mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
                (void)syscall(SYS_mbind, p, len, 1, mask, (unsigned long)(8 * sizeof(mask)), 0u);
            }
            std::memset(p, 0, len);
            if (cudaHostRegister(p, len, cudaHostRegisterPortable) == cudaSuccess) {
                std::lock_guard<std::mutex> guard(g_host_mutex);
                g_host_blocks[p] = HostBlock{len, true};
                *ptr = p;
                return POLCUE_OK;
            }
            (void)cudaGetLastError();
            munmap(p, len);
        }
    }
    // Driver-allocated pinned memory; while it is allocated (and first touched) the calling thread PREFERS the GPU's NUMA
    // node (set_mempolicy MPOL_PREFERRED = 1), so on a two-socket box the pages land next to the GPU that will DMA into
    // them.  A refused syscall (seccomp, single-node VM) changes nothing.
    unsigned long mask[16] = {0}, saved_mask[16] = {0};
    int saved_mode = 0;
    bool switched = false;
    if (bind && syscall(SYS_get_mempolicy, &saved_mode, saved_mask, (unsigned long)(8 * sizeof(saved_mask)), nullptr, 0ul) == 0) {
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
        switched = syscall(SYS_set_mempolicy, 1, mask, (unsigned long)(8 * sizeof(mask))) == 0;
    }
    const cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocPortable);
    if (e == cudaSuccess) std::memset(*ptr, 0, bytes);
    if (switched)      // the caller's own policy comes back, whatever it was
        (void)syscall(SYS_set_mempolicy, saved_mode, saved_mode == 0 ? nullptr : saved_mask, saved_mode == 0 ? 0ul : (unsigned long)(8 * sizeof(saved_mask)));
    if (e != cudaSuccess) return as_code(e);
    std::lock_guard<std::mutex> guard(g_host_mutex);
    g_host_blocks[*ptr] = HostBlock{bytes, false};
    return POLCUE_OK;
}

int polcue_host_alloc(void** ptr, size_t bytes) { return polcue_host_alloc_on(ptr, bytes, -1); }

int polcue_host_free(void* ptr) {
    if (!ptr) return POLCUE_OK;
    HostBlock blk{0, false};
    {
        std::lock_guard<std::mutex> guard(g_host_mutex);
        auto it = g_host_blocks.find(ptr);
        if (it == g_host_blocks.end()) return POLCUE_EINVAL;
        blk = it->second;
        g_host_blocks.erase(it);
    }
    if (blk.mapped) {
        const cudaError_t e = cudaHostUnregister(ptr);
        munmap(ptr, blk.bytes);
        return e == cudaSuccess ? POLCUE_OK : (int)e;
    }
    const cudaError_t e = cudaFreeHost(ptr);
    return e == cudaSuccess ? POLCUE_OK : (int)e;
}

int polcue_fused_mosaic_u8_host(const uint8_t* h_mosaic, int B, int H, int W, const polcue_lut* lut, float* h_iun,
                                float* h_xolp, float* h_normals, int chunk_frames) {
    if (!h_mosaic || !h_xolp || B < 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1)) return POLCUE_EINVAL;
    if (h_normals && (!lut || !lut->d_blob)) return POLCUE_EINVAL;
    if (B == 0) return POLCUE_OK;
    if (chunk_frames <= 0) {
        // ~64 MB of output per chunk: large enough to amortise launches, small enough to pipeline
        const size_t per_frame = (size_t)(H / 2) * (W / 2) * 48;
        chunk_frames = (int)((64u << 20) / (per_frame ? per_frame : 1));
        chunk_frames = chunk_frames < 1 ? 1 : chunk_frames;
    }
    if (chunk_frames > B) chunk_frames = B;

    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return (int)err;
    if (lut && lut->device >= 0 && lut->device != dev) return POLCUE_EINVAL;   // tables live on another device
    std::shared_ptr<MosaicRing> ring;
    const MosaicKey key{dev, chunk_frames, H, W, h_iun ? 1 : 0, h_normals ? 1 : 0};
    int rc = g_mosaic_rings.acquire(key, ring);
    if (rc != POLCUE_OK) return rc;
    std::lock_guard<std::mutex> busy(ring->busy);
    MosaicRing& r = *ring;
    Lanes& q = r.lanes;
    const size_t px = (size_t)(H / 2) * (W / 2), frame = (size_t)H * W;

    int slot = 0;
    for (int first = 0, it = 0; first < B && rc == POLCUE_OK && err == cudaSuccess; first += chunk_frames, ++it, slot = (slot + 1) % kSlots) {
        const int nb = (B - first < chunk_frames) ? B - first : chunk_frames;
        if (it >= kSlots) PC_CHECK(cudaStreamWaitEvent(q.s_in, q.out_done[slot], 0));   // slot's previous results are on the host
        PC_CHECK(cudaMemcpyAsync(r.d_mosaic[slot], h_mosaic + (size_t)first * frame, (size_t)nb * frame, cudaMemcpyHostToDevice, q.s_in));
        PC_CHECK(cudaEventRecord(q.in_done[slot], q.s_in));
        PC_CHECK(cudaStreamWaitEvent(q.s_run, q.in_done[slot], 0));
        if (err != cudaSuccess) break;
        rc = polcue_fused_mosaic_u8(r.d_mosaic[slot], nb, H, W, lut, nullptr, r.d_iun[slot], r.d_xolp[slot], r.d_normals[slot], q.s_run);
        if (rc != POLCUE_OK) break;
        PC_CHECK(cudaEventRecord(q.run_done[slot], q.s_run));
        PC_CHECK(cudaStreamWaitEvent(q.s_out, q.run_done[slot], 0));
        PC_CHECK(cudaMemcpyAsync(h_xolp + (size_t)first * 2 * px, r.d_xolp[slot], (size_t)nb * 2 * px * sizeof(float),
                                 cudaMemcpyDeviceToHost, q.s_out));
        if (h_iun)
            PC_CHECK(cudaMemcpyAsync(h_iun + (size_t)first * px, r.d_iun[slot], (size_t)nb * px * sizeof(float), cudaMemcpyDeviceToHost,
                                     q.s_out));
        if (h_normals)
            PC_CHECK(cudaMemcpyAsync(h_normals + (size_t)first * 9 * px, r.d_normals[slot], (size_t)nb * 9 * px * sizeof(float),
                                     cudaMemcpyDeviceToHost, q.s_out));
        PC_CHECK(cudaEventRecord(q.out_done[slot], q.s_out));
    }
    const cudaError_t drained = q.drain();      // also after a failure: nothing may still be writing into the caller's buffers
    if (rc != POLCUE_OK) return rc;
    if (err != cudaSuccess) return (int)err;
    return drained == cudaSuccess ? POLCUE_OK : (int)drained;
}

int polcue_loader_front_end_u8_host(int in_h, int in_w, int out_h, int out_w, const uint8_t* h_i0, const uint8_t* h_i45,
                                    const uint8_t* h_i90, const uint8_t* h_i135, int B, const uint8_t* h_flip, const polcue_lut* lut,
                                    uint8_t* h_planes, float* h_xolp, float* h_normals, const float* xolp_mean_std, float* h_xolp_norm,
                                    int chunk_samples) {
    if (!h_i0 || !h_i45 || !h_i90 || !h_i135 || !h_xolp || B < 0 || (h_xolp_norm && !xolp_mean_std)) return POLCUE_EINVAL;
    if (in_h <= 0 || in_w <= 0 || out_h <= 0 || out_w <= 0) return POLCUE_EINVAL;
    if (h_normals && (!lut || !lut->d_blob)) return POLCUE_EINVAL;
    if (B == 0) return POLCUE_OK;
    if (chunk_samples <= 0) chunk_samples = 8;
    if (chunk_samples > B) chunk_samples = B;

    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return (int)err;
    if (lut && lut->device >= 0 && lut->device != dev) return POLCUE_EINVAL;
    std::shared_ptr<FrontRing> ring;
    const FrontKey key{dev, chunk_samples, in_h, in_w, out_h, out_w, h_normals ? 1 : 0, h_xolp_norm ? 1 : 0, 1};
    int rc = g_front_rings.acquire(key, ring);
    if (rc != POLCUE_OK) return rc;
    std::lock_guard<std::mutex> busy(ring->busy);
    FrontRing& r = *ring;
    Lanes& q = r.lanes;
    const size_t img = (size_t)in_h * in_w, opx = (size_t)out_h * out_w;
    const uint8_t* h_src[4] = {h_i0, h_i45, h_i90, h_i135};

    int slot = 0;
    for (int first = 0, it = 0; first < B && rc == POLCUE_OK && err == cudaSuccess; first += chunk_samples, ++it, slot = (slot + 1) % kSlots) {
        const int nb = (B - first < chunk_samples) ? B - first : chunk_samples;
        if (it >= kSlots) PC_CHECK(cudaStreamWaitEvent(q.s_in, q.out_done[slot], 0));
        for (int k = 0; k < 4; ++k)
            PC_CHECK(cudaMemcpyAsync(r.d_in[slot] + (size_t)k * chunk_samples * img, h_src[k] + (size_t)first * img, (size_t)nb * img,
                                     cudaMemcpyHostToDevice, q.s_in));
        if (h_flip) PC_CHECK(cudaMemcpyAsync(r.d_flip[slot], h_flip + first, (size_t)nb, cudaMemcpyHostToDevice, q.s_in));   // flags ride with their chunk
        PC_CHECK(cudaEventRecord(q.in_done[slot], q.s_in));
        PC_CHECK(cudaStreamWaitEvent(q.s_run, q.in_done[slot], 0));
        if (err != cudaSuccess) break;
        const uint8_t* d = r.d_in[slot];
        const size_t stride = (size_t)chunk_samples * img;
        rc = polcue_loader_front_end_u8(r.plan, d, d + stride, d + 2 * stride, d + 3 * stride, nb, h_flip ? r.d_flip[slot] : nullptr, lut,
                                        r.d_ws[slot], r.d_planes[slot], nullptr, r.d_xolp[slot], h_normals ? r.d_normals[slot] : nullptr,
                                        xolp_mean_std, h_xolp_norm ? r.d_xnorm[slot] : nullptr, q.s_run);
        if (rc != POLCUE_OK) break;
        PC_CHECK(cudaEventRecord(q.run_done[slot], q.s_run));
        PC_CHECK(cudaStreamWaitEvent(q.s_out, q.run_done[slot], 0));
        PC_CHECK(cudaMemcpyAsync(h_xolp + (size_t)first * 2 * opx, r.d_xolp[slot], (size_t)nb * 2 * opx * sizeof(float), cudaMemcpyDeviceToHost,
                                 q.s_out));
        if (h_planes)
            PC_CHECK(cudaMemcpyAsync(h_planes + (size_t)first * 4 * opx, r.d_planes[slot], (size_t)nb * 4 * opx, cudaMemcpyDeviceToHost, q.s_out));
        if (h_xolp_norm)
            PC_CHECK(cudaMemcpyAsync(h_xolp_norm + (size_t)first * 2 * opx, r.d_xnorm[slot], (size_t)nb * 2 * opx * sizeof(float),
                                     cudaMemcpyDeviceToHost, q.s_out));
        if (h_normals)
            PC_CHECK(cudaMemcpyAsync(h_normals + (size_t)first * 9 * opx, r.d_normals[slot], (size_t)nb * 9 * opx * sizeof(float),
                                     cudaMemcpyDeviceToHost, q.s_out));
        PC_CHECK(cudaEventRecord(q.out_done[slot], q.s_out));
    }
    const cudaError_t drained = q.drain();
    if (rc != POLCUE_OK) return rc;
    if (err != cudaSuccess) return (int)err;
    return drained == cudaSuccess ? POLCUE_OK : (int)drained;
}

}  // extern "C"
