// Host-buffer entry point of the fused pipeline: the call a user of the Python/C API makes when the
// mosaics live in host memory.  Chunks of frames flow through a 3-slot ring on three streams so the
// host->device copy of chunk i+1, the kernel on chunk i and the device->host copy of chunk i-1 overlap
// (PCIe is full duplex).  The device ring is cached per (device, geometry) so repeated calls only pay
// copies + kernels.  bench.py's `e2e` figure times exactly this function.
#include <mutex>

#include "polcue_host.h"

namespace {

constexpr int kSlots = 3;

struct Ring {
    int device = -1, chunk = 0, H = 0, W = 0;
    bool with_iun = false, with_normals = false;
    unsigned char* d_mosaic[kSlots] = {};
    float* d_iun[kSlots] = {};
    float* d_xolp[kSlots] = {};
    float* d_normals[kSlots] = {};
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    cudaEvent_t in_done[kSlots] = {}, run_done[kSlots] = {}, out_done[kSlots] = {};

    void release() {
        for (int i = 0; i < kSlots; ++i) {
            cudaFree(d_mosaic[i]);
            cudaFree(d_iun[i]);
            cudaFree(d_xolp[i]);
            cudaFree(d_normals[i]);
            d_mosaic[i] = nullptr;
            d_iun[i] = d_xolp[i] = d_normals[i] = nullptr;
            if (in_done[i]) cudaEventDestroy(in_done[i]);
            if (run_done[i]) cudaEventDestroy(run_done[i]);
            if (out_done[i]) cudaEventDestroy(out_done[i]);
            in_done[i] = run_done[i] = out_done[i] = nullptr;
        }
        if (s_in) cudaStreamDestroy(s_in);
        if (s_run) cudaStreamDestroy(s_run);
        if (s_out) cudaStreamDestroy(s_out);
        s_in = s_run = s_out = nullptr;
        device = -1;
    }

    cudaError_t ensure(int dev, int chunk_frames, int h, int w, bool iun, bool normals) {
        if (device == dev && chunk == chunk_frames && H == h && W == w && with_iun == iun && with_normals == normals)
            return cudaSuccess;
        release();
        const size_t px = (size_t)(h / 2) * (w / 2);
        cudaError_t e = cudaSuccess;
        for (int i = 0; i < kSlots && e == cudaSuccess; ++i) {
            e = cudaMalloc(&d_mosaic[i], (size_t)chunk_frames * h * w);
            if (e == cudaSuccess) e = cudaMalloc(&d_xolp[i], chunk_frames * 2 * px * sizeof(float));
            if (e == cudaSuccess && iun) e = cudaMalloc(&d_iun[i], chunk_frames * px * sizeof(float));
            if (e == cudaSuccess && normals) e = cudaMalloc(&d_normals[i], chunk_frames * 9 * px * sizeof(float));
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&in_done[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&run_done[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&out_done[i], cudaEventDisableTiming);
        }
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s_run, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            release();
            return e;
        }
        device = dev;
        chunk = chunk_frames;
        H = h;
        W = w;
        with_iun = iun;
        with_normals = normals;
        return cudaSuccess;
    }
};

std::mutex g_ring_mutex;
Ring g_ring;

}  // namespace

extern "C" {

int polcue_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return POLCUE_EINVAL;
    const cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocDefault);
    return e == cudaSuccess ? POLCUE_OK : (int)e;
}

int polcue_host_free(void* ptr) {
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

    std::lock_guard<std::mutex> guard(g_ring_mutex);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    e = g_ring.ensure(dev, chunk_frames, H, W, h_iun != nullptr, h_normals != nullptr);
    if (e != cudaSuccess) return e == cudaErrorMemoryAllocation ? POLCUE_ENOMEM : (int)e;
    Ring& r = g_ring;
    const size_t px = (size_t)(H / 2) * (W / 2), frame = (size_t)H * W;

    int rc = POLCUE_OK;
    int slot = 0;
    for (int first = 0, it = 0; first < B && rc == POLCUE_OK; first += chunk_frames, ++it, slot = (slot + 1) % kSlots) {
        const int nb = (B - first < chunk_frames) ? B - first : chunk_frames;
        if (it >= kSlots) cudaStreamWaitEvent(r.s_in, r.out_done[slot], 0);   // slot's previous results are on the host
        cudaMemcpyAsync(r.d_mosaic[slot], h_mosaic + (size_t)first * frame, (size_t)nb * frame, cudaMemcpyHostToDevice, r.s_in);
        cudaEventRecord(r.in_done[slot], r.s_in);
        cudaStreamWaitEvent(r.s_run, r.in_done[slot], 0);
        rc = polcue_fused_mosaic_u8(r.d_mosaic[slot], nb, H, W, lut, nullptr, r.d_iun[slot], r.d_xolp[slot], r.d_normals[slot],
                                    r.s_run);
        cudaEventRecord(r.run_done[slot], r.s_run);
        cudaStreamWaitEvent(r.s_out, r.run_done[slot], 0);
        cudaMemcpyAsync(h_xolp + (size_t)first * 2 * px, r.d_xolp[slot], (size_t)nb * 2 * px * sizeof(float),
                        cudaMemcpyDeviceToHost, r.s_out);
        if (h_iun)
            cudaMemcpyAsync(h_iun + (size_t)first * px, r.d_iun[slot], (size_t)nb * px * sizeof(float), cudaMemcpyDeviceToHost,
                            r.s_out);
        if (h_normals)
            cudaMemcpyAsync(h_normals + (size_t)first * 9 * px, r.d_normals[slot], (size_t)nb * 9 * px * sizeof(float),
                            cudaMemcpyDeviceToHost, r.s_out);
        cudaEventRecord(r.out_done[slot], r.s_out);
    }
    e = cudaStreamSynchronize(r.s_out);
    if (e == cudaSuccess) e = cudaStreamSynchronize(r.s_run);
    if (e == cudaSuccess) e = cudaStreamSynchronize(r.s_in);
    if (rc != POLCUE_OK) return rc;
    return e == cudaSuccess ? POLCUE_OK : (int)e;
}

}  // extern "C"
