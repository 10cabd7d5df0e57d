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

// The same three-stream ring for the loader front end: four full-resolution image stacks in, resized planes, XOLP,
// normalised XOLP and normals out.
struct FrontRing {
    int device = -1, chunk = 0;
    const polcue_resize_plan* plan = nullptr;
    int in_h = 0, in_w = 0, out_h = 0, out_w = 0;
    unsigned char* d_in[kSlots] = {};        // 4 x chunk x in_h x in_w
    unsigned char* d_ws[kSlots] = {};
    unsigned char* d_planes[kSlots] = {};
    float* d_xolp[kSlots] = {};
    float* d_xnorm[kSlots] = {};
    float* d_normals[kSlots] = {};
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    cudaEvent_t in_done[kSlots] = {}, run_done[kSlots] = {}, out_done[kSlots] = {};

    void release() {
        for (int i = 0; i < kSlots; ++i) {
            cudaFree(d_in[i]); cudaFree(d_ws[i]); cudaFree(d_planes[i]); cudaFree(d_xolp[i]); cudaFree(d_xnorm[i]); cudaFree(d_normals[i]);
            d_in[i] = d_ws[i] = d_planes[i] = nullptr;
            d_xolp[i] = d_xnorm[i] = d_normals[i] = nullptr;
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

    cudaError_t ensure(int dev, int chunk_samples, const polcue_resize_plan* pl, int ih, int iw, int oh, int ow) {
        if (device == dev && chunk == chunk_samples && in_h == ih && in_w == iw && out_h == oh && out_w == ow) return cudaSuccess;
        release();
        const size_t opx = (size_t)oh * ow;
        cudaError_t e = cudaSuccess;
        for (int i = 0; i < kSlots && e == cudaSuccess; ++i) {
            e = cudaMalloc(&d_in[i], (size_t)4 * chunk_samples * ih * iw);
            if (e == cudaSuccess) e = cudaMalloc(&d_ws[i], polcue_resize_workspace_bytes(pl, 4 * chunk_samples));
            if (e == cudaSuccess) e = cudaMalloc(&d_planes[i], (size_t)chunk_samples * 4 * opx);
            if (e == cudaSuccess) e = cudaMalloc(&d_xolp[i], chunk_samples * 2 * opx * sizeof(float));
            if (e == cudaSuccess) e = cudaMalloc(&d_xnorm[i], chunk_samples * 2 * opx * sizeof(float));
            if (e == cudaSuccess) e = cudaMalloc(&d_normals[i], chunk_samples * 9 * opx * sizeof(float));
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
        device = dev; chunk = chunk_samples; plan = pl;
        in_h = ih; in_w = iw; out_h = oh; out_w = ow;
        return cudaSuccess;
    }
};

FrontRing g_front;

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

int polcue_loader_front_end_u8_host(int in_h, int in_w, int out_h, int out_w, const uint8_t* h_i0, const uint8_t* h_i45,
                                    const uint8_t* h_i90, const uint8_t* h_i135, int B, const uint8_t* h_flip, const polcue_lut* lut,
                                    uint8_t* h_planes, float* h_xolp, float* h_normals, const float* xolp_mean_std, float* h_xolp_norm,
                                    int chunk_samples) {
    if (!h_i0 || !h_i45 || !h_i90 || !h_i135 || !h_xolp || B < 0 || (h_xolp_norm && !xolp_mean_std)) return POLCUE_EINVAL;
    if (h_normals && (!lut || !lut->d_blob)) return POLCUE_EINVAL;
    if (B == 0) return POLCUE_OK;
    if (chunk_samples <= 0) chunk_samples = 8;
    if (chunk_samples > B) chunk_samples = B;

    std::lock_guard<std::mutex> guard(g_ring_mutex);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    // the plan (device weights) lives as long as the geometry stays the same
    static polcue_resize_plan* plan = nullptr;
    static int plan_key[5] = {-1, 0, 0, 0, 0};
    if (!plan || plan_key[0] != dev || plan_key[1] != in_h || plan_key[2] != in_w || plan_key[3] != out_h || plan_key[4] != out_w) {
        if (plan) polcue_resize_plan_destroy(plan);
        plan = nullptr;
        const int rc = polcue_resize_plan_create(in_h, in_w, out_h, out_w, &plan);
        if (rc != POLCUE_OK) return rc;
        plan_key[0] = dev; plan_key[1] = in_h; plan_key[2] = in_w; plan_key[3] = out_h; plan_key[4] = out_w;
    }
    e = g_front.ensure(dev, chunk_samples, plan, in_h, in_w, out_h, out_w);
    if (e != cudaSuccess) return e == cudaErrorMemoryAllocation ? POLCUE_ENOMEM : (int)e;
    FrontRing& r = g_front;
    const size_t img = (size_t)in_h * in_w, opx = (size_t)out_h * out_w;
    const uint8_t* h_src[4] = {h_i0, h_i45, h_i90, h_i135};
    // per-sample flip flags: a small device copy per call
    uint8_t* d_flip = nullptr;
    if (h_flip) {
        e = cudaMalloc(&d_flip, (size_t)B);
        if (e == cudaSuccess) e = cudaMemcpy(d_flip, h_flip, (size_t)B, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(d_flip);
            return (int)e;
        }
    }
    int rc = POLCUE_OK;
    int slot = 0;
    for (int first = 0, it = 0; first < B && rc == POLCUE_OK; first += chunk_samples, ++it, slot = (slot + 1) % kSlots) {
        const int nb = (B - first < chunk_samples) ? B - first : chunk_samples;
        if (it >= kSlots) cudaStreamWaitEvent(r.s_in, r.out_done[slot], 0);
        for (int k = 0; k < 4; ++k)
            cudaMemcpyAsync(r.d_in[slot] + (size_t)k * chunk_samples * img, h_src[k] + (size_t)first * img, (size_t)nb * img,
                            cudaMemcpyHostToDevice, r.s_in);
        cudaEventRecord(r.in_done[slot], r.s_in);
        cudaStreamWaitEvent(r.s_run, r.in_done[slot], 0);
        const uint8_t* d = r.d_in[slot];
        const size_t stride = (size_t)chunk_samples * img;
        rc = polcue_loader_front_end_u8(plan, d, d + stride, d + 2 * stride, d + 3 * stride, nb, d_flip ? d_flip + first : nullptr, lut,
                                        r.d_ws[slot], r.d_planes[slot], nullptr, r.d_xolp[slot], h_normals ? r.d_normals[slot] : nullptr,
                                        xolp_mean_std, h_xolp_norm ? r.d_xnorm[slot] : nullptr, r.s_run);
        cudaEventRecord(r.run_done[slot], r.s_run);
        cudaStreamWaitEvent(r.s_out, r.run_done[slot], 0);
        cudaMemcpyAsync(h_xolp + (size_t)first * 2 * opx, r.d_xolp[slot], (size_t)nb * 2 * opx * sizeof(float), cudaMemcpyDeviceToHost, r.s_out);
        if (h_planes) cudaMemcpyAsync(h_planes + (size_t)first * 4 * opx, r.d_planes[slot], (size_t)nb * 4 * opx, cudaMemcpyDeviceToHost, r.s_out);
        if (h_xolp_norm)
            cudaMemcpyAsync(h_xolp_norm + (size_t)first * 2 * opx, r.d_xnorm[slot], (size_t)nb * 2 * opx * sizeof(float), cudaMemcpyDeviceToHost, r.s_out);
        if (h_normals)
            cudaMemcpyAsync(h_normals + (size_t)first * 9 * opx, r.d_normals[slot], (size_t)nb * 9 * opx * sizeof(float), cudaMemcpyDeviceToHost, r.s_out);
        cudaEventRecord(r.out_done[slot], r.s_out);
    }
    e = cudaStreamSynchronize(r.s_out);
    if (e == cudaSuccess) e = cudaStreamSynchronize(r.s_run);
    if (e == cudaSuccess) e = cudaStreamSynchronize(r.s_in);
    cudaFree(d_flip);
    if (rc != POLCUE_OK) return rc;
    return e == cudaSuccess ? POLCUE_OK : (int)e;
}

}  // extern "C"
