// HBM ceiling probe for write-dominated streaming kernels on B200 (measurement tool, not product code).
// The fused polarization kernel moves 4 B in + 44 B out per output pixel (11 float planes).  These kernels move
// the same bytes with the same addresses but no arithmetic, in several read/scheduling flavours, so the
// achievable memory-system ceiling of the ACCESS PATTERN can be told apart from the kernel's own cost.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/hbm_probe tools/hbm_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int T = 512;
constexpr int PLANES = 11;

__device__ __forceinline__ void st4(float* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ unsigned ldu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void store_planes(float* out, size_t plane, size_t g, float4 v) {
#pragma unroll
    for (int p = 0; p < PLANES; ++p) st4(out + (size_t)p * plane + 4 * g, v);
}
__device__ __forceinline__ float4 load_quads(const unsigned* in, size_t groups, size_t g) {
    const unsigned a = ldu(in + g), b = ldu(in + groups + g), c = ldu(in + 2 * groups + g), d = ldu(in + 3 * groups + g);
    return make_float4(__uint_as_float(a | 0x3f800000u), __uint_as_float(b | 0x3f800000u), __uint_as_float(c | 0x3f800000u),
                       __uint_as_float(d | 0x3f800000u));
}

// MODE 0: write only.  1: dependent reads (what the fused kernel does).  2: independent reads (values unused by the stores).
// 3: software prefetch one iteration ahead.  4: one 16-byte read per thread from a single stream.
template <int MODE>
__global__ void __launch_bounds__(T, 2) stat_kernel(const unsigned* __restrict__ in, float* __restrict__ out, size_t groups, size_t plane,
                                                    unsigned* sink) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned acc = 0;
    float4 nxt = make_float4(1.f, 2.f, 3.f, 4.f);
    if (MODE == 3 && g < groups) nxt = load_quads(in, groups, g);
    for (; g < groups; g += stride) {
        float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
        if (MODE == 1) v = load_quads(in, groups, g);
        if (MODE == 2) { const float4 t = load_quads(in, groups, g); acc += __float_as_uint(t.x) ^ __float_as_uint(t.y) ^ __float_as_uint(t.z) ^ __float_as_uint(t.w); }
        if (MODE == 3) { v = nxt; if (g + stride < groups) nxt = load_quads(in, groups, g + stride); }
        if (MODE == 4) { const uint4 t = *reinterpret_cast<const uint4*>(in + 4 * g); v = make_float4(__uint_as_float(t.x | 0x3f800000u), __uint_as_float(t.y | 0x3f800000u), __uint_as_float(t.z | 0x3f800000u), __uint_as_float(t.w | 0x3f800000u)); }
        store_planes(out, plane, g, v);
    }
    if (MODE == 2 && acc == 0x12345678u) *sink = acc;
}

// Dynamic tile scheduler: persistent CTAs pull tiles of T * ITEMS groups from an atomic counter.
// BATCH: all ITEMS reads are issued before the first store (bursty reads per CTA).
template <int ITEMS, bool BATCH>
__global__ void __launch_bounds__(T, 2) dyn_kernel(const unsigned* __restrict__ in, float* __restrict__ out, size_t groups, size_t plane,
                                                   unsigned* counter) {
    __shared__ unsigned tile_s;
    const unsigned tiles = (unsigned)((groups + (size_t)T * ITEMS - 1) / ((size_t)T * ITEMS));
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) tile_s = atomicAdd(counter, 1u);
        __syncthreads();
        const unsigned tile = tile_s;
        if (tile >= tiles) break;
        const size_t g0 = (size_t)tile * T * ITEMS + threadIdx.x;
        if (BATCH) {
            float4 v[ITEMS];
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) if (g0 + (size_t)i * T < groups) v[i] = load_quads(in, groups, g0 + (size_t)i * T);
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) if (g0 + (size_t)i * T < groups) store_planes(out, plane, g0 + (size_t)i * T, v[i]);
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) if (g0 + (size_t)i * T < groups) store_planes(out, plane, g0 + (size_t)i * T, load_quads(in, groups, g0 + (size_t)i * T));
        }
    }
}

// Cluster Launch Control (sm_100): one CTA per tile in the grid; a resident CTA cancels a pending CTA and takes
// over its tile, so tiles are handed out dynamically by the hardware work distributor (no global counter).
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int TT, int MINB>
__global__ void __launch_bounds__(TT, MINB) clc_kernel(const unsigned* __restrict__ in, float* __restrict__ out, size_t groups, size_t plane,
                                                   unsigned* tiles_done) {
    __shared__ __align__(16) uint4 resp;
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned tile = blockIdx.x, phase = 0;
    for (;;) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 16;" ::"r"(s32(&bar)) : "memory");
            asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];" ::"r"(s32(&resp)), "r"(s32(&bar)) : "memory");
        }
        const size_t g = (size_t)tile * TT + threadIdx.x;
        if (g < groups) store_planes(out, plane, g, load_quads(in, groups, g));
        if (threadIdx.x == 0 && tiles_done) atomicAdd(tiles_done, 1u);
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s32(&bar)), "r"(phase) : "memory");
        }
        phase ^= 1;
        unsigned valid, next;
        asm volatile("{\n\t.reg .pred p1;\n\t.reg .b128 r;\n\tld.shared.b128 r, [%2];\n\t"
                     "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\n\tselp.u32 %1, 1, 0, p1;\n\t"
                     "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, _, _, _}, r;\n\t}"
                     : "=r"(next), "=r"(valid) : "r"(s32(&resp)) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();   // everyone has read the response before thread 0 re-arms it
        if (!valid) break;
        tile = next;
    }
}

// CLC tiles + bursty input prefetch: the CTA that owns tile t with t % BURST == 0 asks the L2 to fetch the input of the
// next BURST tiles (one bulk prefetch per quadrant stream), so DRAM sees the reads in large clustered bursts instead
// of a trickle interleaved with the writes.
template <int TT, int MINB, int BURST, int AHEAD>
__global__ void __launch_bounds__(TT, MINB) clc_prefetch_kernel(const unsigned* __restrict__ in, float* __restrict__ out, size_t groups,
                                                                 size_t plane) {
    __shared__ __align__(16) uint4 resp;
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned tile = blockIdx.x, phase = 0;
    for (;;) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 16;" ::"r"(s32(&bar)) : "memory");
            asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];" ::"r"(s32(&resp)), "r"(s32(&bar)) : "memory");
        }
        if (tile % BURST == 0 && threadIdx.x < 4) {
            const size_t g0 = ((size_t)tile + AHEAD) * TT;               // first group of the burst, AHEAD tiles ahead
            if (g0 < groups) {
                size_t n = (size_t)BURST * TT;
                if (g0 + n > groups) n = groups - g0;
                const unsigned* src = in + (size_t)threadIdx.x * groups + g0;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((unsigned)(n * 4)) : "memory");
            }
        }
        const size_t g = (size_t)tile * TT + threadIdx.x;
        if (g < groups) store_planes(out, plane, g, load_quads(in, groups, g));
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s32(&bar)), "r"(phase) : "memory");
        }
        phase ^= 1;
        unsigned valid, next;
        asm volatile("{\n\t.reg .pred p1;\n\t.reg .b128 r;\n\tld.shared.b128 r, [%2];\n\t"
                     "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\n\tselp.u32 %1, 1, 0, p1;\n\t"
                     "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, _, _, _}, r;\n\t}"
                     : "=r"(next), "=r"(valid) : "r"(s32(&resp)) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (!valid) break;
        tile = next;
    }
}

__global__ void __launch_bounds__(T, 2) copy_kernel(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i];
}

template <typename K>
float time_ms(K launch, int reps) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

int main(int argc, char** argv) {
    const size_t pad = argc > 1 ? (size_t)atol(argv[1]) : 0;   // extra floats between planes (DRAM bank-hash sensitivity)
    const size_t px = 64ull * 1024 * 1224;       // output pixels of one cfg2 batch
    const size_t groups = px / 4, plane = px + pad;
    unsigned *in, *counter; float* out;
    CK(cudaMalloc(&in, px * 4));                  // 4 quadrant words (u32) per group of 4 px = 4 B per output pixel
    CK(cudaMalloc(&out, plane * PLANES * sizeof(float)));
    CK(cudaMalloc(&counter, 256));
    CK(cudaMemset(in, 1, px * 4));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int reps = 10;
    printf("plane pad = %zu floats\n", pad);
    auto report = [&](const char* name, double bytes, float ms) { printf("  %-46s %8.3f ms  %8.1f GB/s\n", name, ms, bytes / ms / 1e6); };
    const double W = px * 44.0, RW = px * 48.0;
    for (int cps : {2, 4, 8, 16}) {
        const int grid = sms * cps;
        printf("-- static grid-stride, grid = %d x %d CTAs of %d threads\n", sms, cps, T);
        report("write only", W, time_ms([&] { stat_kernel<0><<<grid, T>>>(in, out, groups, plane, counter); }, reps));
        report("4R+44W dependent reads (fused kernel pattern)", RW, time_ms([&] { stat_kernel<1><<<grid, T>>>(in, out, groups, plane, counter); }, reps));
        report("4R+44W independent reads", RW, time_ms([&] { stat_kernel<2><<<grid, T>>>(in, out, groups, plane, counter); }, reps));
        report("4R+44W prefetch one ahead", RW, time_ms([&] { stat_kernel<3><<<grid, T>>>(in, out, groups, plane, counter); }, reps));
        report("4R+44W single 16-byte read stream", RW, time_ms([&] { stat_kernel<4><<<grid, T>>>(in, out, groups, plane, counter); }, reps));
    }
    printf("-- dynamic tiles, persistent grid = %d x 2\n", sms);
    const int grid = sms * 2;
    auto dyn = [&](auto kern) { return time_ms([&] { cudaMemsetAsync(counter, 0, 4); kern<<<grid, T>>>(in, out, groups, plane, counter); }, reps); };
    report("tile = 1 x 512 groups", RW, dyn(dyn_kernel<1, false>));
    report("tile = 4 x 512 groups", RW, dyn(dyn_kernel<4, false>));
    report("tile = 4 x 512 groups, batched reads", RW, dyn(dyn_kernel<4, true>));
    report("tile = 8 x 512 groups, batched reads", RW, dyn(dyn_kernel<8, true>));
    report("tile = 16 x 512 groups, batched reads", RW, dyn(dyn_kernel<16, true>));
    {
        const unsigned tiles = (unsigned)((groups + T - 1) / T);
        report("CLC (cluster launch control), tile = 512 groups", RW, time_ms([&] { clc_kernel<512, 2><<<tiles, 512>>>(in, out, groups, plane, nullptr); }, reps));
        report("CLC, tile = 128 groups (128 thr, 8 CTA/SM)", RW, time_ms([&] { clc_kernel<128, 8><<<(unsigned)((groups + 127) / 128), 128>>>(in, out, groups, plane, nullptr); }, reps));
        report("CLC, tile = 256 groups (256 thr, 4 CTA/SM)", RW, time_ms([&] { clc_kernel<256, 4><<<(unsigned)((groups + 255) / 256), 256>>>(in, out, groups, plane, nullptr); }, reps));
        report("CLC, tile = 1024 groups (1024 thr, 1 CTA/SM)", RW, time_ms([&] { clc_kernel<1024, 1><<<(unsigned)((groups + 1023) / 1024), 1024>>>(in, out, groups, plane, nullptr); }, reps));
        report("CLC, tile = 512 groups, 4 CTA/SM", RW, time_ms([&] { clc_kernel<512, 4><<<tiles, 512>>>(in, out, groups, plane, nullptr); }, reps));
        {
            const unsigned t256 = (unsigned)((groups + 255) / 256);
            report("CLC 256 thr, prefetch burst 64 tiles, 64 ahead", RW, time_ms([&] { clc_prefetch_kernel<256, 4, 64, 64><<<t256, 256>>>(in, out, groups, plane); }, reps));
            report("CLC 256 thr, prefetch burst 256 tiles, 256 ahead", RW, time_ms([&] { clc_prefetch_kernel<256, 4, 256, 256><<<t256, 256>>>(in, out, groups, plane); }, reps));
            report("CLC 256 thr, prefetch burst 1024 tiles, 1024 ahead", RW, time_ms([&] { clc_prefetch_kernel<256, 4, 1024, 1024><<<t256, 256>>>(in, out, groups, plane); }, reps));
            report("CLC 256 thr, prefetch burst 4096 tiles, 2048 ahead", RW, time_ms([&] { clc_prefetch_kernel<256, 4, 4096, 2048><<<t256, 256>>>(in, out, groups, plane); }, reps));
        }
        report("plain launch, one CTA per 512-group tile", RW, time_ms([&] { stat_kernel<1><<<tiles, T>>>(in, out, groups, plane, counter); }, reps));
        // every tile must be processed exactly once
        CK(cudaMemset(counter, 0, 4));
        CK(cudaMemset(out, 0, plane * PLANES * sizeof(float)));
        clc_kernel<512, 2><<<tiles, T>>>(in, out, groups, plane, counter);
        CK(cudaDeviceSynchronize());
        unsigned done = 0; CK(cudaMemcpy(&done, counter, 4, cudaMemcpyDeviceToHost));
        float probe[4] = {0, 0, 0, 0};
        CK(cudaMemcpy(probe, out + (size_t)10 * plane + 4 * (groups - 1), 16, cudaMemcpyDeviceToHost));
        printf("  CLC check: tiles processed %u of %u, last group of last plane = %g (expect non-zero)\n", done, tiles, probe[0]);
    }
    const size_t n4 = px * 5 / 4;
    report("copy float4 (R+W), 148 x 8", n4 * 32.0, time_ms([&] { copy_kernel<<<sms * 8, T>>>((const float4*)out, (float4*)out + n4 + 1024, n4); }, reps));
    report("cudaMemcpy D2D (R+W)", 2.0 * px * 20.0, time_ms([&] { cudaMemcpyAsync(out, out + px * 5 + 4096, px * 20, cudaMemcpyDeviceToDevice); }, reps));
    return 0;
}
