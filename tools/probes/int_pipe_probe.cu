// Throughput of the integer multiply-add candidates for the Lanczos resize on sm_100a:
// IMAD, IDP.4A (dp4a), IDP.2A (dp2a), and PRMT, each as 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipe_probe int_pipe_probe.cu && ./int_pipe_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void probe(unsigned* out, unsigned a0, unsigned b0, int iters) {
    unsigned acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x + i;
    unsigned a = a0 + threadIdx.x, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) acc[i] = acc[i] * a + b;                                                         // IMAD
            else if (OP == 1) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b));   // IDP.4A
            else if (OP == 2) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b));
            else if (OP == 3) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(a), "r"(b));
            else if (OP == 4) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b));
            else if (OP == 5) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(a), "r"(b));
            // pairs: do the two ops share a pipe?  (64/clk/SM if they do, up to 128 if they do not)
            else if (OP == 10) { if (i & 1) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b)); else asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(a), "r"(b)); }
            else if (OP == 11) { if (i & 1) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b)); else acc[i] = acc[i] * a + b; }
            else if (OP == 12) { if (i & 1) acc[i] = acc[i] * a + b; else asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(a), "r"(b)); }
            else if (OP == 13) { if (i & 1) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b)); else asm volatile("add.u32 %0, %0, %1;" : "+r"(acc[i]) : "r"(a)); }
            else if (OP == 14) { if (i & 1) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b)); else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(a), "r"(b)); }
            else if (OP == 15) { if (i & 1) asm volatile("max.s32 %0, %0, %1;" : "+r"(acc[i]) : "r"(a)); else asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(a), "r"(b)); }
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= acc[i];
    if (s == 0x12345678u) out[0] = s;
}

// Packed FP32 (Blackwell): two FMAs per instruction.  Four 64-bit chains per thread.
template <int OP>
__global__ void probe2(unsigned long long* out, unsigned long long a0, unsigned long long b0, int iters) {
    unsigned long long acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = threadIdx.x + i;
    unsigned long long a = a0 + threadIdx.x, b = b0;
    unsigned x[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (OP == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(a), "l"(b));
            else if (OP == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(acc[i]) : "l"(a));
            else if (OP == 2) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(a), "l"(b)); asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"((unsigned)a), "r"((unsigned)b)); }
            else if (OP == 3) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(a), "l"(b)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"((unsigned)a), "r"((unsigned)b)); }
        }
    }
    unsigned long long s = x[0] ^ x[1] ^ x[2] ^ x[3];
#pragma unroll
    for (int i = 0; i < 4; ++i) s ^= acc[i];
    if (s == 0x12345678u) out[0] = s;
}

template <int OP>
void run2(const char* name, int per_iter) {
    unsigned long long* out;
    cudaMalloc(&out, 8);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 4096, threads = 1024, blocks = sms * 2;
    probe2<OP><<<blocks, threads>>>(out, 3, 5, 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe2<OP><<<blocks, threads>>>(out, 3, 5, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double instr = (double)blocks * threads * iters * per_iter;
    printf("%-10s %8.3f ms  %7.1f thread-instr/clk/SM (at %d MHz nominal)\n", name, ms, instr / (ms * 1e-3) / (khz * 1e3) / sms, khz / 1000);
    cudaFree(out);
}

template <int OP>
void run(const char* name) {
    unsigned* out;
    cudaMalloc(&out, 4);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 4096, threads = 1024, blocks = sms * 2;
    probe<OP><<<blocks, threads>>>(out, 3, 5, 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<OP><<<blocks, threads>>>(out, 3, 5, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double instr = (double)blocks * threads * iters * 8;
    printf("%-10s %8.3f ms  %7.1f thread-instr/clk/SM (at %d MHz nominal)\n", name, ms, instr / (ms * 1e-3) / (khz * 1e3) / sms, khz / 1000);
    cudaFree(out);
}

int main() {
    run<0>("IMAD");
    run<1>("IDP4A.UU");
    run<4>("IDP4A.US");
    run<2>("IDP2A");
    run<3>("PRMT");
    run<5>("FFMA");
    run<10>("IDP+PRMT");
    run<11>("IDP+IMAD");
    run<12>("IMAD+PRMT");
    run<13>("IDP+IADD");
    run<14>("IDP+FFMA");
    run<15>("IMNMX+PRMT");
    run2<0>("FFMA2", 4);
    run2<1>("FADD2", 4);
    run2<2>("FFMA2+PRMT", 8);
    run2<3>("FFMA2+FFMA", 8);
    return 0;
}
