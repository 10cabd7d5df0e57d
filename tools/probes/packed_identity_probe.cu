// Are the packed (f32x2) stencil helpers of normals_loss.cu bit-identical to the scalar ones?  Random tiles, every
// intermediate compared bit for bit.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o packed_identity_probe
//   tools/probes/packed_identity_probe.cu -L<pkg>/polcue -lpolcue -Xlinker -rpath=<pkg>/polcue
#include <cstdio>
#include "../../supervised-depth-estimation-from-polarized-images_b200/csrc/normals_loss.cu"

namespace polcue { std::atomic<unsigned long long> g_launches{0}; }
using namespace polcue;

__global__ void probe(const float* g, const float* pr, int* bad, float* dump) {
    __shared__ __align__(16) float tg[box_rows(8)][kLBoxW];
    __shared__ __align__(16) float tp[box_rows(8)][kLBoxW];
    __shared__ __align__(16) float2 T[8 + 2][kPPitch];
    const int H = 64, W = 256;
    stage_tile<8>(tg, g, H, W, 0, 0);
    stage_tile<8>(tp, pr, H, W, 0, 0);
    stage_pairs<8, 1, 1, kPPitch>(T, g, pr, H, W, 0, 0);
    __syncthreads();
    Cam cam{1.0f / 706.7f, 127.3f, 1.0f / 707.5f, 31.9f};
    const int tx0 = 4 * (threadIdx.x & 31), ty = threadIdx.x >> 5, x = tx0, y = ty;
    float gug[3][4], gvg[3][4], gup[3][4], gvp[3][4];
    gradients4(tg, ty, tx0, x, y, H, W, cam, gug, gvg);
    gradients4(tp, ty, tx0, x, y, H, W, cam, gup, gvp);
    f32x2 fx6[6], fy3[3];
    for (int c = 0; c < 6; ++c) fx6[c] = dup2(((float)min(max(x + c - 1, 0), W - 1) - cam.cx) * cam.inv_fx);
    for (int r = 0; r < 3; ++r) fy3[r] = dup2(((float)min(max(y + r - 1, 0), H - 1) - cam.cy) * cam.inv_fy);
    f32x2 gu[3][4], gv[3][4], centre[4];
    gradients4_pairs(&T[ty][0], tx0, kPPitch, fx6, fy3, gu, gv, centre);
    for (int j = 0; j < 4; ++j) {
        for (int c = 0; c < 3; ++c) {
            float a, b;
            unpk2(gu[c][j], a, b);
            if (__float_as_uint(a) != __float_as_uint(gug[c][j])) { atomicAdd(&bad[0], 1); dump[0] = a; dump[1] = gug[c][j]; dump[2] = c; }
            if (__float_as_uint(b) != __float_as_uint(gup[c][j])) atomicAdd(&bad[1], 1);
            unpk2(gv[c][j], a, b);
            if (__float_as_uint(a) != __float_as_uint(gvg[c][j])) { atomicAdd(&bad[2], 1); dump[3] = a; dump[4] = gvg[c][j]; dump[5] = c; }
            if (__float_as_uint(b) != __float_as_uint(gvp[c][j])) atomicAdd(&bad[3], 1);
        }
        const float ug[3] = {gug[0][j], gug[1][j], gug[2][j]}, vg[3] = {gvg[0][j], gvg[1][j], gvg[2][j]};
        const float up[3] = {gup[0][j], gup[1][j], gup[2][j]}, vp[3] = {gvp[0][j], gvp[1][j], gvp[2][j]};
        float n[3], a3[3], b3[3];
        cross_rn(ug, vg, n); normalize3(n, a3);
        cross_rn(up, vp, n); normalize3(n, b3);
        const f32x2 u[3] = {gu[0][j], gu[1][j], gu[2][j]}, v[3] = {gv[0][j], gv[1][j], gv[2][j]};
        f32x2 nn[3];
        unit_normals_pairs(u, v, nn);
        for (int c = 0; c < 3; ++c) {
            float a, b;
            unpk2(nn[c], a, b);
            if (__float_as_uint(a) != __float_as_uint(a3[c])) atomicAdd(&bad[4], 1);
            if (__float_as_uint(b) != __float_as_uint(b3[c])) atomicAdd(&bad[5], 1);
        }
    }
}

int main() {
    const int H = 64, W = 256;
    float *g, *p, *dump; int* bad;
    cudaMallocManaged(&g, H * W * 4); cudaMallocManaged(&p, H * W * 4); cudaMallocManaged(&bad, 32); cudaMallocManaged(&dump, 64);
    srand(1);
    for (int i = 0; i < H * W; ++i) { g[i] = 0.3f + (rand() % 1000) * 1e-3f; p[i] = 0.3f + (rand() % 1000) * 1.1e-3f; }
    for (int i = 0; i < 8; ++i) bad[i] = 0;
    probe<<<1, 256>>>(g, p, bad, dump);
    printf("sync: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    printf("mismatches gu(gt) %d gu(pred) %d gv(gt) %d gv(pred) %d  normals gt %d pred %d\n", bad[0], bad[1], bad[2], bad[3], bad[4], bad[5]);
    printf("gu sample packed %.9g scalar %.9g comp %g; gv sample packed %.9g scalar %.9g comp %g\n", dump[0], dump[1], dump[2], dump[3], dump[4], dump[5]);
    return 0;
}
