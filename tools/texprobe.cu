// Diagnostic: prints what tld4 returns on this GPU for a uchar4 gather array with border addressing, so that the
// component order and the footprint selection assumed by sample_bilinear() (kernels_warp.cu) can be read off directly.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/texprobe tools/texprobe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void probe(cudaTextureObject_t t, const float2* at, int n, uint4* out) {
    int i = threadIdx.x;
    if (i >= n) return;
    uint32_t a, b, c, d;
    asm("tld4.r.2d.v4.u32.f32 {%0, %1, %2, %3}, [%4, {%5, %6}];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(t), "f"(at[i].x), "f"(at[i].y));
    out[i] = make_uint4(a, b, c, d);
}

int main() {
    const int W = 4, H = 3;
    uchar4 img[H][W];
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) img[y][x] = make_uchar4(10 * (y + 1) + x + 1, 100 + x, 200 + y, 0);
    cudaArray_t arr; cudaChannelFormatDesc f = cudaCreateChannelDesc<uchar4>();
    if (cudaMallocArray(&arr, &f, W, H, cudaArrayTextureGather) != cudaSuccess) { printf("cudaMallocArray failed\n"); return 1; }
    cudaMemcpy2DToArray(arr, 0, 0, img, W * 4, W * 4, H, cudaMemcpyHostToDevice);
    cudaResourceDesc r{}; r.resType = cudaResourceTypeArray; r.res.array.array = arr;
    cudaTextureDesc td{}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder; td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    cudaTextureObject_t t; if (cudaCreateTextureObject(&t, &r, &td, nullptr) != cudaSuccess) { printf("texobj failed\n"); return 1; }
    // (X+1, Y+1) for X,Y = (0,0), (1,1), (-1,0), (3,2), (-2,0), (4,0), (2,-1)
    float2 h_at[] = {{1, 1}, {2, 2}, {0, 1}, {4, 3}, {-1, 1}, {5, 1}, {3, 0},
                     {6.7108864e7f, 1}, {-6.7108864e7f, 1}, {1, 6.7108864e7f}, {1, -6.7108864e7f}, {1e30f, 1}, {-1e30f, -1e30f}, {2147483648.f, 1}, {65536.f, 1}, {-65536.f, 1}, {16777216.f, 16777216.f}};
    const int n = sizeof h_at / sizeof h_at[0];
    float2* d_at; uint4* d_out; cudaMalloc(&d_at, sizeof h_at); cudaMalloc(&d_out, n * sizeof(uint4));
    cudaMemcpy(d_at, h_at, sizeof h_at, cudaMemcpyHostToDevice);
    probe<<<1, 32>>>(t, d_at, n, d_out);
    uint4 o[n]; cudaError_t e = cudaMemcpy(o, d_out, sizeof o, cudaMemcpyDeviceToHost);
    printf("texel value = 10*(y+1)+(x+1); expected order [ (X,Y+1), (X+1,Y+1), (X+1,Y), (X,Y) ], 0 outside; status %s\n", cudaGetErrorString(e));
    for (int i = 0; i < n; ++i) printf("X=%2d Y=%2d -> %u %u %u %u\n", (int)h_at[i].x - 1, (int)h_at[i].y - 1, o[i].x, o[i].y, o[i].z, o[i].w);
    return 0;
}
