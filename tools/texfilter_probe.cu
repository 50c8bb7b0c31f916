// Diagnostic: is the texture unit's bilinear filter (uchar4, normalized-float read mode, unnormalised coordinates,
// border addressing) exact enough to reproduce cv::remap's 8U INTER_LINEAR value
//     (S00*(32-fx)*(32-fy) + S01*fx*(32-fy) + S10*(32-fx)*fy + S11*fx*fy + 512) >> 10        
// for every 1/32-pixel position? The filter weights are 1.8 fixed point, so fx/32 is representable; what is not
// documented is the precision of the weighted sum. This probe measures it: max |255*tex - sum/1024| and the number
// of positions where round-to-nearest of 255*tex differs from the integer formula, over random positions (borders
// included) of a random texture with many extreme values.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/texfilter_probe tools/texfilter_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int W = 256, H = 192;

__device__ __forceinline__ uint32_t rng(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

__device__ __forceinline__ int texel(const uchar4* img, int x, int y, int c) {
    if ((unsigned)x >= (unsigned)W || (unsigned)y >= (unsigned)H) return 0;
    const uchar4 t = img[y * W + x];
    return c == 0 ? t.x : c == 1 ? t.y : t.z;
}

__global__ void probe(cudaTextureObject_t tex, const uchar4* img, int iters, unsigned long long* mism, float* maxerr,
                      unsigned long long* hard, unsigned long long* hard_bad) {
    uint32_t s = 0x9E3779B9u * (blockIdx.x * blockDim.x + threadIdx.x + 1);
    unsigned long long bad = 0, nh = 0, nhb = 0;
    float me = 0.f;
    for (int it = 0; it < iters; ++it) {
        const int sx = (int)(rng(s) % (unsigned)((W + 4) * 32)) - 64, sy = (int)(rng(s) % (unsigned)((H + 4) * 32)) - 64;
        const int X = sx >> 5, Y = sy >> 5, fx = sx & 31, fy = sy & 31;
        const float u = (float)(sx + 16) * (1.f / 32), v = (float)(sy + 16) * (1.f / 32);
        const float4 t = tex2D<float4>(tex, u, v);
        const float tv[3] = {t.x, t.y, t.z};
        for (int c = 0; c < 3; ++c) {
            const int sum = texel(img, X, Y, c) * (32 - fx) * (32 - fy) + texel(img, X + 1, Y, c) * fx * (32 - fy) +
                            texel(img, X, Y + 1, c) * (32 - fx) * fy + texel(img, X + 1, Y + 1, c) * fx * fy;
            const int want = (sum + 512) >> 10;
            const float hv = tv[c] * 255.f;
            const int got = __float2int_rd(hv + 0.5f);
            const float err = fabsf(hv - (float)sum * (1.f / 1024));
            me = fmaxf(me, err);
            const int r = sum & 1023;
            const bool is_hard = r == 511 || r == 512;
            nh += is_hard;
            if (got != want) { ++bad; nhb += is_hard; }
        }
    }
    atomicAdd(mism, bad);
    atomicAdd(hard, nh);
    atomicAdd(hard_bad, nhb);
    atomicMax(reinterpret_cast<int*>(maxerr), __float_as_int(me));
}

int main() {
    uchar4* h = (uchar4*)malloc(W * H * 4);
    srand(7);
    for (int i = 0; i < W * H; ++i) {
        auto v = [] { int r = rand() % 8; return (unsigned char)(r == 0 ? 0 : r == 1 ? 255 : rand() & 255); };
        h[i] = make_uchar4(v(), v(), v(), 0);
    }
    cudaArray_t arr; cudaChannelFormatDesc f = cudaCreateChannelDesc<uchar4>();
    if (cudaMallocArray(&arr, &f, W, H, cudaArrayTextureGather) != cudaSuccess) { printf("{\"error\": \"cudaMallocArray\"}\n"); return 1; }
    cudaMemcpy2DToArray(arr, 0, 0, h, W * 4, W * 4, H, cudaMemcpyHostToDevice);
    uchar4* d_img; cudaMalloc(&d_img, W * H * 4); cudaMemcpy(d_img, h, W * H * 4, cudaMemcpyHostToDevice);
    cudaResourceDesc r{}; r.resType = cudaResourceTypeArray; r.res.array.array = arr;
    cudaTextureDesc td{}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder; td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 0;
    cudaTextureObject_t t; if (cudaCreateTextureObject(&t, &r, &td, nullptr) != cudaSuccess) { printf("{\"error\": \"texobj\"}\n"); return 1; }
    unsigned long long *d_m, *d_h, *d_hb; float* d_e;
    cudaMalloc(&d_m, 8); cudaMalloc(&d_h, 8); cudaMalloc(&d_hb, 8); cudaMalloc(&d_e, 4);
    cudaMemset(d_m, 0, 8); cudaMemset(d_h, 0, 8); cudaMemset(d_hb, 0, 8); cudaMemset(d_e, 0, 4);
    const int blocks = 148 * 8, threads = 256, iters = 512;
    probe<<<blocks, threads>>>(t, d_img, iters, d_m, d_e, d_h, d_hb);
    unsigned long long m = 0, nh = 0, nhb = 0; float e = 0;
    cudaError_t st = cudaMemcpy(&m, d_m, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&nh, d_h, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&nhb, d_hb, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&e, d_e, 4, cudaMemcpyDeviceToHost);
    printf("{\"probe\": \"hw_bilinear_vs_cv_remap_q15\", \"status\": \"%s\", \"samples\": %llu, \"mismatches\": %llu, "
           "\"max_abs_err_levels\": %.9g, \"hard_cases\": %llu, \"hard_mismatches\": %llu}\n",
           cudaGetErrorString(st), (unsigned long long)blocks * threads * iters * 3, m, e, nh, nhb);
    return 0;
}
