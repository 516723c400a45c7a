// image_kernels.cu — dense per-pixel stages: synthetic input, pyramid resize, 7x7 Gaussian blur, border fill.
// All integer fixed-point; HBM/L2-bound streaming kernels (no tensor cores: there is no contraction here).
#include "orbx_internal.cuh"

namespace orbx {

// ---------------------------------------------------------------------------------------------------------------
// orbx-synth-v1: procedural KITTI-like texture, generated directly into level 0.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lowbias32(uint32_t v) {
    v ^= v >> 16; v *= 0x7feb352dU; v ^= v >> 15; v *= 0x846ca68bU; v ^= v >> 16;
    return v;
}
__device__ __forceinline__ uint32_t lattice(uint32_t ix, uint32_t iy, uint32_t S, uint32_t seed) {
    return lowbias32((ix * 73856093U) ^ (iy * 19349663U) ^ (S * 83492791U) ^ (seed * 2654435761U)) & 255U;
}
template <uint32_t S>
__device__ __forceinline__ uint32_t value_noise(uint32_t x, uint32_t y, uint32_t seed) {
    const uint32_t ix = x / S, iy = y / S, fx = x % S, fy = y % S;
    const uint32_t acc = (S - fx) * (S - fy) * lattice(ix, iy, S, seed) + fx * (S - fy) * lattice(ix + 1, iy, S, seed) +
                         (S - fx) * fy * lattice(ix, iy + 1, S, seed) + fx * fy * lattice(ix + 1, iy + 1, S, seed);  // < 2^20 for S <= 64
    return acc / (S * S);
}

__global__ void __launch_bounds__(256) k_synth(uint8_t* __restrict__ l0, int w, int h, int pitch, size_t planeBytes, uint32_t seed0,
                                               int firstImage, int stereo) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= w || y >= h) return;
    const int img = firstImage + blockIdx.z;
    const uint32_t seed = stereo ? seed0 + (uint32_t)(img >> 1) : seed0 + (uint32_t)img;
    const bool right = stereo && (img & 1);
    const uint32_t sx = (uint32_t)x + (right ? 16u : 0u), sy = (uint32_t)y;
    const uint32_t nseed = right ? (seed ^ 0x5bd1e995U) : seed;
    const uint32_t base = (4 * value_noise<64>(sx, sy, seed) + 3 * value_noise<32>(sx, sy, seed) + 2 * value_noise<16>(sx, sy, seed) +
                           value_noise<8>(sx, sy, seed)) / 10;
    const uint32_t blk = lattice(sx / 24, sy / 18, 999, seed);
    const uint32_t noise = lowbias32((sx * 0x9E3779B1U) ^ (sy * 0x85EBCA77U) ^ (nseed * 0xC2B2AE3DU)) & 7U;
    const uint32_t p = (3 * base + 2 * blk) / 5 + noise;
    l0[(size_t)img * planeBytes + (size_t)y * pitch + x] = (uint8_t)min(p, 255u);
}

void launch_synth(uint8_t* l0, int w, int h, int pitch, size_t planeBytes, uint32_t seed0, int firstImage, int nImages, int stereo, cudaStream_t s) {
    dim3 grid((w + 63) / 64, (h + 3) / 4, nImages);
    k_synth<<<grid, 256, 0, s>>>(l0, w, h, pitch, planeBytes, seed0, firstImage, stereo);
}

// ---------------------------------------------------------------------------------------------------------------
// Pyramid level l from level l-1: OpenCV's 8-bit INTER_LINEAR (11-bit coefficients), replacing
// ComputePyramid's cv::resize (ORBextractor.cc:1143-1168).  One thread = 4 destination pixels = one 32-bit store.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_resize(const uint8_t* __restrict__ src, int sw, int sh, int spitch, size_t splane,
                                                uint8_t* __restrict__ dst, int dw, int dh, int dpitch, size_t dplane, ResizeTab t) {
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (y >= dh || x0 >= dw) return;
    const uint8_t* S = src + (size_t)blockIdx.z * splane;
    const int yo = t.yofs[y];
    const int sy0 = min(max(yo, 0), sh - 1), sy1 = min(max(yo + 1, 0), sh - 1);
    const int b0 = t.yb[2 * y], b1 = t.yb[2 * y + 1];
    const uint8_t* __restrict__ r0 = S + (size_t)sy0 * spitch;
    const uint8_t* __restrict__ r1 = S + (size_t)sy1 * spitch;
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = x0 + i;
        if (x < dw) {
            const int sx = t.xofs[x];
            const int sx1 = min(sx + 1, sw - 1);
            const int a0 = t.xa[2 * x], a1 = t.xa[2 * x + 1];
            const int h0 = (int)__ldg(r0 + sx) * a0 + (int)__ldg(r0 + sx1) * a1;
            const int h1 = (int)__ldg(r1 + sx) * a0 + (int)__ldg(r1 + sx1) * a1;
            int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            v = min(max(v, 0), 255);
            out |= (uint32_t)v << (8 * i);
        }
    }
    *reinterpret_cast<uint32_t*>(dst + (size_t)blockIdx.z * dplane + (size_t)y * dpitch + x0) = out;
}

void launch_pyramid_level(const Geom& g, const Buffers& b, int level, const ResizeTab& tab, int nImages, cudaStream_t s) {
    const LevelGeom& S = g.lv[level - 1];
    const LevelGeom& D = g.lv[level];
    dim3 grid((D.w + 127) / 128, (D.h + 7) / 8, nImages), block(32, 8);
    k_resize<<<grid, block, 0, s>>>(b.pyr + S.planeOff + (size_t)b.imgBase * S.planeBytes, S.w, S.h, S.pitch, (size_t)S.planeBytes,
                                   b.pyr + D.planeOff + (size_t)b.imgBase * D.planeBytes, D.w, D.h, D.pitch,
                                   (size_t)D.planeBytes, tab);
}

// ---------------------------------------------------------------------------------------------------------------
// 7x7 sigma=2 Gaussian blur, separable 8.8 fixed point with BORDER_REFLECT_101 at the level's own edge
// (GaussianBlur on a clone, ORBextractor.cc:1105-1106; arithmetic SURVEY.md A.3).
// ---------------------------------------------------------------------------------------------------------------
struct Taps { int k[7]; };

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return min(max(i, 0), n - 1);
}

constexpr int BT_W = 64, BT_H = 16;

__global__ void __launch_bounds__(256) k_blur(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int w, int h, int pitch, size_t plane,
                                              Taps taps) {
    __shared__ uint8_t sp[BT_H + 6][BT_W + 8];
    __shared__ uint16_t sh[BT_H + 6][BT_W];
    const int tx0 = blockIdx.x * BT_W, ty0 = blockIdx.y * BT_H;
    const uint8_t* S = src + (size_t)blockIdx.z * plane;
    for (int i = threadIdx.x; i < (BT_H + 6) * (BT_W + 6); i += 256) {
        const int r = i / (BT_W + 6), c = i - r * (BT_W + 6);
        const int gx = reflect101(tx0 + c - 3, w), gy = reflect101(ty0 + r - 3, h);
        sp[r][c] = __ldg(S + (size_t)gy * pitch + gx);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (BT_H + 6) * BT_W; i += 256) {
        const int r = i / BT_W, c = i - r * BT_W;
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < 7; ++k) acc += (uint32_t)taps.k[k] * sp[r][c + k];
        sh[r][c] = (uint16_t)min(acc, 65535u);
    }
    __syncthreads();
    uint8_t* D = dst + (size_t)blockIdx.z * plane;
    for (int i = threadIdx.x; i < BT_H * BT_W; i += 256) {
        const int r = i / BT_W, c = i - r * BT_W;
        const int gx = tx0 + c, gy = ty0 + r;
        if (gx < w && gy < h) {
            uint32_t acc = 0;
#pragma unroll
            for (int k = 0; k < 7; ++k) acc += (uint32_t)taps.k[k] * sh[r + k][c];
            D[(size_t)gy * pitch + gx] = (uint8_t)min((acc + 32768u) >> 16, 255u);
        }
    }
}

void launch_blur(const Geom& g, const Buffers& b, int level, int nImages, const int taps[7], cudaStream_t s) {
    const LevelGeom& L = g.lv[level];
    Taps t;
    for (int i = 0; i < 7; ++i) t.k[i] = taps[i];
    dim3 grid((L.w + BT_W - 1) / BT_W, (L.h + BT_H - 1) / BT_H, nImages);
    const size_t o = (size_t)L.planeOff + (size_t)b.imgBase * L.planeBytes;
    k_blur<<<grid, 256, 0, s>>>(b.pyr + o, b.blur + o, L.w, L.h, L.pitch, (size_t)L.planeBytes, t);
}

// ---------------------------------------------------------------------------------------------------------------
// BORDER_REFLECT_101 copy of one level into a (w+38) x (h+38) staging image (copyMakeBorder, ORBextractor.cc:1158-1164);
// only used when a caller asks for mvImagePyramid with its border.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_border(const uint8_t* __restrict__ src, int w, int h, int pitch, uint8_t* __restrict__ dst, int dpitch) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w + 2 * kEdge) return;
    dst[(size_t)y * dpitch + x] = src[(size_t)reflect101(y - kEdge, h) * pitch + reflect101(x - kEdge, w)];
}

void launch_border(const uint8_t* src, int w, int h, int pitch, uint8_t* dst, int dpitch, cudaStream_t s) {
    dim3 grid((w + 2 * kEdge + 255) / 256, h + 2 * kEdge);
    k_border<<<grid, 256, 0, s>>>(src, w, h, pitch, dst, dpitch);
}

}  // namespace orbx
