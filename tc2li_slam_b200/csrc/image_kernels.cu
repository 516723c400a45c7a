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

// Fast path for scale factors <= 1.25 (4 destination pixels then span at most 9 source bytes = 3 aligned words): each thread
// owns 4 fixed destination columns and walks down RS_SEG rows.  The warp's whole source footprint (<= 192 bytes x 24 rows)
// arrives with ONE TMA box load (16-byte aligned origin), so enough bytes are in flight to keep HBM busy; column
// coefficients and PRMT selectors live in registers, a source row is three aligned 32-bit shared-memory loads, the
// horizontal interpolation of a pixel is one PRMT + one IDP.2A (coefficients as 16x2 against two source bytes), and the
// horizontally interpolated upper row is reused as the next destination row's lower row whenever the source row repeats.
#ifndef RS_SEG_ROWS
#define RS_SEG_ROWS 16
#endif
#ifndef RS_NWARPS
#define RS_NWARPS 4
#endif
constexpr int RS_SEG = RS_SEG_ROWS, RS_WARPS = RS_NWARPS;
constexpr int RS_BOXW = 192, RS_BOXH = (RS_SEG * 5 + 3) / 4 + 4;  // 1.25 source rows per destination row + the second row of the last one + rounding

__global__ void __launch_bounds__(32 * RS_WARPS) k_resize_s12(const __grid_constant__ CUtensorMap srcMap, int sh, int imgBase, uint8_t* __restrict__ dst,
                                                             int dw, int dh, int dpitch, size_t dplane, ResizeTab t) {
    __shared__ __align__(128) unsigned char tile[RS_WARPS][RS_BOXW * RS_BOXH];
    __shared__ __align__(16) uint4 rowTab[RS_WARPS][RS_SEG];   // per destination row of the segment: {sy0, sy1, b0 << 12, b1 << 12}
    __shared__ __align__(8) uint64_t bars[RS_WARPS];
    const int wy = threadIdx.y;
    const int xw = blockIdx.x * 128;                       // first destination column of the warp
    const int x0 = xw + threadIdx.x * 4;
    const int ys = (blockIdx.y * RS_WARPS + wy) * RS_SEG;
    if (threadIdx.x == 0) mbar_init(&bars[wy], 1);
    mbar_fence_init();
    __syncwarp();
    if (xw >= dw || ys >= dh) return;                      // whole warp
    const int ye = min(ys + RS_SEG, dh);
    // box origin: the 16-byte aligned byte at or before the warp's first source pixel, the first source row of the segment
    const int boxX = (kMarginX + (int)t.xofs[xw]) & ~15;   // byte inside the padded plane row
    const int syMin = min(max((int)t.yofs[ys], 0), sh - 1);
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bars[wy], RS_BOXW * RS_BOXH);
        tma_load_3d(tile[wy], &srcMap, &bars[wy], boxX, kMarginY + syMin, imgBase + blockIdx.z);
    }
    uint32_t coef[4], sel[4];
    int base = 0;
    bool hi3 = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = min(x0 + i, dw - 1);  // (inactive lanes compute with the last column, harmlessly)
        const int sx = t.xofs[x];
        coef[i] = *reinterpret_cast<const uint32_t*>(t.xa + 2 * x);  // (1-fx)*2048 | fx*2048 << 16
        if (i == 0) base = sx & ~3;
        int bi = sx - base;          // 0..7; byte bi + 1 may be byte 8 (third word) for the last pixel only
        if (i == 3 && bi >= 7) { hi3 = true; bi -= 4; }
        sel[i] = (uint32_t)bi | ((uint32_t)(bi + 1) << 4);
    }
    // sel[3], the third-word flag and the *64 factor are per-thread constants of the segment; ptxas otherwise rebuilds them from `bi` / `dw` inside every
    // hrow (five + two instructions per source row, SASS listing profiles/r02/sass_k_resize_s12.txt): an identity shuffle makes them opaque
    sel[3] = __shfl_sync(0xFFFFFFFFu, sel[3], threadIdx.x);
    const uint32_t hi3w = __shfl_sync(0xFFFFFFFFu, hi3 ? 1u : 0u, threadIdx.x);
    // 32-bit shared-window address of the thread's first word (kept in a register; explicit ld.shared below)
    const uint32_t sBase = smem_u32(tile[wy]) + (uint32_t)(kMarginX + base - boxX) - (uint32_t)(syMin * RS_BOXW);
    uint8_t* D = dst + (size_t)blockIdx.z * dplane + x0 + (size_t)ys * dpitch;
    // the row table of the segment (clamped source rows, vertical weights pre-scaled for mad.hi): one row per lane into shared
    // memory, read back in the loop as one broadcast 128-bit load
    if (threadIdx.x < RS_SEG) {
        const int myRow = min(ys + (int)threadIdx.x, dh - 1);
        const int myYo = t.yofs[myRow];
        const uint32_t myBw = *reinterpret_cast<const uint32_t*>(t.yb + 2 * myRow);
        rowTab[wy][threadIdx.x] = make_uint4((uint32_t)min(max(myYo, 0), sh - 1), (uint32_t)min(max(myYo + 1, 0), sh - 1), (myBw & 0xFFFFu) << 12, (myBw >> 16) << 12);
    }
    __syncwarp();
    // Issue balance: the ALU pipe (LOP3 / SHF / PRMT / IADD3) and the FMA pipe (IMAD / IDP) each take one warp instruction
    // every other cycle per scheduler (profiles/tools/pipe_probe.cu), and cv::resize's vertical pass written with shifts is
    // almost all ALU.  Here the horizontally interpolated value keeps its low four bits cleared, S' = S & ~15, instead of
    // being shifted (S >> 4), so that  (b * (S >> 4)) >> 16  ==  mulhi(b << 12, S')  — one IMAD.HI — and the final  >> 2  becomes  * 64  (IMAD) followed by the byte-1 pick of the packing PRMT.
    auto hrow = [&](int sy, uint32_t h[4]) {  // horizontal interpolation of source row sy
        uint32_t w0, w1, w2;
        asm volatile("ld.shared.u32 %0, [%3];\n\tld.shared.u32 %1, [%3+4];\n\tld.shared.u32 %2, [%3+8];" : "=r"(w0), "=r"(w1), "=r"(w2) : "r"(sBase + (uint32_t)(sy * RS_BOXW)));
        h[0] = __dp2a_lo(coef[0], __byte_perm(w0, w1, sel[0]), 0u) & 0xFFFFFFF0u;
        h[1] = __dp2a_lo(coef[1], __byte_perm(w0, w1, sel[1]), 0u) & 0xFFFFFFF0u;
        h[2] = __dp2a_lo(coef[2], __byte_perm(w0, w1, sel[2]), 0u) & 0xFFFFFFF0u;
        h[3] = __dp2a_lo(coef[3], __byte_perm(hi3w ? w1 : w0, hi3w ? w2 : w1, sel[3]), 0u) & 0xFFFFFFF0u;
    };
    const uint32_t k64 = __shfl_sync(0xFFFFFFFFu, 64u + ((uint32_t)dw >> 30), threadIdx.x);  // 64, opaque to the compiler: keeps the multiply an IMAD (FMA pipe) instead of a LEA (ALU pipe)
    mbar_wait(&bars[wy], 0);
    // hX / hY: the two most recent horizontally interpolated source rows (rows rX, rY).  A destination row needs rows sy0 and
    // sy0 + 1; it usually finds sy0 in one of the two and loads sy0 + 1 over the other, so the roles alternate and no
    // register-to-register copies are needed — the vertical pass exists in both orientations.
    int rX = -1, rY = -1;
    uint32_t hX[4] = {0, 0, 0, 0}, hY[4] = {0, 0, 0, 0};
    const int nrows = ye - ys;
    auto vertical = [&](const uint32_t (&A)[4], const uint32_t (&B)[4], uint32_t b0, uint32_t b1) -> uint32_t {
        uint32_t u[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) u[i] = (__umulhi(b1, B[i]) + __umulhi(b0, A[i])) * k64 + 128u;  // (A + B + 2) * 64: byte 1 = the pixel (<= 255 by construction)
        return __byte_perm(__byte_perm(u[0], u[1], 0x0051), __byte_perm(u[2], u[3], 0x0051), 0x5410);
    };
    for (int k = 0; k < nrows; ++k) {
        const uint4 rt = rowTab[wy][k];
        const int sy0 = (int)rt.x, sy1 = (int)rt.y;  // (sy1 == sy0 only where the source row index is clamped at the bottom: the row is then interpolated twice)
        uint32_t out;
        if (sy0 == rY) {
            if (sy1 != rX) { hrow(sy1, hX); rX = sy1; }
            out = vertical(hY, hX, rt.z, rt.w);
        } else {
            if (sy0 != rX) { hrow(sy0, hX); rX = sy0; }
            if (sy1 != rY) { hrow(sy1, hY); rY = sy1; }
            out = vertical(hX, hY, rt.z, rt.w);
        }
        // lanes past the last column store into the row's right padding (rewritten by k_margin / never read)
        *reinterpret_cast<uint32_t*>(D) = out;
        D += dpitch;
    }
}

// tensor maps of the SOURCE levels (level l-1 feeds level l)
int resize_encode_maps(const Geom& g, uint8_t* pyr, int nImages, LevelMaps* out) {
    for (int l = 0; l + 1 < g.nlevels; ++l) {
        const LevelGeom& L = g.lv[l];
        const size_t blockOff = (size_t)L.planeOff - (size_t)kMarginY * L.pitch - kMarginX;
        if (tma_encode_3d(&out->m[l], pyr + blockOff, L.pitch, L.h + 2 * kMarginY, nImages, L.planeBytes, RS_BOXW, RS_BOXH)) return -1;
    }
    return 0;
}

void launch_pyramid_level(const Geom& g, const Buffers& b, const LevelMaps& maps, int level, const ResizeTab& tab, int nImages, cudaStream_t s) {
    const LevelGeom& S = g.lv[level - 1];
    const LevelGeom& D = g.lv[level];
    uint8_t* dp = b.pyr + D.planeOff + (size_t)b.imgBase * D.planeBytes;
    // the box must cover the source footprint of 128 destination columns x 16 rows: scale <= 1.25 in both directions
    if ((double)S.w <= 1.25 * D.w && (double)S.h <= 1.25 * D.h && S.w >= 16) {
        dim3 grid((D.w + 127) / 128, (D.h + RS_SEG * RS_WARPS - 1) / (RS_SEG * RS_WARPS), nImages), block(32, RS_WARPS);
        k_resize_s12<<<grid, block, 0, s>>>(maps.m[level - 1], S.h, b.imgBase, dp, D.w, D.h, D.pitch, (size_t)D.planeBytes, tab);
    } else {
        const uint8_t* sp = b.pyr + S.planeOff + (size_t)b.imgBase * S.planeBytes;
        dim3 grid((D.w + 127) / 128, (D.h + 7) / 8, nImages), block(32, 8);
        k_resize<<<grid, block, 0, s>>>(sp, S.w, S.h, S.pitch, (size_t)S.planeBytes, dp, D.w, D.h, D.pitch, (size_t)D.planeBytes, tab);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Grayscale conversion of the caller: Tracking::GrabImageStereoLidar (Tracking.cc:1646-1671, repeated in the other Grab*
// entry points) runs cvtColor(COLOR_{RGB,BGR}[A]2GRAY) on both images before it builds the Frame.  OpenCV 4.x, 8-bit:
// gray = (R * 9798 + G * 19235 + B * 3735 + 2^14) >> 15.  Here the interleaved colour image is uploaded as it is and converted
// straight into level 0 of the pyramid: four pixels per thread (four aligned words; for three channels the 12 bytes are cut out
// of them with funnel shifts), two IDP.2A per pixel, one 32-bit store.  src: dense [image][h][w * C].
// ---------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(128) k_gray(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int w, int h, int pitch, size_t plane, uint32_t k01,
                                              uint32_t k2) {
    const int x0 = 4 * (blockIdx.x * 128 + threadIdx.x);
    if (x0 >= w) return;
    const int y = blockIdx.y, img = blockIdx.z;
    const size_t off = (((size_t)img * h + y) * (size_t)w + (size_t)x0) * C;
    uint32_t p[4];  // p[i]: bytes c0, c1, c2 of pixel x0 + i in bytes 0..2
    if (C == 4) {
        const uint32_t* a = reinterpret_cast<const uint32_t*>(src + off);  // rows of odd widths are only 4-byte aligned
        p[0] = a[0]; p[1] = a[1]; p[2] = a[2]; p[3] = a[3];
    } else {
        const uint32_t* a = reinterpret_cast<const uint32_t*>(src + (off & ~(size_t)3));  // (the staging buffer has 64 spare bytes at its end)
        const int sh = (int)(off & 3) * 8;
        const uint32_t w0 = a[0], w1 = a[1], w2 = a[2], w3 = a[3];
        const uint32_t q0 = __funnelshift_r(w0, w1, sh), q1 = __funnelshift_r(w1, w2, sh), q2 = __funnelshift_r(w2, w3, sh);
        p[0] = q0; p[1] = __funnelshift_r(q0, q1, 24); p[2] = __funnelshift_r(q1, q2, 16); p[3] = q2 >> 8;
    }
    uint32_t g4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) g4[i] = __dp2a_hi(k2, p[i], __dp2a_lo(k01, p[i], 1u << 14)) >> 15;  // <= 255: the weights sum to 2^15
    // pixels past the last column land in the row's right padding (rewritten by k_margin / never read)
    *reinterpret_cast<uint32_t*>(dst + (size_t)img * plane + (size_t)y * pitch + x0) = g4[0] | (g4[1] << 8) | (g4[2] << 16) | (g4[3] << 24);
}

void launch_gray(const uint8_t* src, int channels, int rgb, uint8_t* l0, int w, int h, int pitch, size_t planeBytes, int nImages, cudaStream_t s) {
    const uint32_t RY = 9798, GY = 19235, BY = 3735;
    const uint32_t k01 = (rgb ? RY : BY) | (GY << 16), k2 = rgb ? BY : RY;  // 16x2 weights of bytes (0, 1) and (2, 3): alpha gets 0
    dim3 grid((w + 511) / 512, h, nImages);
    if (channels == 4)
        k_gray<4><<<grid, 128, 0, s>>>(src, l0, w, h, pitch, planeBytes, k01, k2);
    else
        k_gray<3><<<grid, 128, 0, s>>>(src, l0, w, h, pitch, planeBytes, k01, k2);
}

// ---------------------------------------------------------------------------------------------------------------
// Rectification of the caller: System::TrackStereoLidar (System.cc:238-248) runs cv::remap(im, imToFeed, M1, M2,
// cv::INTER_LINEAR) with the CV_32FC1 maps of cv::initUndistortRectifyMap (Settings.cc:523-526) before Grab* converts to grey.
// OpenCV's fixed-point path: sx = cvRound(mapx * 32), taps at (sx >> 5, sy >> 5) and its right / lower neighbours, weights
// 32 * {(32-fx)(32-fy), fx(32-fy), (32-fx)fy, fx*fy}, dst = (sum + 2^14) >> 15, taps outside the source are 0.
// Here: the raw image is uploaded as it is; the maps stay resident; one thread produces four destination pixels (gathers
// through L1 / L2: the taps of neighbouring pixels share cache lines); colour images are remapped per channel and then
// converted exactly as k_gray does, i.e. in the reference's order remap -> cvtColor.  src: dense [image][sh][sw * C].
// ---------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(128) k_remap(const uint8_t* __restrict__ src, int sw, int sh, const float* __restrict__ mapx, const float* __restrict__ mapy,
                                               uint8_t* __restrict__ dst, int dw, int dh, int pitch, size_t plane, uint32_t k01, uint32_t k2) {
    const int x0 = 4 * (blockIdx.x * 128 + threadIdx.x);
    if (x0 >= dw) return;
    const int y = blockIdx.y, img = blockIdx.z;
    const uint8_t* S = src + (size_t)img * sh * (size_t)sw * C;
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = min(x0 + i, dw - 1);
        const size_t m = (size_t)y * dw + x;
        const int sxq = __float2int_rn(__fmul_rn(__ldg(mapx + m), 32.0f)), syq = __float2int_rn(__fmul_rn(__ldg(mapy + m), 32.0f));
        const int ix = min(max(sxq >> 5, -32768), 32767), iy = min(max(syq >> 5, -32768), 32767);
        const int fx = sxq & 31, fy = syq & 31;
        const int w00 = 32 * (32 - fx) * (32 - fy), w01 = 32 * fx * (32 - fy), w10 = 32 * (32 - fx) * fy, w11 = 32 * fx * fy;
        const bool inx0 = ix >= 0 && ix < sw, inx1 = ix + 1 >= 0 && ix + 1 < sw, iny0 = iy >= 0 && iy < sh, iny1 = iy + 1 >= 0 && iy + 1 < sh;
        const uint8_t* p = S + ((ptrdiff_t)iy * sw + ix) * C;
        uint32_t px = 0;  // the remapped channels in bytes 0..2
#pragma unroll
        for (int c = 0; c < (C == 1 ? 1 : 3); ++c) {
            const int t00 = (iny0 && inx0) ? (int)__ldg(p + c) : 0, t01 = (iny0 && inx1) ? (int)__ldg(p + C + c) : 0;
            const int t10 = (iny1 && inx0) ? (int)__ldg(p + (ptrdiff_t)sw * C + c) : 0, t11 = (iny1 && inx1) ? (int)__ldg(p + (ptrdiff_t)sw * C + C + c) : 0;
            px |= (uint32_t)((t00 * w00 + t01 * w01 + t10 * w10 + t11 * w11 + (1 << 14)) >> 15) << (8 * c);
        }
        const uint32_t g = C == 1 ? px : (__dp2a_hi(k2, px, __dp2a_lo(k01, px, 1u << 14)) >> 15);
        out |= g << (8 * i);
    }
    // pixels past the last column land in the row's right padding (rewritten by k_margin / never read)
    *reinterpret_cast<uint32_t*>(dst + (size_t)img * plane + (size_t)y * pitch + x0) = out;
}

void launch_remap(const uint8_t* src, int sw, int sh, int channels, int rgb, const float* mapx, const float* mapy, uint8_t* l0, int w, int h, int pitch,
                  size_t planeBytes, int nImages, cudaStream_t s) {
    const uint32_t RY = 9798, GY = 19235, BY = 3735;
    const uint32_t k01 = (rgb ? RY : BY) | (GY << 16), k2 = rgb ? BY : RY;
    dim3 grid((w + 511) / 512, h, nImages);
    if (channels == 1)
        k_remap<1><<<grid, 128, 0, s>>>(src, sw, sh, mapx, mapy, l0, w, h, pitch, planeBytes, k01, k2);
    else if (channels == 3)
        k_remap<3><<<grid, 128, 0, s>>>(src, sw, sh, mapx, mapy, l0, w, h, pitch, planeBytes, k01, k2);
    else
        k_remap<4><<<grid, 128, 0, s>>>(src, sw, sh, mapx, mapy, l0, w, h, pitch, planeBytes, k01, k2);
}

// ---------------------------------------------------------------------------------------------------------------
// Input resize of the caller: System::TrackStereoLidar (System.cc:250-252) runs cv::resize(im, imToFeed, settings_->newImSize())
// — INTER_LINEAR, the same 8-bit fixed-point arithmetic as the pyramid's resize, per channel — when
// Settings::needToResize(), before Grab* converts to grey.  Here the raw image is uploaded as it is and resized (+ converted exactly
// as k_gray does, i.e. in the reference's order resize -> cvtColor) straight into level 0.  One thread = four destination pixels.
// src: dense [image][sh][sw * C]; tables as for the pyramid (make_resize_tables for sw -> dw, sh -> dh).
// ---------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(128) k_input_resize(const uint8_t* __restrict__ src, int sw, int sh, ResizeTab t, uint8_t* __restrict__ dst, int dw, int dh,
                                                      int pitch, size_t plane, uint32_t k01, uint32_t k2) {
    const int x0 = 4 * (blockIdx.x * 128 + threadIdx.x);
    if (x0 >= dw) return;
    const int y = blockIdx.y, img = blockIdx.z;
    const uint8_t* S = src + (size_t)img * sh * (size_t)sw * C;
    const int yo = t.yofs[y];
    const int sy0 = min(max(yo, 0), sh - 1), sy1 = min(max(yo + 1, 0), sh - 1);
    const int b0 = t.yb[2 * y], b1 = t.yb[2 * y + 1];
    const uint8_t* __restrict__ r0 = S + (size_t)sy0 * sw * C;
    const uint8_t* __restrict__ r1 = S + (size_t)sy1 * sw * C;
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = min(x0 + i, dw - 1);
        const int sx = t.xofs[x], sx1 = min(sx + 1, sw - 1);
        const int a0 = t.xa[2 * x], a1 = t.xa[2 * x + 1];
        uint32_t px = 0;  // the resized channels in bytes 0..2
#pragma unroll
        for (int c = 0; c < (C == 1 ? 1 : 3); ++c) {
            const int h0 = (int)__ldg(r0 + sx * C + c) * a0 + (int)__ldg(r0 + sx1 * C + c) * a1;
            const int h1 = (int)__ldg(r1 + sx * C + c) * a0 + (int)__ldg(r1 + sx1 * C + c) * a1;
            const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            px |= (uint32_t)min(max(v, 0), 255) << (8 * c);
        }
        const uint32_t g = C == 1 ? px : (__dp2a_hi(k2, px, __dp2a_lo(k01, px, 1u << 14)) >> 15);
        out |= g << (8 * i);
    }
    // pixels past the last column land in the row's right padding (never read)
    *reinterpret_cast<uint32_t*>(dst + (size_t)img * plane + (size_t)y * pitch + x0) = out;
}

void launch_input_resize(const uint8_t* src, int sw, int sh, int channels, int rgb, const ResizeTab& tab, uint8_t* l0, int w, int h, int pitch, size_t planeBytes,
                         int nImages, cudaStream_t s) {
    const uint32_t RY = 9798, GY = 19235, BY = 3735;
    const uint32_t k01 = (rgb ? RY : BY) | (GY << 16), k2 = rgb ? BY : RY;
    dim3 grid((w + 511) / 512, h, nImages);
    if (channels == 1)
        k_input_resize<1><<<grid, 128, 0, s>>>(src, sw, sh, tab, l0, w, h, pitch, planeBytes, k01, k2);
    else if (channels == 3)
        k_input_resize<3><<<grid, 128, 0, s>>>(src, sw, sh, tab, l0, w, h, pitch, planeBytes, k01, k2);
    else
        k_input_resize<4><<<grid, 128, 0, s>>>(src, sw, sh, tab, l0, w, h, pitch, planeBytes, k01, k2);
}

// ---------------------------------------------------------------------------------------------------------------
// Reflect-101 margin ROWS of every pyramid plane (3 above, 3 below the interior): what GaussianBlur(..., BORDER_REFLECT_101)
// reads beyond the top / bottom edge of its (cloned, hence border-less) input (ORBextractor.cc:1105-1106).  They are plain row
// copies (contiguous, 6 x w bytes per plane).  The margin COLUMNS are not materialised in HBM: touching both ends of every
// row costs two 64-byte DRAM accesses per row for 6 bytes (ncu: 0.47 MB per KITTI image, 1.1 ms per 8192 images); the blur
// kernel writes the three mirrored bytes into its shared-memory box instead.  One launch covers all levels of all images.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return min(max(i, 0), n - 1);
}

__global__ void __launch_bounds__(256) k_margin(const __grid_constant__ Geom g, const Buffers b) {
    const int level = blockIdx.y, img = b.imgBase + blockIdx.z;
    const LevelGeom& L = g.lv[level];
    const int t = blockIdx.x * 256 + threadIdx.x;
    const int wWords = (L.w + 3) >> 2;
    const int rowItems = 6 * wWords;
    uint8_t* P = b.pyr + L.planeOff + (size_t)img * L.planeBytes;
    if (t < rowItems) {
        // the three rows above and below: word copies of the mirrored interior rows (-k -> k, h-1+k -> h-1-k)
        const int r = t / wWords, c = t - r * wWords;
        const int y = r < 3 ? r - 3 : L.h + (r - 3);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(P + (size_t)reflect101(y, L.h) * L.pitch);
        if (4 * c + 4 <= L.w)
            reinterpret_cast<uint32_t*>(P + (ptrdiff_t)y * L.pitch)[c] = src[c];
        else  // the last, partial word
            for (int x = 4 * c; x < L.w; ++x) (P + (ptrdiff_t)y * L.pitch)[x] = reinterpret_cast<const uint8_t*>(src)[x];
    }
}

void launch_margin(const Geom& g, const Buffers& b, int nImages, cudaStream_t s) {
    int maxItems = 0;
    for (int l = 0; l < g.nlevels; ++l) maxItems = max(maxItems, 6 * ((g.lv[l].w + 3) / 4));
    dim3 grid((maxItems + 255) / 256, g.nlevels, nImages);
    k_margin<<<grid, 256, 0, s>>>(g, b);
}

// ---------------------------------------------------------------------------------------------------------------
// 7x7 sigma=2 Gaussian blur (GaussianBlur on a clone, ORBextractor.cc:1105-1106; arithmetic SURVEY.md A.3).
// OpenCV's 8-bit path is a separable 8.8 fixed-point filter whose row pass cannot saturate (taps sum to 256 or 257), so the
// result is the exact integer 2-D convolution followed by one rounding, (v + 32768) >> 16.  Here: a warp owns a 128-pixel
// column strip of one image (a thread 4 columns) and slides down the rows; its input arrives as double-buffered TMA boxes (see
// k_blur).  Row pass: the 7 taps of a pixel are two 4-byte dot products (IDP.4A) on byte windows cut out of three aligned
// 32-bit shared-memory loads with PRMT.  Column pass: two input rows at a time as 16x2 lanes, one IDP.2A per touched output
// row against the tap pairs, six running sums per pixel carried across boxes; no inter-thread exchange; a warp writes 128
// contiguous bytes per row.  BORDER_REFLECT_101: the three rows above / below the image exist in memory (k_margin); the three
// columns left / right of it are mirrored inside the shared-memory box by the warps that own an edge strip (lane = box row).
// ---------------------------------------------------------------------------------------------------------------
#ifndef BLUR_NWARPS
#define BLUR_NWARPS 4
#endif
constexpr int BLUR_WARPS = BLUR_NWARPS;

constexpr int BLUR_BOXW = 144;                  // 4 halo bytes + 128 pixels + 4 halo bytes, rounded to a multiple of 16
constexpr int BLUR_BOXH = 32;                   // input rows per TMA box (= lanes: the edge mirroring handles one box row per lane)
constexpr int BLUR_BOX_BYTES = BLUR_BOXW * BLUR_BOXH;  // 4608 = 36 x 128

struct BlurTaps { uint32_t cA, cB, c67, c56, c45, c34, c23, c12, c01, c0x; };

// one warp: output rows [rowBeg, rowBeg + nRows) of the 128-pixel column strip `strip` of image `imgAbs` (dst = that image's pixel (0, 0) of the blurred level)
template <bool SAT>
__device__ __forceinline__ void blur_strip(const CUtensorMap* srcMap, uint8_t* __restrict__ dst, int w, int pitch, int strip, int imgAbs, int rowBeg, int nRows,
                                           unsigned char (*tileW)[BLUR_BOX_BYTES], uint64_t* barW, uint32_t cA, uint32_t cB, uint32_t c67, uint32_t c56, uint32_t c45,
                                           uint32_t c34, uint32_t c23, uint32_t c12, uint32_t c01, uint32_t c0x) {
    // A warp owns one 128-pixel column strip of one image and walks it top to bottom.  Its input arrives as TMA boxes of
    // 144 bytes x 32 rows, double buffered: the next box is in flight while the current one is filtered, so ~9 KB per warp
    // are outstanding (three 128-byte loads per warp would leave HBM two thirds idle at this occupancy), and the running
    // column sums carry across boxes, so there is no per-segment warm-up.
    const int x0 = (strip * 32 + threadIdx.x) * 4;
    const int nChunks = (nRows + 6 + BLUR_BOXH - 1) / BLUR_BOXH;
    auto issue = [&](int c) {  // box origin: byte kMarginX + x - 4 = 128 * strip (16-byte aligned), plane row kMarginY + (rowBeg + 32c - 3) = rowBeg + 32c
        mbar_arrive_expect_tx(&barW[c & 1], BLUR_BOX_BYTES);
        tma_load_3d(tileW[c & 1], srcMap, &barW[c & 1], strip * 128, rowBeg + c * BLUR_BOXH, imgAbs);
    };
    if (threadIdx.x == 0) {
        issue(0);
        if (nChunks > 1) issue(1);
    }
    const bool active = x0 < w;
    // BORDER_REFLECT_101 columns: pixels -3..-1 (first strip) and w..w+2 (every strip whose box reaches past column w - 1) are
    // written into the box in shared memory; eR = box byte of pixel w (the box starts 4 bytes before the strip's first pixel)
    const bool edgeL = strip == 0;
    const int eR = 4 + w - strip * 128;
    const bool edgeR = eR < BLUR_BOXW - 8;  // the filter reads box bytes 1 .. 134
    uint8_t* D = dst + (ptrdiff_t)rowBeg * pitch + x0;
    uint32_t ph0 = 0, ph1 = 0;
    // Column pass two input rows at a time: P = (h[y], h[y+1]) as 16x2 lanes, and the pair's contribution to the eight output
    // rows y-3 .. y+4 is one IDP.2A each against the tap pairs (t6,0) (t5,t6) (t4,t5) (t3,t4) (t2,t3) (t1,t2) (t0,t1) (0,t0).
    // a[i][0..5] = running sums of output rows y-3 .. y+2 before the pair (y, y+1) is added.
    uint32_t a[4][6];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 6; ++k) a[i][k] = 0;  // (the rows fed by these initial sums lie above the image and are not stored)
    auto hpass = [&](const uint32_t wd[3], uint32_t hr[4]) {  // row pass of one input row
        const uint32_t prev = wd[0], cur = wd[1], next = wd[2];
        hr[0] = __dp4a(__byte_perm(prev, cur, 0x4321), cA, __dp4a(__byte_perm(cur, next, 0x4321), cB, 0u));
        hr[1] = __dp4a(__byte_perm(prev, cur, 0x5432), cA, __dp4a(__byte_perm(cur, next, 0x5432), cB, 0u));
        hr[2] = __dp4a(__byte_perm(prev, cur, 0x6543), cA, __dp4a(__byte_perm(cur, next, 0x6543), cB, 0u));
        hr[3] = __dp4a(cur, cA, __dp4a(next, cB, 0u));
    };
    // The rounding term of (v + 32768) >> 16 rides in the two accumulators a chain starts from; the result is then byte 2 of
    // the sum (PRMT picks it; with taps summing to 257 the sum may exceed 255 << 16 and is clamped first).  Drow: the output
    // row y - 3 of the pair being added, advanced by two rows per pair.
    const uint32_t RND = 32768u;
    uint8_t* Drow = D - (ptrdiff_t)6 * pitch;  // the first pair (box row 0 = image row -3) feeds output rows -6, -5
    auto pair = [&](int y, const uint32_t wa[3], const uint32_t wb[3]) {  // input rows y, y+1 -> output rows y-3, y-2
        uint32_t hA[4], hB[4];
        hpass(wa, hA);
        hpass(wb, hB);
        uint32_t o0[4], o1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t P = __byte_perm(hA[i], hB[i], 0x5410);  // both <= 65535 (taps sum to <= 257)
            o0[i] = __dp2a_lo(P, c67, a[i][0]);
            o1[i] = __dp2a_lo(P, c56, a[i][1]);
            a[i][0] = __dp2a_lo(P, c45, a[i][2]);
            a[i][1] = __dp2a_lo(P, c34, a[i][3]);
            a[i][2] = __dp2a_lo(P, c23, a[i][4]);
            a[i][3] = __dp2a_lo(P, c12, a[i][5]);
            a[i][4] = __dp2a_lo(P, c01, RND);
            a[i][5] = __dp2a_lo(P, c0x, RND);
            if (SAT) { o0[i] = min(o0[i], 0x00FFFFFFu); o1[i] = min(o1[i], 0x00FFFFFFu); }
        }
        const uint32_t out0 = __byte_perm(__byte_perm(o0[0], o0[1], 0x0062), __byte_perm(o0[2], o0[3], 0x0062), 0x5410);
        const uint32_t out1 = __byte_perm(__byte_perm(o1[0], o1[1], 0x0062), __byte_perm(o1[2], o1[3], 0x0062), 0x5410);
        if ((unsigned)(y - 3) < (unsigned)nRows) *reinterpret_cast<uint32_t*>(Drow) = out0;   // y: input row relative to rowBeg
        if ((unsigned)(y - 2) < (unsigned)nRows) *reinterpret_cast<uint32_t*>(Drow + pitch) = out1;
        Drow += 2 * (ptrdiff_t)pitch;
    };
    for (int c = 0; c < nChunks; ++c) {
        if (c & 1) { mbar_wait(&barW[1], ph1); ph1 ^= 1; }
        else { mbar_wait(&barW[0], ph0); ph0 ^= 1; }
        if (edgeL | edgeR) {  // warp-uniform: mirror the three pixels beyond the left / right image edge inside the box, lane = box row
            unsigned char* row = tileW[c & 1] + threadIdx.x * BLUR_BOXW;
            if (edgeL) reinterpret_cast<uint32_t*>(row)[0] = __byte_perm(reinterpret_cast<const uint32_t*>(row)[1], 0, 0x1230);  // [pad, p3, p2, p1] in front of p0
            if (edgeR) {
#pragma unroll
                for (int m = 0; m < 3; ++m)
                    if (eR + m < BLUR_BOXW) row[eR + m] = row[eR - 2 - m];  // pixel w + m <- pixel w - 2 - m
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes before the next TMA box lands in this buffer
            __syncwarp();
        }
        if (active) {
            const uint32_t* T = reinterpret_cast<const uint32_t*>(tileW[c & 1]) + threadIdx.x;  // word 0 = pixels x0-4..x0-1
            const int yb = c * BLUR_BOXH - 3;  // row of box row 0, relative to rowBeg
#pragma unroll 2
            for (int k = 0; k < BLUR_BOXH; k += 2) {
                const uint32_t* r0 = T + k * (BLUR_BOXW / 4);
                const uint32_t wa[3] = {r0[0], r0[1], r0[2]}, wb[3] = {r0[BLUR_BOXW / 4], r0[BLUR_BOXW / 4 + 1], r0[BLUR_BOXW / 4 + 2]};
                pair(yb + k, wa, wb);
            }
        }
        __syncwarp();  // every lane is done with this buffer before lane 0 re-arms it
        if (threadIdx.x == 0 && c + 2 < nChunks) issue(c + 2);
    }
}


template <bool SAT>
__global__ void __launch_bounds__(32 * BLUR_WARPS) k_blur(const __grid_constant__ CUtensorMap srcMap, uint8_t* __restrict__ dst, int w, int h, int pitch,
                                                         size_t plane, int imgBase, int nStrips, int nWarps, const __grid_constant__ BlurTaps t) {
    __shared__ __align__(128) unsigned char tile[BLUR_WARPS][2][BLUR_BOX_BYTES];
    __shared__ __align__(8) uint64_t bars[BLUR_WARPS][2];
    const int wy = threadIdx.y;
    if (threadIdx.x == 0) { mbar_init(&bars[wy][0], 1); mbar_init(&bars[wy][1], 1); }
    mbar_fence_init();
    __syncwarp();
    const int gw = blockIdx.x * BLUR_WARPS + wy;
    if (gw >= nWarps) return;   // whole warp
    const int imgRel = gw / nStrips, strip = gw - imgRel * nStrips;
    blur_strip<SAT>(&srcMap, dst + (size_t)imgRel * plane, w, pitch, strip, imgBase + imgRel, 0, h, tile[wy], bars[wy], t.cA, t.cB, t.c67, t.c56, t.c45, t.c34, t.c23,
                    t.c12, t.c01, t.c0x);
}

// All levels of a FEW images in one launch (the single-image handle: Frame's two extractor calls): eight dependent launches of a handful of
// warps each are pure launch latency, and a warp walking a whole strip top to bottom is serial.  Here a warp = (level, image, strip, row segment);
// a segment re-reads the six rows of filter support above / below it, nothing else changes.
struct BlurAllLevels {
    CUtensorMap m[ORBX_MAX_LEVELS];
    unsigned long long dstOff[ORBX_MAX_LEVELS], plane[ORBX_MAX_LEVELS];
    int w[ORBX_MAX_LEVELS], h[ORBX_MAX_LEVELS], pitch[ORBX_MAX_LEVELS], nStrips[ORBX_MAX_LEVELS], nSegs[ORBX_MAX_LEVELS], segRows[ORBX_MAX_LEVELS];
    int warpEnd[ORBX_MAX_LEVELS];   // running total of warps up to and including the level
    int nlevels, imgBase;
};

template <bool SAT>
__global__ void __launch_bounds__(32 * BLUR_WARPS) k_blur_all(const __grid_constant__ BlurAllLevels lv, uint8_t* __restrict__ blur, const __grid_constant__ BlurTaps t) {
    __shared__ __align__(128) unsigned char tile[BLUR_WARPS][2][BLUR_BOX_BYTES];
    __shared__ __align__(8) uint64_t bars[BLUR_WARPS][2];
    const int wy = threadIdx.y;
    if (threadIdx.x == 0) { mbar_init(&bars[wy][0], 1); mbar_init(&bars[wy][1], 1); }
    mbar_fence_init();
    __syncwarp();
    int gw = blockIdx.x * BLUR_WARPS + wy;
    if (gw >= lv.warpEnd[lv.nlevels - 1]) return;   // whole warp
    int level = 0;
    while (gw >= lv.warpEnd[level]) ++level;
    if (level) gw -= lv.warpEnd[level - 1];
    const int perImage = lv.nStrips[level] * lv.nSegs[level];
    const int imgRel = gw / perImage, rem = gw - imgRel * perImage;
    const int seg = rem / lv.nStrips[level], strip = rem - seg * lv.nStrips[level];
    const int rowBeg = seg * lv.segRows[level], nRows = min(lv.segRows[level], lv.h[level] - rowBeg);
    blur_strip<SAT>(&lv.m[level], blur + lv.dstOff[level] + (size_t)imgRel * lv.plane[level], lv.w[level], lv.pitch[level], strip, lv.imgBase + imgRel, rowBeg, nRows,
                    tile[wy], bars[wy], t.cA, t.cB, t.c67, t.c56, t.c45, t.c34, t.c23, t.c12, t.c01, t.c0x);
}

static BlurTaps make_blur_taps(const int taps[7], int* sum) {
    BlurTaps t;
    t.cA = (uint32_t)taps[0] | ((uint32_t)taps[1] << 8) | ((uint32_t)taps[2] << 16) | ((uint32_t)taps[3] << 24);
    t.cB = (uint32_t)taps[4] | ((uint32_t)taps[5] << 8) | ((uint32_t)taps[6] << 16);
    *sum = 0;
    for (int i = 0; i < 7; ++i) *sum += taps[i];
    auto pr = [&](int lo, int hi) { return (uint32_t)(lo < 0 || lo > 6 ? 0 : taps[lo]) | ((uint32_t)(hi < 0 || hi > 6 ? 0 : taps[hi]) << 8); };
    t.c67 = pr(6, 7); t.c56 = pr(5, 6); t.c45 = pr(4, 5); t.c34 = pr(3, 4); t.c23 = pr(2, 3); t.c12 = pr(1, 2); t.c01 = pr(0, 1); t.c0x = pr(-1, 0);
    return t;
}

void launch_blur(const Geom& g, const Buffers& b, const LevelMaps& maps, int level, int nImages, const int taps[7], cudaStream_t s) {
    const LevelGeom& L = g.lv[level];
    int sum;
    const BlurTaps t = make_blur_taps(taps, &sum);
    const int nStrips = (L.w + 127) / 128, nWarps = nStrips * nImages;
    dim3 grid((nWarps + BLUR_WARPS - 1) / BLUR_WARPS), block(32, BLUR_WARPS);
    const size_t o = (size_t)L.planeOff + (size_t)b.imgBase * L.planeBytes;
    if (sum > 256)
        k_blur<true><<<grid, block, 0, s>>>(maps.m[level], b.blur + o, L.w, L.h, L.pitch, (size_t)L.planeBytes, b.imgBase, nStrips, nWarps, t);
    else
        k_blur<false><<<grid, block, 0, s>>>(maps.m[level], b.blur + o, L.w, L.h, L.pitch, (size_t)L.planeBytes, b.imgBase, nStrips, nWarps, t);
}

// every level of images [imgBase, imgBase + nImages) in ONE launch; segRows = output rows per warp (a multiple of 2)
void launch_blur_all(const Geom& g, const Buffers& b, const LevelMaps& maps, int nImages, const int taps[7], int segRows, cudaStream_t s) {
    int sum;
    const BlurTaps t = make_blur_taps(taps, &sum);
    BlurAllLevels lv;
    int total = 0;
    for (int l = 0; l < g.nlevels; ++l) {
        const LevelGeom& L = g.lv[l];
        lv.m[l] = maps.m[l];
        lv.dstOff[l] = L.planeOff + (unsigned long long)b.imgBase * L.planeBytes;
        lv.plane[l] = L.planeBytes;
        lv.w[l] = L.w; lv.h[l] = L.h; lv.pitch[l] = L.pitch;
        lv.nStrips[l] = (L.w + 127) / 128;
        lv.segRows[l] = segRows;
        lv.nSegs[l] = (L.h + segRows - 1) / segRows;
        total += lv.nStrips[l] * lv.nSegs[l] * nImages;
        lv.warpEnd[l] = total;
    }
    for (int l = g.nlevels; l < ORBX_MAX_LEVELS; ++l) lv.warpEnd[l] = total;
    lv.nlevels = g.nlevels; lv.imgBase = b.imgBase;
    dim3 grid((total + BLUR_WARPS - 1) / BLUR_WARPS), block(32, BLUR_WARPS);
    if (sum > 256) k_blur_all<true><<<grid, block, 0, s>>>(lv, b.blur, t);
    else k_blur_all<false><<<grid, block, 0, s>>>(lv, b.blur, t);
}

int blur_encode_maps(const Geom& g, uint8_t* pyr, int nImages, LevelMaps* out) {
    for (int l = 0; l < g.nlevels; ++l) {
        const LevelGeom& L = g.lv[l];
        const size_t blockOff = (size_t)L.planeOff - (size_t)kMarginY * L.pitch - kMarginX;
        if (tma_encode_3d(&out->m[l], pyr + blockOff, L.pitch, L.h + 2 * kMarginY, nImages, L.planeBytes, BLUR_BOXW, BLUR_BOXH)) return -1;
    }
    return 0;
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
