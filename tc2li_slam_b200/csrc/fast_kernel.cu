// fast_kernel.cu — cell-grid FAST-9/16 with the iniThFAST -> minThFAST per-cell fallback.
//
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree (ORBextractor.cc:779-846) and the cv::FAST(cell, kps, th,
// true) calls inside it (ORBextractor.cc:800,819; arithmetic: SURVEY.md A.4).  One CTA = one cell of one pyramid level of
// one image; one launch covers every level of every image of the batch.
//
// Formulation (derived from the OpenCV semantics, not from its code): for a pixel with centre value v and ring r[0..15]
//     S = max( max_s min_{j<9} (r[s+j] - v),  max_s min_{j<9} (v - r[s+j]) )
// the pixel is a corner at threshold t iff S > t, cv::FAST reports response S-1 whatever t is, and 3x3 non-maximum
// suppression inside the cell crop keeps it iff S is strictly greater than the S of its 8 neighbours that lie in the crop's
// interior (3-px rim excluded).  Hence ONE score pass serves both thresholds: the cell uses iniThFAST when any survivor has
// S > iniThFAST, otherwise minThFAST.  Pixels that are not corners at minThFAST get S := 0, which cannot change any
// comparison.  Integer only; HBM traffic = one read of every level, the work is ALU/issue bound.
#include "orbx_internal.cuh"

namespace orbx {

namespace {

constexpr int TS_WORDS = 22;             // tile row stride in 32-bit words: (kMaxCell + 6 + 3 alignment slack) / 4 rounded up
constexpr int TS = TS_WORDS * 4;         // tile row stride in bytes
constexpr int TROWS = kMaxCell + 6;      // rows of the cell crop
constexpr int SS = kMaxCell + 4;         // score row stride (interior + 1-px zero rim, padded)
constexpr int MAXPIX = kMaxCell * kMaxCell;

__device__ __forceinline__ bool has_arc9(uint32_t m) {  // >= 9 contiguous set bits in a circular 16-bit mask
    m |= m << 16;
    m &= m >> 1;
    m &= m >> 2;
    m &= m >> 4;
    m &= m >> 1;
    return (m & 0xFFFFu) != 0;
}

// byte offsets of the 16 ring pixels inside the shared tile
#define RING(k) ((k) == 0 ? 3 * TS : (k) == 1 ? 3 * TS + 1 : (k) == 2 ? 2 * TS + 2 : (k) == 3 ? TS + 3 : (k) == 4 ? 3 : (k) == 5 ? -TS + 3 : \
                 (k) == 6 ? -2 * TS + 2 : (k) == 7 ? -3 * TS + 1 : (k) == 8 ? -3 * TS : (k) == 9 ? -3 * TS - 1 : (k) == 10 ? -2 * TS - 2 : \
                 (k) == 11 ? -TS - 3 : (k) == 12 ? -3 : (k) == 13 ? TS - 3 : (k) == 14 ? 2 * TS - 2 : 3 * TS - 1)

}  // namespace

__global__ void __launch_bounds__(256) k_fast_cells(const __grid_constant__ Geom g, const Buffers b) {
    __shared__ uint32_t tileW[TROWS * TS_WORDS];
    __shared__ uint8_t sc[(kMaxCell + 2) * SS];
    __shared__ uint32_t listBuf[MAXPIX / 2];  // uint16 corner list in passes A-C, packed candidates afterwards
    __shared__ uint16_t surv[MAXPIX / 4 + kMaxCell + 8];
    __shared__ int nList, nSurv, nOut, outBase;

    const uint8_t* tile = reinterpret_cast<const uint8_t*>(tileW);
    uint16_t* list = reinterpret_cast<uint16_t*>(listBuf);
    const int tid = threadIdx.x, img = b.imgBase + blockIdx.y;

    // which level / cell
    int level = 0;
    while (level + 1 < g.nlevels && (int)blockIdx.x >= g.lv[level + 1].cellOff) ++level;
    const LevelGeom& L = g.lv[level];
    const int ci = (int)blockIdx.x - L.cellOff;
    const int i = ci / L.nCols, j = ci - i * L.nCols;
    const int maxBX = L.w - kBorder, maxBY = L.h - kBorder;
    const int iniX = kBorder + j * L.wCell, iniY = kBorder + i * L.hCell;
    const int cw = min(iniX + L.wCell + 6, maxBX) - iniX, ch = min(iniY + L.hCell + 6, maxBY) - iniY;
    uint8_t* flag = b.cellFlags + (size_t)img * g.cellsPerImage + blockIdx.x;
    if (iniY >= maxBY - 3 || iniX >= maxBX - 6) {  // cells the reference skips (ORBextractor.cc:785-796)
        if (tid == 0) *flag = 0;
        return;
    }
    if (cw < 7 || ch < 7) {  // cv::FAST yields nothing on crops thinner than the ring, at either threshold
        if (tid == 0) *flag = 2;
        return;
    }
    const int iw = cw - 6, ih = ch - 6, npix = iw * ih;

    // ---- stage the crop: aligned 32-bit loads, one warp per row ----
    const uint8_t* lev = b.pyr + L.planeOff + (size_t)img * L.planeBytes;
    const int a0 = iniX & ~3, shift = iniX - a0;
    const int nWords = (iniX + cw - a0 + 3) >> 2;
    for (int r = tid >> 5; r < ch; r += 8) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(lev + (size_t)(iniY + r) * L.pitch + a0);
        for (int c = tid & 31; c < nWords; c += 32) tileW[r * TS_WORDS + c] = __ldg(src + c);
    }
    for (int k = tid; k < (ih + 2) * SS; k += 256) sc[k] = 0;
    if (tid == 0) { nList = 0; nSurv = 0; nOut = 0; }
    __syncthreads();

    // ---- pass A: which pixels are corners at minThFAST (cheap reject first) ----
    const int t = g.minTh;
    const uint32_t magic = (0xFFFFFFFFu / (uint32_t)iw) + 1;  // exact p / iw for p < 2^16
    for (int p = tid; p < npix; p += 256) {
        const int y = (int)__umulhi((uint32_t)p, magic), x = p - y * iw;
        const uint8_t* c = tile + (y + 3) * TS + x + 3 + shift;
        const int v = c[0], hi = v + t, lo = v - t;
        const int r0 = c[RING(0)], r8 = c[RING(8)];
        bool br = (r0 > hi) | (r8 > hi), dk = (r0 < lo) | (r8 < lo);
        if (!(br | dk)) continue;
        const int r4 = c[RING(4)], r12 = c[RING(12)];
        br &= (r4 > hi) | (r12 > hi);
        dk &= (r4 < lo) | (r12 < lo);
        if (!(br | dk)) continue;
        uint32_t mb = 0, md = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int r = c[RING(k)];
            mb |= (uint32_t)(r > hi) << k;
            md |= (uint32_t)(r < lo) << k;
        }
        const bool isB = br && has_arc9(mb), isD = dk && has_arc9(md);
        if (isB | isD) list[atomicAdd(&nList, 1)] = (uint16_t)(p | (isD ? 0x8000 : 0));
    }
    __syncthreads();

    // ---- pass B: arc score of the listed pixels (dense over the list: no divergence on the expensive part) ----
    const int nl = nList;
    for (int k = tid; k < nl; k += 256) {
        const int e = list[k], p = e & 0x7FFF;
        const int y = (int)__umulhi((uint32_t)p, magic), x = p - y * iw;
        const uint8_t* c = tile + (y + 3) * TS + x + 3 + shift;
        const int v = c[0];
        const int sgn = (e & 0x8000) ? -1 : 1;
        int d[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) d[q] = sgn * ((int)c[RING(q)] - v);
        int m2[16], m4[16], m8[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) m2[q] = min(d[q], d[(q + 1) & 15]);
#pragma unroll
        for (int q = 0; q < 16; ++q) m4[q] = min(m2[q], m2[(q + 2) & 15]);
#pragma unroll
        for (int q = 0; q < 16; ++q) m8[q] = min(m4[q], m4[(q + 4) & 15]);
        int s = 0;
#pragma unroll
        for (int q = 0; q < 16; ++q) s = max(s, min(m8[q], d[(q + 8) & 15]));
        sc[(y + 1) * SS + x + 1] = (uint8_t)s;  // minTh < s <= 255
    }
    __syncthreads();

    // ---- pass C: cell-local 3x3 non-maximum suppression ----
    int anyIni = 0;
    for (int k = tid; k < nl; k += 256) {
        const int p = list[k] & 0x7FFF;
        const int y = (int)__umulhi((uint32_t)p, magic), x = p - y * iw;
        const uint8_t* q = sc + (y + 1) * SS + x + 1;
        const int s = q[0];
        const bool keep = s > q[-1] && s > q[1] && s > q[-SS - 1] && s > q[-SS] && s > q[-SS + 1] && s > q[SS - 1] && s > q[SS] && s > q[SS + 1];
        if (keep) {
            surv[atomicAdd(&nSurv, 1)] = (uint16_t)p;
            anyIni |= (s > g.iniTh);
        }
    }
    anyIni = __syncthreads_or(anyIni);

    // ---- emit: survivors above the cell's threshold, coordinates relative to (16,16) like the reference (ORBextractor.cc:837-842) ----
    const int th = anyIni ? g.iniTh : g.minTh;
    const int ns = nSurv;
    for (int k = tid; k < ns; k += 256) {
        const int p = surv[k];
        const int y = (int)__umulhi((uint32_t)p, magic), x = p - y * iw;
        const int s = sc[(y + 1) * SS + x + 1];
        if (s > th) listBuf[atomicAdd(&nOut, 1)] = pack_cand(x + 3 + j * L.wCell, y + 3 + i * L.hCell, s - 1);
    }
    __syncthreads();
    const int no = nOut;
    if (tid == 0) {
        *flag = anyIni ? 1 : 2;
        outBase = no ? atomicAdd(b.candCount + img * ORBX_MAX_LEVELS + level, no) : 0;
    }
    __syncthreads();
    if (no) {
        const int base = outBase;
        if (base + no > L.candCap) {
            if (tid == 0) atomicExch(b.errFlag, 1);
            return;
        }
        uint32_t* out = b.cand + L.candOff + (size_t)img * L.candCap + base;
        for (int k = tid; k < no; k += 256) out[k] = listBuf[k];
    }
}

void launch_fast(const Geom& g, const Buffers& b, int nImages, cudaStream_t s) {
    dim3 grid(g.cellsPerImage, nImages);
    k_fast_cells<<<grid, 256, 0, s>>>(g, b);
}

}  // namespace orbx
