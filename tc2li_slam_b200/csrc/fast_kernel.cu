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
#include <algorithm>

#include "orbx_internal.cuh"

namespace orbx {

namespace {

constexpr int FAST_THREADS = 128;
constexpr int TSW = 23;                  // tile row stride in 32-bit words: 1 pad word + (kMaxCell + 6 + 3 alignment slack) / 4, odd
constexpr int TS = TSW * 4;              // tile row stride in bytes (92)
constexpr unsigned FULL = 0xFFFFFFFFu;

// >= 9 contiguous set bits in each circular 16-bit half of m (low half: bright mask, high half: dark mask)
__device__ __forceinline__ uint32_t arc9_halves(uint32_t m) {
    const uint32_t lo = m & 0xFFFFu, hi = m >> 16;
    uint32_t a = lo | (lo << 16), c = hi | (hi << 16);
    a &= a >> 1; c &= c >> 1;
    a &= a >> 2; c &= c >> 2;
    a &= a >> 4; c &= c >> 4;
    a &= a >> 1; c &= c >> 1;
    return (a & 0xFFFFu) | (c << 16);
}

// per byte: 0x80 where the byte of d exceeds t (0 <= t <= 254); kAdd = (0x7f - (t & 0x7f)) * 0x01010101
__device__ __forceinline__ uint32_t bytes_gt(uint32_t d, uint32_t kAdd, bool tHigh) {
    const uint32_t low = (d & 0x7f7f7f7fu) + kAdd;  // bit 7 of each byte: (d & 0x7f) > (t & 0x7f)
    return tHigh ? (low & d) : (low | d);           // t >= 128: also needs d >= 128; t < 128: d >= 128 suffices
}

// byte offsets of the 16 ring pixels inside the shared tile
#define RING(k) ((k) == 0 ? 3 * TS : (k) == 1 ? 3 * TS + 1 : (k) == 2 ? 2 * TS + 2 : (k) == 3 ? TS + 3 : (k) == 4 ? 3 : (k) == 5 ? -TS + 3 : \
                 (k) == 6 ? -2 * TS + 2 : (k) == 7 ? -3 * TS + 1 : (k) == 8 ? -3 * TS : (k) == 9 ? -3 * TS - 1 : (k) == 10 ? -2 * TS - 2 : \
                 (k) == 11 ? -TS - 3 : (k) == 12 ? -3 : (k) == 13 ? TS - 3 : (k) == 14 ? 2 * TS - 2 : 3 * TS - 1)

}  // namespace

// Passes (all over one cell crop staged in shared memory; pixels are addressed by their byte offset inside the tile, the
// score map shares the tile's layout, so no pass needs a division):
//   A  four pixels per thread, packed in one 32-bit word: |ring - centre| > minTh at the four compass points, combined as
//      (N|S)&(E|W) — a necessary condition for a 9-arc of either polarity (VABSDIFF4 + carry tricks);
//      survivors (~15 % of the pixels of a textured frame) are appended to a list with ballot-aggregated offsets;
//   A2 dense over that list: exact bright/dark masks of the 16 ring pixels (both polarities at once in 16x2 lanes: the sign
//      bits of hi - r and r - lo) and the contiguity test -> list of corners;
//   B  dense over the corners: arc score S;   C  cell-local 3x3 NMS, iniTh/minTh decision, packed output.
// Shared memory (dynamic): tile [rows][92] | score map [rows][92] | list1 [maxPix] u16 | list2 [maxPix] u16.
__global__ void __launch_bounds__(FAST_THREADS) k_fast_cells(const __grid_constant__ Geom g, const Buffers b, int tileRows, int maxPix) {
    extern __shared__ __align__(16) uint32_t fsm[];
    __shared__ int n1, n2, nSurv, nOut, outBase;
    uint32_t* tileW = fsm;
    uint32_t* scW = fsm + tileRows * TSW;
    uint16_t* list1 = reinterpret_cast<uint16_t*>(scW + tileRows * TSW);
    uint16_t* list2 = list1 + maxPix;

    const uint8_t* tile = reinterpret_cast<const uint8_t*>(tileW) + 4;  // skip the pad word: tile[r * TS + c] = crop byte (r, a0 + c)
    uint8_t* sc = reinterpret_cast<uint8_t*>(scW) + 4;                   // same layout as the tile
    const int tid = threadIdx.x, lane = tid & 31, img = b.imgBase + blockIdx.y;

    // cell descriptor, precomputed on the host (make_cell_table)
    const uint4 cd = __ldg(b.cellTab + blockIdx.x);
    const int iniX = cd.x & 0xFFFF, iniY = cd.x >> 16;
    const int cw = cd.y & 0xFF, ch = (cd.y >> 8) & 0xFF, level = (cd.y >> 16) & 0xFF, kind = cd.y >> 24;
    uint8_t* flag = b.cellFlags + (size_t)img * g.cellsPerImage + blockIdx.x;
    if (kind != 2) {  // 0: a cell the reference skips (ORBextractor.cc:785-796); 1: crop thinner than the ring, cv::FAST yields nothing
        if (tid == 0) *flag = kind == 0 ? 0 : 2;
        return;
    }
    const LevelGeom& L = g.lv[level];
    const int iw = cw - 6, ih = ch - 6;
    const int nW = cd.w & 0xFF, wFirst = (cd.w >> 8) & 0xFF, shift = (cd.w >> 16) & 0xFF;
    const uint32_t magicW = cd.z;

    // ---- stage the crop: aligned 32-bit loads; tile word (c + 1) of a row = global bytes a0 + 4c .. ----
    {
        const uint8_t* lev = b.pyr + L.planeOff + (size_t)img * L.planeBytes + (size_t)iniY * L.pitch + (iniX - shift);
        const int nWords = (shift + cw + 3) >> 2;
        const uint32_t magicL = (0xFFFFFFFFu / (uint32_t)nWords) + 1;
        const int nItems = ch * nWords;
        for (int e = tid; e < nItems; e += FAST_THREADS) {
            const int r = (int)__umulhi((uint32_t)e, magicL), c = e - r * nWords;
            tileW[r * TSW + 1 + c] = __ldg(reinterpret_cast<const uint32_t*>(lev + (size_t)r * L.pitch) + c);
        }
        uint4* z = reinterpret_cast<uint4*>(scW);
        const int nz = (ch * TSW + 3) >> 2;
        for (int k = tid; k < nz; k += FAST_THREADS) z[k] = make_uint4(0, 0, 0, 0);
        if (tid == 0) { n1 = 0; n2 = 0; nSurv = 0; nOut = 0; }
    }
    __syncthreads();

    // Threshold attempts (ORBextractor.cc:800-820): iniThFAST first; pixels that are not corners at iniThFAST cannot suppress
    // one that is, so the first attempt is exactly cv::FAST(cell, iniThFAST, true).  Only a cell without any survivor is
    // redone at minThFAST (a few per cent of the cells); scores written by the first attempt stay valid for the second.
    uint16_t* surv = list1;
    int attempt = 0;
    for (;; ++attempt) {
        const int t = attempt == 0 ? g.iniTh : g.minTh;

        // ---- pass A: compass test, 4 pixels per thread ----
        {
            const uint32_t kAdd = (uint32_t)(0x7f - (t & 0x7f)) * 0x01010101u;
            const bool tHigh = t >= 128;
            const int items = ih * nW;
            const int xb0 = 4 * wFirst - 3 - shift;  // interior x of byte 0 of the first word (<= 0)
            for (int it = tid; it < items; it += FAST_THREADS) {
                const int y = (int)__umulhi((uint32_t)it, magicW), cc = it - y * nW;
                const uint32_t* row = tileW + (y + 3) * TSW + 1 + wFirst + cc;
                const uint32_t cur = row[0], prev = row[-1], next = row[1], up = row[3 * TSW], dn = row[-3 * TSW];
                const uint32_t e = __funnelshift_r(cur, next, 24);  // pixels at x + 3
                const uint32_t w = __funnelshift_r(prev, cur, 8);   // pixels at x - 3
                const uint32_t ns = bytes_gt(__vabsdiffu4(up, cur), kAdd, tHigh) | bytes_gt(__vabsdiffu4(dn, cur), kAdd, tHigh);
                const uint32_t ew = bytes_gt(__vabsdiffu4(e, cur), kAdd, tHigh) | bytes_gt(__vabsdiffu4(w, cur), kAdd, tHigh);
                uint32_t F = ns & ew & 0x80808080u;
                if (F) {
                    // bytes outside the interior [0, iw) are dropped
                    const int xb = xb0 + 4 * cc;
                    const int nL = max(-xb, 0), nR = max(xb + 4 - iw, 0);
                    F &= (0xFFFFFFFFu << (8 * nL)) & (0xFFFFFFFFu >> (8 * nR));
                    if (F) {
                        const int off = (y + 3) * TS + 4 * (wFirst + cc);  // tile byte offset of byte 0 of this word
                        int pos = atomicAdd(&n1, __popc(F));
                        if (F & 0x80u) list1[pos++] = (uint16_t)off;
                        if (F & 0x8000u) list1[pos++] = (uint16_t)(off + 1);
                        if (F & 0x800000u) list1[pos++] = (uint16_t)(off + 2);
                        if (F & 0x80000000u) list1[pos] = (uint16_t)(off + 3);
                    }
                }
            }
        }
        __syncthreads();

        // ---- pass A2: exact ring masks + contiguity, dense over list1 ----
        const int nl1 = n1;
        for (int k0 = 0; k0 < nl1; k0 += FAST_THREADS) {
            const int k = k0 + tid;
            int off = 0;
            uint32_t arcs = 0;
            if (k < nl1) {
                off = list1[k];
                const uint8_t* c = tile + off;
                const uint32_t v = c[0];
                // 16x2 lanes: low = (v + t + 1) - (r + 1) = hi - r, high = r - lo; (r + 1) * 0xFFFF = (r << 16) | (-(r + 1) & 0xFFFF)
                const uint32_t K = (v + t + 1) | (((uint32_t)(t - (int)v) & 0xFFFFu) << 16);
                uint32_t m = 0;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const uint32_t r = c[RING(q)];
                    const uint32_t sdiff = __vadd2(K, r * 0xFFFFu + 0xFFFFu);
                    m |= ((sdiff >> (15 - q)) & (0x00010001u << q));  // bit q: r > hi, bit 16 + q: r < lo
                }
                arcs = arc9_halves(m);
            }
            const unsigned bal = __ballot_sync(FULL, arcs != 0);
            if (bal) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&n2, __popc(bal));
                base = __shfl_sync(FULL, base, 0);
                if (arcs) list2[base + __popc(bal & ((1u << lane) - 1))] = (uint16_t)(off | ((arcs >> 16) ? 0x8000 : 0));
            }
        }
        __syncthreads();

        // ---- pass B: arc score of the corners ----
        const int nl = n2;
        for (int k = tid; k < nl; k += FAST_THREADS) {
            const int e = list2[k], off = e & 0x7FFF;
            const uint8_t* c = tile + off;
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
            sc[off] = (uint8_t)s;  // t < s <= 255
        }
        __syncthreads();

        // ---- pass C: cell-local 3x3 non-maximum suppression.  Only interior pixels ever get a score, so the rim reads 0. ----
        for (int k0 = 0; k0 < nl; k0 += FAST_THREADS) {
            const int k = k0 + tid;
            bool keep = false;
            int off = 0;
            if (k < nl) {
                off = list2[k] & 0x7FFF;
                const uint8_t* q = sc + off;
                const int s = q[0];
                keep = s > q[-1] && s > q[1] && s > q[-TS - 1] && s > q[-TS] && s > q[-TS + 1] && s > q[TS - 1] && s > q[TS] && s > q[TS + 1];
            }
            const unsigned bal = __ballot_sync(FULL, keep);
            if (bal) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&nSurv, __popc(bal));
                base = __shfl_sync(FULL, base, 0);
                if (keep) surv[base + __popc(bal & ((1u << lane) - 1))] = (uint16_t)off;
            }
        }
        __syncthreads();
        if (nSurv > 0 || attempt == 1 || g.iniTh == g.minTh) break;
        __syncthreads();  // everyone has read nSurv == 0
        if (tid == 0) { n1 = 0; n2 = 0; }
        __syncthreads();
    }
    const int anyIni = attempt == 0 && nSurv > 0;

    // ---- emit the survivors, coordinates relative to (16,16) like the reference (ORBextractor.cc:837-842) ----
    const int ns = nSurv;
    uint32_t* outBuf = reinterpret_cast<uint32_t*>(list2);  // list2 is dead after pass C; <= maxPix/4 strict local maxima
    for (int k = tid; k < ns; k += FAST_THREADS) {
        const int off = surv[k];
        const int r = off / TS, cx = off - r * TS - shift;  // crop coordinates (the reference's cell-local keypoint position)
        outBuf[k] = pack_cand(cx + (iniX - kBorder), r + (iniY - kBorder), (int)sc[off] - 1);
    }
    if (tid == 0) nOut = ns;
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
        for (int k = tid; k < no; k += FAST_THREADS) out[k] = outBuf[k];
    }
}

// one descriptor per cell of every level (ORBextractor.cc:779-796), built once per geometry on the host
void make_cell_table(const Geom& g, uint4* tab, int* tileRows, int* maxPix) {
    int rows = 0, pix = 0;
    for (int l = 0; l < g.nlevels; ++l) {
        const LevelGeom& L = g.lv[l];
        const int maxBX = L.w - kBorder, maxBY = L.h - kBorder;
        for (int i = 0; i < L.nRows; ++i)
            for (int j = 0; j < L.nCols; ++j) {
                const int iniX = kBorder + j * L.wCell, iniY = kBorder + i * L.hCell;
                int cw = std::min(iniX + L.wCell + 6, maxBX) - iniX, ch = std::min(iniY + L.hCell + 6, maxBY) - iniY;
                int kind = 2;
                if (iniY >= maxBY - 3 || iniX >= maxBX - 6) kind = 0;
                else if (cw < 7 || ch < 7) kind = 1;
                uint4 d = make_uint4(0, 0, 0, 0);
                if (kind != 2) { cw = std::max(cw, 0); ch = std::max(ch, 0); }
                const int shift = iniX & 3, iw = cw - 6;
                const int wFirst = (3 + shift) >> 2, nW = kind == 2 ? ((3 + shift + iw - 1) >> 2) - wFirst + 1 : 1;
                d.x = (uint32_t)iniX | ((uint32_t)iniY << 16);
                d.y = (uint32_t)(cw & 0xFF) | ((uint32_t)(ch & 0xFF) << 8) | ((uint32_t)l << 16) | ((uint32_t)kind << 24);
                d.z = (0xFFFFFFFFu / (uint32_t)nW) + 1;  // exact it / nW for it < 2^16
                d.w = (uint32_t)nW | ((uint32_t)wFirst << 8) | ((uint32_t)shift << 16);
                tab[L.cellOff + i * L.nCols + j] = d;
                if (kind == 2) { rows = std::max(rows, ch); pix = std::max(pix, iw * (ch - 6)); }
            }
    }
    *tileRows = (std::max(rows, 7) + 3) & ~3;  // multiple of 4: keeps the score map 16-byte aligned (TSW is odd)
    *maxPix = (std::max(pix, 8) + 7) & ~7;
}

size_t fast_smem_bytes(int tileRows, int maxPix) { return (size_t)tileRows * TSW * 4 * 2 + (size_t)maxPix * 2 * 2 + 16; }

cudaError_t fast_configure(int tileRows, int maxPix) {
    return cudaFuncSetAttribute(k_fast_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(fast_smem_bytes(tileRows, maxPix), 48 * 1024));
}

void launch_fast(const Geom& g, const Buffers& b, int nImages, int tileRows, int maxPix, cudaStream_t s) {
    dim3 grid(g.cellsPerImage, nImages);
    k_fast_cells<<<grid, FAST_THREADS, fast_smem_bytes(tileRows, maxPix), s>>>(g, b, tileRows, maxPix);
}

}  // namespace orbx
