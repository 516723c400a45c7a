// fast_band_kernel.cu — cell-grid FAST-9/16 as a dense band pass (round 2; replaces the warp-per-cell kernel as the default).
//
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree (ORBextractor.cc:779-846) and the cv::FAST(cell, kps, th,
// true) calls inside it (ORBextractor.cc:800,819; arithmetic: SURVEY.md A.4), including the iniThFAST -> minThFAST per-cell
// fallback (ORBextractor.cc:817-820).
//
// What makes a cell-free formulation exact (derived from the OpenCV semantics, not from its code):
//   * the FAST score S of a pixel (max over the 16 arcs of 9 of the arc minimum of |ring - centre| with one sign) does not
//     depend on the threshold or on the crop; the pixel is a corner at t iff S > t and cv::FAST reports S - 1;
//   * the 3-px rims of neighbouring cell crops overlap by 6 px, so the crop INTERIORS tile the level's region [19, w-19) x
//     [19, h-19) without gaps or overlap: every pixel belongs to exactly one cell;
//   * cv::FAST's 3x3 non-maximum suppression compares a corner with the scores of its 8 neighbours inside the SAME crop's
//     interior (rim and non-corners read 0); a non-corner neighbour has S <= t < S(corner), so comparing with the true scores
//     of the neighbours that are corners at t and lie in the same cell is identical;
//   * a cell is redone at minThFAST iff it kept nothing at iniThFAST.
//
// Execution model: persistent CTAs, dynamic work queue; one work item = (image, band) with band = one cell ROW of one level cut
// into chunks of whole cells that fit a shared-memory tile of TP bytes per row (host table, make_fast_bands).  Per item:
//   load   the band (hCell + 6 rows) arrives as one bulk copy (TMA engine, cp.async.bulk -> UBLKCP) per row on one mbarrier;
//   dense  warps are autonomous: a lane takes 16 consecutive pixels (four 32-bit words) of the band's interior per trip: packed
//          compass test (|N - c| > t or |S - c| > t) and (|E - c| > t or |W - c| > t) with VABSDIFF4; the threshold compare is one
//          add per difference whose inter-byte carries can only ADD survivors (the exact score decides later); the 16 flags of a
//          lane are enumerated straight into the WARP's own pixel list (one shuffle scan per trip, no atomics, no barrier);
//   score  whenever the warp's list holds 32 pixels: one pixel per lane, both polarities at once as 16x2 lanes
//          [ring, 255 - ring] (one IMAD per ring pixel), the sixteen 9-windows from two levels of three-input packed minima
//          (VIMNMX3.U16x2), no half swaps; S > t -> score map (bytes, same geometry as the tile) and the warp's corner list;
//   nms    (after the one CTA barrier that completes the score map) one corner per lane against the map, neighbours in other
//          cells masked by a per-column table; survivors -> global candidate list (one atomic per warp and batch), cell hit flags;
//   redo   cells without a survivor: the same dense / score / nms sequence restricted to those cells at minThFAST (the tile is
//          still resident, nothing is re-read).
// A warp whose corner list overflows (textures such as saturated noise) switches the CTA's NMS to a scan of the score map.
// Integer only; HBM traffic = one read of every level (+ 6 halo rows per cell row through L2); the work is issue / ALU-pipe bound.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "orbx_internal.cuh"

namespace orbx {

namespace {

#ifndef FASTB_NT
#define FASTB_NT 256
#endif
#ifndef FASTB_DEFAULT_TP
#define FASTB_DEFAULT_TP 384
#endif
#ifndef FASTB_CAPC
#define FASTB_CAPC 192
#endif
#ifndef FASTB_MINB
#define FASTB_MINB 4
#endif
constexpr int NT = FASTB_NT;        // threads per CTA
constexpr int NWARP = NT / 32;
constexpr int CAPL = 32 + 512;      // a warp's pixel list: fewer than 32 waiting + one trip of 32 lanes x 16 pixels
constexpr int CAPC = FASTB_CAPC;    // a warp's corner list
constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int kMaxBandCells = 63;   // local cell index: 6 bits of the column table

__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {  // a * b + c on the FMA pipe (kept out of the ALU pipe on purpose)
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ int atom_add_global(int* p, int v) {  // (plain atomicAdd makes ptxas wrap the call in its own warp aggregation)
    int r;
    asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "r"(v) : "memory");
    return r;
}

// bit 7 of every byte set where |x - c| > t can hold (superset: a carry out of a byte below can add a survivor at d == t).
// kAdd = (0x7f - t) * 0x01010101 for t < 128, 0 for t >= 128 (then the test is d >= 128).  No false negatives: d >= 128 has its
// own bit 7; below that d + (0x7f - t) <= 254 stays inside the byte and reaches 0x80 iff d > t.
__device__ __forceinline__ uint32_t compass4(uint32_t up, uint32_t dn, uint32_t e, uint32_t w, uint32_t c, uint32_t kAdd, uint32_t m) {
    const uint32_t dN = __vabsdiffu4(up, c), dS = __vabsdiffu4(dn, c), dE = __vabsdiffu4(e, c), dW = __vabsdiffu4(w, c);
    const uint32_t lN = dN + kAdd, lS = dS + kAdd, lE = dE + kAdd, lW = dW + kAdd;
    const uint32_t a = lN | dN | lS, b = lE | dE | lW;
    return ((a | dS) & m) & (b | dW);
}

// byte offsets of the 16 ring pixels inside a tile with row stride TS (SURVEY.md A.4 order)
#define RINGB(TS, k) ((k) == 0 ? 3 * (TS) : (k) == 1 ? 3 * (TS) + 1 : (k) == 2 ? 2 * (TS) + 2 : (k) == 3 ? (TS) + 3 : (k) == 4 ? 3 : (k) == 5 ? -(TS) + 3 : \
                      (k) == 6 ? -2 * (TS) + 2 : (k) == 7 ? -3 * (TS) + 1 : (k) == 8 ? -3 * (TS) : (k) == 9 ? -3 * (TS) - 1 : (k) == 10 ? -2 * (TS) - 2 : \
                      (k) == 11 ? -(TS) - 3 : (k) == 12 ? -3 : (k) == 13 ? (TS) - 3 : (k) == 14 ? 2 * (TS) - 2 : 3 * (TS) - 1)

// exact score of the pixel at tile offset `off`:  S = max( max_s min_{j<9} r[s+j] - v,  v - min_s max_{j<9} r[s+j] ).
// Lanes of a register: low half r[k], high half 255 - r[k]; the minimum over a window then is (min r, 255 - max r).
template <int TP>
__device__ __forceinline__ int fast_score(const uint8_t* tile, int off) {
    const uint8_t* p = tile + off;
    const int v = p[0];
    uint32_t R[18];
#pragma unroll
    for (int k = 0; k < 16; ++k) R[k] = imad((uint32_t)p[RINGB(TP, k)], 0xFFFF0001u, 0x00FF0000u);
    R[16] = R[0];
    R[17] = R[1];
    uint32_t m3[22];
#pragma unroll
    for (int k = 0; k < 16; ++k) m3[k] = __vimin3_u16x2(R[k], R[k + 1], R[k + 2]);
#pragma unroll
    for (int k = 0; k < 6; ++k) m3[16 + k] = m3[k];
    uint32_t W[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) W[k] = __vimin3_u16x2(m3[k], m3[k + 3], m3[k + 6]);
    const uint32_t a0 = __vimax3_u16x2(W[0], W[1], W[2]), a1 = __vimax3_u16x2(W[3], W[4], W[5]), a2 = __vimax3_u16x2(W[6], W[7], W[8]);
    const uint32_t a3 = __vimax3_u16x2(W[9], W[10], W[11]), a4 = __vimax3_u16x2(W[12], W[13], W[14]);
    const uint32_t b0 = __vimax3_u16x2(a0, a1, a2), b1 = __vimax3_u16x2(a3, a4, W[15]);
    const uint32_t best = __vmaxu2(b0, b1);
    const int bright = (int)(best & 0xFFFFu) - v;        // max_s min r - v
    const int dark = (int)(best >> 16) + v - 255;        // v - min_s max r
    return max(bright, dark);
}

struct Ctx {
    uint32_t tileS;      // shared-window address of the tile
    uint8_t* tile;
    uint8_t* sc;         // score byte of tile offset `off` at sc[off] (the map starts two rows into the tile geometry)
    uint16_t* plist;     // this warp's flagged pixels (tile byte offsets)
    uint16_t* clist;     // this warp's corners (tile byte offsets)
    int* ovf2;           // CTA flag: a corner list overflowed
    int n, nc;           // fill levels (warp uniform)
    int lane;
    unsigned lt;         // lanes below this one
};

// the last min(n, 32) entries of the warp's pixel list: exact score, corners at t -> score map + corner list
template <int TP>
__device__ __forceinline__ void score_group(Ctx& c, int t) {
    const int first = max(c.n - 32, 0), k = first + c.lane;
    int off = 3 * TP + 4, S = 0;
    if (k < c.n) {
        off = c.plist[k];
        S = fast_score<TP>(c.tile, off);
    }
    const bool corner = S > t;
    const unsigned bal = __ballot_sync(FULL, corner);
    if (corner) {
        c.sc[off] = (uint8_t)S;  // t < S <= 255
        const int pos = c.nc + __popc(bal & c.lt);
        if (pos < CAPC) c.clist[pos] = (uint16_t)off;
        else *c.ovf2 = 1;
    }
    c.nc += __popc(bal);
    c.n = first;
    __syncwarp();  // the slots are free for the next enumeration
}

// the flags of one lane trip -> the warp's pixel list; full groups of 32 are scored right away.
// H: bit 8 * byte + 7 - word set for pixel 4 * word + byte of the 16 pixels starting at tile offset off16 (words 0..3).
template <int TP>
__device__ __forceinline__ void enumerate(Ctx& c, uint32_t H, uint32_t off16, int t) {
    const int cnt = __popc(H);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int v;
        asm("{\n\t.reg .pred p;\n\tshfl.sync.up.b32 %0|p, %1, %2, 0, 0xffffffff;\n\t@!p mov.b32 %0, 0;\n\t}" : "=r"(v) : "r"(incl), "r"(d));
        incl += v;
    }
    const int tot = __shfl_sync(FULL, incl, 31);
    if (tot == 0) return;  // warp uniform
    int pos = c.n + incl - cnt;
    while (H) {
        const uint32_t bit = 31u - (uint32_t)__clz(H);
        H ^= 1u << bit;
        // pixel = 4 * (7 - (bit & 7)) + (bit >> 3) = 28 - 4 * bit + 33 * (bit >> 3)
        c.plist[pos++] = (uint16_t)imad(bit >> 3, 33u, imad(bit, 0xFFFFFFFCu, off16 + 28u));
    }
    c.n += tot;
    __syncwarp();
    while (c.n >= 32) score_group<TP>(c, t);
}

}  // namespace

template <int TP>
__global__ void __launch_bounds__(NT, FASTB_MINB) k_fast_band(const __grid_constant__ Geom g, const Buffers b, const FastBand* __restrict__ bands,
                                                              const uint8_t* __restrict__ aux, int nBands, int nWork, int tileRows, int* __restrict__ workCtr) {
    extern __shared__ __align__(128) unsigned char fsm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int sNe, sOvf2, sNext[3];
    __shared__ int sHit[kMaxBandCells + 1];
    __shared__ uint8_t sEmpty[kMaxBandCells + 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* tile = fsm;
    uint8_t* scm = tile + (size_t)tileRows * TP;                         // score rows 2 .. tileRows-3
    uint32_t* smask = reinterpret_cast<uint32_t*>(scm + (size_t)(tileRows - 4) * TP);   // [TP / 4 + 8]
    uint8_t* sCol = reinterpret_cast<uint8_t*>(smask + TP / 4 + 8);                     // [TP]
    uint16_t* plistAll = reinterpret_cast<uint16_t*>(sCol + TP);
    uint16_t* clistAll = plistAll + NWARP * CAPL;

    Ctx c;
    c.tile = tile;
    c.tileS = __shfl_sync(FULL, smem_u32(tile), 0);  // through a shuffle: ptxas otherwise re-derives the window address at every use
    c.sc = scm - 2 * TP;
    c.plist = plistAll + warp * CAPL;
    c.clist = clistAll + warp * CAPC;
    c.ovf2 = &sOvf2;
    c.n = 0;
    c.nc = 0;
    c.lane = lane;
    c.lt = (1u << lane) - 1u;
    const uint32_t smaskS = __shfl_sync(FULL, smem_u32(smask), 0);

    for (int k = tid; k < (tileRows - 4) * (TP / 16); k += NT) reinterpret_cast<uint4*>(scm)[k] = make_uint4(0, 0, 0, 0);  // stays all-zero between items
    if (tid == 0) {
        mbar_init(&bar, 1);
        sNe = 0; sOvf2 = 0;
    }
    mbar_fence_init();
    __syncthreads();
    uint32_t phase = 0;

    int w = blockIdx.x, imgRel = w / nBands, bandIdx = w - imgRel * nBands;
    while (w < nWork) {
        const FastBand d = bands[bandIdx];
        const int img = b.imgBase + imgRel;
        const LevelGeom& L = g.lv[d.level];
        const int ih = d.ch - 6, nDr = d.nDr;

        // ---- load: one bulk copy per tile row (warp 0), tables and the next work item meanwhile ----
        if (warp == 0) {
            if (lane == 0) mbar_arrive_expect_tx(&bar, d.rowBytes * d.ch);
            __syncwarp();
            const uint8_t* src = b.pyr + L.planeOff + (size_t)img * L.planeBytes + (size_t)d.iniY * L.pitch + d.iniX - d.shift;
            for (int r = lane; r < d.ch; r += 32)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(c.tileS + r * TP),
                             "l"(src + (size_t)r * L.pitch), "r"(d.rowBytes), "r"(smem_u32(&bar))
                             : "memory");
        } else if (warp == 1 && lane == 0) {
            const int nx = gridDim.x + atom_add_global(workCtr, 1);  // dynamic queue behind the first gridDim.x items
            const int ni = nx / nBands;
            sNext[0] = nx; sNext[1] = ni; sNext[2] = nx - ni * nBands;
        }
        {
            const uint32_t* gm = reinterpret_cast<const uint32_t*>(aux + d.aux);
            const int nMask = 4 * nDr + 4, nCol = (d.cw + 3) >> 2;
            for (int k = tid; k < nMask + nCol; k += NT) {
                const uint32_t v = __ldg(gm + k);
                if (k < nMask) smask[k] = v;
                else reinterpret_cast<uint32_t*>(sCol)[k - nMask] = v;
            }
            if (tid <= kMaxBandCells) sHit[tid] = 0;
        }
        __syncthreads();
        mbar_wait(&bar, phase);
        phase ^= 1;

        uint32_t* cand = b.cand + L.candOff + (size_t)img * L.candCap;
        int* candCount = b.candCount + img * ORBX_MAX_LEVELS + d.level;
        const int relX = (int)d.iniX - kBorder - (int)d.shift, relY = (int)d.iniY - kBorder;  // candidate = (tile byte column + relX, tile row + relY)

        // one corner per lane: it survives iff its score is strictly greater than the scores of its 8 neighbours inside the same
        // cell's interior; survivors -> global candidate list (coordinates relative to (16,16), ORBextractor.cc:837-842)
        auto nms_group = [&](bool act, int off) {
            bool keep = false;
            int s = 0, info = 0, r = 0, cb = 0;
            if (act) {
                r = off / TP;
                cb = off - r * TP;
                info = sCol[cb - d.shift];
                const uint8_t* q = c.sc + off;
                s = q[0];
                int mL = max(max((int)q[-TP - 1], (int)q[-1]), (int)q[TP - 1]);
                int mR = max(max((int)q[-TP + 1], (int)q[1]), (int)q[TP + 1]);
                const int mM = max((int)q[-TP], (int)q[TP]);
                if (info & 1) mL = 0;  // first interior column of its cell: the left neighbours belong to another crop's interior
                if (info & 2) mR = 0;
                keep = s > max(mM, max(mL, mR));
            }
            const unsigned bal = __ballot_sync(FULL, keep);
            if (!bal) return;
            const int ns = __popc(bal);
            int base = 0;
            if (lane == 0) base = atom_add_global(candCount, ns);
            base = __shfl_sync(FULL, base, 0);
            if (base + ns > L.candCap) {
                if (lane == 0) atomicExch(b.errFlag, 1);
                return;
            }
            if (keep) {
                cand[base + __popc(bal & c.lt)] = pack_cand(cb + relX, r + relY, s - 1);
                sHit[info >> 2] = 1;
            }
        };
        // the warp's corner list (or, after an overflow anywhere in the CTA, this warp's share of the whole score map) -> NMS
        auto nms_pass = [&](int ovf) {
            if (!ovf) {
                const int nc = c.nc;
                for (int k0 = 0; k0 < nc; k0 += 32) {
                    const int k = k0 + lane;
                    nms_group(k < nc, k < nc ? (int)c.clist[k] : 0);
                }
            } else {
                const int iw = d.cw - 6, tot = iw * ih;
                for (int p0 = warp * 32; p0 < tot; p0 += NT) {
                    const int p = p0 + lane;
                    int off = 0;
                    bool act = false;
                    if (p < tot) {
                        const int y = p / iw, x = p - y * iw;
                        off = (y + 3) * TP + d.shift + 3 + x;
                        act = c.sc[off] != 0;
                    }
                    nms_group(act, off);
                }
            }
        };
        auto wipe_scores = [&](int ovf) {
            if (!ovf) {
                for (int k = lane; k < c.nc; k += 32) c.sc[c.clist[k]] = 0;
            } else {
                for (int k = tid; k < (d.ch - 4) * (TP / 16); k += NT) reinterpret_cast<uint4*>(scm)[k] = make_uint4(0, 0, 0, 0);
            }
            c.nc = 0;
        };

        // ================= pass 0: the whole band at iniThFAST =================
        {
            const int t = g.iniTh;
            const uint32_t kAdd = t < 128 ? (uint32_t)(0x7f - t) * 0x01010101u : 0u;
            const uint32_t rowBase = 3 * TP + 4 * d.wFirstP;  // tile byte offset of the first word of interior row 0
            const int nTot = nDr * ih;
            for (int base = warp * 32; base < nTot; base += NT) {
                const int i0 = base + lane;
                const bool valid = i0 < nTot;
                const uint32_t i = valid ? (uint32_t)i0 : 0u;
                const uint32_t rr = __umulhi(i, d.magicD), dc = i - rr * (uint32_t)nDr;
                const uint32_t off = rowBase + rr * TP + 16u * dc;
                const uint32_t a = c.tileS + off;
                uint32_t c0, c1, c2, c3, u0, u1, u2, u3, d0, d1, d2, d3, pv, nx, m0, m1, m2, m3;
                asm volatile(
                    "ld.shared.v2.u32 {%0,%1}, [%18];\n\t"
                    "ld.shared.v2.u32 {%2,%3}, [%18+8];\n\t"
                    "ld.shared.v2.u32 {%4,%5}, [%18+-%20];\n\t"
                    "ld.shared.v2.u32 {%6,%7}, [%18+-%21];\n\t"
                    "ld.shared.v2.u32 {%8,%9}, [%18+%20];\n\t"
                    "ld.shared.v2.u32 {%10,%11}, [%18+%22];\n\t"
                    "ld.shared.u32 %12, [%18+-4];\n\t"
                    "ld.shared.u32 %13, [%18+16];\n\t"
                    "ld.shared.v4.u32 {%14,%15,%16,%17}, [%19];"
                    : "=r"(c0), "=r"(c1), "=r"(c2), "=r"(c3), "=r"(u0), "=r"(u1), "=r"(u2), "=r"(u3), "=r"(d0), "=r"(d1), "=r"(d2), "=r"(d3), "=r"(pv),
                      "=r"(nx), "=r"(m0), "=r"(m1), "=r"(m2), "=r"(m3)
                    : "r"(a), "r"(smaskS + 16u * (valid ? dc : (uint32_t)nDr)), "n"(3 * TP), "n"(3 * TP - 8), "n"(3 * TP + 8));
                // pixels at x + 3 / x - 3 of every word
                const uint32_t F0 = compass4(u0, d0, __funnelshift_r(c0, c1, 24), __funnelshift_r(pv, c0, 8), c0, kAdd, m0);
                const uint32_t F1 = compass4(u1, d1, __funnelshift_r(c1, c2, 24), __funnelshift_r(c0, c1, 8), c1, kAdd, m1);
                const uint32_t F2 = compass4(u2, d2, __funnelshift_r(c2, c3, 24), __funnelshift_r(c1, c2, 8), c2, kAdd, m2);
                const uint32_t F3 = compass4(u3, d3, __funnelshift_r(c3, nx, 24), __funnelshift_r(c2, c3, 8), c3, kAdd, m3);
                enumerate<TP>(c, F0 | (F1 >> 1) | (F2 >> 2) | (F3 >> 3), off, t);
            }
            if (c.n > 0) score_group<TP>(c, t);
            __syncthreads();  // the score map is complete
            const int ovf = sOvf2;
            nms_pass(ovf);
            __syncthreads();  // the hit flags are complete; every thread holds its copy of the overflow flag
            if (tid == 0) sOvf2 = 0;
            wipe_scores(ovf);
            // flags (ORBextractor.cc:817-820: a cell that kept nothing is redone at minThFAST) and the list of cells to redo
            if (tid < d.nCells) {
                const bool hit = sHit[tid] != 0;
                b.cellFlags[(size_t)img * g.cellsPerImage + d.cell0 + tid] = hit ? 1 : 2;
                if (!hit && g.iniTh != g.minTh) sEmpty[atomicAdd(&sNe, 1)] = (uint8_t)tid;
            }
            __syncthreads();
        }

        // ================= pass 1: cells that kept nothing, at minThFAST =================
        const int ne = sNe;
        if (ne > 0) {
            const int t = g.minTh;
            const uint32_t kAdd = t < 128 ? (uint32_t)(0x7f - t) * 0x01010101u : 0u;
            const int nwc = d.nwc, perCell = nwc * ih;
            for (int e0 = 0; e0 < ne; e0 += 8) {  // (chunks keep the flattened index below 2^16)
                const int nTot = min(8, ne - e0) * perCell;
                for (int base = warp * 32; base < nTot; base += NT) {
                    const int i0 = base + lane;
                    uint32_t F = 0u, off = 3 * TP;
                    if (i0 < nTot) {
                        const uint32_t i = (uint32_t)i0;
                        const uint32_t ce = __umulhi(i, d.magicCell), rem = i - ce * (uint32_t)perCell;
                        const uint32_t rr = __umulhi(rem, d.magicWc), wc = rem - rr * (uint32_t)nwc;
                        const int lc = sEmpty[e0 + ce];
                        const int x0 = 3 + lc * d.wCell, x1 = min(x0 + (int)d.wCell, (int)d.cw - 3);   // interior columns of the cell (crop coordinates)
                        const int wi = ((d.shift + x0) >> 2) + (int)wc;                             // tile word index
                        if (wi <= ((d.shift + x1 - 1) >> 2)) {
                            const int cb = 4 * wi - d.shift;                                          // crop column of the word's byte 0
                            const int lo = max(0, x0 - cb), hi = max(0, cb + 4 - x1);
                            const uint32_t m = (0x80808080u << (8 * lo)) & (0x80808080u >> (8 * hi));
                            off = (rr + 3u) * TP + 4u * (uint32_t)wi;
                            const uint32_t a = c.tileS + off;
                            uint32_t cur, pv, nx, up, dn;
                            asm volatile(
                                "ld.shared.u32 %0, [%5];\n\t"
                                "ld.shared.u32 %1, [%5+-4];\n\t"
                                "ld.shared.u32 %2, [%5+4];\n\t"
                                "ld.shared.u32 %3, [%5+-%6];\n\t"
                                "ld.shared.u32 %4, [%5+%6];"
                                : "=r"(cur), "=r"(pv), "=r"(nx), "=r"(up), "=r"(dn)
                                : "r"(a), "n"(3 * TP));
                            F = compass4(up, dn, __funnelshift_r(cur, nx, 24), __funnelshift_r(pv, cur, 8), cur, kAdd, m);
                        }
                    }
                    enumerate<TP>(c, F, off, t);
                }
            }
            if (c.n > 0) score_group<TP>(c, t);
            __syncthreads();
            const int ovf = sOvf2;
            nms_pass(ovf);
            __syncthreads();
            if (tid == 0) { sOvf2 = 0; sNe = 0; }
            wipe_scores(ovf);
        }
        w = sNext[0]; imgRel = sNext[1]; bandIdx = sNext[2];
        __syncthreads();  // every thread has read the next item and is done with the tile, the score map and the flags
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// host: the band table of one geometry (ORBextractor.cc:759-796 cell grid, cut into bands)
// ---------------------------------------------------------------------------------------------------------------------------
static uint32_t magic_div(uint32_t d) { return d <= 1 ? 0xFFFFFFFFu : (0xFFFFFFFFu / d) + 1; }  // exact i / d for i < 2^16 (d == 1: handled by callers' ranges)

int make_fast_bands(const Geom& g, int tp, std::vector<FastBand>& bands, std::vector<uint8_t>& auxOut, std::vector<uint8_t>& flagTemplate,
                    int* tileRows) {
    bands.clear();
    auxOut.clear();
    flagTemplate.assign((size_t)g.cellsPerImage, 0);
    int rows = 7;
    const int cwMax = tp - 15;  // a crop row starts up to 15 bytes into its 16-byte aligned tile row
    for (int l = 0; l < g.nlevels; ++l) {
        const LevelGeom& L = g.lv[l];
        const int maxBX = L.w - kBorder, maxBY = L.h - kBorder;
        // cells with FAST work form a prefix of every cell row (ORBextractor.cc:785-796)
        int nValid = 0;
        for (int j = 0; j < L.nCols; ++j) {
            const int iniX = kBorder + j * L.wCell;
            if (iniX >= maxBX - 6) break;
            if (std::min(iniX + L.wCell + 6, maxBX) - iniX < 7) break;
            nValid = j + 1;
        }
        for (int i = 0; i < L.nRows; ++i) {
            const int iniY = kBorder + i * L.hCell;
            const bool rowSkipped = iniY >= maxBY - 3;
            const int ch = std::min(iniY + L.hCell + 6, maxBY) - iniY;
            for (int j = 0; j < L.nCols; ++j) {
                const int iniX = kBorder + j * L.wCell;
                const bool skipped = rowSkipped || iniX >= maxBX - 6;
                // 0: the reference skips the cell; 2: cv::FAST yields nothing at either threshold (crop thinner than the ring);
                // cells with work are overwritten by the kernel on every run
                flagTemplate[(size_t)L.cellOff + i * L.nCols + j] = skipped ? 0 : 2;
            }
            if (rowSkipped || ch < 7 || nValid == 0) continue;
            const int maxCells = std::min(kMaxBandCells, (cwMax - 6) / L.wCell);
            if (maxCells < 1) return -1;
            const int nChunks = (nValid + maxCells - 1) / maxCells;
            for (int q = 0; q < nChunks; ++q) {
                const int c0 = (int)((long long)nValid * q / nChunks), c1 = (int)((long long)nValid * (q + 1) / nChunks);
                FastBand d;
                std::memset(&d, 0, sizeof(d));
                const int iniX = kBorder + c0 * L.wCell;
                const int lastX = kBorder + (c1 - 1) * L.wCell;
                const int cw = std::min(lastX + L.wCell + 6, maxBX) - iniX;
                const int shift = (kMarginX + iniX) & 15;
                if (shift + cw > tp) return -1;
                const int wFirst = (shift + 3) >> 2, wLast = (shift + cw - 4) >> 2;
                const int wFirstP = wFirst & ~1, nPr = ((wLast | 1) - wFirstP + 1) / 2;
                d.iniX = (uint16_t)iniX; d.iniY = (uint16_t)iniY; d.cw = (uint16_t)cw; d.ch = (uint16_t)ch;
                d.cell0 = (uint16_t)(L.cellOff + i * L.nCols + c0);
                const int nDr = (nPr + 1) / 2;
                d.nDr = (uint16_t)nDr;
                d.level = (uint8_t)l; d.nCells = (uint8_t)(c1 - c0); d.shift = (uint8_t)shift; d.wFirstP = (uint8_t)wFirstP;
                d.magicD = magic_div((uint32_t)nDr);
                d.wCell = (uint16_t)L.wCell;
                d.nwc = (uint16_t)((L.wCell + 2) / 4 + 1);   // words touched by wCell interior pixels at any alignment
                d.magicWc = magic_div(d.nwc);
                d.magicCell = magic_div((uint32_t)d.nwc * (uint32_t)(ch - 6));
                d.rowBytes = (uint32_t)((shift + cw + 15) & ~15);
                while (auxOut.size() % 16) auxOut.push_back(0);
                d.aux = (uint32_t)auxOut.size();
                for (int k = 0; k < 4 * nDr + 4; ++k) {  // (the words past the interior and the four of the idle lanes' slot are zero)
                    uint32_t m = 0;
                    for (int by = 0; by < 4 && k < 2 * nPr; ++by) {
                        const int col = 4 * (wFirstP + k) + by - shift;  // crop column of this byte
                        if (col >= 3 && col < cw - 3) m |= 0x80u << (8 * by);
                    }
                    for (int by = 0; by < 4; ++by) auxOut.push_back((uint8_t)(m >> (8 * by)));
                }
                for (int col = 0; col < cw; ++col) {
                    uint8_t info = 0xFF;
                    if (col >= 3 && col < cw - 3) {
                        const int lc = (col - 3) / L.wCell, ph = (col - 3) % L.wCell;
                        info = (uint8_t)((ph == 0 ? 1 : 0) | ((ph == L.wCell - 1 || col == cw - 4) ? 2 : 0) | (lc << 2));
                    }
                    auxOut.push_back(info);
                }
                while (auxOut.size() % 4) auxOut.push_back(0xFF);
                bands.push_back(d);
                rows = std::max(rows, ch);
            }
        }
    }
    if (bands.empty()) return -1;
    *tileRows = rows;
    return 0;
}

size_t fast_band_smem(int tileRows, int tp) {
    return (size_t)tileRows * tp + (size_t)(tileRows - 4) * tp + (size_t)(tp / 4 + 8) * 4 + (size_t)tp + (size_t)NWARP * CAPL * 2 + (size_t)NWARP * CAPC * 2;
}

// row pitch of the shared-memory tile: the widest instantiation whose two planes still leave room for two CTAs per SM
int fast_band_pick_tp(const Geom& g) {
    if (const char* e = std::getenv("ORBX_FAST_TP")) {
        const int v = std::atoi(e);
        if (v == 384 || v == 640) return v;
    }
    (void)g;
    return FASTB_DEFAULT_TP;
}

// the dynamic shared-memory limit is a per-kernel, per-device attribute: raise it once to the device maximum (never lower it for a later handle)
cudaError_t fast_band_configure() {
    int dev = 0, optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
    cudaFuncAttributes fa;
    if ((e = cudaFuncGetAttributes(&fa, k_fast_band<384>)) != cudaSuccess) return e;
    const int dyn = optin - (int)fa.sharedSizeBytes;
    if ((e = cudaFuncSetAttribute(k_fast_band<384>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_fast_band<640>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
}

void launch_fast_band(const Geom& g, const Buffers& b, const FastBand* bands, const uint8_t* aux, int nBands, int tileRows, int tp, int nImages, int numSMs,
                      int* workCtr, cudaStream_t s) {
    const size_t smem = fast_band_smem(tileRows, tp);
    const int ctasPerSM = std::max(1, (int)((227 * 1024) / (smem + 2048)));
    const int nWork = nBands * nImages;
    const int grid = std::min(nWork, numSMs * ctasPerSM);
    cudaMemsetAsync(workCtr, 0, sizeof(int), s);
    if (tp == 384)
        k_fast_band<384><<<grid, NT, smem, s>>>(g, b, bands, aux, nBands, nWork, tileRows, workCtr);
    else
        k_fast_band<640><<<grid, NT, smem, s>>>(g, b, bands, aux, nBands, nWork, tileRows, workCtr);
}

}  // namespace orbx
