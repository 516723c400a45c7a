// fast_strip_kernel.cu — cell-grid FAST-9/16 at iniThFAST as a streaming strip pass (the default FAST stage, round 2).
//
// Replaces, together with the redo mode of fast_kernel.cu, the cell loop of ORBextractor::ComputeKeyPointsOctTree
// (ORBextractor.cc:779-846) and the cv::FAST(cell, kps, th, true) calls inside it (ORBextractor.cc:800,819; arithmetic: SURVEY.md A.4).
// This kernel does every cell at iniThFAST and marks the cells that kept nothing; those are redone at minThFAST
// (ORBextractor.cc:817-820) by k_fast_cells<TS, true>, which is launched right behind it.
//
// What makes a cell-free formulation exact (derived from the OpenCV semantics, not from its code):
//   * the FAST score S of a pixel (max over the 16 arcs of 9 of the arc minimum of |ring - centre| with one sign) does not
//     depend on the threshold or on the crop; the pixel is a corner at t iff S > t and cv::FAST reports S - 1;
//   * the 3-px rims of neighbouring cell crops overlap by 6 px, so the crop INTERIORS tile the level's region [19, w-19) x
//     [19, h-19) without gaps or overlap: every pixel belongs to exactly one cell;
//   * cv::FAST's 3x3 non-maximum suppression compares a corner with the scores of its 8 neighbours inside the SAME crop's
//     interior (rim and non-corners read 0); a non-corner neighbour has S <= t < S(corner), so comparing with the true scores
//     of the neighbours that are corners at t and lie in the same cell is identical.
//
// Execution model: every WARP is autonomous (no block barrier, no shared state between warps).  A work item is a column strip
// of up to three cells (at most 121 interior pixels = 32 words) over a range of cell rows; the warp streams down the strip in
// steps of 8 pixel rows through a ring of 32 rows in its own shared memory, fed by TMA boxes of 160 bytes x 8 rows (one
// mbarrier per ring slot, the box two steps ahead is in flight while a step is computed).  The first / last four ring rows are
// mirrored behind / in front of the ring by two 4-row boxes, so every +-3 row neighbourhood is linear in memory.  Per step:
//   dense  a lane takes 16 consecutive pixels (four 32-bit words) of one of four rows per trip: packed compass test
//          (|N - c| > t or |S - c| > t) and (|E - c| > t or |W - c| > t) with VABSDIFF4; the threshold compare is one add per
//          difference whose inter-byte carries can only ADD survivors (the exact score decides later).  A lane with flags
//          appends ONE record {16 flags, ring offset, row} to the warp's queue (one ballot per trip);
//   score  whenever all 32 lanes can be kept busy: a lane holds a record and scores one of its flagged pixels per round, idle
//          lanes take the next records of the queue; both polarities at once as 16x2 lanes [ring, 255 - ring] (one IMAD per
//          ring pixel), the sixteen 9-windows from two levels of three-input packed minima (VIMNMX3.U16x2), no half swaps;
//          S > t -> score ring (bytes, same geometry as the pixel ring) and the warp's corner list;
//   nms    every second step, for the rows whose neighbours are completely scored: one corner per lane against the score ring,
//          neighbours in other cells masked (per-column table, row phase inside the cell row); survivors -> global candidate
//          list (one atomic per group), hit flag of their cell.
// Records are finished before the ring slot under their pixels is reloaded (at most one step later); a corner list that
// overflows (textures such as saturated noise) makes the NMS scan the score-ring rows concerned instead.
// Integer only; HBM traffic = one read of every level (+ 6 halo rows per item through L2); the work is issue / ALU-pipe bound.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "orbx_internal.cuh"

namespace orbx {

namespace {

#ifndef FASTS_WARPS
#define FASTS_WARPS 1
#endif
constexpr int SW = FASTS_WARPS;     // warps per CTA (independent of one another)
constexpr int TS = 160;             // bytes per ring row = TMA box width
constexpr int RING = 32;            // ring rows: four slots of eight
constexpr int QCAP = 256;           // pixel queue (power of two): fewer than 32 waiting + the flagged pixels of one step (dense textures go lane by lane)
constexpr int CCAP = 256;           // corner list
constexpr int PIX_BYTES = (RING + 8) * TS;   // ring row r at buffer row r + 4: TMA destinations (ring slots, mirrors) are 128-byte aligned
constexpr int SC_BYTES = RING * TS;
constexpr int COL_BYTES = 144;
constexpr int WARP_BYTES = (PIX_BYTES + SC_BYTES + QCAP * 4 + CCAP * 4 + COL_BYTES + 4 * 8 + 127) / 128 * 128;
static_assert((4 * TS) % 128 == 0 && (8 * TS) % 128 == 0, "ring slots must be 128-byte aligned");
constexpr unsigned FULL = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {  // a * b + c on the FMA pipe
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ int atom_add_global(int* p, int v) {
    int r;
    asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "r"(v) : "memory");
    return r;
}

// bit 7 of every byte set where |x - c| > t can hold (superset: a carry out of a byte below can add a survivor at d == t).
// kAdd = (0x7f - t) * 0x01010101 for t < 128, 0 for t >= 128 (then the test is d >= 128).  No false negatives: d >= 128 has its
// own bit 7; below that d + (0x7f - t) <= 254 stays inside the byte and reaches 0x80 iff d > t.
__device__ __forceinline__ uint32_t compass4(uint32_t up, uint32_t dn, uint32_t e, uint32_t w, uint32_t c, uint32_t kAdd, uint32_t m, uint32_t one) {
    const uint32_t dN = __vabsdiffu4(up, c), dS = __vabsdiffu4(dn, c), dE = __vabsdiffu4(e, c), dW = __vabsdiffu4(w, c);
    const uint32_t lN = imad(dN, one, kAdd), lS = imad(dS, one, kAdd), lE = imad(dE, one, kAdd), lW = imad(dW, one, kAdd);  // (`one` == 1, opaque to ptxas)
    const uint32_t a = lN | dN | lS, b = lE | dE | lW;
    return ((a | dS) & m) & (b | dW);
}

// byte offsets of the 16 ring pixels inside a tile with row stride TS (SURVEY.md A.4 order)
#define RINGB(k) ((k) == 0 ? 3 * TS : (k) == 1 ? 3 * TS + 1 : (k) == 2 ? 2 * TS + 2 : (k) == 3 ? TS + 3 : (k) == 4 ? 3 : (k) == 5 ? -TS + 3 : \
                  (k) == 6 ? -2 * TS + 2 : (k) == 7 ? -3 * TS + 1 : (k) == 8 ? -3 * TS : (k) == 9 ? -3 * TS - 1 : (k) == 10 ? -2 * TS - 2 : \
                  (k) == 11 ? -TS - 3 : (k) == 12 ? -3 : (k) == 13 ? TS - 3 : (k) == 14 ? 2 * TS - 2 : 3 * TS - 1)

// exact score of the pixel at byte offset `off` of the pixel buffer:  S = max( max_s min_{j<9} r[s+j] - v,  v - min_s max_{j<9} r[s+j] ).
// Lanes of a register: low half r[k], high half 255 - r[k]; the minimum over a window then is (min r, 255 - max r).
__device__ __forceinline__ int fast_score(const uint8_t* pix, int off) {
    const uint8_t* p = pix + off;
    const int v = p[0];
    uint32_t R[18];
#pragma unroll
    for (int k = 0; k < 16; ++k) R[k] = imad((uint32_t)p[RINGB(k)], 0xFFFF0001u, 0x00FF0000u);
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

// state of one warp while it streams a strip
struct WarpState {
    uint8_t* pix;        // (RING + 8) rows of TS bytes: ring row r at row r + 4, rows 0..3 mirror ring rows 28..31, rows 36..39 ring rows 0..3
    uint8_t* sc;         // RING rows of TS bytes: score of the pixel at pix offset off at sc[off - 4 * TS]
    uint32_t* q;         // flagged pixels waiting for their score (FIFO): pix offset | local row << 13
    uint32_t* cl;        // corners waiting for their NMS: same encoding
    int qh, qt;          // queue head / tail (free running)
    int ncl;             // corners in the list
    int scanRowMax;      // corners were dropped from the list up to this local row (-1: none): those rows are scanned instead
    int maxRow;          // last local row of the step being computed
    int lane;
    unsigned lt;
};

// one scoring round: the first min(32, waiting) pixels of the queue, one per lane; corners at t -> score ring + corner list
__device__ __forceinline__ void score_round(WarpState& w, int t) {
    const int n = min(w.qt - w.qh, 32);
    int off = 4 * TS + 8, S = 0;
    uint32_t e = 0u;
    if (w.lane < n) {
        e = w.q[(w.qh + w.lane) & (QCAP - 1)];
        off = (int)(e & 0x1FFFu);
        S = fast_score(w.pix, off);
    }
    w.qh += n;
    const bool corner = S > t;
    const unsigned bal = __ballot_sync(FULL, corner);
    if (bal) {
        if (corner) {
            w.sc[off - 4 * TS] = (uint8_t)S;  // t < S <= 255
            const int pos = w.ncl + __popc(bal & w.lt);
            if (pos < CCAP) w.cl[pos] = e;
        }
        const int m = w.ncl + __popc(bal);
        if (m > CCAP) w.scanRowMax = max(w.scanRowMax, w.maxRow);
        w.ncl = min(m, CCAP);
    }
}

}  // namespace

__global__ void __launch_bounds__(32 * SW) k_fast_strip(const __grid_constant__ Geom g, const Buffers b, const __grid_constant__ StripMaps maps,
                                                        const FastStrip* __restrict__ strips, const uint8_t* __restrict__ aux, int nStrips, int nImages,
                                                        int* __restrict__ workCtr) {
    extern __shared__ __align__(128) unsigned char ssm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char* my = ssm + (size_t)wid * WARP_BYTES;
    WarpState w;
    w.pix = my;
    w.sc = my + PIX_BYTES;
    w.q = reinterpret_cast<uint32_t*>(my + PIX_BYTES + SC_BYTES);
    w.cl = w.q + QCAP;
    uint8_t* col = reinterpret_cast<uint8_t*>(w.cl + CCAP);
    uint64_t* bars = reinterpret_cast<uint64_t*>(col + COL_BYTES);
    w.lane = lane;
    w.lt = (1u << lane) - 1u;
    const uint32_t pixS = __shfl_sync(FULL, smem_u32(w.pix), 0);  // through a shuffle: ptxas otherwise re-derives the window address at every use
    const uint32_t one = __shfl_sync(FULL, 1u, 0);                // an opaque 1: x * one + y stays an IMAD (FMA pipe) instead of an ALU add
    if (lane == 0)
        for (int k = 0; k < 4; ++k) mbar_init(&bars[k], 1);
    mbar_fence_init();
    __syncwarp();
    uint32_t phaseBits = 0;
    const int nWork = nStrips * nImages;
    const int q4 = lane & 7, rsel = lane >> 3;  // a trip = four rows x eight groups of 16 pixels
    // the score rows a step wipes: 8 rows x 10 uint4, three per lane (lanes 16.. only two)
    int zr[3], zc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int idx = lane + 32 * k, r = idx / 10;
        zr[k] = __shfl_sync(FULL, r, lane);  // (through a shuffle: otherwise recomputed in every step)
        zc[k] = __shfl_sync(FULL, (idx - 10 * r) * 16, lane);
    }

    for (;;) {
        int wi = 0;
        if (lane == 0) wi = atom_add_global(workCtr, 1);
        wi = __shfl_sync(FULL, wi, 0);
        if (wi >= nWork) break;
        const int item = wi / nImages, imgRel = wi - item * nImages;  // all images' copies of the largest strips first
        const FastStrip d = strips[item];
        const int img = b.imgBase + imgRel;
        const LevelGeom& L = g.lv[d.level];
        const int rowEnd = d.rowEnd, nSteps = (rowEnd - 3 + 7) >> 3, nBoxes = (rowEnd + 3 + 7) >> 3;
        const int x0 = kMarginX + (int)d.iniX - (int)d.shift, y0 = kMarginY + kBorder + (int)d.rowBase;
        const int t = g.iniTh;
        const uint32_t kAdd = t < 128 ? (uint32_t)(0x7f - t) * 0x01010101u : 0u;

        auto issue = [&](int bx) {  // lane 0: box bx (pixel rows 8 bx .. 8 bx + 7 of the item) into its ring slot, plus the mirror rows it feeds
            const int slot = bx & 3;
            const bool mirror = slot == 0 || slot == 3;
            mbar_arrive_expect_tx(&bars[slot], (mirror ? 12 : 8) * TS);
            tma_load_3d(w.pix + (4 + 8 * slot) * TS, &maps.b8[d.level], &bars[slot], x0, y0 + 8 * bx, img);
            if (slot == 0) tma_load_3d(w.pix + (RING + 4) * TS, &maps.b4[d.level], &bars[slot], x0, y0 + 8 * bx, img);
            if (slot == 3) tma_load_3d(w.pix, &maps.b4[d.level], &bars[slot], x0, y0 + 8 * bx + 4, img);
        };
        auto wait_box = [&](int bx) {
            const int slot = bx & 3;
            mbar_wait(&bars[slot], (phaseBits >> slot) & 1u);
            phaseBits ^= 1u << slot;
        };

        // ---- per-item setup ----
        if (lane == 0) { issue(0); if (nBoxes > 1) issue(1); }
        for (int k = lane; k < SC_BYTES / 16; k += 32) reinterpret_cast<uint4*>(w.sc)[k] = make_uint4(0, 0, 0, 0);
        for (int k = lane; k < ((d.cw + 3) >> 2); k += 32) reinterpret_cast<uint32_t*>(col)[k] = __ldg(reinterpret_cast<const uint32_t*>(aux + d.aux) + k);
        // every cell of the item starts as "kept nothing" (2); a survivor's cell becomes 1 (ORBextractor.cc:817-820 is decided by these flags)
        uint8_t* flags = b.cellFlags + (size_t)img * g.cellsPerImage + d.cell0;
        if (lane < d.nCells)
            for (int r = 0; r < d.nCellRows; ++r) flags[r * d.nCols + lane] = 2;
        // byte masks of this lane's four words: bit 7 of a byte set iff its crop column is an interior column
        const int wb = 8 * d.win8 + 16 * q4;  // byte of this lane's first pixel inside a ring row
        uint32_t m[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            m[j] = 0u;
#pragma unroll
            for (int by = 0; by < 4; ++by) {
                const int c = wb + 4 * j + by - d.shift;
                if (c >= 3 && c < (int)d.cw - 3) m[j] |= 0x80u << (8 * by);
            }
        }
        w.qh = 0; w.qt = 0; w.ncl = 0; w.scanRowMax = -1;
        int nmsFrom = 3;  // first local row whose corners still wait for their NMS
        __syncwarp();

        // one corner per lane: it survives iff its score is strictly greater than the scores of its 8 neighbours inside the same
        // cell's interior; survivors -> global candidate list (coordinates relative to (16,16), ORBextractor.cc:837-842)
        uint32_t* cand = b.cand + L.candOff + (size_t)img * L.candCap;
        int* candCount = b.candCount + img * ORBX_MAX_LEVELS + d.level;
        const int candX0 = (int)d.iniX - kBorder - (int)d.shift;
        auto nms_group = [&](bool act, int cb, int row) {
            bool keep = false;
            int s = 0, info = 0, ci = 0;
            const int grow = (int)d.rowBase + row;  // row relative to the level's FAST region (level row - 16)
            if (act) {
                const uint8_t* pc = w.sc + (row & (RING - 1)) * TS + cb;
                const uint8_t* pu = w.sc + ((row - 1) & (RING - 1)) * TS + cb;
                const uint8_t* pd = w.sc + ((row + 1) & (RING - 1)) * TS + cb;
                s = pc[0];
                info = col[cb - d.shift];
                ci = (int)__umulhi((uint32_t)(grow - 3), d.magicH);
                const int ph = grow - 3 - ci * (int)d.hCell;
                int ul = pu[-1], uu = pu[0], ur = pu[1], ll = pc[-1], rr = pc[1], dl = pd[-1], dd = pd[0], dr = pd[1];
                // neighbours outside the cell's interior do not count (rim of the crop: cv::FAST never scores it)
                if (ph == 0) ul = uu = ur = 0;
                if (ph == (int)d.hCell - 1 || row == rowEnd - 1) dl = dd = dr = 0;
                if (info & 1) ul = ll = dl = 0;
                if (info & 2) ur = rr = dr = 0;
                const int mx = max(max(max(ul, uu), max(ur, ll)), max(max(rr, dl), max(dd, dr)));
                keep = s > mx;
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
                cand[base + __popc(bal & w.lt)] = pack_cand(cb + candX0, grow, s - 1);
                flags[(ci - (int)d.cellRow0) * (int)d.nCols + (info >> 2)] = 1;
            }
        };
        // NMS of the corners in local rows [nmsFrom, rowsEnd): from the list, or from the score ring for rows whose corners were dropped.
        // lazy: fewer than 32 list corners that could go wait for the next call unless `force`
        auto nms_rows = [&](int rowsEnd, bool force) {
            if (rowsEnd <= nmsFrom) return;
            if (w.scanRowMax >= nmsFrom) {
                force = true;
                const int scanEnd = min(rowsEnd, w.scanRowMax + 1);
                for (int row = nmsFrom; row < scanEnd; ++row) {
                    const int cb0 = 8 * d.win8 + 4 * lane;
                    const uint32_t v = *reinterpret_cast<const uint32_t*>(w.sc + (row & (RING - 1)) * TS + cb0);
#pragma unroll
                    for (int by = 0; by < 4; ++by) nms_group(((v >> (8 * by)) & 0xFFu) != 0u, cb0 + by, row);
                }
            }
            if (!force) {  // count the corners that could go now
                if (w.ncl < 32) return;
                int ready = 0;
                for (int k0 = 0; k0 < w.ncl; k0 += 32) {
                    const int k = k0 + lane;
                    ready += __popc(__ballot_sync(FULL, k < w.ncl && (int)(w.cl[k] >> 13) < rowsEnd));
                }
                if (ready < 32) return;
            }
            int kept = 0;
            for (int k0 = 0; k0 < w.ncl; k0 += 32) {
                const int k = k0 + lane;
                const bool valid = k < w.ncl;
                const uint32_t e = valid ? w.cl[k] : 0u;
                const int row = (int)(e >> 13), off = (int)(e & 0x1FFFu);
                const int cb = off - ((off * 6554) >> 20) * TS;   // off mod TS (exact for off < 8192)
                const bool scanned = row <= w.scanRowMax;  // handled (now or later) by the scan of its row
                nms_group(valid && !scanned && row < rowsEnd, cb, row);
                const bool stay = valid && !scanned && row >= rowsEnd;
                const unsigned bal = __ballot_sync(FULL, stay);
                if (stay) w.cl[kept + __popc(bal & w.lt)] = e;
                kept += __popc(bal);
                __syncwarp();
            }
            w.ncl = kept;
            nmsFrom = rowsEnd;
            if (w.scanRowMax < rowsEnd) w.scanRowMax = -1;
        };

        wait_box(0);
        for (int s = 0; s < nSteps; ++s) {
            if (s + 1 < nBoxes) wait_box(s + 1);
            __syncwarp();  // every lane is done with the slot box s + 2 lands in (pixels queued before step s - 1 are scored)
            if (lane == 0 && s + 2 < nBoxes) issue(s + 2);
            // the score rows this step writes still hold the rows 32 above: the NMS of the rows < 8 s - 18 must be complete
            if (nmsFrom < 8 * s - 18) nms_rows(min(8 * s - 6, rowEnd), true);
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (k < 2 || lane < 16) *reinterpret_cast<uint4*>(w.sc + ((8 * s + 3 + zr[k]) & (RING - 1)) * TS + zc[k]) = make_uint4(0, 0, 0, 0);
            w.maxRow = 8 * s + 10;
            __syncwarp();
            uint32_t H[2], ent[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int row = 8 * s + 3 + 4 * j + rsel;
                const uint32_t off = (uint32_t)(((row & (RING - 1)) + 4) * TS + wb);
                const uint32_t a = pixS + off;
                uint32_t c0, c1, c2, c3, u0, u1, u2, u3, d0, d1, d2, d3, pv, nx;
                asm volatile(
                    "ld.shared.v2.u32 {%0,%1}, [%14];\n\t"
                    "ld.shared.v2.u32 {%2,%3}, [%14+8];\n\t"
                    "ld.shared.v2.u32 {%4,%5}, [%14+-%15];\n\t"
                    "ld.shared.v2.u32 {%6,%7}, [%14+-%16];\n\t"
                    "ld.shared.v2.u32 {%8,%9}, [%14+%15];\n\t"
                    "ld.shared.v2.u32 {%10,%11}, [%14+%17];\n\t"
                    "ld.shared.u32 %12, [%14+-4];\n\t"
                    "ld.shared.u32 %13, [%14+16];"
                    : "=r"(c0), "=r"(c1), "=r"(c2), "=r"(c3), "=r"(u0), "=r"(u1), "=r"(u2), "=r"(u3), "=r"(d0), "=r"(d1), "=r"(d2), "=r"(d3), "=r"(pv),
                      "=r"(nx)
                    : "r"(a), "n"(3 * TS), "n"(3 * TS - 8), "n"(3 * TS + 8));
                // pixels at x + 3 / x - 3 of every word
                const uint32_t F0 = compass4(u0, d0, __funnelshift_r(c0, c1, 24), __funnelshift_r(pv, c0, 8), c0, kAdd, m[0], one);
                const uint32_t F1 = compass4(u1, d1, __funnelshift_r(c1, c2, 24), __funnelshift_r(c0, c1, 8), c1, kAdd, m[1], one);
                const uint32_t F2 = compass4(u2, d2, __funnelshift_r(c2, c3, 24), __funnelshift_r(c1, c2, 8), c2, kAdd, m[2], one);
                const uint32_t F3 = compass4(u3, d3, __funnelshift_r(c3, nx, 24), __funnelshift_r(c2, c3, 8), c3, kAdd, m[3], one);
                H[j] = F0 | (F1 >> 1) | (F2 >> 2) | (F3 >> 3);   // bit 8 * byte + 7 - word: pixel 4 * word + byte
                if (row >= rowEnd) H[j] = 0u;
                ent[j] = (off + 28u) | ((uint32_t)row << 13);
            }
            // the flagged pixels of the step -> queue: one scan over the lanes' counts, then every lane writes its own (most hold 0..2)
            {
                const int cnt = __popc(H[0]) + __popc(H[1]);
                int incl = cnt;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    int v;
                    asm("{\n\t.reg .pred p;\n\tshfl.sync.up.b32 %0|p, %1, %2, 0, 0xffffffff;\n\t@!p mov.b32 %0, 0;\n\t}" : "=r"(v) : "r"(incl), "r"(dd));
                    incl += v;
                }
                // lanes [lo, hi) go together: as many as the queue can take (all 32 unless the texture is dense); scoring in between
                int inclBase = 0, lo = 0;
                for (;;) {
                    const unsigned fits = __ballot_sync(FULL, incl - inclBase <= QCAP - 32);
                    const unsigned over = ~fits & ~((1u << lo) - 1u);
                    const int hi = over ? __ffs(over) - 1 : 32;   // > lo: one lane holds at most 32 pixels
                    const int endIncl = __shfl_sync(FULL, incl, hi - 1);
                    if (endIncl != inclBase) {
                        if (lane >= lo && lane < hi) {
                            int pos = w.qt + incl - cnt - inclBase;
#pragma unroll
                            for (int j = 0; j < 2; ++j)
                                while (H[j]) {
                                    const uint32_t bit = 31u - (uint32_t)__clz(H[j]);
                                    H[j] ^= 1u << bit;
                                    // pixel = 4 * (7 - (bit & 7)) + (bit >> 3) = 28 - 4 * bit + 33 * (bit >> 3)
                                    w.q[pos++ & (QCAP - 1)] = imad(bit >> 3, 33u, imad(bit, 0xFFFFFFFCu, ent[j]));
                                }
                        }
                        w.qt += endIncl - inclBase;
                        __syncwarp();
                        while (w.qt - w.qh >= 32) score_round(w, t);
                    }
                    if (hi == 32) break;
                    inclBase = endIncl;
                    lo = hi;
                }
            }
            // pixels queued before this step must be scored now: the next step reloads the ring slot below them
            while (w.qt != w.qh && (int)(w.q[w.qh & (QCAP - 1)] >> 13) < 8 * s + 3) score_round(w, t);
            // rows up to 8 s + 2 are completely scored (what is queued is this step's): their corners up to row 8 s + 1 can go
            if (s + 1 < nSteps) nms_rows(min(8 * s + 2, rowEnd), false);
        }
        while (w.qt != w.qh) score_round(w, t);
        __syncwarp();
        nms_rows(rowEnd, true);
        __syncwarp();  // the ring and the tables are free for the next item
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// host: the strip table of one geometry (ORBextractor.cc:759-796 cell grid, cut into column strips x cell-row ranges)
// ---------------------------------------------------------------------------------------------------------------------------
int make_fast_strips(const Geom& g, bool wholeStrips, std::vector<FastStrip>& strips, std::vector<uint8_t>& auxOut, std::vector<uint8_t>& flagTemplate) {
    strips.clear();
    auxOut.clear();
    flagTemplate.assign((size_t)g.cellsPerImage, 0);
    for (int l = 0; l < g.nlevels; ++l) {
        const LevelGeom& L = g.lv[l];
        const int maxBX = L.w - kBorder, maxBY = L.h - kBorder;
        const int HR = L.h - 2 * kBorder;
        // cells with FAST work form a prefix of every cell row / column (ORBextractor.cc:785-796)
        int nValidC = 0, nValidR = 0;
        for (int j = 0; j < L.nCols; ++j)
            if (kBorder + j * L.wCell < maxBX - 6) nValidC = j + 1;
        for (int i = 0; i < L.nRows; ++i) {
            const int iniY = kBorder + i * L.hCell;
            if (iniY < maxBY - 3 && std::min(iniY + L.hCell + 6, maxBY) - iniY >= 7) nValidR = i + 1;
        }
        for (int i = 0; i < L.nRows; ++i)
            for (int j = 0; j < L.nCols; ++j) {
                const bool skipped = kBorder + i * L.hCell >= maxBY - 3 || kBorder + j * L.wCell >= maxBX - 6;
                // 0: the reference skips the cell; 2: cv::FAST yields nothing at either threshold (crop thinner than the ring);
                // cells with work are rewritten by the kernels on every run
                flagTemplate[(size_t)L.cellOff + i * L.nCols + j] = skipped ? 0 : 2;
            }
        if (nValidC == 0 || nValidR == 0) continue;
        // column strips: as many cells as fit 32 words from the (8-byte aligned) window start
        int c0 = 0;
        while (c0 < nValidC) {
            const int iniX = kBorder + c0 * L.wCell;
            const int shift = (kMarginX + iniX) & 15, win8 = (shift + 3) >> 3;
            int c1 = c0;
            int cw = 0;
            while (c1 < nValidC && c1 - c0 < 63) {
                const int cwTry = std::min(kBorder + c1 * L.wCell + L.wCell + 6, maxBX) - iniX;
                if (shift + cwTry - 3 > 8 * win8 + 128) break;  // the last interior byte must lie inside the 128-byte window
                cw = cwTry;
                ++c1;
            }
            if (c1 == c0) return -1;  // a single cell wider than the window (wCell > ~120)
            while (auxOut.size() % 16) auxOut.push_back(0);
            const uint32_t auxOff = (uint32_t)auxOut.size();
            for (int col = 0; col < cw; ++col) {
                uint8_t info = 0xFF;
                if (col >= 3 && col < cw - 3) {
                    const int lc = (col - 3) / L.wCell, ph = (col - 3) % L.wCell;
                    info = (uint8_t)((ph == 0 ? 1 : 0) | ((ph == L.wCell - 1 || col == cw - 4) ? 2 : 0) | (lc << 2));
                }
                auxOut.push_back(info);
            }
            while (auxOut.size() % 4) auxOut.push_back(0xFF);
            const int rowsPerItem = wholeStrips ? nValidR : 1;
            for (int r0 = 0; r0 < nValidR; r0 += rowsPerItem) {
                const int r1 = std::min(r0 + rowsPerItem, nValidR);
                FastStrip d;
                std::memset(&d, 0, sizeof(d));
                d.iniX = (uint16_t)iniX;
                d.cw = (uint16_t)cw;
                d.rowBase = (uint16_t)(r0 * L.hCell);                                 // local row 0 = first output row - 3
                d.rowEnd = (uint16_t)(std::min(3 + r1 * L.hCell, HR - 3) - r0 * L.hCell);
                d.cell0 = (uint16_t)(L.cellOff + r0 * L.nCols + c0);
                d.nCols = (uint16_t)L.nCols;
                d.level = (uint8_t)l; d.nCells = (uint8_t)(c1 - c0); d.shift = (uint8_t)shift; d.win8 = (uint8_t)win8;
                d.nCellRows = (uint8_t)(r1 - r0); d.cellRow0 = (uint8_t)r0;
                d.aux = auxOff;
                d.hCell = (uint32_t)L.hCell;
                d.magicH = (0xFFFFFFFFu / (uint32_t)L.hCell) + 1;
                strips.push_back(d);
            }
            c0 = c1;
        }
    }
    if (strips.empty()) return -1;
    // largest items first: the dynamic queue then ends with the small ones
    std::stable_sort(strips.begin(), strips.end(), [](const FastStrip& a, const FastStrip& c) { return (int)a.rowEnd * a.cw > (int)c.rowEnd * c.cw; });
    return 0;
}

int fast_strip_encode_maps(const Geom& g, uint8_t* pyr, int nImages, StripMaps* m) {
    for (int l = 0; l < g.nlevels; ++l) {
        const LevelGeom& L = g.lv[l];
        const size_t blockOff = (size_t)L.planeOff - (size_t)kMarginY * L.pitch - kMarginX;
        if (tma_encode_3d(&m->b8[l], pyr + blockOff, L.pitch, L.h + 2 * kMarginY, nImages, L.planeBytes, TS, 8)) return -1;
        if (tma_encode_3d(&m->b4[l], pyr + blockOff, L.pitch, L.h + 2 * kMarginY, nImages, L.planeBytes, TS, 4)) return -1;
    }
    return 0;
}

cudaError_t fast_strip_configure() {
    int dev = 0, optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_fast_strip, cudaFuncAttributeMaxDynamicSharedMemorySize, std::min(optin, std::max(SW * WARP_BYTES, 48 * 1024)));
}

void launch_fast_strip(const Geom& g, const Buffers& b, const StripMaps& maps, const FastStrip* strips, const uint8_t* aux, int nStrips, int nImages, int numSMs,
                       int* workCtr, cudaStream_t s) {
    const size_t smem = (size_t)SW * WARP_BYTES;
    const int ctasPerSM = std::max(1, std::min(32, (int)((227 * 1024) / (smem + 1024))));
    const int nWork = nStrips * nImages;
    const int grid = std::min((nWork + SW - 1) / SW, numSMs * ctasPerSM);
    cudaMemsetAsync(workCtr, 0, sizeof(int), s);
    k_fast_strip<<<grid, 32 * SW, smem, s>>>(g, b, maps, strips, aux, nStrips, nImages, workCtr);
}

}  // namespace orbx
