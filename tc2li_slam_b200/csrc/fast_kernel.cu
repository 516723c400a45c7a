// fast_kernel.cu — cell-grid FAST-9/16 with the iniThFAST -> minThFAST per-cell fallback.
//
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree (ORBextractor.cc:779-846) and the cv::FAST(cell, kps, th,
// true) calls inside it (ORBextractor.cc:800,819; arithmetic: SURVEY.md A.4).
//
// Formulation (derived from the OpenCV semantics, not from its code): for a pixel with centre value v and ring r[0..15]
//     S = max( max_s min_{j<9} (r[s+j] - v),  max_s min_{j<9} (v - r[s+j]) )
// the pixel is a corner at threshold t iff S > t, cv::FAST reports response S-1 whatever t is, and 3x3 non-maximum
// suppression inside the cell crop keeps it iff S is strictly greater than the S of its 8 neighbours that lie in the crop's
// interior (3-px rim excluded).  Pixels that are not corners at t cannot suppress one that is, so scoring only the corners at
// t (others read as 0) is exact.
//
// Execution model: ONE WARP PER CELL, persistent warps.  A warp walks a strided list of (image, cell) work items; lane 0
// fetches the cell crop with one TMA box load (tensor map per level, 16-byte aligned box origin, 64- or 96-byte box rows)
// into the warp's own shared-memory tile.  The tile is single buffered (NB = 1): prefetching the next cell's box into a
// second tile measured slower (18.3 vs 17.5 ms per 8192 KITTI images), because shared memory per warp sets the number of
// resident warps (15 vs 20 per SM) and the other warps hide the box latency better than the prefetch did.  No block-level
// barrier exists; every pass is warp-synchronous and compacts its survivors in order with ballot prefixes (no atomics):
//   A  four pixels per lane, packed in one 32-bit word: |ring - centre| > t at the four compass points, combined as
//      (N|S)&(E|W) — a necessary condition for a 9-arc of either polarity (VABSDIFF4 + carry tricks) -> listW.  Lanes are
//      laid out as (row of a group of 32 / nW rows, word column of the interior): column, edge mask and row phase are per-lane
//      constants of the cell and a trip advances by one add (47 instead of 65 instructions per word: 16.3 -> 15.4 ms);
//   A2 dense over list1: the exact score S of both polarities at once, ring pixels q and q + 8 paired in 16x2 lanes and
//      the nine-windows built from three-input packed min / max (VIMNMX3.U16x2); S > t -> list2 (corners) and the score
//      goes into a score map that shares the tile's layout;
//   C  dense over list2: cell-local 3x3 NMS; survivors go straight to the global candidate list (one atomic per batch).
// The threshold is iniThFAST first; only a cell without any survivor is redone at minThFAST (ORBextractor.cc:817-820).
// The lists have a fixed capacity; a pass that would overflow flushes the downstream passes first, and a cell whose corner
// list was flushed early falls back to scanning the score map for the NMS (textures such as saturated noise).
// Integer only; HBM traffic = one read of every level, the work is ALU/issue bound.
#include <algorithm>
#include <type_traits>

#include "orbx_internal.cuh"

namespace orbx {

namespace {

constexpr int FW = 1;                    // warps per CTA (every warp is independent; small CTAs pack shared memory best)
#ifndef FAST_CAP
#define FAST_CAP 512
#endif
#ifndef FAST_CAPW
#define FAST_CAPW 256
#endif
constexpr int CAP = FAST_CAP;                 // entries of list1 / list2
constexpr int CAPW = FAST_CAPW;                // entries of listW (32-bit)
#ifndef FAST_NI
#define FAST_NI 2
#endif
constexpr int NB = 1;                    // tile buffers per warp (2 = the next cell's crop is prefetched)
constexpr unsigned FULL = 0xFFFFFFFFu;

// per byte: bit 7 set where the byte of d exceeds t (0 <= t <= 254); kAdd = (0x7f - (t & 0x7f)) * 0x01010101,
// nh = t >= 128 ? 0 : ~0.  low = bit 7 of each byte: (d & 0x7f) > (t & 0x7f); t >= 128 also needs d >= 128 (low & d), below
// that d >= 128 suffices (low | d): both are maj(low, d, nh), one LOP3
__device__ __forceinline__ uint32_t bytes_gt(uint32_t d, uint32_t kAdd, uint32_t nh) {
    const uint32_t low = (d & 0x7f7f7f7fu) + kAdd;
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(low), "r"(d), "r"(nh));
    return r;
}

// byte offsets of the 16 ring pixels inside a tile with row stride TS
#define RING(TS, k) ((k) == 0 ? 3 * (TS) : (k) == 1 ? 3 * (TS) + 1 : (k) == 2 ? 2 * (TS) + 2 : (k) == 3 ? (TS) + 3 : (k) == 4 ? 3 : (k) == 5 ? -(TS) + 3 : \
                     (k) == 6 ? -2 * (TS) + 2 : (k) == 7 ? -3 * (TS) + 1 : (k) == 8 ? -3 * (TS) : (k) == 9 ? -3 * (TS) - 1 : (k) == 10 ? -2 * (TS) - 2 : \
                     (k) == 11 ? -(TS) - 3 : (k) == 12 ? -3 : (k) == 13 ? (TS) - 3 : (k) == 14 ? 2 * (TS) - 2 : 3 * (TS) - 1)

struct CellCtx {
    const uint8_t* tile;   // crop byte (r, c) at tile[r * TS + shift + c]
    uint8_t* sc;           // score map, same layout
    uint32_t* listW;       // words with a compass survivor: tile word index | 4 pixel flags << 24
    uint16_t* list1;       // compass survivors (tile byte offsets); NMS survivors afterwards
    uint16_t* list2;       // corners (tile byte offsets)
    int nw, n1, n2;        // warp-uniform fill levels
    bool flushedB;         // list2 overflowed and restarted: corners are no longer all listed
    int lane;
};

// list1 -> list2 with the exact score of every listed pixel (16x2 lanes, DPX three-input min / max):
//   S = max( max_s min_{j<9} r[s+j] - v,  v - min_s max_{j<9} r[s+j] )        (at most one of the two terms is positive)
// R[q] holds ring pixels q (low half) and q + 8 (high half), R[q + 8] is its half swap, so the nine-window starting at q
// yields the windows starting at q and q + 8 at once: m3[q] = op3(R[q], R[q+1], R[q+2]), W[q] = op3(m3[q], m3[q+3], m3[q+6]).
// A pixel is a corner at t iff S > t; only corners get their score written to the map.
template <int TS>
__device__ __forceinline__ void flush_A2(CellCtx& c, int t) {
    for (int k0 = 0; k0 < c.n1; k0 += 32) {
        const int k = k0 + c.lane;
        int off = 0, S = 0;
        if (k < c.n1) {
            off = c.list1[k];
            const uint8_t* p = c.tile + off;
            const int v = p[0];
            // 16x2 lanes [ring, 255 - ring]: the minimum over a window is (min r, 255 - max r), both polarities without half swaps
            uint32_t R[18];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                uint32_t x;
                asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(x) : "r"((uint32_t)p[RING(TS, q)]), "r"(0xFFFF0001u), "r"(0x00FF0000u));  // IMAD: FMA pipe
                R[q] = x;
            }
            R[16] = R[0];
            R[17] = R[1];
            uint32_t m3[22];
#pragma unroll
            for (int q = 0; q < 16; ++q) m3[q] = __vimin3_u16x2(R[q], R[q + 1], R[q + 2]);
#pragma unroll
            for (int q = 0; q < 6; ++q) m3[16 + q] = m3[q];
            uint32_t W9[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) W9[q] = __vimin3_u16x2(m3[q], m3[q + 3], m3[q + 6]);
            const uint32_t a0 = __vimax3_u16x2(W9[0], W9[1], W9[2]), a1 = __vimax3_u16x2(W9[3], W9[4], W9[5]), a2 = __vimax3_u16x2(W9[6], W9[7], W9[8]);
            const uint32_t a3 = __vimax3_u16x2(W9[9], W9[10], W9[11]), a4 = __vimax3_u16x2(W9[12], W9[13], W9[14]);
            const uint32_t best = __vmaxu2(__vimax3_u16x2(a0, a1, a2), __vimax3_u16x2(a3, a4, W9[15]));
            S = max((int)(best & 0xFFFFu) - v, (int)(best >> 16) + v - 255);   // max_s min r - v,  v - min_s max r
        }
        const bool corner = S > t;
        const unsigned bal = __ballot_sync(FULL, corner);
        if (bal) {
            if (c.n2 + 32 > CAP) {  // warp-uniform: the scores stay in the map, the list restarts
                c.n2 = 0;
                c.flushedB = true;
            }
            if (corner) {
                c.sc[off] = (uint8_t)S;  // t < S <= 255
                c.list2[c.n2 + __popc(bal & ((1u << c.lane) - 1))] = (uint16_t)off;
            }
            c.n2 += __popc(bal);
        }
    }
    __syncwarp();
    c.n1 = 0;
}

// listW -> list1: one entry per flagged pixel, in order (exclusive prefix of the per-lane counts from the ballots of their bits)
template <int TS>
__device__ __forceinline__ void flush_W(CellCtx& c, int t) {
    for (int k0 = 0; k0 < c.nw; k0 += 32) {
        const int k = k0 + c.lane;
        const uint32_t e = k < c.nw ? c.listW[k] : 0u;
        const int f4 = (int)(e >> 24), off = (int)(e & 0xFFFu) * 4;
        const int cnt = __popc(f4);
        if (c.n1 + 128 > CAP) flush_A2<TS>(c, t);  // warp-uniform
        const unsigned q0 = __ballot_sync(FULL, cnt & 1), q1 = __ballot_sync(FULL, cnt & 2), q2 = __ballot_sync(FULL, cnt & 4);
        const unsigned lt = (1u << c.lane) - 1;
        int pos = c.n1 + __popc(q0 & lt) + 2 * __popc(q1 & lt) + 4 * __popc(q2 & lt);
        if (f4 & 1) c.list1[pos++] = (uint16_t)off;
        if (f4 & 2) c.list1[pos++] = (uint16_t)(off + 1);
        if (f4 & 4) c.list1[pos++] = (uint16_t)(off + 2);
        if (f4 & 8) c.list1[pos] = (uint16_t)(off + 3);
        c.n1 += __popc(q0) + 2 * __popc(q1) + 4 * __popc(q2);
    }
    __syncwarp();
    c.nw = 0;
}

// survivors (tile offsets in `surv`) -> global candidate list, coordinates relative to (16,16) (ORBextractor.cc:837-842)
template <int TS>
__device__ __forceinline__ void emit(const CellCtx& c, const uint16_t* surv, int ns, int shift, int relX, int relY, uint32_t* cand, int candCap,
                                     int* candCount, int* errFlag) {
    if (ns == 0) return;
    int base = 0;
    if (c.lane == 0) base = atomicAdd(candCount, ns);
    base = __shfl_sync(FULL, base, 0);
    if (base + ns > candCap) {
        if (c.lane == 0) atomicExch(errFlag, 1);
        return;
    }
    for (int k = c.lane; k < ns; k += 32) {
        const int off = surv[k];
        const int r = off / TS, cx = off - r * TS - shift;  // crop coordinates = the reference's cell-local keypoint position
        cand[base + k] = pack_cand(cx + relX, r + relY, (int)c.sc[off] - 1);
    }
}

}  // namespace

template <int TS>
__global__ void __launch_bounds__(32 * FW) k_fast_cells(const __grid_constant__ Geom g, const Buffers b, const __grid_constant__ FastMaps maps,
                                                       int nCells, int nWork, int tileBytes) {
    extern __shared__ __align__(128) unsigned char fsm[];
    __shared__ __align__(8) uint64_t bars[FW][2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warpBytes = (NB + 1) * tileBytes + CAPW * 4 + 2 * CAP * 2;
    unsigned char* my = fsm + (size_t)wid * warpBytes;
    CellCtx c;
    c.sc = my + NB * tileBytes;
    c.listW = reinterpret_cast<uint32_t*>(my + (NB + 1) * tileBytes);
    c.list1 = reinterpret_cast<uint16_t*>(c.listW + CAPW);
    c.list2 = c.list1 + CAP;
    c.lane = lane;
    for (int k = lane; k < tileBytes / 16; k += 32) reinterpret_cast<uint4*>(c.sc)[k] = make_uint4(0, 0, 0, 0);  // stays all-zero between cells
    if (lane == 0) { mbar_init(&bars[wid][0], 1); mbar_init(&bars[wid][1], 1); }
    mbar_fence_init();
    __syncwarp();
    uint32_t ph0 = 0, ph1 = 0;

    const int stride = gridDim.x * FW;
    auto issue = [&](int img, int cell, int buf) {  // lane 0: the crop of work item (img, cell) into tile buffer `buf` (cells without FAST work load nothing)
        const uint4 cd = __ldg(b.cellTab + cell);
        if ((cd.y >> 24) != 2) return;
        const int iniX = cd.x & 0xFFFF, iniY = cd.x >> 16, level = (cd.y >> 16) & 0xFF;
        mbar_arrive_expect_tx(&bars[wid][buf], TS * (g.lv[level].hCell + 6));
        tma_load_3d(my + buf * tileBytes, &maps.m[level], &bars[wid][buf], (kMarginX + iniX) & ~15, kMarginY + iniY, b.imgBase + img);
    };
    // one cell: wait for its crop, FAST + NMS at iniThFAST (attempt 0) and / or minThFAST (attempt 1), candidates and the cell's flag
    auto run_cell = [&](int imgRel, int cellIdx, int buf, int nextImg, int nextCell) {   // nextImg < 0: no next item
        const int img = b.imgBase + imgRel;
        const uint4 cd = __ldg(b.cellTab + cellIdx);
        const int iniX = cd.x & 0xFFFF, iniY = cd.x >> 16;
        const int cw = cd.y & 0xFF, ch = (cd.y >> 8) & 0xFF, level = (cd.y >> 16) & 0xFF, kind = cd.y >> 24;
        uint8_t* flag = b.cellFlags + (size_t)img * g.cellsPerImage + cellIdx;
        if (kind != 2) {  // 0: a cell the reference skips (ORBextractor.cc:785-796); 1: crop thinner than the ring, cv::FAST yields nothing
            if (lane == 0) {
                *flag = kind == 0 ? 0 : 2;
                if (NB == 1 && nextImg >= 0) issue(nextImg, nextCell, 0);
            }
            return;
        }
        const LevelGeom& L = g.lv[level];
        const int iw = cw - 6, ih = ch - 6;
        const int nW = cd.w & 0xFF, wFirst = (cd.w >> 8) & 0xFF, shift = (cd.w >> 16) & 0xFF;
        const uint32_t magicW = cd.z;
        c.tile = my + buf * tileBytes;
        const uint32_t* tileW = reinterpret_cast<const uint32_t*>(c.tile);
        if (buf == 0) { mbar_wait(&bars[wid][0], ph0); ph0 ^= 1; }
        else { mbar_wait(&bars[wid][1], ph1); ph1 ^= 1; }

        uint32_t* cand = b.cand + L.candOff + (size_t)img * L.candCap;
        int* candCount = b.candCount + img * ORBX_MAX_LEVELS + level;
        const int relX = iniX - kBorder, relY = iniY - kBorder;
        int nsTotal = 0;
        bool issuedNext = false;
        int attempt = 0;
        for (;; ++attempt) {
            const int t = attempt == 0 ? g.iniTh : g.minTh;
            c.nw = 0; c.n1 = 0; c.n2 = 0; c.flushedB = false;

            // ---- pass A: compass test, 4 pixels per lane; words with a survivor are listed with their 4 pixel flags ----
            {
                // Both values go through a shuffle so that ptxas keeps them in registers: it otherwise re-derives kAdd with a move +
                // IMAD at each of its four uses and the shared-window base with four uniform instructions per trip.
                const uint32_t kAdd = __shfl_sync(FULL, (uint32_t)(0x7f - (t & 0x7f)) * 0x01010101u, 0);
                const uint32_t nh = t >= 128 ? 0u : 0xFFFFFFFFu;
                const uint32_t tileS = __shfl_sync(FULL, smem_u32(tileW), 0);  // shared-window address for the explicit ld.shared below
                // bytes outside the interior [0, iw) can only sit in the first / last word of a row
                const int xb0 = 4 * wFirst - 3 - shift;  // interior x of byte 0 of the first word (<= 0)
                const uint32_t maskFirst = 0x80808080u << (8 * (-xb0));
                const uint32_t maskLast = 0x80808080u >> (8 * (xb0 + 4 * nW - iw));
                // Lane = (row r of a group of RT = 32 / nW rows, word col): the word column, its edge mask and the row phase are
                // per-lane constants of the cell, a trip advances every lane by RT rows (one add), and a row past the interior
                // only zeroes the lane's mask (the reads stay inside the warp's own shared memory: the score map follows the tile).
                // (RT <= 4 bounds the rows read past the interior; lanes beyond RT * nW idle on row 0 with an empty mask)
                const int RT = nW == 1 ? 4 : min(4, (int)__umulhi(32u, magicW));
                int r = nW == 1 ? lane : (int)__umulhi((uint32_t)lane, magicW), col = lane - r * nW;
                const bool laneOn = r < RT;
                if (!laneOn) { r = 0; col = 0; }
                const uint32_t mLane = laneOn ? (col == 0 ? maskFirst : 0x80808080u) & (col == nW - 1 ? maskLast : 0x80808080u) : 0u;
                constexpr int NI = FAST_NI;
                const int dW = RT * (TS / 4);
                int y = r, wofs = (r + 3) * (TS / 4) + wFirst + col;
                // LOWT (t < 128, every configuration the reference ships): bytes_gt is low | d, so (N|S) & (E|W) & m folds into four
                // three-input LOP3 instead of four majorities and two combines
                auto compass = [&](auto LOWT, uint32_t m, int wo) -> uint32_t {
                    uint32_t cur, prev, next, up, dn;
                    asm volatile("ld.shared.u32 %0, [%5];\n\tld.shared.u32 %1, [%5+-4];\n\tld.shared.u32 %2, [%5+4];\n\tld.shared.u32 %3, [%5+%6];\n\tld.shared.u32 %4, [%5+-%6];"
                                 : "=r"(cur), "=r"(prev), "=r"(next), "=r"(up), "=r"(dn)
                                 : "r"(tileS + 4u * (uint32_t)wo), "n"(3 * TS));
                    const uint32_t e = __funnelshift_r(cur, next, 24);  // pixels at x + 3
                    const uint32_t wv = __funnelshift_r(prev, cur, 8);  // pixels at x - 3
                    if (decltype(LOWT)::value) {
                        const uint32_t dN = __vabsdiffu4(up, cur), dS = __vabsdiffu4(dn, cur), dE = __vabsdiffu4(e, cur), dW4 = __vabsdiffu4(wv, cur);
                        // (a carry out of a byte >= 128 can only ADD a survivor one byte up, at d == t there: the exact score of pass A2 decides)
                        const uint32_t lN = dN + kAdd, lS = dS + kAdd, lE = dE + kAdd, lW = dW4 + kAdd;
                        const uint32_t a = lN | dN | lS, bq = lE | dE | lW;
                        return ((a | dS) & m) & (bq | dW4);
                    }
                    const uint32_t ns = bytes_gt(__vabsdiffu4(up, cur), kAdd, nh) | bytes_gt(__vabsdiffu4(dn, cur), kAdd, nh);
                    const uint32_t ew = bytes_gt(__vabsdiffu4(e, cur), kAdd, nh) | bytes_gt(__vabsdiffu4(wv, cur), kAdd, nh);
                    return ns & ew & m;
                };
                auto append = [&](uint32_t F, int wo) {  // (the caller has made room for 32 entries)
                    const unsigned bal = __ballot_sync(FULL, F != 0);
                    if (bal) {
                        // bits 7/15/23/31 -> 4 contiguous flag bits 24..27
                        if (F) c.listW[c.nw + __popc(bal & ((1u << lane) - 1))] = (((F >> 7) * 0x01020408u) & 0x0F000000u) | (uint32_t)wo;
                        c.nw += __popc(bal);
                    }
                };
                // full trips (every lane's NI rows lie inside the interior) need no per-row mask select; the last one or two do
                auto passA = [&](auto LOWT) {
                    int base = 0;
                    const int yFullMax = ih - NI * RT;  // largest trip base whose NI * RT rows all lie inside the interior (rows r = 0 .. RT - 1 are all in use)
                    for (; base <= yFullMax; base += RT * NI) {
                        if (c.nw + 32 * NI > CAPW) flush_W<TS>(c, t);  // warp-uniform, once per trip
                        uint32_t F[NI];
#pragma unroll
                        for (int j = 0; j < NI; ++j) F[j] = compass(LOWT, mLane, wofs + j * dW);
#pragma unroll
                        for (int j = 0; j < NI; ++j) append(F[j], wofs + j * dW);
                        y += NI * RT;
                        wofs += NI * dW;
                    }
                    for (; base < ih; base += RT * NI) {
                        if (c.nw + 32 * NI > CAPW) flush_W<TS>(c, t);
                        uint32_t F[NI];
#pragma unroll
                        for (int j = 0; j < NI; ++j) F[j] = compass(LOWT, y + j * RT < ih ? mLane : 0u, wofs + j * dW);
#pragma unroll
                        for (int j = 0; j < NI; ++j) append(F[j], wofs + j * dW);
                        y += NI * RT;
                        wofs += NI * dW;
                    }
                };
                if (t < 128) passA(std::true_type());
                else passA(std::false_type());
                __syncwarp();
            }
            flush_W<TS>(c, t);
            flush_A2<TS>(c, t);

            // ---- pass C: cell-local 3x3 non-maximum suppression (only interior pixels ever get a score: the rim reads 0) ----
            int ns = 0;                  // survivors waiting in list1
            uint16_t* surv = c.list1;    // list1 is free now
            if (!c.flushedB) {
                for (int k0 = 0; k0 < c.n2; k0 += 32) {
                    const int k = k0 + lane;
                    bool keep = false;
                    int off = 0;
                    if (k < c.n2) {
                        off = c.list2[k];
                        const uint8_t* q = c.sc + off;
                        const int s = q[0];
                        const int nmax = max(max(max((int)q[-1], (int)q[1]), max((int)q[-TS - 1], (int)q[-TS])), max(max((int)q[-TS + 1], (int)q[TS - 1]), max((int)q[TS], (int)q[TS + 1])));
                        keep = s > nmax;  // all eight loads issue together (no short-circuit chain)
                    }
                    const unsigned bal = __ballot_sync(FULL, keep);
                    if (keep) surv[ns + __popc(bal & ((1u << lane) - 1))] = (uint16_t)off;
                    ns += __popc(bal);
                }
                __syncwarp();
                // the tile is free once no second attempt can follow: the next cell's crop starts to load under the rest of this cell
                const bool last = ns > 0 || attempt == 1 || g.iniTh == g.minTh;
                if (NB == 1 && last && lane == 0 && nextImg >= 0) { issue(nextImg, nextCell, 0); issuedNext = true; }
                if (ns > 0 && ns <= 32) {
                    // one group of survivors (the usual case): the counter's round trip hides behind the wipe of the score map
                    uint32_t packed = 0u;
                    if (lane < ns) {
                        const int off = surv[lane];
                        const int r = off / TS, cx = off - r * TS - shift;
                        packed = pack_cand(cx + relX, r + relY, (int)c.sc[off] - 1);
                    }
                    int base = 0;
                    if (lane == 0) base = atomicAdd(candCount, ns);
                    __syncwarp();
                    for (int k = lane; k < c.n2; k += 32) c.sc[c.list2[k]] = 0;  // leave the score map clean
                    base = __shfl_sync(FULL, base, 0);
                    if (base + ns > L.candCap) {
                        if (lane == 0) atomicExch(b.errFlag, 1);
                    } else if (lane < ns)
                        cand[base + lane] = packed;
                } else {
                    emit<TS>(c, surv, ns, shift, relX, relY, cand, L.candCap, candCount, b.errFlag);
                    __syncwarp();
                    for (int k = lane; k < c.n2; k += 32) c.sc[c.list2[k]] = 0;  // leave the score map clean
                }
                nsTotal += ns;
            } else {
                // the corner list was flushed early (very dense corners): find the scores by scanning the map
                for (int p0 = 0; p0 < iw * ih; p0 += 32) {
                    const int p = p0 + lane;
                    bool keep = false;
                    int off = 0;
                    if (p < iw * ih) {
                        const int y = p / iw, x = p - y * iw;
                        off = (y + 3) * TS + shift + 3 + x;
                        const uint8_t* q = c.sc + off;
                        const int s = q[0];
                        keep = s > 0 && s > q[-1] && s > q[1] && s > q[-TS - 1] && s > q[-TS] && s > q[-TS + 1] && s > q[TS - 1] && s > q[TS] && s > q[TS + 1];
                    }
                    const unsigned bal = __ballot_sync(FULL, keep);
                    if (bal) {
                        if (ns + 32 > CAP) {
                            __syncwarp();
                            emit<TS>(c, surv, ns, shift, relX, relY, cand, L.candCap, candCount, b.errFlag);
                            nsTotal += ns;
                            ns = 0;
                            __syncwarp();
                        }
                        if (keep) surv[ns + __popc(bal & ((1u << lane) - 1))] = (uint16_t)off;
                        ns += __popc(bal);
                    }
                }
                __syncwarp();
                emit<TS>(c, surv, ns, shift, relX, relY, cand, L.candCap, candCount, b.errFlag);
                nsTotal += ns;
                __syncwarp();
                for (int k = lane; k < ch * (TS / 16); k += 32) reinterpret_cast<uint4*>(c.sc)[k] = make_uint4(0, 0, 0, 0);
            }
            __syncwarp();
            if (nsTotal > 0 || attempt == 1 || g.iniTh == g.minTh) break;
        }
        if (lane == 0) {
            *flag = (attempt == 0) ? 1 : 2;
            if (NB == 1 && !issuedNext && nextImg >= 0) issue(nextImg, nextCell, 0);   // (the dense-corner path: every lane is past the tile here)
        }
        __syncwarp();
    };
    int w = blockIdx.x * FW + wid;
    // (image, cell) of the work item, walked incrementally: `stride` items further is dImg images and dCell cells on
    const int dImg = stride / nCells, dCell = stride - dImg * nCells;
    int imgRel = w / nCells, cellIdx = w - imgRel * nCells;
    if (lane == 0 && w < nWork) issue(imgRel, cellIdx, 0);
    for (int it = 0; w < nWork; w += stride, ++it, imgRel += dImg, cellIdx += dCell) {
        const int buf = NB == 2 ? (it & 1) : 0;
        if (cellIdx >= nCells) { cellIdx -= nCells; ++imgRel; }
        if (NB == 2) {
            if (lane == 0 && w + stride < nWork) {
                const int c2 = cellIdx + dCell, wrap = c2 >= nCells;
                issue(imgRel + dImg + wrap, c2 - (wrap ? nCells : 0), buf ^ 1);
            }
        }
        if (cellIdx >= nCells) { cellIdx -= nCells; ++imgRel; }
        int ni = -1, nc = 0;
        if (w + stride < nWork) {
            nc = cellIdx + dCell;
            ni = imgRel + dImg + (nc >= nCells ? 1 : 0);
            if (nc >= nCells) nc -= nCells;
        }
        run_cell(imgRel, cellIdx, buf, ni, nc);
    }
}

// one descriptor per cell of every level (ORBextractor.cc:779-796), built once per geometry on the host.
// x = iniX | iniY << 16;  y = cw | ch << 8 | level << 16 | kind << 24;  z = magic for / nW;  w = nW | wFirst << 8 | shift << 16
// (shift = byte of the crop's first column inside its 16-byte aligned box row; words are counted from the box row start)
void make_cell_table(const Geom& g, uint4* tab, int* tileRows, int* maxCw) {
    int rows = 0, mcw = 0;
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
                const int shift = (kMarginX + iniX) & 15, iw = cw - 6;
                const int wFirst = (3 + shift) >> 2, nW = kind == 2 ? ((3 + shift + iw - 1) >> 2) - wFirst + 1 : 1;
                d.x = (uint32_t)iniX | ((uint32_t)iniY << 16);
                d.y = (uint32_t)(cw & 0xFF) | ((uint32_t)(ch & 0xFF) << 8) | ((uint32_t)l << 16) | ((uint32_t)kind << 24);
                d.z = (0xFFFFFFFFu / (uint32_t)nW) + 1;  // exact it / nW for it < 2^16
                d.w = (uint32_t)nW | ((uint32_t)wFirst << 8) | ((uint32_t)shift << 16);
                tab[L.cellOff + i * L.nCols + j] = d;
                if (kind == 2) { rows = std::max(rows, L.hCell + 6); mcw = std::max(mcw, cw); }  // the TMA box always has hCell + 6 rows
            }
    }
    *tileRows = (std::max(rows, 7) + 3) & ~3;  // multiple of 4: keeps every per-warp piece 128-byte aligned for both row strides
    *maxCw = mcw;
}

// box rows of 64 bytes hold crops up to 49 columns (15 alignment bytes + crop), 96 bytes up to 81 (the largest possible: 78)
int fast_tile_stride(int maxCw) { return maxCw <= 49 ? 64 : 96; }
#ifndef FAST_PAD
#define FAST_PAD 0   // residency experiment: unused bytes of shared memory per warp
#endif
static size_t fast_warp_bytes(int tileRows, int ts) { return (size_t)(NB + 1) * tileRows * ts + CAPW * 4 + 2 * CAP * 2 + FAST_PAD; }

cudaError_t fast_configure(int tileRows, int ts) {
    const int smem = (int)(FW * fast_warp_bytes(tileRows, ts));
    cudaError_t e = raise_dynamic_smem(k_fast_cells<64>, (size_t)smem);
    if (e == cudaSuccess) e = raise_dynamic_smem(k_fast_cells<96>, (size_t)smem);
    return e;
}

int fast_encode_maps(const Geom& g, uint8_t* pyr, int nImages, int ts, FastMaps* out) {
    for (int l = 0; l < g.nlevels; ++l) {
        const LevelGeom& L = g.lv[l];
        const size_t blockOff = (size_t)L.planeOff - (size_t)kMarginY * L.pitch - kMarginX;
        if (tma_encode_3d(&out->m[l], pyr + blockOff, L.pitch, L.h + 2 * kMarginY, nImages, L.planeBytes, ts, L.hCell + 6)) return -1;
    }
    return 0;
}

void launch_fast(const Geom& g, const Buffers& b, const FastMaps& maps, int nImages, int tileRows, int ts, int numSMs, cudaStream_t s) {
    const size_t smem = FW * fast_warp_bytes(tileRows, ts);
    const int ctasPerSM = std::max(1, std::min(32, (int)((227 * 1024) / (smem + 1024 + 64))));
    const int nWork = g.cellsPerImage * nImages;
    const int grid = std::min((nWork + FW - 1) / FW, numSMs * ctasPerSM);
    if (ts == 64)
        k_fast_cells<64><<<grid, 32 * FW, smem, s>>>(g, b, maps, g.cellsPerImage, nWork, tileRows * ts);
    else
        k_fast_cells<96><<<grid, 32 * FW, smem, s>>>(g, b, maps, g.cellsPerImage, nWork, tileRows * ts);
}

}  // namespace orbx
