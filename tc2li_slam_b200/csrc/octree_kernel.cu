// octree_kernel.cu — DistributeOctTree on the device: one CTA per (image, pyramid level).
//
// Replaces ORBextractor::DistributeOctTree / ExtractorNode::DivideNode / compareNodes (ORBextractor.cc:454-753).
// The reference's result depends on (1) the order of its std::list (children are push_front-ed in the order n1..n4, the
// parent is erased), (2) the phase-2 processing order, which is libstdc++'s unstable std::sort on (nKeys, UL.x), and
// (3) "first maximum wins" inside a node, i.e. the order of the candidate vector.  All three are reproduced:
//   * the node table is kept in list order; every round computes each node's next position with prefix sums — the closed
//     form of the reference's push_front / erase sequence (see k_octree);
//   * the sort is the element-exact emulation in stdsort_emul.cuh;
//   * candidates may sit in memory in ANY order: a node's key set does not depend on it, and the winner is taken as
//     max (response, -reference position) with the position recomputed from the coordinates (cand_order_key).
// What runs in parallel is everything proportional to the number of candidates: root assignment, the child census of
// every splittable node (shared-memory atomics), the re-labelling after a round of splits, and the final arg-max.
#include <algorithm>

#include "orbx_internal.cuh"
#include "parsort_emul.cuh"
#include "stdsort_emul.cuh"

namespace orbx {

namespace {

constexpr int OT_MAX_WARPS = PS_MAX_WARPS;
#ifndef OT_BATCH_THREADS
#define OT_BATCH_THREADS 128
#endif
constexpr int OT_T_BATCH = OT_BATCH_THREADS;       // threads per CTA when many images are in flight (shared memory bounds the CTAs per SM)
constexpr int OT_T_FEW = 256;         // threads per CTA for a few large frames (long candidate lists, few CTAs)
constexpr int OT_STAGE_MAX = 2048;    // node labels of candidate lists up to the launch's stage capacity live in shared memory (2 bytes each)
typedef StdSortEmul<16> NodeSort;    // element = (nKeys << 12 | UL.x) << 16 | node index: any key size
typedef StdSortEmul<32> NodeSort32;  // element = (nKeys << 12 | UL.x) << 32 | node index: keys below 2^32 (fewer than 2^20 candidates)

struct Box { short x0, y0, x1, y1; };

// Shared-memory state of one (image, level).  The node table is kept IN LIST ORDER (index == position in the reference's
// std::list) and double buffered: every round writes the next table from the current one.
struct OctShared {
    Box* box[2];               // [cap]
    int* cnt[2];               // [cap]   keys in the node
    int* ccnt;                 // [cap*4] census: keys per child quadrant of every splittable node
    unsigned long long* srt;   // [cap]   sort buffer / final arg-max
    uint16_t* childIdx;        // [cap*4] index of each created child in the next table
    uint16_t* tailIdx;         // [cap]   index of an unsplit node in the next table
    uint16_t* v0;              // [cap]   per-node scan inputs / exclusive prefixes (children, kept, splittable children)
    uint16_t* v1;
    uint16_t* v2;
    uint16_t* blkStart;        // [cap]   phase 2: start of a split node's child block, start of its entries in the next record
    uint16_t* recStart;
    uint16_t* rec[2];          // [cap]   splittable nodes in creation order (vSizeAndPointerToNode)
    uint8_t* split;            // [cap]   node is split in this phase-2 round
};

__host__ __device__ inline size_t oct_node_bytes(int cap) {
    size_t c = (size_t)((cap + 7) & ~7);
    return c * (2 * sizeof(Box) + 2 * 4 + 16 + 8 + 8 + 2 * 8 + 1) + 128;
}
__host__ __device__ inline size_t oct_smem_bytes(int cap, int stageCap) { return oct_node_bytes(cap) + (size_t)stageCap * 2; }
// labels staged for lists up to four times the level's quota (at least 512, at most OT_STAGE_MAX entries); longer lists
// keep their labels in global memory
static int oct_stage_cap(int quota) { return std::min(OT_STAGE_MAX, std::max(512, (4 * quota + 63) & ~63)); }

__device__ inline OctShared carve(unsigned char* base, int capIn) {
    const size_t cap = (size_t)((capIn + 7) & ~7);
    OctShared s;
    size_t o = 0;
    s.srt = reinterpret_cast<unsigned long long*>(base + o); o += cap * 8;
    s.box[0] = reinterpret_cast<Box*>(base + o); o += cap * sizeof(Box);
    s.box[1] = reinterpret_cast<Box*>(base + o); o += cap * sizeof(Box);
    s.ccnt = reinterpret_cast<int*>(base + o); o += cap * 16;
    s.cnt[0] = reinterpret_cast<int*>(base + o); o += cap * 4;
    s.cnt[1] = reinterpret_cast<int*>(base + o); o += cap * 4;
    s.childIdx = reinterpret_cast<uint16_t*>(base + o); o += cap * 8;
    s.tailIdx = reinterpret_cast<uint16_t*>(base + o); o += cap * 2;
    s.v0 = reinterpret_cast<uint16_t*>(base + o); o += cap * 2;
    s.v1 = reinterpret_cast<uint16_t*>(base + o); o += cap * 2;
    s.v2 = reinterpret_cast<uint16_t*>(base + o); o += cap * 2;
    s.blkStart = reinterpret_cast<uint16_t*>(base + o); o += cap * 2;
    s.recStart = reinterpret_cast<uint16_t*>(base + o); o += cap * 2;
    s.rec[0] = reinterpret_cast<uint16_t*>(base + o); o += cap * 2;
    s.rec[1] = reinterpret_cast<uint16_t*>(base + o); o += cap * 2;
    s.split = reinterpret_cast<uint8_t*>(base + o);
    return s;
}

// In-place exclusive scan of three uint16 arrays of length n by the whole CTA (chunk per thread, warp-shuffle scan of the
// chunk sums, then the warp totals).  Returns the three totals to every thread.  `part` is a [3][OT_MAX_WARPS] int scratch in
// shared memory.
template <int OT_THREADS>
__device__ inline void scan3(uint16_t* a0, uint16_t* a1, uint16_t* a2, int n, int (*part)[OT_MAX_WARPS], int& t0, int& t1, int& t2) {
    constexpr int NW = OT_THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = (n + OT_THREADS - 1) / OT_THREADS;
    const int lo = min(tid * per, n), hi = min(lo + per, n);
    int s0 = 0, s1 = 0, s2 = 0;
    for (int i = lo; i < hi; ++i) { s0 += a0[i]; s1 += a1[i]; s2 += a2[i]; }
    int i0 = s0, i1 = s1, i2 = s2;  // inclusive warp scans of the chunk sums
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u0 = __shfl_up_sync(0xFFFFFFFFu, i0, o), u1 = __shfl_up_sync(0xFFFFFFFFu, i1, o), u2 = __shfl_up_sync(0xFFFFFFFFu, i2, o);
        if (lane >= o) { i0 += u0; i1 += u1; i2 += u2; }
    }
    if (lane == 31) { part[0][wid] = i0; part[1][wid] = i1; part[2][wid] = i2; }
    __syncthreads();
    int b0 = 0, b1 = 0, b2 = 0;
    t0 = 0; t1 = 0; t2 = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const int p0 = part[0][w], p1 = part[1][w], p2 = part[2][w];
        if (w < wid) { b0 += p0; b1 += p1; b2 += p2; }
        t0 += p0; t1 += p1; t2 += p2;
    }
    s0 = b0 + i0 - s0; s1 = b1 + i1 - s1; s2 = b2 + i2 - s2;  // exclusive prefix of this thread's chunk
    for (int i = lo; i < hi; ++i) {
        const int x0 = a0[i], x1 = a1[i], x2 = a2[i];
        a0[i] = (uint16_t)s0; a1[i] = (uint16_t)s1; a2[i] = (uint16_t)s2;
        s0 += x0; s1 += x1; s2 += x2;
    }
    __syncthreads();
}

// children of node `i` (ExtractorNode::DivideNode, ORBextractor.cc:454-510) written into the next table as the block that
// the reference's push_front sequence n1, n2, n3, n4 leaves at the list front: order n4, n3, n2, n1 (non-empty ones only);
// the splittable children go to the next record in creation order n1..n4 (ORBextractor.cc:611-650 / 680-716)
__device__ __forceinline__ void emit_children(const OctShared& s, int cur, int i, int blockStart, int recPos, uint16_t* recNext) {
    const Box bx = s.box[cur][i];
    const int mx = bx.x0 + ((bx.x1 - bx.x0 + 1) >> 1), my = bx.y0 + ((bx.y1 - bx.y0 + 1) >> 1);  // ceil(float(size)/2) on non-negative ints
    int k[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) k[q] = s.ccnt[i * 4 + q];
    int pos = blockStart;
#pragma unroll
    for (int q = 3; q >= 0; --q) {
        if (k[q] == 0) continue;
        Box cb;
        cb.x0 = (short)((q & 1) ? mx : bx.x0);
        cb.x1 = (short)((q & 1) ? bx.x1 : mx);
        cb.y0 = (short)((q & 2) ? my : bx.y0);
        cb.y1 = (short)((q & 2) ? bx.y1 : my);
        s.box[cur ^ 1][pos] = cb;
        s.cnt[cur ^ 1][pos] = k[q];
        s.childIdx[i * 4 + q] = (uint16_t)pos;
        ++pos;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (k[q] > 1) recNext[recPos++] = s.childIdx[i * 4 + q];
}

}  // namespace

// Rounds are data parallel: a census of every splittable node's four quadrants (shared-memory atomics over the
// candidates), then one pass over the nodes that computes, with CTA-wide prefix sums, where each child / each surviving
// node lands in the next list — the closed form of the reference's "push_front the children, erase the parent" loop:
//   phase 1 (every splittable node is split, list order):  children of node i start at  sum_{j>i} nch_j,
//   phase 2 (nodes split in descending sorted order until the list has N nodes): children of the r-th processed node start
//            at  sum_{r<r'<t} nch_r',  t = number of nodes processed before the reference's `break`;
//   nodes that are not split follow all child blocks in their old relative order.
// The libstdc++-exact sort of phase 2 (ties decide results) is CTA-parallel too (parsort_emul.cuh).
template <int OT_THREADS>
__global__ void __launch_bounds__(OT_THREADS) k_octree(const __grid_constant__ Geom g, const Buffers b, int levelBase, int stageCap) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    __shared__ int part[3][OT_MAX_WARPS];
    __shared__ int sh_t;
    __shared__ ParSortShared psh;

    const int level = levelBase + blockIdx.x, img = b.imgBase + blockIdx.y, tid = threadIdx.x;
    const LevelGeom& L = g.lv[level];
    const int cap = L.nodeCap, N = L.quota;
    const int n = min(b.candCount[img * ORBX_MAX_LEVELS + level], L.candCap);
    int* selCount = b.selCount + img * ORBX_MAX_LEVELS + level;
    if (n == 0) {
        if (tid == 0) *selCount = 0;
        return;
    }
    const OctShared s = carve(smemRaw, cap);
    // The CTA is latency bound (two warps, a barrier every few instructions): what matters is how many CTAs an SM holds, i.e.
    // shared memory per CTA.  The candidates are read-only and come through L1 (ld.global.nc: every round walks them twice and
    // they stay resident); their node labels, rewritten every round, are staged in shared memory when the list fits.
    const uint32_t* __restrict__ cand = b.cand + L.candOff + (size_t)img * L.candCap;
    uint16_t* pnode = b.pnode + L.candOff + (size_t)img * L.candCap;
    if (n <= stageCap) pnode = reinterpret_cast<uint16_t*>(smemRaw + oct_node_bytes(cap));

    // ---- roots (ORBextractor.cc:533-575): nIni vertical strips; empty ones are dropped ----
    const int nIni = L.nIni;
    int* rootCnt = s.ccnt;  // scratch
    for (int k = tid; k < nIni; k += OT_THREADS) rootCnt[k] = 0;
    __syncthreads();
    for (int c = tid; c < n; c += OT_THREADS) {
        int r = (int)__fdiv_rn((float)cand_x(__ldg(cand + c)), L.hX);  // vpIniNodes[kp.pt.x/hX]
        r = min(r, nIni - 1);
        pnode[c] = (uint16_t)r;
        atomicAdd(&rootCnt[r], 1);
    }
    __syncthreads();
    for (int k = tid; k < nIni; k += OT_THREADS) { s.v0[k] = rootCnt[k] > 0; s.v1[k] = 0; s.v2[k] = 0; }
    __syncthreads();
    int cur = 0, size, d1, d2;
    scan3<OT_THREADS>(s.v0, s.v1, s.v2, nIni, part, size, d1, d2);
    for (int k = tid; k < nIni; k += OT_THREADS) {
        s.tailIdx[k] = s.v0[k];
        if (rootCnt[k] > 0) {
            Box rb;
            rb.x0 = (short)(int)__fmul_rn(L.hX, (float)k);
            rb.x1 = (short)(int)__fmul_rn(L.hX, (float)(k + 1));
            rb.y0 = 0;
            rb.y1 = (short)(L.h - 2 * kBorder);
            s.box[0][s.v0[k]] = rb;
            s.cnt[0][s.v0[k]] = rootCnt[k];
        }
    }
    __syncthreads();
    for (int c = tid; c < n; c += OT_THREADS) pnode[c] = s.tailIdx[pnode[c]];
    __syncthreads();

    // ---- rounds (ORBextractor.cc:584-729) ----
    int phase = 1, nRec = 0, recSel = 0;
    bool err = false;
    for (int round = 0; round < 4096; ++round) {
        const int prevSize = size;
        for (int k = tid; k < size * 4; k += OT_THREADS) s.ccnt[k] = 0;
        __syncthreads();
        // census of every splittable node (a node is splittable iff it holds more than one key)
        for (int c = tid; c < n; c += OT_THREADS) {
            const int node = pnode[c] & 0x3FFF;
            if (s.cnt[cur][node] > 1) {
                const uint32_t pc = __ldg(cand + c);
                const Box bx = s.box[cur][node];
                const int mx = bx.x0 + ((bx.x1 - bx.x0 + 1) >> 1), my = bx.y0 + ((bx.y1 - bx.y0 + 1) >> 1);
                const int q = (cand_x(pc) < mx ? 0 : 1) | (cand_y(pc) < my ? 0 : 2);
                pnode[c] = (uint16_t)(node | (q << 14));
                atomicAdd(&s.ccnt[node * 4 + q], 1);
            }
        }
        __syncthreads();
        uint16_t* recCur = s.rec[recSel];
        uint16_t* recNext = s.rec[recSel ^ 1];
        int newSize, nNew;
        if (phase == 1) {
            // every splittable node is split, in list order
            for (int i = tid; i < size; i += OT_THREADS) {
                int nch = 0, nexp = 0;
                const bool sp = s.cnt[cur][i] > 1;
                if (sp) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) { const int k = s.ccnt[i * 4 + q]; nch += k > 0; nexp += k > 1; }
                }
                s.v0[i] = (uint16_t)nch; s.v1[i] = (uint16_t)(sp ? 0 : 1); s.v2[i] = (uint16_t)nexp;
                s.split[i] = sp;
                s.blkStart[i] = (uint16_t)nch;  // keep the count: the scan overwrites v0
            }
            __syncthreads();
            int C, T, R;
            scan3<OT_THREADS>(s.v0, s.v1, s.v2, size, part, C, T, R);
            newSize = C + T; nNew = R;
            if (newSize > cap) { err = true; break; }
            for (int i = tid; i < size; i += OT_THREADS) {
                if (s.split[i])
                    emit_children(s, cur, i, C - s.v0[i] - s.blkStart[i], s.v2[i], recNext);
                else {
                    const int pos = C + s.v1[i];
                    s.box[cur ^ 1][pos] = s.box[cur][i];
                    s.cnt[cur ^ 1][pos] = s.cnt[cur][i];
                    s.tailIdx[i] = (uint16_t)pos;
                }
            }
        } else {
            // phase 2: nodes of the record, largest first (libstdc++ std::sort order, ties included), until the list has N nodes
            const bool small = n < (1 << 20);  // every node count then fits 20 bits: 32-bit keys, single-word compares
            for (int k = tid; k < nRec; k += OT_THREADS) {
                const int i = recCur[k];
                const unsigned long long key = ((unsigned long long)(unsigned)s.cnt[cur][i] << 12) | (unsigned)s.box[cur][i].x0;
                s.srt[k] = small ? (key << 32) | (unsigned)i : (key << 16) | (unsigned)i;
            }
            if (tid == 0) sh_t = nRec;
            __syncthreads();
            // scratch: the next node table's boxes (not written yet) and two of the scan arrays
            unsigned long long* tmp = reinterpret_cast<unsigned long long*>(s.box[cur ^ 1]);
            if (small) par_std_sort<32, OT_THREADS>(s.srt, nRec, tmp, s.v0, s.v1, psh);
            else par_std_sort<16, OT_THREADS>(s.srt, nRec, tmp, s.v0, s.v1, psh);
            for (int i = tid; i < size; i += OT_THREADS) s.split[i] = 0;
            __syncthreads();
            // rank r = 0 .. nRec-1 <-> node srt[nRec-1-r]; growth of the list when it is split = non-empty children - 1
            for (int r = tid; r < nRec; r += OT_THREADS) {
                const int i = (int)(s.srt[nRec - 1 - r] & 0xFFFF);
                int nch = 0, nexp = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) { const int k = s.ccnt[i * 4 + q]; nch += k > 0; nexp += k > 1; }
                s.v0[r] = (uint16_t)nch; s.v1[r] = (uint16_t)(nch - 1); s.v2[r] = (uint16_t)nexp;
                s.tailIdx[r] = (uint16_t)nch;  // scratch: keep the counts, the scan overwrites v0
            }
            __syncthreads();
            int C, G, R;
            scan3<OT_THREADS>(s.v0, s.v1, s.v2, nRec, part, C, G, R);
            // the reference breaks after the first split that brings the list to >= N nodes
            for (int r = tid; r < nRec; r += OT_THREADS) {
                const int after = size + s.v1[r] + (int)s.tailIdx[r] - 1;  // list size after the r-th split
                if (after >= N) atomicMin(&sh_t, r + 1);
            }
            __syncthreads();
            const int t = sh_t;
            const int Ct = t < nRec ? s.v0[t] : C;   // children of the t processed nodes
            nNew = t < nRec ? s.v2[t] : R;
            for (int r = tid; r < t; r += OT_THREADS) {
                const int i = (int)(s.srt[nRec - 1 - r] & 0xFFFF);
                s.split[i] = 1;
                s.blkStart[i] = (uint16_t)(Ct - s.v0[r] - (int)s.tailIdx[r]);  // blocks of later ranks sit in front
                s.recStart[i] = s.v2[r];
            }
            __syncthreads();
            for (int i = tid; i < size; i += OT_THREADS) { s.v0[i] = (uint16_t)(s.split[i] ? 0 : 1); s.v1[i] = 0; s.v2[i] = 0; }
            __syncthreads();
            int T, u1, u2;
            scan3<OT_THREADS>(s.v0, s.v1, s.v2, size, part, T, u1, u2);
            newSize = Ct + T;
            if (newSize > cap) { err = true; break; }
            for (int i = tid; i < size; i += OT_THREADS) {
                if (s.split[i])
                    emit_children(s, cur, i, s.blkStart[i], s.recStart[i], recNext);
                else {
                    const int pos = Ct + s.v0[i];
                    s.box[cur ^ 1][pos] = s.box[cur][i];
                    s.cnt[cur ^ 1][pos] = s.cnt[cur][i];
                    s.tailIdx[i] = (uint16_t)pos;
                }
            }
        }
        __syncthreads();
        // re-label the keys
        for (int c = tid; c < n; c += OT_THREADS) {
            const int e = pnode[c], node = e & 0x3FFF;
            pnode[c] = s.split[node] ? s.childIdx[node * 4 + (e >> 14)] : s.tailIdx[node];
        }
        __syncthreads();
        cur ^= 1;
        recSel ^= 1;
        size = newSize;
        nRec = nNew;
        if (size >= N || size == prevSize) break;
        if (phase == 1 && size + 3 * nNew > N) phase = 2;
    }

    // ---- retain the best key of every node, in list order (ORBextractor.cc:732-750) ----
    if (err || size > L.selCap) {
        if (tid == 0) { atomicExch(b.errFlag, 2); *selCount = 0; }
        return;
    }
    for (int k = tid; k < size; k += OT_THREADS) s.srt[k] = 0;
    __syncthreads();
    const unsigned long long mask40 = (1ull << 40) - 1;
    for (int c = tid; c < n; c += OT_THREADS) {
        const uint32_t pc = __ldg(cand + c);
        const unsigned long long v = ((unsigned long long)cand_r(pc) << 40) | (~cand_order_key(pc, L.wCell, L.hCell) & mask40);
        atomicMax(&s.srt[pnode[c] & 0x3FFF], v);
    }
    __syncthreads();
    uint32_t* sel = b.sel + (size_t)img * g.K + L.selOff;
    for (int k = tid; k < size; k += OT_THREADS) {
        const unsigned long long v = s.srt[k];
        const unsigned long long key = ~v & mask40;
        sel[k] = pack_cand((int)(key & 0xFFF), (int)((key >> 12) & 0xFFF), (int)(v >> 40));
    }
    if (tid == 0) *selCount = size;
}

size_t octree_smem_bytes(int nodeCap) { return oct_smem_bytes(nodeCap, OT_STAGE_MAX); }

// opt in to > 48 KB of dynamic shared memory (quotas above ~700 nodes per level); per device, called at handle creation
cudaError_t octree_configure(int maxNodeCap) {
    const size_t smem = oct_smem_bytes(maxNodeCap, OT_STAGE_MAX);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = raise_dynamic_smem(k_octree<OT_T_BATCH>, smem);
    if (e != cudaSuccess) return e;
    return raise_dynamic_smem(k_octree<OT_T_FEW>, smem);
}

// Many images: one launch per level, each with exactly the shared memory its node capacity needs — the small levels then fit
// three to four times as many CTAs per SM as the largest one.  Few images (large single frames): one launch over all levels,
// so that the levels of an image run side by side instead of one after the other.
cudaError_t launch_octree(const Geom& g, const Buffers& b, int nImages, int numSMs, cudaStream_t s) {
    if (nImages * g.nlevels <= 8 * numSMs) {
        int maxCap = 0;
        for (int l = 0; l < g.nlevels; ++l) maxCap = std::max(maxCap, g.lv[l].nodeCap);
        dim3 grid(g.nlevels, nImages);
        k_octree<OT_T_FEW><<<grid, OT_T_FEW, oct_smem_bytes(maxCap, OT_STAGE_MAX), s>>>(g, b, 0, OT_STAGE_MAX);
        return cudaGetLastError();
    }
    for (int l = 0; l < g.nlevels; ++l) {
        const int stageCap = oct_stage_cap(g.lv[l].quota);
        dim3 grid(1, nImages);
        k_octree<OT_T_BATCH><<<grid, OT_T_BATCH, oct_smem_bytes(g.lv[l].nodeCap, stageCap), s>>>(g, b, l, stageCap);
    }
    return cudaGetLastError();
}
int octree_launch_count(const Geom& g, int nImages, int numSMs) { return nImages * g.nlevels <= 8 * numSMs ? 1 : g.nlevels; }

// test hooks: the sort emulation in one device thread, and the CTA-parallel version k_octree runs
__global__ void k_debug_sort(unsigned long long* buf, int n, int forceDepth) {
    if (threadIdx.x == 0 && blockIdx.x == 0) NodeSort::sort(buf, n, forceDepth);
}
__global__ void __launch_bounds__(PS_THREADS) k_debug_parsort(unsigned long long* buf, unsigned long long* tmp, uint16_t* p0, uint16_t* p1, int n) {
    __shared__ ParSortShared psh;
    par_std_sort<16>(buf, n, tmp, p0, p1, psh);
}
void launch_debug_sort_impl(unsigned long long* buf, int n, int forceDepth, cudaStream_t s) {
    if (forceDepth == -2) {  // the CTA-parallel emulation (global scratch)
        unsigned long long* tmp = nullptr;
        uint16_t* p = nullptr;
        cudaMalloc((void**)&tmp, (size_t)(n + 1) * 8);
        cudaMalloc((void**)&p, (size_t)(n + 1) * 4);
        k_debug_parsort<<<1, PS_THREADS, 0, s>>>(buf, tmp, p, p + n + 1, n);
        cudaStreamSynchronize(s);
        cudaFree(tmp);
        cudaFree(p);
        return;
    }
    k_debug_sort<<<1, 32, 0, s>>>(buf, n, forceDepth);
}

}  // namespace orbx
