// octree_kernel.cu — DistributeOctTree on the device: one CTA per (image, pyramid level).
//
// Replaces ORBextractor::DistributeOctTree / ExtractorNode::DivideNode / compareNodes (ORBextractor.cc:454-753).
// The reference's result depends on (1) the order of its std::list (children are push_front-ed in the order n1..n4, the
// parent is erased), (2) the phase-2 processing order, which is libstdc++'s unstable std::sort on (nKeys, UL.x), and
// (3) "first maximum wins" inside a node, i.e. the order of the candidate vector.  All three are reproduced:
//   * the node list is a doubly linked list in shared memory, driven by one thread exactly like the reference loop;
//   * the sort is the element-exact emulation in stdsort_emul.cuh;
//   * candidates may sit in memory in ANY order: a node's key set does not depend on it, and the winner is taken as
//     max (response, -reference position) with the position recomputed from the coordinates (cand_order_key).
// What runs in parallel is everything proportional to the number of candidates: root assignment, the child census of
// every splittable node (shared-memory atomics), the re-labelling after a round of splits, and the final arg-max.
#include <algorithm>

#include "orbx_internal.cuh"
#include "stdsort_emul.cuh"

namespace orbx {

namespace {

constexpr int OT_THREADS = 64;
constexpr uint16_t NIL = 0xFFFF;
typedef StdSortEmul<16> NodeSort;  // element = (nKeys << 12 | UL.x) << 16 | slot

struct Box { short x0, y0, x1, y1; };

struct OctShared {
    Box* box;                  // [cap]
    int* cnt;                  // [cap]   keys in the node
    int* ccnt;                 // [cap*4] census: keys per child quadrant
    unsigned long long* srt;   // [cap]   sort buffer / final arg-max (aliases nothing)
    uint16_t* cslot;           // [cap*4] slot of each created child
    uint16_t* nxt;             // [cap]
    uint16_t* prv;             // [cap]
    uint16_t* freeStack;       // [cap]
    uint16_t* recA;            // [cap]   splittable nodes in creation order (vSizeAndPointerToNode)
    uint16_t* recB;            // [cap]
    uint16_t* outPos;          // [cap]
    uint8_t* split;            // [cap]   node was split in the current round
};

__device__ __forceinline__ size_t align8(size_t v) { return (v + 7) & ~(size_t)7; }

__host__ __device__ inline size_t oct_smem_bytes(int cap) {
    size_t c = (size_t)cap;
    return c * (sizeof(Box) + 4 + 16 + 8 + 8 + 2 * 6) + ((c + 7) & ~(size_t)7) + 64;
}

__device__ inline OctShared carve(unsigned char* base, int cap) {
    OctShared s;
    size_t o = 0;
    s.srt = reinterpret_cast<unsigned long long*>(base + o); o += (size_t)cap * 8;
    s.box = reinterpret_cast<Box*>(base + o); o += (size_t)cap * sizeof(Box);
    s.ccnt = reinterpret_cast<int*>(base + o); o += (size_t)cap * 16;
    s.cnt = reinterpret_cast<int*>(base + o); o += (size_t)cap * 4;
    s.cslot = reinterpret_cast<uint16_t*>(base + o); o += (size_t)cap * 8;
    s.nxt = reinterpret_cast<uint16_t*>(base + o); o += (size_t)cap * 2;
    s.prv = reinterpret_cast<uint16_t*>(base + o); o += (size_t)cap * 2;
    s.freeStack = reinterpret_cast<uint16_t*>(base + o); o += (size_t)cap * 2;
    s.recA = reinterpret_cast<uint16_t*>(base + o); o += (size_t)cap * 2;
    s.recB = reinterpret_cast<uint16_t*>(base + o); o += (size_t)cap * 2;
    s.outPos = reinterpret_cast<uint16_t*>(base + o); o += (size_t)cap * 2;
    s.split = reinterpret_cast<uint8_t*>(base + o);
    return s;
}

// list state, owned by thread 0
struct ListCtl {
    int head, tail, size, nFree;
};

__device__ __forceinline__ void push_front(const OctShared& s, ListCtl& c, int slot) {
    s.prv[slot] = NIL;
    s.nxt[slot] = (uint16_t)(c.head < 0 ? NIL : c.head);
    if (c.head >= 0) s.prv[c.head] = (uint16_t)slot; else c.tail = slot;
    c.head = slot;
    ++c.size;
}
__device__ __forceinline__ void erase(const OctShared& s, ListCtl& c, int slot) {
    const int p = s.prv[slot], n = s.nxt[slot];
    if (p != NIL) s.nxt[p] = (uint16_t)n; else c.head = (n == NIL ? -1 : n);
    if (n != NIL) s.prv[n] = (uint16_t)p; else c.tail = (p == NIL ? -1 : p);
    --c.size;
    s.freeStack[c.nFree++] = (uint16_t)slot;
}

// ExtractorNode::DivideNode + the four "push_front if non-empty" blocks (ORBextractor.cc:454-510, 611-650 / 680-716).
// The census ccnt[] was taken in parallel before.  Returns false when the slot pool is exhausted.
__device__ __forceinline__ bool split_node(const OctShared& s, ListCtl& c, int node, uint16_t* rec, int& nRec) {
    const Box b = s.box[node];
    const int halfX = (b.x1 - b.x0 + 1) >> 1;  // ceil(float(UR.x-UL.x)/2) on non-negative ints
    const int halfY = (b.y1 - b.y0 + 1) >> 1;
    const int mx = b.x0 + halfX, my = b.y0 + halfY;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int k = s.ccnt[node * 4 + q];
        if (k == 0) continue;
        if (c.nFree == 0) return false;
        const int slot = s.freeStack[--c.nFree];
        Box cb;
        cb.x0 = (short)((q & 1) ? mx : b.x0);
        cb.x1 = (short)((q & 1) ? b.x1 : mx);
        cb.y0 = (short)((q & 2) ? my : b.y0);
        cb.y1 = (short)((q & 2) ? b.y1 : my);
        s.box[slot] = cb;
        s.cnt[slot] = k;
        s.cslot[node * 4 + q] = (uint16_t)slot;
        push_front(s, c, slot);
        if (k > 1) rec[nRec++] = (uint16_t)slot;
    }
    s.split[node] = 1;
    erase(s, c, node);
    return true;
}

}  // namespace

__global__ void __launch_bounds__(OT_THREADS) k_octree(const __grid_constant__ Geom g, const Buffers b, int levelBase) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    __shared__ int sh_finish, sh_size, sh_err;

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
    const uint32_t* __restrict__ cand = b.cand + L.candOff + (size_t)img * L.candCap;
    uint16_t* __restrict__ pnode = b.pnode + L.candOff + (size_t)img * L.candCap;

    // ---- roots (ORBextractor.cc:533-575) ----
    const int nIni = L.nIni;
    for (int k = tid; k < cap; k += OT_THREADS) { s.cnt[k] = 0; s.split[k] = 0; }
    __syncthreads();
    for (int c = tid; c < n; c += OT_THREADS) {
        int r = (int)__fdiv_rn((float)cand_x(cand[c]), L.hX);  // vpIniNodes[kp.pt.x/hX]
        r = min(r, nIni - 1);
        pnode[c] = (uint16_t)r;
        atomicAdd(&s.cnt[r], 1);
    }
    __syncthreads();

    ListCtl ctl;
    int nRec = 0;        // thread 0: entries of the current splittable-node record
    uint16_t* rec = s.recA;
    uint16_t* recNext = s.recB;
    int phase = 1;
    if (tid == 0) {
        ctl.head = ctl.tail = -1; ctl.size = 0; ctl.nFree = 0;
        for (int k = cap - 1; k >= nIni; --k) s.freeStack[ctl.nFree++] = (uint16_t)k;
        for (int i = nIni - 1; i >= 0; --i) {  // push_back in order 0..nIni-1 == push_front in reverse; empty roots are dropped
            if (s.cnt[i] == 0) { s.freeStack[ctl.nFree++] = (uint16_t)i; continue; }
            Box rb;
            rb.x0 = (short)(int)__fmul_rn(L.hX, (float)i);
            rb.x1 = (short)(int)__fmul_rn(L.hX, (float)(i + 1));
            rb.y0 = 0;
            rb.y1 = (short)(L.h - 2 * kBorder);
            s.box[i] = rb;
            push_front(s, ctl, i);
        }
        sh_finish = 0;
        sh_err = 0;
    }
    __syncthreads();

    // ---- rounds: one phase-1 sweep or one phase-2 pass per iteration (ORBextractor.cc:584-729) ----
    for (int round = 0;; ++round) {
        for (int k = tid; k < cap * 4; k += OT_THREADS) s.ccnt[k] = 0;
        __syncthreads();
        // census of every splittable node (a node is splittable iff it holds more than one key)
        for (int c = tid; c < n; c += OT_THREADS) {
            const int node = pnode[c] & 0x3FFF;
            if (s.cnt[node] > 1) {
                const uint32_t pc = cand[c];
                const Box bx = s.box[node];
                const int mx = bx.x0 + ((bx.x1 - bx.x0 + 1) >> 1), my = bx.y0 + ((bx.y1 - bx.y0 + 1) >> 1);
                const int q = (cand_x(pc) < mx ? 0 : 1) | (cand_y(pc) < my ? 0 : 2);
                pnode[c] = (uint16_t)(node | (q << 14));
                atomicAdd(&s.ccnt[node * 4 + q], 1);
            }
        }
        __syncthreads();
        if (tid == 0) {
            const int prevSize = ctl.size;
            bool ok = true;
            int nNew = 0;
            if (phase == 1) {
                int nToExpand = 0;
                int it = ctl.head;
                while (it >= 0 && ok) {
                    const int nx = s.nxt[it] == NIL ? -1 : s.nxt[it];
                    if (s.cnt[it] > 1) {
                        const int before = nNew;
                        ok = split_node(s, ctl, it, recNext, nNew);
                        nToExpand += nNew - before;
                    }
                    it = nx;
                }
                if (ctl.size >= N || ctl.size == prevSize) sh_finish = 1;
                else if (ctl.size + nToExpand * 3 > N) phase = 2;
            } else {
                for (int k = 0; k < nRec; ++k) {
                    const int slot = rec[k];
                    s.srt[k] = ((unsigned long long)(((unsigned)s.cnt[slot] << 12) | (unsigned)s.box[slot].x0) << 16) | (unsigned)slot;
                }
                NodeSort::sort(s.srt, nRec);
                for (int jj = nRec - 1; jj >= 0 && ok; --jj) {
                    ok = split_node(s, ctl, (int)(s.srt[jj] & 0xFFFF), recNext, nNew);
                    if (ctl.size >= N) break;
                }
                if (ctl.size >= N || ctl.size == prevSize) sh_finish = 1;
            }
            if (!ok || round > 4096) { sh_err = 1; sh_finish = 1; }  // slot pool exhausted / runaway guard (never expected)
            uint16_t* tmp = rec; rec = recNext; recNext = tmp;
            nRec = nNew;
        }
        __syncthreads();
        // re-label the keys of the nodes that were split
        for (int c = tid; c < n; c += OT_THREADS) {
            const int e = pnode[c], node = e & 0x3FFF;
            if (s.split[node]) pnode[c] = s.cslot[node * 4 + (e >> 14)];
        }
        const int fin = sh_finish;
        __syncthreads();
        if (fin) break;
        for (int k = tid; k < cap; k += OT_THREADS) s.split[k] = 0;
        // (the ccnt reset at the top of the loop is followed by a barrier, which also orders this reset)
    }

    // ---- retain the best key of every node, in list order (ORBextractor.cc:732-750) ----
    if (tid == 0) {
        int pos = 0;
        for (int it = ctl.head; it >= 0; it = (s.nxt[it] == NIL ? -1 : s.nxt[it])) s.outPos[it] = (uint16_t)pos++;
        sh_size = pos;
    }
    for (int k = tid; k < cap; k += OT_THREADS) s.srt[k] = 0;
    __syncthreads();
    const unsigned long long mask40 = (1ull << 40) - 1;
    for (int c = tid; c < n; c += OT_THREADS) {
        const uint32_t pc = cand[c];
        const unsigned long long v = ((unsigned long long)cand_r(pc) << 40) | (~cand_order_key(pc, L.wCell, L.hCell) & mask40);
        atomicMax(&s.srt[s.outPos[pnode[c] & 0x3FFF]], v);
    }
    __syncthreads();
    const int size = sh_size;
    if (size > L.selCap || sh_err) {
        if (tid == 0) { atomicExch(b.errFlag, 2); *selCount = 0; }
        return;
    }
    uint32_t* sel = b.sel + (size_t)img * g.K + L.selOff;
    for (int k = tid; k < size; k += OT_THREADS) {
        const unsigned long long v = s.srt[k];
        const unsigned long long key = ~v & mask40;
        sel[k] = pack_cand((int)(key & 0xFFF), (int)((key >> 12) & 0xFFF), (int)(v >> 40));
    }
    if (tid == 0) *selCount = size;
}

size_t octree_smem_bytes(int nodeCap) { return oct_smem_bytes(nodeCap); }

// opt in to > 48 KB of dynamic shared memory (quotas above ~700 nodes per level); per device, called at handle creation
cudaError_t octree_configure(int maxNodeCap) {
    const size_t smem = oct_smem_bytes(maxNodeCap);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    return cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024));
}

// Levels are launched in two groups (large / small quotas) so that the small levels do not reserve the shared memory of the
// largest one: the serial controller of a CTA is latency bound, what matters is how many CTAs are resident per SM.
cudaError_t launch_octree(const Geom& g, const Buffers& b, int nImages, cudaStream_t s) {
    int maxCap = 0;
    for (int l = 0; l < g.nlevels; ++l) maxCap = std::max(maxCap, g.lv[l].nodeCap);
    // node capacities shrink with the level (quota decay): split where the capacity first drops below half of the largest
    int split = g.nlevels;
    for (int l = 0; l < g.nlevels; ++l)
        if (2 * g.lv[l].nodeCap <= maxCap) { split = l; break; }
    for (int l = split; l < g.nlevels; ++l)
        if (2 * g.lv[l].nodeCap > maxCap) { split = g.nlevels; break; }  // not monotone: one group
    if (split > 0) {
        dim3 grid(split, nImages);
        k_octree<<<grid, OT_THREADS, oct_smem_bytes(maxCap), s>>>(g, b, 0);
    }
    if (split < g.nlevels) {
        int cap2 = 0;
        for (int l = split; l < g.nlevels; ++l) cap2 = std::max(cap2, g.lv[l].nodeCap);
        dim3 grid(g.nlevels - split, nImages);
        k_octree<<<grid, OT_THREADS, oct_smem_bytes(cap2), s>>>(g, b, split);
    }
    return cudaGetLastError();
}

// test hook: the sort emulation in one device thread, the way k_octree runs it
__global__ void k_debug_sort(unsigned long long* buf, int n, int forceDepth) {
    if (threadIdx.x == 0 && blockIdx.x == 0) NodeSort::sort(buf, n, forceDepth);
}
void launch_debug_sort_impl(unsigned long long* buf, int n, int forceDepth, cudaStream_t s) { k_debug_sort<<<1, 32, 0, s>>>(buf, n, forceDepth); }

}  // namespace orbx
