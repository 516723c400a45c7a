// parsort_emul.cuh — the same libstdc++ std::sort emulation as stdsort_emul.cuh, executed by a whole CTA.
//
// What makes std::sort's result unique (ties included) is the sequence of Hoare partitions and the final insertion sort;
// both have closed forms that a CTA can evaluate in parallel:
//  * __unguarded_partition(first+1, last, pivot): with L_k the k-th position from the left holding an element >= pivot and
//    R_k the k-th from the right holding an element <= pivot (both on the array as it is when the call starts — swapped
//    positions always lie outside the window still to be scanned), the sequential loop swaps (L_k, R_k) for k < K,
//    K = min{k : L_k >= R_k}, and returns min(L_K, R_{K-1}) (R_{K-1} holds a swapped-in element >= pivot by then).
//    L and R come from two prefix counts.
//  * after __introsort_loop every element sits in a block of <= 16 elements that is <= all later blocks, and insertion sort
//    is stable and moves an element only past strictly greater ones: the final position of x is
//    x - #{p in [x-16, x) : key_p > key_x} + #{p in (x, x+16] : key_p < key_x}.
// The control flow (range stack, depth limit, median-of-three, heap-sort fallback) is the serial one, run by thread 0.
// Ranges of at most PS_SMALL elements are not partitioned by the CTA: they are parked and, at the end, each is finished by one
// thread running the plain serial loop (StdSortEmul::introsort_range) — all parked ranges at the same time; ranges are
// independent, so the result is the same, and the number of CTA-wide partition steps (four barriers each) drops several-fold.
// Call with all threads of the CTA (blockDim.x == PS_THREADS); contains __syncthreads().
#pragma once
#include "stdsort_emul.cuh"

namespace orbx {

constexpr int PS_THREADS = 64;      // default CTA size of the callers
constexpr int PS_MAX_WARPS = 8;

constexpr int PS_SMALL = 48;        // ranges up to this length are finished by ONE thread each, all of them concurrently (measured: 32 -> 4.07, 48 -> 3.95, 96 -> 4.11 ms)
constexpr int PS_SMALL_MAX = 128;   // capacity of the small-range list; beyond it ranges are partitioned by the CTA as usual

struct ParSortShared {
    int stF[48], stL[48], stD[48];
    int smF[PS_SMALL_MAX], smL[PS_SMALL_MAX], smD[PS_SMALL_MAX], nSmall;
    int sp, first, last, depth, done, K, totGe, totLe, warpGe[PS_MAX_WARPS], warpLe[PS_MAX_WARPS];
};

template <int KeyShift, int THREADS = PS_THREADS>
__device__ inline void par_std_sort(unsigned long long* a, int n, unsigned long long* tmp, uint16_t* posL, uint16_t* posR, ParSortShared& sh) {
    typedef StdSortEmul<KeyShift> S;
    constexpr int PS_THREADS = THREADS, NW = THREADS / 32;
    static_assert(NW >= 1 && NW <= PS_MAX_WARPS, "CTA size");
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (n <= 16) {
        if (tid == 0) S::insertion_sort(a, 0, n);
        __syncthreads();
        return;
    }
    if (tid == 0) {
        int lg = 0;
        while ((1 << (lg + 1)) <= n) ++lg;
        sh.sp = 0; sh.first = 0; sh.last = n; sh.depth = 2 * lg; sh.done = 0; sh.nSmall = 0;
    }
    __syncthreads();
    while (true) {
        if (tid == 0) {
            // next range that needs a partition (ranges of <= 16 elements are left to the final insertion sort)
            int first = sh.first, last = sh.last, depth = sh.depth, sp = sh.sp;
            while (true) {
                if (last - first > 16 && last - first <= PS_SMALL && sh.nSmall < PS_SMALL_MAX) {
                    // short range: parked, finished below by a single thread while other threads finish the other short ranges
                    const int k = sh.nSmall++;
                    sh.smF[k] = first; sh.smL[k] = last; sh.smD[k] = depth;
                    last = first;
                    continue;
                }
                if (last - first > 16) {
                    if (depth == 0) { S::heap_sort(a, first, last); last = first; continue; }
                    --depth;
                    S::move_median_to_first(a, first, first + 1, first + (last - first) / 2, last - 1);
                    break;
                }
                if (sp == 0) { sh.done = 1; break; }
                --sp;
                first = sh.stF[sp]; last = sh.stL[sp]; depth = sh.stD[sp];
            }
            sh.first = first; sh.last = last; sh.depth = depth; sh.sp = sp;
            sh.K = 0x7FFFFFFF;
        }
        __syncthreads();
        if (sh.done) break;
        const int first = sh.first, last = sh.last;
        const int lo = first + 1, len = last - lo;
        const unsigned long long pv = a[first];
        // ---- prefix counts of (>= pivot) from the left and (<= pivot), chunk per thread ----
        const int per = (len + PS_THREADS - 1) / PS_THREADS;
        const int c0 = lo + min(tid * per, len), c1 = lo + min(tid * per + per, len);
        int ge = 0, le = 0;
        for (int x = c0; x < c1; ++x) { const unsigned long long e = a[x]; ge += !S::lt(e, pv); le += !S::lt(pv, e); }
        int ige = ge, ile = le;  // inclusive warp scans
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int g2 = __shfl_up_sync(0xFFFFFFFFu, ige, o), l2 = __shfl_up_sync(0xFFFFFFFFu, ile, o);
            if (lane >= o) { ige += g2; ile += l2; }
        }
        if (lane == 31) { sh.warpGe[wid] = ige; sh.warpLe[wid] = ile; }
        __syncthreads();
        int totGe = 0, totLe = 0, befGe = 0, befLe = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const int g2 = sh.warpGe[w], l2 = sh.warpLe[w];
            if (w < wid) { befGe += g2; befLe += l2; }
            totGe += g2; totLe += l2;
        }
        int rg = ige - ge + befGe;   // # >= pivot before this chunk
        int rl = ile - le + befLe;   // # <= pivot before this chunk
        for (int x = c0; x < c1; ++x) {
            const unsigned long long e = a[x];
            if (!S::lt(e, pv)) posL[rg++] = (uint16_t)x;                 // L_k, k counted from the left
            if (!S::lt(pv, e)) posR[totLe - 1 - rl++] = (uint16_t)x;     // R_k, k counted from the right
        }
        __syncthreads();
        // ---- K = first k with L_k >= R_k (exists: the median-of-three pivot leaves sentinels on both sides) ----
        const int m = min(totGe, totLe);
        for (int k = tid; k < m; k += PS_THREADS)
            if (!(posL[k] < posR[k])) atomicMin(&sh.K, k);
        __syncthreads();
        const int K = min(sh.K, m);
        for (int k = tid; k < K; k += PS_THREADS) {
            const int i = posL[k], j = posR[k];
            const unsigned long long t = a[i]; a[i] = a[j]; a[j] = t;
        }
        __syncthreads();
        if (tid == 0) {
            // the left scan after the last swap stops at the next original >= pivot position or at R_{K-1}, which now holds a
            // swapped-in element >= pivot, whichever comes first
            int cut = K < totGe ? (int)posL[K] : last;
            if (K > 0) cut = min(cut, (int)posR[K - 1]);
            // libstdc++ recurses on [cut, last) and loops on [first, cut); order is irrelevant: keep the larger pending
            int sp = sh.sp;
            if (cut - first < last - cut) {
                sh.stF[sp] = cut; sh.stL[sp] = last; sh.stD[sp] = sh.depth; ++sp;
                sh.last = cut;
            } else {
                sh.stF[sp] = first; sh.stL[sp] = cut; sh.stD[sp] = sh.depth; ++sp;
                sh.first = cut;
            }
            sh.sp = sp;
        }
        __syncthreads();
    }
    // ---- the parked short ranges: the serial introsort loop, one range per thread (the ranges are disjoint) ----
    for (int r = tid; r < sh.nSmall; r += PS_THREADS) S::template introsort_range<16>(a, sh.smF[r], sh.smL[r], sh.smD[r]);
    __syncthreads();
    // ---- final insertion sort == stable placement inside windows of 16 ----
    for (int x = tid; x < n; x += PS_THREADS) {
        const unsigned long long e = a[x];
        int pos = x;
        for (int p = max(x - 16, 0); p < x; ++p) pos -= S::lt(e, a[p]);
        for (int p = x + 1; p <= min(x + 16, n - 1); ++p) pos += S::lt(a[p], e);
        tmp[pos] = e;
    }
    __syncthreads();
    for (int x = tid; x < n; x += PS_THREADS) a[x] = tmp[x];
    __syncthreads();
}

}  // namespace orbx
