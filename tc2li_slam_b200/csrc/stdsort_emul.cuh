// stdsort_emul.cuh — element-for-element emulation of libstdc++'s std::sort (introsort + final insertion sort).
//
// ORBextractor::DistributeOctTree sorts its expandable nodes with std::sort and an ordering that has ties
// (ORBextractor.cc:512-527, 674); std::sort is unstable, so WHICH of two tied nodes is split first — and with it the
// selected keypoints — is a property of the libstdc++ algorithm.  This is a restatement of that algorithm
// (bits/stl_algo.h, bits/stl_heap.h: __introsort_loop, __move_median_to_first, __unguarded_partition,
// __final_insertion_sort, __partial_sort/heap fallback; threshold 16, depth limit 2*floor(log2 n)); SURVEY.md A.7.
//
// Elements are 64-bit words; the comparator looks only at the bits above `KeyShift` (the payload below rides along).
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#define __device__
#endif
#endif

namespace orbx {

template <int KeyShift>
struct StdSortEmul {
    typedef unsigned long long E;
    // KeyShift == 32: the key is the upper word, one 32-bit compare; otherwise a shifted 64-bit compare
    static __host__ __device__ inline bool lt(E a, E b) {
        if (KeyShift == 32) return (uint32_t)(a >> 32) < (uint32_t)(b >> 32);
        return (a >> KeyShift) < (b >> KeyShift);
    }
    static __host__ __device__ inline void swp(E* a, int i, int j) { E t = a[i]; a[i] = a[j]; a[j] = t; }

    static __host__ __device__ void unguarded_linear_insert(E* a, int last) {
        E val = a[last];
        int next = last - 1;
        while (lt(val, a[next])) { a[last] = a[next]; last = next; --next; }
        a[last] = val;
    }
    static __host__ __device__ void insertion_sort(E* a, int first, int last) {
        if (first == last) return;
        for (int i = first + 1; i != last; ++i) {
            if (lt(a[i], a[first])) {
                E val = a[i];
                for (int k = i; k > first; --k) a[k] = a[k - 1];
                a[first] = val;
            } else
                unguarded_linear_insert(a, i);
        }
    }
    // ---- heap fallback (__partial_sort(first, last, last)) ----
    static __host__ __device__ void push_heap(E* a, int first, int hole, int top, E value) {
        int parent = (hole - 1) / 2;
        while (hole > top && lt(a[first + parent], value)) {
            a[first + hole] = a[first + parent];
            hole = parent;
            parent = (hole - 1) / 2;
        }
        a[first + hole] = value;
    }
    static __host__ __device__ void adjust_heap(E* a, int first, int hole, int len, E value) {
        const int top = hole;
        int child = hole;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            if (lt(a[first + child], a[first + child - 1])) --child;
            a[first + hole] = a[first + child];
            hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            a[first + hole] = a[first + child - 1];
            hole = child - 1;
        }
        push_heap(a, first, hole, top, value);
    }
    static __host__ __device__ void heap_sort(E* a, int first, int last) {
        const int len = last - first;
        if (len >= 2) {
            int parent = (len - 2) / 2;
            while (true) {
                E value = a[first + parent];
                adjust_heap(a, first, parent, len, value);
                if (parent == 0) break;
                --parent;
            }
        }
        while (last - first > 1) {
            --last;
            E value = a[last];
            a[last] = a[first];
            adjust_heap(a, first, 0, last - first, value);
        }
    }
    static __host__ __device__ void move_median_to_first(E* a, int result, int x, int y, int z) {
        if (lt(a[x], a[y])) {
            if (lt(a[y], a[z])) swp(a, result, y);
            else if (lt(a[x], a[z])) swp(a, result, z);
            else swp(a, result, x);
        } else if (lt(a[x], a[z])) swp(a, result, x);
        else if (lt(a[y], a[z])) swp(a, result, z);
        else swp(a, result, y);
    }
    static __host__ __device__ int unguarded_partition(E* a, int first, int last, int pivot) {
        while (true) {
            while (lt(a[first], a[pivot])) ++first;
            --last;
            while (lt(a[pivot], a[last])) --last;
            if (!(first < last)) return first;
            swp(a, first, last);
            ++first;
        }
    }
    // __introsort_loop on the sub-range [first0, last0) with `depth0` splits left: everything std::sort does to that range before
    // the final insertion sort.  Ranges are independent, so different threads may run this on disjoint ranges concurrently.
    template <int StackN>
    static __host__ __device__ void introsort_range(E* a, int first0, int last0, int depth0) {
        int stackFirst[StackN], stackLast[StackN], stackDepth[StackN], sp = 0;
        stackFirst[0] = first0; stackLast[0] = last0; stackDepth[0] = depth0; sp = 1;
        while (sp > 0) {
            --sp;
            int first = stackFirst[sp], last = stackLast[sp], depth = stackDepth[sp];
            while (last - first > 16) {
                if (depth == 0) { heap_sort(a, first, last); break; }
                --depth;
                const int mid = first + (last - first) / 2;
                move_median_to_first(a, first, first + 1, mid, last - 1);
                const int cut = unguarded_partition(a, first + 1, last, first);
                if (cut - first < last - cut) {  // keep the larger one pending: the stack stays below log2 of the range length
                    stackFirst[sp] = cut; stackLast[sp] = last; stackDepth[sp] = depth; ++sp;
                    last = cut;
                } else {
                    stackFirst[sp] = first; stackLast[sp] = cut; stackDepth[sp] = depth; ++sp;
                    first = cut;
                }
            }
        }
    }
    // forceHeapDepth >= 0 overrides the depth limit (tests only: exercises the heap fallback)
    static __host__ __device__ void sort(E* a, int n, int forceDepth = -1) {
        if (n <= 0) return;
        int lg = 0;
        while ((1 << (lg + 1)) <= n) ++lg;
        int stackFirst[64], stackLast[64], stackDepth[64], sp = 0;
        stackFirst[0] = 0; stackLast[0] = n; stackDepth[0] = forceDepth >= 0 ? forceDepth : 2 * lg; sp = 1;
        while (sp > 0) {
            --sp;
            int first = stackFirst[sp], last = stackLast[sp], depth = stackDepth[sp];
            while (last - first > 16) {
                if (depth == 0) { heap_sort(a, first, last); break; }
                --depth;
                const int mid = first + (last - first) / 2;
                move_median_to_first(a, first, first + 1, mid, last - 1);
                const int cut = unguarded_partition(a, first + 1, last, first);
                // libstdc++ recurses on [cut, last) and loops on [first, cut); the two ranges are independent, so the
                // result does not depend on which is finished first: keep the larger one pending (stack <= log2 n).
                if (cut - first < last - cut) {
                    stackFirst[sp] = cut; stackLast[sp] = last; stackDepth[sp] = depth; ++sp;
                    last = cut;
                } else {
                    stackFirst[sp] = first; stackLast[sp] = cut; stackDepth[sp] = depth; ++sp;
                    first = cut;
                }
            }
        }
        if (n > 16) {
            insertion_sort(a, 0, 16);
            for (int i = 16; i < n; ++i) unguarded_linear_insert(a, i);
        } else
            insertion_sort(a, 0, n);
    }
};

}  // namespace orbx
