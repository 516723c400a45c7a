// grid_kernel.cu — the feature grid of a Frame on the device (widening).
//
//   k_assign_grid      : Frame::AssignFeaturesToGrid (Frame.cc:412-443) + Frame::PosInGrid (Frame.cc:751-762) for the single-grid
//                        case (Nleft == -1) on the keypoints the extractor just produced; with rectified input (zero
//                        distortion) Frame::UndistortKeyPoints leaves mvKeysUn == mvKeys (Frame.cc:777-782).  mGrid[i][j]
//                        (FRAME_GRID_COLS = 64 x FRAME_GRID_ROWS = 48 vectors, Frame.h:59-60, 278) becomes CSR: cell c = i * 48 + j
//                        holds items[start[c] .. start[c + 1]) = keypoint indices in push_back order (ascending).
//                        One CTA per image: shared-memory histogram, block scan, scatter through shared cursors, then every
//                        cell list (0.65 entries on average) is put in ascending order by the thread that owns the cell.
//   k_features_in_area : Frame::GetFeaturesInArea (Frame.cc:687-749), one warp per query.  The cells of one grid column that a
//                        query touches are adjacent in the CSR, so the reference's (ix, iy, in-cell) visiting order is one
//                        contiguous scan per column; hits are appended in order with ballot prefixes.
// Float expressions follow the reference operation by operation (sub, sub/add, mul in float32; round half away from zero;
// floor / ceil; strict |d| < r); no FMA contraction (--fmad=false) — although none of them has a multiply-add shape.
#include <algorithm>

#include "orbx_internal.cuh"

namespace orbx {

namespace {

constexpr int GC = ORBX_GRID_COLS, GR = ORBX_GRID_ROWS, NC = GC * GR;
constexpr int GT = 256;             // threads per CTA of k_assign_grid
constexpr int CPT = NC / GT;        // cells per thread (12)
static_assert(NC % GT == 0, "cells per thread");

__device__ __forceinline__ int cell_of(float x, float y, const GridParams& p) {
    const int px = (int)roundf(__fmul_rn(__fsub_rn(x, p.minX), p.invW));
    const int py = (int)roundf(__fmul_rn(__fsub_rn(y, p.minY), p.invH));
    return (px < 0 || px >= GC || py < 0 || py >= GR) ? -1 : px * GR + py;
}

}  // namespace

__global__ void __launch_bounds__(GT) k_assign_grid(const Buffers b, int K, GridParams p) {
    extern __shared__ int items[];       // [K]
    __shared__ int cnt[NC];              // histogram, then the running cursor of every cell
    __shared__ int start[NC + 1];
    __shared__ int warpSum[GT / 32];
    const int img = b.imgBase + blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = b.outCount[img];
    const orbx_keypoint* kps = b.kps + (size_t)img * K;
    for (int c = tid; c < NC; c += GT) cnt[c] = 0;
    __syncthreads();
    const float2* un = b.kpsUn ? b.kpsUn + (size_t)img * K : nullptr;   // Frame::mvKeysUn when the caller undistorted the keypoints
    for (int i = tid; i < n; i += GT) {
        const int c = un ? cell_of(un[i].x, un[i].y, p) : cell_of(kps[i].x, kps[i].y, p);
        if (c >= 0) atomicAdd(&cnt[c], 1);
    }
    __syncthreads();
    // exclusive scan over the cells: thread t owns cells [t * CPT, (t + 1) * CPT)
    int local[CPT], sum = 0;
#pragma unroll
    for (int k = 0; k < CPT; ++k) { local[k] = sum; sum += cnt[tid * CPT + k]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warpSum[wid] = incl;
    __syncthreads();
    int base = incl - sum;
    for (int w = 0; w < wid; ++w) base += warpSum[w];
#pragma unroll
    for (int k = 0; k < CPT; ++k) start[tid * CPT + k] = base + local[k];
    if (tid == GT - 1) start[NC] = base + sum;
    __syncthreads();
    for (int c = tid; c < NC; c += GT) cnt[c] = start[c];
    __syncthreads();
    for (int i = tid; i < n; i += GT) {
        const int c = un ? cell_of(un[i].x, un[i].y, p) : cell_of(kps[i].x, kps[i].y, p);
        if (c >= 0) items[atomicAdd(&cnt[c], 1)] = i;
    }
    __syncthreads();
    // push_back order = ascending keypoint index: insertion sort of every (short) cell list by its owner
    for (int k = 0; k < CPT; ++k) {
        const int c = tid * CPT + k, s = start[c], e = start[c + 1];
        for (int a = s + 1; a < e; ++a) {
            const int v = items[a];
            int q = a - 1;
            while (q >= s && items[q] > v) { items[q + 1] = items[q]; --q; }
            items[q + 1] = v;
        }
    }
    __syncthreads();
    int* gs = b.gridStart + (size_t)img * (NC + 1);
    int* gi = b.gridItems + (size_t)img * K;
    for (int c = tid; c <= NC; c += GT) gs[c] = start[c];
    const int total = start[NC];
    for (int i = tid; i < total; i += GT) gi[i] = items[i];
}

// queries: {image, x, y, r, minLevel, maxLevel}; out: [nQueries][cap] indices, counts[nQueries] = what the reference would return
__global__ void __launch_bounds__(128) k_features_in_area(const Buffers b, int K, GridParams p, const orbx_area_query* __restrict__ queries, int nQueries, int cap,
                                                         int* __restrict__ counts, int* __restrict__ out) {
    const int q = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= nQueries) return;
    const orbx_area_query Q = queries[q];
    const float x = Q.x, y = Q.y, r = Q.r;
    int n = 0;
    // (x - mnMinX - factorX) * mfGridElementWidthInv etc. (Frame.cc:695-717), left to right in float32
    const int minCX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, p.minX), r), p.invW)));
    const int maxCX = min(GC - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, p.minX), r), p.invW)));
    const int minCY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, p.minY), r), p.invH)));
    const int maxCY = min(GR - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, p.minY), r), p.invH)));
    if (!(minCX >= GC || maxCX < 0 || minCY >= GR || maxCY < 0)) {
        const bool checkLevels = Q.min_level > 0 || Q.max_level >= 0;
        const int* gs = b.gridStart + (size_t)Q.image * (NC + 1);
        const int* gi = b.gridItems + (size_t)Q.image * K;
        const orbx_keypoint* kps = b.kps + (size_t)Q.image * K;
        int* o = out + (size_t)q * cap;
        for (int ix = minCX; ix <= maxCX; ++ix) {
            const int s = gs[ix * GR + minCY], e = gs[ix * GR + maxCY + 1];
            for (int j0 = s; j0 < e; j0 += 32) {
                const int j = j0 + lane;
                bool hit = false;
                int idx = 0;
                if (j < e) {
                    idx = gi[j];
                    orbx_keypoint kp = kps[idx];
                    if (b.kpsUn) { const float2 u = b.kpsUn[(size_t)Q.image * K + idx]; kp.x = u.x; kp.y = u.y; }
                    bool lv = true;
                    if (checkLevels) lv = !(kp.octave < Q.min_level) && !(Q.max_level >= 0 && kp.octave > Q.max_level);
                    hit = lv && fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r;
                }
                const unsigned bal = __ballot_sync(0xFFFFFFFFu, hit);
                const int pos = n + __popc(bal & ((1u << lane) - 1));
                if (hit && pos < cap) o[pos] = idx;
                n += __popc(bal);
            }
        }
    }
    if (lane == 0) counts[q] = n;
}

// Frame::UndistortKeyPoints (Frame.cc:777-810) = cv::undistortPoints(mat, mat, toK(), mDistCoef, cv::Mat(), mK): OpenCV's fixed-point inversion of the distortion
// model, five iterations, in double, operation by operation (every product and sum an explicit round-to-nearest double op: no contraction), rounded to float once.
__global__ void __launch_bounds__(128) k_undistort(const Buffers b, int K, const __grid_constant__ UndistortParams p, float2* __restrict__ out) {
    const int img = b.imgBase + blockIdx.y, i = blockIdx.x * 128 + threadIdx.x;
    if (i >= b.outCount[img]) return;
    const orbx_keypoint kp = b.kps[(size_t)img * K + i];
    const double u = (double)kp.x, v = (double)kp.y;
    double x = __dmul_rn(__dsub_rn(u, p.cx), p.ifx), y = __dmul_rn(__dsub_rn(v, p.cy), p.ify);
    const double x0 = x, y0 = y;
    const double* k = p.k;
    for (int j = 0; j < 5; ++j) {
        const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
        const double num = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[7], r2), k[6]), r2), k[5]), r2));
        const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[4], r2), k[1]), r2), k[0]), r2));
        const double icdist = __ddiv_rn(num, den);
        if (icdist < 0) {
            x = __dmul_rn(__dsub_rn(u, p.cx), p.ifx);
            y = __dmul_rn(__dsub_rn(v, p.cy), p.ify);
            break;
        }
        // deltaX = 2*k[2]*x*y + k[3]*(r2 + 2*x*x) + k[8]*r2 + k[9]*r2*r2;  deltaY = k[2]*(r2 + 2*y*y) + 2*k[3]*x*y + k[10]*r2 + k[11]*r2*r2  (left to right)
        const double dX = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, k[2]), x), y), __dmul_rn(k[3], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x)))),
                                              __dmul_rn(k[8], r2)),
                                    __dmul_rn(__dmul_rn(k[9], r2), r2));
        const double dY = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(k[2], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))), __dmul_rn(__dmul_rn(__dmul_rn(2.0, k[3]), x), y)),
                                              __dmul_rn(k[10], r2)),
                                    __dmul_rn(__dmul_rn(k[11], r2), r2));
        x = __dmul_rn(__dsub_rn(x0, dX), icdist);
        y = __dmul_rn(__dsub_rn(y0, dY), icdist);
    }
    // RR = P * I: xx = fx' x + 0 y + cx', yy = 0 x + fy' y + cy', ww = 1 / (0 x + 0 y + 1)
    const double xx = __dadd_rn(__dadd_rn(__dmul_rn(p.nfx, x), __dmul_rn(0.0, y)), p.ncx);
    const double yy = __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(p.nfy, y)), p.ncy);
    const double ww = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(0.0, y)), 1.0));
    out[(size_t)img * K + i] = make_float2((float)__dmul_rn(xx, ww), (float)__dmul_rn(yy, ww));
}

void launch_undistort(const Buffers& b, int K, const UndistortParams& p, float2* out, int nImages, cudaStream_t s) {
    dim3 grid((unsigned)((K + 127) / 128), (unsigned)nImages);
    k_undistort<<<grid, 128, 0, s>>>(b, K, p, out);
}

size_t grid_smem_bytes(int K) { return (size_t)K * sizeof(int); }

cudaError_t grid_configure(int K) {
    return raise_dynamic_smem(k_assign_grid, grid_smem_bytes(K));
}

void launch_assign_grid(const Buffers& b, int K, const GridParams& p, int nImages, cudaStream_t s) {
    k_assign_grid<<<nImages, GT, grid_smem_bytes(K), s>>>(b, K, p);
}

void launch_features_in_area(const Buffers& b, int K, const GridParams& p, const orbx_area_query* queries, int nQueries, int cap, int* counts, int* out,
                             cudaStream_t s) {
    k_features_in_area<<<(nQueries + 3) / 4, 128, 0, s>>>(b, K, p, queries, nQueries, cap, counts, out);
}

}  // namespace orbx
