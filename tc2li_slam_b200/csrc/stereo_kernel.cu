// stereo_kernel.cu — Frame::ComputeStereoMatches on the device (the step right after the extractor).
//
// Replaces slam_framework/src/Frame.cc:841-1011 (+ ORBmatcher::DescriptorDistance, ORBmatcher.cc:2067-2083): for every left
// keypoint the best right keypoint by Hamming distance among those whose row band covers the left keypoint's row, same or
// adjacent octave, disparity in [0, mbf/mb]; 11x11 SAD refinement over 11 shifts on the left keypoint's pyramid level of
// both images; parabola fit; rejection of matches whose SAD reaches 2.1 x the median SAD.  Keypoints, descriptors and both
// pyramids are already resident (the extractor produced them), so nothing crosses PCIe except the two float results.
//
// One CTA per stereo frame.
//   1  right keypoints into shared memory, counting-sorted by integer row.  The reference builds per-row candidate lists; a
//      contiguous bucket range plus the exact band test yields the same candidate SET, and its "first minimum in index
//      order" is the minimum over (distance, index).
//   2  one warp per left keypoint: lanes stride the candidates (XOR + POPC on 8 words), warp min-reduce; then the SAD search:
//      both windows staged in shared memory, one lane per shift, VABSDIFF4 + IDP.4A on funnel-shifted words; parabola fit.
//   3  radix select of the median SAD over the matches, rejection.
// Float expressions follow the reference's operation order with explicit _rn intrinsics; results are bit-identical.
#include <algorithm>

#include "orbx_internal.cuh"

namespace orbx {

namespace {

constexpr int SM_THREADS = 256, SM_WARPS = 8;
constexpr int TH_HIGH = 100, TH_LOW = 50;  // ORBmatcher.cc:44-45
constexpr unsigned FULL = 0xFFFFFFFFu;

}  // namespace

size_t stereo_smem_bytes(int K, int H) { return (size_t)K * (4 + 4 + 4 + 2 + 1) + (size_t)(H + 2) * 8 + 64; }

__global__ void __launch_bounds__(SM_THREADS) k_stereo_match(const __grid_constant__ Geom g, const Buffers bl, const Buffers br, int lMul, int lAdd, int rMul,
                                                            int rAdd, float mb, float mbf) {
    extern __shared__ __align__(16) unsigned char ssm[];
    __shared__ uint32_t winL[SM_WARPS][11 * 3];   // 11 rows x 12 bytes (11 used)
    __shared__ uint32_t winR[SM_WARPS][11 * 6];   // 11 rows x 24 bytes (21 used)
    __shared__ int hist[256];
    __shared__ int part[SM_THREADS + 1];
    __shared__ int sh_nvalid, sh_bin, sh_before, sh_median;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int imgL = bl.imgBase + lMul * blockIdx.x + lAdd, imgR = br.imgBase + rMul * blockIdx.x + rAdd;
    const int K = g.K, H = g.lv[0].h;
    float* xR = reinterpret_cast<float*>(ssm);
    float* yR = xR + K;
    int* sad = reinterpret_cast<int*>(yR + K);
    int* rowStart = sad + K;            // [H + 2]
    int* rowCur = rowStart + (H + 2);   // [H + 2]
    uint16_t* rowIdx = reinterpret_cast<uint16_t*>(rowCur + (H + 2));
    uint8_t* octR = reinterpret_cast<uint8_t*>(rowIdx + K);

    const int N = min(bl.outCount[imgL], K), Nr = min(br.outCount[imgR], K);
    const orbx_keypoint* kpsL = bl.kps + (size_t)imgL * K;
    const orbx_keypoint* kpsR = br.kps + (size_t)imgR * K;
    const uint4* descL = reinterpret_cast<const uint4*>(bl.desc + (size_t)imgL * K * 32);
    const uint4* descR = reinterpret_cast<const uint4*>(br.desc + (size_t)imgR * K * 32);
    float* uRight = bl.uRight + (size_t)imgL * K;
    float* depth = bl.depth + (size_t)imgL * K;

    // ---- phase 1: right keypoints, counting sort by integer row ----
    for (int k = tid; k < H + 2; k += SM_THREADS) rowStart[k] = 0;
    for (int k = tid; k < N; k += SM_THREADS) { uRight[k] = -1.0f; depth[k] = -1.0f; sad[k] = -1; }
    if (tid == 0) sh_nvalid = 0;
    __syncthreads();
    for (int k = tid; k < Nr; k += SM_THREADS) {
        const orbx_keypoint kp = kpsR[k];
        xR[k] = kp.x; yR[k] = kp.y; octR[k] = (uint8_t)kp.octave;
        atomicAdd(&rowStart[min(max((int)kp.y, 0), H - 1) + 1], 1);
    }
    __syncthreads();
    {   // exclusive scan: rowStart[r] = number of right keypoints in rows < r
        const int n = H + 1, per = (n + SM_THREADS - 1) / SM_THREADS;
        const int lo = min(tid * per, n), hi = min(lo + per, n);
        int sum = 0;
        for (int k = lo; k < hi; ++k) sum += rowStart[k];
        part[tid] = sum;
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int k = 0; k < SM_THREADS; ++k) { const int v = part[k]; part[k] = run; run += v; }
        }
        __syncthreads();
        int run = part[tid];
        for (int k = lo; k < hi; ++k) { run += rowStart[k]; rowStart[k] = run; }  // inclusive: rowStart[r] = count in rows <= r-1
        __syncthreads();
        for (int k = tid; k < H + 1; k += SM_THREADS) rowCur[k] = rowStart[k];
        __syncthreads();
    }
    for (int k = tid; k < Nr; k += SM_THREADS) rowIdx[atomicAdd(&rowCur[min(max((int)yR[k], 0), H - 1)], 1)] = (uint16_t)k;
    __syncthreads();

    // ---- phase 2: one warp per left keypoint ----
    const int bandRows = (int)ceilf(2.0f * g.lv[g.nlevels - 1].sf) + 2;
    const float minD = 0.f, maxD = __fdiv_rn(mbf, mb);
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    for (int iL = wid; iL < N; iL += SM_WARPS) {
        const orbx_keypoint kp = kpsL[iL];
        const int levelL = kp.octave;
        const float vL = kp.y, uL = kp.x;
        const int vLi = (int)vL;  // vRowIndices[vL]
        const float minU = __fsub_rn(uL, maxD), maxU = __fsub_rn(uL, minD);
        int best = (TH_HIGH << 16) | 0xFFFF;
        if (!(maxU < 0)) {
            const uint4 a0 = __ldg(descL + 2 * iL), a1 = __ldg(descL + 2 * iL + 1);
            const int c0 = rowStart[max(vLi - bandRows, 0)], c1 = rowStart[min(vLi + bandRows, H - 1) + 1];
            for (int c = c0 + lane; c < c1; c += 32) {
                const int iR = rowIdx[c];
                const int oc = octR[iR];
                if (oc < levelL - 1 || oc > levelL + 1) continue;
                const float yr = yR[iR], r = __fmul_rn(2.0f, g.lv[oc].sf);
                const int maxr = (int)ceilf(__fadd_rn(yr, r)), minr = (int)floorf(__fsub_rn(yr, r));
                if (vLi < minr || vLi > maxr) continue;  // not in vRowIndices[vL]
                const float xr = xR[iR];
                if (!(xr >= minU && xr <= maxU)) continue;
                const uint4 b0 = __ldg(descR + 2 * iR), b1 = __ldg(descR + 2 * iR + 1);
                const int dist = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
                                 __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
                best = min(best, (dist << 16) | iR);
            }
            best = __reduce_min_sync(FULL, best);
        }
        const int bestDist = best >> 16;
        if (bestDist >= thOrbDist) continue;  // warp-uniform

        // ---- sub-pixel match by correlation (Frame.cc:927-1000) ----
        const float uR0 = xR[best & 0xFFFF];
        const float sfL = g.lv[levelL].sf, isf = __fdiv_rn(1.0f, sfL);  // mvInvScaleFactors[octave] = 1.0f / mvScaleFactor (ORBextractor.cc:402)
        const float scaleduL = roundf(__fmul_rn(uL, isf)), scaledvL = roundf(__fmul_rn(vL, isf)), scaleduR0 = roundf(__fmul_rn(uR0, isf));
        const float iniu = __fsub_rn(__fadd_rn(scaleduR0, 5.f), 5.f), endu = __fadd_rn(__fadd_rn(__fadd_rn(scaleduR0, 5.f), 5.f), 1.f);
        const LevelGeom& LV = g.lv[levelL];
        if (iniu < 0 || endu >= (float)LV.w) continue;
        const int y0 = (int)__fsub_rn(scaledvL, 5.f), xl0 = (int)__fsub_rn(scaleduL, 5.f), xr0 = (int)scaleduR0 - 10;
        {
            const uint8_t* PL = bl.pyr + LV.planeOff + (size_t)imgL * LV.planeBytes + (ptrdiff_t)y0 * LV.pitch + xl0;
            const uint8_t* PR = br.pyr + LV.planeOff + (size_t)imgR * LV.planeBytes + (ptrdiff_t)y0 * LV.pitch + xr0;
            uint8_t* wl = reinterpret_cast<uint8_t*>(winL[wid]);
            uint8_t* wr = reinterpret_cast<uint8_t*>(winR[wid]);
            for (int e = lane; e < 11 * 12; e += 32) {
                const int r = e / 12, c = e - r * 12;
                wl[e] = c < 11 ? __ldg(PL + (ptrdiff_t)r * LV.pitch + c) : 0;
            }
            for (int e = lane; e < 11 * 24; e += 32) {
                const int r = e / 24, c = e - r * 24;
                wr[e] = c < 21 ? __ldg(PR + (ptrdiff_t)r * LV.pitch + c) : 0;
            }
        }
        __syncwarp();
        int mySad = 0x7FFFFFF;  // lanes >= 11 do not take part
        if (lane < 11) {
            int acc = 0;
#pragma unroll
            for (int r = 0; r < 11; ++r) {
                const uint32_t* lw = winL[wid] + r * 3;
                const uint32_t* rw = winR[wid] + r * 6;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int off = lane + 4 * k;
                    uint32_t rv = __funnelshift_r(rw[off >> 2], rw[(off >> 2) + 1], 8 * (off & 3));
                    uint32_t lv = lw[k];
                    if (k == 2) { rv &= 0x00FFFFFFu; lv &= 0x00FFFFFFu; }
                    acc = (int)__dp4a(__vabsdiffu4(lv, rv), 0x01010101u, (uint32_t)acc);
                }
            }
            mySad = acc;  // == cv::norm(IL, IR, NORM_L1), exact in float
        }
        // first minimum over the shifts in order incR = -5 .. 5
        const int bestPacked = __reduce_min_sync(FULL, (mySad << 4) | (lane & 15));
        const int sBest = bestPacked & 15, sadBest = bestPacked >> 4;
        const int d1i = __shfl_sync(FULL, mySad, max(sBest - 1, 0)), d3i = __shfl_sync(FULL, mySad, min(sBest + 1, 10));
        __syncwarp();  // the windows are free again
        if (sBest == 0 || sBest == 10) continue;  // bestincR == -L || bestincR == L
        const float dist1 = (float)d1i, dist2 = (float)sadBest, dist3 = (float)d3i;
        const float deltaR = __fdiv_rn(__fsub_rn(dist1, dist3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(dist1, dist3), __fmul_rn(2.0f, dist2))));
        if (deltaR < -1 || deltaR > 1) continue;
        float bestuR = __fmul_rn(sfL, __fadd_rn(__fadd_rn(scaleduR0, (float)(sBest - 5)), deltaR));
        float disparity = __fsub_rn(uL, bestuR);
        if (disparity >= minD && disparity < maxD) {
            if (disparity <= 0) {
                disparity = 0.01f;
                bestuR = (float)((double)uL - 0.01);
            }
            if (lane == 0) {
                depth[iL] = __fdiv_rn(mbf, disparity);
                uRight[iL] = bestuR;
                sad[iL] = sadBest;
                atomicAdd(&sh_nvalid, 1);
            }
        }
    }
    __syncthreads();

    // ---- phase 3: matches whose SAD reaches 1.5 * 1.4 * median are rejected (Frame.cc:1003-1018) ----
    const int nvalid = sh_nvalid;
    if (nvalid == 0) return;  // (the reference indexes an empty vector here)
    const int rank = nvalid / 2;
    for (int k = tid; k < 256; k += SM_THREADS) hist[k] = 0;
    __syncthreads();
    for (int k = tid; k < N; k += SM_THREADS)
        if (sad[k] >= 0) atomicAdd(&hist[sad[k] >> 7], 1);  // SAD <= 121 * 255 < 2^15
    __syncthreads();
    if (tid == 0) {
        int run = 0, bin = 0;
        for (; bin < 256; ++bin) { if (run + hist[bin] > rank) break; run += hist[bin]; }
        sh_bin = bin; sh_before = run;
    }
    __syncthreads();
    const int bin = sh_bin;
    for (int k = tid; k < 256; k += SM_THREADS) hist[k] = 0;
    __syncthreads();
    for (int k = tid; k < N; k += SM_THREADS)
        if (sad[k] >= 0 && (sad[k] >> 7) == bin) atomicAdd(&hist[sad[k] & 127], 1);
    __syncthreads();
    if (tid == 0) {
        int run = sh_before, lo = 0;
        for (; lo < 128; ++lo) { if (run + hist[lo] > rank) break; run += hist[lo]; }
        sh_median = (bin << 7) | lo;
    }
    __syncthreads();
    const float thDist = __fmul_rn(__fmul_rn(1.5f, 1.4f), (float)sh_median);
    for (int k = tid; k < N; k += SM_THREADS)
        if (sad[k] >= 0 && !((float)sad[k] < thDist)) { uRight[k] = -1.0f; depth[k] = -1.0f; }
}

cudaError_t stereo_configure(int K, int H) {
    return raise_dynamic_smem(k_stereo_match, stereo_smem_bytes(K, H));
}

void launch_stereo_match(const Geom& g, const Buffers& bl, const Buffers& br, int nFrames, int lMul, int lAdd, int rMul, int rAdd, float mb, float mbf,
                         cudaStream_t s) {
    k_stereo_match<<<nFrames, SM_THREADS, stereo_smem_bytes(g.K, g.lv[0].h), s>>>(g, bl, br, lMul, lAdd, rMul, rAdd, mb, mbf);
}

}  // namespace orbx
