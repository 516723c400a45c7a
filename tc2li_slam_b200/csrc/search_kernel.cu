// search_kernel.cu — the search of ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono)
// (ORBmatcher.cc:1685-1895) on the device, from the point where a map point has been projected (ORBmatcher.cc:1729 on), for the
// single-grid case (CurrentFrame.Nleft == -1).  The pose algebra in front of it (Sophus / Eigen: Tcw * x3Dw, project, the
// bForward / bBackward decision) stays with the caller, who sends one orbx_proj_query per surviving map point; keypoints,
// descriptors, right coordinates (orbx_batch_stereo_match) and the feature grid (orbx_batch_assign_grid) are already resident.
//
// One CTA of 1024 threads per call (a frame).  The frame's grid, keypoint coordinates / levels / angles, right coordinates and — when
// they fit — its descriptors are staged in shared memory first, so nothing below waits for global memory (the walks are chains of
// dependent loads).  Phase 1, a warp per query, twice over the same candidate walk (Frame::GetFeaturesInArea order,
// see grid_kernel.cu): first the candidate counts, then — after a block scan gave every query its slice of a shared-memory
// CSR — the candidates themselves as (Hamming distance << 16 | keypoint index), with the right-coordinate gate
// (ORBmatcher.cc:1760-1766) applied.  Phase 2 is the reference's own sequential dependency: a keypoint taken by a point with
// observations is skipped by every later point (ORBmatcher.cc:1756-1758), so one warp walks the queries in order, lanes over
// the candidates of a query, first minimum by a (distance, position) key, acceptance at bestDist <= TH_HIGH; everything it
// touches is in shared memory.  Phase 3: rotation histogram, ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:2021-2062), removal
// of the matches outside the three dominant bins (ORBmatcher.cc:1872-1892).  Integer work except the window arithmetic, which
// follows the reference operation by operation in float32.  The template parameter switches to the local-map overload
// ORBmatcher::SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints) (ORBmatcher.cc:52-222): best / second best with levels
// and the mfNNratio test instead of the rotation check, keypoints taken before the search as an input.
//
// Two-grid frames (fisheye stereo, Frame::Nleft != -1; ORBmatcher.cc:153-218 and :1803-1868): image2 >= 0 names the right image of the
// frame.  Both images' grids / keypoints / descriptors are staged (the right ones behind the left ones, keypoint k of the right image
// is entry K + k), a map point arrives as up to two sub-queries in a row told apart by side[i] (0 = left image, 1 = right image after
// its left part, 2 = right image alone), and the sequential pass replays what the reference's `continue`s do to the right part: an EMPTY
// left window cancels it in the frame-to-frame search (:1744-1745), a failed ratio test in the local-map search (:133-134).  An accepted
// local-map match also takes the partner keypoint of mvLeftToRightMatch / mvRightToLeftMatch (:139-143, :205-209).  No right-coordinate gate.
#include <algorithm>

#include "orbx_internal.cuh"

namespace orbx {

namespace {

constexpr int ST = 1024;  // threads per CTA: the candidate walks are latency bound, one CTA serves a frame
constexpr int GC = ORBX_GRID_COLS, GR = ORBX_GRID_ROWS, NC = GC * GR;
constexpr int HISTO = 30;  // ORBmatcher::HISTO_LENGTH (ORBmatcher.cc:46)

struct Window {
    int minCX, maxCX, minCY, maxCY;
    bool empty;
};

__device__ __forceinline__ Window window_of(float x, float y, float r, const GridParams& p) {  // Frame.cc:695-717
    Window w;
    w.minCX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, p.minX), r), p.invW)));
    w.maxCX = min(GC - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, p.minX), r), p.invW)));
    w.minCY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, p.minY), r), p.invH)));
    w.maxCY = min(GR - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, p.minY), r), p.invH)));
    w.empty = w.minCX >= GC || w.maxCX < 0 || w.minCY >= GR || w.maxCY < 0;
    return w;
}

// the candidate walk of one query by one warp; EMIT = false: count only.  Returns the number of candidates (after the gate).
template <bool EMIT>
__device__ __forceinline__ int walk(const orbx_proj_query& Q, const Window& W, const int* gs, const uint16_t* gi, const float2* kxy, const uint8_t* koct,
                                    const uint8_t* desc, const float* uRight, uint32_t* out, int lane, int base) {
    if (W.empty) return 0;
    const bool checkLevels = Q.min_level > 0 || Q.max_level >= 0;
    uint32_t qd[8];
    if (EMIT) {
#pragma unroll
        for (int k = 0; k < 8; ++k) qd[k] = reinterpret_cast<const uint32_t*>(Q.desc)[k];
    }
    int n = 0;
    for (int ix = W.minCX; ix <= W.maxCX; ++ix) {
        const int s = gs[ix * GR + W.minCY], e = gs[ix * GR + W.maxCY + 1];
        for (int j0 = s; j0 < e; j0 += 32) {
            const int j = j0 + lane;
            bool hit = false;
            int idx = 0, kpOctave = 0;
            if (j < e) {
                idx = base + gi[j];
                const float2 kp = kxy[idx];
                kpOctave = koct[idx];
                bool lv = true;
                if (checkLevels) lv = !(kpOctave < Q.min_level) && !(Q.max_level >= 0 && kpOctave > Q.max_level);
                hit = lv && fabsf(__fsub_rn(kp.x, Q.u)) < Q.radius && fabsf(__fsub_rn(kp.y, Q.v)) < Q.radius;
                if (hit && uRight) {  // CurrentFrame.mvuRight[i2] > 0: the predicted right coordinate must fall inside the window too
                    const float ur2 = uRight[idx];
                    if (ur2 > 0.f && fabsf(__fsub_rn(Q.ur, ur2)) > Q.radius) hit = false;
                }
            }
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, hit);
            if (EMIT && hit) {
                const uint4* d = reinterpret_cast<const uint4*>(desc + (size_t)(idx - base) * 32);   // desc: this image's rows
                const uint4 a = d[0], c = d[1];
                const int dist = __popc(a.x ^ qd[0]) + __popc(a.y ^ qd[1]) + __popc(a.z ^ qd[2]) + __popc(a.w ^ qd[3]) + __popc(c.x ^ qd[4]) + __popc(c.y ^ qd[5]) +
                                 __popc(c.z ^ qd[6]) + __popc(c.w ^ qd[7]);
                out[n + __popc(bal & ((1u << lane) - 1))] = ((uint32_t)dist << 20) | ((uint32_t)kpOctave << 16) | (uint32_t)idx;
            }
            n += __popc(bal);
        }
    }
    return n;
}

}  // namespace

// shared memory (KK = K per image of the frame): match[KK] | candStart[nq + 1] | cand[candCap] | gs[sides][3073] | kxy[KK] | ur[KK] | kang[KK] | qang[nq] | gi[KK] (u16) |
// entIdx[nq] (u16) | link[KK] (i16) | entBin[nq] | taken[KK] | koct[KK] | qblk[nq] | qside[nq]
// LOCAL = false: frame-to-frame search (best distance, rotation check); LOCAL = true: local-map search (best / second best with
// their levels and the mfNNratio test, keypoints taken before the search in taken0, no rotation check)
template <bool LOCAL>
__global__ void __launch_bounds__(ST) k_search_by_projection(const Buffers b, int image, int K, GridParams p, const orbx_proj_query* __restrict__ queries, int nq,
                                                            int useURight, int checkOrientation, int thHigh, float nnRatio, const uint8_t* __restrict__ taken0,
                                                            int candCap, int stageDesc, int* __restrict__ outMatch, int* __restrict__ outInfo, int image2,
                                                            const uint8_t* __restrict__ side, const int* __restrict__ links) {
    extern __shared__ __align__(16) unsigned char ssm[];
    __shared__ int warpSum[ST / 32];
    __shared__ int hist[HISTO];
    __shared__ int sInfo[4];  // nmatches, entries, total candidates, keep-bin mask
    const int sides = image2 >= 0 ? 2 : 1, KK = K * sides;
    int* match = reinterpret_cast<int*>(ssm);
    int* candStart = match + KK;
    uint32_t* cand = reinterpret_cast<uint32_t*>(candStart + nq + 1);
    int* gsS = reinterpret_cast<int*>(cand + candCap);
    float2* kxy = reinterpret_cast<float2*>((reinterpret_cast<uintptr_t>(gsS + sides * (NC + 1)) + 7) & ~(uintptr_t)7);
    float* urS = reinterpret_cast<float*>(kxy + KK);
    float* kang = urS + KK;     // keypoint angles (rotation histogram)
    float* qang = kang + KK;    // per query: kpLF.angle
    uint16_t* giS = reinterpret_cast<uint16_t*>(qang + nq);
    uint16_t* entIdx = giS + KK;
    int16_t* link = reinterpret_cast<int16_t*>(entIdx + nq);   // two-grid local map: partner keypoint in the other image, -1 = none
    uint8_t* entBin = reinterpret_cast<uint8_t*>(link + (links ? KK : 0));   // (no link / side arrays for single-grid frames)
    uint8_t* taken = entBin + nq;
    uint8_t* koct = taken + KK;
    uint8_t* qblk = koct + KK;  // per query: Observations() > 0
    uint8_t* qside = qblk + nq; // per query: 0 left image, 1 right image after its left part, 2 right image alone
    // the frame's descriptors (32 bytes per keypoint) when they fit next to the rest: then nothing below touches global memory
    uint4* descS = reinterpret_cast<uint4*>((reinterpret_cast<uintptr_t>(qside + (side ? nq : 0)) + 15) & ~(uintptr_t)15);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = b.outCount[image];
    const int n2 = sides == 2 ? b.outCount[image2] : 0;
    const uint8_t* desc = b.desc + (size_t)image * K * 32;
    const uint8_t* desc2 = sides == 2 ? b.desc + (size_t)image2 * K * 32 : nullptr;
    if (stageDesc) {
        const uint4* g = reinterpret_cast<const uint4*>(desc);
        for (int k = tid; k < 2 * n; k += ST) descS[k] = g[k];
        if (sides == 2) {
            const uint4* g2 = reinterpret_cast<const uint4*>(desc2);
            for (int k = tid; k < 2 * n2; k += ST) descS[2 * K + k] = g2[k];
        }
        desc = reinterpret_cast<const uint8_t*>(descS);
        desc2 = desc + (size_t)K * 32;
    }
    const float* uRight = useURight ? b.uRight + (size_t)image * K : nullptr;
    // the frame's grid, keypoint coordinates / levels and right coordinates move to shared memory once: the walks below touch global
    // memory only for the 32-byte descriptors of their hits
    for (int kk = tid; kk < KK; kk += ST) {
        const int sd = kk >= K ? 1 : 0, k = kk - sd * K, img = sd ? image2 : image, cnt = sd ? n2 : n;
        match[kk] = -1;
        taken[kk] = (taken0 && k < cnt) ? (taken0[kk] != 0) : 0;
        if (links) link[kk] = k < cnt ? (int16_t)links[kk] : (int16_t)-1;
        if (k < cnt) {
            const orbx_keypoint kp = b.kps[(size_t)img * K + k];
            kxy[kk] = (b.kpsUn && sides == 1) ? b.kpsUn[(size_t)img * K + k] : make_float2(kp.x, kp.y);   // mvKeysUn for single-grid frames, mvKeys / mvKeysRight for two-grid ones
            koct[kk] = (uint8_t)(kp.octave & 15);
            urS[kk] = uRight ? uRight[k] : -1.f;
            kang[kk] = kp.angle;
            giS[kk] = (uint16_t)b.gridItems[(size_t)img * K + k];
        }
    }
    for (int c = tid; c < sides * (NC + 1); c += ST) {
        const int sd = c > NC ? 1 : 0;
        gsS[c] = b.gridStart[(size_t)(sd ? image2 : image) * (NC + 1) + (c - sd * (NC + 1))];
    }
    for (int i = tid; i < nq; i += ST) { qang[i] = queries[i].angle; qblk[i] = queries[i].blocks != 0; if (side) qside[i] = side[i]; }
    if (tid < HISTO) hist[tid] = 0;
    if (tid < 4) sInfo[tid] = 0;
    __syncthreads();
    // ---- phase 1a: candidate counts ----
    for (int i = wid; i < nq; i += ST / 32) {
        const orbx_proj_query Q = queries[i];
        const int sd = (side && qside[i]) ? 1 : 0;
        const int c = walk<false>(Q, window_of(Q.u, Q.v, Q.radius, p), gsS + sd * (NC + 1), giS + sd * K, kxy, koct, sd ? desc2 : desc, uRight ? urS : nullptr, nullptr, lane, sd * K);
        if (lane == 0) candStart[i] = c;
    }
    __syncthreads();
    // exclusive scan of the counts (chunks of ST with a running carry)
    int carry = 0;
    for (int base = 0; base < nq; base += ST) {
        const int i = base + tid;
        const int v = i < nq ? candStart[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warpSum[wid] = incl;
        __syncthreads();
        int off = carry;
        for (int w = 0; w < wid; ++w) off += warpSum[w];
        if (i < nq) candStart[i] = off + incl - v;
        int tot = 0;
        for (int w = 0; w < ST / 32; ++w) tot += warpSum[w];
        carry += tot;
        __syncthreads();
    }
    if (tid == 0) { candStart[nq] = carry; sInfo[2] = carry; }
    __syncthreads();
    const int total = carry;
    // Queries are processed in chunks [i0, i1) whose candidates fit the shared-memory list (one chunk unless the windows are
    // very wide, e.g. th = 15 after tracking was lost); the taken / match state carries over, so chunking does not change the result.
    int nm = 0, ne = 0;  // (live in warp 0)
    bool skipRight = false;  // (warp 0) the left part of this map point hit a `continue` that also skips its right part
    for (int i0 = 0; i0 < nq;) {
        if (tid == 0) {
            int i1 = i0;
            const int base = candStart[i0];
            while (i1 < nq && candStart[i1 + 1] - base <= candCap) ++i1;
            sInfo[2] = i1;
        }
        __syncthreads();
        const int i1 = sInfo[2], cbase = candStart[i0];
        if (i1 == i0) {  // a single window holds more candidates than the list: the caller fails with ORBX_ERR_CAPACITY
            if (tid == 0) { outInfo[0] = -1; outInfo[1] = candStart[i0 + 1] - cbase; }
            return;
        }
        // ---- phase 1b: the candidates, in visiting order, with their distances ----
        for (int i = i0 + wid; i < i1; i += ST / 32) {
            const orbx_proj_query Q = queries[i];
            const int sd = (side && qside[i]) ? 1 : 0;
            walk<true>(Q, window_of(Q.u, Q.v, Q.radius, p), gsS + sd * (NC + 1), giS + sd * K, kxy, koct, sd ? desc2 : desc, uRight ? urS : nullptr, cand + (candStart[i] - cbase), lane, sd * K);
        }
        __syncthreads();
        // ---- phase 2: the reference's loop over the map points, in order (one warp) ----
        if (wid == 0) {
            for (int i = i0; i < i1; ++i) {
                const int sd = side ? qside[i] : 0;
                if (sd == 0) skipRight = false;
                else if (sd == 1 && skipRight) continue;
                const int s = candStart[i] - cbase, e = candStart[i + 1] - cbase;
                if (s == e) {
                    if (!LOCAL && sd == 0) skipRight = true;   // if(vIndices2.empty()) continue;  (ORBmatcher.cc:1744-1745)
                    continue;
                }
                if (LOCAL) {
                    // ORBmatcher.cc:85-137: the reference's update order decides who is "second best", so every lane replays the short
                    // candidate list as written (warp-uniform, shared-memory broadcasts, no shuffles)
                    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
                    for (int c = s; c < e; ++c) {
                        const uint32_t en = cand[c];
                        const int idx = (int)(en & 0xFFFFu);
                        if (taken[idx]) continue;
                        const int dist = (int)(en >> 20), level = (int)((en >> 16) & 15u);
                        if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = level; bestIdx = idx; }
                        else if (dist < bestDist2) { bestLevel2 = level; bestDist2 = dist; }
                    }
                    if (bestDist <= thHigh) {
                        const float lim = __fmul_rn(nnRatio, (float)bestDist2);
                        const bool ratioFail = bestLevel == bestLevel2 && (float)bestDist > lim;
                        if (ratioFail) {
                            if (sd == 0) skipRight = true;       // the `continue` of ORBmatcher.cc:133-134 leaves the loop body
                        } else if (bestLevel != bestLevel2 || (float)bestDist <= lim) {
                            const int partner = links ? link[bestIdx] : -1;   // mvLeftToRightMatch / mvRightToLeftMatch (two-grid frames), -1 = none
                            if (lane == 0) {
                                match[bestIdx] = i; taken[bestIdx] = qblk[i];
                                if (partner >= 0) { const int pk = (sd == 0 ? K : 0) + partner; match[pk] = i; taken[pk] = qblk[i]; }
                            }
                            nm += partner >= 0 ? 2 : 1;
                            __syncwarp();
                        }
                    }
                    continue;
                }
                uint32_t best = 0xFFFFFFFFu;  // (distance << 16 | position): the first of the smallest distances
                for (int c0 = s; c0 < e; c0 += 32) {
                    const int c = c0 + lane;
                    if (c < e) {
                        const uint32_t en = cand[c];
                        const int idx = (int)(en & 0xFFFFu);
                        if (!taken[idx]) best = min(best, ((en >> 20) << 16) | (uint32_t)(c - s));
                    }
                }
    #pragma unroll
                for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));
                if (best != 0xFFFFFFFFu && (int)(best >> 16) <= thHigh) {
                    const int idx = (int)(cand[s + (int)(best & 0xFFFFu)] & 0xFFFFu);
                    if (lane == 0) {
                        match[idx] = i;
                        taken[idx] = qblk[i];
                        if (checkOrientation) {
                            float rot = __fsub_rn(qang[i], kang[idx]);  // ORBmatcher.cc:1793-1798
                            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                            int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO));
                            if (bin == HISTO) bin = 0;
                            entIdx[ne] = (uint16_t)idx;
                            entBin[ne] = (uint8_t)bin;
                            ++hist[bin];
                        }
                    }
                    ++nm;
                    ++ne;
                    __syncwarp();
                }
            }
        }
        __syncthreads();
        i0 = i1;
    }
    if (wid == 0 && lane == 0) { sInfo[0] = nm; sInfo[1] = checkOrientation ? ne : 0; }
    __syncthreads();
    // ---- phase 3: rotation consistency ----
    if (!LOCAL && checkOrientation) {
        if (tid == 0) {  // ORBmatcher::ComputeThreeMaxima
            int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
            for (int i = 0; i < HISTO; ++i) {
                const int sz = hist[i];
                if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
                else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
                else if (sz > max3) { max3 = sz; ind3 = i; }
            }
            if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
            else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
            int keep = 0;
            if (ind1 >= 0) keep |= 1 << ind1;
            if (ind2 >= 0) keep |= 1 << ind2;
            if (ind3 >= 0) keep |= 1 << ind3;
            sInfo[3] = keep;
        }
        __syncthreads();
        const int keep = sInfo[3], ne = sInfo[1];
        for (int k = tid; k < ne; k += ST)
            if (!((keep >> entBin[k]) & 1)) {
                match[entIdx[k]] = -1;
                atomicSub(&sInfo[0], 1);
            }
        __syncthreads();
    }
    for (int kk = tid; kk < KK; kk += ST) outMatch[kk] = (kk < K ? kk < n : kk - K < n2) ? match[kk] : -1;
    if (tid == 0) { outInfo[0] = sInfo[0]; outInfo[1] = total; }
}

size_t search_smem_bytes(int K, int nq, int candCap, int stageDesc, int sides, int hasLinks) {
    const size_t KK = (size_t)K * sides;
    return (stageDesc ? KK * 32 + 16 : 0) + KK * 4 + ((size_t)nq + 1) * 4 + (size_t)candCap * 4 + (size_t)sides * (NC + 1) * 4 + 8 + KK * (8 + 4 + 4 + 2 + 1 + 1) +
           (hasLinks ? KK * 2 : 0) + (size_t)nq * (4 + 2 + 1 + 1 + (sides == 2 ? 1 : 0)) + 32;
}

cudaError_t search_configure() {
    cudaError_t e = raise_dynamic_smem(k_search_by_projection<false>, kSearchSmemBudget);
    if (e != cudaSuccess) return e;
    return raise_dynamic_smem(k_search_by_projection<true>, kSearchSmemBudget);
}

void launch_search_by_projection(const Buffers& b, int image, int K, const GridParams& p, const orbx_proj_query* queries, int nq, int useURight, int checkOrientation,
                                 int thHigh, const uint8_t* taken0, int candCap, int stageDesc, int* outMatch, int* outInfo, cudaStream_t s, int image2, const uint8_t* side) {
    k_search_by_projection<false><<<1, ST, search_smem_bytes(K, nq, candCap, stageDesc, image2 >= 0 ? 2 : 1, 0), s>>>(b, image, K, p, queries, nq, useURight, checkOrientation,
                                                                                                                  thHigh, 0.f, taken0, candCap, stageDesc, outMatch, outInfo,
                                                                                                                  image2, side, nullptr);
}

void launch_search_local_points(const Buffers& b, int image, int K, const GridParams& p, const orbx_proj_query* queries, int nq, int useURight, int thHigh, float nnRatio,
                                const uint8_t* taken0, int candCap, int stageDesc, int* outMatch, int* outInfo, cudaStream_t s, int image2, const uint8_t* side,
                                const int* links) {
    k_search_by_projection<true><<<1, ST, search_smem_bytes(K, nq, candCap, stageDesc, image2 >= 0 ? 2 : 1, links ? 1 : 0), s>>>(b, image, K, p, queries, nq, useURight, 0, thHigh, nnRatio,
                                                                                                                 taken0, candCap, stageDesc, outMatch, outInfo, image2, side,
                                                                                                                 links);
}

}  // namespace orbx
