// lapmatch_kernel.cu — descriptor stage of Frame::ComputeStereoFishEyeMatches on the device.
//
// Replaces slam_framework/src/Frame.cc:1162-1187: BFmatcher.knnMatch(stereoDescLeft, stereoDescRight, matches, 2) with
// BFmatcher = cv::BFMatcher(cv::NORM_HAMMING) (Frame.cc:52) over the two lapping-area tails — rows [monoLeft, Nleft) of the
// left and [monoRight, Nright) of the right descriptors, exactly the partition operator() leaves behind
// (ORBextractor.cc:1126-1135) — and Lowe's gate `m[0].distance < m[1].distance * 0.7` (Frame.cc:1187).  What follows in
// the reference (KannalaBrandt8::TriangulateMatches per accepted pair, an Eigen SVD on a handful of matches) stays on the
// host: it consumes {queryIdx, trainIdx} and is not data parallel.
//
// cv::batchDistance with K = 2 scans the train rows in index order and lets a row into the K-best list only when its
// distance is strictly smaller than the current K-th, behind equal distances: the result is the two smallest
// (distance, trainIdx) pairs in lexicographic order.  A 32-bit key distance << 16 | trainIdx makes that a plain min.
//
// One CTA per stereo frame; the train descriptors sit in shared memory word-transposed ([word][train]: conflict-free), in
// tiles of LM_TILE rows; one warp per query, lanes stride the train rows (8 x XOR + POPC), per-lane best two keys, warp
// merge of the sorted pairs by shuffles.  Both descriptor sets are already resident; only 17 bytes per query leave the GPU.
#include <algorithm>

#include "orbx_internal.cuh"

namespace orbx {

namespace {

constexpr int LM_THREADS = 256, LM_WARPS = 8;
constexpr int LM_TILE = 2048;  // train rows per shared-memory tile (64 KB)
constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr uint32_t NONE = 0xFFFFFFFFu;

}  // namespace

__global__ void __launch_bounds__(LM_THREADS) k_lap_match(const __grid_constant__ Geom g, const Buffers bl, const Buffers br, int lMul, int lAdd, int rMul,
                                                         int rAdd) {
    extern __shared__ __align__(16) uint32_t tr[];  // [8][LM_TILE]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int imgL = bl.imgBase + lMul * blockIdx.x + lAdd, imgR = br.imgBase + rMul * blockIdx.x + rAdd;
    const int K = g.K;
    const int NL = min(bl.outCount[imgL], K), NR = min(br.outCount[imgR], K);
    const int monoL = min(max(bl.outMono[imgL], 0), NL), monoR = min(max(br.outMono[imgR], 0), NR);
    const int nq = NL - monoL, nt = NR - monoR;
    const uint32_t* dL = reinterpret_cast<const uint32_t*>(bl.desc + ((size_t)imgL * K + monoL) * 32);
    const uint32_t* dR = reinterpret_cast<const uint32_t*>(br.desc + ((size_t)imgR * K + monoR) * 32);
    int4* out = bl.lapKnn + (size_t)imgL * K;
    uint8_t* ok = bl.lapOk + (size_t)imgL * K;

    for (int t0 = 0; t0 < max(nt, 1); t0 += LM_TILE) {
        const int tn = min(nt - t0, LM_TILE);
        __syncthreads();  // the previous tile is no longer read
        for (int k = tid; k < tn * 8; k += LM_THREADS) tr[(k & 7) * LM_TILE + (k >> 3)] = __ldg(dR + (size_t)t0 * 8 + k);
        __syncthreads();
        for (int q = wid; q < nq; q += LM_WARPS) {
            uint32_t qd[8];
            {
                const uint4 a = __ldg(reinterpret_cast<const uint4*>(dL + (size_t)q * 8)), b = __ldg(reinterpret_cast<const uint4*>(dL + (size_t)q * 8) + 1);
                qd[0] = a.x; qd[1] = a.y; qd[2] = a.z; qd[3] = a.w; qd[4] = b.x; qd[5] = b.y; qd[6] = b.z; qd[7] = b.w;
            }
            uint32_t k0 = NONE, k1 = NONE;
            for (int j = lane; j < tn; j += 32) {
                int d = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) d += __popc(qd[w] ^ tr[w * LM_TILE + j]);
                const uint32_t key = ((uint32_t)d << 16) | (uint32_t)(t0 + j);
                k1 = min(k1, max(k0, key));
                k0 = min(k0, key);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const uint32_t b0 = __shfl_xor_sync(FULL, k0, o), b1 = __shfl_xor_sync(FULL, k1, o);
                k1 = min(max(k0, b0), min(k1, b1));
                k0 = min(k0, b0);
            }
            if (lane == 0) {
                if (t0 > 0) {  // merge with the best two of the earlier tiles
                    const int4 p = out[q];
                    const uint32_t p0 = p.x >= 0 ? ((uint32_t)p.y << 16) | (uint32_t)p.x : NONE, p1 = p.z >= 0 ? ((uint32_t)p.w << 16) | (uint32_t)p.z : NONE;
                    k1 = min(max(k0, p0), min(k1, p1));
                    k0 = min(k0, p0);
                }
                int4 r;
                r.x = k0 != NONE ? (int)(k0 & 0xFFFF) : -1; r.y = k0 != NONE ? (int)(k0 >> 16) : -1;
                r.z = k1 != NONE ? (int)(k1 & 0xFFFF) : -1; r.w = k1 != NONE ? (int)(k1 >> 16) : -1;
                out[q] = r;
                // (float)d0 < (float)d1 * 0.7 with the product in double (Frame.cc:1187)
                ok[q] = (k1 != NONE && (double)r.y < __dmul_rn((double)r.w, 0.7)) ? 1 : 0;
            }
        }
    }
}

size_t lap_match_smem_bytes() { return (size_t)8 * LM_TILE * 4; }

cudaError_t lap_match_configure() { return raise_dynamic_smem(k_lap_match, lap_match_smem_bytes()); }

void launch_lap_match(const Geom& g, const Buffers& bl, const Buffers& br, int nFrames, int lMul, int lAdd, int rMul, int rAdd, cudaStream_t s) {
    k_lap_match<<<nFrames, LM_THREADS, lap_match_smem_bytes(), s>>>(g, bl, br, lMul, lAdd, rMul, rAdd);
}

}  // namespace orbx
