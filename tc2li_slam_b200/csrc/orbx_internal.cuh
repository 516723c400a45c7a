// orbx_internal.cuh — shared declarations of the sm_100a ORB front-end (host orchestration + kernels).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <utility>

#include "../../include/orbx.h"
#include "tma.cuh"

namespace orbx {

constexpr int kEdge = 19;      // EDGE_THRESHOLD (ORBextractor.cc:47)
constexpr int kBorder = 16;    // minBorderX/Y = EDGE_THRESHOLD - 3 (ORBextractor.cc:763)
constexpr int kHalfPatch = 15; // HALF_PATCH_SIZE (ORBextractor.cc:46)
constexpr int kMarginX = 4;    // padding bytes in front of every plane row (keeps rows word-aligned; the blur mirrors pixels 3..1 into its box copy of them)
constexpr int kMarginY = 3;    // rows of margin above and below every plane
constexpr int kMaxCell = 72;   // wCell/hCell < 70 for any level with >= 1 cell of 35 px (ORBextractor.cc:771-777)

// Per-level geometry, passed by value to every kernel (fits well inside the 4 KB parameter space).
struct LevelGeom {
    int w, h, pitch;                  // level size and row pitch in bytes (pitch % 128 == 0)
    int nCols, nRows, wCell, hCell;   // FAST cell grid (ORBextractor.cc:771-777)
    int quota;                        // mnFeaturesPerLevel[level]
    int nIni;                         // initial quadtree nodes (ORBextractor.cc:533)
    float hX;                         // root node width (ORBextractor.cc:535)
    int candCap;                      // candidate capacity per image (NMS keeps <= 1 pixel per 2x2 block of a cell interior)
    int nodeCap;                      // quadtree node-slot capacity
    int selCap, selOff;               // selected keypoints: capacity, offset inside the per-image array
    int cellOff;                      // offset inside the per-image cell array
    float sf;                         // mvScaleFactor[level]
    float kpSize;                     // (float)(int)(31 * sf)
    unsigned long long planeOff;      // byte offset of pixel (0,0) of image 0 of this level in the pyramid / blur buffers
    unsigned long long planeBytes;    // image-to-image stride: pitch * (h + 2*kMarginY); every plane carries a margin (see k_margin)
    unsigned long long candOff;       // element offset of this level's block in the candidate buffers
};

struct Geom {
    int nlevels;
    int K;               // per-image keypoint capacity = sum selCap
    int cellsPerImage;   // sum nCols*nRows
    int iniTh, minTh;
    int maxImages;
    LevelGeom lv[ORBX_MAX_LEVELS];
};

// resize tables of one level (device pointers): per destination column / row
struct ResizeTab {
    const short* xofs;  // [w_dst]
    const short* xa;    // [w_dst][2]
    const short* yofs;  // [h_dst]
    const short* yb;    // [h_dst][2]
};

// candidate packing: x' = x - 16 (12 bits) | y' = y - 16 (12 bits) | response (8 bits)
__host__ __device__ inline uint32_t pack_cand(int xr, int yr, int resp) { return (uint32_t)xr | ((uint32_t)yr << 12) | ((uint32_t)resp << 24); }
__host__ __device__ inline int cand_x(uint32_t c) { return (int)(c & 0xFFFu); }
__host__ __device__ inline int cand_y(uint32_t c) { return (int)((c >> 12) & 0xFFFu); }
__host__ __device__ inline int cand_r(uint32_t c) { return (int)(c >> 24); }

// Position of a candidate in the reference's candidate vector order: cell row, cell column, then raster inside the
// cell (ORBextractor.cc:779-846).  Smaller key = earlier in vToDistributeKeys; only its ORDER matters (tie-break of
// the per-node maximum, ORBextractor.cc:737-747).  40 bits.
__host__ __device__ inline unsigned long long cand_order_key(uint32_t c, int wCell, int hCell) {
    const int x = cand_x(c), y = cand_y(c);
    const unsigned long long ci = (unsigned long long)((y - 3) / hCell), cj = (unsigned long long)((x - 3) / wCell);
    return (ci << 32) | (cj << 24) | ((unsigned long long)y << 12) | (unsigned long long)x;
}

// work item of the orient+describe kernel
struct KpWork {
    short x, y;      // level coordinates
    short level;
    short dst;       // output slot (keypoints[dst], descriptors[dst]); -1 = none
};

struct Buffers {
    uint8_t* pyr;           // all levels, all images (level-major: [level][image][h][pitch])
    uint8_t* blur;          // same layout
    uint32_t* cand;         // [level][image][candCap]
    uint16_t* pnode;        // same layout: quadtree node slot of each candidate
    int* candCount;         // [image][ORBX_MAX_LEVELS]
    uint8_t* cellFlags;     // [image][cellsPerImage]: 0 skipped, 1 had a corner above iniThFAST, 2 fell back to minThFAST
    uint32_t* sel;          // [image][K] packed selected keypoints (level coordinates - 16), final list order
    int* selCount;          // [image][ORBX_MAX_LEVELS]
    int* lapping;           // [image][2]
    orbx_keypoint* kps;     // [image][K]
    uint8_t* desc;          // [image][K][32]
    KpWork* work;           // [image][K]
    int* outCount;          // [image]
    int* outMono;           // [image]
    float* uRight;          // [image][K] stereo matches of LEFT images: right-image x of the match, -1 = none (Frame::mvuRight)
    float* depth;           // [image][K] depth of the match, -1 = none (Frame::mvDepth)
    int4* lapKnn;           // [image][K] lapping-tail kNN of LEFT images: {trainIdx0, distance0, trainIdx1, distance1}, -1 = absent
    uint8_t* lapOk;         // [image][K] Lowe ratio gate of that pair (Frame.cc:1187)
    int* gridStart;         // [image][64 * 48 + 1] CSR of Frame::mGrid (cell = column * 48 + row)
    int* gridItems;         // [image][K] keypoint indices, cell by cell, ascending inside a cell
    const float2* kpsUn;    // [image][K] undistorted keypoint coordinates (Frame::mvKeysUn), or NULL: the grid and the single-grid searches then read kps
    int* errFlag;           // device-side error flag (capacity overflow etc.)
    const uint4* cellTab;   // [cellsPerImage] FAST cell descriptors (make_cell_table)
    int imgBase;            // first image of this launch set (kernels index images as imgBase + blockIdx)
};

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to (kernel, device), not to a handle: a second handle with smaller parameters must
// not lower the limit a live, larger handle launches with.  Every *_configure goes through this: the limit is only ever raised.
template <typename Kernel>
inline cudaError_t raise_dynamic_smem(Kernel kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> limit;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    size_t& cur = limit[std::make_pair(reinterpret_cast<const void*>(kernel), dev)];
    const size_t want = std::max<size_t>(bytes, 48 * 1024);
    if (want <= cur) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want);
    if (e == cudaSuccess) cur = want;
    return e;
}

// kernel launchers (one translation unit per stage)
void launch_synth(uint8_t* l0, int w, int h, int pitch, size_t planeBytes, uint32_t seed0, int firstImage, int nImages, int stereo, cudaStream_t s);

struct FastMaps {  // per level: {tile row bytes x (hCell + 6)} boxes of the pyramid (FAST cell crops)
    CUtensorMap m[ORBX_MAX_LEVELS];
};
void launch_fast(const Geom& g, const Buffers& b, const FastMaps& maps, int nImages, int tileRows, int ts, int numSMs, cudaStream_t s);
void make_cell_table(const Geom& g, uint4* tab, int* tileRows, int* maxCw);
int fast_tile_stride(int maxCw);
cudaError_t fast_configure(int tileRows, int ts);
int fast_encode_maps(const Geom& g, uint8_t* pyr, int nImages, int ts, FastMaps* out);
struct LevelMaps {  // one tensor map per pyramid level
    CUtensorMap m[ORBX_MAX_LEVELS];
};
void launch_blur(const Geom& g, const Buffers& b, const LevelMaps& maps, int level, int nImages, const int taps[7], cudaStream_t s);
void launch_blur_all(const Geom& g, const Buffers& b, const LevelMaps& maps, int nImages, const int taps[7], int segRows, cudaStream_t s);
int blur_encode_maps(const Geom& g, uint8_t* pyr, int nImages, LevelMaps* out);
void launch_pyramid_level(const Geom& g, const Buffers& b, const LevelMaps& maps, int level, const ResizeTab& tab, int nImages, cudaStream_t s);
int resize_encode_maps(const Geom& g, uint8_t* pyr, int nImages, LevelMaps* out);
cudaError_t launch_octree(const Geom& g, const Buffers& b, int nImages, int numSMs, cudaStream_t s);
int octree_launch_count(const Geom& g, int nImages, int numSMs);
void launch_assemble(const Geom& g, const Buffers& b, int nImages, cudaStream_t s);
#ifndef DESC_WINDOW_ROW
#define DESC_WINDOW_ROW 80
#endif
constexpr int kDescWindowRow = DESC_WINDOW_ROW;  // bytes per row of the staged rBRIEF window (TMA box width, multiple of 16, >= 52)
struct DescMaps {  // per level: 48x31 boxes of the pyramid (IC_Angle patch), kDescWindowRow x 37 boxes of the blurred pyramid (rBRIEF window)
    CUtensorMap ic[ORBX_MAX_LEVELS];
    CUtensorMap bl[ORBX_MAX_LEVELS];
};
void launch_describe(const Geom& g, const Buffers& b, const DescMaps& maps, int nImages, cudaStream_t s);
cudaError_t describe_configure();
void launch_margin(const Geom& g, const Buffers& b, int nImages, cudaStream_t s);
void launch_remap(const uint8_t* src, int sw, int sh, int channels, int rgb, const float* mapx, const float* mapy, uint8_t* l0, int w, int h, int pitch,
                  size_t planeBytes, int nImages, cudaStream_t s);
void launch_input_resize(const uint8_t* src, int sw, int sh, int channels, int rgb, const ResizeTab& tab, uint8_t* l0, int w, int h, int pitch, size_t planeBytes,
                         int nImages, cudaStream_t s);
void launch_gray(const uint8_t* src, int channels, int rgb, uint8_t* l0, int w, int h, int pitch, size_t planeBytes, int nImages, cudaStream_t s);
void launch_border(const uint8_t* src, int w, int h, int pitch, uint8_t* dst, int dpitch, cudaStream_t s);
size_t octree_smem_bytes(int nodeCap);
cudaError_t upload_describe_constants(const signed char pattern[1024], const int umax[16]);

size_t stereo_smem_bytes(int K, int H);
cudaError_t stereo_configure(int K, int H);
void launch_stereo_match(const Geom& g, const Buffers& bl, const Buffers& br, int nFrames, int lMul, int lAdd, int rMul, int rAdd, float mb, float mbf,
                         cudaStream_t s);

size_t lap_match_smem_bytes();
cudaError_t lap_match_configure();
void launch_lap_match(const Geom& g, const Buffers& bl, const Buffers& br, int nFrames, int lMul, int lAdd, int rMul, int rAdd, cudaStream_t s);

struct GridParams {  // Frame::mnMinX / mnMinY and mfGridElementWidthInv / HeightInv (Frame.cc:179-180)
    float minX, minY, invW, invH;
};
size_t grid_smem_bytes(int K);
cudaError_t grid_configure(int K);
struct UndistortParams {   // cv::undistortPoints(src, dst, K, distCoeffs, noArray(), P) as Frame::UndistortKeyPoints calls it (Frame.cc:777-810)
    double cx, cy, ifx, ify, k[12], nfx, nfy, ncx, ncy;
};
void launch_undistort(const Buffers& b, int K, const UndistortParams& p, float2* out, int nImages, cudaStream_t s);
void launch_assign_grid(const Buffers& b, int K, const GridParams& p, int nImages, cudaStream_t s);
void launch_features_in_area(const Buffers& b, int K, const GridParams& p, const orbx_area_query* queries, int nQueries, int cap, int* counts, int* out,
                             cudaStream_t s);

constexpr size_t kSearchSmemBudget = 224 * 1024;   // dynamic shared memory of the search kernel (B200: 227 KB per CTA minus its static arrays)
size_t search_smem_bytes(int K, int nq, int candCap, int stageDesc, int sides, int hasLinks);
cudaError_t search_configure();
// image2 >= 0: a two-grid frame (left image `image`, right image `image2`); side[nq] (device) tells the sub-queries apart; taken0 / links / outMatch then hold
// 2 * K entries (left image first); links = partner keypoint in the other image (mvLeftToRightMatch | mvRightToLeftMatch), local-map search only
void launch_search_by_projection(const Buffers& b, int image, int K, const GridParams& p, const orbx_proj_query* queries, int nq, int useURight, int checkOrientation,
                                 int thHigh, const uint8_t* taken0, int candCap, int stageDesc, int* outMatch, int* outInfo, cudaStream_t s, int image2 = -1,
                                 const uint8_t* side = nullptr);

void launch_search_local_points(const Buffers& b, int image, int K, const GridParams& p, const orbx_proj_query* queries, int nq, int useURight, int thHigh, float nnRatio,
                                const uint8_t* taken0, int candCap, int stageDesc, int* outMatch, int* outInfo, cudaStream_t s, int image2 = -1,
                                const uint8_t* side = nullptr, const int* links = nullptr);

// DBoW2 vocabulary transform (bow_kernel.cu): the tree by node id (0 = root); children of node n = childList[childStart[n] .. childStart[n + 1])
struct BowVocDev {
    int k, L, scoring, weighting;
    const uint4* desc;        // [nodes][2]
    const int* childStart;    // [nodes + 1]
    const int* childList;     // [nodes - 1]
    const double* weight;     // [nodes]
    const int* wordId;        // [nodes]
};
struct BowOut {               // per image: capacity K entries each (nodeStart: K + 1)
    uint32_t* words; double* values; int* nWords;
    uint32_t* nodes; int* nodeStart; uint32_t* features; int* nNodes;
};
int bow_sort_size(int K);
cudaError_t bow_configure(int K);
void launch_bow(const BowVocDev& v, const uint8_t* desc, const int* counts, int fixedCount, int K, int nImages, int levelsup, uint32_t* leaf, uint32_t* nid,
                const BowOut& out, cudaStream_t s);
const BowVocDev& vocabulary_device(const orbx_vocabulary* v);
int vocabulary_device_ordinal(const orbx_vocabulary* v);
// error reporting shared by the translation units that hold C-ABI entry points (thread-local message behind orbx_last_error)
int set_error(int code, const std::string& msg);
int set_cuda_error(cudaError_t e, const char* what);

// debug / test hooks
void launch_debug_sincos(const float* anglesDeg, float* outSin, float* outCos, int n, cudaStream_t s);
void launch_debug_atan2(const float* y, const float* x, float* out, int n, cudaStream_t s);

}  // namespace orbx
