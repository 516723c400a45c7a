// orbx_host.cu — host side of the C ABI (include/orbx.h): geometry, device buffers, launch sequence, copies.
//
// Mirrors, on the host, what TC2LI_SLAM::ORBextractor computes once per instance (ORBextractor.cc:383-443: scale tables,
// per-level quotas, umax) and per call outside the pixel loops (level sizes ORBextractor.cc:1147-1148, cell grid
// :759-777, quadtree roots :533-535), and OpenCV's bilinear-resize coefficient tables.  Everything per
// pixel / per candidate / per keypoint runs in the CUDA kernels of this directory.  There is no CPU fallback: without a
// usable device every entry point fails with ORBX_ERR_CUDA.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "orbx_internal.cuh"
#include "stdsort_emul.cuh"
#include "tma.cuh"

namespace orbx {
void launch_debug_sort_impl(unsigned long long* buf, int n, int forceDepth, cudaStream_t s);
cudaError_t octree_configure(int maxNodeCap);
}

using namespace orbx;

namespace orbx {
// cuTensorMapEncodeTiled through the runtime's driver entry point (liborbx links cudart statically and not libcuda)
int tma_encode_3d(CUtensorMap* out, void* base, uint64_t pitch, uint64_t rows, uint64_t images, uint64_t imageStride, uint32_t boxBytes,
                  uint32_t boxRows) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return -1;
        fn = (EncodeFn)p;
    }
    const cuuint64_t gdim[3] = {pitch, rows, images};
    const cuuint64_t gstride[2] = {pitch, imageStride};
    const cuuint32_t box[3] = {boxBytes, boxRows, 1};
    const cuuint32_t estride[3] = {1, 1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}
}  // namespace orbx

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return ORBX_ERR_CUDA;
}
#define CU(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return cuda_fail(e__, #x); } while (0)
}  // namespace
namespace orbx {
int set_error(int code, const std::string& msg) { return fail(code, msg); }
int set_cuda_error(cudaError_t e, const char* what) { return cuda_fail(e, what); }
}  // namespace orbx
namespace {

static const signed char kPatternHost[1024] = {
#include "orb_pattern.inc"
};

inline int cv_round(float v) { return (int)lrintf(v); }    // cvRound: round half to even
inline int cv_round_d(double v) { return (int)lrint(v); }
inline short sat_short(int v) { return (short)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

struct Params {
    int nfeatures, nlevels, iniTh, minTh;
    float scaleFactorF;
    std::vector<float> sf, isf, sig2, isig2;
    std::vector<int> quota;
    int umax[16];
};

// ORBextractor::ORBextractor (ORBextractor.cc:383-443)
int make_params(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, Params& p) {
    if (nlevels < 1 || nlevels > ORBX_MAX_LEVELS) return fail(ORBX_ERR_ARG, "nlevels must be in [1, ORBX_MAX_LEVELS]");
    if (nfeatures < 1 || nfeatures > 30000) return fail(ORBX_ERR_ARG, "nfeatures must be in [1, 30000]");
    if (!(scaleFactor > 1.0f) || scaleFactor > 4.f) return fail(ORBX_ERR_ARG, "scaleFactor must be in (1, 4]");
    if (minTh < 1 || iniTh < minTh || iniTh > 254) return fail(ORBX_ERR_ARG, "need 1 <= minThFAST <= iniThFAST <= 254");
    p.nfeatures = nfeatures; p.nlevels = nlevels; p.iniTh = iniTh; p.minTh = minTh; p.scaleFactorF = scaleFactor;
    const double sfd = (double)scaleFactor;  // the member is a double initialised from the float argument (ORBextractor.h:105)
    p.sf.assign(nlevels, 1.f); p.sig2.assign(nlevels, 1.f); p.isf.assign(nlevels, 1.f); p.isig2.assign(nlevels, 1.f);
    for (int i = 1; i < nlevels; ++i) {
        p.sf[i] = (float)((double)p.sf[i - 1] * sfd);
        p.sig2[i] = p.sf[i] * p.sf[i];
    }
    for (int i = 0; i < nlevels; ++i) { p.isf[i] = 1.0f / p.sf[i]; p.isig2[i] = 1.0f / p.sig2[i]; }
    p.quota.assign(nlevels, 0);
    const float factor = (float)(1.0 / sfd);  // 1.0f / double
    float nDesired = (float)nfeatures * (1.f - factor) / (1.f - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        p.quota[l] = cv_round(nDesired);
        sum += p.quota[l];
        nDesired *= factor;
    }
    p.quota[nlevels - 1] = std::max(nfeatures - sum, 0);
    // circular patch: umax[v] = half width of row v (ORBextractor.cc:427-442)
    int um[kHalfPatch + 2] = {0};
    const int vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    const int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = (double)kHalfPatch * kHalfPatch;
    for (int v = 0; v <= vmax; ++v) um[v] = cv_round_d(std::sqrt(hp2 - (double)v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (um[v0] == um[v0 + 1]) ++v0;
        um[v] = v0;
        ++v0;
    }
    for (int v = 0; v < 16; ++v) p.umax[v] = um[v];
    return ORBX_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
struct orbx_batch {
    Params p;
    Geom g;
    Buffers buf = {};
    int device = 0, width = 0, height = 0, maxImages = 0;
    int taps[7] = {18, 34, 48, 56, 48, 34, 18};
    unsigned tapsGeneration = 0;        // bumped by every change of the taps (orbx_extractor re-captures its graph)
    cudaStream_t own = nullptr, stream = nullptr;
    cudaStream_t side = nullptr;        // blur runs here, concurrently with FAST + quadtree (both only need the pyramid)
    cudaEvent_t evPyr = nullptr, evBlur = nullptr;
    int sched = 0;
    // feature grid (orbx_batch_assign_grid): buffers are allocated on first use
    GridParams gridParams = {0.f, 0.f, 0.f, 0.f};
    int gridImages = 0;                 // images the grids were last built for (0 = never)
    orbx_area_query* qDev = nullptr; int* qCounts = nullptr; int* qOut = nullptr;
    size_t qCap = 0, qOutCap = 0;
    orbx_proj_query* spDev = nullptr; int* spOut = nullptr;   // projection search: queries; match[K] + info[2]
    size_t spCap = 0;
    bool spConfigured = false;
    // Frame::UndistortKeyPoints (orbx_batch_undistort_keypoints): undistorted coordinates of images [0, unImages); 0 = mvKeysUn == mvKeys
    float2* kpsUnBuf = nullptr;
    int unImages = 0;
    bool gridUn = false;   // the feature grid was built from the undistorted coordinates: its readers use them too
    // bag of words (orbx_batch_compute_bow): buffers are allocated on first use
    uint32_t *bowLeaf = nullptr, *bowNid = nullptr;
    BowOut bow = {};
    int bowImages = 0;
    // extra lanes for orbx_batch_run_pipelined: chunks of a resident batch are issued round-robin over the lanes so that the
    // latency-bound stages of one chunk (quadtree) overlap the issue-bound stages of another (FAST)
    static const int kLanes = 3;
    cudaStream_t laneMain[kLanes] = {}, laneSide[kLanes] = {};
    cudaEvent_t lanePyr[kLanes] = {}, laneBlur[kLanes] = {}, laneDone[kLanes] = {}, evFork = nullptr;
    std::vector<void*> allocs;
    ResizeTab tabs[ORBX_MAX_LEVELS];
    size_t pyrBytes = 0, candElems = 0;
    uint8_t* borderScratch = nullptr;
    size_t borderScratchBytes = 0;
    bool timing = false;
    cudaEvent_t ev[ORBX_NUM_STAGES + 1] = {};
    int launches = 0;
    int lastImages = 0;
    std::vector<int> hostLap;
    bool lapStable = false;   // the `lapping` pointer given to run stays valid and pinned (single-image handle): copy from it directly
    int fastTileRows = 0, fastTs = 64, numSMs = 148;
    FastMaps fastMaps;
    LevelMaps blurMaps, resizeMaps;
    DescMaps descMaps;
    float *rectX = nullptr, *rectY = nullptr;  // resident rectification maps [height][width] (System.cc:238-248), null = none
    int srcW = 0, srcH = 0;                    // size of the raw images the maps sample
    bool inResize = false;                     // orbx_batch_set_input_resize: raw images are resized (System.cc:250-252) instead of remapped
    ResizeTab inTab = {nullptr, nullptr, nullptr, nullptr};
    int inTabW = 0, inTabH = 0;                // raw size the tables were built for
    uint8_t* stageIn = nullptr;  // dense device staging of uploaded images (PCIe copies of odd-width rows are ~6x slower than one contiguous copy)
    size_t stageBytes = 0;
};

namespace {

template <typename T>
int dev_alloc(orbx_batch* b, T** out, size_t count) {
    void* p = nullptr;
    CU(cudaMalloc(&p, std::max<size_t>(count * sizeof(T), 256)));
    b->allocs.push_back(p);
    *out = (T*)p;
    return ORBX_OK;
}

// level sizes, FAST grid, quadtree roots, capacities
int make_geom(const Params& p, int W, int H, int maxImages, Geom& g) {
    std::memset(&g, 0, sizeof(g));
    if (W > ORBX_MAX_DIM || H > ORBX_MAX_DIM) return fail(ORBX_ERR_TOO_LARGE, "image larger than ORBX_MAX_DIM");
    g.nlevels = p.nlevels; g.iniTh = p.iniTh; g.minTh = p.minTh; g.maxImages = maxImages;
    unsigned long long planeOff = 0, candOff = 0;
    int selOff = 0, cellOff = 0;
    for (int l = 0; l < p.nlevels; ++l) {
        LevelGeom& L = g.lv[l];
        L.w = cv_round((float)W * p.isf[l]);  // ORBextractor.cc:1147-1148: always from the ORIGINAL size
        L.h = cv_round((float)H * p.isf[l]);
        const int bw = L.w - 2 * kBorder, bh = L.h - 2 * kBorder;  // maxBorder - minBorder
        if (bw < 35 || bh < 35) return fail(ORBX_ERR_TOO_SMALL, "a pyramid level is smaller than one 35-px FAST cell");
        L.pitch = (int)align_up((size_t)L.w + 2 * kMarginX, 128);
        const float width = (float)bw, height = (float)bh;
        L.nCols = (int)(width / 35.f);
        L.nRows = (int)(height / 35.f);
        L.wCell = (int)std::ceil(width / (float)L.nCols);
        L.hCell = (int)std::ceil(height / (float)L.nRows);
        if (L.wCell > kMaxCell || L.hCell > kMaxCell) return fail(ORBX_ERR_TOO_LARGE, "FAST cell larger than the kernel tile");
        L.quota = p.quota[l];
        L.nIni = (int)std::round((float)bw / (float)bh);  // ORBextractor.cc:533
        if (L.nIni < 1) return fail(ORBX_ERR_TOO_SMALL, "portrait level: the reference's quadtree has no root node (nIni == 0)");
        L.hX = (float)bw / (float)L.nIni;
        // strict 3x3 maxima: at most one per 2x2 block of each cell interior
        L.candCap = L.nCols * L.nRows * ((L.wCell + 1) / 2) * ((L.hCell + 1) / 2);
        L.selCap = std::max(L.quota + 3, 4 * L.nIni);
        L.nodeCap = L.selCap + 8;
        if (L.nodeCap > 16383) return fail(ORBX_ERR_TOO_LARGE, "per-level quota too large");
        L.selOff = selOff; selOff += L.selCap;
        L.cellOff = cellOff; cellOff += L.nCols * L.nRows;
        L.sf = p.sf[l];
        L.kpSize = (float)(int)(31 * p.sf[l]);  // PATCH_SIZE*mvScaleFactor[level] truncated (ORBextractor.cc:854)
        L.planeBytes = (unsigned long long)L.pitch * (L.h + 2 * kMarginY);
        L.planeOff = planeOff + (unsigned long long)kMarginY * L.pitch + kMarginX;  // pixel (0,0) of image 0
        planeOff += L.planeBytes * maxImages;
        L.candOff = candOff; candOff += (unsigned long long)L.candCap * maxImages;
    }
    g.K = selOff;
    g.cellsPerImage = cellOff;
    if (g.K > 32767) return fail(ORBX_ERR_TOO_LARGE, "keypoint capacity per image exceeds 32767");
    return ORBX_OK;
}

// cv::resize INTER_LINEAR 8U coefficient tables: 11-bit fixed point, clamped source index on x
void make_resize_tables(int sn, int dn, bool clampIndex, std::vector<short>& ofs, std::vector<short>& coef) {
    ofs.resize(dn); coef.resize(2 * (size_t)dn);
    const double scale = 1.0 / ((double)dn / (double)sn);
    for (int d = 0; d < dn; ++d) {
        float f = (float)(((double)d + 0.5) * scale - 0.5);
        int s = (int)std::floor(f);
        f -= (float)s;
        if (clampIndex) {
            if (s < 0) { f = 0.f; s = 0; }
            if (s >= sn - 1) { f = 0.f; s = sn - 1; }
        }
        ofs[d] = (short)s;
        coef[2 * d] = sat_short(cv_round((1.f - f) * 2048.f));
        coef[2 * d + 1] = sat_short(cv_round(f * 2048.f));
    }
}

int batch_alloc(orbx_batch* b) {
    const Geom& g = b->g;
    const int n = b->maxImages;
    const LevelGeom& last = g.lv[g.nlevels - 1];
    b->pyrBytes = (size_t)(last.planeOff + last.planeBytes * n);  // (a few bytes more than needed: planeOff includes the margin offset)
    b->candElems = (size_t)(last.candOff + (unsigned long long)last.candCap * n);
    Buffers& B = b->buf;
    int r;
    if ((r = dev_alloc(b, &B.pyr, b->pyrBytes))) return r;
    if ((r = dev_alloc(b, &B.blur, b->pyrBytes))) return r;
    if ((r = dev_alloc(b, &B.cand, b->candElems))) return r;
    if ((r = dev_alloc(b, &B.pnode, b->candElems))) return r;
    if ((r = dev_alloc(b, &B.candCount, (size_t)n * ORBX_MAX_LEVELS))) return r;
    if ((r = dev_alloc(b, &B.cellFlags, (size_t)n * g.cellsPerImage))) return r;
    if ((r = dev_alloc(b, &B.sel, (size_t)n * g.K))) return r;
    if ((r = dev_alloc(b, &B.selCount, (size_t)n * ORBX_MAX_LEVELS))) return r;
    if ((r = dev_alloc(b, &B.lapping, (size_t)n * 2))) return r;
    if ((r = dev_alloc(b, &B.kps, (size_t)n * g.K))) return r;
    if ((r = dev_alloc(b, &B.desc, (size_t)n * g.K * 32))) return r;
    if ((r = dev_alloc(b, &B.work, (size_t)n * g.K))) return r;
    if (n == 1) {   // the single-image handle reads {count, monoIndex, error flag} back with ONE copy: keep them adjacent
        if ((r = dev_alloc(b, &B.outCount, (size_t)4))) return r;
        B.outMono = B.outCount + 1;
        B.errFlag = B.outCount + 2;
    } else if (n == 2) {   // the stereo-pair batch: {count[2], monoIndex[2], error flag}
        if ((r = dev_alloc(b, &B.outCount, (size_t)8))) return r;
        B.outMono = B.outCount + 2;
        B.errFlag = B.outCount + 4;
    } else {
        if ((r = dev_alloc(b, &B.outCount, (size_t)n))) return r;
        if ((r = dev_alloc(b, &B.outMono, (size_t)n))) return r;
    }
    if ((r = dev_alloc(b, &B.uRight, (size_t)n * g.K))) return r;
    if ((r = dev_alloc(b, &B.depth, (size_t)n * g.K))) return r;
    if ((r = dev_alloc(b, &B.lapKnn, (size_t)n * g.K))) return r;
    if ((r = dev_alloc(b, &B.lapOk, (size_t)n * g.K))) return r;
    if (n > 2 && (r = dev_alloc(b, &B.errFlag, 1))) return r;
    B.imgBase = 0;
    for (int l = 0; l < g.nlevels; ++l) {
        const LevelGeom& L = g.lv[l];
        const size_t blockOff = (size_t)L.planeOff - (size_t)kMarginY * L.pitch - kMarginX;  // first byte of this level's block
        if (tma_encode_3d(&b->descMaps.ic[l], B.pyr + blockOff, L.pitch, L.h + 2 * kMarginY, n, L.planeBytes, 48, 31) ||
            tma_encode_3d(&b->descMaps.bl[l], B.blur + blockOff, L.pitch, L.h + 2 * kMarginY, n, L.planeBytes, kDescWindowRow, 37))
            return fail(ORBX_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    }
    {
        std::vector<uint4> tab((size_t)g.cellsPerImage);
        int maxCw = 0;
        make_cell_table(g, tab.data(), &b->fastTileRows, &maxCw);
        b->fastTs = fast_tile_stride(maxCw);
        uint4* dtab = nullptr;
        if ((r = dev_alloc(b, &dtab, tab.size()))) return r;
        CU(cudaMemcpy(dtab, tab.data(), tab.size() * sizeof(uint4), cudaMemcpyHostToDevice));
        B.cellTab = dtab;
        CU(fast_configure(b->fastTileRows, b->fastTs));
        if (fast_encode_maps(g, B.pyr, n, b->fastTs, &b->fastMaps)) return fail(ORBX_ERR_CUDA, "cuTensorMapEncodeTiled failed (FAST tiles)");
        if (blur_encode_maps(g, B.pyr, n, &b->blurMaps)) return fail(ORBX_ERR_CUDA, "cuTensorMapEncodeTiled failed (blur tiles)");
        if (resize_encode_maps(g, B.pyr, n, &b->resizeMaps)) return fail(ORBX_ERR_CUDA, "cuTensorMapEncodeTiled failed (resize tiles)");
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, b->device));
        b->numSMs = prop.multiProcessorCount;
        CU(describe_configure());
        if (stereo_smem_bytes(g.K, g.lv[0].h) <= 220 * 1024) CU(stereo_configure(g.K, g.lv[0].h));
        CU(lap_match_configure());
    }
    CU(cudaMemset(B.errFlag, 0, sizeof(int)));
    CU(cudaMemset(B.lapping, 0, (size_t)n * 2 * sizeof(int)));
    CU(cudaMemset(B.pyr, 0, b->pyrBytes));
    // resize tables
    for (int l = 1; l < g.nlevels; ++l) {
        std::vector<short> xo, xa, yo, yb;
        make_resize_tables(g.lv[l - 1].w, g.lv[l].w, true, xo, xa);
        make_resize_tables(g.lv[l - 1].h, g.lv[l].h, false, yo, yb);
        short *dxo, *dxa, *dyo, *dyb;
        if ((r = dev_alloc(b, &dxo, xo.size()))) return r;
        if ((r = dev_alloc(b, &dxa, xa.size()))) return r;
        if ((r = dev_alloc(b, &dyo, yo.size()))) return r;
        if ((r = dev_alloc(b, &dyb, yb.size()))) return r;
        CU(cudaMemcpy(dxo, xo.data(), xo.size() * 2, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(dxa, xa.data(), xa.size() * 2, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(dyo, yo.data(), yo.size() * 2, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(dyb, yb.data(), yb.size() * 2, cudaMemcpyHostToDevice));
        b->tabs[l] = ResizeTab{dxo, dxa, dyo, dyb};
    }
    return ORBX_OK;
}

int check_images(const orbx_batch* b, int first, int n) {
    if (!b) return fail(ORBX_ERR_ARG, "null batch");
    if (first < 0 || n < 0 || first + n > b->maxImages) return fail(ORBX_ERR_ARG, "image range outside [0, max_images)");
    return ORBX_OK;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
extern "C" {

const char* orbx_version(void) { return "orbx 0.1 (sm_100a)"; }
const char* orbx_last_error(void) { return g_err.c_str(); }

int orbx_batch_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast, int device, int width, int height,
                      int max_images, orbx_batch** out) {
    if (!out) return fail(ORBX_ERR_ARG, "null out pointer");
    *out = nullptr;
    if (width <= 0 || height <= 0) return fail(ORBX_ERR_EMPTY, "empty image");
    if (max_images < 1 || max_images > 65535) return fail(ORBX_ERR_ARG, "max_images must be in [1, 65535]");
    orbx_batch* b = new orbx_batch();
    int r = make_params(nfeatures, scale_factor, nlevels, ini_th_fast, min_th_fast, b->p);
    if (r == ORBX_OK) r = make_geom(b->p, width, height, max_images, b->g);
    if (r != ORBX_OK) { delete b; return r; }
    b->device = device; b->width = width; b->height = height; b->maxImages = max_images;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) {
        delete b;
        return fail(ORBX_ERR_CUDA, e != cudaSuccess ? std::string("no usable CUDA device: ") + cudaGetErrorString(e)
                                                    : std::string("CUDA device ordinal out of range (there is no CPU fallback)"));
    }
    DeviceGuard dg(device);
    auto bail = [&](int code) { orbx_batch_destroy(b); return code; };
    if ((e = cudaStreamCreateWithFlags(&b->own, cudaStreamNonBlocking)) != cudaSuccess) return bail(cuda_fail(e, "cudaStreamCreate"));
    b->stream = b->own;
    {
        const char* sv = std::getenv("ORBX_SCHED");  // experiment hook (profiles/tools): how the blur is scheduled against FAST -> quadtree
        b->sched = sv ? std::atoi(sv) : 0;
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        const int prio = b->sched == 3 ? hi : b->sched == 4 ? lo : 0;
        if ((e = cudaStreamCreateWithPriority(&b->side, cudaStreamNonBlocking, prio)) != cudaSuccess) return bail(cuda_fail(e, "cudaStreamCreate"));
    }
    if ((e = cudaEventCreateWithFlags(&b->evPyr, cudaEventDisableTiming)) != cudaSuccess) return bail(cuda_fail(e, "cudaEventCreate"));
    if ((e = cudaEventCreateWithFlags(&b->evBlur, cudaEventDisableTiming)) != cudaSuccess) return bail(cuda_fail(e, "cudaEventCreate"));
    if ((e = cudaEventCreateWithFlags(&b->evFork, cudaEventDisableTiming)) != cudaSuccess) return bail(cuda_fail(e, "cudaEventCreate"));
    for (int k = 1; k < orbx_batch::kLanes; ++k) {
        if ((e = cudaStreamCreateWithFlags(&b->laneMain[k], cudaStreamNonBlocking)) != cudaSuccess) return bail(cuda_fail(e, "cudaStreamCreate"));
        if ((e = cudaStreamCreateWithFlags(&b->laneSide[k], cudaStreamNonBlocking)) != cudaSuccess) return bail(cuda_fail(e, "cudaStreamCreate"));
    }
    for (int k = 0; k < orbx_batch::kLanes; ++k) {
        if ((e = cudaEventCreateWithFlags(&b->lanePyr[k], cudaEventDisableTiming)) != cudaSuccess) return bail(cuda_fail(e, "cudaEventCreate"));
        if ((e = cudaEventCreateWithFlags(&b->laneBlur[k], cudaEventDisableTiming)) != cudaSuccess) return bail(cuda_fail(e, "cudaEventCreate"));
        if ((e = cudaEventCreateWithFlags(&b->laneDone[k], cudaEventDisableTiming)) != cudaSuccess) return bail(cuda_fail(e, "cudaEventCreate"));
    }
    if ((r = batch_alloc(b)) != ORBX_OK) return bail(r);
    if ((e = upload_describe_constants(kPatternHost, b->p.umax)) != cudaSuccess) return bail(cuda_fail(e, "constant upload"));
    int maxCap = 0;
    for (int l = 0; l < b->g.nlevels; ++l) maxCap = std::max(maxCap, b->g.lv[l].nodeCap);
    if ((e = octree_configure(maxCap)) != cudaSuccess) return bail(cuda_fail(e, "octree shared-memory configuration"));
    for (auto& ev : b->ev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail(cuda_fail(e, "cudaEventCreate"));
    *out = b;
    return ORBX_OK;
}

void orbx_batch_destroy(orbx_batch* b) {
    if (!b) return;
    DeviceGuard dg(b->device);
    if (b->own) cudaStreamSynchronize(b->own);
    for (void* p : b->allocs) cudaFree(p);
    if (b->spDev) cudaFree(b->spDev);
    if (b->spOut) cudaFree(b->spOut);
    if (b->qDev) cudaFree(b->qDev);
    if (b->qCounts) cudaFree(b->qCounts);
    if (b->qOut) cudaFree(b->qOut);
    if (b->borderScratch) cudaFree(b->borderScratch);
    if (b->stageIn) cudaFree(b->stageIn);
    for (auto& ev : b->ev) if (ev) cudaEventDestroy(ev);
    if (b->side) { cudaStreamSynchronize(b->side); cudaStreamDestroy(b->side); }
    if (b->evPyr) cudaEventDestroy(b->evPyr);
    if (b->evBlur) cudaEventDestroy(b->evBlur);
    if (b->evFork) cudaEventDestroy(b->evFork);
    for (int k = 0; k < orbx_batch::kLanes; ++k) {
        if (b->laneMain[k]) { cudaStreamSynchronize(b->laneMain[k]); cudaStreamDestroy(b->laneMain[k]); }
        if (b->laneSide[k]) { cudaStreamSynchronize(b->laneSide[k]); cudaStreamDestroy(b->laneSide[k]); }
        if (b->lanePyr[k]) cudaEventDestroy(b->lanePyr[k]);
        if (b->laneBlur[k]) cudaEventDestroy(b->laneBlur[k]);
        if (b->laneDone[k]) cudaEventDestroy(b->laneDone[k]);
    }
    if (b->own) cudaStreamDestroy(b->own);
    delete b;
}

int orbx_batch_max_keypoints(const orbx_batch* b) { return b ? b->g.K : ORBX_ERR_ARG; }
int orbx_batch_max_images(const orbx_batch* b) { return b ? b->maxImages : ORBX_ERR_ARG; }
int orbx_batch_set_blur_taps(orbx_batch* b, const int taps[7]) {
    if (!b || !taps) return fail(ORBX_ERR_ARG, "null argument");
    int sum = 0;
    for (int i = 0; i < 7; ++i) { if (taps[i] < 0 || taps[i] > 255) return fail(ORBX_ERR_ARG, "tap outside [0,255] (the kernel packs taps into bytes)"); sum += taps[i]; }
    if (sum > 257) return fail(ORBX_ERR_ARG, "taps must sum to about 256 (8.8 fixed point)");
    if (std::memcmp(b->taps, taps, sizeof(b->taps)) != 0) ++b->tapsGeneration;  // a captured graph has the old taps baked into its kernel arguments
    std::memcpy(b->taps, taps, sizeof(b->taps));
    return ORBX_OK;
}
int orbx_batch_set_stream(orbx_batch* b, void* s) {
    if (!b) return fail(ORBX_ERR_ARG, "null batch");
    b->stream = s ? (cudaStream_t)s : b->own;
    return ORBX_OK;
}
void* orbx_batch_get_stream(const orbx_batch* b) { return b ? (void*)b->stream : nullptr; }

// n images [height][row_stride] spaced image_stride apart -> level 0 of images [first, first+n): one 3-D strided copy
static int copy_images_in(orbx_batch* b, const uint8_t* images, size_t row_stride, size_t image_stride, int first, int n, cudaMemcpyKind kind) {
    const LevelGeom& L0 = b->g.lv[0];
    uint8_t* dst0 = b->buf.pyr + L0.planeOff + (size_t)first * L0.planeBytes;
    if (kind == cudaMemcpyHostToDevice && row_stride == (size_t)b->width && (n == 1 || image_stride == row_stride * (size_t)b->height)) {
        // dense host images: one contiguous PCIe copy into device staging, then a device-side strided re-pack
        const size_t bytes = (size_t)n * b->width * b->height;
        if (bytes > b->stageBytes) {
            CU(cudaStreamSynchronize(b->stream));
            if (b->stageIn) cudaFree(b->stageIn);
            b->stageIn = nullptr; b->stageBytes = 0;
            CU(cudaMalloc((void**)&b->stageIn, bytes));
            b->stageBytes = bytes;
        }
        CU(cudaMemcpyAsync(b->stageIn, images, bytes, cudaMemcpyHostToDevice, b->stream));
        images = b->stageIn;
        image_stride = row_stride * (size_t)b->height;
        kind = cudaMemcpyDeviceToDevice;
    }
    if (n == 1 || image_stride % row_stride != 0) {
        for (int i = 0; i < n; ++i)
            CU(cudaMemcpy2DAsync(dst0 + (size_t)i * L0.planeBytes, L0.pitch, images + (size_t)i * image_stride, row_stride, b->width, b->height, kind,
                                 b->stream));
        return ORBX_OK;
    }
    cudaMemcpy3DParms p = {};
    p.srcPtr = make_cudaPitchedPtr((void*)images, row_stride, b->width, image_stride / row_stride);
    p.dstPtr = make_cudaPitchedPtr((void*)dst0, L0.pitch, b->width, (size_t)b->height + 2 * kMarginY);
    p.extent = make_cudaExtent(b->width, b->height, n);
    p.kind = kind;
    CU(cudaMemcpy3DAsync(&p, b->stream));
    return ORBX_OK;
}

int orbx_batch_upload(orbx_batch* b, const uint8_t* host_images, size_t row_stride, size_t image_stride, int first, int n) {
    int r = check_images(b, first, n);
    if (r) return r;
    if (!host_images) return fail(ORBX_ERR_EMPTY, "null image");
    if (row_stride < (size_t)b->width) return fail(ORBX_ERR_ARG, "row stride smaller than the width");
    if (n == 0) return ORBX_OK;
    DeviceGuard dg(b->device);
    return copy_images_in(b, host_images, row_stride, image_stride, first, n, cudaMemcpyHostToDevice);
}

// Tracking::GrabImageStereoLidar's cvtColor (Tracking.cc:1646-1671) fused into the upload: interleaved 8-bit colour images
// (channels 3 or 4; rgb != 0: red first = COLOR_RGB[A]2GRAY, else COLOR_BGR[A]2GRAY) -> device staging -> level 0 as grey
int orbx_batch_upload_color(orbx_batch* b, const uint8_t* host_images, size_t row_stride, size_t image_stride, int channels, int rgb, int first, int n) {
    int r = check_images(b, first, n);
    if (r) return r;
    if (!host_images) return fail(ORBX_ERR_EMPTY, "null image");
    if (channels != 3 && channels != 4) return fail(ORBX_ERR_ARG, "colour images have 3 or 4 interleaved channels");
    const size_t rowBytes = (size_t)b->width * channels;
    if (row_stride < rowBytes) return fail(ORBX_ERR_ARG, "row stride smaller than width * channels");
    if (n == 0) return ORBX_OK;
    DeviceGuard dg(b->device);
    const size_t imgBytes = rowBytes * b->height, bytes = (size_t)n * imgBytes + 64;
    if (bytes > b->stageBytes) {
        CU(cudaStreamSynchronize(b->stream));
        if (b->stageIn) cudaFree(b->stageIn);
        b->stageIn = nullptr; b->stageBytes = 0;
        CU(cudaMalloc((void**)&b->stageIn, bytes));
        b->stageBytes = bytes;
    }
    if (row_stride == rowBytes && (n == 1 || image_stride == imgBytes))
        CU(cudaMemcpyAsync(b->stageIn, host_images, (size_t)n * imgBytes, cudaMemcpyHostToDevice, b->stream));
    else
        for (int i = 0; i < n; ++i)
            CU(cudaMemcpy2DAsync(b->stageIn + (size_t)i * imgBytes, rowBytes, host_images + (size_t)i * image_stride, row_stride, rowBytes, b->height,
                                 cudaMemcpyHostToDevice, b->stream));
    const LevelGeom& L0 = b->g.lv[0];
    launch_gray(b->stageIn, channels, rgb, b->buf.pyr + L0.planeOff + (size_t)first * L0.planeBytes, L0.w, L0.h, L0.pitch, (size_t)L0.planeBytes, n, b->stream);
    CU(cudaGetLastError());
    return ORBX_OK;
}

// System::TrackStereoLidar's rectification (System.cc:238-248): the CV_32FC1 maps of Settings::M1l()/M2l() (or M1r()/M2r()), one
// float per pixel of the handle's width x height, sampling raw images of src_width x src_height.  NULL maps clear it.  Synchronises.
int orbx_batch_set_rectification(orbx_batch* b, const float* map_x, const float* map_y, size_t map_stride_bytes, int src_width, int src_height) {
    if (!b) return fail(ORBX_ERR_ARG, "null batch");
    DeviceGuard dg(b->device);
    CU(cudaStreamSynchronize(b->stream));
    b->inResize = false;
    if (!map_x || !map_y) {
        b->srcW = b->srcH = 0;
        return ORBX_OK;
    }
    if (src_width <= 0 || src_height <= 0 || src_width > ORBX_MAX_DIM || src_height > ORBX_MAX_DIM) return fail(ORBX_ERR_ARG, "raw image size out of range");
    const size_t rowBytes = (size_t)b->width * sizeof(float);
    if (map_stride_bytes < rowBytes) return fail(ORBX_ERR_ARG, "map stride smaller than width * sizeof(float)");
    if (!b->rectX) {
        int r;
        if ((r = dev_alloc(b, &b->rectX, (size_t)b->width * b->height))) return r;
        if ((r = dev_alloc(b, &b->rectY, (size_t)b->width * b->height))) return r;
    }
    CU(cudaMemcpy2D(b->rectX, rowBytes, map_x, map_stride_bytes, rowBytes, b->height, cudaMemcpyHostToDevice));
    CU(cudaMemcpy2D(b->rectY, rowBytes, map_y, map_stride_bytes, rowBytes, b->height, cudaMemcpyHostToDevice));
    b->srcW = src_width; b->srcH = src_height;
    return ORBX_OK;
}

// cv::resize(raw, newImSize) of System::TrackStereoLidar (System.cc:250-252) in front of the extractor: raw images of src_width x src_height
// given to orbx_batch_upload_raw are resized on the device to the handle's width x height; 0 x 0 switches it off
int orbx_batch_set_input_resize(orbx_batch* b, int src_width, int src_height) {
    if (!b) return fail(ORBX_ERR_ARG, "null batch");
    DeviceGuard dg(b->device);
    CU(cudaStreamSynchronize(b->stream));
    if (src_width == 0 && src_height == 0) {
        if (b->inResize) b->srcW = b->srcH = 0;
        b->inResize = false;
        return ORBX_OK;
    }
    if (src_width <= 1 || src_height <= 1 || src_width > ORBX_MAX_DIM || src_height > ORBX_MAX_DIM) return fail(ORBX_ERR_ARG, "raw image size out of range");
    if (b->inTabW != src_width || b->inTabH != src_height) {
        std::vector<short> xo, xa, yo, yb;
        make_resize_tables(src_width, b->width, true, xo, xa);
        make_resize_tables(src_height, b->height, false, yo, yb);
        int r;
        short *dxo, *dxa, *dyo, *dyb;  // (a handle changes its raw size rarely: earlier tables stay allocated until the handle is destroyed)
        if ((r = dev_alloc(b, &dxo, xo.size()))) return r;
        if ((r = dev_alloc(b, &dxa, xa.size()))) return r;
        if ((r = dev_alloc(b, &dyo, yo.size()))) return r;
        if ((r = dev_alloc(b, &dyb, yb.size()))) return r;
        CU(cudaMemcpy(dxo, xo.data(), xo.size() * 2, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(dxa, xa.data(), xa.size() * 2, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(dyo, yo.data(), yo.size() * 2, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(dyb, yb.data(), yb.size() * 2, cudaMemcpyHostToDevice));
        b->inTab = ResizeTab{dxo, dxa, dyo, dyb};
        b->inTabW = src_width; b->inTabH = src_height;
    }
    b->srcW = src_width; b->srcH = src_height;
    b->inResize = true;
    return ORBX_OK;
}

// raw (unrectified) images of src_width x src_height, channels 1, 3 or 4 -> cv::remap -> [cvtColor] -> level 0, all on the device
int orbx_batch_upload_raw(orbx_batch* b, const uint8_t* host_images, size_t row_stride, size_t image_stride, int channels, int rgb, int first, int n) {
    int r = check_images(b, first, n);
    if (r) return r;
    if (!host_images) return fail(ORBX_ERR_EMPTY, "null image");
    if (!b->srcW) return fail(ORBX_ERR_ARG, "no rectification maps (orbx_batch_set_rectification) or input resize (orbx_batch_set_input_resize) set");
    if (channels != 1 && channels != 3 && channels != 4) return fail(ORBX_ERR_ARG, "images have 1, 3 or 4 interleaved channels");
    const size_t rowBytes = (size_t)b->srcW * channels;
    if (row_stride < rowBytes) return fail(ORBX_ERR_ARG, "row stride smaller than width * channels");
    if (n == 0) return ORBX_OK;
    DeviceGuard dg(b->device);
    const size_t imgBytes = rowBytes * b->srcH, bytes = (size_t)n * imgBytes + 64;
    if (bytes > b->stageBytes) {
        CU(cudaStreamSynchronize(b->stream));
        if (b->stageIn) cudaFree(b->stageIn);
        b->stageIn = nullptr; b->stageBytes = 0;
        CU(cudaMalloc((void**)&b->stageIn, bytes));
        b->stageBytes = bytes;
    }
    if (row_stride == rowBytes && (n == 1 || image_stride == imgBytes))
        CU(cudaMemcpyAsync(b->stageIn, host_images, (size_t)n * imgBytes, cudaMemcpyHostToDevice, b->stream));
    else
        for (int i = 0; i < n; ++i)
            CU(cudaMemcpy2DAsync(b->stageIn + (size_t)i * imgBytes, rowBytes, host_images + (size_t)i * image_stride, row_stride, rowBytes, b->srcH,
                                 cudaMemcpyHostToDevice, b->stream));
    const LevelGeom& L0 = b->g.lv[0];
    if (b->inResize)
        launch_input_resize(b->stageIn, b->srcW, b->srcH, channels, rgb, b->inTab, b->buf.pyr + L0.planeOff + (size_t)first * L0.planeBytes, L0.w, L0.h, L0.pitch,
                            (size_t)L0.planeBytes, n, b->stream);
    else
        launch_remap(b->stageIn, b->srcW, b->srcH, channels, rgb, b->rectX, b->rectY, b->buf.pyr + L0.planeOff + (size_t)first * L0.planeBytes, L0.w, L0.h, L0.pitch,
                     (size_t)L0.planeBytes, n, b->stream);
    CU(cudaGetLastError());
    return ORBX_OK;
}

int orbx_batch_upload_device(orbx_batch* b, const uint8_t* dev_images, size_t row_stride, size_t image_stride, int first, int n) {
    int r = check_images(b, first, n);
    if (r) return r;
    if (!dev_images) return fail(ORBX_ERR_EMPTY, "null image");
    if (row_stride < (size_t)b->width) return fail(ORBX_ERR_ARG, "row stride smaller than the width");
    if (n == 0) return ORBX_OK;
    DeviceGuard dg(b->device);
    return copy_images_in(b, dev_images, row_stride, image_stride, first, n, cudaMemcpyDeviceToDevice);
}

// level-0 images of the handle back to the host (e.g. frames produced by orbx_batch_synth): one strided copy
int orbx_batch_download_images(orbx_batch* b, int first, int n, uint8_t* host_images, size_t row_stride, size_t image_stride) {
    int r = check_images(b, first, n);
    if (r) return r;
    if (!host_images || row_stride < (size_t)b->width || image_stride != row_stride * (size_t)b->height)
        return fail(ORBX_ERR_ARG, "need a dense [n][height][row_stride] destination");
    DeviceGuard dg(b->device);
    if (n == 0) return ORBX_OK;
    const LevelGeom& L0 = b->g.lv[0];
    cudaMemcpy3DParms p = {};
    p.srcPtr = make_cudaPitchedPtr((void*)(b->buf.pyr + L0.planeOff + (size_t)first * L0.planeBytes), L0.pitch, b->width, (size_t)b->height + 2 * kMarginY);
    p.dstPtr = make_cudaPitchedPtr((void*)host_images, row_stride, b->width, b->height);
    p.extent = make_cudaExtent(b->width, b->height, n);
    p.kind = cudaMemcpyDeviceToHost;
    CU(cudaMemcpy3DAsync(&p, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return ORBX_OK;
}

int orbx_batch_synth(orbx_batch* b, uint32_t seed0, int first, int n, int stereo) {
    int r = check_images(b, first, n);
    if (r) return r;
    if (n == 0) return ORBX_OK;
    DeviceGuard dg(b->device);
    const LevelGeom& L0 = b->g.lv[0];
    launch_synth(b->buf.pyr + L0.planeOff, L0.w, L0.h, L0.pitch, (size_t)L0.planeBytes, seed0, first, n, stereo, b->stream);
    CU(cudaGetLastError());
    return ORBX_OK;
}

// the whole launch sequence for images [first, first+n) on lane `lane` (lane 0 = the handle's own / caller's stream)
static int run_range_lane(orbx_batch* b, int first, int n, const int* lapping, int lane) {
    const Geom& g = b->g;
    cudaStream_t s = lane == 0 ? b->stream : b->laneMain[lane];
    cudaStream_t side = lane == 0 ? b->side : b->laneSide[lane];
    cudaEvent_t evPyr = lane == 0 ? b->evPyr : b->lanePyr[lane], evBlur = lane == 0 ? b->evBlur : b->laneBlur[lane];
    Buffers buf = b->buf;
    buf.imgBase = first;
    int launches = 0;
    if (lapping && b->lapStable) {
        CU(cudaMemcpyAsync(buf.lapping + 2 * (size_t)first, lapping, (size_t)n * 2 * sizeof(int), cudaMemcpyHostToDevice, s));
    } else if (lapping) {
        // (the staging vector must stay untouched until the copy has run: one slot per image of the handle)
        b->hostLap.resize(2 * (size_t)b->maxImages);
        std::memcpy(b->hostLap.data() + 2 * (size_t)first, lapping, 2 * (size_t)n * sizeof(int));
        CU(cudaMemcpyAsync(buf.lapping + 2 * (size_t)first, b->hostLap.data() + 2 * (size_t)first, (size_t)n * 2 * sizeof(int), cudaMemcpyHostToDevice, s));
    } else
        CU(cudaMemsetAsync(buf.lapping + 2 * (size_t)first, 0, (size_t)n * 2 * sizeof(int), s));
    CU(cudaMemsetAsync(buf.candCount + (size_t)first * ORBX_MAX_LEVELS, 0, (size_t)n * ORBX_MAX_LEVELS * sizeof(int), s));
    int evi = 0;
    auto mark = [&]() { if (b->timing && lane == 0) cudaEventRecord(b->ev[evi++], s); };
    mark();
    for (int l = 1; l < g.nlevels; ++l) { launch_pyramid_level(g, buf, b->resizeMaps, l, b->tabs[l], n, s); ++launches; }
    // FAST -> quadtree and the blur both depend only on the pyramid: the blur goes to the side stream unless per-stage
    // timing was requested (then everything is serialised on one stream so that the events bracket single stages).
    // The reflect-101 margin rows are read by the blur alone (FAST and the descriptor stage stay 16 / 19 pixels inside the image),
    // so k_margin travels with it and leaves the FAST -> quadtree -> descriptor chain.
    // (only for a few images, where it takes a node out of the latency chain; with many images the margin kernel beside FAST cost 0.3 ms per 8192)
    const bool overlap = !b->timing || lane != 0;
    const bool marginOnSide = overlap && n <= 16;
    if (!marginOnSide) { launch_margin(g, buf, n, s); ++launches; }
    mark();
    // a few images (the single-image handle): every level in ONE launch, strips cut into row segments so that the handful of warps becomes a
    // few hundred; many images: one launch per level (each already fills the machine, and whole strips have no re-read rows)
    int sumStrips = 0;
    for (int l = 0; l < g.nlevels; ++l) sumStrips += (g.lv[l].w + 127) / 128;
    const int blurSeg = (long long)n * sumStrips <= 256 ? 64 : (long long)n * sumStrips <= 512 ? 128 : 0;
    auto issue_blur = [&](cudaStream_t st) {
        if (marginOnSide) { launch_margin(g, buf, n, st); ++launches; }
        if (blurSeg) { launch_blur_all(g, buf, b->blurMaps, n, b->taps, blurSeg, st); ++launches; }
        else for (int l = 0; l < g.nlevels; ++l) { launch_blur(g, buf, b->blurMaps, l, n, b->taps, st); ++launches; }
    };
    auto blur_on_side = [&]() -> int {
        CU(cudaEventRecord(evPyr, s));
        CU(cudaStreamWaitEvent(side, evPyr, 0));
        issue_blur(side);
        CU(cudaEventRecord(evBlur, side));
        return ORBX_OK;
    };
    const bool blurFirst = b->sched == 0 || b->sched == 3 || b->sched == 4;
    if (overlap && blurFirst) { int r = blur_on_side(); if (r) return r; }
    if (overlap && b->sched == 1) CU(cudaEventRecord(evPyr, s));
    launch_fast(g, buf, b->fastMaps, n, b->fastTileRows, b->fastTs, b->numSMs, s); ++launches;
    if (overlap && b->sched == 1) {  // the blur is enqueued after FAST but still only waits for the pyramid
        CU(cudaStreamWaitEvent(side, evPyr, 0));
        issue_blur(side);
        CU(cudaEventRecord(evBlur, side));
    }
    if (overlap && b->sched == 2) { int r = blur_on_side(); if (r) return r; }  // the blur waits for FAST: it runs beside the quadtree only
    mark();
    if (!overlap) issue_blur(s);
    mark();
    {
        cudaError_t e = launch_octree(g, buf, n, b->numSMs, s); launches += octree_launch_count(g, n, b->numSMs);
        if (e != cudaSuccess) return cuda_fail(e, "octree launch");
    }
    mark();
    launch_assemble(g, buf, n, s); ++launches;
    mark();
    if (overlap) CU(cudaStreamWaitEvent(s, evBlur, 0));
    launch_describe(g, buf, b->descMaps, n, s); ++launches;
    mark();
    CU(cudaGetLastError());
    b->launches = launches;
    b->lastImages = n;
    return ORBX_OK;
}

int orbx_batch_run_range(orbx_batch* b, int first, int n, const int* lapping) {
    int r = check_images(b, first, n);
    if (r) return r;
    if (n == 0) return ORBX_OK;
    DeviceGuard dg(b->device);
    b->unImages = 0;   // new keypoints: mvKeysUn is mvKeys again until the caller undistorts them
    return run_range_lane(b, first, n, lapping, 0);
}

// images [0, n) in chunks of `chunk` images issued round-robin over the handle's lanes (streams); returns when everything is
// ENQUEUED; the handle's stream is made to wait for all lanes, so orbx_batch_sync / download behave as after orbx_batch_run
int orbx_batch_run_pipelined(orbx_batch* b, int n, const int* lapping, int chunk) {
    int r = check_images(b, 0, n);
    if (r) return r;
    if (n == 0) return ORBX_OK;
    if (chunk <= 0) return fail(ORBX_ERR_ARG, "chunk must be positive");
    DeviceGuard dg(b->device);
    b->unImages = 0;
    if (b->timing || chunk >= n) return run_range_lane(b, 0, n, lapping, 0);
    CU(cudaEventRecord(b->evFork, b->stream));
    for (int k = 1; k < orbx_batch::kLanes; ++k) CU(cudaStreamWaitEvent(b->laneMain[k], b->evFork, 0));
    int c = 0, launches = 0;
    for (int first = 0; first < n; first += chunk, ++c) {
        const int m = std::min(chunk, n - first);
        if ((r = run_range_lane(b, first, m, lapping ? lapping + 2 * (size_t)first : nullptr, c % orbx_batch::kLanes))) return r;
        launches += b->launches;
    }
    for (int k = 1; k < orbx_batch::kLanes; ++k) {
        CU(cudaEventRecord(b->laneDone[k], b->laneMain[k]));
        CU(cudaStreamWaitEvent(b->stream, b->laneDone[k], 0));
    }
    b->launches = launches;
    b->lastImages = n;
    return ORBX_OK;
}

int orbx_batch_run(orbx_batch* b, int n, const int* lapping) { return orbx_batch_run_range(b, 0, n, lapping); }

// Frame::ComputeStereoMatches (Frame.cc:841-1011) for frames [0, n_frames): image 2f is the left, 2f+1 the right image of frame f.
// Asynchronous on the handle's stream, after orbx_batch_run of the same images.
int orbx_batch_stereo_match(orbx_batch* b, int n_frames, float mb, float mbf) {
    int r = check_images(b, 0, 2 * n_frames);
    if (r) return r;
    if (n_frames == 0) return ORBX_OK;
    if (!(mb > 0.f) || !(mbf > 0.f)) return fail(ORBX_ERR_ARG, "mb and mbf must be positive");
    if (b->p.scaleFactorF > 1.9f) return fail(ORBX_ERR_ARG, "stereo matching needs scaleFactor <= 1.9 (beyond it the reference's SAD window, Frame.cc:950, leaves the image and cv::Mat::colRange asserts)");
    if (stereo_smem_bytes(b->g.K, b->g.lv[0].h) > 220 * 1024) return fail(ORBX_ERR_TOO_LARGE, "keypoint capacity too large for the stereo matcher's shared memory");
    DeviceGuard dg(b->device);
    Buffers buf = b->buf;
    buf.imgBase = 0;
    launch_stereo_match(b->g, buf, buf, n_frames, 2, 0, 2, 1, mb, mbf, b->stream);
    CU(cudaGetLastError());
    return ORBX_OK;
}

// uright / depth: [n_frames][K] floats (row f = left image of frame f), -1 where the left keypoint has no match; synchronises
int orbx_batch_download_stereo(orbx_batch* b, int n_frames, float* uright, float* depth) {
    int r = check_images(b, 0, 2 * n_frames);
    if (r) return r;
    DeviceGuard dg(b->device);
    const size_t K = (size_t)b->g.K;
    if (n_frames) {
        // left images are the even ones: a strided copy of every second row block
        if (uright) CU(cudaMemcpy2DAsync(uright, K * 4, b->buf.uRight, 2 * K * 4, K * 4, n_frames, cudaMemcpyDeviceToHost, b->stream));
        if (depth) CU(cudaMemcpy2DAsync(depth, K * 4, b->buf.depth, 2 * K * 4, K * 4, n_frames, cudaMemcpyDeviceToHost, b->stream));
    }
    CU(cudaStreamSynchronize(b->stream));
    return ORBX_OK;
}



// Frame::UndistortKeyPoints (Frame.cc:777-810) for images [0, n_images) on the resident keypoints: cv::undistortPoints(mat, mat, toK(), mDistCoef, cv::Mat(), mK).
// camera / new_camera = {fx, fy, cx, cy} of toK() / mK; dist = mDistCoef (4, 5, 8 or 12 floats).  dist == NULL or dist[0] == 0 means mvKeysUn = mvKeys
// (the reference's early return).  Asynchronous; the grid and the single-grid searches built afterwards read the undistorted coordinates.
int orbx_batch_undistort_keypoints(orbx_batch* b, int n_images, const float camera[4], const float* dist, int n_dist, const float new_camera[4]) {
    int r = check_images(b, 0, n_images);
    if (r) return r;
    if (!dist || n_dist <= 0 || dist[0] == 0.0f) { b->unImages = 0; return ORBX_OK; }
    if (!camera || !new_camera || (n_dist != 4 && n_dist != 5 && n_dist != 8 && n_dist != 12)) return fail(ORBX_ERR_ARG, "camera matrices and 4, 5, 8 or 12 distortion coefficients are needed");
    if (!(camera[0] != 0.f) || !(camera[1] != 0.f)) return fail(ORBX_ERR_ARG, "zero focal length");
    DeviceGuard dg(b->device);
    if (!b->kpsUnBuf && (r = dev_alloc(b, &b->kpsUnBuf, (size_t)b->maxImages * b->g.K))) return r;
    UndistortParams p;
    p.cx = camera[2]; p.cy = camera[3];
    p.ifx = 1. / (double)camera[0]; p.ify = 1. / (double)camera[1];
    for (int i = 0; i < 12; ++i) p.k[i] = i < n_dist ? (double)dist[i] : 0.0;
    p.nfx = new_camera[0]; p.nfy = new_camera[1]; p.ncx = new_camera[2]; p.ncy = new_camera[3];
    b->unImages = n_images;
    if (n_images == 0) return ORBX_OK;
    Buffers buf = b->buf;
    buf.imgBase = 0;
    launch_undistort(buf, b->g.K, p, b->kpsUnBuf, n_images, b->stream);
    CU(cudaGetLastError());
    return ORBX_OK;
}

// xy: [n_images][K][2] floats (Frame::mvKeysUn[i].pt); synchronises
int orbx_batch_download_undistorted(orbx_batch* b, int n_images, float* xy) {
    int r = check_images(b, 0, n_images);
    if (r) return r;
    if (!xy) return fail(ORBX_ERR_ARG, "null output");
    if (n_images > b->unImages) return fail(ORBX_ERR_ARG, "orbx_batch_undistort_keypoints has not been run for these images (or the coefficients were zero: mvKeysUn = mvKeys)");
    DeviceGuard dg(b->device);
    if (n_images) CU(cudaMemcpyAsync(xy, b->kpsUnBuf, (size_t)n_images * b->g.K * sizeof(float2), cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return ORBX_OK;
}

// Frame::AssignFeaturesToGrid (Frame.cc:412-443) for images [0, n_images) on the resident keypoints; asynchronous on the handle's stream
int orbx_batch_assign_grid(orbx_batch* b, int n_images, float min_x, float max_x, float min_y, float max_y) {
    int r = check_images(b, 0, n_images);
    if (r) return r;
    if (!(max_x > min_x) || !(max_y > min_y)) return fail(ORBX_ERR_ARG, "grid bounds must satisfy max > min");
    DeviceGuard dg(b->device);
    if (!b->buf.gridStart) {
        if (grid_smem_bytes(b->g.K) > 180 * 1024) return fail(ORBX_ERR_TOO_LARGE, "keypoint capacity too large for the grid kernel's shared memory");
        if ((r = dev_alloc(b, &b->buf.gridStart, (size_t)b->maxImages * (ORBX_GRID_CELLS + 1)))) return r;
        if ((r = dev_alloc(b, &b->buf.gridItems, (size_t)b->maxImages * b->g.K))) return r;
        CU(grid_configure(b->g.K));
    }
    // mfGridElementWidthInv / HeightInv exactly as the Frame constructors compute them (Frame.cc:179-180): float / (float - float)
    b->gridParams.minX = min_x; b->gridParams.minY = min_y;
    b->gridParams.invW = static_cast<float>(ORBX_GRID_COLS) / (max_x - min_x);
    b->gridParams.invH = static_cast<float>(ORBX_GRID_ROWS) / (max_y - min_y);
    if (b->unImages > 0 && n_images > b->unImages) return fail(ORBX_ERR_ARG, "orbx_batch_undistort_keypoints covered fewer images than this grid asks for");
    b->gridImages = n_images;
    b->gridUn = b->unImages > 0;
    if (n_images == 0) return ORBX_OK;
    Buffers buf = b->buf;
    buf.imgBase = 0;
    buf.kpsUn = b->gridUn ? b->kpsUnBuf : nullptr;
    launch_assign_grid(buf, b->g.K, b->gridParams, n_images, b->stream);
    CU(cudaGetLastError());
    return ORBX_OK;
}

int orbx_batch_download_grid(orbx_batch* b, int n_images, int* cell_start, int* items) {
    int r = check_images(b, 0, n_images);
    if (r) return r;
    if (n_images > b->gridImages) return fail(ORBX_ERR_ARG, "orbx_batch_assign_grid has not been run for these images");
    DeviceGuard dg(b->device);
    if (n_images) {
        if (cell_start) CU(cudaMemcpyAsync(cell_start, b->buf.gridStart, (size_t)n_images * (ORBX_GRID_CELLS + 1) * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
        if (items) CU(cudaMemcpyAsync(items, b->buf.gridItems, (size_t)n_images * b->g.K * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    }
    CU(cudaStreamSynchronize(b->stream));
    return ORBX_OK;
}

// Frame::GetFeaturesInArea (Frame.cc:687-749), batched: one warp per query
int orbx_batch_features_in_area(orbx_batch* b, const orbx_area_query* queries, int n_queries, int capacity, int* counts, int* indices) {
    if (!b) return fail(ORBX_ERR_ARG, "null handle");
    if (n_queries < 0 || capacity < 0 || (n_queries && (!queries || !counts)) || (n_queries && capacity && !indices)) return fail(ORBX_ERR_ARG, "bad query arguments");
    if (n_queries == 0) return ORBX_OK;
    for (int q = 0; q < n_queries; ++q)
        if (queries[q].image < 0 || queries[q].image >= b->gridImages) return fail(ORBX_ERR_ARG, "query refers to an image without a grid (run orbx_batch_assign_grid first)");
    DeviceGuard dg(b->device);
    const size_t outElems = std::max<size_t>((size_t)n_queries * (size_t)capacity, 1);
    if ((size_t)n_queries > b->qCap) {
        if (b->qDev) cudaFree(b->qDev);
        if (b->qCounts) cudaFree(b->qCounts);
        b->qDev = nullptr; b->qCounts = nullptr; b->qCap = 0;
        CU(cudaMalloc(&b->qDev, (size_t)n_queries * sizeof(orbx_area_query)));
        CU(cudaMalloc(&b->qCounts, (size_t)n_queries * sizeof(int)));
        b->qCap = (size_t)n_queries;
    }
    if (outElems > b->qOutCap) {
        if (b->qOut) cudaFree(b->qOut);
        b->qOut = nullptr; b->qOutCap = 0;
        CU(cudaMalloc(&b->qOut, outElems * sizeof(int)));
        b->qOutCap = outElems;
    }
    CU(cudaMemcpyAsync(b->qDev, queries, (size_t)n_queries * sizeof(orbx_area_query), cudaMemcpyHostToDevice, b->stream));
    Buffers buf = b->buf;
    buf.imgBase = 0;
    buf.kpsUn = b->gridUn ? b->kpsUnBuf : nullptr;
    launch_features_in_area(buf, b->g.K, b->gridParams, b->qDev, n_queries, capacity, b->qCounts, b->qOut, b->stream);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(counts, b->qCounts, (size_t)n_queries * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    if (capacity) CU(cudaMemcpyAsync(indices, b->qOut, (size_t)n_queries * capacity * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return ORBX_OK;
}

// both projection searches (search_kernel.cu); local = the local-map overload; image2 >= 0 = a two-grid frame (side: n_queries host bytes; taken / taken2: host
// bytes [K] per image or NULL; l2r / r2l: host ints [K] or NULL; match / match2: host ints [K] per image); synchronous
struct SearchArgs {
    bool local = false;
    int image = 0, image2 = -1;
    const orbx_proj_query* queries = nullptr;
    const uint8_t* side = nullptr;
    int n_queries = 0, use_uright = 0, check_orientation = 0, th_high = 100;
    float nn_ratio = 0.f;
    const uint8_t *taken = nullptr, *taken2 = nullptr;
    const int *l2r = nullptr, *r2l = nullptr;
    int *match = nullptr, *match2 = nullptr, *n_matches = nullptr;
};

static int search_common(orbx_batch* b, const SearchArgs& a) {
    if (!b) return fail(ORBX_ERR_ARG, "null handle");
    const int n_queries = a.n_queries;
    const bool two = a.image2 >= 0;
    if (n_queries < 0 || (n_queries && !a.queries) || !a.match || !a.n_matches || (two && (!a.match2 || (n_queries && !a.side)))) return fail(ORBX_ERR_ARG, "bad search arguments");
    if (a.image < 0 || a.image >= b->gridImages || (two && a.image2 >= b->gridImages)) return fail(ORBX_ERR_ARG, "image has no feature grid (run orbx_batch_assign_grid first)");
    if (two && a.image2 == a.image) return fail(ORBX_ERR_ARG, "a two-grid frame needs two different images");
    if (n_queries > 65535) return fail(ORBX_ERR_TOO_LARGE, "more than 65535 map points in one search");
    if (two)
        for (int i = 0; i < n_queries; ++i)
            if (a.side[i] > 2 || (a.side[i] == 1 && (i == 0 || a.side[i - 1] != 0))) return fail(ORBX_ERR_ARG, "side[i] must be 0, 2, or 1 right after a 0");
    DeviceGuard dg(b->device);
    const int K = b->g.K, sides = two ? 2 : 1, KK = K * sides;
    if (KK > (two ? 32767 : 65535)) return fail(ORBX_ERR_TOO_LARGE, "keypoint capacity too large for the search kernel's 16-bit indices");   // (two-grid: signed 16-bit partner links)
    const size_t budget = kSearchSmemBudget;
    const bool anyTaken = a.taken || a.taken2, anyLinks = two && a.local && (a.l2r || a.r2l);
    // the frame's descriptors are staged in shared memory when at least 32 KB remain for the candidate list next to them
    const int stageDesc = search_smem_bytes(K, n_queries, 0, 1, sides, anyLinks) + 32 * 1024 <= budget ? 1 : 0;
    const size_t fixed = search_smem_bytes(K, n_queries, 0, stageDesc, sides, anyLinks);
    if (fixed + 4096 > budget) return fail(ORBX_ERR_TOO_LARGE, "keypoint capacity / number of map points too large for the search kernel's shared memory");
    const int candCap = (int)((budget - fixed) / 4);
    if (!b->spConfigured) { CU(search_configure()); b->spConfigured = true; }
    // per-call scratch: match[2K], info[2], links[2K] (ints) | taken0[2K] (bytes); side[] rides behind the queries
    if (!b->spOut) CU(cudaMalloc(&b->spOut, ((size_t)4 * K + 2) * sizeof(int) + (size_t)2 * K));
    if ((size_t)n_queries > b->spCap) {
        if (b->spDev) cudaFree(b->spDev);
        b->spDev = nullptr; b->spCap = 0;
        CU(cudaMalloc(&b->spDev, (size_t)n_queries * (sizeof(orbx_proj_query) + 1)));
        b->spCap = (size_t)n_queries;
    }
    uint8_t* sideDev = reinterpret_cast<uint8_t*>(b->spDev + b->spCap);
    if (n_queries) CU(cudaMemcpyAsync(b->spDev, a.queries, (size_t)n_queries * sizeof(orbx_proj_query), cudaMemcpyHostToDevice, b->stream));
    if (n_queries && two) CU(cudaMemcpyAsync(sideDev, a.side, (size_t)n_queries, cudaMemcpyHostToDevice, b->stream));
    int* infoDev = b->spOut + 2 * K;
    int* linksDev = infoDev + 2;
    uint8_t* takenDev = reinterpret_cast<uint8_t*>(linksDev + 2 * K);
    if (anyTaken) {
        if (a.taken) CU(cudaMemcpyAsync(takenDev, a.taken, (size_t)K, cudaMemcpyHostToDevice, b->stream)); else CU(cudaMemsetAsync(takenDev, 0, (size_t)K, b->stream));
        if (two) { if (a.taken2) CU(cudaMemcpyAsync(takenDev + K, a.taken2, (size_t)K, cudaMemcpyHostToDevice, b->stream)); else CU(cudaMemsetAsync(takenDev + K, 0, (size_t)K, b->stream)); }
    }
    if (anyLinks) {
        if (a.l2r) CU(cudaMemcpyAsync(linksDev, a.l2r, (size_t)K * sizeof(int), cudaMemcpyHostToDevice, b->stream)); else CU(cudaMemsetAsync(linksDev, 0xFF, (size_t)K * sizeof(int), b->stream));
        if (a.r2l) CU(cudaMemcpyAsync(linksDev + K, a.r2l, (size_t)K * sizeof(int), cudaMemcpyHostToDevice, b->stream)); else CU(cudaMemsetAsync(linksDev + K, 0xFF, (size_t)K * sizeof(int), b->stream));
    }
    Buffers buf = b->buf;
    buf.imgBase = 0;
    buf.kpsUn = (b->gridUn && !two) ? b->kpsUnBuf : nullptr;
    if (a.local)
        launch_search_local_points(buf, a.image, K, b->gridParams, b->spDev, n_queries, (a.use_uright && !two) ? 1 : 0, a.th_high, a.nn_ratio, anyTaken ? takenDev : nullptr, candCap,
                                   stageDesc, b->spOut, infoDev, b->stream, a.image2, two ? sideDev : nullptr, anyLinks ? linksDev : nullptr);
    else
        launch_search_by_projection(buf, a.image, K, b->gridParams, b->spDev, n_queries, (a.use_uright && !two) ? 1 : 0, a.check_orientation ? 1 : 0, a.th_high,
                                    anyTaken ? takenDev : nullptr, candCap, stageDesc, b->spOut, infoDev, b->stream, a.image2, two ? sideDev : nullptr);
    CU(cudaGetLastError());
    int info[2] = {0, 0};
    CU(cudaMemcpyAsync(a.match, b->spOut, (size_t)K * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    if (two) CU(cudaMemcpyAsync(a.match2, b->spOut + K, (size_t)K * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CU(cudaMemcpyAsync(info, infoDev, 2 * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    if (info[0] < 0) return fail(ORBX_ERR_CAPACITY, "the search windows hold more candidates than the kernel's shared memory (smaller th / fewer points)");
    *a.n_matches = info[0];
    return ORBX_OK;
}

// ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) from ORBmatcher.cc:1729 on
int orbx_batch_search_by_projection(orbx_batch* b, int image, const orbx_proj_query* queries, int n_queries, int use_uright, int check_orientation, int th_high,
                                    int* match, int* n_matches) {
    SearchArgs a;
    a.image = image; a.queries = queries; a.n_queries = n_queries; a.use_uright = use_uright; a.check_orientation = check_orientation; a.th_high = th_high;
    a.match = match; a.n_matches = n_matches;
    return search_common(b, a);
}

// ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1898-2019), the relocalisation search, behind its projection
int orbx_batch_search_keyframe_points(orbx_batch* b, int image, const orbx_proj_query* queries, int n_queries, int check_orientation, int orb_dist,
                                      const uint8_t* taken, int* match, int* n_matches) {
    SearchArgs a;
    a.image = image; a.queries = queries; a.n_queries = n_queries; a.check_orientation = check_orientation; a.th_high = orb_dist; a.taken = taken;
    a.match = match; a.n_matches = n_matches;
    return search_common(b, a);
}

// ORBmatcher::SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints) (ORBmatcher.cc:52-222), the local-map search
int orbx_batch_search_local_points(orbx_batch* b, int image, const orbx_proj_query* queries, int n_queries, int use_uright, float nn_ratio, int th_high,
                                   const uint8_t* taken, int* match, int* n_matches) {
    SearchArgs a;
    a.local = true; a.image = image; a.queries = queries; a.n_queries = n_queries; a.use_uright = use_uright; a.nn_ratio = nn_ratio; a.th_high = th_high; a.taken = taken;
    a.match = match; a.n_matches = n_matches;
    return search_common(b, a);
}

// the two-grid (Frame::Nleft != -1) branches: ORBmatcher.cc:1803-1868 and :153-218
int orbx_batch_search_by_projection_stereo(orbx_batch* b, int image_left, int image_right, const orbx_proj_query* queries, const uint8_t* side, int n_queries,
                                           int check_orientation, int th_high, int* match_left, int* match_right, int* n_matches) {
    if (image_right < 0) return fail(ORBX_ERR_ARG, "bad right image");
    SearchArgs a;
    a.image = image_left; a.image2 = image_right; a.queries = queries; a.side = side; a.n_queries = n_queries; a.check_orientation = check_orientation; a.th_high = th_high;
    a.match = match_left; a.match2 = match_right; a.n_matches = n_matches;
    return search_common(b, a);
}

int orbx_batch_search_local_points_stereo(orbx_batch* b, int image_left, int image_right, const orbx_proj_query* queries, const uint8_t* side, int n_queries,
                                          float nn_ratio, int th_high, const uint8_t* taken_left, const uint8_t* taken_right, const int* left_to_right,
                                          const int* right_to_left, int* match_left, int* match_right, int* n_matches) {
    if (image_right < 0) return fail(ORBX_ERR_ARG, "bad right image");
    SearchArgs a;
    a.local = true; a.image = image_left; a.image2 = image_right; a.queries = queries; a.side = side; a.n_queries = n_queries; a.nn_ratio = nn_ratio; a.th_high = th_high;
    a.taken = taken_left; a.taken2 = taken_right; a.l2r = left_to_right; a.r2l = right_to_left;
    a.match = match_left; a.match2 = match_right; a.n_matches = n_matches;
    return search_common(b, a);
}

// Descriptor stage of Frame::ComputeStereoFishEyeMatches (Frame.cc:1162-1187) for frames [0, n_frames): kNN (k = 2) of the left
// lapping tail against the right lapping tail + Lowe's gate.  Asynchronous, after orbx_batch_run of the same images.
int orbx_batch_lapping_match(orbx_batch* b, int n_frames) {
    int r = check_images(b, 0, 2 * n_frames);
    if (r) return r;
    if (n_frames == 0) return ORBX_OK;
    if (b->g.K > 65535) return fail(ORBX_ERR_TOO_LARGE, "keypoint capacity above 65535: train indices do not fit the matcher's 16-bit field");
    DeviceGuard dg(b->device);
    Buffers buf = b->buf;
    buf.imgBase = 0;
    launch_lap_match(b->g, buf, buf, n_frames, 2, 0, 2, 1, b->stream);
    CU(cudaGetLastError());
    return ORBX_OK;
}

// knn [n_frames][K], ratio_ok [n_frames][K] (row f = left image of frame f; entries [0, N_left - mono_left) are valid); synchronises
int orbx_batch_download_lapping(orbx_batch* b, int n_frames, orbx_knn2* knn, uint8_t* ratio_ok) {
    int r = check_images(b, 0, 2 * n_frames);
    if (r) return r;
    DeviceGuard dg(b->device);
    const size_t K = (size_t)b->g.K;
    if (n_frames) {
        if (knn) CU(cudaMemcpy2DAsync(knn, K * 16, b->buf.lapKnn, 2 * K * 16, K * 16, n_frames, cudaMemcpyDeviceToHost, b->stream));
        if (ratio_ok) CU(cudaMemcpy2DAsync(ratio_ok, K, b->buf.lapOk, 2 * K, K, n_frames, cudaMemcpyDeviceToHost, b->stream));
    }
    CU(cudaStreamSynchronize(b->stream));
    return ORBX_OK;
}

// Frame::ComputeBoW (Frame.cc:768-775) for images [0, n_images) on the resident descriptors; asynchronous on the handle's stream
int orbx_batch_compute_bow(orbx_batch* b, const orbx_vocabulary* v, int n_images, int levelsup) {
    int r = check_images(b, 0, n_images);
    if (r) return r;
    if (!v) return fail(ORBX_ERR_ARG, "null vocabulary");
    if (vocabulary_device_ordinal(v) != b->device) return fail(ORBX_ERR_ARG, "the vocabulary lives on another device than the handle");
    const int K = b->g.K;
    if (K > ORBX_BOW_MAX_FEATURES) return fail(ORBX_ERR_TOO_LARGE, "keypoint capacity above ORBX_BOW_MAX_FEATURES");
    DeviceGuard dg(b->device);
    if (!b->bowLeaf) {
        const size_t n = (size_t)b->maxImages;
        if ((r = dev_alloc(b, &b->bow.values, n * K))) return r;
        if ((r = dev_alloc(b, &b->bowLeaf, n * K))) return r;
        if ((r = dev_alloc(b, &b->bowNid, n * K))) return r;
        if ((r = dev_alloc(b, &b->bow.words, n * K))) return r;
        if ((r = dev_alloc(b, &b->bow.nodes, n * K))) return r;
        if ((r = dev_alloc(b, &b->bow.features, n * K))) return r;
        if ((r = dev_alloc(b, &b->bow.nodeStart, n * ((size_t)K + 1)))) return r;
        if ((r = dev_alloc(b, &b->bow.nWords, n))) return r;
        if ((r = dev_alloc(b, &b->bow.nNodes, n))) return r;
        CU(bow_configure(K));
    }
    b->bowImages = n_images;
    if (n_images == 0) return ORBX_OK;
    launch_bow(vocabulary_device(v), b->buf.desc, b->buf.outCount, 0, K, n_images, levelsup, b->bowLeaf, b->bowNid, b->bow, b->stream);
    CU(cudaGetLastError());
    return ORBX_OK;
}

int orbx_batch_download_bow(orbx_batch* b, int n_images, uint32_t* words, double* values, int* n_words, uint32_t* nodes, int* node_start, uint32_t* features,
                            int* n_nodes) {
    int r = check_images(b, 0, n_images);
    if (r) return r;
    if (n_images > b->bowImages) return fail(ORBX_ERR_ARG, "orbx_batch_compute_bow has not been run for these images");
    DeviceGuard dg(b->device);
    const size_t n = (size_t)n_images, K = (size_t)b->g.K;
    if (n_images) {
        if (words) CU(cudaMemcpyAsync(words, b->bow.words, n * K * 4, cudaMemcpyDeviceToHost, b->stream));
        if (values) CU(cudaMemcpyAsync(values, b->bow.values, n * K * 8, cudaMemcpyDeviceToHost, b->stream));
        if (n_words) CU(cudaMemcpyAsync(n_words, b->bow.nWords, n * 4, cudaMemcpyDeviceToHost, b->stream));
        if (nodes) CU(cudaMemcpyAsync(nodes, b->bow.nodes, n * K * 4, cudaMemcpyDeviceToHost, b->stream));
        if (node_start) CU(cudaMemcpyAsync(node_start, b->bow.nodeStart, n * (K + 1) * 4, cudaMemcpyDeviceToHost, b->stream));
        if (features) CU(cudaMemcpyAsync(features, b->bow.features, n * K * 4, cudaMemcpyDeviceToHost, b->stream));
        if (n_nodes) CU(cudaMemcpyAsync(n_nodes, b->bow.nNodes, n * 4, cudaMemcpyDeviceToHost, b->stream));
    }
    CU(cudaStreamSynchronize(b->stream));
    return ORBX_OK;
}

int orbx_batch_sync(orbx_batch* b) {
    if (!b) return fail(ORBX_ERR_ARG, "null batch");
    DeviceGuard dg(b->device);
    CU(cudaStreamSynchronize(b->stream));
    return ORBX_OK;
}

int orbx_batch_download_async(orbx_batch* b, int first, int n, orbx_keypoint* keypoints, uint8_t* descriptors, int* counts, int* mono_index) {
    int r = check_images(b, first, n);
    if (r) return r;
    if (n == 0) return ORBX_OK;
    DeviceGuard dg(b->device);
    cudaStream_t s = b->stream;
    const size_t K = (size_t)b->g.K;
    if (keypoints) CU(cudaMemcpyAsync(keypoints, b->buf.kps + first * K, n * K * sizeof(orbx_keypoint), cudaMemcpyDeviceToHost, s));
    if (descriptors) CU(cudaMemcpyAsync(descriptors, b->buf.desc + first * K * 32, n * K * 32, cudaMemcpyDeviceToHost, s));
    if (counts) CU(cudaMemcpyAsync(counts, b->buf.outCount + first, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (mono_index) CU(cudaMemcpyAsync(mono_index, b->buf.outMono + first, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    return ORBX_OK;
}

int orbx_batch_check(orbx_batch* b) {
    if (!b) return fail(ORBX_ERR_ARG, "null batch");
    DeviceGuard dg(b->device);
    int err = 0;
    CU(cudaMemcpyAsync(&err, b->buf.errFlag, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    if (err) {
        cudaMemsetAsync(b->buf.errFlag, 0, sizeof(int), b->stream);
        return fail(ORBX_ERR_CAPACITY, err == 1 ? "device candidate buffer overflow" : "device quadtree capacity exceeded");
    }
    return ORBX_OK;
}

int orbx_batch_download(orbx_batch* b, int n, orbx_keypoint* keypoints, uint8_t* descriptors, int* counts, int* mono_index) {
    int r = orbx_batch_download_async(b, 0, n, keypoints, descriptors, counts, mono_index);
    if (r) return r;
    return orbx_batch_check(b);
}

// ---- CUDA events on the handles' streams: device-side timing and cross-handle ordering for pipelined callers ----
void* orbx_event_create(const orbx_batch* b) {
    if (!b) { g_err = "null batch"; return nullptr; }
    DeviceGuard dg(b->device);  // an event belongs to the device that is current when it is created
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) { g_err = "cudaEventCreate failed"; return nullptr; }
    return (void*)e;
}
void orbx_event_destroy(void* ev) { if (ev) cudaEventDestroy((cudaEvent_t)ev); }
int orbx_event_record(void* ev, orbx_batch* b) {
    if (!ev || !b) return fail(ORBX_ERR_ARG, "null argument");
    DeviceGuard dg(b->device);
    CU(cudaEventRecord((cudaEvent_t)ev, b->stream));
    return ORBX_OK;
}
int orbx_event_wait(orbx_batch* b, void* ev) {
    if (!ev || !b) return fail(ORBX_ERR_ARG, "null argument");
    DeviceGuard dg(b->device);
    CU(cudaStreamWaitEvent(b->stream, (cudaEvent_t)ev, 0));
    return ORBX_OK;
}
int orbx_event_synchronize(void* ev) {
    if (!ev) return fail(ORBX_ERR_ARG, "null argument");
    CU(cudaEventSynchronize((cudaEvent_t)ev));
    return ORBX_OK;
}
int orbx_event_elapsed_ms(void* start, void* stop, float* ms) {
    if (!start || !stop || !ms) return fail(ORBX_ERR_ARG, "null argument");
    CU(cudaEventSynchronize((cudaEvent_t)stop));
    CU(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return ORBX_OK;
}

int orbx_batch_get_level_size(const orbx_batch* b, int level, int* width, int* height) {
    if (!b || level < 0 || level >= b->g.nlevels) return fail(ORBX_ERR_ARG, "bad level");
    if (width) *width = b->g.lv[level].w;
    if (height) *height = b->g.lv[level].h;
    return ORBX_OK;
}

static int copy_plane(orbx_batch* b, const uint8_t* base, int image, int level, uint8_t* dst, size_t dst_stride, int border) {
    if (!b || level < 0 || level >= b->g.nlevels || image < 0 || image >= b->maxImages || !dst) return fail(ORBX_ERR_ARG, "bad argument");
    if (border != 0 && border != kEdge) return fail(ORBX_ERR_ARG, "border must be 0 or 19");
    DeviceGuard dg(b->device);
    const LevelGeom& L = b->g.lv[level];
    const uint8_t* src = base + L.planeOff + (size_t)image * L.planeBytes;
    if (border == 0) {
        if (dst_stride < (size_t)L.w) return fail(ORBX_ERR_ARG, "dst stride too small");
        CU(cudaMemcpy2DAsync(dst, dst_stride, src, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost, b->stream));
    } else {
        const int bw = L.w + 2 * kEdge, bh = L.h + 2 * kEdge;
        if (dst_stride < (size_t)bw) return fail(ORBX_ERR_ARG, "dst stride too small");
        const size_t need = (size_t)bw * bh;
        if (need > b->borderScratchBytes) {
            if (b->borderScratch) cudaFree(b->borderScratch);
            b->borderScratch = nullptr; b->borderScratchBytes = 0;
            CU(cudaMalloc((void**)&b->borderScratch, need));
            b->borderScratchBytes = need;
        }
        launch_border(src, L.w, L.h, L.pitch, b->borderScratch, bw, b->stream);
        CU(cudaGetLastError());
        CU(cudaMemcpy2DAsync(dst, dst_stride, b->borderScratch, bw, bw, bh, cudaMemcpyDeviceToHost, b->stream));
    }
    CU(cudaStreamSynchronize(b->stream));
    return ORBX_OK;
}

int orbx_batch_get_level(orbx_batch* b, int image, int level, uint8_t* dst, size_t dst_stride, int border) {
    return copy_plane(b, b ? b->buf.pyr : nullptr, image, level, dst, dst_stride, border);
}
int orbx_batch_debug_blurred(orbx_batch* b, int image, int level, uint8_t* dst, size_t dst_stride) {
    return copy_plane(b, b ? b->buf.blur : nullptr, image, level, dst, dst_stride, 0);
}

int orbx_batch_enable_timing(orbx_batch* b, int on) {
    if (!b) return fail(ORBX_ERR_ARG, "null batch");
    b->timing = on != 0;
    return ORBX_OK;
}
int orbx_batch_stage_ms(orbx_batch* b, float ms[ORBX_NUM_STAGES]) {
    if (!b || !ms) return fail(ORBX_ERR_ARG, "null argument");
    if (!b->timing) return fail(ORBX_ERR_ARG, "timing not enabled");
    DeviceGuard dg(b->device);
    CU(cudaEventSynchronize(b->ev[ORBX_NUM_STAGES]));
    for (int i = 0; i < ORBX_NUM_STAGES; ++i) CU(cudaEventElapsedTime(&ms[i], b->ev[i], b->ev[i + 1]));
    return ORBX_OK;
}
int orbx_batch_kernel_launches(const orbx_batch* b) { return b ? b->launches : ORBX_ERR_ARG; }

int orbx_batch_debug_candidates(orbx_batch* b, int image, int level, float* xyr, int capacity, int* n) {
    if (!b || level < 0 || level >= b->g.nlevels || image < 0 || image >= b->maxImages || !n) return fail(ORBX_ERR_ARG, "bad argument");
    DeviceGuard dg(b->device);
    const LevelGeom& L = b->g.lv[level];
    CU(cudaStreamSynchronize(b->stream));
    int cnt = 0;
    CU(cudaMemcpy(&cnt, b->buf.candCount + image * ORBX_MAX_LEVELS + level, sizeof(int), cudaMemcpyDeviceToHost));
    cnt = std::min(cnt, L.candCap);
    *n = cnt;
    if (!xyr || cnt == 0) return ORBX_OK;
    std::vector<uint32_t> c(cnt);
    CU(cudaMemcpy(c.data(), b->buf.cand + L.candOff + (size_t)image * L.candCap, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
    // the device keeps candidates unordered; present them in the reference's vector order
    std::sort(c.begin(), c.end(), [&](uint32_t a, uint32_t d) { return cand_order_key(a, L.wCell, L.hCell) < cand_order_key(d, L.wCell, L.hCell); });
    for (int i = 0; i < std::min(cnt, capacity); ++i) {
        xyr[3 * i] = (float)cand_x(c[i]); xyr[3 * i + 1] = (float)cand_y(c[i]); xyr[3 * i + 2] = (float)cand_r(c[i]);
    }
    return ORBX_OK;
}

int orbx_batch_debug_selected(orbx_batch* b, int image, int level, float* xyr, int capacity, int* n) {
    if (!b || level < 0 || level >= b->g.nlevels || image < 0 || image >= b->maxImages || !n) return fail(ORBX_ERR_ARG, "bad argument");
    DeviceGuard dg(b->device);
    const LevelGeom& L = b->g.lv[level];
    CU(cudaStreamSynchronize(b->stream));
    int cnt = 0;
    CU(cudaMemcpy(&cnt, b->buf.selCount + image * ORBX_MAX_LEVELS + level, sizeof(int), cudaMemcpyDeviceToHost));
    *n = cnt;
    if (!xyr || cnt == 0) return ORBX_OK;
    std::vector<uint32_t> c(cnt);
    CU(cudaMemcpy(c.data(), b->buf.sel + (size_t)image * b->g.K + L.selOff, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < std::min(cnt, capacity); ++i) {
        xyr[3 * i] = (float)cand_x(c[i]); xyr[3 * i + 1] = (float)cand_y(c[i]); xyr[3 * i + 2] = (float)cand_r(c[i]);
    }
    return ORBX_OK;
}

int orbx_batch_debug_fallback_cells(orbx_batch* b, int image, int level, int* n_fallback) {
    if (!b || level < 0 || level >= b->g.nlevels || image < 0 || image >= b->maxImages || !n_fallback) return fail(ORBX_ERR_ARG, "bad argument");
    DeviceGuard dg(b->device);
    const LevelGeom& L = b->g.lv[level];
    CU(cudaStreamSynchronize(b->stream));
    std::vector<uint8_t> f((size_t)L.nCols * L.nRows);
    CU(cudaMemcpy(f.data(), b->buf.cellFlags + (size_t)image * b->g.cellsPerImage + L.cellOff, f.size(), cudaMemcpyDeviceToHost));
    int c = 0;
    for (uint8_t v : f) c += (v == 2);
    *n_fallback = c;
    return ORBX_OK;
}

// ---- pinned host memory for callers that want overlapped copies ----
void* orbx_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { g_err = "cudaHostAlloc failed"; return nullptr; }
    return p;
}
void orbx_free_pinned(void* p) { if (p) cudaFreeHost(p); }

// ---- constants / host-logic hooks (no GPU needed) ----
int orbx_tables(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast, float* sf, float* isf, float* sig2, float* isig2,
                int* quota, int* umax) {
    Params p;
    int r = make_params(nfeatures, scale_factor, nlevels, ini_th_fast, min_th_fast, p);
    if (r) return r;
    for (int i = 0; i < nlevels; ++i) {
        if (sf) sf[i] = p.sf[i];
        if (isf) isf[i] = p.isf[i];
        if (sig2) sig2[i] = p.sig2[i];
        if (isig2) isig2[i] = p.isig2[i];
        if (quota) quota[i] = p.quota[i];
    }
    if (umax) std::memcpy(umax, p.umax, sizeof(p.umax));
    return ORBX_OK;
}

int orbx_level_geometry(int nfeatures, float scale_factor, int nlevels, int width, int height, int level, int out[8]) {
    Params p;
    int r = make_params(nfeatures, scale_factor, nlevels, 20, 7, p);
    if (r) return r;
    Geom g;
    if ((r = make_geom(p, width, height, 1, g))) return r;
    if (level < 0 || level >= nlevels || !out) return fail(ORBX_ERR_ARG, "bad level");
    const LevelGeom& L = g.lv[level];
    out[0] = L.w; out[1] = L.h; out[2] = L.nCols; out[3] = L.nRows; out[4] = L.wCell; out[5] = L.hCell; out[6] = L.nIni; out[7] = L.selCap;
    return ORBX_OK;
}

// sorts n (key, index) pairs exactly like libstdc++ std::sort with a key-only comparator; perm[i] = input index at slot i.
// on_device != 0 runs the same routine in a single CUDA thread (the way the quadtree kernel uses it).
int orbx_debug_sort(const uint32_t* keys, int n, int* perm, int force_depth, int on_device) {
    if (!keys || !perm || n < 0 || n > 65535) return fail(ORBX_ERR_ARG, "bad argument");
    std::vector<unsigned long long> e(n);
    for (int i = 0; i < n; ++i) e[i] = ((unsigned long long)keys[i] << 16) | (unsigned)i;
    if (on_device) {
        unsigned long long* d = nullptr;
        CU(cudaMalloc((void**)&d, std::max<size_t>(8, (size_t)n * 8)));
        cudaError_t er = cudaMemcpy(d, e.data(), (size_t)n * 8, cudaMemcpyHostToDevice);
        if (er == cudaSuccess) { launch_debug_sort_impl(d, n, force_depth, 0); er = cudaDeviceSynchronize(); }
        if (er == cudaSuccess) er = cudaMemcpy(e.data(), d, (size_t)n * 8, cudaMemcpyDeviceToHost);
        cudaFree(d);
        if (er != cudaSuccess) return cuda_fail(er, "device sort");
    } else
        StdSortEmul<16>::sort(e.data(), n, force_depth);
    for (int i = 0; i < n; ++i) perm[i] = (int)(e[i] & 0xFFFF);
    return ORBX_OK;
}

int orbx_debug_sincos(const float* angles_deg, float* out_sin, float* out_cos, int n) {
    float *d = nullptr;
    CU(cudaMalloc((void**)&d, std::max<size_t>(4, (size_t)n * 12)));
    cudaError_t er = cudaMemcpy(d, angles_deg, (size_t)n * 4, cudaMemcpyHostToDevice);
    if (er == cudaSuccess) { launch_debug_sincos(d, d + n, d + 2 * (size_t)n, n, 0); er = cudaDeviceSynchronize(); }
    if (er == cudaSuccess) er = cudaMemcpy(out_sin, d + n, (size_t)n * 4, cudaMemcpyDeviceToHost);
    if (er == cudaSuccess) er = cudaMemcpy(out_cos, d + 2 * (size_t)n, (size_t)n * 4, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (er != cudaSuccess) return cuda_fail(er, "device sincos");
    return ORBX_OK;
}

int orbx_debug_atan2(const float* y, const float* x, float* out, int n) {
    float *d = nullptr;
    CU(cudaMalloc((void**)&d, std::max<size_t>(4, (size_t)n * 12)));
    cudaError_t er = cudaMemcpy(d, y, (size_t)n * 4, cudaMemcpyHostToDevice);
    if (er == cudaSuccess) er = cudaMemcpy(d + n, x, (size_t)n * 4, cudaMemcpyHostToDevice);
    if (er == cudaSuccess) { launch_debug_atan2(d, d + n, d + 2 * (size_t)n, n, 0); er = cudaDeviceSynchronize(); }
    if (er == cudaSuccess) er = cudaMemcpy(out, d + 2 * (size_t)n, (size_t)n * 4, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (er != cudaSuccess) return cuda_fail(er, "device atan2");
    return ORBX_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// Multi-GPU pool (include/orbx.h): one host thread per device, pipelined batch handles, frames sharded by chunk
// ------------------------------------------------------------------------------------------------------------------
struct orbx_pool {
    std::vector<int> devices;
    std::vector<std::vector<orbx_batch*>> handles;   // [device][handle]
    std::vector<cudaEvent_t> evStart, evStop;        // per device, on its first handle's stream / after the join of all its handles
    std::vector<float> lastMs;
    int chunk = 0, K = 0, width = 0, height = 0, launches = 0;
};

extern "C" {

int orbx_pool_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast, const int* devices, int n_devices, int width,
                     int height, int chunk_images, int handles_per_device, orbx_pool** out) {
    if (!out) return fail(ORBX_ERR_ARG, "null out pointer");
    *out = nullptr;
    if (!devices || n_devices < 1 || n_devices > 64) return fail(ORBX_ERR_ARG, "need 1..64 device ordinals");
    if (chunk_images < 1 || handles_per_device < 1 || handles_per_device > 8) return fail(ORBX_ERR_ARG, "chunk_images >= 1 and 1 <= handles_per_device <= 8");
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return fail(ORBX_ERR_ARG, "a device appears twice");
    orbx_pool* p = new orbx_pool();
    p->devices.assign(devices, devices + n_devices);
    p->handles.resize((size_t)n_devices);
    p->evStart.assign((size_t)n_devices, nullptr);
    p->evStop.assign((size_t)n_devices, nullptr);
    p->lastMs.assign((size_t)n_devices, 0.f);
    p->chunk = chunk_images; p->width = width; p->height = height;
    for (int d = 0; d < n_devices; ++d) {
        for (int k = 0; k < handles_per_device; ++k) {
            orbx_batch* b = nullptr;
            const int r = orbx_batch_create(nfeatures, scale_factor, nlevels, ini_th_fast, min_th_fast, devices[d], width, height, chunk_images, &b);
            if (r != ORBX_OK) { const std::string msg = g_err; orbx_pool_destroy(p); return fail(r, msg); }
            p->handles[(size_t)d].push_back(b);
        }
        DeviceGuard dg(devices[d]);
        if (cudaEventCreate(&p->evStart[(size_t)d]) != cudaSuccess || cudaEventCreate(&p->evStop[(size_t)d]) != cudaSuccess) {
            orbx_pool_destroy(p);
            return fail(ORBX_ERR_CUDA, "cudaEventCreate failed");
        }
    }
    p->K = p->handles[0][0]->g.K;
    *out = p;
    return ORBX_OK;
}

void orbx_pool_destroy(orbx_pool* p) {
    if (!p) return;
    for (size_t d = 0; d < p->handles.size(); ++d) {
        for (orbx_batch* b : p->handles[d]) orbx_batch_destroy(b);
        DeviceGuard dg(p->devices[d]);
        if (p->evStart[d]) cudaEventDestroy(p->evStart[d]);
        if (p->evStop[d]) cudaEventDestroy(p->evStop[d]);
    }
    delete p;
}

int orbx_pool_max_keypoints(const orbx_pool* p) { return p ? p->K : ORBX_ERR_ARG; }
int orbx_pool_devices(const orbx_pool* p) { return p ? (int)p->devices.size() : ORBX_ERR_ARG; }
int orbx_pool_kernel_launches(const orbx_pool* p) { return p ? p->launches : ORBX_ERR_ARG; }
int orbx_pool_last_ms(const orbx_pool* p, float* ms) {
    if (!p || !ms) return fail(ORBX_ERR_ARG, "null argument");
    for (size_t d = 0; d < p->lastMs.size(); ++d) ms[d] = p->lastMs[d];
    return ORBX_OK;
}

int orbx_pool_extract(orbx_pool* p, const uint8_t* images, size_t row_stride, size_t image_stride, int n_images, const int* lapping,
                      orbx_keypoint* keypoints, uint8_t* descriptors, int* counts, int* mono_index) {
    if (!p) return fail(ORBX_ERR_ARG, "null pool");
    if (n_images < 0) return fail(ORBX_ERR_ARG, "negative image count");
    if (n_images == 0) return ORBX_OK;
    if (!images) return fail(ORBX_ERR_EMPTY, "null image");
    if (row_stride < (size_t)p->width || image_stride < row_stride * (size_t)p->height) return fail(ORBX_ERR_ARG, "strides smaller than the image");
    const int G = (int)p->devices.size(), chunk = p->chunk, nChunks = (n_images + chunk - 1) / chunk;
    const size_t K = (size_t)p->K;
    std::vector<int> rc((size_t)G, ORBX_OK), nl((size_t)G, 0);
    std::vector<std::string> err((size_t)G);
    // one host thread per device (the scaled-out form of the two extraction threads of Frame.cc:139-142): it only enqueues,
    // every copy and kernel is asynchronous on the handles' streams
    auto worker = [&](int d) {
        std::vector<orbx_batch*>& hs = p->handles[(size_t)d];
        const int nh = (int)hs.size();
        int r = ORBX_OK, mine = 0;
        DeviceGuard dg(p->devices[(size_t)d]);
        cudaEventRecord(p->evStart[(size_t)d], hs[0]->stream);
        for (int k = 1; k < nh; ++k) cudaStreamWaitEvent(hs[(size_t)k]->stream, p->evStart[(size_t)d], 0);   // nobody starts before the clock does
        for (int c = d; c < nChunks && r == ORBX_OK; c += G, ++mine) {
            orbx_batch* b = hs[(size_t)(mine % nh)];
            const int first = c * chunk, n = std::min(chunk, n_images - first);
            r = orbx_batch_upload(b, images + (size_t)first * image_stride, row_stride, image_stride, 0, n);
            if (r == ORBX_OK) r = orbx_batch_run(b, n, lapping ? lapping + 2 * (size_t)first : nullptr);
            if (r == ORBX_OK)
                r = orbx_batch_download_async(b, 0, n, keypoints ? keypoints + (size_t)first * K : nullptr, descriptors ? descriptors + (size_t)first * K * 32 : nullptr,
                                              counts ? counts + first : nullptr, mono_index ? mono_index + first : nullptr);
            nl[(size_t)d] += orbx_batch_kernel_launches(b);
        }
        // join the handles' streams on the first one, stop the clock there, wait
        for (int k = 1; k < nh; ++k) {
            cudaEventRecord(hs[(size_t)k]->evFork, hs[(size_t)k]->stream);
            cudaStreamWaitEvent(hs[0]->stream, hs[(size_t)k]->evFork, 0);
        }
        cudaEventRecord(p->evStop[(size_t)d], hs[0]->stream);
        for (int k = 0; k < nh; ++k) {
            const int rk = orbx_batch_check(hs[(size_t)k]);   // synchronises the handle's stream, reports device-side capacity errors
            if (r == ORBX_OK) r = rk;
        }
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p->evStart[(size_t)d], p->evStop[(size_t)d]) == cudaSuccess) p->lastMs[(size_t)d] = ms;
        rc[(size_t)d] = r;
        if (r != ORBX_OK) err[(size_t)d] = g_err;   // (g_err is thread local)
    };
    std::vector<std::thread> th;
    for (int d = 1; d < G; ++d) th.emplace_back(worker, d);
    worker(0);
    for (std::thread& t : th) t.join();
    p->launches = 0;
    for (int d = 0; d < G; ++d) p->launches += nl[(size_t)d];
    for (int d = 0; d < G; ++d)
        if (rc[(size_t)d] != ORBX_OK) return fail(rc[(size_t)d], "device " + std::to_string(p->devices[(size_t)d]) + ": " + err[(size_t)d]);
    return ORBX_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// Single-image handle == one TC2LI_SLAM::ORBextractor instance: a batch of one, created lazily for the image size
// ------------------------------------------------------------------------------------------------------------------
struct orbx_extractor {
    Params p;
    int device = 0;
    int taps[7] = {18, 34, 48, 56, 48, 34, 18};
    orbx_batch* b = nullptr;
    // pinned staging so that H2D / D2H are truly asynchronous with respect to the other handle of a stereo pair
    uint8_t* pinImg = nullptr; size_t pinImgBytes = 0;
    orbx_keypoint* pinKps = nullptr; uint8_t* pinDesc = nullptr; int* pinCnt = nullptr; size_t pinK = 0;
    // The whole per-image sequence (H2D, 20 kernels on two streams, D2H) is launch-bound for one image: after a first eager
    // call it is captured once into a CUDA graph and replayed; pinCnt[4..5] is the pinned lapping window the graph reads.
    cudaGraphExec_t exec = nullptr;
    int channels = 1, rgb = 0;  // input format of the last call (1 = grey); a change re-captures the graph
    // rectification (orbx_set_rectification): host copies of the maps, applied to the batch-of-one when it is (re)created
    std::vector<float> rectX, rectY;
    int rectW = 0, rectH = 0, srcW = 0, srcH = 0;   // fed (rectified / resized) size and raw size; rectW > 0: the extract calls take RAW images
    bool inResize = false;                          // raw images are resized (orbx_set_input_resize) instead of remapped
    bool rectDirty = false;
    unsigned tapsGeneration = 0;  // generation of the batch's taps the graph was captured with
    int graphState = 0;  // 0: next call runs eagerly (lazy allocations), 1: capture on the next call, 2: replay `exec`, 3: capture failed, stay eager
    // Where the results of the last call live: (vb, vi) = (b, 0) after a single-image call, (the pair batch of the left handle, 0 | 1) after
    // orbx_extract_stereo took the pair path; vCount / vMono = N and monoIndex of that call.  Everything that reads results goes through the view.
    orbx_batch* vb = nullptr;
    int vi = 0, vCount = 0, vMono = 0;
    // orbx_extract_stereo on two handles with identical parameters: ONE batch of two images, one graph, one launch sequence for the pair — the
    // launch sequence is latency bound (a pair costs what one image costs: profiles/tools/latency_probe.py).  Owned by the LEFT handle.
    struct Pair {
        orbx_batch* b = nullptr;
        uint8_t* pinImg = nullptr; size_t pinImgBytes = 0;
        orbx_keypoint* pinKps = nullptr; uint8_t* pinDesc = nullptr; int* pinCnt = nullptr; size_t pinK = 0;  // pinCnt: count[2], mono[2], err | [8..11] lapping
        cudaGraphExec_t exec = nullptr;
        int graphState = 0;
        int taps[7] = {0, 0, 0, 0, 0, 0, 0};
    } pair;
    orbx_extractor* peer = nullptr;   // the other handle of the pair whose view points into pair.b (cleared when either is destroyed)
};

namespace {

int extractor_prepare(orbx_extractor* h, int w, int hgt, int channels, int rgb) {
    if (h->b && (h->b->width != w || h->b->height != hgt)) {
        if (h->exec) { cudaGraphExecDestroy(h->exec); h->exec = nullptr; }
        h->graphState = 0;
        orbx_batch_destroy(h->b); h->b = nullptr;
    }
    if (h->channels != channels || h->rgb != rgb) {  // another input format: other copies and kernels, start over with an eager call
        if (h->exec) { cudaGraphExecDestroy(h->exec); h->exec = nullptr; }
        h->graphState = 0;
        h->channels = channels; h->rgb = rgb;
    }
    if (!h->b) {
        int r = orbx_batch_create(h->p.nfeatures, h->p.scaleFactorF, h->p.nlevels, h->p.iniTh, h->p.minTh, h->device, w, hgt, 1, &h->b);
        if (r) return r;
        orbx_batch_set_blur_taps(h->b, h->taps);
        h->b->lapStable = true;
        h->rectDirty = h->rectW > 0;
    }
    if (h->rectDirty) {  // new maps (or a new batch-of-one): other kernels and pointers, start over with an eager call
        if (h->exec) { cudaGraphExecDestroy(h->exec); h->exec = nullptr; }
        h->graphState = 0;
        int r = h->rectW > 0 ? (h->inResize ? orbx_batch_set_input_resize(h->b, h->srcW, h->srcH)
                                            : orbx_batch_set_rectification(h->b, h->rectX.data(), h->rectY.data(), (size_t)h->rectW * sizeof(float), h->srcW, h->srcH))
                             : orbx_batch_set_rectification(h->b, nullptr, nullptr, 0, 0, 0);
        if (r) return r;
        h->rectDirty = false;
    }
    DeviceGuard dg(h->device);
    const size_t need = h->rectW > 0 ? (size_t)h->srcW * h->srcH * channels : (size_t)w * hgt * channels;
    if (need > h->pinImgBytes) {
        if (h->pinImg) cudaFreeHost(h->pinImg);
        h->pinImg = nullptr; h->pinImgBytes = 0;
        CU(cudaHostAlloc((void**)&h->pinImg, need, cudaHostAllocDefault));
        h->pinImgBytes = need;
    }
    const size_t K = (size_t)h->b->g.K;
    if (K > h->pinK) {
        if (h->pinKps) cudaFreeHost(h->pinKps);
        if (h->pinDesc) cudaFreeHost(h->pinDesc);
        if (h->pinCnt) cudaFreeHost(h->pinCnt);
        h->pinKps = nullptr; h->pinDesc = nullptr; h->pinCnt = nullptr; h->pinK = 0;
        CU(cudaHostAlloc((void**)&h->pinKps, K * sizeof(orbx_keypoint), cudaHostAllocDefault));
        CU(cudaHostAlloc((void**)&h->pinDesc, K * 32, cudaHostAllocDefault));
        CU(cudaHostAlloc((void**)&h->pinCnt, 8 * sizeof(int), cudaHostAllocDefault));
        h->pinK = K;
    }
    return ORBX_OK;
}

// stage + enqueue everything for one image; nothing waits here
// everything one image needs, enqueued on the handle's streams (also the body of the captured graph)
int extractor_enqueue(orbx_extractor* h, int w, int hgt) {
    int r;
    if (h->rectW > 0) r = orbx_batch_upload_raw(h->b, h->pinImg, (size_t)h->srcW * h->channels, (size_t)h->srcW * h->channels * h->srcH, h->channels, h->rgb, 0, 1);
    else if (h->channels == 1) r = orbx_batch_upload(h->b, h->pinImg, (size_t)w, (size_t)w * hgt, 0, 1);
    else r = orbx_batch_upload_color(h->b, h->pinImg, (size_t)w * h->channels, (size_t)w * h->channels * hgt, h->channels, h->rgb, 0, 1);
    if (r) return r;
    if ((r = orbx_batch_run(h->b, 1, h->pinCnt + 4))) return r;
    orbx_batch* b = h->b;
    cudaStream_t s = b->stream;
    const size_t K = (size_t)b->g.K;
    CU(cudaMemcpyAsync(h->pinKps, b->buf.kps, K * sizeof(orbx_keypoint), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(h->pinDesc, b->buf.desc, K * 32, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(h->pinCnt, b->buf.outCount, 3 * sizeof(int), cudaMemcpyDeviceToHost, s));   // count, monoIndex, error flag (adjacent in a batch of one)
    return ORBX_OK;
}

// stage + enqueue everything for one image; nothing waits here
int extractor_issue(orbx_extractor* h, const uint8_t* image, int w, int hgt, size_t stride, int lap0, int lap1, int channels = 1, int rgb = 0) {
    // with rectification maps the caller hands in the RAW image (System.cc:238-248); the extractor works on the maps' size
    const int rawH = hgt;
    if (h->rectW > 0) {
        if (w != h->srcW || hgt != h->srcH) return fail(ORBX_ERR_ARG, "image size differs from the raw size given to orbx_set_rectification");
        w = h->rectW; hgt = h->rectH;
    }
    int r = extractor_prepare(h, w, hgt, channels, rgb);
    if (r) return r;
    const size_t rowBytes = (size_t)(h->rectW > 0 ? h->srcW : w) * channels;
    for (int y = 0; y < rawH; ++y) std::memcpy(h->pinImg + (size_t)y * rowBytes, image + (size_t)y * stride, rowBytes);
    h->pinCnt[4] = lap0; h->pinCnt[5] = lap1;
    DeviceGuard dg(h->device);
    orbx_batch* b = h->b;
    static const bool noGraph = std::getenv("ORBX_NO_GRAPH") != nullptr;
    if (noGraph || b->timing || b->stream != b->own) return extractor_enqueue(h, w, hgt);
    if (h->graphState == 2 && h->tapsGeneration != b->tapsGeneration) {  // taps changed through the batch handle: the graph is stale
        if (h->exec) { cudaGraphExecDestroy(h->exec); h->exec = nullptr; }
        h->graphState = 1;
    }
    h->tapsGeneration = b->tapsGeneration;
    if (h->graphState == 0) { h->graphState = 1; return extractor_enqueue(h, w, hgt); }  // first call eager: lazy allocations happen here
    if (h->graphState == 3) return extractor_enqueue(h, w, hgt);
    if (h->graphState == 1) {
        cudaGraph_t graph = nullptr;
        CU(cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeThreadLocal));
        r = extractor_enqueue(h, w, hgt);
        const cudaError_t e = cudaStreamEndCapture(b->stream, &graph);
        cudaError_t e2 = cudaErrorUnknown;
        if (r == ORBX_OK && e == cudaSuccess && graph) e2 = cudaGraphInstantiate(&h->exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (e2 != cudaSuccess) {
            h->exec = nullptr;
            h->graphState = 3;
            cudaGetLastError();
            return extractor_enqueue(h, w, hgt);
        }
        h->graphState = 2;
    }
    CU(cudaGraphLaunch(h->exec, b->stream));
    return ORBX_OK;
}

int extractor_collect(orbx_extractor* h, orbx_keypoint* kps, uint8_t* desc, int capacity, int* n, int* mono) {
    DeviceGuard dg(h->device);
    CU(cudaStreamSynchronize(h->b->stream));
    if (h->pinCnt[2]) {
        cudaMemsetAsync(h->b->buf.errFlag, 0, sizeof(int), h->b->stream);
        return fail(ORBX_ERR_CAPACITY, "device capacity exceeded");
    }
    const int N = h->pinCnt[0];
    h->vb = h->b; h->vi = 0; h->vCount = N; h->vMono = h->pinCnt[1];
    if (N > capacity) return fail(ORBX_ERR_CAPACITY, "output capacity smaller than the number of keypoints");
    if (kps) std::memcpy(kps, h->pinKps, (size_t)N * sizeof(orbx_keypoint));
    if (desc) std::memcpy(desc, h->pinDesc, (size_t)N * 32);
    if (n) *n = N;
    if (mono) *mono = h->pinCnt[1];
    return ORBX_OK;
}

// ---- the pair path of orbx_extract_stereo ----
bool pair_eligible(const orbx_extractor* l, const orbx_extractor* r) {
    static const bool off = std::getenv("ORBX_NO_PAIR") != nullptr;
    return !off && l->device == r->device && l->p.nfeatures == r->p.nfeatures && l->p.scaleFactorF == r->p.scaleFactorF && l->p.nlevels == r->p.nlevels &&
           l->p.iniTh == r->p.iniTh && l->p.minTh == r->p.minTh && std::memcmp(l->taps, r->taps, sizeof(l->taps)) == 0 && l->rectW == 0 && r->rectW == 0;
}

void pair_release(orbx_extractor* h) {
    orbx_extractor::Pair& P = h->pair;
    if (P.exec) { cudaGraphExecDestroy(P.exec); P.exec = nullptr; }
    if (P.b) { orbx_batch_destroy(P.b); P.b = nullptr; }
    if (P.pinImg) cudaFreeHost(P.pinImg);
    if (P.pinKps) cudaFreeHost(P.pinKps);
    if (P.pinDesc) cudaFreeHost(P.pinDesc);
    if (P.pinCnt) cudaFreeHost(P.pinCnt);
    P = orbx_extractor::Pair();
}

int pair_enqueue(orbx_extractor* left, int w, int hgt) {
    orbx_extractor::Pair& P = left->pair;
    int r = orbx_batch_upload(P.b, P.pinImg, (size_t)w, (size_t)w * hgt, 0, 2);
    if (r) return r;
    if ((r = orbx_batch_run(P.b, 2, P.pinCnt + 8))) return r;
    cudaStream_t s = P.b->stream;
    const size_t K = (size_t)P.b->g.K;
    CU(cudaMemcpyAsync(P.pinKps, P.b->buf.kps, 2 * K * sizeof(orbx_keypoint), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(P.pinDesc, P.b->buf.desc, 2 * K * 32, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(P.pinCnt, P.b->buf.outCount, 5 * sizeof(int), cudaMemcpyDeviceToHost, s));   // count[2], monoIndex[2], error flag (adjacent in a batch of two)
    return ORBX_OK;
}

int pair_extract(orbx_extractor* left, orbx_extractor* right, const uint8_t* imL, const uint8_t* imR, int w, int hgt, size_t strideL, size_t strideR, const int lapL[2],
                 const int lapR[2], orbx_keypoint* kpsL, uint8_t* descL, int capL, int* nL, int* monoL, orbx_keypoint* kpsR, uint8_t* descR, int capR, int* nR, int* monoR) {
    orbx_extractor::Pair& P = left->pair;
    DeviceGuard dg(left->device);
    if (P.b && (P.b->width != w || P.b->height != hgt)) {
        if (P.exec) { cudaGraphExecDestroy(P.exec); P.exec = nullptr; }
        orbx_batch_destroy(P.b); P.b = nullptr; P.graphState = 0;
        if (right->vb && right->peer == left) right->vb = nullptr;
        left->vb = nullptr;
    }
    if (!P.b) {
        int r = orbx_batch_create(left->p.nfeatures, left->p.scaleFactorF, left->p.nlevels, left->p.iniTh, left->p.minTh, left->device, w, hgt, 2, &P.b);
        if (r) return r;
        orbx_batch_set_blur_taps(P.b, left->taps);
        std::memcpy(P.taps, left->taps, sizeof(P.taps));
        P.b->lapStable = true;
        P.graphState = 0;
    }
    if (std::memcmp(P.taps, left->taps, sizeof(P.taps)) != 0) {   // taps changed since the graph was captured (they are baked into kernel arguments)
        orbx_batch_set_blur_taps(P.b, left->taps);
        std::memcpy(P.taps, left->taps, sizeof(P.taps));
        if (P.exec) { cudaGraphExecDestroy(P.exec); P.exec = nullptr; }
        if (P.graphState == 2) P.graphState = 1;
    }
    const size_t need = 2 * (size_t)w * hgt, K = (size_t)P.b->g.K;
    if (need > P.pinImgBytes) {
        if (P.pinImg) cudaFreeHost(P.pinImg);
        P.pinImg = nullptr; P.pinImgBytes = 0;
        CU(cudaHostAlloc((void**)&P.pinImg, need, cudaHostAllocDefault));
        P.pinImgBytes = need;
        if (P.exec) { cudaGraphExecDestroy(P.exec); P.exec = nullptr; }
        if (P.graphState == 2) P.graphState = 1;
    }
    if (K > P.pinK) {
        if (P.pinKps) cudaFreeHost(P.pinKps);
        if (P.pinDesc) cudaFreeHost(P.pinDesc);
        if (P.pinCnt) cudaFreeHost(P.pinCnt);
        P.pinKps = nullptr; P.pinDesc = nullptr; P.pinCnt = nullptr; P.pinK = 0;
        CU(cudaHostAlloc((void**)&P.pinKps, 2 * K * sizeof(orbx_keypoint), cudaHostAllocDefault));
        CU(cudaHostAlloc((void**)&P.pinDesc, 2 * K * 32, cudaHostAllocDefault));
        CU(cudaHostAlloc((void**)&P.pinCnt, 16 * sizeof(int), cudaHostAllocDefault));
        P.pinK = K;
        if (P.exec) { cudaGraphExecDestroy(P.exec); P.exec = nullptr; }
        if (P.graphState == 2) P.graphState = 1;
    }
    const size_t plane = (size_t)w * hgt;
    if (strideL == (size_t)w) std::memcpy(P.pinImg, imL, plane);
    else for (int y = 0; y < hgt; ++y) std::memcpy(P.pinImg + (size_t)y * w, imL + (size_t)y * strideL, (size_t)w);
    if (strideR == (size_t)w) std::memcpy(P.pinImg + plane, imR, plane);
    else for (int y = 0; y < hgt; ++y) std::memcpy(P.pinImg + plane + (size_t)y * w, imR + (size_t)y * strideR, (size_t)w);
    P.pinCnt[8] = lapL[0]; P.pinCnt[9] = lapL[1]; P.pinCnt[10] = lapR[0]; P.pinCnt[11] = lapR[1];
    orbx_batch* b = P.b;
    static const bool noGraph = std::getenv("ORBX_NO_GRAPH") != nullptr;
    int r = ORBX_OK;
    if (noGraph || b->timing || b->stream != b->own || P.graphState == 0 || P.graphState == 3) {
        if (P.graphState == 0) P.graphState = 1;   // first call eager: lazy allocations happen here
        r = pair_enqueue(left, w, hgt);
    } else {
        if (P.graphState == 1) {
            cudaGraph_t graph = nullptr;
            CU(cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeThreadLocal));
            r = pair_enqueue(left, w, hgt);
            const cudaError_t e = cudaStreamEndCapture(b->stream, &graph);
            cudaError_t e2 = cudaErrorUnknown;
            if (r == ORBX_OK && e == cudaSuccess && graph) e2 = cudaGraphInstantiate(&P.exec, graph, 0);
            if (graph) cudaGraphDestroy(graph);
            if (e2 != cudaSuccess) {
                P.exec = nullptr;
                P.graphState = 3;
                cudaGetLastError();
                r = pair_enqueue(left, w, hgt);
            } else {
                P.graphState = 2;
                r = ORBX_OK;
            }
        }
        if (P.graphState == 2) CU(cudaGraphLaunch(P.exec, b->stream));
    }
    if (r) return r;
    CU(cudaStreamSynchronize(b->stream));
    if (P.pinCnt[4]) {
        cudaMemsetAsync(b->buf.errFlag, 0, sizeof(int), b->stream);
        return fail(ORBX_ERR_CAPACITY, "device capacity exceeded");
    }
    // the views of both handles now point into the pair batch
    if (left->peer && left->peer != right) { left->peer->peer = nullptr; if (left->peer->vb == b) left->peer->vb = nullptr; }
    if (right->peer && right->peer != left) { right->peer->peer = nullptr; }
    left->peer = right; right->peer = left;
    left->vb = b; left->vi = 0; left->vCount = P.pinCnt[0]; left->vMono = P.pinCnt[2];
    right->vb = b; right->vi = 1; right->vCount = P.pinCnt[1]; right->vMono = P.pinCnt[3];
    if (nL) *nL = P.pinCnt[0];
    if (monoL) *monoL = P.pinCnt[2];
    if (nR) *nR = P.pinCnt[1];
    if (monoR) *monoR = P.pinCnt[3];
    if (P.pinCnt[0] > capL || P.pinCnt[1] > capR) return fail(ORBX_ERR_CAPACITY, "output capacity smaller than the number of keypoints");
    if (kpsL) std::memcpy(kpsL, P.pinKps, (size_t)P.pinCnt[0] * sizeof(orbx_keypoint));
    if (descL) std::memcpy(descL, P.pinDesc, (size_t)P.pinCnt[0] * 32);
    if (kpsR) std::memcpy(kpsR, P.pinKps + K, (size_t)P.pinCnt[1] * sizeof(orbx_keypoint));
    if (descR) std::memcpy(descR, P.pinDesc + K * 32, (size_t)P.pinCnt[1] * 32);
    return ORBX_OK;
}

}  // namespace

extern "C" {

int orbx_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast, int device, orbx_extractor** out) {
    if (!out) return fail(ORBX_ERR_ARG, "null out pointer");
    *out = nullptr;
    orbx_extractor* h = new orbx_extractor();
    int r = make_params(nfeatures, scale_factor, nlevels, ini_th_fast, min_th_fast, h->p);
    if (r) { delete h; return r; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) {
        delete h;
        return fail(ORBX_ERR_CUDA, e != cudaSuccess ? std::string("no usable CUDA device: ") + cudaGetErrorString(e)
                                                    : std::string("CUDA device ordinal out of range (there is no CPU fallback)"));
    }
    h->device = device;
    *out = h;
    return ORBX_OK;
}

void orbx_destroy(orbx_extractor* h) {
    if (!h) return;
    if (h->peer) {   // the other handle of a pair must not keep a view into a batch that goes away with this one
        if (h->peer->vb == h->pair.b || h->peer->vb == h->vb) h->peer->vb = nullptr;
        h->peer->peer = nullptr;
    }
    pair_release(h);
    if (h->exec) cudaGraphExecDestroy(h->exec);
    if (h->b) orbx_batch_destroy(h->b);
    if (h->pinImg) cudaFreeHost(h->pinImg);
    if (h->pinKps) cudaFreeHost(h->pinKps);
    if (h->pinDesc) cudaFreeHost(h->pinDesc);
    if (h->pinCnt) cudaFreeHost(h->pinCnt);
    delete h;
}

int orbx_get_levels(const orbx_extractor* h) { return h ? h->p.nlevels : ORBX_ERR_ARG; }
float orbx_get_scale_factor(const orbx_extractor* h) { return h ? h->p.scaleFactorF : 0.f; }
static int copy_table(const orbx_extractor* h, const std::vector<float> Params::*m, float* out) {
    if (!h || !out) return fail(ORBX_ERR_ARG, "null argument");
    std::memcpy(out, (h->p.*m).data(), (size_t)h->p.nlevels * sizeof(float));
    return ORBX_OK;
}
int orbx_get_scale_factors(const orbx_extractor* h, float* out) { return copy_table(h, &Params::sf, out); }
int orbx_get_inverse_scale_factors(const orbx_extractor* h, float* out) { return copy_table(h, &Params::isf, out); }
int orbx_get_scale_sigma_squares(const orbx_extractor* h, float* out) { return copy_table(h, &Params::sig2, out); }
int orbx_get_inverse_scale_sigma_squares(const orbx_extractor* h, float* out) { return copy_table(h, &Params::isig2, out); }
int orbx_get_features_per_level(const orbx_extractor* h, int* out) {
    if (!h || !out) return fail(ORBX_ERR_ARG, "null argument");
    std::memcpy(out, h->p.quota.data(), (size_t)h->p.nlevels * sizeof(int));
    return ORBX_OK;
}
int orbx_max_keypoints(const orbx_extractor* h) {
    // image-size independent bound: per level max(quota + 3, 4 * nIni) with nIni <= ORBX_MAX_DIM / 35
    if (!h) return ORBX_ERR_ARG;
    if (h->b) return h->b->g.K;
    int k = 0;
    for (int q : h->p.quota) k += std::max(q + 3, 4 * (ORBX_MAX_DIM / 35));
    return k;
}
int orbx_set_blur_taps(orbx_extractor* h, const int taps[7]) {
    if (!h || !taps) return fail(ORBX_ERR_ARG, "null argument");
    {
        int sum = 0;
        for (int i = 0; i < 7; ++i) { if (taps[i] < 0 || taps[i] > 255) return fail(ORBX_ERR_ARG, "tap outside [0,255] (the kernel packs taps into bytes)"); sum += taps[i]; }
        if (sum > 257) return fail(ORBX_ERR_ARG, "taps must sum to about 256 (8.8 fixed point)");
    }
    std::memcpy(h->taps, taps, sizeof(h->taps));
    if (h->exec) { cudaGraphExecDestroy(h->exec); h->exec = nullptr; }  // the taps are baked into the captured kernel arguments
    if (h->graphState == 2) h->graphState = 1;
    if (h->b) return orbx_batch_set_blur_taps(h->b, taps);
    return ORBX_OK;
}

int orbx_extract(orbx_extractor* h, const uint8_t* image, int width, int height, size_t stride, int lap0, int lap1, orbx_keypoint* keypoints,
                 uint8_t* descriptors, int capacity, int* n_keypoints, int* mono_index) {
    if (!h) return fail(ORBX_ERR_ARG, "null handle");
    if (n_keypoints) *n_keypoints = 0;
    if (mono_index) *mono_index = 0;
    if (!image || width <= 0 || height <= 0) return fail(ORBX_ERR_EMPTY, "empty image");  // ORBextractor.cc:1063-1064
    if (stride < (size_t)width) return fail(ORBX_ERR_ARG, "row stride smaller than the width");
    int r = extractor_issue(h, image, width, height, stride, lap0, lap1);
    if (r) return r;
    return extractor_collect(h, keypoints, descriptors, capacity, n_keypoints, mono_index);
}

int orbx_extract_stereo(orbx_extractor* left, orbx_extractor* right, const uint8_t* image_left, const uint8_t* image_right, int width, int height,
                        size_t stride_left, size_t stride_right, const int lap_left[2], const int lap_right[2], orbx_keypoint* kps_left,
                        uint8_t* desc_left, int cap_left, int* n_left, int* mono_left, orbx_keypoint* kps_right, uint8_t* desc_right, int cap_right,
                        int* n_right, int* mono_right) {
    if (!left || !right || left == right) return fail(ORBX_ERR_ARG, "need two distinct handles");
    if (!image_left || !image_right || width <= 0 || height <= 0) return fail(ORBX_ERR_EMPTY, "empty image");
    if (stride_left < (size_t)width || stride_right < (size_t)width) return fail(ORBX_ERR_ARG, "row stride smaller than the width");
    const int zero[2] = {0, 0};
    if (!lap_left) lap_left = zero;
    if (!lap_right) lap_right = zero;
    if (n_left) *n_left = 0;
    if (n_right) *n_right = 0;
    if (pair_eligible(left, right))
        return pair_extract(left, right, image_left, image_right, width, height, stride_left, stride_right, lap_left, lap_right, kps_left, desc_left, cap_left, n_left,
                            mono_left, kps_right, desc_right, cap_right, n_right, mono_right);
    int r = extractor_issue(left, image_left, width, height, stride_left, lap_left[0], lap_left[1]);
    if (r) return r;
    if ((r = extractor_issue(right, image_right, width, height, stride_right, lap_right[0], lap_right[1]))) return r;
    r = extractor_collect(left, kps_left, desc_left, cap_left, n_left, mono_left);
    const int r2 = extractor_collect(right, kps_right, desc_right, cap_right, n_right, mono_right);
    return r ? r : r2;
}

// System::TrackStereoLidar's cv::remap (System.cc:238-248) in front of this extractor: map_x / map_y are the CV_32FC1 maps of the
// camera this extractor serves (Settings::M1l()/M2l() or M1r()/M2r()), width x height floats; the orbx_extract* calls then take
// the RAW src_width x src_height image.  NULL maps switch rectification off.
int orbx_set_rectification(orbx_extractor* h, const float* map_x, const float* map_y, int width, int height, size_t map_stride_bytes, int src_width,
                           int src_height) {
    if (!h) return fail(ORBX_ERR_ARG, "null handle");
    h->inResize = false;
    if (!map_x || !map_y) {
        h->rectW = h->rectH = h->srcW = h->srcH = 0;
        h->rectX.clear(); h->rectY.clear();
        h->rectDirty = true;
        return ORBX_OK;
    }
    if (width <= 0 || height <= 0 || src_width <= 0 || src_height <= 0 || width > ORBX_MAX_DIM || height > ORBX_MAX_DIM || src_width > ORBX_MAX_DIM ||
        src_height > ORBX_MAX_DIM)
        return fail(ORBX_ERR_ARG, "image size out of range");
    if (map_stride_bytes < (size_t)width * sizeof(float)) return fail(ORBX_ERR_ARG, "map stride smaller than width * sizeof(float)");
    h->rectX.resize((size_t)width * height);
    h->rectY.resize((size_t)width * height);
    for (int y = 0; y < height; ++y) {
        std::memcpy(h->rectX.data() + (size_t)y * width, (const char*)map_x + (size_t)y * map_stride_bytes, (size_t)width * sizeof(float));
        std::memcpy(h->rectY.data() + (size_t)y * width, (const char*)map_y + (size_t)y * map_stride_bytes, (size_t)width * sizeof(float));
    }
    h->rectW = width; h->rectH = height; h->srcW = src_width; h->srcH = src_height;
    h->rectDirty = true;
    return ORBX_OK;
}

// System::TrackStereoLidar's cv::resize(im, imToFeed, settings_->newImSize()) (System.cc:250-252) in front of this extractor: the orbx_extract*
// calls then take the RAW src_width x src_height image and feed width x height.  width == 0 switches it off.  Replaces a rectification.
int orbx_set_input_resize(orbx_extractor* h, int width, int height, int src_width, int src_height) {
    if (!h) return fail(ORBX_ERR_ARG, "null handle");
    h->rectX.clear(); h->rectY.clear();
    if (width == 0 && height == 0) {
        h->rectW = h->rectH = h->srcW = h->srcH = 0;
        h->inResize = false;
        h->rectDirty = true;
        return ORBX_OK;
    }
    if (width <= 0 || height <= 0 || src_width <= 1 || src_height <= 1 || width > ORBX_MAX_DIM || height > ORBX_MAX_DIM || src_width > ORBX_MAX_DIM ||
        src_height > ORBX_MAX_DIM)
        return fail(ORBX_ERR_ARG, "image size out of range");
    h->rectW = width; h->rectH = height; h->srcW = src_width; h->srcH = src_height;
    h->inResize = true;
    h->rectDirty = true;
    return ORBX_OK;
}

// the same with the caller's cvtColor (Tracking.cc:1646-1671) done on the device: interleaved 8-bit colour input
int orbx_extract_color(orbx_extractor* h, const uint8_t* image, int width, int height, size_t stride, int channels, int rgb, int lap0, int lap1,
                       orbx_keypoint* keypoints, uint8_t* descriptors, int capacity, int* n_keypoints, int* mono_index) {
    if (!h) return fail(ORBX_ERR_ARG, "null handle");
    if (n_keypoints) *n_keypoints = 0;
    if (mono_index) *mono_index = 0;
    if (!image || width <= 0 || height <= 0) return fail(ORBX_ERR_EMPTY, "empty image");
    if (channels != 3 && channels != 4) return fail(ORBX_ERR_ARG, "colour images have 3 or 4 interleaved channels");
    if (stride < (size_t)width * channels) return fail(ORBX_ERR_ARG, "row stride smaller than width * channels");
    int r = extractor_issue(h, image, width, height, stride, lap0, lap1, channels, rgb != 0);
    if (r) return r;
    return extractor_collect(h, keypoints, descriptors, capacity, n_keypoints, mono_index);
}

int orbx_extract_stereo_color(orbx_extractor* left, orbx_extractor* right, const uint8_t* image_left, const uint8_t* image_right, int width, int height,
                              size_t stride_left, size_t stride_right, int channels, int rgb, const int lap_left[2], const int lap_right[2],
                              orbx_keypoint* kps_left, uint8_t* desc_left, int cap_left, int* n_left, int* mono_left, orbx_keypoint* kps_right,
                              uint8_t* desc_right, int cap_right, int* n_right, int* mono_right) {
    if (!left || !right || left == right) return fail(ORBX_ERR_ARG, "need two distinct handles");
    if (!image_left || !image_right || width <= 0 || height <= 0) return fail(ORBX_ERR_EMPTY, "empty image");
    if (channels != 3 && channels != 4) return fail(ORBX_ERR_ARG, "colour images have 3 or 4 interleaved channels");
    if (stride_left < (size_t)width * channels || stride_right < (size_t)width * channels) return fail(ORBX_ERR_ARG, "row stride smaller than width * channels");
    const int zero[2] = {0, 0};
    if (!lap_left) lap_left = zero;
    if (!lap_right) lap_right = zero;
    int r = extractor_issue(left, image_left, width, height, stride_left, lap_left[0], lap_left[1], channels, rgb != 0);
    if (r) return r;
    if ((r = extractor_issue(right, image_right, width, height, stride_right, lap_right[0], lap_right[1], channels, rgb != 0))) return r;
    r = extractor_collect(left, kps_left, desc_left, cap_left, n_left, mono_left);
    const int r2 = extractor_collect(right, kps_right, desc_right, cap_right, n_right, mono_right);
    return r ? r : r2;
}

// replaces Frame::ComputeStereoMatches (Frame.cc:841-1011) for the stereo pair last extracted by `left` and `right`
// (mpORBextractorLeft / mpORBextractorRight): uright[N_left] / depth[N_left] = Frame::mvuRight / mvDepth, -1 = no match.
int orbx_stereo_match(orbx_extractor* left, orbx_extractor* right, float mb, float mbf, float* uright, float* depth, int capacity, int* n_left) {
    if (!left || !right || left == right || !left->vb || !right->vb) return fail(ORBX_ERR_ARG, "need two distinct handles that have extracted an image");
    orbx_batch *bl = left->vb, *br = right->vb;
    if (bl->device != br->device || bl->width != br->width || bl->height != br->height || bl->g.K != br->g.K || bl->g.nlevels != br->g.nlevels)
        return fail(ORBX_ERR_ARG, "left and right extractors differ in image size, parameters or device");
    if (!(mb > 0.f) || !(mbf > 0.f)) return fail(ORBX_ERR_ARG, "mb and mbf must be positive");
    if (bl->p.scaleFactorF > 1.9f) return fail(ORBX_ERR_ARG, "stereo matching needs scaleFactor <= 1.9 (beyond it the reference's SAD window, Frame.cc:950, leaves the image and cv::Mat::colRange asserts)");
    if (stereo_smem_bytes(bl->g.K, bl->g.lv[0].h) > 220 * 1024) return fail(ORBX_ERR_TOO_LARGE, "keypoint capacity too large for the stereo matcher's shared memory");
    DeviceGuard dg(bl->device);
    if (br != bl) CU(cudaStreamSynchronize(br->stream));  // the right image's results must be complete (the left stream orders its own)
    Buffers l = bl->buf, r = br->buf;
    l.imgBase = 0; r.imgBase = 0;
    launch_stereo_match(bl->g, l, r, 1, 0, left->vi, 0, right->vi, mb, mbf, bl->stream);
    CU(cudaGetLastError());
    const int N = left->vCount;
    const size_t lo = (size_t)left->vi * bl->g.K;
    if (n_left) *n_left = N;
    if (N > capacity) return fail(ORBX_ERR_CAPACITY, "output capacity smaller than the number of left keypoints");
    if (uright && N) CU(cudaMemcpyAsync(uright, l.uRight + lo, (size_t)N * 4, cudaMemcpyDeviceToHost, bl->stream));
    if (depth && N) CU(cudaMemcpyAsync(depth, l.depth + lo, (size_t)N * 4, cudaMemcpyDeviceToHost, bl->stream));
    CU(cudaStreamSynchronize(bl->stream));
    return ORBX_OK;
}

// descriptor stage of Frame::ComputeStereoFishEyeMatches (Frame.cc:1162-1187) for the pair last extracted by `left` and `right`:
// knn[q] / ratio_ok[q] for the *n_query rows of the left lapping tail (query q = left keypoint mono_left + q, train index t =
// right keypoint mono_right + t), i.e. BFmatcher.knnMatch(stereoDescLeft, stereoDescRight, matches, 2) and Lowe's gate.
int orbx_lapping_match(orbx_extractor* left, orbx_extractor* right, orbx_knn2* knn, uint8_t* ratio_ok, int capacity, int* n_query, int* n_train) {
    if (!left || !right || left == right || !left->vb || !right->vb) return fail(ORBX_ERR_ARG, "need two distinct handles that have extracted an image");
    orbx_batch *bl = left->vb, *br = right->vb;
    if (bl->device != br->device || bl->g.K != br->g.K) return fail(ORBX_ERR_ARG, "left and right extractors differ in parameters or device");
    if (bl->g.K > 65535) return fail(ORBX_ERR_TOO_LARGE, "keypoint capacity above 65535: train indices do not fit the matcher's 16-bit field");
    DeviceGuard dg(bl->device);
    if (br != bl) CU(cudaStreamSynchronize(br->stream));  // the right image's results must be complete (the left stream orders its own)
    Buffers l = bl->buf, r = br->buf;
    l.imgBase = 0; r.imgBase = 0;
    launch_lap_match(bl->g, l, r, 1, 0, left->vi, 0, right->vi, bl->stream);
    CU(cudaGetLastError());
    const int nq = left->vCount - left->vMono, nt = right->vCount - right->vMono;
    const size_t lo = (size_t)left->vi * bl->g.K;
    if (n_query) *n_query = nq;
    if (n_train) *n_train = nt;
    if (nq > capacity) return fail(ORBX_ERR_CAPACITY, "output capacity smaller than the number of left lapping keypoints");
    if (knn && nq) CU(cudaMemcpyAsync(knn, l.lapKnn + lo, (size_t)nq * 16, cudaMemcpyDeviceToHost, bl->stream));
    if (ratio_ok && nq) CU(cudaMemcpyAsync(ratio_ok, l.lapOk + lo, (size_t)nq, cudaMemcpyDeviceToHost, bl->stream));
    CU(cudaStreamSynchronize(bl->stream));
    return ORBX_OK;
}

int orbx_get_level_size(const orbx_extractor* h, int level, int* width, int* height) {
    if (!h || !h->vb) return fail(ORBX_ERR_ARG, "no image has been extracted yet");
    return orbx_batch_get_level_size(h->vb, level, width, height);
}
int orbx_get_level(orbx_extractor* h, int level, uint8_t* dst, size_t dst_stride, int border) {
    if (!h || !h->vb) return fail(ORBX_ERR_ARG, "no image has been extracted yet");
    return orbx_batch_get_level(h->vb, h->vi, level, dst, dst_stride, border);
}
orbx_batch* orbx_extractor_batch(orbx_extractor* h) { return h ? (h->vb ? h->vb : h->b) : nullptr; }

}  // extern "C"
