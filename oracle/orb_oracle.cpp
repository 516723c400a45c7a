// orb_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A from-scratch CPU restatement of TC2LI_SLAM::ORBextractor (reference:
// slam_framework/src/ORBextractor.cc, include/ORBextractor.h) together with the OpenCV / libstdc++ /
// glibc primitives it calls (restated from their published semantics; OpenCV is
// NOT vendored by the reference and its C++ SDK is absent from this image).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this
// library; it is the checker and the timed CPU baseline, never a product code path.  The product
// (tc2li_slam_b200/csrc) shares no code with this file.
//
// Pinning: the reference ships no tests or golden vectors for this path ("parity unpinned" by the
// reference itself).  This oracle is pinned instead by (1) bit-exact agreement of every primitive with
// cv2 4.13 (tests/test_oracle_vs_cv2.py), (2) the six known-answer cases of SURVEY.md Appendix C
// (tests/test_oracle_known_answers.py), (3) whole-pipeline agreement with the *verbatim* reference
// translation unit compiled against stand-in OpenCV headers (oracle/_ref, tests/test_oracle_vs_ref.py).
//
// Build: g++ -O3 -ffp-contract=off (no FMA contraction: the canonical arithmetic), libstdc++ std::sort,
// glibc sinf/cosf.

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <cstdio>
#include <map>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// A.0 integer rounding helpers (cvRound = round-half-to-even via the current rounding mode)
// ---------------------------------------------------------------------------------------------
inline int cv_round(float v) { return (int)lrintf(v); }
inline int cv_round(double v) { return (int)lrint(v); }
inline short sat_short(int v) { return (short)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

const int kEdge = 19;       // EDGE_THRESHOLD, ORBextractor.cc:47
const int kHalfPatch = 15;  // HALF_PATCH_SIZE, ORBextractor.cc:46
const int kPatch = 31;      // PATCH_SIZE, ORBextractor.cc:45

static const signed char kPattern[1024] = {
#include "orb_pattern.inc"
};

struct Plane {
    int w = 0, h = 0;
    std::vector<uint8_t> d;
    void alloc(int w_, int h_) { w = w_; h = h_; d.assign((size_t)w * h, 0); }
    uint8_t* row(int y) { return d.data() + (size_t)y * w; }
    const uint8_t* row(int y) const { return d.data() + (size_t)y * w; }
};

struct KeyPoint {  // layout-identical to cv::KeyPoint (28 bytes)
    float x, y, size, angle, response;
    int32_t octave, class_id;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct Cand {  // FAST candidate relative to (minBorderX, minBorderY), ORBextractor.cc:837-842
    float x, y, response;
};

// ---------------------------------------------------------------------------------------------
// Synthetic input generator orbx-synth-v1 — pure integer arithmetic.
// ---------------------------------------------------------------------------------------------
inline uint32_t lowbias32(uint32_t v) {
    v ^= v >> 16; v *= 0x7feb352dU; v ^= v >> 15; v *= 0x846ca68bU; v ^= v >> 16;
    return v;
}
inline uint32_t lat(uint32_t ix, uint32_t iy, uint32_t S, uint32_t seed) {
    return lowbias32((ix * 73856093U) ^ (iy * 19349663U) ^ (S * 83492791U) ^ (seed * 2654435761U)) & 255U;
}
inline uint32_t octv(uint32_t x, uint32_t y, uint32_t S, uint32_t seed) {
    uint32_t ix = x / S, iy = y / S, fx = x % S, fy = y % S;
    uint64_t acc = (uint64_t)(S - fx) * (S - fy) * lat(ix, iy, S, seed) + (uint64_t)fx * (S - fy) * lat(ix + 1, iy, S, seed) +
                   (uint64_t)(S - fx) * fy * lat(ix, iy + 1, S, seed) + (uint64_t)fx * fy * lat(ix + 1, iy + 1, S, seed);
    return (uint32_t)(acc / ((uint64_t)S * S));
}
void synth_v1(uint8_t* out, int W, int H, uint32_t seed, int xoff, uint32_t noise_xor) {
    for (int yy = 0; yy < H; ++yy)
        for (int xx = 0; xx < W; ++xx) {
            uint32_t x = (uint32_t)(xx + xoff), y = (uint32_t)yy;
            uint32_t base = (4 * octv(x, y, 64, seed) + 3 * octv(x, y, 32, seed) + 2 * octv(x, y, 16, seed) + octv(x, y, 8, seed)) / 10;
            uint32_t blk = lat(x / 24, y / 18, 999, seed);
            uint32_t noise = lowbias32((x * 0x9E3779B1U) ^ (y * 0x85EBCA77U) ^ ((seed ^ noise_xor) * 0xC2B2AE3DU)) & 7U;
            uint32_t p = (3 * base + 2 * blk) / 5 + noise;
            out[(size_t)yy * W + xx] = (uint8_t)(p > 255 ? 255 : p);
        }
}

// ---------------------------------------------------------------------------------------------
// A.1 cv::resize(INTER_LINEAR) for 8UC1 (called at ORBextractor.cc:1156).  Fixed-point bilinear with
// 11-bit coefficients; the vertical pass drops 4 bits before the 16-bit multiply.
// ---------------------------------------------------------------------------------------------
void resize_linear_8u(const uint8_t* src, int sw, int sh, size_t sstride, uint8_t* dst, int dw, int dh, size_t dstride) {
    // cv::resize derives inv_scale = dsize/ssize when dsize is given, and hal::resize then uses scale = 1/inv_scale.
    const double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> ax(2 * (size_t)dw), by(2 * (size_t)dh);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)std::floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ax[2 * dx] = sat_short(cv_round((1.f - fx) * 2048.f));
        ax[2 * dx + 1] = sat_short(cv_round(fx * 2048.f));
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)std::floor(fy);
        fy -= sy;
        yofs[dy] = sy;
        by[2 * dy] = sat_short(cv_round((1.f - fy) * 2048.f));
        by[2 * dy + 1] = sat_short(cv_round(fy * 2048.f));
    }
    std::vector<int> r0(dw), r1(dw);
    int have0 = -1000000, have1 = -1000000;
    auto hpass = [&](int sy, std::vector<int>& out) {
        const uint8_t* S = src + (size_t)sy * sstride;
        for (int dx = 0; dx < dw; ++dx) {
            int sx = xofs[dx];
            int sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
            out[dx] = S[sx] * ax[2 * dx] + S[sx1] * ax[2 * dx + 1];
        }
    };
    for (int dy = 0; dy < dh; ++dy) {
        int sy0 = std::min(std::max(yofs[dy], 0), sh - 1);
        int sy1 = std::min(std::max(yofs[dy] + 1, 0), sh - 1);
        if (sy0 == have1) { r0.swap(r1); std::swap(have0, have1); }
        if (sy0 != have0) { hpass(sy0, r0); have0 = sy0; }
        if (sy1 == have0) { r1 = r0; have1 = sy1; }
        else if (sy1 != have1) { hpass(sy1, r1); have1 = sy1; }
        const int b0 = by[2 * dy], b1 = by[2 * dy + 1];
        uint8_t* D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; ++dx) {
            int v = (((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2;
            D[dx] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
}

inline int reflect101(int i, int n) {  // A.2 BORDER_REFLECT_101
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * (n - 1) - i;
    }
    return i;
}

// ---------------------------------------------------------------------------------------------
// A.3 GaussianBlur(7x7, sigma 2) on 8U (ORBextractor.cc:1106): separable 8.8 fixed-point filter.
// ---------------------------------------------------------------------------------------------
void gaussian_blur7_8u(const uint8_t* src, int w, int h, size_t sstride, uint8_t* dst, size_t dstride, const int taps[7]) {
    std::vector<uint16_t> hbuf((size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* S = src + (size_t)y * sstride;
        uint16_t* Hh = hbuf.data() + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 0;
            if (x >= 3 && x + 3 < w)
                for (int k = 0; k < 7; ++k) acc += (uint32_t)taps[k] * S[x + k - 3];
            else
                for (int k = 0; k < 7; ++k) acc += (uint32_t)taps[k] * S[reflect101(x + k - 3, w)];
            Hh[x] = (uint16_t)(acc > 65535 ? 65535 : acc);
        }
    }
    for (int y = 0; y < h; ++y) {
        const uint16_t* R[7];
        for (int k = 0; k < 7; ++k) R[k] = hbuf.data() + (size_t)reflect101(y + k - 3, h) * w;
        uint8_t* D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 0;
            for (int k = 0; k < 7; ++k) acc += (uint32_t)taps[k] * R[k][x];
            uint32_t v = (acc + 32768U) >> 16;
            D[x] = (uint8_t)(v > 255 ? 255 : v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// A.4 cv::FAST(img, kps, t, true) TYPE_9_16 on a w x h crop (ORBextractor.cc:800,819).
// Output in raster order: (x, y, score) with score = (largest t' at which the pixel is still a corner),
// i.e. max over the 16 arcs of 9 of min |ring - v| on one side, minus 1.
// ---------------------------------------------------------------------------------------------
static const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

inline bool has_arc9(uint32_t m) {  // 9 contiguous set bits in a circular 16-bit mask
    m |= m << 16;
    m &= m >> 1;  // runs of >=2
    m &= m >> 2;  // >=4
    m &= m >> 4;  // >=8
    m &= m >> 1;  // >=9
    return (m & 0xFFFFu) != 0;
}

inline int arc_score(const int d[16]) {  // max over starts of min over 9 contiguous of d
    int m2[16], m4[16], m8[16], best = -1000;
    for (int i = 0; i < 16; ++i) m2[i] = std::min(d[i], d[(i + 1) & 15]);
    for (int i = 0; i < 16; ++i) m4[i] = std::min(m2[i], m2[(i + 2) & 15]);
    for (int i = 0; i < 16; ++i) m8[i] = std::min(m4[i], m4[(i + 4) & 15]);
    for (int i = 0; i < 16; ++i) best = std::max(best, std::min(m8[i], d[(i + 8) & 15]));
    return best;
}

struct FastKp { int x, y, score; };

void fast9_nms(const uint8_t* img, int w, int h, size_t stride, int t, std::vector<FastKp>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    int off[16];
    for (int k = 0; k < 16; ++k) off[k] = kRingDy[k] * (int)stride + kRingDx[k];
    // score rows with a zero rim; three-row rolling window for the 3x3 non-max suppression
    std::vector<int> buf((size_t)3 * w, 0);
    auto score_row = [&](int y, int* sc) {
        std::fill(sc, sc + w, 0);
        if (y < 3 || y >= h - 3) return;
        const uint8_t* p = img + (size_t)y * stride;
        for (int x = 3; x < w - 3; ++x) {
            const int v = p[x], lo = v - t, hi = v + t;
            // an arc of 9 covers one of every opposite pair: cheap necessary test first
            const int a0 = p[x + off[0]], a8 = p[x + off[8]];
            bool br = (a0 > hi) || (a8 > hi), dk = (a0 < lo) || (a8 < lo);
            if (!br && !dk) continue;
            const int a4 = p[x + off[4]], a12 = p[x + off[12]];
            br = br && ((a4 > hi) || (a12 > hi));
            dk = dk && ((a4 < lo) || (a12 < lo));
            if (!br && !dk) continue;
            uint32_t mb = 0, md = 0;
            int r[16];
            for (int k = 0; k < 16; ++k) {
                r[k] = p[x + off[k]];
                mb |= (uint32_t)(r[k] > hi) << k;
                md |= (uint32_t)(r[k] < lo) << k;
            }
            int d[16], s = 0;
            if (br && has_arc9(mb)) {
                for (int k = 0; k < 16; ++k) d[k] = r[k] - v;
                s = arc_score(d);
            } else if (dk && has_arc9(md)) {
                for (int k = 0; k < 16; ++k) d[k] = v - r[k];
                s = arc_score(d);
            } else
                continue;
            sc[x] = s - 1;  // s > t >= 0 here
        }
    };
    int* prev = buf.data();
    int* cur = buf.data() + w;
    int* next = buf.data() + 2 * (size_t)w;
    score_row(2, prev);
    score_row(3, cur);
    for (int y = 3; y < h - 3; ++y) {
        score_row(y + 1, next);
        for (int x = 3; x < w - 3; ++x) {
            const int s = cur[x];
            if (s <= 0) continue;  // corners have score >= t >= 1
            if (s > cur[x - 1] && s > cur[x + 1] && s > prev[x - 1] && s > prev[x] && s > prev[x + 1] && s > next[x - 1] && s > next[x] &&
                s > next[x + 1])
                out.push_back({x, y, s});
        }
        int* tmp = prev; prev = cur; cur = next; next = tmp;
    }
}

// ---------------------------------------------------------------------------------------------
// A.5 cv::fastAtan2 (scalar path, float32, no FMA; degrees in [0,360)) — used at ORBextractor.cc:76
// ---------------------------------------------------------------------------------------------
float fast_atan2(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale, p5 = 0.1555786518463281f * scale,
                p7 = -0.04432655554792128f * scale;
    const float eps = (float)2.2204460492503131e-16;  // (float)DBL_EPSILON
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---------------------------------------------------------------------------------------------
// The extractor (ORBextractor.cc:383-443 ctor, :755-870 key points, :529-753 octree, :1060-1168)
// ---------------------------------------------------------------------------------------------
struct QNode {  // ExtractorNode, ORBextractor.h:39-50 (rectangle kept as two corners; the reference's
                // four corners always form an axis-aligned rectangle)
    int x0, y0, x1, y1;
    std::vector<int> keys;  // indices into the candidate vector, order preserved
    bool noMore = false;
    std::list<QNode>::iterator self;
};

struct Extractor {
    int nfeatures, nlevels, iniTh, minTh;
    double scaleFactor;  // ORBextractor.h:105 — double member
    std::vector<float> sf, isf, sig2, isig2;
    std::vector<int> quota, umax;
    int taps[7] = {18, 34, 48, 56, 48, 34, 18};

    // per-call state (stage dumps)
    std::vector<Plane> pyr, blurred;
    std::vector<std::vector<Cand>> cands;
    std::vector<int> fallbackCells;
    std::vector<std::vector<KeyPoint>> selected;  // level coordinates, with angle
    std::vector<KeyPoint> kps;
    std::vector<uint8_t> desc;
    int monoIndex = 0;

    Extractor(int nf, float s, int nl, int ini, int mn) : nfeatures(nf), nlevels(nl), iniTh(ini), minTh(mn), scaleFactor(s) {
        sf.resize(nl); sig2.resize(nl); isf.resize(nl); isig2.resize(nl);
        sf[0] = 1.0f; sig2[0] = 1.0f;
        for (int i = 1; i < nl; ++i) {
            sf[i] = (float)(sf[i - 1] * scaleFactor);  // float * double -> double -> float (ORBextractor.cc:394)
            sig2[i] = sf[i] * sf[i];
        }
        for (int i = 0; i < nl; ++i) { isf[i] = 1.0f / sf[i]; isig2[i] = 1.0f / sig2[i]; }
        quota.resize(nl);
        float factor = (float)(1.0f / scaleFactor);
        float nDesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
        int sum = 0;
        for (int l = 0; l < nl - 1; ++l) {
            quota[l] = cv_round(nDesired);
            sum += quota[l];
            nDesired *= factor;
        }
        quota[nl - 1] = std::max(nfeatures - sum, 0);
        // circular patch row half-widths (ORBextractor.cc:427-442)
        umax.assign(kHalfPatch + 1, 0);
        int vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
        int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
        const double hp2 = kHalfPatch * kHalfPatch;
        for (int v = 0; v <= vmax; ++v) umax[v] = cv_round(std::sqrt(hp2 - v * v));
        for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
    }

    // ORBextractor.cc:1143-1168 (the 19-px border is not materialised: nothing on the path reads it)
    void computePyramid(const uint8_t* img, int w, int h, size_t stride) {
        pyr.assign(nlevels, Plane());
        for (int l = 0; l < nlevels; ++l) {
            float scale = isf[l];
            int lw = cv_round((float)w * scale), lh = cv_round((float)h * scale);
            pyr[l].alloc(lw, lh);
            if (l == 0)
                for (int y = 0; y < h; ++y) std::memcpy(pyr[0].row(y), img + (size_t)y * stride, w);
            else
                resize_linear_8u(pyr[l - 1].d.data(), pyr[l - 1].w, pyr[l - 1].h, pyr[l - 1].w, pyr[l].d.data(), lw, lh, lw);
        }
    }

    // ORBextractor.cc:454-510
    static void divide(const QNode& n, const std::vector<Cand>& c, QNode ch[4]) {
        const int halfX = (int)std::ceil((float)(n.x1 - n.x0) / 2);
        const int halfY = (int)std::ceil((float)(n.y1 - n.y0) / 2);
        const int mx = n.x0 + halfX, my = n.y0 + halfY;
        ch[0].x0 = n.x0; ch[0].y0 = n.y0; ch[0].x1 = mx;   ch[0].y1 = my;
        ch[1].x0 = mx;   ch[1].y0 = n.y0; ch[1].x1 = n.x1; ch[1].y1 = my;
        ch[2].x0 = n.x0; ch[2].y0 = my;   ch[2].x1 = mx;   ch[2].y1 = n.y1;
        ch[3].x0 = mx;   ch[3].y0 = my;   ch[3].x1 = n.x1; ch[3].y1 = n.y1;
        for (int k : n.keys) {
            const Cand& p = c[k];
            int q = (p.x < mx) ? ((p.y < my) ? 0 : 2) : ((p.y < my) ? 1 : 3);
            ch[q].keys.push_back(k);
        }
        for (int q = 0; q < 4; ++q) ch[q].noMore = (ch[q].keys.size() == 1);
    }

    // ORBextractor.cc:529-753.  Returns indices into `c`, in final list order.
    std::vector<int> distributeOctTree(const std::vector<Cand>& c, int minX, int maxX, int minY, int maxY, int N) {
        const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
        const float hX = (float)(maxX - minX) / nIni;
        std::list<QNode> nodes;
        std::vector<QNode*> ini(nIni);
        for (int i = 0; i < nIni; ++i) {
            QNode n;
            n.x0 = (int)(hX * (float)i); n.x1 = (int)(hX * (float)(i + 1));
            n.y0 = 0; n.y1 = maxY - minY;
            nodes.push_back(n);
            ini[i] = &nodes.back();
        }
        for (size_t i = 0; i < c.size(); ++i) ini[(int)(c[i].x / hX)]->keys.push_back((int)i);
        for (auto it = nodes.begin(); it != nodes.end();) {
            if (it->keys.size() == 1) { it->noMore = true; ++it; }
            else if (it->keys.empty()) it = nodes.erase(it);
            else ++it;
        }
        typedef std::pair<int, QNode*> SP;
        auto cmp = [](const SP& a, const SP& b) {  // compareNodes, ORBextractor.cc:512-527
            if (a.first < b.first) return true;
            if (a.first > b.first) return false;
            return a.second->x0 < b.second->x0;
        };
        auto pushChildren = [&](QNode ch[4], std::vector<SP>& rec, int* nToExpand) {
            for (int q = 0; q < 4; ++q) {
                if (ch[q].keys.empty()) continue;
                nodes.push_front(ch[q]);
                if (ch[q].keys.size() > 1) {
                    if (nToExpand) ++*nToExpand;
                    rec.push_back(SP((int)ch[q].keys.size(), &nodes.front()));
                    nodes.front().self = nodes.begin();
                }
            }
        };
        bool finish = false;
        std::vector<SP> rec;
        while (!finish) {
            int prevSize = (int)nodes.size();
            int nToExpand = 0;
            rec.clear();
            for (auto it = nodes.begin(); it != nodes.end();) {
                if (it->noMore) { ++it; continue; }
                QNode ch[4];
                divide(*it, c, ch);
                pushChildren(ch, rec, &nToExpand);
                it = nodes.erase(it);
            }
            if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) finish = true;
            else if ((int)nodes.size() + nToExpand * 3 > N) {
                while (!finish) {
                    prevSize = (int)nodes.size();
                    std::vector<SP> prev = rec;
                    rec.clear();
                    std::sort(prev.begin(), prev.end(), cmp);  // libstdc++ introsort: tie order is part of the result
                    for (int j = (int)prev.size() - 1; j >= 0; --j) {
                        QNode ch[4];
                        divide(*prev[j].second, c, ch);
                        pushChildren(ch, rec, nullptr);
                        nodes.erase(prev[j].second->self);
                        if ((int)nodes.size() >= N) break;
                    }
                    if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) finish = true;
                }
            }
        }
        std::vector<int> out;
        out.reserve(nodes.size());
        for (const QNode& n : nodes) {  // best response per node, first wins ties (ORBextractor.cc:732-750)
            int best = n.keys[0];
            float br = c[best].response;
            for (size_t k = 1; k < n.keys.size(); ++k)
                if (c[n.keys[k]].response > br) { best = n.keys[k]; br = c[best].response; }
            out.push_back(best);
        }
        return out;
    }

    // ORBextractor.cc:50-77
    float icAngle(const Plane& im, int x, int y) const {
        int m01 = 0, m10 = 0;
        const uint8_t* center = im.row(y) + x;
        for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * center[u];
        const int step = im.w;
        for (int v = 1; v <= kHalfPatch; ++v) {
            int vsum = 0, d = umax[v];
            for (int u = -d; u <= d; ++u) {
                int vp = center[u + v * step], vm = center[u - v * step];
                vsum += (vp - vm);
                m10 += u * (vp + vm);
            }
            m01 += v * vsum;
        }
        return fast_atan2((float)m01, (float)m10);
    }

    // ORBextractor.cc:755-870.  Returns false when a level is too small for the cell grid (the reference
    // would divide by zero there).
    bool computeKeyPoints() {
        cands.assign(nlevels, {});
        fallbackCells.assign(nlevels, 0);
        selected.assign(nlevels, {});
        const float W = 35;
        std::vector<FastKp> cell;
        for (int l = 0; l < nlevels; ++l) {
            const Plane& im = pyr[l];
            const int minBX = kEdge - 3, minBY = minBX, maxBX = im.w - kEdge + 3, maxBY = im.h - kEdge + 3;
            const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
            const int nCols = (int)(width / W), nRows = (int)(height / W);
            if (nCols < 1 || nRows < 1) return false;
            const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
            std::vector<Cand>& vc = cands[l];
            for (int i = 0; i < nRows; ++i) {
                const int iniY = minBY + i * hCell;
                int maxY = iniY + hCell + 6;
                if (iniY >= maxBY - 3) continue;
                if (maxY > maxBY) maxY = maxBY;
                for (int j = 0; j < nCols; ++j) {
                    const int iniX = minBX + j * wCell;
                    int maxX = iniX + wCell + 6;
                    if (iniX >= maxBX - 6) continue;
                    if (maxX > maxBX) maxX = maxBX;
                    const uint8_t* crop = im.row(iniY) + iniX;
                    fast9_nms(crop, maxX - iniX, maxY - iniY, im.w, iniTh, cell);
                    if (cell.empty()) {
                        fast9_nms(crop, maxX - iniX, maxY - iniY, im.w, minTh, cell);
                        ++fallbackCells[l];
                    }
                    for (const FastKp& k : cell) vc.push_back({(float)(k.x + j * wCell), (float)(k.y + i * hCell), (float)k.score});
                }
            }
            std::vector<int> keep = distributeOctTree(vc, minBX, maxBX, minBY, maxBY, quota[l]);
            const int scaledPatch = (int)(kPatch * sf[l]);
            for (int idx : keep) {
                KeyPoint kp;
                kp.x = vc[idx].x + minBX; kp.y = vc[idx].y + minBY;
                kp.size = (float)scaledPatch; kp.angle = -1.f; kp.response = vc[idx].response;
                kp.octave = l; kp.class_id = -1;
                selected[l].push_back(kp);
            }
        }
        for (int l = 0; l < nlevels; ++l)
            for (KeyPoint& kp : selected[l]) kp.angle = icAngle(pyr[l], cv_round(kp.x), cv_round(kp.y));
        return true;
    }

    // ORBextractor.cc:81-120
    static void descriptor(const KeyPoint& kp, const Plane& im, uint8_t* out) {
        const float factorPI = (float)(3.14159265358979323846 / 180.f);
        float angle = (float)kp.angle * factorPI;
        float a = cosf(angle), b = sinf(angle);
        const uint8_t* center = im.row(cv_round(kp.y)) + cv_round(kp.x);
        const int step = im.w;
        const signed char* pat = kPattern;
        auto val = [&](int idx) -> int {
            float px = (float)pat[2 * idx], py = (float)pat[2 * idx + 1];
            return center[cv_round(px * b + py * a) * step + cv_round(px * a - py * b)];
        };
        for (int i = 0; i < 32; ++i, pat += 32) {
            int v = 0;
            for (int k = 0; k < 8; ++k) v |= (val(2 * k) < val(2 * k + 1)) << k;
            out[i] = (uint8_t)v;
        }
    }

    // ORBextractor.cc:1060-1141.  Returns monoIndex, -1 for an empty image, -2 when the image is too small.
    int extract(const uint8_t* img, int w, int h, size_t stride, int lap0, int lap1) {
        kps.clear(); desc.clear(); monoIndex = 0;
        if (!img || w <= 0 || h <= 0) return -1;
        computePyramid(img, w, h, stride);
        blurred.assign(nlevels, Plane());
        if (!computeKeyPoints()) return -2;
        int n = 0;
        for (int l = 0; l < nlevels; ++l) n += (int)selected[l].size();
        kps.assign(n, KeyPoint());
        desc.assign((size_t)n * 32, 0);
        int mono = 0, stereo = n - 1;
        uint8_t d[32];
        for (int l = 0; l < nlevels; ++l) {
            if (selected[l].empty()) continue;
            blurred[l].alloc(pyr[l].w, pyr[l].h);
            gaussian_blur7_8u(pyr[l].d.data(), pyr[l].w, pyr[l].h, pyr[l].w, blurred[l].d.data(), pyr[l].w, taps);
            const float scale = sf[l];
            for (const KeyPoint& src : selected[l]) {
                descriptor(src, blurred[l], d);
                KeyPoint kp = src;
                if (l != 0) { kp.x *= scale; kp.y *= scale; }
                int dst = (kp.x >= lap0 && kp.x <= lap1) ? stereo-- : mono++;
                kps[dst] = kp;
                std::memcpy(&desc[(size_t)dst * 32], d, 32);
            }
        }
        monoIndex = mono;
        return mono;
    }
};

// ---------------------------------------------------------------------------------------------
// Frame::ComputeStereoMatches (slam_framework/src/Frame.cc:841-1011) restated: row-banded Hamming search of the right
// keypoints (ORBmatcher::DescriptorDistance, ORBmatcher.cc:2067-2083; TH_HIGH 100, TH_LOW 50, ORBmatcher.cc:44-45), 11x11 SAD
// refinement over 11 shifts on the pyramid level of the left keypoint, parabola fit, median-based outlier rejection.
// uRight / depth = -1 where the left keypoint has no match.
// ---------------------------------------------------------------------------------------------
inline int hamming256(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 32; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

void stereo_match(const Extractor& L, const Extractor& R, float mb, float mbf, std::vector<float>& uRight, std::vector<float>& depth) {
    const int N = (int)L.kps.size(), Nr = (int)R.kps.size();
    uRight.assign(N, -1.0f);
    depth.assign(N, -1.0f);
    if (N == 0 || L.pyr.empty()) return;
    const int thOrbDist = (100 + 50) / 2;
    const int nRows = L.pyr[0].h;
    std::vector<std::vector<int>> rows(nRows);
    for (int iR = 0; iR < Nr; ++iR) {
        const KeyPoint& kp = R.kps[iR];
        const float r = 2.0f * R.sf[kp.octave];
        const int maxr = (int)std::ceil(kp.y + r), minr = (int)std::floor(kp.y - r);
        for (int yi = minr; yi <= maxr; ++yi)
            if (yi >= 0 && yi < nRows) rows[yi].push_back(iR);  // (always inside for valid keypoints)
    }
    const float minZ = mb, minD = 0, maxD = mbf / minZ;
    std::vector<std::pair<int, int>> distIdx;
    for (int iL = 0; iL < N; ++iL) {
        const KeyPoint& kpL = L.kps[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y, uL = kpL.x;
        const std::vector<int>& cands = rows[(size_t)vL];
        if (cands.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = 100;
        int bestIdxR = 0;
        const uint8_t* dL = &L.desc[(size_t)iL * 32];
        for (int iR : cands) {
            const KeyPoint& kpR = R.kps[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = hamming256(dL, &R.desc[(size_t)iR * 32]);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {
            const float uR0 = R.kps[bestIdxR].x;
            const float scaleFactor = L.isf[kpL.octave];
            const float scaleduL = std::round(kpL.x * scaleFactor), scaledvL = std::round(kpL.y * scaleFactor), scaleduR0 = std::round(uR0 * scaleFactor);
            const int w = 5, Lw = 5;
            const Plane& PL = L.pyr[kpL.octave];
            const Plane& PR = R.pyr[kpL.octave];
            int best = 2147483647, bestincR = 0;
            float vDists[11];
            const float iniu = scaleduR0 + Lw - w, endu = scaleduR0 + Lw + w + 1;
            if (iniu < 0 || endu >= PR.w) continue;
            const int y0 = (int)(scaledvL - w), xl0 = (int)(scaleduL - w);
            for (int incR = -Lw; incR <= Lw; ++incR) {
                const int xr0 = (int)(scaleduR0 + incR - w);
                int sad = 0;
                for (int yy = 0; yy < 2 * w + 1; ++yy) {
                    const uint8_t* a = PL.row(y0 + yy) + xl0;
                    const uint8_t* b = PR.row(y0 + yy) + xr0;
                    for (int xx = 0; xx < 2 * w + 1; ++xx) sad += std::abs((int)a[xx] - (int)b[xx]);
                }
                const float dist = (float)sad;  // cv::norm(IL, IR, NORM_L1) as float
                if (dist < best) { best = (int)dist; bestincR = incR; }
                vDists[Lw + incR] = dist;
            }
            if (bestincR == -Lw || bestincR == Lw) continue;
            const float dist1 = vDists[Lw + bestincR - 1], dist2 = vDists[Lw + bestincR], dist3 = vDists[Lw + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = L.sf[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) {
                    disparity = 0.01;
                    bestuR = uL - 0.01;
                }
                depth[iL] = mbf / disparity;
                uRight[iL] = bestuR;
                distIdx.push_back(std::pair<int, int>(best, iL));
            }
        }
    }
    if (distIdx.empty()) return;  // (the reference indexes an empty vector here)
    std::sort(distIdx.begin(), distIdx.end());
    const float median = distIdx[distIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    for (int i = (int)distIdx.size() - 1; i >= 0; --i) {
        if (distIdx[i].first < thDist) break;
        uRight[distIdx[i].second] = -1;
        depth[distIdx[i].second] = -1;
    }
}


// ---------------------------------------------------------------------------------------------
// The caller's grayscale conversion, Tracking::GrabImageStereoLidar (Tracking.cc:1646-1671; the other Grab* entry points
// repeat it): cvtColor(im, im, COLOR_RGB2GRAY | COLOR_BGR2GRAY | COLOR_RGBA2GRAY | COLOR_BGRA2GRAY) on 8-bit images.
// OpenCV 4.x (imgproc/src/color_rgb.simd.hpp, RGB2Gray<uchar>): 15-bit fixed point,
//     gray = (R * 9798 + G * 19235 + B * 3735 + (1 << 14)) >> 15        (RY15, GY15, BY15; alpha ignored).
// Pinned against cv2.cvtColor for all four codes, IPP on and off, over all 2^24 colours (tests/test_oracle_vs_cv2.py).
// ---------------------------------------------------------------------------------------------
void cvt_gray_8u(const uint8_t* src, int w, int h, size_t sstride, int channels, int rgb, uint8_t* dst, size_t dstride) {
    const int RY = 9798, GY = 19235, BY = 3735;
    const int k0 = rgb ? RY : BY, k2 = rgb ? BY : RY;
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = src + (size_t)y * sstride;
        uint8_t* d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x, s += channels) d[x] = (uint8_t)((s[0] * k0 + s[1] * GY + s[2] * k2 + (1 << 14)) >> 15);
    }
}

// ---------------------------------------------------------------------------------------------
// The rectification in front of the Grab* call, System::TrackStereoLidar (System.cc:238-248):
// cv::remap(im, imToFeed, M1, M2, cv::INTER_LINEAR) with the CV_32FC1 maps of cv::initUndistortRectifyMap (Settings.cc:523-526),
// default border (BORDER_CONSTANT, value 0), 8-bit images of 1, 3 or 4 channels.  OpenCV's fixed-point path:
//   sx = cvRound(mapx * 32), sy = cvRound(mapy * 32)  (float product, round half to even); integer parts sx >> 5, sy >> 5
//   (saturated to short), fractions fx = sx & 31, fy = sy & 31; weights 32*(32-fx)*(32-fy), 32*fx*(32-fy), 32*(32-fx)*fy,
//   32*fx*fy (the BilinearTab_i entries: products of multiples of 1/32 scaled by 2^15 are integers, so the table needs no
//   correction step); dst = (sum of weight * tap + 2^14) >> 15, taps outside the source read as 0.
// Pinned against cv2.remap (tests/test_oracle_vs_cv2.py).
// ---------------------------------------------------------------------------------------------
void remap_linear_8u(const uint8_t* src, int sw, int sh, size_t sstride, int channels, const float* mapx, const float* mapy, size_t mstride, uint8_t* dst,
                     int dw, int dh, size_t dstride) {
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x) {
            const int sxq = (int)lrintf(mapx[(size_t)y * mstride + x] * 32.0f), syq = (int)lrintf(mapy[(size_t)y * mstride + x] * 32.0f);
            const int ix = std::min(std::max(sxq >> 5, -32768), 32767), iy = std::min(std::max(syq >> 5, -32768), 32767);
            const int fx = sxq & 31, fy = syq & 31;
            const int w00 = 32 * (32 - fx) * (32 - fy), w01 = 32 * fx * (32 - fy), w10 = 32 * (32 - fx) * fy, w11 = 32 * fx * fy;
            for (int c = 0; c < channels; ++c) {
                auto tap = [&](int yy, int xx) -> int {
                    return (xx >= 0 && xx < sw && yy >= 0 && yy < sh) ? src[(size_t)yy * sstride + (size_t)xx * channels + c] : 0;
                };
                const int v = tap(iy, ix) * w00 + tap(iy, ix + 1) * w01 + tap(iy + 1, ix) * w10 + tap(iy + 1, ix + 1) * w11;
                dst[(size_t)y * dstride + (size_t)x * channels + c] = (uint8_t)((v + (1 << 14)) >> 15);
            }
        }
}

// ---------------------------------------------------------------------------------------------
// Frame::ComputeStereoFishEyeMatches, descriptor stage (Frame.cc:1162-1187; BFmatcher = cv::BFMatcher(cv::NORM_HAMMING),
// Frame.cc:52): knnMatch(k = 2) of the left lapping tail [monoLeft, Nleft) against the right tail [monoRight, Nright)
// and Lowe's ratio gate `m[0].distance < m[1].distance * 0.7` (float * double -> double compare, Frame.cc:1187).
// cv::BFMatcher::knnMatch -> cv::batchDistance(K = 2): the train rows are scanned in index order and a row enters the
// K-best list only if its distance is STRICTLY smaller than the current K-th, sliding in behind equal distances; the
// result is therefore the two smallest (distance, trainIdx) pairs in lexicographic order.  Fewer than two train rows
// give fewer matches (the reference then skips the query: `size() >= 2`); no train rows give none.
// out[q] = {trainIdx0, distance0, trainIdx1, distance1} (-1 where absent), ok[q] = passes the ratio gate.
// Pinned against cv2.BFMatcher(NORM_HAMMING).knnMatch in tests/test_oracle_vs_cv2.py.
// ---------------------------------------------------------------------------------------------
void lapping_knn_match(const uint8_t* dl, int nq, const uint8_t* dr, int nt, int* out, uint8_t* ok) {
    for (int q = 0; q < nq; ++q) {
        int idx[2] = {-1, -1}, dist[2] = {INT_MAX, INT_MAX};
        for (int j = 0; j < nt; ++j) {
            const int d = hamming256(dl + (size_t)q * 32, dr + (size_t)j * 32);
            if (d < dist[1]) {
                int k = 0;
                if (dist[0] > d) { idx[1] = idx[0]; dist[1] = dist[0]; k = -1; }
                idx[k + 1] = j;
                dist[k + 1] = d;
            }
        }
        int* o = out + 4 * (size_t)q;
        o[0] = idx[0]; o[1] = idx[0] >= 0 ? dist[0] : -1;
        o[2] = idx[1]; o[3] = idx[1] >= 0 ? dist[1] : -1;
        ok[q] = (idx[1] >= 0 && (float)dist[0] < (float)dist[1] * 0.7) ? 1 : 0;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C API (ctypes) — used by tests/, smoke() and bench.py's CPU-baseline legs only.
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------------------
// Feature grid of a Frame (widening).  Restates Frame::PosInGrid (Frame.cc:751-762),
// Frame::AssignFeaturesToGrid (Frame.cc:412-443) for the single-grid case (Nleft == -1, mvKeysUn == mvKeys: rectified input,
// zero distortion, Frame::UndistortKeyPoints returns the keys unchanged) and Frame::GetFeaturesInArea (Frame.cc:687-749).
// mGrid[i][j] (FRAME_GRID_COLS = 64 columns, FRAME_GRID_ROWS = 48 rows, Frame.h:59-60) is laid out as CSR: cell c = i * 48 + j
// holds items[cellStart[c] .. cellStart[c + 1]) = the keypoint indices in push_back order (ascending).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int GRID_COLS = 64, GRID_ROWS = 48;

static inline bool pos_in_grid(float x, float y, float minX, float minY, float invW, float invH, int& px, int& py) {
    px = (int)std::round((x - minX) * invW);  // round half away from zero, as ::round / std::round do (Frame.cc:753-754)
    py = (int)std::round((y - minY) * invH);
    return !(px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS);
}

void assign_features_to_grid(const KeyPoint* kps, int n, float minX, float maxX, float minY, float maxY, int* cellStart, int* items) {
    const float invW = (float)GRID_COLS / (maxX - minX), invH = (float)GRID_ROWS / (maxY - minY);  // Frame.cc:179-180
    std::vector<std::vector<int>> grid(GRID_COLS * GRID_ROWS);
    for (int i = 0; i < n; ++i) {
        int px, py;
        if (pos_in_grid(kps[i].x, kps[i].y, minX, minY, invW, invH, px, py)) grid[px * GRID_ROWS + py].push_back(i);
    }
    int pos = 0;
    for (int c = 0; c < GRID_COLS * GRID_ROWS; ++c) {
        cellStart[c] = pos;
        for (int i : grid[c]) items[pos++] = i;
    }
    cellStart[GRID_COLS * GRID_ROWS] = pos;
}

// returns the number of indices the reference would return; the first min(count, cap) are written to out, in the reference's order
int features_in_area(const KeyPoint* kps, const int* cellStart, const int* items, float minX, float maxX, float minY, float maxY, float x, float y, float r,
                     int minLevel, int maxLevel, int* out, int cap) {
    const float invW = (float)GRID_COLS / (maxX - minX), invH = (float)GRID_ROWS / (maxY - minY);
    const int nMinCellX = std::max(0, (int)std::floor((x - minX - r) * invW));
    if (nMinCellX >= GRID_COLS) return 0;
    const int nMaxCellX = std::min(GRID_COLS - 1, (int)std::ceil((x - minX + r) * invW));
    if (nMaxCellX < 0) return 0;
    const int nMinCellY = std::max(0, (int)std::floor((y - minY - r) * invH));
    if (nMinCellY >= GRID_ROWS) return 0;
    const int nMaxCellY = std::min(GRID_ROWS - 1, (int)std::ceil((y - minY + r) * invH));
    if (nMaxCellY < 0) return 0;
    const bool checkLevels = minLevel > 0 || maxLevel >= 0;
    int n = 0;
    for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
        for (int iy = nMinCellY; iy <= nMaxCellY; ++iy) {
            const int c = ix * GRID_ROWS + iy;
            for (int j = cellStart[c]; j < cellStart[c + 1]; ++j) {
                const KeyPoint& kp = kps[items[j]];
                if (checkLevels) {
                    if (kp.octave < minLevel) continue;
                    if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                }
                const float dx = kp.x - x, dy = kp.y - y;
                if (std::fabs(dx) < r && std::fabs(dy) < r) {
                    if (n < cap) out[n] = items[j];
                    ++n;
                }
            }
        }
    return n;
}


// ---------------------------------------------------------------------------------------------------------------------
// The search of ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (ORBmatcher.cc:1685-1895)
// from the point where a map point has been projected (ORBmatcher.cc:1729 on) — single-grid case (CurrentFrame.Nleft == -1).
// What precedes that line is pose algebra on Sophus / Eigen types and stays with the caller; a query carries its outcome:
//   u, v = uv(0), uv(1);  radius = th * mvScaleFactors[nLastOctave];  ur = uv(0) - mbf * invzc;
//   [minLevel, maxLevel] = the GetFeaturesInArea level window of the bForward / bBackward / default branch (:1737-1742);
//   angle = kpLF.angle;  blocks = pMP->Observations() > 0;  desc = pMP->GetDescriptor().
// Restated: the candidate walk (GetFeaturesInArea order), the skip of keypoints already taken by a point with observations
// (:1756-1758; CurrentFrame.mvpMapPoints starts all-NULL as in Tracking::TrackWithMotionModel), the right-coordinate gate
// (:1760-1766), first-minimum Hamming distance (:1768-1776), acceptance at bestDist <= TH_HIGH (:1779-1783), the rotation
// histogram (:1784-1801, bin = round(rot / HISTO_LENGTH)), ComputeThreeMaxima (ORBmatcher.cc:2021-2062) and the removal of the
// matches outside the three dominant bins (:1872-1892).  match[k] = index of the query whose point keypoint k ends up with, -1 = none.
// ---------------------------------------------------------------------------------------------------------------------
struct ProjQuery {
    float u, v, radius, ur;
    int32_t minLevel, maxLevel;
    float angle;
    int32_t blocks;
    uint8_t desc[32];
};
static_assert(sizeof(ProjQuery) == 64, "orbx_proj_query layout");

// The same loop serves ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1898-2019, relocalisation) from
// :1947 on: no right-coordinate gate (uRight = nullptr), every assigned keypoint is skipped (:1962-1963: blocks = 1 in every query and taken0 marks the
// keypoints that hold a point before the search), thHigh = ORBdist, angle = pKF->mvKeysUn[i].angle.
int search_by_projection_tail(const KeyPoint* kps, int nKps, const uint8_t* desc, const float* uRight, const uint8_t* taken0, const int* cellStart, const int* items,
                              float minX, float maxX, float minY, float maxY, const ProjQuery* q, int nq, int checkOrientation, int thHigh, int* match) {
    constexpr int HISTO_LENGTH = 30;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<int> cand((size_t)std::max(nKps, 1));
    std::vector<char> taken((size_t)std::max(nKps, 1), 0);  // mvpMapPoints[i2] && mvpMapPoints[i2]->Observations() > 0
    for (int k = 0; k < nKps; ++k) { match[k] = -1; taken[k] = taken0 ? (taken0[k] != 0) : 0; }
    int nmatches = 0;
    for (int i = 0; i < nq; ++i) {
        const ProjQuery& Q = q[i];
        const int nc = features_in_area(kps, cellStart, items, minX, maxX, minY, maxY, Q.u, Q.v, Q.radius, Q.minLevel, Q.maxLevel, cand.data(), nKps);
        if (nc == 0) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int c = 0; c < nc; ++c) {
            const int i2 = cand[c];
            if (taken[i2]) continue;
            if (uRight && uRight[i2] > 0) {
                const float er = std::fabs(Q.ur - uRight[i2]);
                if (er > Q.radius) continue;
            }
            const int dist = hamming256(Q.desc, desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= thHigh) {
            match[bestIdx2] = i;
            taken[bestIdx2] = Q.blocks != 0;
            ++nmatches;
            if (checkOrientation) {
                float rot = Q.angle - kps[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (checkOrientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
        for (int i = 0; i < HISTO_LENGTH; ++i) {  // ORBmatcher::ComputeThreeMaxima
            const int sz = (int)rotHist[i].size();
            if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
            else if (sz > max3) { max3 = sz; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int k : rotHist[i]) { match[k] = -1; --nmatches; }
    }
    return nmatches;
}


// ---------------------------------------------------------------------------------------------------------------------
// ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, th, bFarPoints, thFarPoints) (ORBmatcher.cc:52-222),
// the local-map search of Tracking::SearchLocalPoints (Tracking.cc:3282), single-grid case (F.Nleft == -1).  Its body has no pose
// algebra: the projections were stored in the map points by Frame::isInFrustum.  A query = one map point that passes the early
// tests (:61-68, mbTrackInView): u, v = mTrackProjX / mTrackProjY; radius = r * F.mvScaleFactors[nPredictedLevel] with
// r = RadiusByViewingCos(mTrackViewCos) (* th if th != 1) (:75-81); ur = mTrackProjXR; [minLevel, maxLevel] = [nPredictedLevel - 1,
// nPredictedLevel]; blocks = Observations() > 0; desc = GetDescriptor().  taken0[k] != 0 where F.mvpMapPoints[k] already holds a point
// with observations when the search starts.  Restated: candidate walk, taken skip (:95-97), right-coordinate gate (:99-104), best /
// second-best distance with their levels in the reference's update order (:110-127), acceptance at bestDist <= TH_HIGH unless both
// are on one level and bestDist > mfNNratio * bestDist2 (:131-137).  match[k] = query index written to F.mvpMapPoints[k] by THIS
// search, -1 = left as it was.
// ---------------------------------------------------------------------------------------------------------------------
int search_local_points(const KeyPoint* kps, int nKps, const uint8_t* desc, const float* uRight, const uint8_t* taken0, const int* cellStart, const int* items,
                        float minX, float maxX, float minY, float maxY, const ProjQuery* q, int nq, float nnRatio, int thHigh, int* match) {
    std::vector<int> cand((size_t)std::max(nKps, 1));
    std::vector<char> taken((size_t)std::max(nKps, 1), 0);
    for (int k = 0; k < nKps; ++k) { match[k] = -1; taken[k] = taken0 ? (taken0[k] != 0) : 0; }
    int nmatches = 0;
    for (int i = 0; i < nq; ++i) {
        const ProjQuery& Q = q[i];
        const int nc = features_in_area(kps, cellStart, items, minX, maxX, minY, maxY, Q.u, Q.v, Q.radius, Q.minLevel, Q.maxLevel, cand.data(), nKps);
        if (nc == 0) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int c = 0; c < nc; ++c) {
            const int idx = cand[c];
            if (taken[idx]) continue;
            if (uRight && uRight[idx] > 0) {
                const float er = std::fabs(Q.ur - uRight[idx]);
                if (er > Q.radius) continue;
            }
            const int dist = hamming256(Q.desc, desc + (size_t)idx * 32);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = kps[idx].octave; bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = kps[idx].octave; bestDist2 = dist;
            }
        }
        if (bestDist <= thHigh) {
            if (bestLevel == bestLevel2 && (float)bestDist > nnRatio * (float)bestDist2) continue;
            if (bestLevel != bestLevel2 || (float)bestDist <= nnRatio * (float)bestDist2) {
                match[bestIdx] = i;
                taken[bestIdx] = Q.blocks != 0;
                ++nmatches;
            }
        }
    }
    return nmatches;
}



// ---------------------------------------------------------------------------------------------------------------------
// The two-grid (fisheye stereo, Frame::Nleft != -1) branches of both searches.  The frame is a left and a right image with one grid
// each (Frame::mGrid / mGridRight, Frame.cc:428-441), mvpMapPoints holds N = Nleft + Nright entries (left keypoints first), the
// right-coordinate gate does not apply (Nleft == -1 is false, ORBmatcher.cc:99, :1760).  A map point becomes up to two sub-queries in
// a row, told apart by side[i]: 0 = its search in the left image, 1 = its search in the right image after a left one ("paired"),
// 2 = a right-image search of a point that has no left part (local map: mbTrackInViewR without mbTrackInView).
// Frame to frame (ORBmatcher.cc:1729-1868): the left part as in the single-grid case; `if(vIndices2.empty()) continue;` (:1744-1745)
// also skips the right part of that point; the right part (:1803-1866) walks mGridRight with (u, v) of the point re-projected into the
// right camera (the caller's pose algebra), writes mvpMapPoints[bestIdx2 + Nleft] and votes in the same rotation histogram with
// mvKeysRight[bestIdx2].angle.
// Local map (ORBmatcher.cc:72-219): the ratio-test `continue` of the left part (:133-134) skips the right part too; an accepted left
// match also takes the right keypoint mvLeftToRightMatch[bestIdx] (+1 match, :139-143), an accepted right match the left keypoint
// mvRightToLeftMatch[bestIdx] (:205-209).
// matchL / matchR[k] = index of the SUB-QUERY whose point the keypoint ends up with (-1 = none / untouched).
// ---------------------------------------------------------------------------------------------------------------------
struct StereoSide {
    const KeyPoint* kps; int n; const uint8_t* desc; const int* cellStart; const int* items; int* match; std::vector<char> taken;
};

int search_by_projection_tail_stereo(StereoSide& L, StereoSide& R, float minX, float maxX, float minY, float maxY, const ProjQuery* q, const uint8_t* side, int nq,
                                     int checkOrientation, int thHigh) {
    constexpr int HISTO_LENGTH = 30;
    std::vector<std::pair<int, int>> rotHist[HISTO_LENGTH];   // (side, keypoint)
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<int> cand((size_t)std::max(std::max(L.n, R.n), 1));
    for (StereoSide* S : {&L, &R}) { S->taken.assign((size_t)std::max(S->n, 1), 0); for (int k = 0; k < S->n; ++k) S->match[k] = -1; }
    int nmatches = 0;
    bool skipRight = false;
    for (int i = 0; i < nq; ++i) {
        const ProjQuery& Q = q[i];
        const int sd = side[i];
        if (sd == 0) skipRight = false;
        else if (sd == 1 && skipRight) continue;
        StereoSide& S = sd == 0 ? L : R;
        const int nc = features_in_area(S.kps, S.cellStart, S.items, minX, maxX, minY, maxY, Q.u, Q.v, Q.radius, Q.minLevel, Q.maxLevel, cand.data(), S.n);
        if (nc == 0) { if (sd == 0) skipRight = true; continue; }
        int bestDist = 256, bestIdx2 = -1;
        for (int c = 0; c < nc; ++c) {
            const int i2 = cand[c];
            if (S.taken[i2]) continue;
            const int dist = hamming256(Q.desc, S.desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= thHigh) {
            S.match[bestIdx2] = i;
            S.taken[bestIdx2] = Q.blocks != 0;
            ++nmatches;
            if (checkOrientation) {
                float rot = Q.angle - S.kps[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(std::make_pair(sd == 0 ? 0 : 1, bestIdx2));
            }
        }
    }
    if (checkOrientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
        for (int i = 0; i < HISTO_LENGTH; ++i) {  // ORBmatcher::ComputeThreeMaxima
            const int sz = (int)rotHist[i].size();
            if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
            else if (sz > max3) { max3 = sz; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i)
            if (i != ind1 && i != ind2 && i != ind3)
                for (auto& e : rotHist[i]) { (e.first ? R : L).match[e.second] = -1; --nmatches; }
    }
    return nmatches;
}

int search_local_points_stereo(StereoSide& L, StereoSide& R, const uint8_t* takenL0, const uint8_t* takenR0, const int* leftToRight, const int* rightToLeft, float minX,
                               float maxX, float minY, float maxY, const ProjQuery* q, const uint8_t* side, int nq, float nnRatio, int thHigh) {
    std::vector<int> cand((size_t)std::max(std::max(L.n, R.n), 1));
    L.taken.assign((size_t)std::max(L.n, 1), 0); R.taken.assign((size_t)std::max(R.n, 1), 0);
    for (int k = 0; k < L.n; ++k) { L.match[k] = -1; L.taken[k] = takenL0 ? (takenL0[k] != 0) : 0; }
    for (int k = 0; k < R.n; ++k) { R.match[k] = -1; R.taken[k] = takenR0 ? (takenR0[k] != 0) : 0; }
    int nmatches = 0;
    bool skipRight = false;
    for (int i = 0; i < nq; ++i) {
        const ProjQuery& Q = q[i];
        const int sd = side[i];
        if (sd == 0) skipRight = false;
        else if (sd == 1 && skipRight) continue;
        StereoSide& S = sd == 0 ? L : R;
        StereoSide& O = sd == 0 ? R : L;
        const int nc = features_in_area(S.kps, S.cellStart, S.items, minX, maxX, minY, maxY, Q.u, Q.v, Q.radius, Q.minLevel, Q.maxLevel, cand.data(), S.n);
        if (nc == 0) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int c = 0; c < nc; ++c) {
            const int idx = cand[c];
            if (S.taken[idx]) continue;
            const int dist = hamming256(Q.desc, S.desc + (size_t)idx * 32);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = S.kps[idx].octave; bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = S.kps[idx].octave; bestDist2 = dist;
            }
        }
        if (bestDist <= thHigh) {
            if (bestLevel == bestLevel2 && (float)bestDist > nnRatio * (float)bestDist2) { if (sd == 0) skipRight = true; continue; }
            const int partner = (sd == 0 ? leftToRight : rightToLeft) ? (sd == 0 ? leftToRight : rightToLeft)[bestIdx] : -1;
            if (partner != -1) { O.match[partner] = i; O.taken[partner] = Q.blocks != 0; ++nmatches; }
            S.match[bestIdx] = i;
            S.taken[bestIdx] = Q.blocks != 0;
            ++nmatches;
        }
    }
    return nmatches;
}


// ---------------------------------------------------------------------------------------------------------------------
// Frame::UndistortKeyPoints (Frame.cc:777-810): cv::undistortPoints(mat, mat, toK(), mDistCoef, cv::Mat(), mK) on the keypoint coordinates, i.e.
// OpenCV's cvUndistortPointsInternal with its default criteria (5 fixed-point iterations, no epsilon test), no tilt, R = identity, P = mK: all in
// double, left to right, no FMA contraction; the result is rounded to float once.  cam / newCam = {fx, fy, cx, cy} (float, as the cv::Mat_<float>
// matrices hold them); dist = k1, k2, p1, p2[, k3[, k4, k5, k6[, s1, s2, s3, s4]]] (nDist = 4, 5, 8 or 12).  Pinned against cv2.undistortPoints.
// ---------------------------------------------------------------------------------------------------------------------
void undistort_points(const float* xyIn, int n, const float cam[4], const float* dist, int nDist, const float newCam[4], float* xyOut) {
    const double fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3];
    const double ifx = 1. / fx, ify = 1. / fy;
    double k[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < nDist && i < 12; ++i) k[i] = dist[i];
    const double nfx = newCam[0], nfy = newCam[1], ncx = newCam[2], ncy = newCam[3];
    for (int i = 0; i < n; ++i) {
        double x = xyIn[2 * i], y = xyIn[2 * i + 1];
        const double u = x, v = y;
        x = (x - cx) * ifx;
        y = (y - cy) * ify;
        if (nDist > 0) {
            const double x0 = x, y0 = y;
            for (int j = 0; j < 5; ++j) {
                const double r2 = x * x + y * y;
                const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
                if (icdist < 0) {   // (OpenCV: "test: undistortPoints.regression_14583")
                    x = (u - cx) * ifx;
                    y = (v - cy) * ify;
                    break;
                }
                const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
                const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
                x = (x0 - deltaX) * icdist;
                y = (y0 - deltaY) * icdist;
            }
        }
        const double xx = nfx * x + 0.0 * y + ncx, yy = 0.0 * x + nfy * y + ncy, ww = 1. / (0.0 * x + 0.0 * y + 1.0);
        xyOut[2 * i] = (float)(xx * ww);
        xyOut[2 * i + 1] = (float)(yy * ww);
    }
}


// ------------------------------------------------------------------------------------------------------------------
// DBoW2 vocabulary: what Frame::ComputeBoW (Frame.cc:768-775) runs on the descriptors the extractor produced.  Restated from the
// vendored library (slam_framework/Thirdparty/DBoW2/DBoW2): the text loader the reference uses for ORBvoc.txt
// (TemplatedVocabulary.h:1350-1433), the tree descent (:1230-1276) with FORB::distance (FORB.cpp:81-100), the vector assembly
// (:1139-1214), BowVector::addWeight / addIfNotExist / normalize (BowVector.cpp:34-84), mustNormalize (ScoringObject.h:74-89).
// Pinned against that code compiled here (oracle/_ref/libref_bow.so, tests/test_oracle_vs_ref.py).
// One deliberate difference: the reference's `while(!f.eof())` turns a trailing empty line into one more child of the root with an
// UNINITIALISED descriptor (FORB::fromString sets no byte when the string is empty); here empty lines are skipped.  The results
// are the same for files that do not end in a blank line (the tests write such files).
// ------------------------------------------------------------------------------------------------------------------
struct Vocabulary {
    int k = 0, L = 0, scoring = 0, weighting = 0;
    std::vector<int> parent;                 // node 0 is the root
    std::vector<std::vector<int>> children;  // in file order (= ascending node id)
    std::vector<uint8_t> desc;               // [node][32]
    std::vector<double> weight;
    std::vector<int> wordId;                 // -1 for inner nodes
    int nWords = 0;
};

static bool vocabulary_load_text(const char* path, Vocabulary& v) {
    FILE* f = std::fopen(path, "r");
    if (!f) return false;
    std::vector<char> line(1 << 12);
    if (!std::fgets(line.data(), (int)line.size(), f)) { std::fclose(f); return false; }
    if (std::sscanf(line.data(), "%d %d %d %d", &v.k, &v.L, &v.scoring, &v.weighting) != 4 || v.k < 0 || v.k > 20 || v.L < 1 || v.L > 10 || v.scoring < 0 ||
        v.scoring > 5 || v.weighting < 0 || v.weighting > 3) { std::fclose(f); return false; }
    v.parent.assign(1, 0); v.children.assign(1, {}); v.desc.assign(32, 0); v.weight.assign(1, 0.0); v.wordId.assign(1, -1);
    while (std::fgets(line.data(), (int)line.size(), f)) {
        char* p = line.data();
        char* e = nullptr;
        const long pid = std::strtol(p, &e, 10);
        if (e == p) continue;  // blank line
        p = e;
        const long leaf = std::strtol(p, &e, 10); p = e;
        const int nid = (int)v.parent.size();
        v.parent.push_back((int)pid);
        v.children.emplace_back();
        v.children[(size_t)pid].push_back(nid);
        for (int i = 0; i < 32; ++i) { const long b = std::strtol(p, &e, 10); p = e; v.desc.push_back((uint8_t)b); }
        v.weight.push_back(std::strtod(p, &e));
        v.wordId.push_back(leaf > 0 ? v.nWords++ : -1);
    }
    std::fclose(f);
    return true;
}


// one descriptor down the tree: the child with the smallest distance, the FIRST of equals (strict <); nid = the node at level L - levelsup
static void vocabulary_descend(const Vocabulary& v, const uint8_t* d, int levelsup, int* word, double* w, int* nid) {
    const int nidLevel = v.L - levelsup;
    *nid = 0;  // (the reference sets it only when nidLevel <= 0 or the descent passes that level: complete trees always do)
    int node = 0, level = 0;
    do {
        ++level;
        const std::vector<int>& ch = v.children[(size_t)node];
        node = ch[0];
        int best = hamming256(d, &v.desc[(size_t)node * 32]);
        for (size_t c = 1; c < ch.size(); ++c) {
            const int dist = hamming256(d, &v.desc[(size_t)ch[c] * 32]);
            if (dist < best) { best = dist; node = ch[c]; }
        }
        if (level == nidLevel) *nid = node;
    } while (!v.children[(size_t)node].empty());
    *word = v.wordId[(size_t)node] < 0 ? 0 : v.wordId[(size_t)node];   // (Node::word_id defaults to 0)
    *w = v.weight[(size_t)node];
}

static void vocabulary_transform(const Vocabulary& v, const uint8_t* desc, int n, int levelsup, std::map<unsigned, double>& bow,
                                 std::map<unsigned, std::vector<unsigned>>& fv) {
    bow.clear(); fv.clear();
    if (v.parent.size() <= 1) return;
    const bool must = v.scoring != 5;             // DotProductScoring is the one that does not normalise
    const bool l2 = v.scoring == 1;
    const bool tf = v.weighting == 0 || v.weighting == 1;   // TF_IDF, TF: weights add up; IDF, BINARY: a word counts once
    for (int i = 0; i < n; ++i) {
        int word, nid; double w;
        vocabulary_descend(v, desc + (size_t)i * 32, levelsup, &word, &w, &nid);
        if (w > 0) {
            auto it = bow.lower_bound((unsigned)word);
            if (it != bow.end() && it->first == (unsigned)word) { if (tf) it->second += w; }
            else bow.insert(it, std::make_pair((unsigned)word, w));
            fv[(unsigned)nid].push_back((unsigned)i);
        }
    }
    if (tf && !bow.empty() && !must) {
        const double nd = (double)bow.size();
        for (auto& e : bow) e.second /= nd;
    }
    if (must) {
        double norm = 0.0;
        if (l2) { for (auto& e : bow) norm += e.second * e.second; norm = std::sqrt(norm); }
        else for (auto& e : bow) norm += std::fabs(e.second);
        if (norm > 0.0) for (auto& e : bow) e.second /= norm;
    }
}

extern "C" {
void* orbo_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    return new Extractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void orbo_destroy(void* h) { delete (Extractor*)h; }
void orbo_set_blur_taps(void* h, const int* taps) { std::memcpy(((Extractor*)h)->taps, taps, 7 * sizeof(int)); }
int orbo_extract(void* h, const uint8_t* img, int w, int hh, int stride, int lap0, int lap1) {
    return ((Extractor*)h)->extract(img, w, hh, (size_t)stride, lap0, lap1);
}
int orbo_num_keypoints(void* h) { return (int)((Extractor*)h)->kps.size(); }
void orbo_get_keypoints(void* h, void* out) {
    Extractor* e = (Extractor*)h;
    if (!e->kps.empty()) std::memcpy(out, e->kps.data(), e->kps.size() * sizeof(KeyPoint));
}
void orbo_get_descriptors(void* h, uint8_t* out) {
    Extractor* e = (Extractor*)h;
    if (!e->desc.empty()) std::memcpy(out, e->desc.data(), e->desc.size());
}
void orbo_get_tables(void* h, float* sf, float* isf, float* sig2, float* isig2, int* quota, int* umax) {
    Extractor* e = (Extractor*)h;
    for (int i = 0; i < e->nlevels; ++i) { sf[i] = e->sf[i]; isf[i] = e->isf[i]; sig2[i] = e->sig2[i]; isig2[i] = e->isig2[i]; quota[i] = e->quota[i]; }
    for (int i = 0; i < 16; ++i) umax[i] = e->umax[i];
}
void orbo_level_size(void* h, int l, int* w, int* hh) { Extractor* e = (Extractor*)h; *w = e->pyr[l].w; *hh = e->pyr[l].h; }
void orbo_get_level(void* h, int l, uint8_t* out) { Extractor* e = (Extractor*)h; std::memcpy(out, e->pyr[l].d.data(), e->pyr[l].d.size()); }
int orbo_get_blurred(void* h, int l, uint8_t* out) {
    Extractor* e = (Extractor*)h;
    if (e->blurred[l].d.empty()) return 0;
    std::memcpy(out, e->blurred[l].d.data(), e->blurred[l].d.size());
    return 1;
}
int orbo_num_candidates(void* h, int l) { return (int)((Extractor*)h)->cands[l].size(); }
void orbo_get_candidates(void* h, int l, float* xyr) {
    Extractor* e = (Extractor*)h;
    if (!e->cands[l].empty()) std::memcpy(xyr, e->cands[l].data(), e->cands[l].size() * sizeof(Cand));
}
int orbo_num_fallback_cells(void* h, int l) { return ((Extractor*)h)->fallbackCells[l]; }
int orbo_num_selected(void* h, int l) { return (int)((Extractor*)h)->selected[l].size(); }
void orbo_get_selected(void* h, int l, void* out) {
    Extractor* e = (Extractor*)h;
    if (!e->selected[l].empty()) std::memcpy(out, e->selected[l].data(), e->selected[l].size() * sizeof(KeyPoint));
}
int orbo_stereo_match(void* hl, void* hr, float mb, float mbf, float* uright, float* depth) {
    std::vector<float> u, d;
    stereo_match(*(Extractor*)hl, *(Extractor*)hr, mb, mbf, u, d);
    if (!u.empty()) { std::memcpy(uright, u.data(), u.size() * 4); std::memcpy(depth, d.data(), d.size() * 4); }
    return (int)u.size();
}
void orbo_assign_grid(const void* kps, int n, float minX, float maxX, float minY, float maxY, int* cellStart, int* items) {
    assign_features_to_grid((const KeyPoint*)kps, n, minX, maxX, minY, maxY, cellStart, items);
}
int orbo_features_in_area(const void* kps, const int* cellStart, const int* items, float minX, float maxX, float minY, float maxY, float x, float y, float r,
                          int minLevel, int maxLevel, int* out, int cap) {
    return features_in_area((const KeyPoint*)kps, cellStart, items, minX, maxX, minY, maxY, x, y, r, minLevel, maxLevel, out, cap);
}
int orbo_search_by_projection(const void* kps, int nKps, const uint8_t* desc, const float* uRight, const uint8_t* taken0, const int* cellStart, const int* items,
                              float minX, float maxX, float minY, float maxY, const void* q, int nq, int checkOrientation, int thHigh, int* match) {
    return search_by_projection_tail((const KeyPoint*)kps, nKps, desc, uRight, taken0, cellStart, items, minX, maxX, minY, maxY, (const ProjQuery*)q, nq,
                                     checkOrientation, thHigh, match);
}
int orbo_search_local_points(const void* kps, int nKps, const uint8_t* desc, const float* uRight, const uint8_t* taken0, const int* cellStart, const int* items,
                             float minX, float maxX, float minY, float maxY, const void* q, int nq, float nnRatio, int thHigh, int* match) {
    return search_local_points((const KeyPoint*)kps, nKps, desc, uRight, taken0, cellStart, items, minX, maxX, minY, maxY, (const ProjQuery*)q, nq, nnRatio, thHigh, match);
}
void orbo_lapping_match(const uint8_t* dl, int nq, const uint8_t* dr, int nt, int* out, uint8_t* ok) { lapping_knn_match(dl, nq, dr, nt, out, ok); }
void orbo_cvt_gray(const uint8_t* src, int w, int h, int sstride, int channels, int rgb, uint8_t* dst, int dstride) {
    cvt_gray_8u(src, w, h, (size_t)sstride, channels, rgb, dst, (size_t)dstride);
}
void orbo_remap(const uint8_t* src, int sw, int sh, int sstride, int channels, const float* mapx, const float* mapy, int mstride, uint8_t* dst, int dw, int dh,
                int dstride) {
    remap_linear_8u(src, sw, sh, (size_t)sstride, channels, mapx, mapy, (size_t)mstride, dst, dw, dh, (size_t)dstride);
}
// primitives, for direct checks against cv2
int orbo_search_by_projection_stereo(const void* kpsL, int nL, const uint8_t* descL, const int* startL, const int* itemsL, const void* kpsR, int nR, const uint8_t* descR,
                                     const int* startR, const int* itemsR, float minX, float maxX, float minY, float maxY, const void* q, const uint8_t* side, int nq,
                                     int checkOrientation, int thHigh, int* matchL, int* matchR) {
    StereoSide L{(const KeyPoint*)kpsL, nL, descL, startL, itemsL, matchL, {}}, R{(const KeyPoint*)kpsR, nR, descR, startR, itemsR, matchR, {}};
    return search_by_projection_tail_stereo(L, R, minX, maxX, minY, maxY, (const ProjQuery*)q, side, nq, checkOrientation, thHigh);
}
int orbo_search_local_points_stereo(const void* kpsL, int nL, const uint8_t* descL, const uint8_t* takenL0, const int* startL, const int* itemsL, const int* leftToRight,
                                    const void* kpsR, int nR, const uint8_t* descR, const uint8_t* takenR0, const int* startR, const int* itemsR, const int* rightToLeft,
                                    float minX, float maxX, float minY, float maxY, const void* q, const uint8_t* side, int nq, float nnRatio, int thHigh, int* matchL,
                                    int* matchR) {
    StereoSide L{(const KeyPoint*)kpsL, nL, descL, startL, itemsL, matchL, {}}, R{(const KeyPoint*)kpsR, nR, descR, startR, itemsR, matchR, {}};
    return search_local_points_stereo(L, R, takenL0, takenR0, leftToRight, rightToLeft, minX, maxX, minY, maxY, (const ProjQuery*)q, side, nq, nnRatio, thHigh);
}
void orbo_undistort_points(const float* xy, int n, const float* cam, const float* dist, int nDist, const float* newCam, float* out) {
    undistort_points(xy, n, cam, dist, nDist, newCam, out);
}
// vocabulary (DBoW2) -------------------------------------------------------------------------------------------------
void* orbo_voc_load_text(const char* path) {
    Vocabulary* v = new Vocabulary();
    if (!vocabulary_load_text(path, *v)) { delete v; return nullptr; }
    return v;
}
void orbo_voc_destroy(void* h) { delete (Vocabulary*)h; }
void orbo_voc_info(void* h, int out[6]) {
    Vocabulary* v = (Vocabulary*)h;
    out[0] = v->k; out[1] = v->L; out[2] = v->scoring; out[3] = v->weighting; out[4] = (int)v->parent.size(); out[5] = v->nWords;
}
// the tree as flat arrays (node 0 = root): parent[n], is_word[n], descriptors[n][32], weights[n]
void orbo_voc_export(void* h, int* parent, uint8_t* isWord, uint8_t* desc, double* weight) {
    Vocabulary* v = (Vocabulary*)h;
    const size_t n = v->parent.size();
    for (size_t i = 0; i < n; ++i) { parent[i] = v->parent[i]; isWord[i] = v->wordId[i] >= 0; weight[i] = v->weight[i]; }
    std::memcpy(desc, v->desc.data(), n * 32);
}
int orbo_bow_transform(void* h, const uint8_t* desc, int n, int levelsup, uint32_t* words, double* values, int* nWords, uint32_t* nodes, int* nodeStart,
                       uint32_t* features, int* nNodes) {
    std::map<unsigned, double> bow;
    std::map<unsigned, std::vector<unsigned>> fv;
    vocabulary_transform(*(Vocabulary*)h, desc, n, levelsup, bow, fv);
    int k = 0;
    for (auto& e : bow) { words[k] = e.first; values[k] = e.second; ++k; }
    *nWords = k;
    int m = 0, pos = 0;
    for (auto& e : fv) {
        nodes[m] = e.first; nodeStart[m] = pos;
        for (unsigned f : e.second) features[pos++] = f;
        ++m;
    }
    nodeStart[m] = pos;
    *nNodes = m;
    return 0;
}
void orbo_synth_v1(uint8_t* out, int W, int H, uint32_t seed, int xoff, uint32_t noise_xor) { synth_v1(out, W, H, seed, xoff, noise_xor); }
void orbo_resize(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride) {
    resize_linear_8u(src, sw, sh, (size_t)sstride, dst, dw, dh, (size_t)dstride);
}
void orbo_blur(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, const int* taps) {
    gaussian_blur7_8u(src, w, h, (size_t)sstride, dst, (size_t)dstride, taps);
}
int orbo_fast(const uint8_t* img, int w, int h, int stride, int t, int* xys, int cap) {
    std::vector<FastKp> v;
    fast9_nms(img, w, h, (size_t)stride, t, v);
    int n = (int)std::min<size_t>(v.size(), (size_t)cap);
    for (int i = 0; i < n; ++i) { xys[3 * i] = v[i].x; xys[3 * i + 1] = v[i].y; xys[3 * i + 2] = v[i].score; }
    return (int)v.size();
}
float orbo_fast_atan2(float y, float x) { return fast_atan2(y, x); }
void orbo_sincos(float angle, float* s, float* c) { *s = sinf(angle); *c = cosf(angle); }
// std::sort on (key, id) pairs with the reference comparator semantics (key = (nKeys, UL.x) packed by the caller):
// used to check the product's libstdc++-exact sort port.
void orbo_std_sort_pairs(int* firsts, int* seconds, int n) {
    std::vector<std::pair<std::pair<int, int>, int>> v(n);
    for (int i = 0; i < n; ++i) v[i] = {{firsts[i], seconds[i]}, i};
    std::sort(v.begin(), v.end(), [](const std::pair<std::pair<int, int>, int>& a, const std::pair<std::pair<int, int>, int>& b) {
        if (a.first.first < b.first.first) return true;
        if (a.first.first > b.first.first) return false;
        return a.first.second < b.first.second;
    });
    for (int i = 0; i < n; ++i) { firsts[i] = v[i].second; }  // permutation: output slot i holds input index
    (void)seconds;
}
}
