// cvstub.hpp — the handful of OpenCV types and functions that TC2LI-SLAM's ORBextractor.cc touches, written from the
// documented OpenCV semantics so that the reference's OWN translation unit (slam_framework/src/ORBextractor.cc, compiled
// where it lies, never copied) can be built in an image without the OpenCV SDK.  TEST INFRASTRUCTURE ONLY (oracle/_ref).
//
// Containers (Mat with shared buffers and ROI views, KeyPoint, Point_, Size, Rect, Input/OutputArray) are implemented
// here; the pixel primitives (resize INTER_LINEAR, copyMakeBorder REFLECT_101, FAST 9/16 + NMS, GaussianBlur 7x7, fastAtan2)
// forward to the restated primitives of oracle/orb_oracle.cpp, each of which is checked bit for bit against the real
// OpenCV (cv2 4.13) in tests/test_oracle_vs_cv2.py.  So: control flow = the reference's, primitives = cv2-verified.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

typedef unsigned char uchar;
#define CV_8U 0
#define CV_8UC1 0
#define CV_PI 3.1415926535897932384626433832795

extern "C" {
void orbo_resize(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
void orbo_blur(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, const int* taps);
int orbo_fast(const uint8_t* img, int w, int h, int stride, int t, int* xys, int cap);
float orbo_fast_atan2(float y, float x);
}

inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16, INTER_LINEAR = 1, NORM_L1 = 2 };

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
    int x, y, width, height;
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

struct KeyPoint {  // same field order as cv::KeyPoint (28 bytes)
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

struct MatStep {
    size_t v;
    operator size_t() const { return v; }
};

class Mat {  // 8-bit single-channel only; shared buffer + region of interest
public:
    int rows, cols;
    uchar* data;
    MatStep step;
    Mat() : rows(0), cols(0), data(nullptr), step{0} {}
    Mat(int r, int c, int) : rows(0), cols(0), data(nullptr), step{0} { create(r, c, 0); }
    Mat(Size s, int) : rows(0), cols(0), data(nullptr), step{0} { create(s.height, s.width, 0); }
    Mat(int r, int c, int, void* ext, size_t st) : rows(r), cols(c), data((uchar*)ext), step{st} {}  // non-owning
    void create(int r, int c, int) {
        if (data && r == rows && c == cols) return;
        buf = std::shared_ptr<uchar>(new uchar[(size_t)std::max(r * c, 1)], std::default_delete<uchar[]>());
        rows = r; cols = c; data = buf.get(); step.v = (size_t)c;
    }
    static Mat zeros(int r, int c, int t) { Mat m(r, c, t); std::memset(m.data, 0, (size_t)r * c); return m; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    size_t step1() const { return step.v; }
    void release() { buf.reset(); data = nullptr; rows = cols = 0; step.v = 0; }
    Mat operator()(const Rect& r) const { Mat m = *this; m.data = data + (size_t)r.y * step.v + r.x; m.rows = r.height; m.cols = r.width; return m; }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
    Mat row(int y) const { return rowRange(y, y + 1); }
    Mat clone() const {
        Mat m(rows, cols, 0);
        for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step.v, data + (size_t)y * step.v, (size_t)cols);
        return m;
    }
    void copyTo(Mat dst) const {  // the reference only copies equally sized rows into existing matrices
        assert(dst.rows == rows && dst.cols == cols);
        for (int y = 0; y < rows; ++y) std::memcpy(dst.data + (size_t)y * dst.step.v, data + (size_t)y * step.v, (size_t)cols);
    }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step.v + x); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step.v + x); }
    uchar* ptr(int y = 0) { return data + (size_t)y * step.v; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step.v; }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step.v); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step.v); }
    bool sameBuffer(const Mat& o) const { return buf && buf == o.buf; }

private:
    std::shared_ptr<uchar> buf;
};

class _InputArray {
public:
    _InputArray(const Mat& m) : m_(m) {}
    bool empty() const { return m_.empty(); }
    Mat getMat() const { return m_; }
private:
    Mat m_;
};
typedef const _InputArray& InputArray;

class _OutputArray {
public:
    _OutputArray(Mat& m) : m_(&m) {}
    void release() const { m_->release(); }
    void create(int r, int c, int t) const { m_->create(r, c, t); }
    Mat getMat() const { return *m_; }
private:
    Mat* m_;
};
typedef const _OutputArray& OutputArray;

inline float fastAtan2(float y, float x) { return orbo_fast_atan2(y, x); }

// cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR): dst keeps its buffer when it already has the requested size (the reference
// relies on this: mvImagePyramid[level] is a view into the bordered buffer `temp`)
inline void resize(const Mat& src, Mat& dst, Size dsize, double, double, int) {
    dst.create(dsize.height, dsize.width, 0);
    orbo_resize(src.data, src.cols, src.rows, (int)src.step.v, dst.data, dst.cols, dst.rows, (int)dst.step.v);
}

// cv::copyMakeBorder(..., BORDER_REFLECT_101 [+ BORDER_ISOLATED]); src may be the interior view of dst
inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int) {
    dst.create(src.rows + top + bottom, src.cols + left + right, 0);
    auto refl = [](int i, int n) { if (n == 1) return 0; while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i; return i; };
    // interior first (a no-op when src already is that interior), then the frame from the interior
    const bool inPlace = src.data == dst.data + (size_t)top * dst.step.v + left;
    if (!inPlace)
        for (int y = 0; y < src.rows; ++y) std::memmove(dst.data + (size_t)(y + top) * dst.step.v + left, src.data + (size_t)y * src.step.v, (size_t)src.cols);
    for (int y = 0; y < dst.rows; ++y) {
        const int sy = refl(y - top, src.rows);
        uchar* d = dst.data + (size_t)y * dst.step.v;
        const uchar* s = dst.data + (size_t)(sy + top) * dst.step.v + left;
        for (int x = 0; x < dst.cols; ++x) {
            const bool interior = y >= top && y < top + src.rows && x >= left && x < left + src.cols;
            if (!interior) d[x] = s[refl(x - left, src.cols)];
        }
    }
}

// cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) on 8-bit data; in-place allowed.  Taps as OpenCV >= 4.5.
inline const int* cvstub_blur_taps() { static int taps[7] = {18, 34, 48, 56, 48, 34, 18}; return taps; }
inline void cvstub_set_blur_taps(const int* t) { std::memcpy(const_cast<int*>(cvstub_blur_taps()), t, 7 * sizeof(int)); }
inline void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sx, double sy, int) {
    assert(ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2);
    Mat in = src.clone();
    dst.create(src.rows, src.cols, 0);
    orbo_blur(in.data, in.cols, in.rows, (int)in.step.v, dst.data, (int)dst.step.v, cvstub_blur_taps());
}

// cv::FAST(image, keypoints, threshold, nonmaxSuppression = true), TYPE_9_16
inline void FAST(const Mat& img, std::vector<KeyPoint>& kps, int threshold, bool nms) {
    assert(nms);
    kps.clear();
    const int cap = img.rows * img.cols / 4 + 16;
    std::vector<int> xys((size_t)cap * 3);
    const int n = orbo_fast(img.data, img.cols, img.rows, (int)img.step.v, threshold, xys.data(), cap);
    for (int i = 0; i < n; ++i) kps.push_back(KeyPoint((float)xys[3 * i], (float)xys[3 * i + 1], 7.f, -1.f, (float)xys[3 * i + 2]));
}

// cv::norm(a, b, NORM_L1) on equally sized 8-bit matrices: sum of absolute differences (as double, like OpenCV)
inline double norm(const Mat& a, const Mat& b, int type) {
    assert(type == NORM_L1 && a.rows == b.rows && a.cols == b.cols);
    long long s = 0;
    for (int y = 0; y < a.rows; ++y) {
        const uchar *pa = a.ptr(y), *pb = b.ptr(y);
        for (int x = 0; x < a.cols; ++x) s += pa[x] > pb[x] ? pa[x] - pb[x] : pb[x] - pa[x];
    }
    return (double)s;
}

struct KeyPointsFilter {  // only referenced by the reference's dead ComputeKeyPointsOld
    static void retainBest(std::vector<KeyPoint>& k, int n) {
        if ((int)k.size() <= n) return;
        std::stable_sort(k.begin(), k.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
        k.resize(n);
    }
};

}  // namespace cv
