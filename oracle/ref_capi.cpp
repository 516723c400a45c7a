// ref_capi.cpp — C entry points around the reference's own TC2LI_SLAM::ORBextractor, compiled verbatim from
// /root/reference/slam_framework/src/ORBextractor.cc against oracle/cvstub.  TEST INFRASTRUCTURE ONLY (oracle/_ref).
#include <cstring>
#include <vector>

#include "ORBextractor.h"

extern "C" {
void* refo_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    return new TC2LI_SLAM::ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void refo_destroy(void* h) { delete (TC2LI_SLAM::ORBextractor*)h; }
void refo_set_blur_taps(const int* taps) { cv::cvstub_set_blur_taps(taps); }

struct RefResult {
    std::vector<cv::KeyPoint> kps;
    cv::Mat desc;
};
static thread_local RefResult g_res;

// System::TrackStereoLidar hands the extractor a clone (continuous Mat); so do we.  Returns operator()'s return value.
int refo_extract(void* h, const unsigned char* img, int w, int hh, int stride, int lap0, int lap1) {
    TC2LI_SLAM::ORBextractor* e = (TC2LI_SLAM::ORBextractor*)h;
    cv::Mat in;
    if (img && w > 0 && hh > 0) in = cv::Mat(hh, w, CV_8UC1, (void*)img, (size_t)stride).clone();
    std::vector<int> lap = {lap0, lap1};
    g_res.kps.clear();
    g_res.desc = cv::Mat();
    return (*e)(in, cv::Mat(), g_res.kps, g_res.desc, lap);
}
int refo_num_keypoints() { return (int)g_res.kps.size(); }
void refo_get_keypoints(void* out) { if (!g_res.kps.empty()) std::memcpy(out, g_res.kps.data(), g_res.kps.size() * sizeof(cv::KeyPoint)); }
void refo_get_descriptors(unsigned char* out) {
    for (int y = 0; y < g_res.desc.rows; ++y) std::memcpy(out + (size_t)y * 32, g_res.desc.ptr(y), 32);
}
void refo_tables(void* h, float* sf, float* isf, float* sig2, float* isig2) {
    TC2LI_SLAM::ORBextractor* e = (TC2LI_SLAM::ORBextractor*)h;
    const int n = e->GetLevels();
    std::vector<float> a = e->GetScaleFactors(), b = e->GetInverseScaleFactors(), c = e->GetScaleSigmaSquares(), d = e->GetInverseScaleSigmaSquares();
    for (int i = 0; i < n; ++i) { sf[i] = a[i]; isf[i] = b[i]; sig2[i] = c[i]; isig2[i] = d[i]; }
}
void refo_level_size(void* h, int l, int* w, int* hh) {
    TC2LI_SLAM::ORBextractor* e = (TC2LI_SLAM::ORBextractor*)h;
    *w = e->mvImagePyramid[l].cols; *hh = e->mvImagePyramid[l].rows;
}
// level l with `border` pixels of the buffer around it (0 or 19: mvImagePyramid[l] is the interior view of a bordered Mat)
void refo_get_level(void* h, int l, int border, unsigned char* out) {
    TC2LI_SLAM::ORBextractor* e = (TC2LI_SLAM::ORBextractor*)h;
    const cv::Mat& m = e->mvImagePyramid[l];
    const int W = m.cols + 2 * border;
    for (int y = -border; y < m.rows + border; ++y)
        std::memcpy(out + (size_t)(y + border) * W, m.data + (ptrdiff_t)y * (ptrdiff_t)(size_t)m.step - border, (size_t)W);
}
}
