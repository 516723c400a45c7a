// compat_demo.cpp — uses the drop-in ORBextractor.h exactly the way Frame::ExtractORB does (Frame.cc:445-455, two threads as
// Frame.cc:139-142).  Compiled by tests/test_compat_shim.py against the stand-in OpenCV headers of oracle/cvstub (this image
// has no OpenCV SDK).  Prints "N monoIndex sum(desc)" per image; with argument "compile-only" it does nothing.
#include <cstdio>
#include <cstring>
#include <thread>

#include "ORBextractor.h"

extern "C" void orbo_synth_v1(unsigned char* out, int W, int H, unsigned seed, int xoff, unsigned noise_xor);

int main(int argc, char** argv) {
    if (argc > 1 && !std::strcmp(argv[1], "compile-only")) return 0;
    const int W = 1241, H = 376;
    cv::Mat imL(H, W, CV_8UC1), imR(H, W, CV_8UC1);
    orbo_synth_v1(imL.data, W, H, 42, 0, 0);
    orbo_synth_v1(imR.data, W, H, 42, 16, 0x5bd1e995u);
    TC2LI_SLAM::ORBextractor* left = new TC2LI_SLAM::ORBextractor(2000, 1.2f, 8, 20, 7);
    TC2LI_SLAM::ORBextractor* right = new TC2LI_SLAM::ORBextractor(2000, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> kL, kR;
    cv::Mat dL, dR;
    int mL = 0, mR = 0;
    auto run = [](TC2LI_SLAM::ORBextractor* e, const cv::Mat& im, std::vector<cv::KeyPoint>* k, cv::Mat* d, int* mono) {
        std::vector<int> lap = {0, 0};
        *mono = (*e)(im, cv::Mat(), *k, *d, lap);
    };
    std::thread tl(run, left, std::cref(imL), &kL, &dL, &mL), tr(run, right, std::cref(imR), &kR, &dR, &mR);
    tl.join();
    tr.join();
    for (int s = 0; s < 2; ++s) {
        const std::vector<cv::KeyPoint>& k = s ? kR : kL;
        const cv::Mat& d = s ? dR : dL;
        unsigned long long sum = 0;
        for (int i = 0; i < d.rows; ++i)
            for (int j = 0; j < 32; ++j) sum += d.ptr(i)[j] * (unsigned long long)(i * 32 + j + 1);
        double sx = 0;
        for (const cv::KeyPoint& p : k) sx += p.pt.x + 2 * p.pt.y + 3 * p.angle + p.octave;
        std::printf("%zu %d %llu %.6f %d %d\n", k.size(), s ? mR : mL, sum, sx, left->mvImagePyramid[7].cols, left->mvImagePyramid[7].rows);
    }
    std::printf("levels %d scale %.6f\n", left->GetLevels(), left->GetScaleFactors()[7]);
    return 0;
}
