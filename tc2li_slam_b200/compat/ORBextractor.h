/* ORBextractor.h — drop-in replacement for TC2LI-SLAM's slam_framework/include/ORBextractor.h (+ src/ORBextractor.cc).
 *
 * Same namespace, class name, constructor, operator(), getters and public mvImagePyramid as the reference
 * (ORBextractor.h:52-118), so Frame / Tracking / ORBmatcher compile and behave unchanged; the work is done by the
 * sm_100a CUDA path behind the C ABI of include/orbx.h (link liborbx.so).  Header-only; needs the application's own
 * OpenCV headers for cv::Mat / cv::KeyPoint.  See INTEGRATION.md.
 *
 * Differences a maintainer should know about:
 *   - no CPU fallback: without a usable GPU the constructor throws std::runtime_error;
 *   - images whose smallest pyramid level is below one 35-px FAST cell make the reference divide by zero; here
 *     operator() returns -1 and leaves the outputs empty;
 *   - device selection: environment variable ORBX_DEVICE (default 0);
 *   - mvImagePyramid is refreshed eagerly after every call (one device->host copy per level, with the 19-px
 *     REFLECT_101 border of ORBextractor.cc:1149-1165); define ORBX_LAZY_PYRAMID to skip that and call SyncPyramid()
 *     only where the levels are read (Frame::ComputeStereoMatches, Frame.cc:848,938,948,953).
 */
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <cstring>
#include <stdexcept>
#include <string>
#include <cassert>
#include <cstdlib>
#include <list>
#include <stdexcept>
#include <string>
#include <vector>

#include <opencv2/opencv.hpp>

#include "orbx.h"

namespace TC2LI_SLAM {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures_, float scaleFactor_, int nlevels_, int iniThFAST_, int minThFAST_)
        : nfeatures(nfeatures_), scaleFactor(scaleFactor_), nlevels(nlevels_), iniThFAST(iniThFAST_), minThFAST(minThFAST_), h_(nullptr) {
        static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "cv::KeyPoint layout");
        const char* dev = std::getenv("ORBX_DEVICE");
        const int rc = orbx_create(nfeatures, scaleFactor_, nlevels, iniThFAST, minThFAST, dev ? std::atoi(dev) : 0, &h_);
        if (rc != ORBX_OK) throw std::runtime_error(std::string("ORBextractor (orbx): ") + orbx_last_error());
        mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        mnFeaturesPerLevel.resize(nlevels);
        orbx_get_scale_factors(h_, mvScaleFactor.data());
        orbx_get_inverse_scale_factors(h_, mvInvScaleFactor.data());
        orbx_get_scale_sigma_squares(h_, mvLevelSigma2.data());
        orbx_get_inverse_scale_sigma_squares(h_, mvInvLevelSigma2.data());
        orbx_get_features_per_level(h_, mnFeaturesPerLevel.data());
        mvImagePyramid.resize(nlevels);
    }
    ~ORBextractor() { orbx_destroy(h_); }
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image.  Mask is ignored, as in the reference.
    int operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors,
                   std::vector<int>& vLappingArea) {
        if (_image.empty()) return -1;
        cv::Mat image = _image.getMat();
        assert(image.type() == CV_8UC1);
        const int cap = orbx_max_keypoints(h_);
        kps_.resize((size_t)cap);
        desc_.resize((size_t)cap * 32);
        int n = 0, mono = 0;
        const int rc = orbx_extract(h_, image.data, image.cols, image.rows, (size_t)image.step, vLappingArea[0], vLappingArea[1], kps_.data(),
                                    desc_.data(), cap, &n, &mono);
        if (rc != ORBX_OK) {
            _keypoints.clear();
            _descriptors.release();
            // an empty / too small image is the reference's "-1"; a device fault or an exceeded capacity must not look like "no features"
            if (rc == ORBX_ERR_EMPTY || rc == ORBX_ERR_TOO_SMALL) return -1;
            throw std::runtime_error(std::string("orbx: ") + orbx_last_error());
        }
        _keypoints = std::vector<cv::KeyPoint>((size_t)n);
        if (n == 0)
            _descriptors.release();
        else {
            std::memcpy((void*)_keypoints.data(), kps_.data(), (size_t)n * sizeof(orbx_keypoint));
            _descriptors.create(n, 32, CV_8U);
            cv::Mat d = _descriptors.getMat();
            for (int i = 0; i < n; ++i) std::memcpy(d.ptr(i), desc_.data() + (size_t)i * 32, 32);
        }
#ifndef ORBX_LAZY_PYRAMID
        SyncPyramid();
#endif
        return mono;
    }

    // refresh mvImagePyramid from the device: each level is the interior view of a Mat carrying the 19-px border
    void SyncPyramid() {
        for (int l = 0; l < nlevels; ++l) {
            int w = 0, hgt = 0;
            if (orbx_get_level_size(h_, l, &w, &hgt) != ORBX_OK) return;
            cv::Mat temp(cv::Size(w + 38, hgt + 38), CV_8UC1);
            orbx_get_level(h_, l, temp.data, (size_t)temp.step, 19);
            mvImagePyramid[l] = temp(cv::Rect(19, 19, w, hgt));
        }
    }

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    std::vector<cv::Mat> mvImagePyramid;

    orbx_extractor* handle() { return h_; }  // for orbx_extract_stereo (both images issued before either is awaited)

protected:
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

private:
    orbx_extractor* h_;
    std::vector<orbx_keypoint> kps_;
    std::vector<unsigned char> desc_;
};

}  // namespace TC2LI_SLAM

#endif  // ORBEXTRACTOR_H
