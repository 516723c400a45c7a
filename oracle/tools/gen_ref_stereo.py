#!/usr/bin/env python3
"""Builds oracle/_ref/libref_stereo.so: the reference's OWN Frame::ComputeStereoMatches (slam_framework/src/Frame.cc) and
ORBmatcher::DescriptorDistance (slam_framework/src/ORBmatcher.cc), compiled verbatim around a stand-in `Frame` that has
exactly the members the function touches.  Frame.cc as a whole cannot be compiled here (ROS / Eigen / g2o / DBoW2 headers are
absent), so the two function bodies are cut out of the mounted reference at BUILD time into a temporary file OUTSIDE the
repository, compiled, and only the shared object is kept (oracle/_ref is git-ignored).  TEST INFRASTRUCTURE ONLY.

usage: gen_ref_stereo.py /root/reference OUT.so"""
import os, re, subprocess, sys, tempfile


def cut_function(src, signature):
    i = src.index(signature)
    j = src.index("{", i)
    depth, k = 0, j
    while True:
        c = src[k]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return src[i:k + 1]
        k += 1


def main():
    ref, out = sys.argv[1], sys.argv[2]
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    frame = open(os.path.join(ref, "slam_framework/src/Frame.cc")).read()
    matcher = open(os.path.join(ref, "slam_framework/src/ORBmatcher.cc")).read()
    body_match = cut_function(frame, "void Frame::ComputeStereoMatches()")
    body_dist = cut_function(matcher, "int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b)")
    consts = "\n".join(l for l in matcher.splitlines() if re.match(r"\s*const int ORBmatcher::(TH_HIGH|TH_LOW|HISTO_LENGTH)\s*=", l))
    assert "TH_HIGH" in consts and "TH_LOW" in consts
    glue_top = r'''
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <vector>
#include "opencv2/cvstub.hpp"
using namespace std;
namespace TC2LI_SLAM {
struct ORBextractor { std::vector<cv::Mat> mvImagePyramid; };
struct ORBmatcher {
    static const int TH_LOW, TH_HIGH, HISTO_LENGTH;
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
};
struct Frame {   // the members Frame::ComputeStereoMatches touches (Frame.h:140, 239-261, 307-308)
    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
    cv::Mat mDescriptors, mDescriptorsRight;
    std::vector<float> mvuRight, mvDepth, mvScaleFactors, mvInvScaleFactors;
    float mbf, mb;
    ORBextractor *mpORBextractorLeft, *mpORBextractorRight;
    void ComputeStereoMatches();
};
'''
    glue_bottom = r'''
}  // namespace TC2LI_SLAM

extern "C" int refs_stereo_match(const void* kL, int nL, const unsigned char* dL, const void* kR, int nR, const unsigned char* dR, int nlevels,
                                 const unsigned char* const* pyrL, const unsigned char* const* pyrR, const int* w, const int* h, const float* sf,
                                 const float* isf, float mb, float mbf, float* uright, float* depth) {
    using namespace TC2LI_SLAM;
    Frame f;
    ORBextractor eL, eR;
    f.N = nL;
    f.mvKeys.assign((const cv::KeyPoint*)kL, (const cv::KeyPoint*)kL + nL);
    f.mvKeysRight.assign((const cv::KeyPoint*)kR, (const cv::KeyPoint*)kR + nR);
    f.mDescriptors = cv::Mat(nL, 32, CV_8U, (void*)dL, 32);
    f.mDescriptorsRight = cv::Mat(nR, 32, CV_8U, (void*)dR, 32);
    f.mvScaleFactors.assign(sf, sf + nlevels);
    f.mvInvScaleFactors.assign(isf, isf + nlevels);
    f.mb = mb; f.mbf = mbf;
    for (int l = 0; l < nlevels; ++l) {
        eL.mvImagePyramid.push_back(cv::Mat(h[l], w[l], CV_8UC1, (void*)pyrL[l], (size_t)w[l]));
        eR.mvImagePyramid.push_back(cv::Mat(h[l], w[l], CV_8UC1, (void*)pyrR[l], (size_t)w[l]));
    }
    f.mpORBextractorLeft = &eL; f.mpORBextractorRight = &eR;
    f.ComputeStereoMatches();
    std::memcpy(uright, f.mvuRight.data(), (size_t)nL * 4);
    std::memcpy(depth, f.mvDepth.data(), (size_t)nL * 4);
    return nL;
}
'''
    tmp = tempfile.mkdtemp(prefix="orbx_ref_stereo_")
    gen = os.path.join(tmp, "ref_stereo_gen.cpp")
    with open(gen, "w") as f:
        f.write(glue_top + consts + "\n" + body_dist + "\n" + body_match + "\n" + glue_bottom)
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    cmd = ["g++", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-std=c++14", "-w", "-shared", "-o", out, "-I" + os.path.join(here, "cvstub"), gen,
           "-L" + here, "-lorb_oracle", "-Wl,-rpath,$ORIGIN/.."]
    subprocess.check_call(cmd)
    os.remove(gen)
    os.rmdir(tmp)
    print("built", out)


if __name__ == "__main__":
    main()
