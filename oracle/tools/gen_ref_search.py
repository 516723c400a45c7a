#!/usr/bin/env python3
"""Builds oracle/_ref/libref_search.so: the reference's OWN ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono),
ORBmatcher::ComputeThreeMaxima, ORBmatcher::DescriptorDistance (slam_framework/src/ORBmatcher.cc) and Frame::GetFeaturesInArea /
AssignFeaturesToGrid / PosInGrid (slam_framework/src/Frame.cc), compiled verbatim around stand-ins for the types they touch.
Eigen is not in this image and Sophus needs it, so the pose algebra in front of the search is replaced by stand-ins that make it
the identity: a map point's "world position" is (u, v, z), `Tcw * x` returns x, the camera's project() returns (x(0), x(1)); the
last-frame pose stand-in returns a chosen tlc, which selects the bForward / bBackward / default branch.  Everything from
ORBmatcher.cc:1729 on (windows, level filters, the taken-keypoint skip, the right-coordinate gate, best distance, rotation
histogram, three maxima) runs as written.  As for the other generators the bodies are cut out of the mounted reference at BUILD
time into a temporary file OUTSIDE the repository; only the shared object is kept (oracle/_ref is git-ignored).  TEST INFRASTRUCTURE ONLY.

usage: gen_ref_search.py /root/reference OUT.so"""
import os, re, subprocess, sys, tempfile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_ref_stereo import cut_function


def main():
    ref, out = sys.argv[1], sys.argv[2]
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    frame = open(os.path.join(ref, "slam_framework/src/Frame.cc")).read()
    matcher = open(os.path.join(ref, "slam_framework/src/ORBmatcher.cc")).read()
    header = open(os.path.join(ref, "slam_framework/include/Frame.h")).read()
    defs = "\n".join(l for l in header.splitlines() if re.match(r"\s*#define FRAME_GRID_(ROWS|COLS)\s+\d+", l))
    consts = "\n".join(l.strip() for l in matcher.splitlines() if re.match(r"\s*const int ORBmatcher::(TH_HIGH|TH_LOW|HISTO_LENGTH)\s*=", l))
    assert "TH_HIGH" in consts and "HISTO_LENGTH" in consts
    frame_bodies = [cut_function(frame, "void Frame::AssignFeaturesToGrid()"),
                    cut_function(frame, "bool Frame::PosInGrid(const cv::KeyPoint &kp, int &posX, int &posY)"),
                    cut_function(frame, "vector<size_t> Frame::GetFeaturesInArea(")]
    matcher_bodies = [cut_function(matcher, "int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b)"),
                      cut_function(matcher, "void ORBmatcher::ComputeThreeMaxima(vector<int>* histo, const int L, int &ind1, int &ind2, int &ind3)"),
                      cut_function(matcher, "int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)"),
                      cut_function(matcher, "float ORBmatcher::RadiusByViewingCos(const float &viewCos)"),
                      cut_function(matcher, "int ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, const float th , const int ORBdist)"),
                      cut_function(matcher, "int ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th, const bool bFarPoints, const float thFarPoints)")]
    glue_top = r'''
#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstring>
#include <set>
#include <vector>
#include "opencv2/cvstub.hpp"
using namespace std;
''' + defs + r'''
namespace Eigen {   // stand-ins: plain containers; the only arithmetic is the distance the relocalisation search range-checks
struct Vector3f {
    float d[3];
    float operator()(int i) const { return d[i]; }
    Vector3f operator-(const Vector3f& o) const { return Vector3f{{d[0] - o.d[0], d[1] - o.d[1], d[2] - o.d[2]}}; }
    float norm() const { return std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]); }
};
struct Vector2f { float d[2]; float operator()(int i) const { return d[i]; } };
}
namespace Sophus {  // stand-in: `T * x` is the identity, or a constant for the last-frame pose (-> tlc)
struct SE3f {
    bool constant = false, offset = false;   // offset: T * p = p + t (the right camera's view of a point in the two-grid tests)
    Eigen::Vector3f t = {{0.f, 0.f, 0.f}};
    SE3f inverse() const { return *this; }
    Eigen::Vector3f translation() const { return t; }
    Eigen::Vector3f operator*(const Eigen::Vector3f& p) const { return constant ? t : offset ? Eigen::Vector3f{{p(0) + t(0), p(1) + t(1), p(2) + t(2)}} : p; }
};
}
namespace TC2LI_SLAM {
struct MapPoint {
    Eigen::Vector3f pos; cv::Mat desc; int nObs;
    // what Frame::isInFrustum leaves in a map point for the local-map search (MapPoint.h)
    bool mbTrackInView = false, mbTrackInViewR = false, bad = false;
    float mTrackProjX = 0, mTrackProjY = 0, mTrackDepth = 0, mTrackDepthR = 0, mTrackProjXR = 0, mTrackProjYR = 0, mTrackViewCos = 0, mTrackViewCosR = 0;
    int mnTrackScaleLevel = 0, mnTrackScaleLevelR = -1;
    bool isBad() { return bad; }
    int predictedLevel = 0;     // relocalisation search: what PredictScale returns
    float GetMaxDistanceInvariance() { return 1e30f; }
    float GetMinDistanceInvariance() { return 0.f; }
    template <class F> int PredictScale(const float&, F*) { return predictedLevel; }
    Eigen::Vector3f GetWorldPos() { return pos; }
    cv::Mat GetDescriptor() { return desc; }
    int Observations() { return nObs; }
};
struct GeometricCamera { Eigen::Vector2f project(const Eigen::Vector3f& p) { return Eigen::Vector2f{{p(0), p(1)}}; } };
struct Frame {
    int N, Nleft;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<float> mvuRight, mvScaleFactors;
    std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
    cv::Mat mDescriptors;
    float mb, mbf;
    GeometricCamera* mpCamera;
    Sophus::SE3f mTcw;
    Sophus::SE3f GetPose() const { return mTcw; }
    Sophus::SE3f mTrl;
    Sophus::SE3f GetRelativePoseTrl() { return mTrl; }
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    std::vector<std::size_t> mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    void AssignFeaturesToGrid();
    bool PosInGrid(const cv::KeyPoint &kp, int &posX, int &posY);
    vector<size_t> GetFeaturesInArea(const float &x, const float  &y, const float  &r, const int minLevel=-1, const int maxLevel=-1, const bool bRight = false) const;
};
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv, Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
struct KeyFrame {
    std::vector<MapPoint*> mps;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<MapPoint*> GetMapPointMatches() { return mps; }
};
struct ORBmatcher {
    static const int TH_LOW, TH_HIGH, HISTO_LENGTH;
    float mfNNratio; bool mbCheckOrientation;
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int &ind1, int &ind2, int &ind3);
    int SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono);
    int SearchByProjection(Frame &F, const std::vector<MapPoint*> &vpMapPoints, const float th=3, const bool bFarPoints = false, const float thFarPoints = 50.0f);
    int SearchByProjection(Frame &CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*> &sAlreadyFound, const float th, const int ORBdist);
    float RadiusByViewingCos(const float &viewCos);
};
''' + consts + "\n"
    glue_bottom = r'''
}  // namespace TC2LI_SLAM

using namespace TC2LI_SLAM;
// points: nq x {u, v, z, angle} floats; octaves / nobs: nq ints; pdesc: nq x 32 bytes; tlcz selects the branch (vs mb)
extern "C" int refsp_search(const void* kps, int n, const unsigned char* desc, const float* uright, float minX, float maxX, float minY, float maxY,
                            const float* scaleFactors, int nlevels, float mb, float mbf, const float* points, const int* octaves, const int* nobs,
                            const unsigned char* pdesc, int nq, float th, float tlcz, int bMono, int checkOrientation, int* match) {
    Frame cur, last;
    GeometricCamera cam;
    cur.N = n; cur.Nleft = -1;
    cur.mvKeys.assign((const cv::KeyPoint*)kps, (const cv::KeyPoint*)kps + n);
    cur.mvKeysUn = cur.mvKeys;
    cur.mvpMapPoints.assign(n, nullptr);                 // Tracking::TrackWithMotionModel: fill(..., NULL) before the search
    cur.mvuRight.assign(uright, uright + n);
    cur.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    cur.mDescriptors = cv::Mat(n, 32, CV_8U, (void*)desc, 32);
    cur.mb = mb; cur.mbf = mbf; cur.mpCamera = &cam;
    Frame::mnMinX = minX; Frame::mnMaxX = maxX; Frame::mnMinY = minY; Frame::mnMaxY = maxY;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (Frame::mnMaxY - Frame::mnMinY);
    cur.AssignFeaturesToGrid();
    std::vector<MapPoint> pts(nq);
    last.N = nq; last.Nleft = -1;
    last.mvKeys.resize(nq); last.mvKeysUn.resize(nq);
    last.mvpMapPoints.resize(nq); last.mvbOutlier.assign(nq, false);
    last.mTcw.constant = true; last.mTcw.t = Eigen::Vector3f{{0.f, 0.f, tlcz}};
    for (int i = 0; i < nq; ++i) {
        pts[i].pos = Eigen::Vector3f{{points[4 * i], points[4 * i + 1], points[4 * i + 2]}};
        pts[i].desc = cv::Mat(1, 32, CV_8U, (void*)(pdesc + 32 * (size_t)i), 32);
        pts[i].nObs = nobs[i];
        last.mvpMapPoints[i] = &pts[i];
        last.mvKeys[i].octave = octaves[i]; last.mvKeysUn[i].octave = octaves[i];
        last.mvKeys[i].angle = points[4 * i + 3]; last.mvKeysUn[i].angle = points[4 * i + 3];
    }
    ORBmatcher m;
    m.mfNNratio = 0.9f; m.mbCheckOrientation = checkOrientation != 0;
    const int nm = m.SearchByProjection(cur, last, th, bMono != 0);
    for (int k = 0; k < n; ++k) match[k] = cur.mvpMapPoints[k] ? (int)(cur.mvpMapPoints[k] - pts.data()) : -1;
    return nm;
}
'''
    glue_bottom += r'''
// local-map overload: points: nq x {mTrackProjX, mTrackProjY, mTrackProjXR, mTrackViewCos, mTrackDepth} floats; levels / nobs / flags: nq ints
// (flags: bit 0 = mbTrackInView, bit 1 = isBad()); taken0[k] != 0: F.mvpMapPoints[k] holds a point with observations before the search
extern "C" int refsp_search_local(const void* kps, int n, const unsigned char* desc, const float* uright, const unsigned char* taken0, float minX, float maxX,
                                  float minY, float maxY, const float* scaleFactors, int nlevels, const float* points, const int* levels, const int* nobs,
                                  const int* flags, const unsigned char* pdesc, int nq, float th, float nnRatio, int bFarPoints, float thFarPoints, int* match) {
    Frame F;
    F.N = n; F.Nleft = -1;
    F.mvKeys.assign((const cv::KeyPoint*)kps, (const cv::KeyPoint*)kps + n);
    F.mvKeysUn = F.mvKeys;
    F.mvuRight.assign(uright, uright + n);
    F.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    F.mDescriptors = cv::Mat(n, 32, CV_8U, (void*)desc, 32);
    Frame::mnMinX = minX; Frame::mnMaxX = maxX; Frame::mnMinY = minY; Frame::mnMaxY = maxY;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (Frame::mnMaxY - Frame::mnMinY);
    F.AssignFeaturesToGrid();
    MapPoint held, heldFree;       // earlier matches of the frame: with / without observations
    held.nObs = 1; heldFree.nObs = 0;
    F.mvpMapPoints.assign(n, nullptr);
    for (int k = 0; k < n; ++k) if (taken0[k]) F.mvpMapPoints[k] = taken0[k] == 1 ? &held : &heldFree;
    std::vector<MapPoint> pts(nq);
    std::vector<MapPoint*> vp(nq);
    for (int i = 0; i < nq; ++i) {
        MapPoint& P = pts[i];
        P.mTrackProjX = points[5 * i]; P.mTrackProjY = points[5 * i + 1]; P.mTrackProjXR = points[5 * i + 2]; P.mTrackViewCos = points[5 * i + 3];
        P.mTrackDepth = points[5 * i + 4];
        P.mnTrackScaleLevel = levels[i]; P.nObs = nobs[i];
        P.mbTrackInView = (flags[i] & 1) != 0; P.bad = (flags[i] & 2) != 0;
        P.desc = cv::Mat(1, 32, CV_8U, (void*)(pdesc + 32 * (size_t)i), 32);
        vp[i] = &P;
    }
    ORBmatcher m;
    m.mfNNratio = nnRatio; m.mbCheckOrientation = true;
    const int nm = m.SearchByProjection(F, vp, th, bFarPoints != 0, thFarPoints);
    for (int k = 0; k < n; ++k) {
        MapPoint* p = F.mvpMapPoints[k];
        match[k] = (p && p != &held && p != &heldFree) ? (int)(p - pts.data()) : -1;
    }
    return nm;
}
'''
    glue_bottom += r'''
// relocalisation overload: points: nq x {u, v, z, keyframe keypoint angle} floats; levels: predicted levels; flags: bit 0 = NULL entry of the keyframe,
// bit 1 = isBad(), bit 2 = in sAlreadyFound; taken0[k] != 0: CurrentFrame.mvpMapPoints[k] is non-NULL before the search
extern "C" int refsp_search_kf(const void* kps, int n, const unsigned char* desc, const unsigned char* taken0, float minX, float maxX, float minY, float maxY,
                               const float* scaleFactors, int nlevels, const float* points, const int* levels, const int* flags, const unsigned char* pdesc, int nq,
                               float th, int orbDist, int checkOrientation, int* match) {
    Frame cur;
    GeometricCamera cam;
    cur.N = n; cur.Nleft = -1;
    cur.mvKeys.assign((const cv::KeyPoint*)kps, (const cv::KeyPoint*)kps + n);
    cur.mvKeysUn = cur.mvKeys;
    cur.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    cur.mDescriptors = cv::Mat(n, 32, CV_8U, (void*)desc, 32);
    cur.mpCamera = &cam;
    Frame::mnMinX = minX; Frame::mnMaxX = maxX; Frame::mnMinY = minY; Frame::mnMaxY = maxY;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (Frame::mnMaxY - Frame::mnMinY);
    cur.AssignFeaturesToGrid();
    MapPoint held;
    cur.mvpMapPoints.assign(n, nullptr);
    for (int k = 0; k < n; ++k) if (taken0[k]) cur.mvpMapPoints[k] = &held;
    std::vector<MapPoint> pts(nq);
    KeyFrame kf;
    kf.mps.resize(nq); kf.mvKeysUn.resize(nq);
    std::set<MapPoint*> found;
    for (int i = 0; i < nq; ++i) {
        MapPoint& P = pts[i];
        P.pos = Eigen::Vector3f{{points[4 * i], points[4 * i + 1], points[4 * i + 2]}};
        P.desc = cv::Mat(1, 32, CV_8U, (void*)(pdesc + 32 * (size_t)i), 32);
        P.predictedLevel = levels[i]; P.bad = (flags[i] & 2) != 0; P.nObs = 1;
        kf.mps[i] = (flags[i] & 1) ? nullptr : &P;
        if (flags[i] & 4) found.insert(&P);
        kf.mvKeysUn[i].angle = points[4 * i + 3];
    }
    ORBmatcher m;
    m.mfNNratio = 0.9f; m.mbCheckOrientation = checkOrientation != 0;
    const int nm = m.SearchByProjection(cur, &kf, found, th, orbDist);
    for (int k = 0; k < n; ++k) {
        MapPoint* p = cur.mvpMapPoints[k];
        match[k] = (p && p != &held) ? (int)(p - pts.data()) : -1;
    }
    return nm;
}
'''
    glue_bottom += r'''
// ---- two-grid (Nleft != -1) variants: left keypoints [nL], right keypoints [nR], descLR = the vconcat of both descriptor matrices ----
static void two_grid_frame(Frame& F, const void* kpsL, int nL, const void* kpsR, int nR, const unsigned char* descLR, float minX, float maxX, float minY, float maxY,
                           const float* scaleFactors, int nlevels) {
    F.Nleft = nL; F.N = nL + nR;
    F.mvKeys.assign((const cv::KeyPoint*)kpsL, (const cv::KeyPoint*)kpsL + nL);
    F.mvKeysRight.assign((const cv::KeyPoint*)kpsR, (const cv::KeyPoint*)kpsR + nR);
    F.mvuRight.assign(F.N, -1.f);
    F.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    F.mDescriptors = cv::Mat(F.N, 32, CV_8U, (void*)descLR, 32);
    Frame::mnMinX = minX; Frame::mnMaxX = maxX; Frame::mnMinY = minY; Frame::mnMaxY = maxY;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (Frame::mnMaxY - Frame::mnMinY);
    F.AssignFeaturesToGrid();
}
// frame to frame: the right camera sees a point at (u + dxR, v + dyR) (GetRelativePoseTrl stand-in = that offset); match[N]: left entries first
extern "C" int refsp_search_stereo(const void* kpsL, int nL, const void* kpsR, int nR, const unsigned char* descLR, float minX, float maxX, float minY, float maxY,
                                   const float* scaleFactors, int nlevels, float mb, float mbf, const float* points, const int* octaves, const int* nobs,
                                   const unsigned char* pdesc, int nq, float th, float tlcz, float dxR, float dyR, int checkOrientation, int* match) {
    Frame cur, last;
    GeometricCamera cam;
    two_grid_frame(cur, kpsL, nL, kpsR, nR, descLR, minX, maxX, minY, maxY, scaleFactors, nlevels);
    cur.mvpMapPoints.assign(cur.N, nullptr);
    cur.mb = mb; cur.mbf = mbf; cur.mpCamera = &cam;
    cur.mTrl.offset = true; cur.mTrl.t = Eigen::Vector3f{{dxR, dyR, 0.f}};
    std::vector<MapPoint> pts(nq);
    last.N = nq; last.Nleft = -1;
    last.mvKeys.resize(nq); last.mvKeysUn.resize(nq);
    last.mvpMapPoints.resize(nq); last.mvbOutlier.assign(nq, false);
    last.mTcw.constant = true; last.mTcw.t = Eigen::Vector3f{{0.f, 0.f, tlcz}};
    for (int i = 0; i < nq; ++i) {
        pts[i].pos = Eigen::Vector3f{{points[4 * i], points[4 * i + 1], points[4 * i + 2]}};
        pts[i].desc = cv::Mat(1, 32, CV_8U, (void*)(pdesc + 32 * (size_t)i), 32);
        pts[i].nObs = nobs[i];
        last.mvpMapPoints[i] = &pts[i];
        last.mvKeys[i].octave = octaves[i]; last.mvKeysUn[i].octave = octaves[i];
        last.mvKeys[i].angle = points[4 * i + 3]; last.mvKeysUn[i].angle = points[4 * i + 3];
    }
    ORBmatcher m;
    m.mfNNratio = 0.9f; m.mbCheckOrientation = checkOrientation != 0;
    const int nm = m.SearchByProjection(cur, last, th, false);
    for (int k = 0; k < cur.N; ++k) match[k] = cur.mvpMapPoints[k] ? (int)(cur.mvpMapPoints[k] - pts.data()) : -1;
    return nm;
}
// local map: points nq x {mTrackProjX, mTrackProjY, mTrackViewCos, mTrackDepth, mTrackProjXR, mTrackProjYR, mTrackViewCosR}; levels / levelsR (-1 = none);
// flags bit 0 = mbTrackInView, bit 1 = isBad(), bit 2 = mbTrackInViewR; taken0[N] as in refsp_search_local; l2r[nL] / r2l[nR] = mvLeftToRightMatch / mvRightToLeftMatch
extern "C" int refsp_search_local_stereo(const void* kpsL, int nL, const void* kpsR, int nR, const unsigned char* descLR, const unsigned char* taken0, const int* l2r,
                                         const int* r2l, float minX, float maxX, float minY, float maxY, const float* scaleFactors, int nlevels, const float* points,
                                         const int* levels, const int* levelsR, const int* nobs, const int* flags, const unsigned char* pdesc, int nq, float th,
                                         float nnRatio, int bFarPoints, float thFarPoints, int* match) {
    Frame F;
    two_grid_frame(F, kpsL, nL, kpsR, nR, descLR, minX, maxX, minY, maxY, scaleFactors, nlevels);
    F.mvLeftToRightMatch.assign(l2r, l2r + nL);
    F.mvRightToLeftMatch.assign(r2l, r2l + nR);
    MapPoint held, heldFree;
    held.nObs = 1; heldFree.nObs = 0;
    F.mvpMapPoints.assign(F.N, nullptr);
    for (int k = 0; k < F.N; ++k) if (taken0[k]) F.mvpMapPoints[k] = taken0[k] == 1 ? &held : &heldFree;
    std::vector<MapPoint> pts(nq);
    std::vector<MapPoint*> vp(nq);
    for (int i = 0; i < nq; ++i) {
        MapPoint& P = pts[i];
        const float* p = points + 7 * (size_t)i;
        P.mTrackProjX = p[0]; P.mTrackProjY = p[1]; P.mTrackViewCos = p[2]; P.mTrackDepth = p[3]; P.mTrackProjXR = p[4]; P.mTrackProjYR = p[5]; P.mTrackViewCosR = p[6];
        P.mnTrackScaleLevel = levels[i]; P.mnTrackScaleLevelR = levelsR[i]; P.nObs = nobs[i];
        P.mbTrackInView = (flags[i] & 1) != 0; P.bad = (flags[i] & 2) != 0; P.mbTrackInViewR = (flags[i] & 4) != 0;
        P.desc = cv::Mat(1, 32, CV_8U, (void*)(pdesc + 32 * (size_t)i), 32);
        vp[i] = &P;
    }
    ORBmatcher m;
    m.mfNNratio = nnRatio; m.mbCheckOrientation = true;
    const int nm = m.SearchByProjection(F, vp, th, bFarPoints != 0, thFarPoints);
    for (int k = 0; k < F.N; ++k) {
        MapPoint* p = F.mvpMapPoints[k];
        match[k] = (p && p != &held && p != &heldFree) ? (int)(p - pts.data()) : -1;
    }
    return nm;
}
'''
    tmp = tempfile.mkdtemp(prefix="orbx_ref_search_")
    gen = os.path.join(tmp, "ref_search_gen.cpp")
    with open(gen, "w") as f:
        f.write(glue_top + "\n".join(frame_bodies + matcher_bodies) + "\n" + glue_bottom)
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    cmd = ["g++", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-std=c++14", "-w", "-shared", "-o", out, "-I" + os.path.join(here, "cvstub"), gen]
    try:
        subprocess.check_call(cmd)
    finally:
        os.remove(gen)
        os.rmdir(tmp)
    print("built", out)


if __name__ == "__main__":
    main()
