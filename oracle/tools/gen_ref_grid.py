#!/usr/bin/env python3
"""Builds oracle/_ref/libref_grid.so: the reference's OWN Frame::AssignFeaturesToGrid, Frame::PosInGrid and
Frame::GetFeaturesInArea (slam_framework/src/Frame.cc), compiled verbatim around a stand-in `Frame` that has exactly the members
those functions touch (Frame.h:59-60, 140, 239-241, 276-278, 313-316, 364-370).  Like gen_ref_stereo.py the bodies are cut out of
the mounted reference at BUILD time into a temporary file OUTSIDE the repository; only the shared object is kept (oracle/_ref is
git-ignored).  TEST INFRASTRUCTURE ONLY.

usage: gen_ref_grid.py /root/reference OUT.so"""
import os, re, subprocess, sys, tempfile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_ref_stereo import cut_function


def main():
    ref, out = sys.argv[1], sys.argv[2]
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    frame = open(os.path.join(ref, "slam_framework/src/Frame.cc")).read()
    header = open(os.path.join(ref, "slam_framework/include/Frame.h")).read()
    defs = "\n".join(l for l in header.splitlines() if re.match(r"\s*#define FRAME_GRID_(ROWS|COLS)\s+\d+", l))
    assert "FRAME_GRID_ROWS" in defs and "FRAME_GRID_COLS" in defs
    bodies = [cut_function(frame, "void Frame::AssignFeaturesToGrid()"),
              cut_function(frame, "bool Frame::PosInGrid(const cv::KeyPoint &kp, int &posX, int &posY)"),
              cut_function(frame, "vector<size_t> Frame::GetFeaturesInArea(")]
    glue_top = r'''
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include "opencv2/cvstub.hpp"
using namespace std;
''' + defs + r'''
namespace TC2LI_SLAM {
struct Frame {   // the members the three functions touch
    int N, Nleft;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    std::vector<std::size_t> mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    void AssignFeaturesToGrid();
    bool PosInGrid(const cv::KeyPoint &kp, int &posX, int &posY);
    vector<size_t> GetFeaturesInArea(const float &x, const float  &y, const float  &r, const int minLevel=-1, const int maxLevel=-1, const bool bRight = false) const;
};
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv, Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
'''
    glue_bottom = r'''
}  // namespace TC2LI_SLAM

using namespace TC2LI_SLAM;
static Frame* make_frame(const void* kps, int n, const void* kpsRight, int nRight, float minX, float maxX, float minY, float maxY) {
    Frame* f = new Frame;
    f->mvKeys.assign((const cv::KeyPoint*)kps, (const cv::KeyPoint*)kps + n);
    if (nRight < 0) {          // one grid: Nleft == -1, keys undistorted (Frame.cc:126-147, 217)
        f->N = n; f->Nleft = -1;
        f->mvKeysUn = f->mvKeys;   // Frame::UndistortKeyPoints with mDistCoef(0) == 0 (Frame.cc:777-782)
    } else {                   // two grids: the fisheye stereo constructor (Frame.cc:1096-1104): N = Nleft + Nright, distorted keys
        f->Nleft = n; f->N = n + nRight;
        f->mvKeysRight.assign((const cv::KeyPoint*)kpsRight, (const cv::KeyPoint*)kpsRight + nRight);
    }
    Frame::mnMinX = minX; Frame::mnMaxX = maxX; Frame::mnMinY = minY; Frame::mnMaxY = maxY;
    // the reference's constructors, Frame.cc:179-180
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (Frame::mnMaxY - Frame::mnMinY);
    f->AssignFeaturesToGrid();
    return f;
}
extern "C" void* refg_create(const void* kps, int n, const void* kpsRight, int nRight, float minX, float maxX, float minY, float maxY) {
    return make_frame(kps, n, kpsRight, nRight, minX, maxX, minY, maxY);
}
extern "C" void refg_destroy(void* h) { delete (Frame*)h; }
extern "C" void refg_grid(void* h, int right, int* cellStart, int* items) {   // mGrid[i][j] / mGridRight[i][j] -> CSR, cell = i * FRAME_GRID_ROWS + j
    Frame* f = (Frame*)h;
    int pos = 0;
    for (int i = 0; i < FRAME_GRID_COLS; ++i)
        for (int j = 0; j < FRAME_GRID_ROWS; ++j) {
            cellStart[i * FRAME_GRID_ROWS + j] = pos;
            for (size_t v : (right ? f->mGridRight[i][j] : f->mGrid[i][j])) items[pos++] = (int)v;
        }
    cellStart[FRAME_GRID_COLS * FRAME_GRID_ROWS] = pos;
}
extern "C" int refg_features_in_area(void* h, int right, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap) {
    const vector<size_t> v = ((Frame*)h)->GetFeaturesInArea(x, y, r, minLevel, maxLevel, right != 0);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int)v[i];
    return (int)v.size();
}
'''
    tmp = tempfile.mkdtemp(prefix="orbx_ref_grid_")
    gen = os.path.join(tmp, "ref_grid_gen.cpp")
    with open(gen, "w") as f:
        f.write(glue_top + "\n".join(bodies) + "\n" + glue_bottom)
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    cmd = ["g++", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-std=c++14", "-w", "-shared", "-o", out, "-I" + os.path.join(here, "cvstub"), gen]
    subprocess.check_call(cmd)
    os.remove(gen)
    os.rmdir(tmp)
    print("built", out)


if __name__ == "__main__":
    main()
