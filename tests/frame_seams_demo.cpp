// frame_seams_demo.cpp — the Frame / ORBmatcher seams of INTEGRATION.md 2b, 2d, 2e written the way a maintainer would write them in
// slam_framework/src/Frame.cc and ORBmatcher.cc: two drop-in ORBextractor instances, orbx_stereo_match (Frame::ComputeStereoMatches), the
// feature grid into std::vector<size_t> mGrid[64][48] (Frame::AssignFeaturesToGrid), and both projection searches with the map points
// taken from the frame's own keypoints.  Compiled by tests/test_frame_seams.py against oracle/cvstub (this image has no OpenCV SDK);
// prints checksums that the test compares with the CPU oracle.  With argument "compile-only" it does nothing.
#include <cstdio>
#include <cstring>
#include <vector>

#include "ORBextractor.h"

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64
static_assert(FRAME_GRID_ROWS == ORBX_GRID_ROWS && FRAME_GRID_COLS == ORBX_GRID_COLS, "Frame.h:59-60");

extern "C" void orbo_synth_v1(unsigned char* out, int W, int H, unsigned seed, int xoff, unsigned noise_xor);

int main(int argc, char** argv) {
    if (argc > 1 && !std::strcmp(argv[1], "compile-only")) return 0;
    const int W = 1241, H = 376;
    const float mb = 0.5372f, mbf = 387.5744f;
    cv::Mat imL(H, W, CV_8UC1), imR(H, W, CV_8UC1);
    orbo_synth_v1(imL.data, W, H, 77, 0, 0);
    orbo_synth_v1(imR.data, W, H, 77, 16, 0x5bd1e995u);
    TC2LI_SLAM::ORBextractor left(2000, 1.2f, 8, 20, 7), right(2000, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
    cv::Mat mDescriptors, mDescriptorsRight;
    std::vector<int> lap = {0, 0};
    left(imL, cv::Mat(), mvKeys, mDescriptors, lap);
    right(imR, cv::Mat(), mvKeysRight, mDescriptorsRight, lap);
    const int N = (int)mvKeys.size();

    // Frame::ComputeStereoMatches (Frame.cc:841-1011)
    std::vector<float> mvuRight(N), mvDepth(N);
    int nLeft = 0;
    if (orbx_stereo_match(left.handle(), right.handle(), mb, mbf, mvuRight.data(), mvDepth.data(), N, &nLeft) != ORBX_OK || nLeft != N) return 2;

    // Frame::AssignFeaturesToGrid (Frame.cc:412-443): bounds of Frame::ComputeImageBounds without distortion (Frame.cc:834-837)
    orbx_batch* hb = orbx_extractor_batch(left.handle());
    if (orbx_batch_assign_grid(hb, 1, 0.0f, (float)W, 0.0f, (float)H) != ORBX_OK) return 3;
    std::vector<int> cellStart(ORBX_GRID_CELLS + 1), items(orbx_batch_max_keypoints(hb));
    if (orbx_batch_download_grid(hb, 1, cellStart.data(), items.data()) != ORBX_OK) return 4;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    for (int i = 0; i < FRAME_GRID_COLS; ++i)
        for (int j = 0; j < FRAME_GRID_ROWS; ++j)
            mGrid[i][j].assign(items.begin() + cellStart[i * FRAME_GRID_ROWS + j], items.begin() + cellStart[i * FRAME_GRID_ROWS + j + 1]);
    unsigned long long gsum = 0, gtot = 0;
    for (int i = 0; i < FRAME_GRID_COLS; ++i)
        for (int j = 0; j < FRAME_GRID_ROWS; ++j)
            for (std::size_t k = 0; k < mGrid[i][j].size(); ++k) { gsum += (unsigned long long)(i * FRAME_GRID_ROWS + j + 1) * (mGrid[i][j][k] + 1) * (k + 1); ++gtot; }
    std::printf("grid %llu %llu\n", gtot, gsum);

    // ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th = 15, bMono = false): one query per "map point" (here: the frame's own keypoints)
    const std::vector<float> mvScaleFactors = left.GetScaleFactors();
    std::vector<orbx_proj_query> q(N);
    for (int i = 0; i < N; ++i) {
        const int o = mvKeys[i].octave;
        orbx_proj_query e;
        e.u = mvKeys[i].pt.x; e.v = mvKeys[i].pt.y; e.radius = 15.0f * mvScaleFactors[o]; e.ur = mvKeys[i].pt.x - 1.0f;
        e.min_level = o - 1; e.max_level = o + 1; e.angle = mvKeys[i].angle; e.blocks = 1;
        std::memcpy(e.desc, mDescriptors.ptr(i), 32);
        q[i] = e;
    }
    std::vector<int> match(orbx_batch_max_keypoints(hb));
    int nmatches = 0;
    if (orbx_batch_search_by_projection(hb, 0, q.data(), N, 1, 1, 100, match.data(), &nmatches) != ORBX_OK) return 5;
    long long msum = 0;
    for (int k = 0; k < N; ++k) msum += (long long)(match[k] + 2) * (k + 1);
    std::printf("frame %d %lld\n", nmatches, msum);

    // ORBmatcher::SearchByProjection(F, vpMapPoints, th = 3): r = RadiusByViewingCos(0.9) * th, window [o - 1, o]
    for (int i = 0; i < N; ++i) { const int o = mvKeys[i].octave; q[i].radius = 4.0f * 3.0f * mvScaleFactors[o]; q[i].min_level = o - 1; q[i].max_level = o; }
    std::vector<uint8_t> taken(orbx_batch_max_keypoints(hb), 0);
    for (int k = 0; k < N; k += 5) taken[k] = 1;
    if (orbx_batch_search_local_points(hb, 0, q.data(), N, 1, 0.8f, 100, taken.data(), match.data(), &nmatches) != ORBX_OK) return 6;
    msum = 0;
    for (int k = 0; k < N; ++k) msum += (long long)(match[k] + 2) * (k + 1);
    std::printf("local %d %lld\n", nmatches, msum);
    return 0;
}
