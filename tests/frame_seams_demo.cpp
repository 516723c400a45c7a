// frame_seams_demo.cpp — the Frame / ORBmatcher seams of INTEGRATION.md 2b, 2d, 2e written the way a maintainer would write them in
// slam_framework/src/Frame.cc and ORBmatcher.cc: two drop-in ORBextractor instances, orbx_stereo_match (Frame::ComputeStereoMatches), the
// feature grid into std::vector<size_t> mGrid[64][48] (Frame::AssignFeaturesToGrid), and both projection searches with the map points
// taken from the frame's own keypoints.  Compiled by tests/test_frame_seams.py against oracle/cvstub (this image has no OpenCV SDK);
// prints checksums that the test compares with the CPU oracle.  With argument "compile-only" it does nothing.
#include <cstdio>
#include <cstring>
#include <map>
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

    // Frame::ComputeBoW (Frame.cc:768-775): mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4) on the resident descriptors, here against a
    // small synthetic vocabulary (k = 4, L = 3) handed over as arrays; the containers are rebuilt the way INTEGRATION.md 2f shows
    {
        const int kk = 4, LL = 3, nNodes = 1 + 4 + 16 + 64;
        std::vector<int> parent(nNodes, 0);
        std::vector<uint8_t> isWord(nNodes, 0), vdesc((size_t)nNodes * 32, 0);
        std::vector<double> weight(nNodes, 0.0);
        for (int i = 1; i < nNodes; ++i) {
            parent[i] = (i - 1) / kk;
            isWord[i] = i >= 21;
            for (int j = 0; j < 32; ++j) vdesc[(size_t)i * 32 + j] = (uint8_t)((i * 131 + j * 29 + (i * j) % 7) & 255);
            weight[i] = isWord[i] ? 0.5 + (i % 13) * 0.37 : 0.0;
        }
        orbx_vocabulary* voc = nullptr;
        if (orbx_vocabulary_create(0, kk, LL, 0, 0, nNodes, parent.data(), isWord.data(), vdesc.data(), weight.data(), &voc) != ORBX_OK) return 7;
        const int K = orbx_batch_max_keypoints(hb);
        std::vector<uint32_t> words(K), nodes(K), feats(K);
        std::vector<double> values(K);
        std::vector<int> start(K + 1);
        int nw = 0, nn = 0;
        if (orbx_batch_compute_bow(hb, voc, 1, 2) != ORBX_OK) return 8;
        if (orbx_batch_download_bow(hb, 1, words.data(), values.data(), &nw, nodes.data(), start.data(), feats.data(), &nn) != ORBX_OK) return 9;
        std::map<unsigned, double> mBowVec;
        std::map<unsigned, std::vector<unsigned> > mFeatVec;
        for (int i = 0; i < nw; ++i) mBowVec.insert(mBowVec.end(), std::make_pair(words[i], values[i]));
        for (int j = 0; j < nn; ++j) mFeatVec.insert(mFeatVec.end(), std::make_pair(nodes[j], std::vector<unsigned>(feats.begin() + start[j], feats.begin() + start[j + 1])));
        unsigned long long wsum = 0, vsum = 0, fsum = 0;
        for (std::map<unsigned, double>::const_iterator it = mBowVec.begin(); it != mBowVec.end(); ++it) {
            unsigned long long bits;
            std::memcpy(&bits, &it->second, 8);
            wsum += (unsigned long long)(it->first + 1);
            vsum += bits >> 20;   // (the sum of the top 44 bits of every value: any differing value changes it)
        }
        for (std::map<unsigned, std::vector<unsigned> >::const_iterator it = mFeatVec.begin(); it != mFeatVec.end(); ++it)
            for (size_t j = 0; j < it->second.size(); ++j) fsum += (unsigned long long)(it->first + 1) * (it->second[j] + 1) * (j + 1);
        std::printf("bow %d %d %llu %llu %llu\n", nw, nn, wsum, vsum, fsum);
        orbx_vocabulary_destroy(voc);
    }
    return 0;
}
