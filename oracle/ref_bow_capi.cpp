// ref_bow_capi.cpp — C entry points around the reference's OWN vocabulary code: the vendored DBoW2
// (slam_framework/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h, FORB.cpp, BowVector.cpp, FeatureVector.cpp, ScoringObject.cpp), compiled
// where it lies against oracle/dbowstub + oracle/cvstub -> oracle/_ref/libref_bow.so.  What Frame::ComputeBoW (Frame.cc:768-775) does:
//   vector<cv::Mat> vCurrentDesc = Converter::toDescriptorVector(mDescriptors);  mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4);
// TEST INFRASTRUCTURE ONLY: pins oracle/orb_oracle.cpp's restatement (tests/test_oracle_vs_ref.py).
#include <cstdint>
#include <cstring>
#include <vector>
#include "ORBVocabulary.h"

using TC2LI_SLAM::ORBVocabulary;

extern "C" {
void* refb_load_text(const char* path) {
    ORBVocabulary* v = new ORBVocabulary();
    if (!v->loadFromTextFile(path)) { delete v; return nullptr; }
    return v;
}
void refb_destroy(void* h) { delete (ORBVocabulary*)h; }
int refb_words(void* h) { return (int)((ORBVocabulary*)h)->size(); }
// descriptors: n rows of 32 bytes.  Out: the BowVector as ascending (word, value) pairs, the FeatureVector as CSR over ascending node ids.
int refb_transform(void* h, const uint8_t* desc, int n, int levelsup, uint32_t* words, double* values, int* nWords, uint32_t* nodes, int* nodeStart,
                   uint32_t* features, int* nNodes) {
    ORBVocabulary* voc = (ORBVocabulary*)h;
    std::vector<cv::Mat> v;   // Converter::toDescriptorVector (Converter.cc): one 1 x 32 row per keypoint
    v.reserve(n);
    for (int i = 0; i < n; ++i) {
        cv::Mat m(1, 32, CV_8U);
        std::memcpy(m.data, desc + (size_t)i * 32, 32);
        v.push_back(m);
    }
    DBoW2::BowVector bow;
    DBoW2::FeatureVector fv;
    voc->transform(v, bow, fv, levelsup);
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
}
