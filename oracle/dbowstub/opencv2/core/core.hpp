// stand-in for <opencv2/core/core.hpp> as the vendored DBoW2 (slam_framework/Thirdparty/DBoW2) sees it: cv::Mat from oracle/cvstub plus
// a cv::FileStorage / cv::FileNode that compile and never open anything (the YAML save / load members are virtual, hence instantiated;
// the tests go through loadFromTextFile, the loader ORB-SLAM itself uses for ORBvoc.txt).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>   // (the real header pulls these in; DBoW2 relies on it)
#include <string>
#include "opencv2/cvstub.hpp"
#ifndef CV_32F
#define CV_32F 5
#endif
namespace cv {
struct FileNode {
    size_t size() const { return 0; }
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    operator int() const { return 0; }
    operator double() const { return 0.0; }
    operator std::string() const { return std::string(); }
};
struct FileStorage {
    enum { READ = 0, WRITE = 1 };
    FileStorage(const char*, int) {}
    bool isOpened() const { return false; }
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](const char*) const { return FileNode(); }
    void release() {}
};
template <class T> FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }
}  // namespace cv
