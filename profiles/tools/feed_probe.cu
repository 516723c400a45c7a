// feed_probe.cu — the host <-> device feed ceiling of this box for the e2e path: aggregate H2D-only, D2H-only and duplex copy
// bandwidth with 1 / 2 / 4 / 8 GPUs fed at once (one host thread per GPU, pinned buffers of the benchmark's chunk sizes, two copies
// in flight per direction).  Prints one JSON object.   nvcc -O2 -std=c++17 -o feed_probe feed_probe.cu
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

static const size_t kInBytes = (size_t)128 * 1241 * 376;          // one chunk of 128 KITTI images
static const size_t kOutBytes = (size_t)128 * 2024 * (28 + 32);   // its keypoints + descriptors

struct Dev {
    int id;
    unsigned char *hin[2], *hout[2], *din[2], *dout[2];
    cudaStream_t up, down;
};

static double run(std::vector<Dev>& devs, int mode /*1 h2d, 2 d2h, 3 both*/, double seconds, double* h2d_gbs, double* d2h_gbs) {
    const int n = (int)devs.size();
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    std::vector<double> upB(n, 0), downB(n, 0), secs(n, 0);
    auto worker = [&](int k) {
        Dev& d = devs[k];
        cudaSetDevice(d.id);
        ready.fetch_add(1);
        while (!go.load()) {}
        const auto t0 = std::chrono::steady_clock::now();
        double up = 0, down = 0;
        int it = 0;
        for (;;) {
            if (mode & 1) { cudaMemcpyAsync(d.din[it & 1], d.hin[it & 1], kInBytes, cudaMemcpyHostToDevice, d.up); up += kInBytes; }
            if (mode & 2) { cudaMemcpyAsync(d.hout[it & 1], d.dout[it & 1], kOutBytes, cudaMemcpyDeviceToHost, d.down); down += kOutBytes; }
            ++it;
            if ((it & 1) == 0) {  // two copies per direction in flight, then wait for them
                cudaStreamSynchronize(d.up);
                cudaStreamSynchronize(d.down);
                if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > seconds) break;
            }
        }
        cudaStreamSynchronize(d.up);
        cudaStreamSynchronize(d.down);
        secs[k] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        upB[k] = up; downB[k] = down;
    };
    std::vector<std::thread> th;
    for (int k = 0; k < n; ++k) th.emplace_back(worker, k);
    while (ready.load() < n) {}
    go.store(true);
    for (auto& t : th) t.join();
    double u = 0, dn = 0, smax = 0;
    for (int k = 0; k < n; ++k) { u += upB[k] / secs[k]; dn += downB[k] / secs[k]; smax = secs[k] > smax ? secs[k] : smax; }
    *h2d_gbs = u / 1e9; *d2h_gbs = dn / 1e9;
    return smax;
}

int main(int argc, char** argv) {
    const double seconds = argc > 1 ? atof(argv[1]) : 0.4;
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    std::vector<Dev> all(ndev);
    for (int i = 0; i < ndev; ++i) {
        Dev& d = all[i];
        d.id = i;
        cudaSetDevice(i);
        for (int k = 0; k < 2; ++k) {
            if (cudaHostAlloc((void**)&d.hin[k], kInBytes, cudaHostAllocPortable) != cudaSuccess || cudaHostAlloc((void**)&d.hout[k], kOutBytes, cudaHostAllocPortable) != cudaSuccess ||
                cudaMalloc((void**)&d.din[k], kInBytes) != cudaSuccess || cudaMalloc((void**)&d.dout[k], kOutBytes) != cudaSuccess) { printf("{\"error\": \"allocation failed on device %d\"}\n", i); return 1; }
            for (size_t j = 0; j < kInBytes; j += 4096) d.hin[k][j] = (unsigned char)j;   // touch the pages
        }
        cudaStreamCreateWithFlags(&d.up, cudaStreamNonBlocking);
        cudaStreamCreateWithFlags(&d.down, cudaStreamNonBlocking);
    }
    std::vector<std::vector<int>> sets = {{0}};
    const int only = argc > 2 ? atoi(argv[2]) : 0;   // > 0: just the set {0 .. only - 1} (bench.py's live ceiling)
    if (only > 0) {
        sets.clear();
        std::vector<int> s;
        for (int i = 0; i < only && i < ndev; ++i) s.push_back(i);
        sets.push_back(s);
        ndev = 0;  // (skips the fixed sets below)
    }
    if (ndev >= 2) { sets.push_back({0, 1}); }
    if (ndev >= 8) { sets.push_back({0, 4}); sets.push_back({0, 7}); }
    if (ndev >= 4) { sets.push_back({0, 1, 2, 3}); }
    if (ndev >= 8) { sets.push_back({0, 2, 4, 6}); sets.push_back({4, 5, 6, 7}); sets.push_back({0, 1, 2, 3, 4, 5, 6, 7}); }
    cudaGetDeviceCount(&ndev);
    printf("{\"devices\": %d, \"chunk_in_bytes\": %zu, \"chunk_out_bytes\": %zu, \"seconds_per_run\": %.2f, \"runs\": [\n", ndev, kInBytes, kOutBytes, seconds);
    bool first = true;
    for (auto& s : sets) {
        std::vector<Dev> devs;
        std::string name;
        for (int id : s) { devs.push_back(all[id]); name += (name.empty() ? "" : ",") + std::to_string(id); }
        double r[3][2];
        for (int mode = 1; mode <= 3; ++mode) {
            double a, b;
            run(devs, mode, 0.1, &a, &b);  // warm-up
            run(devs, mode, seconds, &r[mode - 1][0], &r[mode - 1][1]);
        }
        printf("%s {\"gpus\": [%s], \"n\": %zu, \"h2d_only_GBps\": %.1f, \"d2h_only_GBps\": %.1f, \"duplex_h2d_GBps\": %.1f, \"duplex_d2h_GBps\": %.1f}", first ? "" : ",\n", name.c_str(), s.size(),
               r[0][0], r[1][1], r[2][0], r[2][1]);
        first = false;
    }
    printf("\n]}\n");
    return 0;
}
