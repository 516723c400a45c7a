// feed_probe.cu — the host <-> device feed ceiling of this box for the e2e path: aggregate H2D-only, D2H-only and duplex copy
// bandwidth with 1 / 2 / 4 / 8 GPUs fed at once, pinned buffers of the benchmark's chunk sizes, two copies in flight per direction.
// One host PROCESS per GPU by default (fork before any CUDA call): that is how the benchmark's ranks feed their GPUs, and it is
// faster than one process with a thread per GPU on this driver (mode "threads": profiles/r02/feed_ceiling_one_process.json).
//   feed_probe [seconds] [all | comma separated device list] [threads]        prints one JSON object
//   nvcc -O2 -std=c++17 -o feed_probe feed_probe.cu
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>
#include <cuda_runtime.h>

static const size_t kInBytes = (size_t)128 * 1241 * 376;          // one chunk of 128 KITTI images
static const size_t kOutBytes = (size_t)128 * 2024 * (28 + 32);   // its keypoints + descriptors

struct Shared {   // lives in an anonymous shared mapping
    std::atomic<int> arrived[64];
    std::atomic<int> failed;
    double up[8], down[8], secs[8];
};

struct Bufs {
    unsigned char *hin[2], *hout[2], *din[2], *dout[2];
    cudaStream_t up, down;
};

static bool setup(int dev, Bufs& b) {
    if (cudaSetDevice(dev) != cudaSuccess) return false;
    for (int k = 0; k < 2; ++k) {
        if (cudaHostAlloc((void**)&b.hin[k], kInBytes, cudaHostAllocPortable) != cudaSuccess || cudaHostAlloc((void**)&b.hout[k], kOutBytes, cudaHostAllocPortable) != cudaSuccess ||
            cudaMalloc((void**)&b.din[k], kInBytes) != cudaSuccess || cudaMalloc((void**)&b.dout[k], kOutBytes) != cudaSuccess)
            return false;
        for (size_t j = 0; j < kInBytes; j += 4096) b.hin[k][j] = (unsigned char)j;   // touch the pages
    }
    return cudaStreamCreateWithFlags(&b.up, cudaStreamNonBlocking) == cudaSuccess && cudaStreamCreateWithFlags(&b.down, cudaStreamNonBlocking) == cudaSuccess;
}

static void meet(Shared* sh, int phase, int n) {   // all n participants
    sh->arrived[phase].fetch_add(1);
    while (sh->arrived[phase].load() < n) {}
}

static void copy_loop(Bufs& d, int mode, double seconds, double* up, double* down, double* secs) {
    const auto t0 = std::chrono::steady_clock::now();
    double u = 0, dn = 0;
    int it = 0;
    for (;;) {
        if (mode & 1) { cudaMemcpyAsync(d.din[it & 1], d.hin[it & 1], kInBytes, cudaMemcpyHostToDevice, d.up); u += kInBytes; }
        if (mode & 2) { cudaMemcpyAsync(d.hout[it & 1], d.dout[it & 1], kOutBytes, cudaMemcpyDeviceToHost, d.down); dn += kOutBytes; }
        ++it;
        if ((it & 1) == 0) {  // two copies per direction in flight, then wait for them
            cudaStreamSynchronize(d.up);
            cudaStreamSynchronize(d.down);
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > seconds) break;
        }
    }
    cudaStreamSynchronize(d.up);
    cudaStreamSynchronize(d.down);
    *secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    *up = u; *down = dn;
}

// one participant (process or thread): warm-up + timed run of the three modes; results of mode m into sh slot k
static void participant(Shared* sh, int k, int n, int dev, double seconds, double res[3][2]) {
    Bufs b;
    if (!setup(dev, b)) sh->failed.store(1);
    int phase = 0;
    for (int mode = 1; mode <= 3; ++mode) {
        double u, d, s;
        meet(sh, phase++, n);
        if (!sh->failed.load()) copy_loop(b, mode, 0.1, &u, &d, &s);
        meet(sh, phase++, n);
        u = d = 0; s = 1;
        if (!sh->failed.load()) copy_loop(b, mode, seconds, &u, &d, &s);
        res[mode - 1][0] = u / s / 1e9;
        res[mode - 1][1] = d / s / 1e9;
    }
}

static bool run_set(const std::vector<int>& set, double seconds, bool threads, double out[3][2]) {
    const int n = (int)set.size();
    Shared* sh = (Shared*)mmap(nullptr, sizeof(Shared) + 8 * 6 * sizeof(double), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    new (sh) Shared();
    for (auto& a : sh->arrived) a.store(0);
    sh->failed.store(0);
    double (*res)[3][2] = (double (*)[3][2])(sh + 1);
    if (threads) {
        std::vector<std::thread> th;
        for (int k = 0; k < n; ++k) th.emplace_back([&, k] { participant(sh, k, n, set[k], seconds, res[k]); });
        for (auto& t : th) t.join();
    } else {
        std::vector<pid_t> pids;
        for (int k = 0; k < n; ++k) {
            const pid_t pid = fork();   // (no CUDA call has been made in this process: every child creates its own context)
            if (pid == 0) { participant(sh, k, n, set[k], seconds, res[k]); _exit(0); }
            pids.push_back(pid);
        }
        for (pid_t p : pids) { int st; waitpid(p, &st, 0); }
    }
    for (int m = 0; m < 3; ++m) { out[m][0] = out[m][1] = 0; for (int k = 0; k < n; ++k) { out[m][0] += res[k][m][0]; out[m][1] += res[k][m][1]; } }
    const bool ok = !sh->failed.load();
    munmap(sh, sizeof(Shared) + 8 * 6 * sizeof(double));
    return ok;
}

int main(int argc, char** argv) {
    const double seconds = argc > 1 ? atof(argv[1]) : 0.4;
    const char* only = argc > 2 && strcmp(argv[2], "all") ? argv[2] : nullptr;   // a comma separated device list: just that set (bench.py's live ceiling)
    const bool threads = argc > 3 && !strcmp(argv[3], "threads");
    int ndev = 8;
    if (const char* e = getenv("FEED_PROBE_NDEV")) ndev = atoi(e);   // (the device count must be known without a CUDA call in the parent)
    else {
        FILE* f = popen("nvidia-smi -L 2>/dev/null | wc -l", "r");
        if (f) { if (fscanf(f, "%d", &ndev) != 1) ndev = 1; pclose(f); }
    }
    std::vector<std::vector<int>> sets;
    if (only) {
        std::vector<int> s;
        for (const char* q = only; *q;) { s.push_back(atoi(q)); while (*q && *q != ',') ++q; if (*q) ++q; }
        sets.push_back(s);
    } else {
        sets.push_back({0});
        if (ndev >= 2) sets.push_back({0, 1});
        if (ndev >= 8) { sets.push_back({0, 4}); sets.push_back({0, 7}); }
        if (ndev >= 4) sets.push_back({0, 1, 2, 3});
        if (ndev >= 8) { sets.push_back({0, 2, 4, 6}); sets.push_back({4, 5, 6, 7}); sets.push_back({0, 1, 2, 3, 4, 5, 6, 7}); }
    }
    printf("{\"devices\": %d, \"feeders\": \"%s\", \"chunk_in_bytes\": %zu, \"chunk_out_bytes\": %zu, \"seconds_per_run\": %.2f, \"runs\": [\n", ndev,
           threads ? "one process, one thread per GPU" : "one process per GPU", kInBytes, kOutBytes, seconds);
    bool first = true;
    for (auto& s : sets) {
        std::string name;
        for (int id : s) name += (name.empty() ? "" : ",") + std::to_string(id);
        double r[3][2];
        const bool ok = run_set(s, seconds, threads, r);
        printf("%s {\"gpus\": [%s], \"n\": %zu, \"ok\": %s, \"h2d_only_GBps\": %.1f, \"d2h_only_GBps\": %.1f, \"duplex_h2d_GBps\": %.1f, \"duplex_d2h_GBps\": %.1f}", first ? "" : ",\n", name.c_str(),
               s.size(), ok ? "true" : "false", r[0][0], r[1][1], r[2][0], r[2][1]);
        first = false;
        fflush(stdout);
    }
    printf("\n]}\n");
    return 0;
}
