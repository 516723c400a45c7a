// bow_kernel.cu — Frame::ComputeBoW on the device (widening).
//
// Replaces, for the descriptors the extractor left resident (or a host array of them),
//     mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4)                       (Frame.cc:768-775, KeyFrame::ComputeBoW)
// i.e. DBoW2's TemplatedVocabulary<FORB::TDescriptor, FORB>::transform(features, v, fv, levelsup)
// (slam_framework/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1139-1214), its per-feature tree descent (:1230-1276) with
// FORB::distance (FORB.cpp:81-100), BowVector::addWeight / addIfNotExist / normalize (BowVector.cpp:34-84), FeatureVector::addFeature
// (FeatureVector.cpp:29-43) and the scoring objects' mustNormalize (ScoringObject.h:74-89).  The vocabulary comes from the text
// format ORB-SLAM ships ORBvoc.txt in (loadFromTextFile, TemplatedVocabulary.h:1350-1433; System.cc loads it once at start-up).
//
//   k_bow_descend  : one thread per descriptor.  Level by level the thread reads the node's children (two 16-byte loads each; the
//                    whole k = 10, L = 6 tree is 35 MB and lives in L2), XOR + POPC over eight words, keeps the FIRST minimum
//                    (strict <, :1256-1260), remembers the node at level L - levelsup.  Output: leaf node and that node.
//   k_bow_assemble : one CTA per image.  (leaf, feature) keys are bitonic-sorted in shared memory; a run of equal leaves is one
//                    word (word ids ascend with leaf node ids: both are numbered in file order).  The value of a word is its weight
//                    added c times IN SEQUENCE (v += w per feature, as the reference's map accumulates it: c * w would round
//                    differently), or the weight once for IDF / BINARY.  The norm is the sequential sum over ascending word ids (the
//                    std::map iteration order), L1 or L2, every operation an explicit round-to-nearest double op, division and
//                    square root IEEE — the values are bit-identical to the CPU's.  A second sort on (node, feature) gives the
//                    FeatureVector as CSR over ascending node ids with ascending feature indices.
// Stopped words (weight 0) are left out of both vectors (:1170, :1198).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "orbx_internal.cuh"

namespace orbx {

namespace {

constexpr int BT = 256;   // threads per CTA of k_bow_assemble

// A node descriptor is one 32-byte sector: ONE 256-bit load (LDG.E.256, sm_100) instead of two 128-bit ones.  The lanes of a warp walk different nodes, so
// every load is its own L1 access, and that pipe — not L2 (the tree's home) or POPC — is what bounds the descent (ncu: 2 x 123 M sector accesses per 2 M
// descriptors with the 128-bit pair, 0.69 ms per 1024 images).
__device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const uint4* __restrict__ d) {
    uint32_t b[8];
    asm("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]) : "l"(d));
    return __popc(a0.x ^ b[0]) + __popc(a0.y ^ b[1]) + __popc(a0.z ^ b[2]) + __popc(a0.w ^ b[3]) + __popc(a1.x ^ b[4]) + __popc(a1.y ^ b[5]) +
           __popc(a1.z ^ b[6]) + __popc(a1.w ^ b[7]);
}

// in-place ascending bitonic sort of P = 2^m 64-bit keys in shared memory
__device__ void bitonic_sort(unsigned long long* key, int P) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (P >> 1); t += BT) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = key[lo], c = key[hi];
                if ((a > c) == up) { key[lo] = c; key[hi] = a; }
            }
        }
    }
    __syncthreads();
}

// exclusive prefix over the CTA of one int per thread; returns the thread's offset, *total = sum over the CTA
__device__ int block_exclusive(int v, int* warpSum, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += u;
    }
    __syncthreads();   // (warpSum may still be read from the previous use)
    if (lane == 31) warpSum[wid] = incl;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < BT / 32; ++w) { const int s = warpSum[w]; if (w < wid) base += s; tot += s; }
    *total = tot;
    return base + incl - v;
}

}  // namespace

__global__ void __launch_bounds__(128) k_bow_descend(const uint8_t* __restrict__ desc, const int* __restrict__ counts, int fixedCount, int K, BowVocDev v,
                                                    int nidLevel, uint32_t* __restrict__ leafOut, uint32_t* __restrict__ nidOut) {
    const int img = blockIdx.y, i = blockIdx.x * 128 + threadIdx.x;
    const int n = counts ? counts[img] : fixedCount;
    if (i >= n) return;
    const uint4* q = reinterpret_cast<const uint4*>(desc + ((size_t)img * K + i) * 32);
    const uint4 a0 = q[0], a1 = q[1];
    int node = 0, level = 0, nid = 0;
    int cs = __ldg(v.childStart), cn = __ldg(v.childStart + 1) - cs;
    do {
        ++level;
        int best = __ldg(v.childList + cs);
        int bestD = hamming256(a0, a1, v.desc + 2 * (size_t)best);
        for (int c = 1; c < cn; ++c) {
            const int id = __ldg(v.childList + cs + c);
            const int d = hamming256(a0, a1, v.desc + 2 * (size_t)id);
            if (d < bestD) { bestD = d; best = id; }
        }
        node = best;
        if (level == nidLevel) nid = node;
        cs = __ldg(v.childStart + node);
        cn = __ldg(v.childStart + node + 1) - cs;
    } while (cn > 0);
    leafOut[(size_t)img * K + i] = (uint32_t)node;
    nidOut[(size_t)img * K + i] = (uint32_t)nid;
}

__global__ void __launch_bounds__(BT) k_bow_assemble(const int* __restrict__ counts, int fixedCount, int K, int P, BowVocDev v, const uint32_t* __restrict__ leafIn,
                                                    const uint32_t* __restrict__ nidIn, BowOut o) {
    extern __shared__ unsigned long long key[];   // [P]
    __shared__ int warpSum[BT / 32];
    __shared__ double chunk[BT];
    __shared__ double normSh;
    const int img = blockIdx.x, tid = threadIdx.x;
    const int n = min(counts ? counts[img] : fixedCount, K);
    const uint32_t* leaf = leafIn + (size_t)img * K;
    const uint32_t* nidp = nidIn + (size_t)img * K;
    uint32_t* words = o.words + (size_t)img * K;
    double* values = o.values + (size_t)img * K;
    uint32_t* nodes = o.nodes + (size_t)img * K;
    int* nodeStart = o.nodeStart + (size_t)img * (K + 1);
    uint32_t* feats = o.features + (size_t)img * K;
    const bool tf = v.weighting == 0 || v.weighting == 1;   // TF_IDF, TF accumulate; IDF, BINARY count a word once (:1156, :1183)
    const bool must = v.scoring != 5;                        // every scoring object but DotProductScoring normalises (ScoringObject.h:74-89)
    const bool l2 = v.scoring == 1;

    // ---- BowVector ----
    for (int i = tid; i < P; i += BT) {
        unsigned long long k = ~0ull;
        if (i < n) {
            const uint32_t lf = leaf[i];
            if (__ldg(v.weight + lf) > 0.0) k = ((unsigned long long)lf << 32) | (unsigned)i;
        }
        key[i] = k;
    }
    bitonic_sort(key, P);
    int nWords = 0;
    for (int base = 0; base < P; base += BT) {   // P is a multiple of BT or smaller than it: every thread runs the same trips
        const int i = base + tid;
        bool head = false;
        uint32_t lf = 0;
        if (i < P) {
            const unsigned long long k = key[i];
            if (k != ~0ull) {
                lf = (uint32_t)(k >> 32);
                head = i == 0 || (uint32_t)(key[i - 1] >> 32) != lf;
            }
        }
        int tot;
        const int u = nWords + block_exclusive(head ? 1 : 0, warpSum, &tot);
        if (head) {
            int c = 1;
            while (i + c < P && (uint32_t)(key[i + c] >> 32) == lf && key[i + c] != ~0ull) ++c;
            const double w = __ldg(v.weight + lf);
            double val = w;
            if (tf) for (int r = 1; r < c; ++r) val = __dadd_rn(val, w);   // BowVector::addWeight: vit->second += v, feature by feature
            words[u] = (uint32_t)__ldg(v.wordId + lf);
            values[u] = val;
        }
        nWords += tot;
    }
    __syncthreads();
    if (nWords > 0 && (must || tf)) {
        if (must) {
            // BowVector::normalize (BowVector.cpp:61-84): sequential sum in ascending word order
            double norm = 0.0;
            for (int base = 0; base < nWords; base += BT) {
                __syncthreads();
                if (base + tid < nWords) {
                    const double x = values[base + tid];
                    chunk[tid] = l2 ? __dmul_rn(x, x) : fabs(x);
                }
                __syncthreads();
                if (tid == 0) {
                    const int m = min(BT, nWords - base);
                    for (int j = 0; j < m; ++j) norm = __dadd_rn(norm, chunk[j]);
                }
            }
            if (tid == 0) normSh = l2 ? __dsqrt_rn(norm) : norm;
        } else if (tid == 0) {
            normSh = (double)nWords;   // TF / TF_IDF without normalisation: vit->second /= v.size() (:1175-1181)
        }
        __syncthreads();
        const double norm = normSh;
        if (!must || norm > 0.0)
            for (int j = tid; j < nWords; j += BT) values[j] = __ddiv_rn(values[j], norm);
    }
    if (tid == 0) o.nWords[img] = nWords;
    __syncthreads();

    // ---- FeatureVector ----
    for (int i = tid; i < P; i += BT) {
        unsigned long long k = ~0ull;
        if (i < n && __ldg(v.weight + leaf[i]) > 0.0) k = ((unsigned long long)nidp[i] << 32) | (unsigned)i;
        key[i] = k;
    }
    bitonic_sort(key, P);
    int nNodes = 0, nFeat = 0;
    for (int base = 0; base < P; base += BT) {
        const int i = base + tid;
        bool head = false, valid = false;
        unsigned long long k = ~0ull;
        if (i < P) {
            k = key[i];
            valid = k != ~0ull;
            head = valid && (i == 0 || (uint32_t)(key[i - 1] >> 32) != (uint32_t)(k >> 32));
        }
        int tot;
        const int u = nNodes + block_exclusive(head ? 1 : 0, warpSum, &tot);
        if (head) { nodes[u] = (uint32_t)(k >> 32); nodeStart[u] = i; }   // valid keys sort to the front: position = offset
        if (valid) { feats[i] = (uint32_t)k; }
        nNodes += tot;
        int totV;
        block_exclusive(valid ? 1 : 0, warpSum, &totV);
        nFeat += totV;
    }
    if (tid == 0) { nodeStart[nNodes] = nFeat; o.nNodes[img] = nNodes; }
}

void launch_bow(const BowVocDev& v, const uint8_t* desc, const int* counts, int fixedCount, int K, int nImages, int levelsup, uint32_t* leaf, uint32_t* nid,
                const BowOut& out, cudaStream_t s) {
    if (nImages <= 0 || K <= 0) return;
    dim3 grid((unsigned)((K + 127) / 128), (unsigned)nImages);
    k_bow_descend<<<grid, 128, 0, s>>>(desc, counts, fixedCount, K, v, v.L - levelsup, leaf, nid);
    const int P = bow_sort_size(K);
    k_bow_assemble<<<nImages, BT, (size_t)P * sizeof(unsigned long long), s>>>(counts, fixedCount, K, P, v, leaf, nid, out);
}

int bow_sort_size(int K) {
    int P = 32;
    while (P < K) P <<= 1;
    return P;
}

cudaError_t bow_configure(int K) { return raise_dynamic_smem(k_bow_assemble, (size_t)bow_sort_size(K) * sizeof(unsigned long long)); }

}  // namespace orbx

// ------------------------------------------------------------------------------------------------------------------
// host side: the vocabulary object
// ------------------------------------------------------------------------------------------------------------------
using namespace orbx;

struct orbx_vocabulary {
    int device = 0;
    int nNodes = 0, nWords = 0;
    BowVocDev dev = {};
    // scratch of orbx_vocabulary_transform (host descriptors in, host vectors out)
    cudaStream_t stream = nullptr;
    void* scratch = nullptr;
    size_t scratchBytes = 0;
    int scratchK = 0;
};

namespace {
struct DevGuard {
    int prev = -1;
    explicit DevGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define CUV(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return orbx::set_cuda_error(e__, #x); } while (0)
}  // namespace

extern "C" {

void orbx_vocabulary_destroy(orbx_vocabulary* v) {
    if (!v) return;
    DevGuard dg(v->device);
    if (v->stream) { cudaStreamSynchronize(v->stream); cudaStreamDestroy(v->stream); }
    cudaFree((void*)v->dev.desc); cudaFree((void*)v->dev.childStart); cudaFree((void*)v->dev.childList); cudaFree((void*)v->dev.weight); cudaFree((void*)v->dev.wordId);
    if (v->scratch) cudaFree(v->scratch);
    delete v;
}

// The tree as flat arrays, node 0 = the root: parent[n] (parent[0] ignored), is_word[n] (a leaf that is a word), descriptors[n][32], weights[n].
// Children keep the order of their node ids (the order loadFromTextFile pushes them in); word ids are handed out to the words in
// ascending node id (TemplatedVocabulary.h:1414-1421).
int orbx_vocabulary_create(int device, int k, int L, int scoring, int weighting, int n_nodes, const int* parent, const uint8_t* is_word, const uint8_t* descriptors,
                           const double* weights, orbx_vocabulary** out) {
    if (!out) return set_error(ORBX_ERR_ARG, "null out pointer");
    *out = nullptr;
    // the bounds loadFromTextFile accepts (TemplatedVocabulary.h:1364-1369)
    if (k < 0 || k > 20 || L < 1 || L > 10 || scoring < 0 || scoring > 5 || weighting < 0 || weighting > 3) return set_error(ORBX_ERR_ARG, "vocabulary header out of range");
    if (n_nodes < 2 || !parent || !is_word || !descriptors || !weights) return set_error(ORBX_ERR_ARG, "vocabulary needs a root and at least one more node");
    std::vector<int> childStart((size_t)n_nodes + 1, 0), childList((size_t)n_nodes - 1), wordId((size_t)n_nodes, 0);
    for (int n = 1; n < n_nodes; ++n) {
        if (parent[n] < 0 || parent[n] >= n) return set_error(ORBX_ERR_ARG, "a node's parent must have a smaller id (file order)");
        ++childStart[(size_t)parent[n] + 1];
    }
    for (int n = 0; n < n_nodes; ++n) childStart[(size_t)n + 1] += childStart[(size_t)n];
    {
        std::vector<int> cur(childStart.begin(), childStart.end() - 1);
        for (int n = 1; n < n_nodes; ++n) childList[(size_t)cur[(size_t)parent[n]]++] = n;
    }
    int nWords = 0;
    for (int n = 1; n < n_nodes; ++n) {
        const bool leaf = childStart[(size_t)n + 1] == childStart[(size_t)n];
        if (is_word[n]) {
            if (!leaf) return set_error(ORBX_ERR_ARG, "a word node has children");
            wordId[(size_t)n] = nWords++;
        }   // (a childless node that is not marked as a word keeps DBoW2's default word id 0)
    }
    if (childStart[1] == 0) return set_error(ORBX_ERR_ARG, "the root has no children");
    orbx_vocabulary* v = new orbx_vocabulary();
    v->device = device; v->nNodes = n_nodes; v->nWords = nWords;
    DevGuard dg(device);
    auto up = [&](const void* src, size_t bytes, const void** dst) -> cudaError_t {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(bytes, 256));
        if (e != cudaSuccess) return e;
        *dst = p;
        return cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice);
    };
    cudaError_t e;
    if ((e = up(descriptors, (size_t)n_nodes * 32, (const void**)&v->dev.desc)) != cudaSuccess ||
        (e = up(childStart.data(), childStart.size() * sizeof(int), (const void**)&v->dev.childStart)) != cudaSuccess ||
        (e = up(childList.data(), childList.size() * sizeof(int), (const void**)&v->dev.childList)) != cudaSuccess ||
        (e = up(weights, (size_t)n_nodes * sizeof(double), (const void**)&v->dev.weight)) != cudaSuccess ||
        (e = up(wordId.data(), wordId.size() * sizeof(int), (const void**)&v->dev.wordId)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        orbx_vocabulary_destroy(v);
        return set_cuda_error(e, "vocabulary upload");
    }
    v->dev.k = k; v->dev.L = L; v->dev.scoring = scoring; v->dev.weighting = weighting;
    *out = v;
    return ORBX_OK;
}

// ORBVocabulary::loadFromTextFile (TemplatedVocabulary.h:1350-1433): "k L scoring weighting", then one line per node
// "parent is_leaf b0 .. b31 weight" in creation order.  Blank lines are skipped (the reference's eof loop would turn a trailing
// blank line into a node with an uninitialised descriptor).
int orbx_vocabulary_load_text(int device, const char* path, orbx_vocabulary** out) {
    if (!out) return set_error(ORBX_ERR_ARG, "null out pointer");
    *out = nullptr;
    if (!path) return set_error(ORBX_ERR_ARG, "null path");
    FILE* f = std::fopen(path, "rb");
    if (!f) return set_error(ORBX_ERR_ARG, std::string("cannot open vocabulary file ") + path);
    std::fseek(f, 0, SEEK_END);
    const long size = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::string text((size_t)std::max(size, 0L), '\0');
    const size_t got = size > 0 ? std::fread(&text[0], 1, (size_t)size, f) : 0;
    std::fclose(f);
    text.resize(got);
    const char* p = text.c_str();
    char* e = nullptr;
    long hdr[4];
    for (int i = 0; i < 4; ++i) {
        hdr[i] = std::strtol(p, &e, 10);
        if (e == p) return set_error(ORBX_ERR_ARG, "vocabulary file: bad header");
        p = e;
    }
    while (*p && *p != '\n') ++p;
    std::vector<int> parent(1, 0);
    std::vector<uint8_t> isWord(1, 0), desc(32, 0);
    std::vector<double> weight(1, 0.0);
    while (*p) {
        const char* eol = std::strchr(p, '\n');
        const char* next = eol ? eol + 1 : p + std::strlen(p);
        const long pid = std::strtol(p, &e, 10);
        if (e == p || e > next) { p = next; continue; }   // blank line
        p = e;
        if (pid < 0 || pid >= (long)parent.size()) return set_error(ORBX_ERR_ARG, "vocabulary file: parent id out of range");
        const long leaf = std::strtol(p, &e, 10); p = e;
        parent.push_back((int)pid);
        isWord.push_back(leaf > 0);
        for (int i = 0; i < 32; ++i) { const long b = std::strtol(p, &e, 10); p = e; desc.push_back((uint8_t)b); }
        weight.push_back(std::strtod(p, &e));
        p = next;
    }
    return orbx_vocabulary_create(device, (int)hdr[0], (int)hdr[1], (int)hdr[2], (int)hdr[3], (int)parent.size(), parent.data(), isWord.data(), desc.data(),
                                  weight.data(), out);
}

// {k, L, scoring, weighting, nodes (with the root), words}
int orbx_vocabulary_info(const orbx_vocabulary* v, int out[6]) {
    if (!v || !out) return set_error(ORBX_ERR_ARG, "null argument");
    out[0] = v->dev.k; out[1] = v->dev.L; out[2] = v->dev.scoring; out[3] = v->dev.weighting; out[4] = v->nNodes; out[5] = v->nWords;
    return ORBX_OK;
}

// transform() of n host descriptors (KeyFrame::ComputeBoW, or any caller that holds descriptors on the host); synchronous
int orbx_vocabulary_transform(orbx_vocabulary* v, const uint8_t* descriptors, int n, int levelsup, uint32_t* words, double* values, int* n_words, uint32_t* nodes,
                              int* node_start, uint32_t* features, int* n_nodes) {
    if (!v || n < 0 || (n && !descriptors) || !n_words || !n_nodes || !node_start) return set_error(ORBX_ERR_ARG, "bad transform arguments");
    if (n && (!words || !values || !nodes || !features)) return set_error(ORBX_ERR_ARG, "bad transform arguments");
    *n_words = 0; *n_nodes = 0; node_start[0] = 0;
    if (n == 0) return ORBX_OK;
    if (n > ORBX_BOW_MAX_FEATURES) return set_error(ORBX_ERR_TOO_LARGE, "more than ORBX_BOW_MAX_FEATURES descriptors in one transform");
    DevGuard dg(v->device);
    const int K = n;
    // scratch: desc [K][32] | values [K] f64 | leaf, nid, words, nodes, features [K] u32 | nodeStart [K + 1] | counts [2]
    const size_t need = (size_t)K * 32 + (size_t)K * 4 * 5 + ((size_t)K + 1 + 2) * 4 + 8 + (size_t)K * 8;
    if (need > v->scratchBytes) {
        if (v->scratch) cudaFree(v->scratch);
        v->scratch = nullptr; v->scratchBytes = 0;
        CUV(cudaMalloc(&v->scratch, need));
        v->scratchBytes = need;
    }
    if (K > v->scratchK) { CUV(bow_configure(K)); v->scratchK = K; }
    uint8_t* base = (uint8_t*)v->scratch;
    uint8_t* dDesc = base; base += (size_t)K * 32;          // (16-byte aligned rows: the descent reads them as two uint4)
    double* dValues = (double*)base; base += (size_t)K * 8;
    uint32_t* dLeaf = (uint32_t*)base; base += (size_t)K * 4;
    uint32_t* dNid = (uint32_t*)base; base += (size_t)K * 4;
    BowOut o;
    o.values = dValues;
    o.words = (uint32_t*)base; base += (size_t)K * 4;
    o.nodes = (uint32_t*)base; base += (size_t)K * 4;
    o.features = (uint32_t*)base; base += (size_t)K * 4;
    o.nodeStart = (int*)base; base += ((size_t)K + 1) * 4;
    o.nWords = (int*)base; o.nNodes = o.nWords + 1;
    CUV(cudaMemcpyAsync(dDesc, descriptors, (size_t)n * 32, cudaMemcpyHostToDevice, v->stream));
    launch_bow(v->dev, dDesc, nullptr, n, K, 1, levelsup, dLeaf, dNid, o, v->stream);
    CUV(cudaGetLastError());
    int cnt[2];
    CUV(cudaMemcpyAsync(cnt, o.nWords, sizeof(cnt), cudaMemcpyDeviceToHost, v->stream));
    CUV(cudaStreamSynchronize(v->stream));
    *n_words = cnt[0]; *n_nodes = cnt[1];
    if (cnt[0]) {
        CUV(cudaMemcpyAsync(words, o.words, (size_t)cnt[0] * 4, cudaMemcpyDeviceToHost, v->stream));
        CUV(cudaMemcpyAsync(values, o.values, (size_t)cnt[0] * 8, cudaMemcpyDeviceToHost, v->stream));
    }
    CUV(cudaMemcpyAsync(node_start, o.nodeStart, ((size_t)cnt[1] + 1) * 4, cudaMemcpyDeviceToHost, v->stream));
    if (cnt[1]) CUV(cudaMemcpyAsync(nodes, o.nodes, (size_t)cnt[1] * 4, cudaMemcpyDeviceToHost, v->stream));
    CUV(cudaStreamSynchronize(v->stream));
    const int nf = node_start[cnt[1]];
    if (nf) {
        CUV(cudaMemcpyAsync(features, o.features, (size_t)nf * 4, cudaMemcpyDeviceToHost, v->stream));
        CUV(cudaStreamSynchronize(v->stream));
    }
    return ORBX_OK;
}

}  // extern "C"

namespace orbx {
const BowVocDev& vocabulary_device(const orbx_vocabulary* v) { return v->dev; }
int vocabulary_device_ordinal(const orbx_vocabulary* v) { return v->device; }
}  // namespace orbx
