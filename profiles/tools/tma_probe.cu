// tma_probe.cu — isolates the TMA box load used by k_describe / k_fast_cells (debug tool)
#include <cstdio>
#include <cstdlib>
static int g_variant = 0;
#include <vector>
#include "../../tc2li_slam_b200/csrc/tma.cuh"
using namespace orbx;
namespace orbx {
int tma_encode_3d(CUtensorMap* out, void* base, uint64_t pitch, uint64_t rows, uint64_t images, uint64_t imageStride, uint32_t boxBytes, uint32_t boxRows) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return -1;
    const cuuint64_t gdim[3] = {pitch, rows, images}; const cuuint64_t gstride[2] = {pitch, imageStride};
    const cuuint32_t box[3] = {boxBytes, boxRows, 1}; const cuuint32_t estride[3] = {1, 1, 1};
    CUresult r = ((EncodeFn)p)(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, g_variant == 1 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return (int)r;
}
}
struct Maps { CUtensorMap m[2]; };
__global__ void k(const __grid_constant__ Maps maps, int which, int x, int y, int img, int boxB, int boxR, unsigned char* out, int* status) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); }
    mbar_fence_init();
    __syncthreads();
    if (threadIdx.x == 0) {
        status[1] = (int)(smem_u32(sm) & 127);
        mbar_arrive_expect_tx(&bar, boxB * boxR);
        tma_load_3d(sm, &maps.m[which], &bar, x, y, img);
    }
    int spins = 0; bool ok = false;
    for (; spins < (1 << 20); ++spins) if ((ok = mbar_try_wait(&bar, 0))) break;
    if (threadIdx.x == 0) { status[0] = ok ? 1 : 0; status[2] = spins; }
    if (ok) for (int i = threadIdx.x; i < boxB * boxR; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int pitch = 1280, rows = 382, n = 3; const size_t plane = (size_t)pitch * rows;
    std::vector<unsigned char> h(plane * n);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)((i * 2654435761u) >> 24);
    unsigned char *d, *o; int* st; cudaMalloc(&d, h.size()); cudaMalloc(&o, 4096); cudaMalloc(&st, 16);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    g_variant = variant;
    Maps maps;
    const int b0w = variant == 2 ? 64 : 32, b1w = variant == 2 ? 64 : 48;
    int r0 = tma_encode_3d(&maps.m[0], d, pitch, rows, n, plane, b0w, 31), r1 = tma_encode_3d(&maps.m[1], d, pitch, rows, n, plane, b1w, 37);
    { const unsigned* w = (const unsigned*)&maps.m[0]; for (int i = 0; i < 32; ++i) printf("%08x%c", w[i], i % 8 == 7 ? '\n' : ' '); }
    printf("encode rc %d %d\n", r0, r1);
    for (int which = 0; which < 2; ++which) {
        const int bb = which ? b1w : b0w, br = which ? 37 : 31, x = 777, y = 100, img = 2;
        cudaMemset(st, 0, 16);
        k<<<1, 64, 4096>>>(maps, which, x, y, img, bb, br, o, st);
        cudaError_t e = cudaDeviceSynchronize();
        int s[4]; cudaMemcpy(s, st, 16, cudaMemcpyDeviceToHost);
        std::vector<unsigned char> got(bb * br); cudaMemcpy(got.data(), o, got.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r = 0; r < br; ++r) for (int c = 0; c < bb; ++c) bad += got[r * bb + c] != h[(size_t)img * plane + (size_t)(y + r) * pitch + x + c];
        printf("which %d: %s ok=%d smemAlign=%d spins=%d mismatches=%d\n", which, cudaGetErrorString(e), s[0], s[1], s[2], bad);
    }
    return 0;
}
