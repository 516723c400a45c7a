// tma_probe2.cu — bisects the TMA path: mode 0 mbarrier only, 1 bulk 1-D copy, 2 tensor map in param space, 3 tensor map in global memory
#include <cstdio>
#include <cstdlib>
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
    return (int)((EncodeFn)p)(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}
}
struct Maps { CUtensorMap m[2]; };
__global__ void k(const __grid_constant__ Maps maps, const CUtensorMap* gmap, int mode, int x, int y, int img, const unsigned char* src, unsigned char* out, int* status) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    mbar_fence_init();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (mode == 0) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
        } else if (mode == 1) {
            mbar_arrive_expect_tx(&bar, 992);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm)), "l"(src), "r"(992), "r"(smem_u32(&bar)) : "memory");
        } else if (mode == 2) {
            mbar_arrive_expect_tx(&bar, 992);
            tma_load_3d(sm, &maps.m[0], &bar, x, y, img);
        } else if (mode == 3) {
            mbar_arrive_expect_tx(&bar, 992);
            tma_load_3d(sm, gmap, &bar, x, y, img);
        } else {
            mbar_arrive_expect_tx(&bar, mode == 6 ? 512 : 992);
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(sm)), "l"(&maps.m[1]), "r"(smem_u32(&bar)), "r"(mode >= 5 ? 192 : x), "r"(y + img * 382) : "memory");
        }
    }
    int spins = 0; bool ok = false;
    for (; spins < (1 << 20); ++spins) if ((ok = mbar_try_wait(&bar, 0))) break;
    if (threadIdx.x == 0) { status[0] = ok ? 1 : 0; status[2] = spins; }
    if (ok) for (int i = threadIdx.x; i < 992; i += blockDim.x) out[i] = sm[i];
}
int encode2d(CUtensorMap* out, void* base, uint64_t pitch, uint64_t rows, uint32_t boxBytes, uint32_t boxRows, int dt) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return -1;
    const cuuint64_t gdim[2] = {pitch, rows}; const cuuint64_t gstride[1] = {1280};
    const cuuint32_t box[2] = {boxBytes, boxRows}; const cuuint32_t estride[2] = {1, 1};
    return (int)((EncodeFn)p)(out, (CUtensorMapDataType)dt, 2, base, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}
int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int pitch = 1280, rows = 382, n = 3; const size_t plane = (size_t)pitch * rows;
    std::vector<unsigned char> h(plane * n);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)((i * 2654435761u) >> 24);
    unsigned char *d, *o; int* st; CUtensorMap* gm; cudaMalloc(&d, h.size()); cudaMalloc(&o, 4096); cudaMalloc(&st, 16); cudaMalloc(&gm, 256);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    Maps maps;
    printf("encode rc %d\n", tma_encode_3d(&maps.m[0], d, pitch, rows, n, plane, 32, 31));
    maps.m[1] = maps.m[0];
    if (mode == 4) printf("encode2d rc %d\n", encode2d(&maps.m[1], d, pitch, (uint64_t)rows * n, 32, 31, CU_TENSOR_MAP_DATA_TYPE_UINT8));
    if (mode == 5) printf("encode2d u32 rc %d\n", encode2d(&maps.m[1], d, pitch / 4, (uint64_t)rows * n, 8, 31, CU_TENSOR_MAP_DATA_TYPE_UINT32));
    if (mode == 6) printf("encode2d f32 128B rc %d\n", encode2d(&maps.m[1], d, pitch / 4, (uint64_t)rows * n, 32, 4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32));
    cudaMemcpy(gm, &maps.m[0], 128, cudaMemcpyHostToDevice);
    const int x = 777, y = 100, img = 2;
    cudaMemset(st, 0, 16);
    k<<<1, 64, 4096>>>(maps, gm, mode, x, y, img, d + 1024, o, st);
    cudaError_t e = cudaDeviceSynchronize();
    int s[4] = {0, 0, 0, 0}; cudaMemcpy(s, st, 16, cudaMemcpyDeviceToHost);
    std::vector<unsigned char> got(992); cudaMemcpy(got.data(), o, got.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    if (mode >= 2) { for (int r = 0; r < 31; ++r) for (int c = 0; c < 32; ++c) bad += got[r * 32 + c] != h[(size_t)img * plane + (size_t)(y + r) * pitch + x + c]; }
    else if (mode == 1) for (int i = 0; i < 992; ++i) bad += got[i] != h[1024 + i];
    printf("mode %d: %s ok=%d spins=%d mismatches=%d\n", mode, cudaGetErrorString(e), s[0], s[2], bad);
    return 0;
}
