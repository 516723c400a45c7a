// tma.cuh — Tensor Memory Accelerator helpers (sm_100a): bulk tensor loads of image tiles into shared memory, completion
// through an mbarrier.  Used to stage FAST cell crops and keypoint patches with ONE instruction per tile instead of a
// per-thread gather (address arithmetic + LDG + STS per word).  Tensor maps describe one pyramid level of the whole batch
// as a 3-D byte tensor {row pitch, rows incl. margin, images}.  The global address of a box origin must be 16-byte aligned
// (an unaligned origin faults as "illegal instruction"; profiles/tools/tma_probe2.cu), rows and images are arbitrary.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace orbx {

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded spin: a box that never lands (bad tensor map) traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (int spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1 << 22)) __trap();
}
// box of the tensor map at coordinates (c0 = byte in row, c1 = row, c2 = image) -> dense [boxRows][boxBytes] in shared memory
__device__ __forceinline__ void tma_load_3d(void* smemDst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     smem_u32(smemDst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
#endif

// host: encode a 3-D uint8 tensor map {pitch, rows, images} with a {boxBytes, boxRows, 1} box; returns 0 on success
int tma_encode_3d(CUtensorMap* out, void* base, uint64_t pitch, uint64_t rows, uint64_t images, uint64_t imageStride, uint32_t boxBytes,
                  uint32_t boxRows);

}  // namespace orbx
