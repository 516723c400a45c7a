// describe_kernel.cu — output assembly, IC_Angle orientation and 256-bit rBRIEF descriptors.
//
//   k_assemble : ORBextractor::operator() bookkeeping (ORBextractor.cc:1093-1138): level-major keypoint order, coordinates
//                scaled by mvScaleFactor[level] in float32, the vLappingArea partition (outside the window: filled from
//                the front in encounter order; inside: from the back), octave/size/response fields (ORBextractor.cc:854-864).
//   k_describe : one warp per keypoint.  IC_Angle (ORBextractor.cc:50-77) on the UNBLURRED level: lane u-15 owns patch
//                column u, the two moments are reduced with warp shuffles (exact in int32), angle = cv::fastAtan2.
//                computeOrbDescriptor (ORBextractor.cc:81-120) on the BLURRED level: lane i builds descriptor byte i from 16
//                rotated samples; pattern staged in shared memory; cosf/sinf as glibc evaluates them (float_exact.cuh).
// Both stages are gathers served by L2 (749 + 512 bytes per keypoint); no FMA contraction anywhere.
#include <cstdlib>
#include <type_traits>

#include "float_exact.cuh"
#include "orbx_internal.cuh"

namespace orbx {

__constant__ signed char c_pattern[1024];  // bit_pattern_31_ (ORBextractor.cc:123-381) as int8 x0,y0,x1,y1 per pair
__constant__ int c_umax[16];               // circular patch half-widths (ORBextractor.cc:427-442)

cudaError_t upload_describe_constants(const signed char pattern[1024], const int umax[16]) {
    cudaError_t e = cudaMemcpyToSymbol(c_pattern, pattern, 1024);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_umax, umax, 16 * sizeof(int));
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_assemble(const __grid_constant__ Geom g, const Buffers b) {
    __shared__ int warpTot[8];
    __shared__ int carry;  // keypoints outside the lapping window so far
    const int img = b.imgBase + blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int* selCount = b.selCount + img * ORBX_MAX_LEVELS;
    int total = 0;
    for (int l = 0; l < g.nlevels; ++l) total += selCount[l];
    const float lap0 = (float)b.lapping[2 * img], lap1 = (float)b.lapping[2 * img + 1];
    orbx_keypoint* kps = b.kps + (size_t)img * g.K;
    KpWork* work = b.work + (size_t)img * g.K;
    if (tid == 0) carry = 0;
    __syncthreads();
    int seen = 0;  // keypoints of the previous levels
    for (int l = 0; l < g.nlevels; ++l) {
        const LevelGeom& L = g.lv[l];
        const int nl = selCount[l];
        const uint32_t* sel = b.sel + (size_t)img * g.K + L.selOff;
        for (int base = 0; base < nl; base += 256) {
            const int k = base + tid;
            const bool valid = k < nl;
            int X = 0, Y = 0, R = 0;
            float sx = 0.f, sy = 0.f;
            bool inLap = false;
            if (valid) {
                const uint32_t pc = sel[k];
                X = cand_x(pc) + kBorder; Y = cand_y(pc) + kBorder; R = cand_r(pc);
                sx = (float)X; sy = (float)Y;
                if (l != 0) { sx = __fmul_rn(sx, L.sf); sy = __fmul_rn(sy, L.sf); }  // keypoint->pt *= scale (ORBextractor.cc:1122-1124)
                inLap = sx >= lap0 && sx <= lap1;
            }
            const unsigned monoMask = __ballot_sync(0xFFFFFFFFu, valid && !inLap);
            if (lane == 0) warpTot[wid] = __popc(monoMask);
            __syncthreads();
            int before = carry;
            for (int w = 0; w < wid; ++w) before += warpTot[w];
            const int monoBefore = before + __popc(monoMask & ((1u << lane) - 1));  // mono keypoints preceding this one
            if (valid) {
                const int ord = seen + k;  // encounter index
                const int dst = inLap ? (total - 1 - (ord - monoBefore)) : monoBefore;
                orbx_keypoint kp;
                kp.x = sx; kp.y = sy; kp.size = L.kpSize; kp.angle = -1.f; kp.response = (float)R; kp.octave = l; kp.class_id = -1;
                kps[dst] = kp;
                KpWork wk;
                wk.x = (short)X; wk.y = (short)Y; wk.level = (short)l; wk.dst = (short)dst;
                work[ord] = wk;
            }
            __syncthreads();
            if (tid == 0) {
                int s = 0;
                for (int w = 0; w < 8; ++w) s += warpTot[w];
                carry += s;
            }
            __syncthreads();
        }
        seen += nl;
    }
    if (tid == 0) {
        b.outCount[img] = total;
        b.outMono[img] = carry;
    }
}

void launch_assemble(const Geom& g, const Buffers& b, int nImages, cudaStream_t s) { k_assemble<<<nImages, 256, 0, s>>>(g, b); }

// ---------------------------------------------------------------------------------------------------------------
// k_describe: a CTA of DW = 2 warps handles 64 consecutive keypoints, KPW = 32 per warp, in three phases (measured, ms per 8192 KITTI
// images: DW x KPW = 8 x 8: 9.50, 4 x 16: 8.89, 8 x 32: 8.86, 2 x 32: 8.77 — long per-warp runs amortise the per-lane weight / pattern
// setup and phase 2 then keeps every thread of the CTA busy).
//   1  moments: per keypoint ONE lane issues a TMA box load of the 31x31 patch of the unblurred level (box rows of 48 bytes:
//      box origins must be 16-byte aligned in global memory, so the box starts at the aligned byte at or before column
//      -15) into the warp's double-buffered shared memory; the next box is in flight while the current one is reduced.
//      Lane j issues for the warp's keypoint j with operands it computed once (SIMD over the eight keypoints).
//      Lane v+15 owns patch row v: 9 words are funnel-shifted into 8 patch-aligned words, sum u*I and sum I are 2 x 8
//      IDP.4A against per-lane constant weight words (circular mask folded in), then two warp-shuffle reductions.
//   2  64 threads turn the 64 moment pairs into angle (cv::fastAtan2), cosf, sinf: the scalar float/double chain runs once
//      per keypoint instead of once per lane.  Meanwhile the first two 37x37 windows of the blurred level are in flight.
//   3  rBRIEF: lane i builds descriptor byte i from 16 rotated samples (|offset| <= 18: pattern radius 18.38) of the staged
//      window; its eight pattern pairs {x0,y0,x1,y1} stay in registers; cvRound as add-magic (RNE, exact below 2^22).  The
//      window rows are 80 bytes apart (TMA box width): with 64 the random byte gathers of a warp fall into 20 of the 32 banks
//      (5.1 wavefronts per load, ncu), with 80 the row offsets cycle through all of them.
// Timing experiment (boxes cut to 2 rows, wrong results): 9.17 instead of 9.55 ms, i.e. the kernel is not bound by the box loads
// or by L2 / HBM but by instruction issue and dependency latency (~470 warp instructions per keypoint).
#ifndef DESC_DW
#define DESC_DW 2
#endif
#ifndef DESC_KPW
#define DESC_KPW 32
#endif
constexpr int DW = DESC_DW;      // warps per CTA
constexpr int KPW = DESC_KPW;    // consecutive keypoints per warp
constexpr int PA_ROW = 48, PB_ROW = kDescWindowRow;
constexpr int PA_BYTES = PA_ROW * 31, PB_BYTES = PB_ROW * 37;
constexpr int ST_BYTES = (PB_ROW * 37 + 127) & ~127;   // one stage holds either box, 128-byte aligned
constexpr int SA_BYTES = 1536;   // stage stride of the moment phase (one 48 x 31 patch)
constexpr int NA = 2;            // patch stages = the lane parity of the issuing lane (three or four stages in the same memory measured slower: 10.3 vs 9.6 ms)
static_assert(NA * SA_BYTES <= 2 * ST_BYTES && PA_BYTES <= SA_BYTES, "moment stages");

// barrier waits on a 32-bit shared-window address (kept in a register; through generic pointers ptxas re-derives the window base, S2UR + ULEA,
// at every use: 7 % of the kernel's stall samples in the r02h profile)
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    for (int spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (spin > (1 << 22)) __trap();
    }
}

// the same for the box issue and the small per-keypoint records: everything in the loops addresses shared memory by register
__device__ __forceinline__ void tma_issue_a(uint32_t dst, const CUtensorMap* map, uint32_t bar, uint32_t bytes, int c0, int c1, int c2) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(map), "r"(bar),
                 "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ KpWork lds_work(uint32_t a) {
    uint32_t lo, hi;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(a));
    KpWork w;
    w.x = (short)(lo & 0xFFFFu); w.y = (short)(lo >> 16); w.level = (short)(hi & 0xFFFFu); w.dst = (short)(hi >> 16);
    return w;
}

__device__ __forceinline__ int dp4a_us(uint32_t pix, uint32_t wgt, int acc) {  // unsigned pixels x signed weights
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(pix), "r"(wgt), "r"(acc));
    return d;
}

__global__ void __launch_bounds__(32 * DW) k_describe(const __grid_constant__ Geom g, const Buffers b, const __grid_constant__ DescMaps maps) {
    extern __shared__ __align__(128) unsigned char dsm[];
    __shared__ __align__(16) float4 pat4[256];
    __shared__ __align__(8) uint64_t bars[DW][2];
    __shared__ int mom[DW * KPW][2];
    __shared__ float2 rot[DW * KPW];  // (sin, cos)
    __shared__ KpWork wks[DW * KPW];  // the CTA's work items (one coalesced read instead of a global load per use)
    const int img = b.imgBase + blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = b.outCount[img];
    const int kBase = blockIdx.x * DW * KPW;
    if (kBase >= n) return;
    if (tid < DW * KPW && kBase + tid < n) wks[tid] = b.work[(size_t)img * g.K + kBase + tid];
    {   // pair p = 8 * byte + bit is read by lane `byte` in iteration `bit`: store it at [bit][byte] so that a warp reads 32
        // consecutive float4 (a [byte][bit] layout makes every lane hit the same banks: 8-way conflicts)
        for (int t = tid; t < 256; t += 32 * DW) {
            const signed char* q = c_pattern + 4 * t;
            pat4[(t & 7) * 32 + (t >> 3)] = make_float4((float)q[0], (float)q[1], (float)q[2], (float)q[3]);
        }
    }
    if (lane == 0) { mbar_init(&bars[wid][0], 1); mbar_init(&bars[wid][1], 1); }
    mbar_fence_init();
    __syncthreads();
    const int k0 = kBase + wid * KPW;
    const int cnt = max(0, min(KPW, n - k0));
    unsigned char* stage = dsm + (size_t)wid * 2 * ST_BYTES;
    const int wk0 = wid * KPW;   // this warp's work items: wks[wk0 + i] (indexed, not through a pointer: the address stays a shared-window constant)
    uint32_t phbits = 0;  // bit s = parity of stage barrier s (uniform per warp)
    // shared-window addresses of the warp's stages and barriers, through a shuffle so that ptxas keeps them in registers
    const uint32_t stageS = __shfl_sync(0xFFFFFFFFu, smem_u32(stage), 0);
    const uint32_t barS = __shfl_sync(0xFFFFFFFFu, smem_u32(&bars[wid][0]), 0);
    const uint32_t wksS = __shfl_sync(0xFFFFFFFFu, smem_u32(&wks[wk0]), 0);
    const uint32_t rotS = __shfl_sync(0xFFFFFFFFu, smem_u32(&rot[wk0]), 0);
    const uint32_t momS = __shfl_sync(0xFFFFFFFFu, smem_u32(&mom[wk0][0]), 0);

    // Lane j owns the box loads of the warp's keypoint j: the box coordinates and the tensor-map offset are computed once for
    // all eight keypoints (SIMD) and stay in the lane's registers, the stage is the lane's parity, and an issue is a one-lane
    // branch around expect_tx + UTMALDG (the operands go to uniform registers either way) instead of ~20 scalar setup instructions.
    const KpWork myWk = wks[wid * KPW + (lane & (KPW - 1))];
    const int myLvOff = (int)myWk.level * (int)sizeof(CUtensorMap);
    const int myAx = (kMarginX + myWk.x - kHalfPatch) & ~15, myBx = (kMarginX + myWk.x - 18) & ~15, myAy = kMarginY + myWk.y - kHalfPatch;
    const uint32_t myBar = barS + 8u * (lane & 1);
    const uint32_t myStageA = stageS + (lane & 1) * SA_BYTES, myStageB = stageS + (lane & 1) * ST_BYTES;
    auto issueB = [&](int i) {  // executed by lane i only
        tma_issue_a(myStageB, reinterpret_cast<const CUtensorMap*>(reinterpret_cast<const char*>(&maps.bl[0]) + myLvOff), myBar, PB_BYTES, myBx, myAy - 3, img);
    };
    auto wait_stage = [&](int st) {
        mbar_wait_a(barS + 8u * (uint32_t)st, (phbits >> st) & 1u);
        phbits ^= 1u << st;
    };

    // ---- phase 1: moments ----
    {
        // per-lane weights of patch row v = lane - 15: bytes u = 4c + j - 15, kept iff |u| <= umax[|v|] (ORBextractor.cc:57-74)
        uint32_t wu[8], w1[8];
        const int v = lane - kHalfPatch, av = v < 0 ? -v : v;
        const int lim = lane < 31 ? c_umax[av] : -1;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            uint32_t a = 0, o = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int u = 4 * c + j - kHalfPatch;
                const bool in = (u < 0 ? -u : u) <= lim && u <= kHalfPatch;
                a |= (in ? (uint32_t)(u & 0xFF) : 0u) << (8 * j);
                o |= (in ? 1u : 0u) << (8 * j);
            }
            wu[c] = a; w1[c] = o;
        }
        auto issueMyA = [&]() {
            tma_issue_a(myStageA, reinterpret_cast<const CUtensorMap*>(reinterpret_cast<const char*>(&maps.ic[0]) + myLvOff), myBar, PA_BYTES, myAx, myAy, img);
        };
        if (lane == 0 && cnt > 0) issueMyA();
        const uint32_t rowOff = (uint32_t)(lane < 31 ? lane : 0) * PA_ROW;
        int st = 0;  // stage of keypoint i
        for (int i = 0; i < cnt; ++i) {
            if (lane == i + 1 && i + 1 < cnt) issueMyA();
            const int sa = (kMarginX + lds_work(wksS + 8u * (uint32_t)i).x - kHalfPatch) & 15;  // patch column -15 sits at byte sa of the box row
            wait_stage(st);
            // the whole 48-byte box row as three 128-bit loads (conflict-free at this stride), then a warp-uniform word offset
            uint32_t rw[12];
            {
                const uint32_t a = stageS + (uint32_t)st * SA_BYTES + rowOff;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rw[0]), "=r"(rw[1]), "=r"(rw[2]), "=r"(rw[3]) : "r"(a));
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+16];" : "=r"(rw[4]), "=r"(rw[5]), "=r"(rw[6]), "=r"(rw[7]) : "r"(a));
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+32];" : "=r"(rw[8]), "=r"(rw[9]), "=r"(rw[10]), "=r"(rw[11]) : "r"(a));
            }
            const int bs = (sa & 3) * 8;
            int m10 = 0, s1 = 0;
            auto reduce_row = [&](auto W0) {
                constexpr int w0 = decltype(W0)::value;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t wd = __funnelshift_r(rw[w0 + c], rw[w0 + c + 1], bs);
                    m10 = dp4a_us(wd, wu[c], m10);
                    s1 = (int)__dp4a(wd, w1[c], (uint32_t)s1);
                }
            };
            switch (sa >> 2) {
                case 0: reduce_row(std::integral_constant<int, 0>()); break;
                case 1: reduce_row(std::integral_constant<int, 1>()); break;
                case 2: reduce_row(std::integral_constant<int, 2>()); break;
                default: reduce_row(std::integral_constant<int, 3>()); break;
            }
            int m01 = v * s1;  // lane 31 has all-zero weights
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                m10 += __shfl_xor_sync(0xFFFFFFFFu, m10, o);
                m01 += __shfl_xor_sync(0xFFFFFFFFu, m01, o);
            }
            if (lane == 0) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(momS + 8u * (uint32_t)i), "r"(m01), "r"(m10) : "memory");
            __syncwarp();  // every lane is done with this stage before lane 0 re-arms it
            st = st + 1 == NA ? 0 : st + 1;
        }
        // both stages are free: put the first two blurred windows in flight while phase 2 runs
        if (lane < 2 && lane < cnt) issueB(lane);
    }
    __syncthreads();

    // ---- phase 2: angle, sin, cos — one thread per keypoint ----
    if (tid < DW * KPW && kBase + tid < n) {
        const float angle = fast_atan2_deg((float)mom[tid][0], (float)mom[tid][1]);
        const float factorPI = (float)(3.14159265358979323846 / 180.0);  // == (float)(CV_PI/180.f): the divisor is promoted to double
        float sn, cs;
        sinf_cosf_glibc(__fmul_rn(angle, factorPI), &sn, &cs);
        rot[tid] = make_float2(sn, cs);
        b.kps[(size_t)img * g.K + wks[tid].dst].angle = angle;
    }
    __syncthreads();

    // ---- phase 3: steered BRIEF ----
    const float RND = 12582912.f;  // 1.5 * 2^23: (x + RND) has the round-half-even integer of x in its low mantissa bits
    float4 pq[8];  // this lane's eight pattern pairs stay in registers over the warp's keypoints
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) pq[bit] = pat4[bit * 32 + lane];
    for (int i = 0; i < cnt; ++i) {
        const KpWork wk = lds_work(wksS + 8u * (uint32_t)i);
        float sn, cs;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(sn), "=f"(cs) : "r"(rotS + 8u * (uint32_t)i));
        wait_stage(i & 1);
        // sample index r * PB_ROW + c from the biased float bits (bias B = 0x4B400000 each), in wrap-around 32-bit arithmetic
        const uint32_t BIAS = (uint32_t)(0x4B400000ull * (PB_ROW + 1));
        // the window centre's shared address with both float biases taken off
        const uint32_t cbK = stageS + (uint32_t)(i & 1) * ST_BYTES + (18 * PB_ROW + 18) + (uint32_t)((kMarginX + wk.x - 18) & 15) - BIAS;
        int val = 0;
#pragma unroll
        for (int bit = 0; bit < 8; ++bit) {
            const float4 q = pq[bit];
            const int r0 = __float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(q.x, sn), __fmul_rn(q.y, cs)), RND));
            const int c0 = __float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(q.x, cs), __fmul_rn(q.y, sn)), RND));
            const int r1 = __float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(q.z, sn), __fmul_rn(q.w, cs)), RND));
            const int c1 = __float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(q.z, cs), __fmul_rn(q.w, sn)), RND));
            uint32_t t0, t1;  // address = r * PB_ROW + (c + cbK): one add + one IMAD per sample
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t0) : "r"((uint32_t)r0 * PB_ROW + ((uint32_t)c0 + cbK)));
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t1) : "r"((uint32_t)r1 * PB_ROW + ((uint32_t)c1 + cbK)));
            val |= (t0 < t1) << bit;
        }
        b.desc[((size_t)img * g.K + wk.dst) * 32 + lane] = (uint8_t)val;
        __syncwarp();
        if (lane == i + 2 && i + 2 < cnt) issueB(i + 2);
    }
}

cudaError_t describe_configure() {
    return raise_dynamic_smem(k_describe, (size_t)DW * 2 * ST_BYTES);
}

void launch_describe(const Geom& g, const Buffers& b, const DescMaps& maps, int nImages, cudaStream_t s) {
    dim3 grid((g.K + DW * KPW - 1) / (DW * KPW), nImages);
    k_describe<<<grid, 32 * DW, DW * 2 * ST_BYTES, s>>>(g, b, maps);
}

// ---------------------------------------------------------------------------------------------------------------
// test hooks: the float routines and the sort emulation evaluated on the device
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_debug_sincos(const float* deg, float* os, float* oc, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float factorPI = (float)(3.14159265358979323846 / 180.0);
    sinf_cosf_glibc(__fmul_rn(deg[i], factorPI), os + i, oc + i);
}
__global__ void k_debug_atan2(const float* y, const float* x, float* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fast_atan2_deg(y[i], x[i]);
}
void launch_debug_sincos(const float* deg, float* os, float* oc, int n, cudaStream_t s) {
    k_debug_sincos<<<(n + 255) / 256, 256, 0, s>>>(deg, os, oc, n);
}
void launch_debug_atan2(const float* y, const float* x, float* out, int n, cudaStream_t s) {
    k_debug_atan2<<<(n + 255) / 256, 256, 0, s>>>(y, x, out, n);
}

}  // namespace orbx
