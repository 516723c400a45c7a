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
// One warp per keypoint.  Both patches are first staged in shared memory with aligned 32-bit loads that are all in flight
// together (9 + 13 loads per lane instead of 31 + 16 dependent byte gathers): the 31x31 patch of the unblurred level for the
// moments, the 37x37 patch of the blurred level for the 512 rotated samples (|offset| <= 18: pattern radius 18.38).
constexpr int DW = 8;                      // warps (keypoints) per CTA
constexpr int PA_W = 9, PA_ROWS = 31;      // IC_Angle patch: 31 rows x 9 words
constexpr int PB_W = 11, PB_ROWS = 37;     // descriptor patch: 37 rows x 11 words

__global__ void __launch_bounds__(32 * DW) k_describe(const __grid_constant__ Geom g, const Buffers b) {
    __shared__ signed char pat[1024];
    __shared__ uint32_t patchA[DW][PA_ROWS * PA_W];
    __shared__ uint32_t patchB[DW][PB_ROWS * PB_W];
    const int img = b.imgBase + blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    reinterpret_cast<int*>(pat)[tid] = reinterpret_cast<const int*>(c_pattern)[tid];
    __syncthreads();
    const int n = b.outCount[img];
    const int k = blockIdx.x * DW + wid;
    if (k >= n) return;
    const KpWork wk = b.work[(size_t)img * g.K + k];
    const LevelGeom& L = g.lv[wk.level];
    const size_t plane = L.planeOff + (size_t)img * L.planeBytes;

    // ---- stage both patches ----
    const int xa = (wk.x - kHalfPatch) & ~3, sa = (wk.x - kHalfPatch) - xa;
    const int xb = (wk.x - 18) & ~3, sb = (wk.x - 18) - xb;
    {
        const uint8_t* A = b.pyr + plane + (size_t)(wk.y - kHalfPatch) * L.pitch + xa;
        const uint8_t* B = b.blur + plane + (size_t)(wk.y - 18) * L.pitch + xb;
        uint32_t va[9], vb[13];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int e = lane + 32 * i, r = e / PA_W, c = e - r * PA_W;
            va[i] = e < PA_ROWS * PA_W ? __ldg(reinterpret_cast<const uint32_t*>(A + (size_t)r * L.pitch) + c) : 0u;
        }
#pragma unroll
        for (int i = 0; i < 13; ++i) {
            const int e = lane + 32 * i, r = e / PB_W, c = e - r * PB_W;
            vb[i] = e < PB_ROWS * PB_W ? __ldg(reinterpret_cast<const uint32_t*>(B + (size_t)r * L.pitch) + c) : 0u;
        }
#pragma unroll
        for (int i = 0; i < 9; ++i)
            if (lane + 32 * i < PA_ROWS * PA_W) patchA[wid][lane + 32 * i] = va[i];
#pragma unroll
        for (int i = 0; i < 13; ++i)
            if (lane + 32 * i < PB_ROWS * PB_W) patchB[wid][lane + 32 * i] = vb[i];
    }
    __syncwarp();

    // ---- IC_Angle: lane u-15 owns patch column u ----
    const int u = lane - kHalfPatch;
    const int au = u < 0 ? -u : u;
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const uint8_t* pa = reinterpret_cast<const uint8_t*>(patchA[wid]) + sa + lane;
        int colsum = 0;
#pragma unroll
        for (int v = -kHalfPatch; v <= kHalfPatch; ++v) {
            const int av = v < 0 ? -v : v;
            if (au <= c_umax[av]) {
                const int I = pa[(v + kHalfPatch) * (PA_W * 4)];
                colsum += I;
                m01 += v * I;
            }
        }
        m10 = u * colsum;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xFFFFFFFFu, m10, o);
        m01 += __shfl_xor_sync(0xFFFFFFFFu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- steered BRIEF ----
    const float factorPI = (float)(3.14159265358979323846 / 180.0);  // == (float)(CV_PI/180.f): the divisor is promoted to double
    float sn, cs;
    sinf_cosf_glibc(__fmul_rn(angle, factorPI), &sn, &cs);
    const uint8_t* cb = reinterpret_cast<const uint8_t*>(patchB[wid]) + 18 * (PB_W * 4) + 18 + sb;  // the keypoint inside the staged patch
    const signed char* p = pat + lane * 32;
    int val = 0;
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        const float x0 = (float)p[4 * bit], y0 = (float)p[4 * bit + 1], x1 = (float)p[4 * bit + 2], y1 = (float)p[4 * bit + 3];
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, sn), __fmul_rn(y0, cs)));
        const int q0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, cs), __fmul_rn(y0, sn)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, sn), __fmul_rn(y1, cs)));
        const int q1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, cs), __fmul_rn(y1, sn)));
        const int t0 = cb[r0 * (PB_W * 4) + q0], t1 = cb[r1 * (PB_W * 4) + q1];
        val |= (t0 < t1) << bit;
    }
    b.desc[((size_t)img * g.K + wk.dst) * 32 + lane] = (uint8_t)val;
    if (lane == 0) b.kps[(size_t)img * g.K + wk.dst].angle = angle;
}

void launch_describe(const Geom& g, const Buffers& b, int nImages, cudaStream_t s) {
    dim3 grid((g.K + DW - 1) / DW, nImages);
    k_describe<<<grid, 32 * DW, 0, s>>>(g, b);
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
