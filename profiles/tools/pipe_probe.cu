// pipe_probe.cu — issue rate of the integer instructions the ORB kernels lean on, alone and interleaved with IMAD, on one B200.
// For every op: 8 independent dependency chains per thread, 1024 threads per CTA, one CTA per SM; reports warp-instructions
// per clock per SM (the ALU and FMA pipes each sustain 2 at best; 4 = both pipes busy).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 512
template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    if (OP == 0) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 1) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 2) asm volatile("mad.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 3) asm volatile("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 4) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 5) r = __vimin3_u16x2(a, b, c);
    else if (OP == 6) r = __vabsdiffu4(a, b) + 0 * c;
    else if (OP == 7) r = __popc(a ^ c) ;
    else if (OP == 8) r = __byte_perm(a, b, c);
    else if (OP == 9) r = __funnelshift_r(a, b, 8) ;
    else if (OP == 10) asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    else if (OP == 11) asm volatile("shr.u32 %0, %1, 4;" : "=r"(r) : "r"(a));
    else if (OP == 12) r = __vminu2(a, b);
    else if (OP == 13) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 14) asm volatile("add.u32 %0, %1, 0x6b6b6b6b;" : "=r"(r) : "r"(a));                       // VIADD with an immediate
    else if (OP == 15) asm volatile("add.u32 %0, %1, %2;\n\tadd.u32 %0, %0, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));   // IADD3
    else if (OP == 16) asm volatile("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));                         // HMNMX2
    else if (OP == 17) asm volatile("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));     // HFMA2.RELU
    else if (OP == 18) asm volatile("min.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));                           // VIMNMX.U32
    else if (OP == 19) asm volatile("{.reg .pred p; setp.ne.u32 p, %1, %2; selp.u32 %0, %1, %3, p;}" : "=r"(r) : "r"(a), "r"(b), "r"(c));  // ISETP + SEL
    else if (OP == 20) asm volatile("{.reg .pred p; setp.ne.u32 p, %1, %2; vote.sync.ballot.b32 %0, p, 0xffffffff;}" : "=r"(r) : "r"(a), "r"(b));  // ISETP + VOTE
    else if (OP == 21) asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(r) : "r"(a), "r"(b));                      // what ptxas makes of x * 1 + y
    else if (OP == 22) asm volatile("lop3.b32 %0, %1, %2, %3, 0xfe;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 23) asm volatile("shf.r.wrap.b32 %0, %1, %2, 24;" : "=r"(r) : "r"(a), "r"(b));
    else if (OP == 24) asm volatile("vmin2.u32.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else r = a;
    return r;
}

template <int OP, int MIX>
__global__ void __launch_bounds__(1024) k(uint32_t* out, uint32_t s0, uint32_t s1, long long* clk) {
    uint32_t x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 7 + i + s0; y[i] = threadIdx.x + i * 3 + s1; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i] = op<OP>(x[i], s0, s1);
            if (MIX) y[i] = op<1>(y[i], s1, s0);
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += x[i] + y[i];
    out[blockIdx.x * 1024 + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP, int MIX>
void run(const char* name, uint32_t* out, long long* clk) {
    k<OP, MIX><<<148, 1024>>>(out, 3, 5, clk);
    k<OP, MIX><<<148, 1024>>>(out, 3, 5, clk);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i];
    avg /= 148;
    const double winstr = 32.0 * ITERS * 8 * (MIX ? 2 : 1);
    printf("%-28s %s  %.2f warp-instr/clk/SM\n", name, MIX ? "+IMAD" : "alone", winstr / avg);
}

// shared-memory byte gathers: 17 scattered LDS.U8 per lane (the exact-score stage's access pattern) against conflict-free word loads
template <int MODE>
__global__ void __launch_bounds__(256) kg(uint32_t* out, long long* clk, int stride) {
    __shared__ uint8_t tile[48 * 640];
    for (int i = threadIdx.x; i < 48 * 640; i += 256) tile[i] = (uint8_t)(i * 7);
    __syncthreads();
    uint32_t acc = 0;
    uint32_t off = 4 * 640 + ((threadIdx.x * 2654435761u) >> 8) % (40 * 640 - 8);
    if (MODE == 1) off = 4 * 640 + (threadIdx.x & 31) * 4 + (threadIdx.x >> 5) * 640;     // conflict free, consecutive words
    if (MODE == 2) off = 4 * 640 + (threadIdx.x & 31) * 9 + (threadIdx.x >> 5) * 640;     // pixels ~9 apart in one row (flagged pixels in raster order)
    const long long t0 = clock64();
    for (int it = 0; it < 256; ++it) {
        const uint8_t* p = tile + off;
#pragma unroll
        for (int k = 0; k < 17; ++k) acc += p[(k % 7 - 3) * 640 + (k % 5) - 2];
        off = (off + stride) % (40 * 640) + 4 * 640;
    }
    const long long t1 = clock64();
    out[blockIdx.x * 256 + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int MODE>
void gather1(const char* name, uint32_t* out, long long* clk) {
    kg<MODE><<<148 * 4, 256>>>(out, clk, 1283);
    kg<MODE><<<148 * 4, 256>>>(out, clk, 1283);
    cudaDeviceSynchronize();
    long long h[148 * 4];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148 * 4; ++i) avg += (double)h[i];
    avg /= 148 * 4;
    // 4 CTAs x 8 warps per SM, 256 x 17 LDS.U8 per warp
    printf("%-40s %.2f clk per warp-wide LDS.U8 per SM (4 CTAs x 8 warps resident)\n", name, avg / (256.0 * 17 * 32));
}
void gather() {
    uint32_t* out; long long* clk;
    cudaMalloc(&out, 148 * 4 * 256 * 4); cudaMalloc(&clk, 148 * 4 * 8);
    gather1<0>("LDS.U8 gather, random offsets", out, clk);
    gather1<1>("LDS.U8 gather, consecutive words", out, clk);
    gather1<2>("LDS.U8 gather, pixels 9 apart in a row", out, clk);
}

int main() {
    uint32_t* out; long long* clk;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
#define R(OP, NAME) run<OP, 0>(NAME, out, clk); run<OP, 1>(NAME, out, clk);
    R(0, "LOP3") R(1, "IMAD (mad.lo)") R(2, "IMAD.HI (mad.hi)") R(3, "IDP.4A") R(4, "IDP.2A") R(5, "VIMNMX3.U16x2") R(6, "VABSDIFF4")
    R(7, "POPC(+LOP)") R(8, "PRMT") R(9, "SHF (funnel)") R(10, "IADD") R(11, "SHR") R(12, "VIMNMX.U16x2") R(13, "VABSDIFF4.add")
    R(14, "add imm (VIADD)") R(15, "add a+b+c (IADD3)") R(16, "min.f16x2 (HMNMX2)") R(17, "fma.relu.f16x2") R(18, "min.u32") R(19, "ISETP+SEL (2 instr)")
    R(20, "ISETP+VOTE (2 instr)") R(21, "mad x*1+y") R(22, "LOP3 0xfe") R(23, "SHF.R.W imm") R(24, "vmin2 (emulated?)")
    gather();
    return 0;
}
