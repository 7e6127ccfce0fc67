// Measured issue rates of the instructions that bind the ICSPCodec kernels on B200 (sm_100a).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_peaks tools/microbench_peaks.cu
//   tools/microbench_peaks > profiles/r02_peaks.json          (on a GPU box)
//
// Every test is one CTA of 1024 threads per SM (8 warps per SM sub-partition), each warp running a long unrolled
// sequence of INDEPENDENT instructions of one kind (8 chains), so the pipe of that instruction is the only limit.
// Rates are reported as warp-instructions per clock per SM, measured with clock64() inside the kernel (median over
// the CTAs), i.e. they do not depend on the SM clock.  bench.py reads profiles/r02_peaks.json for its rooflines:
//   * VABSDIFF4.U8.ACC  -> the INT ceiling of the motion-estimation kernel (SURVEY.md 8d asks for it to be measured)
//   * LDS.32 / LDS.64 / LDS.128 conflict-free -> the shared-memory ceiling of the same kernel
//   * DADD / DMUL / DFMA -> the FP64 ceiling of the DCT / IDCT kernels
//   * I2F.F64 / F2I.F64 and FP64 + integer mixes -> what the non-FP64 instructions of those kernels cost
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int THREADS = 1024;
constexpr int ITERS = 256;     // loop trips
constexpr int UNROLL = 8;      // groups per trip; each group = 8 independent instructions

struct Out { long long clocks; unsigned sink; };

#define BEGIN_TEST(name)                                                                         \
    __global__ void __launch_bounds__(THREADS) name(Out* out, const int* __restrict__ seed)       \
    {                                                                                             \
        extern __shared__ __align__(16) unsigned char smem[];                                     \
        (void)smem;                                                                               \
        const int s0 = seed[threadIdx.x & 31];
#define TIMED_LOOP_BEGIN                                                                          \
        __syncthreads();                                                                          \
        const long long t0 = clock64();                                                           \
        _Pragma("unroll 1") for (int it = 0; it < ITERS; it++) {                                  \
            _Pragma("unroll") for (int u = 0; u < UNROLL; u++) {
#define TIMED_LOOP_END(sinkexpr)                                                                  \
            }                                                                                     \
        }                                                                                         \
        __syncthreads();                                                                          \
        const long long t1 = clock64();                                                           \
        if (threadIdx.x == 0) out[blockIdx.x].clocks = t1 - t0;                                   \
        if ((sinkexpr) == 0x7fffffffu) out[blockIdx.x].sink = 1;                                  \
    }

// ---- FP64 ------------------------------------------------------------------------------------------------------
#define D8(op)  asm volatile(op " %0, %0, %8;\n\t" op " %1, %1, %8;\n\t" op " %2, %2, %8;\n\t" op " %3, %3, %8;\n\t" \
                             op " %4, %4, %8;\n\t" op " %5, %5, %8;\n\t" op " %6, %6, %8;\n\t" op " %7, %7, %8;"      \
                             : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+d"(a5), "+d"(a6), "+d"(a7) : "d"(k))
#define DECL_D8 double a0 = s0, a1 = s0 + 1, a2 = s0 + 2, a3 = s0 + 3, a4 = s0 + 4, a5 = s0 + 5, a6 = s0 + 6, a7 = s0 + 7; const double k = 1.0000001 + s0 * 1e-9;
#define SINK_D8 (unsigned)__double2loint(a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7)

BEGIN_TEST(t_dadd) DECL_D8 TIMED_LOOP_BEGIN D8("add.rn.f64"); TIMED_LOOP_END(SINK_D8)
BEGIN_TEST(t_dmul) DECL_D8 TIMED_LOOP_BEGIN D8("mul.rn.f64"); TIMED_LOOP_END(SINK_D8)
BEGIN_TEST(t_dfma) DECL_D8 TIMED_LOOP_BEGIN
    asm volatile("fma.rn.f64 %0, %0, %8, %8;\n\tfma.rn.f64 %1, %1, %8, %8;\n\tfma.rn.f64 %2, %2, %8, %8;\n\tfma.rn.f64 %3, %3, %8, %8;\n\t"
                 "fma.rn.f64 %4, %4, %8, %8;\n\tfma.rn.f64 %5, %5, %8, %8;\n\tfma.rn.f64 %6, %6, %8, %8;\n\tfma.rn.f64 %7, %7, %8, %8;"
                 : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+d"(a5), "+d"(a6), "+d"(a7) : "d"(k));
TIMED_LOOP_END(SINK_D8)
// DADD with the round-toward-zero modifier (the magic-constant truncation uses it)
BEGIN_TEST(t_dadd_rz) DECL_D8 TIMED_LOOP_BEGIN D8("add.rz.f64"); TIMED_LOOP_END(SINK_D8)

// ---- conversions -----------------------------------------------------------------------------------------------
#define DECL_I8 int i0 = s0, i1 = s0 + 1, i2 = s0 + 2, i3 = s0 + 3, i4 = s0 + 4, i5 = s0 + 5, i6 = s0 + 6, i7 = s0 + 7;
BEGIN_TEST(t_i2f64) DECL_I8 double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0; TIMED_LOOP_BEGIN
    asm volatile("cvt.rn.f64.s32 %0, %8;\n\tcvt.rn.f64.s32 %1, %9;\n\tcvt.rn.f64.s32 %2, %10;\n\tcvt.rn.f64.s32 %3, %11;\n\t"
                 "cvt.rn.f64.s32 %4, %12;\n\tcvt.rn.f64.s32 %5, %13;\n\tcvt.rn.f64.s32 %6, %14;\n\tcvt.rn.f64.s32 %7, %15;"
                 : "=d"(a0), "=d"(a1), "=d"(a2), "=d"(a3), "=d"(a4), "=d"(a5), "=d"(a6), "=d"(a7)
                 : "r"(i0), "r"(i1), "r"(i2), "r"(i3), "r"(i4), "r"(i5), "r"(i6), "r"(i7));
    i0 += __double2hiint(a0) & 1; i1 += __double2hiint(a1) & 1; i2 += __double2hiint(a2) & 1; i3 += __double2hiint(a3) & 1;   // loop-variant inputs
    i4 += __double2hiint(a4) & 1; i5 += __double2hiint(a5) & 1; i6 += __double2hiint(a6) & 1; i7 += __double2hiint(a7) & 1;   // (8 LOP3.LUT-class ALU ops per 8 converts)
TIMED_LOOP_END(SINK_D8)
BEGIN_TEST(t_f2i64) DECL_D8 int i0 = 0, i1 = 0, i2 = 0, i3 = 0, i4 = 0, i5 = 0, i6 = 0, i7 = 0; (void)k; TIMED_LOOP_BEGIN
    asm volatile("cvt.rzi.s32.f64 %0, %8;\n\tcvt.rzi.s32.f64 %1, %9;\n\tcvt.rzi.s32.f64 %2, %10;\n\tcvt.rzi.s32.f64 %3, %11;\n\t"
                 "cvt.rzi.s32.f64 %4, %12;\n\tcvt.rzi.s32.f64 %5, %13;\n\tcvt.rzi.s32.f64 %6, %14;\n\tcvt.rzi.s32.f64 %7, %15;"
                 : "=r"(i0), "=r"(i1), "=r"(i2), "=r"(i3), "=r"(i4), "=r"(i5), "=r"(i6), "=r"(i7)
                 : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(a4), "d"(a5), "d"(a6), "d"(a7));
    a0 = __hiloint2double(__double2hiint(a0), i0); a1 = __hiloint2double(__double2hiint(a1), i1); a2 = __hiloint2double(__double2hiint(a2), i2);
    a3 = __hiloint2double(__double2hiint(a3), i3); a4 = __hiloint2double(__double2hiint(a4), i4); a5 = __hiloint2double(__double2hiint(a5), i5);
    a6 = __hiloint2double(__double2hiint(a6), i6); a7 = __hiloint2double(__double2hiint(a7), i7);   // low word := result (a register move at most)
TIMED_LOOP_END((unsigned)(i0 + i1 + i2 + i3 + i4 + i5 + i6 + i7))
// I2F.F32 (full-rate ALU conversion) for comparison
BEGIN_TEST(t_i2f32) DECL_I8 float a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0; TIMED_LOOP_BEGIN
    asm volatile("cvt.rn.f32.s32 %0, %8;\n\tcvt.rn.f32.s32 %1, %9;\n\tcvt.rn.f32.s32 %2, %10;\n\tcvt.rn.f32.s32 %3, %11;\n\t"
                 "cvt.rn.f32.s32 %4, %12;\n\tcvt.rn.f32.s32 %5, %13;\n\tcvt.rn.f32.s32 %6, %14;\n\tcvt.rn.f32.s32 %7, %15;"
                 : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3), "=f"(a4), "=f"(a5), "=f"(a6), "=f"(a7)
                 : "r"(i0), "r"(i1), "r"(i2), "r"(i3), "r"(i4), "r"(i5), "r"(i6), "r"(i7));
    i0 += __float_as_int(a0) & 1; i1 += __float_as_int(a1) & 1; i2 += __float_as_int(a2) & 1; i3 += __float_as_int(a3) & 1;
    i4 += __float_as_int(a4) & 1; i5 += __float_as_int(a5) & 1; i6 += __float_as_int(a6) & 1; i7 += __float_as_int(a7) & 1;
TIMED_LOOP_END((unsigned)__float_as_int(a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7))

// ---- integer ---------------------------------------------------------------------------------------------------
#define I8(op) asm volatile(op " %0, %0, %8, %0;\n\t" op " %1, %1, %8, %1;\n\t" op " %2, %2, %8, %2;\n\t" op " %3, %3, %8, %3;\n\t" \
                            op " %4, %4, %8, %4;\n\t" op " %5, %5, %8, %5;\n\t" op " %6, %6, %8, %6;\n\t" op " %7, %7, %8, %7;"      \
                            : "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3), "+r"(i4), "+r"(i5), "+r"(i6), "+r"(i7) : "r"(kk))
#define SINK_I8 (unsigned)(i0 + i1 + i2 + i3 + i4 + i5 + i6 + i7)
BEGIN_TEST(t_vabsdiff4) DECL_I8 const int kk = s0 * 0x01020304 + 0x10203040; TIMED_LOOP_BEGIN I8("vabsdiff4.u32.u32.u32.add"); TIMED_LOOP_END(SINK_I8)
BEGIN_TEST(t_imad) DECL_I8 const int kk = s0 | 3; TIMED_LOOP_BEGIN I8("mad.lo.s32"); TIMED_LOOP_END(SINK_I8)
BEGIN_TEST(t_imadhi) DECL_I8 const int kk = s0 | 0x10000003; TIMED_LOOP_BEGIN I8("mad.hi.s32"); TIMED_LOOP_END(SINK_I8)
BEGIN_TEST(t_iadd3) DECL_I8 const int kk = s0 | 3; TIMED_LOOP_BEGIN
    asm volatile("add.s32 %0, %0, %8;\n\tadd.s32 %1, %1, %8;\n\tadd.s32 %2, %2, %8;\n\tadd.s32 %3, %3, %8;\n\t"
                 "add.s32 %4, %4, %8;\n\tadd.s32 %5, %5, %8;\n\tadd.s32 %6, %6, %8;\n\tadd.s32 %7, %7, %8;"
                 : "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3), "+r"(i4), "+r"(i5), "+r"(i6), "+r"(i7) : "r"(kk));
TIMED_LOOP_END(SINK_I8)
BEGIN_TEST(t_lop3) DECL_I8 const int kk = s0 | 0x55; TIMED_LOOP_BEGIN
    asm volatile("lop3.b32 %0, %0, %8, %1, 0x96;\n\tlop3.b32 %1, %1, %8, %2, 0x96;\n\tlop3.b32 %2, %2, %8, %3, 0x96;\n\tlop3.b32 %3, %3, %8, %4, 0x96;\n\t"
                 "lop3.b32 %4, %4, %8, %5, 0x96;\n\tlop3.b32 %5, %5, %8, %6, 0x96;\n\tlop3.b32 %6, %6, %8, %7, 0x96;\n\tlop3.b32 %7, %7, %8, %0, 0x96;"
                 : "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3), "+r"(i4), "+r"(i5), "+r"(i6), "+r"(i7) : "r"(kk));
TIMED_LOOP_END(SINK_I8)
BEGIN_TEST(t_prmt) DECL_I8 const int kk = s0 | 0x3210; TIMED_LOOP_BEGIN
    asm volatile("prmt.b32 %0, %0, %1, %8;\n\tprmt.b32 %1, %1, %2, %8;\n\tprmt.b32 %2, %2, %3, %8;\n\tprmt.b32 %3, %3, %4, %8;\n\t"
                 "prmt.b32 %4, %4, %5, %8;\n\tprmt.b32 %5, %5, %6, %8;\n\tprmt.b32 %6, %6, %7, %8;\n\tprmt.b32 %7, %7, %0, %8;"
                 : "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3), "+r"(i4), "+r"(i5), "+r"(i6), "+r"(i7) : "r"(kk));
TIMED_LOOP_END(SINK_I8)
BEGIN_TEST(t_shf) DECL_I8 const int kk = (s0 & 7) + 8; TIMED_LOOP_BEGIN
    asm volatile("shf.r.wrap.b32 %0, %0, %1, %8;\n\tshf.r.wrap.b32 %1, %1, %2, %8;\n\tshf.r.wrap.b32 %2, %2, %3, %8;\n\tshf.r.wrap.b32 %3, %3, %4, %8;\n\t"
                 "shf.r.wrap.b32 %4, %4, %5, %8;\n\tshf.r.wrap.b32 %5, %5, %6, %8;\n\tshf.r.wrap.b32 %6, %6, %7, %8;\n\tshf.r.wrap.b32 %7, %7, %0, %8;"
                 : "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3), "+r"(i4), "+r"(i5), "+r"(i6), "+r"(i7) : "r"(kk));
TIMED_LOOP_END(SINK_I8)

// ---- FP64 mixed with integer work (what the DCT / IDCT kernels look like to the scheduler) -----------------------
// NI integer instructions (alternating IADD3-class on the ALU pipe and IMAD on the FMA pipe) per DADD
template <int NI>
__global__ void __launch_bounds__(THREADS) t_dadd_mix(Out* out, const int* __restrict__ seed)
{
    const int s0 = seed[threadIdx.x & 31];
    DECL_D8 DECL_I8 const int kk = s0 | 3;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            // 8 DADD interleaved with 8*NI integer instructions
#define MIXSTEP(A, I)                                                                                  \
            asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(A) : "d"(k));                                  \
            if (NI >= 1) asm volatile("add.s32 %0, %0, %1;" : "+r"(I) : "r"(kk));                        \
            if (NI >= 2) asm volatile("mad.lo.s32 %0, %0, %1, %0;" : "+r"(I) : "r"(kk));                 \
            if (NI >= 3) asm volatile("xor.b32 %0, %0, %1;" : "+r"(I) : "r"(kk));                        \
            if (NI >= 4) asm volatile("mad.lo.s32 %0, %0, %1, %0;" : "+r"(I) : "r"(kk));
            MIXSTEP(a0, i0) MIXSTEP(a1, i1) MIXSTEP(a2, i2) MIXSTEP(a3, i3) MIXSTEP(a4, i4) MIXSTEP(a5, i5) MIXSTEP(a6, i6) MIXSTEP(a7, i7)
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x].clocks = t1 - t0;
    if (SINK_D8 + SINK_I8 == 0x7fffffffu) out[blockIdx.x].sink = 1;
}


// do the 64-bit conversions share the FP64 pipe?  8 DADD + 2 conversions per group: if they run on their own pipe the DADD
// rate stays ~2.0/clk/SM, if they share it drops to ~1.0
template <int KIND>   // 0: I2F.F64.S32, 1: F2I.S32.F64
__global__ void __launch_bounds__(THREADS) t_dadd_cvt(Out* out, const int* __restrict__ seed)
{
    const int s0 = seed[threadIdx.x & 31];
    DECL_D8 int i0 = s0, i1 = s0 + 1; double c0 = s0, c1 = s0 + 1;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            D8("add.rn.f64");
            if (KIND == 0) {
                asm volatile("cvt.rn.f64.s32 %0, %2;\n\tcvt.rn.f64.s32 %1, %3;" : "=d"(c0), "=d"(c1) : "r"(i0), "r"(i1));
                i0 += __double2hiint(c0) & 1; i1 += __double2hiint(c1) & 1;
            } else {
                asm volatile("cvt.rzi.s32.f64 %0, %2;\n\tcvt.rzi.s32.f64 %1, %3;" : "=r"(i0), "=r"(i1) : "d"(c0), "d"(c1));
                c0 = __hiloint2double(__double2hiint(c0), i0); c1 = __hiloint2double(__double2hiint(c1), i1);
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x].clocks = t1 - t0;
    if (SINK_D8 + (unsigned)i0 + (unsigned)i1 + (unsigned)__double2loint(c0 + c1) == 0x7fffffffu) out[blockIdx.x].sink = 1;
}

// ---- shared memory -----------------------------------------------------------------------------------------------
// every lane reads consecutive words/vectors (conflict free); addresses advance so that nothing is hoisted
template <int WIDTH>   // bytes per lane: 4, 8, 16
__global__ void __launch_bounds__(THREADS) t_lds(Out* out, const int* __restrict__ seed)
{
    extern __shared__ __align__(16) unsigned char smem[];
    for (int i = threadIdx.x; i < 65536 / 4; i += THREADS) ((int*)smem)[i] = seed[i & 31];
    // lane l of every warp reads bytes [l*WIDTH, (l+1)*WIDTH) of a 32*WIDTH-byte row: conflict free for every width
    const unsigned base = (unsigned)__cvta_generic_to_shared(smem) + (threadIdx.x & 31) * WIDTH;
    unsigned v0 = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0, v5 = 0, v6 = 0, v7 = 0, w = 0;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        const unsigned a = base + (it & 3) * 8192;
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const unsigned ad = a + (u * 8 + j) * 32 * WIDTH;   // 64 distinct rows of 32*WIDTH bytes per trip
                unsigned x, y, z, q;
                if (WIDTH == 4) { asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(ad)); v0 ^= x; }
                if (WIDTH == 8) { asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(ad)); v0 = v0 ^ x ^ y; }
                if (WIDTH == 16) { asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(q) : "r"(ad)); v0 = v0 ^ x ^ y; v1 = v1 ^ z ^ q; }
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x].clocks = t1 - t0;
    if ((v0 ^ v1 ^ v2 ^ v3 ^ v4 ^ v5 ^ v6 ^ v7 ^ w) == 0x7fffffffu) out[blockIdx.x].sink = 1;
}
// the motion-estimation inner loop in its ideal form: one conflict-free LDS.32 feeding one VABSDIFF4.U8.ACC
__global__ void __launch_bounds__(THREADS) t_lds_sad(Out* out, const int* __restrict__ seed)
{
    extern __shared__ __align__(16) unsigned char smem[];
    for (int i = threadIdx.x; i < 32768 / 4; i += THREADS) ((int*)smem)[i] = seed[i & 31];
    const unsigned base = (unsigned)__cvta_generic_to_shared(smem) + (threadIdx.x & 31) * 4 + (threadIdx.x >> 5) * 128;
    unsigned a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    const unsigned c0 = seed[0], c1 = seed[1], c2 = seed[2], c3 = seed[3];
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        const unsigned a = base + (it & 3) * 4096;
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            unsigned v0, v1, v2, v3, v4, v5, v6, v7;
            asm volatile("ld.shared.u32 %0, [%8];\n\tld.shared.u32 %1, [%8+416];\n\tld.shared.u32 %2, [%8+832];\n\tld.shared.u32 %3, [%8+1248];\n\t"
                         "ld.shared.u32 %4, [%8+1664];\n\tld.shared.u32 %5, [%8+2080];\n\tld.shared.u32 %6, [%8+2496];\n\tld.shared.u32 %7, [%8+2912];"
                         : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3), "=r"(v4), "=r"(v5), "=r"(v6), "=r"(v7) : "r"(a + u * 4));
            asm volatile("vabsdiff4.u32.u32.u32.add %0, %4, %8, %0;\n\tvabsdiff4.u32.u32.u32.add %1, %5, %9, %1;\n\t"
                         "vabsdiff4.u32.u32.u32.add %2, %6, %10, %2;\n\tvabsdiff4.u32.u32.u32.add %3, %7, %11, %3;"
                         : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3) : "r"(v0), "r"(v1), "r"(v2), "r"(v3), "r"(c0), "r"(c1), "r"(c2), "r"(c3));
            asm volatile("vabsdiff4.u32.u32.u32.add %0, %4, %8, %0;\n\tvabsdiff4.u32.u32.u32.add %1, %5, %9, %1;\n\t"
                         "vabsdiff4.u32.u32.u32.add %2, %6, %10, %2;\n\tvabsdiff4.u32.u32.u32.add %3, %7, %11, %3;"
                         : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3) : "r"(v4), "r"(v5), "r"(v6), "r"(v7), "r"(c0), "r"(c1), "r"(c2), "r"(c3));
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x].clocks = t1 - t0;
    if (a0 + a1 + a2 + a3 == 0x7fffffffu) out[blockIdx.x].sink = 1;
}

// -----------------------------------------------------------------------------------------------------------------
struct Result { std::string name; double per_clk_sm; double extra; std::string note; };

template <typename K>
static double run(K kern, int nsm, Out* d_out, const int* d_seed, size_t smem, double warp_instr_per_warp)
{
    std::vector<Out> h(nsm);
    for (int rep = 0; rep < 3; rep++) {
        kern<<<nsm, THREADS, smem>>>(d_out, d_seed);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h.data(), d_out, nsm * sizeof(Out), cudaMemcpyDeviceToHost));
    std::vector<long long> c;
    for (auto& o : h) c.push_back(o.clocks);
    std::sort(c.begin(), c.end());
    const double clocks = (double)c[c.size() / 2];
    return warp_instr_per_warp * (THREADS / 32) / clocks;     // warp-instructions per clock per SM
}

int main()
{
    int dev = 0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    const int nsm = prop.multiProcessorCount;
    Out* d_out; int* d_seed;
    CK(cudaMalloc(&d_out, nsm * sizeof(Out)));
    CK(cudaMemset(d_out, 0, nsm * sizeof(Out)));
    int seed[32];
    for (int i = 0; i < 32; i++) seed[i] = 3 + i;
    CK(cudaMalloc(&d_seed, sizeof(seed)));
    CK(cudaMemcpy(d_seed, seed, sizeof(seed), cudaMemcpyHostToDevice));
    const double n8 = (double)ITERS * UNROLL * 8;     // instructions of the measured kind per warp
    std::vector<Result> r;
    r.push_back({"DADD", run(t_dadd, nsm, d_out, d_seed, 0, n8), 0, "add.rn.f64"});
    r.push_back({"DMUL", run(t_dmul, nsm, d_out, d_seed, 0, n8), 0, "mul.rn.f64"});
    r.push_back({"DFMA", run(t_dfma, nsm, d_out, d_seed, 0, n8), 0, "fma.rn.f64"});
    r.push_back({"DADD.RZ", run(t_dadd_rz, nsm, d_out, d_seed, 0, n8), 0, "add.rz.f64"});
    r.push_back({"I2F.F64.S32", run(t_i2f64, nsm, d_out, d_seed, 0, n8), 0, "cvt.rn.f64.s32 (+2 ALU ops per 8)"});
    r.push_back({"F2I.S32.F64.TRUNC", run(t_f2i64, nsm, d_out, d_seed, 0, n8), 0, "cvt.rzi.s32.f64 (+2 ALU ops per 8)"});
    r.push_back({"I2F.F32.S32", run(t_i2f32, nsm, d_out, d_seed, 0, n8), 0, "cvt.rn.f32.s32 (+2 ALU ops per 8)"});
    r.push_back({"VABSDIFF4.U8.ACC", run(t_vabsdiff4, nsm, d_out, d_seed, 0, n8), 0, "vabsdiff4.u32.u32.u32.add"});
    r.push_back({"IMAD", run(t_imad, nsm, d_out, d_seed, 0, n8), 0, "mad.lo.s32"});
    r.push_back({"IMAD.HI", run(t_imadhi, nsm, d_out, d_seed, 0, n8), 0, "mad.hi.s32"});
    r.push_back({"IADD", run(t_iadd3, nsm, d_out, d_seed, 0, n8), 0, "add.s32"});
    r.push_back({"LOP3", run(t_lop3, nsm, d_out, d_seed, 0, n8), 0, "lop3.b32"});
    r.push_back({"PRMT", run(t_prmt, nsm, d_out, d_seed, 0, n8), 0, "prmt.b32"});
    r.push_back({"SHF", run(t_shf, nsm, d_out, d_seed, 0, n8), 0, "shf.r.wrap.b32"});
    r.push_back({"DADD+1int", run(t_dadd_mix<1>, nsm, d_out, d_seed, 0, n8), 1, "DADD rate with 1 integer instruction per DADD"});
    r.push_back({"DADD+2int", run(t_dadd_mix<2>, nsm, d_out, d_seed, 0, n8), 2, "DADD rate with 2 integer instructions per DADD"});
    r.push_back({"DADD+3int", run(t_dadd_mix<3>, nsm, d_out, d_seed, 0, n8), 3, "DADD rate with 3 integer instructions per DADD"});
    r.push_back({"DADD+4int", run(t_dadd_mix<4>, nsm, d_out, d_seed, 0, n8), 4, "DADD rate with 4 integer instructions per DADD"});
    r.push_back({"DADD+0.25xI2F.F64", run(t_dadd_cvt<0>, nsm, d_out, d_seed, 0, n8), 0, "DADD rate with 2 I2F.F64.S32 per 8 DADD (separate pipes if ~2.0)"});
    r.push_back({"DADD+0.25xF2I.F64", run(t_dadd_cvt<1>, nsm, d_out, d_seed, 0, n8), 0, "DADD rate with 2 F2I.S32.F64 per 8 DADD (separate pipes if ~2.0)"});
    CK(cudaFuncSetAttribute(t_lds<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(t_lds<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(t_lds<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(t_lds_sad, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    r.push_back({"LDS.32", run(t_lds<4>, nsm, d_out, d_seed, 65536, n8), 128, "conflict-free, 128 B per warp instruction"});
    r.push_back({"LDS.64", run(t_lds<8>, nsm, d_out, d_seed, 65536, n8), 256, "conflict-free, 256 B per warp instruction"});
    r.push_back({"LDS.128", run(t_lds<16>, nsm, d_out, d_seed, 65536, n8), 512, "conflict-free, 512 B per warp instruction"});
    r.push_back({"LDS.32+VABSDIFF4", run(t_lds_sad, nsm, d_out, d_seed, 65536, n8), 0, "pairs per clock: one LDS.32 feeding one VABSDIFF4.U8.ACC"});

    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
    printf("{\n \"gpu\": \"%s\", \"sms\": %d, \"sm_clock_khz_max\": %d, \"unit\": \"warp-instructions per clock per SM (32 lanes each)\",\n \"rates\": {\n", prop.name, nsm, clk_khz);
    for (size_t i = 0; i < r.size(); i++)
        printf("  \"%s\": {\"per_clk_sm\": %.4f, \"lanes_per_clk_sm\": %.2f, \"note\": \"%s\"}%s\n", r[i].name.c_str(), r[i].per_clk_sm, r[i].per_clk_sm * 32,
               r[i].note.c_str(), i + 1 < r.size() ? "," : "");
    printf(" }\n}\n");
    return 0;
}
