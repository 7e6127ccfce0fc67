// Transform kernels of libicspcuda, second generation (round 2).  Same arithmetic as icsp_device.cuh, bit for bit; what
// changed is everything AROUND the FP64 instructions, because the first generation issued 2.1 integer instructions per
// FP64 instruction and the measured issue model (tools/microbench_peaks.cu: DADD alone 2.0 warp-instr/clk/SM, with 2 / 3
// integer instructions per DADD 1.21 / 0.94) makes the FP64 pipe idle half of the time at that ratio:
//   * addressing: per macroblock the lane computes two byte offsets (luma / chroma) for the current frame and two for the
//     motion-compensated prediction, plus an "entirely inside the picture" flag per plane; block k only adds the uniform
//     delta Geom::d_pix[k].  The clamped fetch with the reference's zero last row/column (ENC:2244-2266) is the slow path.
//   * residual: packed 16-bit arithmetic (two pixels per instruction, biased so that no borrow crosses the halves) produces
//     the eight butterfly inputs of the exact first DCT stage; they become doubles by the 2^52 trick (one DADD each; the
//     measured I2F.F64.S32 rate is 0.49 warp-instr/clk/SM, a quarter of DADD's).
//   * scaling: the final "* 0.25" of both transforms (ENC:2738-2744, 2885-2891) is a power of two, so it commutes with every
//     rounding: it is folded into the second-stage constants (0.25*|T|), 8 DMUL less per lane and block.
//   * quantiser: trunc/floor by one F2I, division by an exact 19-bit reciprocal multiply (valid while |x|*Q < 2^19:
//     |AC| <= 4080 for 8-bit video, Q <= 128; larger Q takes the first-generation path), zig-zag scatter through eight
//     shared-memory addresses kept in registers.
#pragma once
#include "icsp_device.cuh"
#include <type_traits>

namespace icsp {

#ifndef ICSP_MAGIC_CVT
#define ICSP_MAGIC_CVT 0     // 1: ints -> doubles by 2^52 + DADD instead of I2F.F64 (measured slower: the FP64 pipe is the scarcer one)
#endif
#ifndef ICSP_MAGIC_FLOOR
#define ICSP_MAGIC_FLOOR 0   // 1: chroma floor() by DADD.RZ with 1.5*2^52 instead of F2I.F64.FLOOR (measured equal)
#endif

// [table][0..7] = irt2, |T| magnitudes (as g_mag); [table][8..15] = the same times 0.25
__device__ double g_mag2[2][16];
struct Mags2 { double m[8]; double q[8]; };
template <int TAB>
__device__ __forceinline__ Mags2 load_mags2()
{
    Mags2 M;
#pragma unroll
    for (int i = 0; i < 8; i++) { M.m[i] = __ldg(&g_mag2[TAB][i]); M.q[i] = __ldg(&g_mag2[TAB][8 + i]); }
    return M;
}
#define ICSP_TQ(M, u, x) (kTS(u, x) > 0 ? (M).q[kTI(u, x)] : -(M).q[kTI(u, x)])

// (double)(h - bias) for a small non-negative h: exact
__device__ __forceinline__ double biased_to_double(uint32_t h, int bias)
{
#if ICSP_MAGIC_CVT
    return __dsub_rn(__hiloint2double(0x43300000, (int)h), 4503599627370496.0 + (double)bias);
#else
    return (double)((int)h - bias);
#endif
}
// (double)v for any int32: exact
__device__ __forceinline__ double int_to_double(int v)
{
#if ICSP_MAGIC_CVT
    return __dsub_rn(__hiloint2double(0x43300000, v ^ 0x80000000), 4503599627370496.0 + 2147483648.0);
#else
    return (double)v;
#endif
}
__device__ __forceinline__ int floor_to_int(double x)
{
#if ICSP_MAGIC_FLOOR
    return __double2loint(__dadd_rz(x, 6755399441055744.0));   // 1.5*2^52: the sum is positive, so RZ is floor; low word = integer
#else
    return __double2int_rd(x);
#endif
}
// r / q toward zero by multiplication: m19 = floor(2^19/q) + 1, exact while |r|*q < 2^19 and |r| <= 4080
__device__ __forceinline__ int div_m19(int r, int m19) { return ((r * m19) >> 19) + (int)((unsigned)r >> 31); }

__device__ __forceinline__ void sts_u16(unsigned addr, int v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ int lds_s16(unsigned addr)
{
    int v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

// ---- addressing --------------------------------------------------------------------------------------------------
// Per macroblock a lane keeps ONE running byte offset (row r of the block being fetched, relative to the frame start)
// and the offset of the motion-compensated prediction relative to it (mvoff: -my*pitch - mx of the plane); block k+1 is
// reached by a uniform increment (Y0->Y1 +8, Y1->Y2 +8w-8, Y2->Y3 +8, Cb->Cr +cw*ch) or, once per macroblock, by
// switching to the chroma plane.  `inside` says that the prediction of the whole macroblock lies inside the picture
// (the common case): then the 8 prediction pixels are three aligned words and two funnel shifts.  Otherwise the
// reference's padding (clamp + zero last row / column, getPaddingImage ENC:2227-2269) is evaluated out of line.
__device__ __noinline__ uint2 pred_fetch_slow(const uint8_t* __restrict__ prevf, int w, int h, int cw, int ch, int k, int r, int mbx, int mby, int mx, int my)
{
    if (k < 4) return ref_row8_packed(prevf, w, h, 16, 16 + (2 * mby + (k >> 1)) * 8 + r - my, 16 + (2 * mbx + (k & 1)) * 8 - mx);
    return ref_row8_packed(prevf + w * h + (k - 4) * cw * ch, cw, ch, 8, 8 + mby * 8 + r - my / 2, 8 + mbx * 8 - mx / 2);
}
// 8 prediction pixels as fetched: three aligned words and the bit shift that aligns them.  The funnel shifts are applied
// where the pixels are USED (one block later), so that nothing waits for the loads while they are in flight.
struct PredRaw { uint32_t w0, w1, w2; int sh; };
__device__ __forceinline__ uint2 pred_pixels(const PredRaw& q) { return make_uint2(__funnelshift_r(q.w0, q.w1, q.sh), __funnelshift_r(q.w1, q.w2, q.sh)); }
__device__ __forceinline__ PredRaw pred_fetch_fast(const uint8_t* __restrict__ q)
{
    const uintptr_t u = (uintptr_t)q;
    const uint32_t* w = (const uint32_t*)(u & ~(uintptr_t)3);
    PredRaw o;
    o.sh = (int)(u & 3) * 8;
    o.w0 = __ldg(w); o.w1 = __ldg(w + 1); o.w2 = __ldg(w + 2);   // w2 may lie past the row: every buffer has 64 bytes of slack
    return o;
}
struct MbWalk {
    int off;              // byte offset of row r of the current block inside a frame
    int mvoff;            // prediction offset relative to `off` in the current plane
    int off_c, mvoff_c;   // the same for the chroma planes (Cb)
    unsigned inside;      // bit k: the 8 prediction pixels of row r of block k lie inside the picture (no clamping, no zero row/column)
    int mbx, mby, mx, my;
};
__device__ __forceinline__ MbWalk mb_walk(const Geom& g, int mbx, int mby, int r, int mx, int my, bool intra)
{
    MbWalk a;
    const int cmx = mx / 2, cmy = my / 2;                      // CmotionCompensation ENC:2538-2539: toward zero
    a.mbx = mbx; a.mby = mby; a.mx = mx; a.my = my;
    a.off_c = g.w * g.h + (mby * 8 + r) * g.cw + mbx * 8;
    a.mvoff_c = -cmy * g.cw - cmx;
    const unsigned cin = ((unsigned)(mbx * 8 - cmx) <= (unsigned)(g.cw - 8) && (unsigned)(mby * 8 + r - cmy) < (unsigned)g.ch) ? 0x30u : 0u;
    if (intra) { a.off = a.off_c; a.mvoff = 0; a.inside = 0x30u; return a; }
    a.off = (mby * 16 + r) * g.w + mbx * 16;
    a.mvoff = -my * g.w - mx;                                   // motionCompensation ENC:2185-2186
    const int ux = mbx * 16 - mx, uy = mby * 16 + r - my;
    const unsigned x0 = (unsigned)ux <= (unsigned)(g.w - 8) ? 5u : 0u, x1 = (unsigned)(ux + 8) <= (unsigned)(g.w - 8) ? 10u : 0u;   // blocks 0,2 | 1,3
    const unsigned y0 = (unsigned)uy < (unsigned)g.h ? 3u : 0u, y1 = (unsigned)(uy + 8) < (unsigned)g.h ? 12u : 0u;                 // blocks 0,1 | 2,3
    a.inside = ((x0 | x1) & (y0 | y1)) | cin;
    return a;
}
// Macroblock handled by 8-lane group number i of a frame: the 2*mbw + 2*(mbh-2) border macroblocks first, the interior
// ones after them.  Only a border macroblock can have a prediction that reaches outside the picture (|mv| <= 16), so the
// clamped-fetch path is confined to the first few warps of a frame instead of dragging one warp in five through it.
__device__ __forceinline__ int mb_border_first(const Geom& g, int i, int& mbx, int& mby)
{
    const int nb = g.mbh > 2 ? 2 * g.mbw + 2 * (g.mbh - 2) : g.nmb;
    if (g.mbw <= 2) { mby = (int)__umulhi((unsigned)i, g.magic_mbw); mbx = i - mby * g.mbw; }
    else if (i < g.mbw) { mbx = i; mby = 0; }
    else if (i < 2 * g.mbw && g.mbh > 1) { mbx = i - g.mbw; mby = g.mbh - 1; }
    else if (i < nb) { const int j = i - 2 * g.mbw; mby = 1 + (j >> 1); mbx = (j & 1) ? g.mbw - 1 : 0; }
    else { const int j = i - nb; const int q = (int)__umulhi((unsigned)j, g.magic_mbw2); mby = 1 + q; mbx = 1 + j - q * (g.mbw - 2); }
    return mby * g.mbw + mbx;
}
// move from block k to block k+1 (k < 5)
__device__ __forceinline__ void mb_walk_next(const Geom& g, MbWalk& a, int k)
{
    if (k < 3) a.off += (k & 1) ? 8 * g.w - 8 : 8;
    else if (k == 3) { a.off = a.off_c; a.mvoff = a.mvoff_c; }
    else a.off += g.cw * g.ch;
}
__device__ __forceinline__ PredRaw mb_walk_pred(const Geom& g, const MbWalk& a, const uint8_t* __restrict__ prevf, int k, int r)
{
    if ((a.inside >> k) & 1u) return pred_fetch_fast(prevf + (a.off + a.mvoff));
    const uint2 v = pred_fetch_slow(prevf, g.w, g.h, g.cw, g.ch, k, r, a.mbx, a.mby, a.mx, a.my);
    PredRaw o; o.w0 = v.x; o.w1 = v.y; o.w2 = 0u; o.sh = 0;
    return o;
}

#ifndef ICSP_TR_UNROLL
#define ICSP_TR_UNROLL 0     // 1: six copies of the body (measured slower: 43 KB of code, instruction-cache misses dominate the stalls)
#endif
#if ICSP_TR_UNROLL
#define ICSP_TR_PRAGMA _Pragma("unroll")
#else
#define ICSP_TR_PRAGMA _Pragma("unroll 1")
#endif

// =====================================================================================================
// Kernel A2: residual + forward DCT + AC quantisation + zig-zag (R3, R5, R7, R8).  8 lanes per 8x8 block, one 8-lane
// group per macroblock walking over its blocks (inter: Y0..Y3, Cb, Cr; intra: Cb, Cr of the raw pixels).
// Outputs: AC levels (the DC slot is filled by the DC chain kernel), ACflag, the scaled DC as a double in dcraw
// ([G][nmb][6], macroblock major like the levels).
// =====================================================================================================
constexpr int TR2_THREADS = 128;

template <bool INTRA, bool QFAST, bool TAP>
__global__ void __launch_bounds__(TR2_THREADS, 8) fdct_quant_kernel2(const __grid_constant__ Geom g, const __grid_constant__ FramePtrs p,
                                                                      const __grid_constant__ Step st)
{
    __shared__ double s_tile[TR2_THREADS / 8][72];
    __shared__ __align__(16) int16_t s_lv[TR2_THREADS / 8][72];   // 144-byte rows: the 4 groups of a warp land on different banks
    const int grp = threadIdx.x >> 3, r = threadIdx.x & 7;
    const int mbi = blockIdx.x * (TR2_THREADS / 8) + grp;
    const bool valid = mbi < g.nmb;
    int mbx, mby;
    const int mb = mb_border_first(g, valid ? mbi : 0, mbx, mby);
    const int gop = blockIdx.y;
    const size_t f = (size_t)gop * st.gop_len + st.t;
    const uint8_t* curf = p.cur + f * g.fb;
    const uint8_t* prevf = p.rec + (f - (INTRA ? 0 : 1)) * g.fb;
    const Mags2 M = load_mags2<0>();
    MbWalk a;
    {
        int mx = 0, my = 0;
        if (!INTRA) {
            const int mvw = *(const int*)(p.mv + (f * g.nmb + mb) * 2);
            mx = (int)(int16_t)(mvw & 0xffff); my = mvw >> 16;
        }
        a = mb_walk(g, mbx, mby, r, mx, my, INTRA);
    }
    // shared-memory addresses of the zig-zag positions of column r: element v of the column goes to IZ[v*8 + r]
    unsigned zz[8];
    {
        const uint2 izc = __ldg(&g_izcol[r]);
        const unsigned base = (unsigned)__cvta_generic_to_shared(&s_lv[grp][0]);
#pragma unroll
        for (int v = 0; v < 8; v++) zz[v] = base + 2u * (((v < 4 ? izc.x : izc.y) >> (8 * (v & 3))) & 255u);
    }
    const unsigned row_addr = (unsigned)__cvta_generic_to_shared(&s_lv[grp][8 * r]);
    const double mu = r == 0 ? M.m[0] : 1.0;     // D[i][0] *= irt2 (ENC:2732-2736); * 1.0 is exact for the other columns
    constexpr int K0 = INTRA ? 4 : 0;
    // outputs, 32-bit element offsets from frame / GOP bases
    int16_t* const lvf = p.levels + f * g.nmb * 384;
    uint8_t* const acf = p.acflag + f * g.nmb * 6;
    double* const dcf = p.dcraw + (size_t)gop * g.nmb * 6;
    const int mb6 = mb * 6;
    uint2 cw = __ldg((const uint2*)(curf + a.off));
    PredRaw prq{0u, 0u, 0u, 0};
    if (!INTRA) prq = mb_walk_pred(g, a, prevf, K0, r);
    // one block; LUMA is a compile-time flag so that the luma and chroma loops carry no code of each other (one body with
    // run-time selects executes both quantiser roundings / both reconstruction rules; six unrolled bodies miss the
    // instruction cache)
    auto body = [&](auto luma_tag, int k) {
        constexpr bool luma = decltype(luma_tag)::value;
        // ---- residual and the butterfly inputs of stage 1, two pixels per instruction ---------------------------
        // halves: [c0 | c1], [c3 | c2], [c7 | c6], [c4 | c5]; the biases keep every half positive (no borrow between halves)
        const uint2 pr = pred_pixels(prq);
        const uint32_t EA = __byte_perm(cw.x, 0, 0x4140) - __byte_perm(pr.x, 0, 0x4140) + 0x02000200u;   // e0, e1  (+512)
        const uint32_t EC = __byte_perm(cw.x, 0, 0x4243) - __byte_perm(pr.x, 0, 0x4243) + 0x02000200u;   // e3, e2
        const uint32_t EB = __byte_perm(cw.y, 0, 0x4243) - __byte_perm(pr.y, 0, 0x4243) + 0x02000200u;   // e7, e6
        const uint32_t ED = __byte_perm(cw.y, 0, 0x4140) - __byte_perm(pr.y, 0, 0x4140) + 0x02000200u;   // e4, e5
        if (k + 1 < 6) {                                   // next block's pixels are in flight during this block's math
            mb_walk_next(g, a, k);
            cw = __ldg((const uint2*)(curf + a.off));
            if (!INTRA) prq = mb_walk_pred(g, a, prevf, k + 1, r);
        }
        const uint32_t S01 = EA + EB, S32 = EC + ED;                               // s0, s1 | s3, s2   (+1024)
        const uint32_t D01 = EA - EB + 0x04000400u, D32 = EC - ED + 0x04000400u;   // d0, d1 | d3, d2   (+1024)
        const uint32_t PP = S01 - S32 + 0x08000800u;                               // s0-s3, s1-s2      (+2048)
        const uint32_t QQ = S01 + S32;                                             // s0+s3, s1+s2      (+2048)
        const double d0 = biased_to_double(D01 & 0xffffu, 1024), d1 = biased_to_double(D01 >> 16, 1024);
        const double d3 = biased_to_double(D32 & 0xffffu, 1024), d2 = biased_to_double(D32 >> 16, 1024);
        const double p03 = biased_to_double(PP & 0xffffu, 2048), p12 = biased_to_double(PP >> 16, 2048);
        const double sum4 = biased_to_double((QQ & 0xffffu) + (QQ >> 16), 4096);
        const double dif4 = biased_to_double((QQ & 0xffffu) - (QQ >> 16) + 4096u, 4096);
        // ---- stage 1 (ENC:2709-2718): every product and partial sum is exact (see fdct_row in icsp_device.cuh) ----
        double t[8];
        {
            const double ca = M.m[1], cb = M.m[3], cc = M.m[5], cd = M.m[7], cE = M.m[2], cF = M.m[6], cG = M.m[4];
            t[0] = sum4;
            t[4] = __dmul_rn(cG, dif4);
            t[2] = __fma_rn(cF, p12, __dmul_rn(cE, p03));
            t[6] = __fma_rn(-cE, p12, __dmul_rn(cF, p03));
            t[1] = __fma_rn(cd, d3, __fma_rn(cc, d2, __fma_rn(cb, d1, __dmul_rn(ca, d0))));
            t[3] = __fma_rn(-cc, d3, __fma_rn(-ca, d2, __fma_rn(-cd, d1, __dmul_rn(cb, d0))));
            t[5] = __fma_rn(cb, d3, __fma_rn(cd, d2, __fma_rn(-ca, d1, __dmul_rn(cc, d0))));
            t[7] = __fma_rn(-ca, d3, __fma_rn(cb, d2, __fma_rn(-cc, d1, __dmul_rn(cd, d0))));
        }
        group_transpose(t, s_tile[grp], r);   // lane r now holds column u = r: t[y][r]
        // ---- stage 2 (ENC:2720-2744) with the 0.25 folded into the constants, quantiser, zig-zag scatter -------
        int nz = 0;
#pragma unroll
        for (int v = 0; v < 8; v++) {
            double acc;
            if (v == 0) {
                acc = t[0];
#pragma unroll
                for (int y = 1; y < 8; y++) acc = __dadd_rn(acc, t[y]);      // T[0][y] == 1.0
                acc = __dmul_rn(acc, M.q[0]);                                // * irt2 * 0.25
            } else {
                acc = __dmul_rn(t[0], ICSP_TQ(M, v, 0));
#pragma unroll
                for (int y = 1; y < 8; y++) acc = __dadd_rn(acc, __dmul_rn(t[y], ICSP_TQ(M, v, y)));
            }
            const double D = __dmul_rn(acc, mu);
            if (v == 0 && r == 0 && valid) dcf[mb6 + k] = D;                 // the DC goes to the chain kernel as a double
            if (TAP && valid) p.dct_tap[((f * g.nmb + mb) * 6 + k) * 64 + v * 8 + r] = D;
            const double x = __dadd_rn(D, 0.5);
            int q;
            if (luma) q = __double2int_rz(x); else q = floor_to_int(x);      // luma truncates (ENC:2780), chroma floors (ENC:4642)
            const int L = QFAST ? div_m19(q, st.m19_ac) : div_magic(q, st.magic_ac);
            sts_u16(zz[v], L);                         // the DC slot (v == 0, r == 0) is rewritten by the DC chain kernel
            if (v == 0) nz |= r == 0 ? 0 : L; else nz |= L;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, nz != 0);
        __syncwarp();
        if (valid) {
            uint4 row;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(row.x), "=r"(row.y), "=r"(row.z), "=r"(row.w) : "r"(row_addr) : "memory");
            *(uint4*)(lvf + ((mb6 + k) * 64 + 8 * r)) = row;
            if (r == 0) acf[mb6 + k] = ((bal >> (8 * (grp & 3))) & 0xffu) ? 0 : 1;
        }
        // s_lv is rewritten by the next block only after the __syncwarp()s of its transposition
    };
    if (!INTRA) {
        ICSP_TR_PRAGMA
        for (int k = 0; k < 4; k++) body(std::true_type{}, k);
    }
    ICSP_TR_PRAGMA
    for (int k = 4; k < 6; k++) body(std::false_type{}, k);
}

// =====================================================================================================
// Kernel C2: dequantisation + IDCT + reconstruction (R9, R10, R13).  TAB 0: encoder table (float widened), TAB 1: decoder
// table (binary64).  Luma and intra chroma: clip(pred + (int)idct) (mergeBlock ENC:4812 + interYReconstruct ENC:2343-2346;
// for intra chroma pred == 0 and intraImgReconstruct ENC:1964-1971 gives the same value); inter chroma: clip((int)(pred +
// idct)) with the sum formed in double (interCbCrReconstruct ENC:2605-2607).
// =====================================================================================================
template <int TAB, bool INTRA>
__global__ void __launch_bounds__(TR2_THREADS, 8) idct_recon_kernel2(const __grid_constant__ Geom g, const __grid_constant__ FramePtrs p,
                                                                      const __grid_constant__ Step st)
{
    __shared__ double s_tile[TR2_THREADS / 8][72];
    __shared__ __align__(16) int16_t s_lv[TR2_THREADS / 8][72];
    __shared__ __align__(8) uint8_t s_px[TR2_THREADS / 8][72];
    const int grp = threadIdx.x >> 3, r = threadIdx.x & 7;
    const int mbi = blockIdx.x * (TR2_THREADS / 8) + grp;
    const bool valid = mbi < g.nmb;
    int mbx, mby;
    const int mb = mb_border_first(g, valid ? mbi : 0, mbx, mby);
    const int gop = blockIdx.y;
    const size_t f = (size_t)gop * st.gop_len + st.t;
    const uint8_t* prevf = p.rec + (f - (INTRA ? 0 : 1)) * g.fb;
    uint8_t* recf = p.rec + f * g.fb;
    const Mags2 M = load_mags2<TAB>();
    MbWalk a;
    {
        int mx = 0, my = 0;
        if (!INTRA) {
            const int mvw = *(const int*)(p.mv + (f * g.nmb + mb) * 2);
            mx = (int)(int16_t)(mvw & 0xffff); my = mvw >> 16;
        }
        a = mb_walk(g, mbx, mby, r, mx, my, INTRA);
    }
    unsigned zz[8];      // shared-memory addresses of row r's coefficients: element u of the row sits at IZ[r*8 + u]
    {
        const uint2 izr = __ldg(&g_izrow[r]);
        const unsigned base = (unsigned)__cvta_generic_to_shared(&s_lv[grp][0]);
#pragma unroll
        for (int u = 0; u < 8; u++) zz[u] = base + 2u * (((u < 4 ? izr.x : izr.y) >> (8 * (u & 3))) & 255u);
    }
    const unsigned lv_row = (unsigned)__cvta_generic_to_shared(&s_lv[grp][8 * r]);
    const unsigned px_row = (unsigned)__cvta_generic_to_shared(&s_px[grp][8 * r]);
    const unsigned px_col = (unsigned)__cvta_generic_to_shared(&s_px[grp][r]);
    constexpr int K0 = INTRA ? 4 : 0;
    const int16_t* const lvf = p.levels + f * g.nmb * 384;
    const int32_t* const dcf = p.dcrec + (size_t)gop * g.nmb * 6;
    const int mb6 = mb * 6;
    uint4 lv = __ldg((const uint4*)(lvf + ((mb6 + K0) * 64 + 8 * r)));
    PredRaw prq{0u, 0u, 0u, 0};
    if (!INTRA) prq = mb_walk_pred(g, a, prevf, K0, r);
    int dc = 0;
    if (r == 0) dc = dcf[mb6 + K0];                       // level*QstepDC + P from the DC chain
    auto body = [&](auto luma_tag, int k) {
        constexpr bool luma = decltype(luma_tag)::value;
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(lv_row), "r"(lv.x), "r"(lv.y), "r"(lv.z), "r"(lv.w) : "memory");
        const uint2 pr = pred_pixels(prq);
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(px_row), "r"(pr.x), "r"(pr.y) : "memory");
        const int dck = dc;
        const int off_k = a.off;
        __syncwarp();
        if (k + 1 < 6) {                                   // next block's loads are in flight during this block's math
            mb_walk_next(g, a, k);
            lv = __ldg((const uint4*)(lvf + ((mb6 + k + 1) * 64 + 8 * r)));
            if (!INTRA) prq = mb_walk_pred(g, a, prevf, k + 1, r);
            if (r == 0) dc = dcf[mb6 + k + 1];
        }
        if (k + 2 < 6) asm volatile("prefetch.global.L2 [%0];" ::"l"(lvf + ((mb6 + k + 2) * 64 + 8 * r)));
        double qd[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            int q = lds_s16(zz[u]) * st.qac;                // IQuantization_block
            if (u == 0 && r == 0) q = dck;
            qd[u] = int_to_double(q);
        }
        // ---- stage 1 (ENC:2858-2867) for row y = r ------------------------------------------------------------
        double t[8];
        {
            const double a0 = __dmul_rn(M.m[0], qd[0]);
#pragma unroll
            for (int x = 0; x < 8; x++) {
                double s = a0;
#pragma unroll
                for (int u = 1; u < 8; u++) {
                    if (TAB == 0) s = __fma_rn(qd[u], ICSP_T(M, u, x), s);        // exact product: fused == unfused
                    else s = __dadd_rn(s, __dmul_rn(qd[u], ICSP_T(M, u, x)));
                }
                t[x] = s;
            }
        }
        group_transpose(t, s_tile[grp], r);   // lane r holds column x = r: t[v][r]
        // ---- stage 2 (ENC:2869-2891) with the 0.25 folded into the constants; 21 products feed the 56 adds --------
        double lo[4], hi[4];
        {
            const double a0 = __dmul_rn(M.q[0], t[0]);
#pragma unroll
            for (int y = 0; y < 4; y++) lo[y] = hi[y] = a0;
#pragma unroll
            for (int v = 1; v < 8; v++) {
                double pm[8];
                bool have[8] = {false, false, false, false, false, false, false, false};
#pragma unroll
                for (int y = 0; y < 4; y++) {
                    const int kk = kTI(v, y);
                    if (!have[kk]) { pm[kk] = __dmul_rn(t[v], M.q[kk]); have[kk] = true; }
                    const bool neg = kTS(v, y) < 0;
                    lo[y] = neg ? __dsub_rn(lo[y], pm[kk]) : __dadd_rn(lo[y], pm[kk]);
                    hi[y] = (neg != ((v & 1) != 0)) ? __dsub_rn(hi[y], pm[kk]) : __dadd_rn(hi[y], pm[kk]);
                }
            }
        }
        // ---- reconstruction of column r -------------------------------------------------------------------------
        constexpr bool sum_first = !INTRA && !luma;
#pragma unroll
        for (int y = 0; y < 8; y++) {
            const double R = y < 4 ? lo[y] : hi[7 - y];
            unsigned pv;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(pv) : "r"(px_col + 8u * y) : "memory");
            int o;
            if (!sum_first) o = (int)pv + __double2int_rz(R);
            else o = __double2int_rz(__dadd_rn(biased_to_double(pv, 0), R));
            o = min(255, max(0, o));
            asm volatile("st.shared.u8 [%0], %1;" ::"r"(px_col + 8u * y), "r"(o) : "memory");
        }
        __syncwarp();
        if (valid) {
            uint2 o;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(o.x), "=r"(o.y) : "r"(px_row) : "memory");
            *(uint2*)(recf + off_k) = o;
        }
        // s_lv / s_px rows are rewritten by their own lane; the column reads of the next block follow its __syncwarp()
    };
    if (!INTRA) {
        ICSP_TR_PRAGMA
        for (int k = 0; k < 4; k++) body(std::true_type{}, k);
    }
    ICSP_TR_PRAGMA
    for (int k = 4; k < 6; k++) body(std::false_type{}, k);
}

}  // namespace icsp
