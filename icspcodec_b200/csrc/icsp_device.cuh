// Device-side building blocks shared by every kernel of libicspcuda (sm_100a).
//
// Numerical contract: IEEE binary64, one rounding per multiply and per add
// exactly where the reference (compiled without FMA) rounds.  Every FP64 operation below is an explicit
// __dmul_rn/__dadd_rn/__dsub_rn (never contracted by the compiler); __fma_rn appears only where the
// product is exactly representable (int x float-widened constant), so fused == unfused bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace icsp {

struct Geom {
    int w, h;      // luma size
    int mbw, mbh;  // macroblock grid
    int nmb;
    int cw, ch;    // chroma size
    int bw, bh;    // luma 8x8 grid
    int fb;        // bytes per I420 frame
    unsigned magic_bw, magic_mbw;  // ceil(2^32/bw), ceil(2^32/mbw): n/d == umulhi(n, magic) while n*d < 2^32
    unsigned magic_mbw2;           // ceil(2^32/(mbw-2)) (interior macroblocks per row), 0 if mbw <= 2
    // block k of a macroblock (Y0..Y3, Cb, Cr) relative to block 0 of its plane group (luma: Y0, chroma: Cb):
    int d_pix[6];                  // byte offset inside a frame: luma (k>>1)*8*w + (k&1)*8; Cb 0, Cr cw*ch
    int d_dc[6];                   // index into the plane-raster DC maps: luma (k>>1)*bw + (k&1); Cb 4*nmb, Cr 5*nmb
};

__constant__ signed char c_cand[8][64][2];  // spiral visiting order per carried start state: (dx,dy) (ENC:2101-2143)
__constant__ unsigned char c_next[8][65];   // start state, moves made -> state handed to the next macroblock

// compile-time copies of the zig-zag tables (indices are compile-time constants after full unrolling)
__host__ __device__ constexpr int kZZ(int k)
{
    constexpr unsigned char t[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                     41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                     30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    return t[k];
}
__host__ __device__ constexpr int kIZ(int raster)
{
    for (int k = 0; k < 64; k++)
        if (kZZ(k) == raster) return k;
    return 0;
}

// x / q (C semantics, truncation toward zero), branch free: magic = ceil(2^31/q), result = umulhi(2|x|, magic).
// Exact while |x|*q < 2^31; transform coefficients of 8-bit video are bounded by |x| <= 4080+4080 and q <= 255.
__device__ __forceinline__ int div_magic(int x, unsigned magic)
{
    const unsigned r = __umulhi((unsigned)abs(x) << 1, magic);
    return x < 0 ? -(int)r : (int)r;
}

__device__ __forceinline__ int med3(int a, int b, int c)
{  // ENC:3677-3679
    if (a > b && a > c) return b > c ? b : c;
    if (b > a && b > c) return a > c ? a : c;
    return a > b ? a : b;
}
__device__ __forceinline__ int clip255(int v) { return min(255, max(0, v)); }

// R1 (getPaddingImage, ENC:2227-2269): padded(y,x) = src(clamp) except the LAST padded row/column, which stay 0.
__device__ __forceinline__ int ref_px(const uint8_t* __restrict__ P, int w, int h, int pad, int y, int x)
{
    if (y == h + 2 * pad - 1 || x == w + 2 * pad - 1) return 0;
    int yy = min(max(y - pad, 0), h - 1), xx = min(max(x - pad, 0), w - 1);
    return P[yy * w + xx];
}

// 8 consecutive prediction pixels of one row, starting at padded coordinate (y, x0), packed as two little-endian
// words (byte i of the pair = pixel x0+i)
__device__ __forceinline__ uint2 ref_row8_packed(const uint8_t* __restrict__ P, int w, int h, int pad, int y, int x0)
{
    const int ux = x0 - pad, uy = y - pad;
    if (uy >= 0 && uy < h && ux >= 0 && ux + 7 < w) {
        const uint8_t* p = P + uy * w + ux;
        const uintptr_t a = (uintptr_t)p;
        const uint32_t* q = (const uint32_t*)(a & ~(uintptr_t)3);
        const int sh = (int)(a & 3) * 8;
        const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = sh ? __ldg(q + 2) : 0u;
        return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
    }
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        lo |= (uint32_t)ref_px(P, w, h, pad, y, x0 + i) << (8 * i);
        hi |= (uint32_t)ref_px(P, w, h, pad, y, x0 + 4 + i) << (8 * i);
    }
    return make_uint2(lo, hi);
}

// ---- A.5 DC predictors on the 8x8 grid of one plane (dc = reconstructed DCs, row-major [bh][bw]) ----
// S = element stride of the map in ints (2 when the DC shares an 8-byte slot with the level, see dc_chain_kernel)
template <int S = 1>
__device__ __forceinline__ int dc_pred_luma(const int* dc, int bw, int bx, int by)
{  // ENC:3643-3990 reduced to geometry
    if (bx == 0 && by == 0) return 1024;
    if (by == 0) return dc[S * (bx - 1)];
    if (bx == 0) return dc[S * ((by - 1) * bw)];
    const int L = dc[S * (by * bw + bx - 1)], U = dc[S * ((by - 1) * bw + bx)];
    if ((bx & 1) == 0 || ((by & 1) == 0 && bx != bw - 1)) return med3(L, U, dc[S * ((by - 1) * bw + bx + 1)]);
    return med3(L, dc[S * ((by - 1) * bw + bx - 1)], U);
}
template <int S = 1>
__device__ __forceinline__ int dc_pred_chroma(const int* dc, int bw, int bx, int by)
{  // ENC:4482-4513
    if (bx == 0 && by == 0) return 1024;
    if (by == 0) return dc[S * (bx - 1)];
    if (bx == 0) return dc[S * ((by - 1) * bw)];
    const int L = dc[S * (by * bw + bx - 1)], U = dc[S * ((by - 1) * bw + bx)];
    if (bx == bw - 1) return med3(L, dc[S * ((by - 1) * bw + bx - 1)], U);
    return med3(L, U, dc[S * ((by - 1) * bw + bx + 1)]);
}

// ---- quantiser (A.4) ----------------------------------------------------------------------------
// luma ENC:2780  (int)(D+0.5)/Q ; chroma ENC:4642  (int)floor(D+0.5)/Q ; both divisions truncate toward zero
__device__ __forceinline__ int quant_magic(double D, unsigned magic, bool chroma)
{
    const double x = __dadd_rn(D, 0.5);
    const int r = chroma ? __double2int_rd(x) : __double2int_rz(x);
    return div_magic(r, magic);
}

// ---- 8x8 transforms, one 8-lane group per block --------------------------------------------------
// costable[u][x] = kTS(u,x) * LIT[kTI(u,x)] with LIT[k] ~ cos(k*pi/16) (ENC.h:191-198): index and sign are compile-time,
// the seven magnitudes (and irt2) live in REGISTERS (Mags, loaded once per thread through the read-only path) instead
// of being re-fetched from the constant bank into uniform registers before every FP64 instruction.
__host__ __device__ constexpr int kTI(int u, int x)
{
    int k = ((2 * x + 1) * u) % 32;
    if (k > 16) k = 32 - k;
    if (k > 8) k = 16 - k;
    return k;
}
__host__ __device__ constexpr int kTS(int u, int x)
{
    int k = ((2 * x + 1) * u) % 32;
    if (k > 16) k = 32 - k;
    return k > 8 ? -1 : 1;
}
__device__ double g_mag[2][8];   // [table][k]: |costable| magnitudes, table 0 = float widened, 1 = binary64; [t][0] = irt2
struct Mags { double m[8]; };    // m[0] = irt2, m[1..7] = magnitudes a(1) e(2) b(3) g(4) c(5) f(6) d(7)
template <int TAB>
__device__ __forceinline__ Mags load_mags()
{
    Mags M;
#pragma unroll
    for (int i = 0; i < 8; i++) M.m[i] = __ldg(&g_mag[TAB][i]);
    return M;
}
#define ICSP_T(M, u, x) (kTS(u, x) > 0 ? (M).m[kTI(u, x)] : -(M).m[kTI(u, x)])

// Stage 1 of the forward DCT (ENC:2709-2718) for one row: t[u] = sum_x E[x]*T[u][x].
// Every term and every partial sum is exactly representable (|E|<=255, 24-bit constants, < 38 bits), so
// the even/odd factorisation below returns the reference's value bit for bit.
__device__ __forceinline__ void fdct_row(const int e[8], double t[8], const Mags& M)
{
    const double a = M.m[1], b = M.m[3], c = M.m[5], d = M.m[7], E = M.m[2], F = M.m[6], G = M.m[4];
    const int s0 = e[0] + e[7], s1 = e[1] + e[6], s2 = e[2] + e[5], s3 = e[3] + e[4];
    const double d0 = (double)(e[0] - e[7]), d1 = (double)(e[1] - e[6]), d2 = (double)(e[2] - e[5]), d3 = (double)(e[3] - e[4]);
    const double p03 = (double)(s0 - s3), p12 = (double)(s1 - s2);
    t[0] = (double)(s0 + s1 + s2 + s3);
    t[4] = __dmul_rn(G, (double)(s0 - s1 - s2 + s3));
    t[2] = __fma_rn(F, p12, __dmul_rn(E, p03));
    t[6] = __fma_rn(-E, p12, __dmul_rn(F, p03));
    t[1] = __fma_rn(d, d3, __fma_rn(c, d2, __fma_rn(b, d1, __dmul_rn(a, d0))));
    t[3] = __fma_rn(-c, d3, __fma_rn(-a, d2, __fma_rn(-d, d1, __dmul_rn(b, d0))));
    t[5] = __fma_rn(b, d3, __fma_rn(d, d2, __fma_rn(-a, d1, __dmul_rn(c, d0))));
    t[7] = __fma_rn(-a, d3, __fma_rn(b, d2, __fma_rn(-c, d1, __dmul_rn(d, d0))));
}
// Stage 2 (ENC:2720-2729) for one column u: D[v] = sum_y fl(t[y]*T[v][y]), y ascending, rounded per op;
// then the irt2 / 0.25 scaling of ENC:2732-2744 (element [0][0] is scaled by irt2 twice, in sequence).
__device__ __forceinline__ void fdct_col(const double t[8], int u, double D[8], const Mags& M)
{
    double s = t[0];
#pragma unroll
    for (int y = 1; y < 8; y++) s = __dadd_rn(s, t[y]);  // T[0][y] == 1.0
    D[0] = s;
#pragma unroll
    for (int v = 1; v < 8; v++) {
        double acc = __dmul_rn(t[0], ICSP_T(M, v, 0));
#pragma unroll
        for (int y = 1; y < 8; y++) acc = __dadd_rn(acc, __dmul_rn(t[y], ICSP_T(M, v, y)));
        D[v] = acc;
    }
    D[0] = __dmul_rn(D[0], M.m[0]);
    if (u == 0) {
#pragma unroll
        for (int v = 0; v < 8; v++) D[v] = __dmul_rn(D[v], M.m[0]);
    }
#pragma unroll
    for (int v = 0; v < 8; v++) D[v] = __dmul_rn(D[v], 0.25);
}

// Stage 1 of the IDCT (ENC:2858-2867) for one row y: t[x] = sum_u (C[u]*Q[u])*T[u][x].
// TAB==0: Q[u]*T[u][x] is exact (int x float-widened constant) so an FMA equals mul-then-add;
// TAB==1 (decoder, binary64 table): the product is inexact and must be rounded separately.
template <int TAB>
__device__ __forceinline__ void idct_row(const int q[8], double t[8], const Mags& M)
{
    const double a0 = __dmul_rn(M.m[0], (double)q[0]);
    double qd[8];
#pragma unroll
    for (int u = 1; u < 8; u++) qd[u] = (double)q[u];
#pragma unroll
    for (int x = 0; x < 8; x++) {
        double s = a0;
#pragma unroll
        for (int u = 1; u < 8; u++) {
            if (TAB == 0) s = __fma_rn(qd[u], ICSP_T(M, u, x), s);
            else s = __dadd_rn(s, __dmul_rn(qd[u], ICSP_T(M, u, x)));
        }
        t[x] = s;
    }
}
// Stage 2 (ENC:2869-2878) for one column x: R[y] = sum_v fl((C[v]*t[v])*T[v][y]), then *0.25 (ENC:2885-2891).
// T[v][7-y] == (-1)^v T[v][y] exactly, so each rounded product serves two outputs; rows 2, 4 and 6 of the table hold
// only 2, 1 and 2 distinct magnitudes in their first four entries, so 21 multiplications (not 28) feed the 56 adds.
template <int TAB>
__device__ __forceinline__ void idct_col(const double t[8], double R[8], const Mags& M)
{
    const double a0 = __dmul_rn(M.m[0], t[0]);
    double lo[4], hi[4];
#pragma unroll
    for (int y = 0; y < 4; y++) lo[y] = hi[y] = a0;
#pragma unroll
    for (int v = 1; v < 8; v++) {
        double pm[8];   // product by magnitude index, computed on first use (all indices are compile-time)
        bool have[8] = {false, false, false, false, false, false, false, false};
#pragma unroll
        for (int y = 0; y < 4; y++) {
            const int k = kTI(v, y);
            if (!have[k]) { pm[k] = __dmul_rn(t[v], M.m[k]); have[k] = true; }
            const bool neg = kTS(v, y) < 0;       // fl(a * -m) == -fl(a * m)
            lo[y] = neg ? __dsub_rn(lo[y], pm[k]) : __dadd_rn(lo[y], pm[k]);
            hi[y] = (neg != ((v & 1) != 0)) ? __dsub_rn(hi[y], pm[k]) : __dadd_rn(hi[y], pm[k]);
        }
    }
#pragma unroll
    for (int y = 0; y < 4; y++) {
        R[y] = __dmul_rn(lo[y], 0.25);
        R[7 - y] = __dmul_rn(hi[y], 0.25);
    }
}

// 8x8 transposition inside one 8-lane group through a padded shared tile (pitch 9).
template <typename T>
__device__ __forceinline__ void group_transpose(T v[8], T* tile, int r)
{
#pragma unroll
    for (int k = 0; k < 8; k++) tile[r * 9 + k] = v[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = tile[k * 9 + r];
    __syncwarp();
}

// sum over the 8 lanes of a group (groups are aligned to 8 lanes inside a warp)
__device__ __forceinline__ int group_sum(int v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

}  // namespace icsp
