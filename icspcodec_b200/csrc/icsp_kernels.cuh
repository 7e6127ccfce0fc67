// Kernels of libicspcuda (sm_100a).  One launch processes the same intra-GOP frame index t of EVERY GOP in the
// batch (blockIdx.y = GOP), because closed GOPs share no state (README.md:152) while frames inside a GOP are
// strictly sequential.  Frame index of GOP g at step t: f = g*gop_len + t.
#pragma once
#include "icsp_device.cuh"

namespace icsp {

struct FramePtrs {
    const uint8_t* cur;  // [F][fb] input frames
    uint8_t* rec;        // [F][fb] reconstruction
    int16_t* levels;     // [F][nmb][6][64]
    uint8_t* acflag;     // [F][nmb][6]
    uint8_t* mpm;        // [F][nmb][4]
    uint8_t* ipm;        // [F][nmb][4]
    int16_t* mvd;        // [F][nmb][2]
    int16_t* mv;         // [F][nmb][2]
    int32_t* minsad;     // [F][nmb]
    double* dcraw;       // [G][nmb][6] scaled forward-DCT DC, macroblock major like the levels (Y0..Y3, Cb, Cr)
    int32_t* dcrec;      // [G][nmb][6] reconstructed DC, same order
    uint8_t* mestate;    // [G][nmb]  spiral start state per macroblock
    uint8_t* memoves;    // [G][nmb]  moves made by the search (64 = no early break)
    uint32_t* meflag;    // [G]       number of macroblocks of the frame whose search broke early
    unsigned long long* mezero;  // [G][nmb][8] zero-SAD masks per start state (exact fallback only)
    double* dct_tap;     // optional debug tap [F][nmb][6][64]: forward DCT output, raster order, before the DC predictor (NULL = off)
};

struct Step {
    int gop_len, t;  // f = g*gop_len + t
    int qdc, qac;
    int intra;       // 1: intra frame (transform kernels touch chroma only; luma is the wavefront kernel)
    unsigned magic_ac, magic_dc;  // ceil(2^31/q), see div_magic
    int qmode_ac, m19_ac;         // qmode_ac == 0: AC levels by div_m19 with m19_ac = floor(2^19/qac)+1 (qac <= 128); 1: div_magic
};

// item -> (mb, k, plane geometry)
struct BlockId {
    int mb, k, plane;  // plane 0 Y, 1 Cb, 2 Cr
    int bx, by;        // position in the plane's 8x8 grid
    int pw, ph;        // plane size
    int poff;          // plane offset inside a frame
    int dcidx;         // index into dcraw/dcrec (macroblock major)
};
// luma launches: item = mb*4 + k ; chroma launches: item = mb*2 + (k-4)
template <bool CHROMA>
__device__ __forceinline__ BlockId block_id(const Geom& g, int item)
{
    BlockId b;
    if (CHROMA) { b.mb = item >> 1; b.k = 4 + (item & 1); }
    else { b.mb = item >> 2; b.k = item & 3; }
    const int mby = (int)__umulhi((unsigned)b.mb, g.magic_mbw), mbx = b.mb - mby * g.mbw;
    if (!CHROMA) {
        b.plane = 0; b.bx = 2 * mbx + (b.k & 1); b.by = 2 * mby + (b.k >> 1);
        b.pw = g.w; b.ph = g.h; b.poff = 0; b.dcidx = b.mb * 6 + b.k;
    } else {
        b.plane = b.k - 3; b.bx = mbx; b.by = mby; b.pw = g.cw; b.ph = g.ch;
        b.poff = g.w * g.h + (b.k - 4) * g.cw * g.ch;
        b.dcidx = b.mb * 6 + b.k;
    }
    return b;
}

// prediction row r of a block, 8 pixels packed in two words (inter: motion compensated from the previous
// reconstruction; intra chroma: 0)
template <bool CHROMA>
__device__ __forceinline__ uint2 pred_row_packed(const Geom& g, const BlockId& b, const uint8_t* prevf, const int16_t* mvf, int r, int intra)
{
    if (intra) return make_uint2(0u, 0u);
    const int mvw = *(const int*)(mvf + 2 * b.mb);
    const int mx = (int)(int16_t)(mvw & 0xffff), my = mvw >> 16;
    // luma: motionCompensation ENC:2185-2186, pad 16; chroma: CmotionCompensation ENC:2538-2539, mv/2 toward zero, pad 8
    if (!CHROMA) return ref_row8_packed(prevf, g.w, g.h, 16, 16 + b.by * 8 + r - my, 16 + b.bx * 8 - mx);
    return ref_row8_packed(prevf + b.poff, g.cw, g.ch, 8, 8 + b.by * 8 + r - my / 2, 8 + b.bx * 8 - mx / 2);
}

// =====================================================================================================
// Kernel A: residual + forward DCT + AC quantisation + zig-zag (R3, R5, R7, R8).  8 lanes per 8x8 block.
// Writes AC levels (DC slot is filled by the DC chain kernel), ACflag and the scaled DC as a double.
// =====================================================================================================
constexpr int TR_THREADS = 128;  // 16 blocks per CTA
// zig-zag positions needed by lane r of a block group, packed one byte per entry: g_izcol[r] byte v = IZ[v*8+r]
// (column r, used after the forward transform), g_izrow[r] byte u = IZ[r*8+u] (row r, used before the inverse).
// Global memory + read-only path: a lane-indexed __constant__ table would serialise 8 ways.
__device__ uint2 g_izcol[8], g_izrow[8];
__device__ __forceinline__ int iz_byte(const uint2& t, int i) { return (int)(((i < 4 ? t.x : t.y) >> (8 * (i & 3))) & 255u); }
// One 8-lane group owns a whole macroblock: its blocks k0..k1-1 (inter frames: Y0..Y3, Cb, Cr; intra frames: Cb, Cr
// only) are processed one after the other, so the macroblock-level work (coordinates, motion vector, base addresses)
// is paid once, and the pixel loads of block k+1 are issued before the arithmetic of block k.  All groups of a warp are
// at the same k, so the luma/chroma differences (plane geometry, pad, mv/2, quantiser rounding) are warp-uniform.
struct MbBlock { int poff, bx, by, pw, ph, pad, mx, my, dcidx; };
__device__ __forceinline__ MbBlock mb_block(const Geom& g, int mb, int mbx, int mby, int k, int mx, int my)
{
    MbBlock b;
    if (k < 4) {
        b.poff = 0; b.bx = 2 * mbx + (k & 1); b.by = 2 * mby + (k >> 1); b.pw = g.w; b.ph = g.h; b.pad = 16;
        b.mx = mx; b.my = my;                               // motionCompensation ENC:2185-2186
        b.dcidx = mb * 6 + k;
    } else {
        b.poff = g.w * g.h + (k - 4) * g.cw * g.ch; b.bx = mbx; b.by = mby; b.pw = g.cw; b.ph = g.ch; b.pad = 8;
        b.mx = mx / 2; b.my = my / 2;                       // CmotionCompensation ENC:2538-2539: truncation toward zero
        b.dcidx = mb * 6 + k;
    }
    return b;
}


__global__ void __launch_bounds__(TR_THREADS) fdct_quant_kernel(Geom g, FramePtrs p, Step st)
{
    __shared__ double s_tile[TR_THREADS / 8][72];
    __shared__ __align__(16) int16_t s_lv[TR_THREADS / 8][72];   // 144-byte rows: the 4 groups of a warp land on different banks
    const int grp = threadIdx.x >> 3, r = threadIdx.x & 7;
    const int mbi = blockIdx.x * (TR_THREADS / 8) + grp;
    const bool valid = mbi < g.nmb;
    const int mb = valid ? mbi : 0;
    const int gop = blockIdx.y;
    const size_t f = (size_t)gop * st.gop_len + st.t;
    const int mby = (int)__umulhi((unsigned)mb, g.magic_mbw), mbx = mb - mby * g.mbw;
    const uint8_t* curf = p.cur + f * g.fb;
    const uint8_t* prevf = p.rec + (f - (st.intra ? 0 : 1)) * g.fb;
    const Mags M = load_mags<0>();
    const uint2 izc = __ldg(&g_izcol[r]);
    int mx = 0, my = 0;
    if (!st.intra) {
        const int mvw = *(const int*)(p.mv + (f * g.nmb + mb) * 2);
        mx = (int)(int16_t)(mvw & 0xffff); my = mvw >> 16;
    }
    const int k0 = st.intra ? 4 : 0;
    auto fetch = [&](int k, uint2& cw, uint2& pr) {
        const MbBlock b = mb_block(g, mb, mbx, mby, k, mx, my);
        cw = __ldg((const uint2*)(curf + b.poff + (unsigned)((b.by * 8 + r) * b.pw + b.bx * 8)));
        pr = make_uint2(0u, 0u);
        if (!st.intra) pr = ref_row8_packed(prevf + b.poff, b.pw, b.ph, b.pad, b.pad + b.by * 8 + r - b.my, b.pad + b.bx * 8 - b.mx);
    };
    uint2 cw, pr;
    fetch(k0, cw, pr);
    int16_t* lvmb = p.levels + (f * g.nmb + mb) * 384;
#pragma unroll 1
    for (int k = k0; k < 6; k++) {
        int e[8];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            e[i] = (int)((cw.x >> (8 * i)) & 255u) - (int)((pr.x >> (8 * i)) & 255u);
            e[4 + i] = (int)((cw.y >> (8 * i)) & 255u) - (int)((pr.y >> (8 * i)) & 255u);
        }
        if (k + 1 < 6) fetch(k + 1, cw, pr);               // next block's pixels are in flight during this block's math
        double t[8], D[8];
        fdct_row(e, t, M);
        group_transpose(t, s_tile[grp], r);   // lane r now holds column u=r: t[y][r]
        fdct_col(t, r, D, M);                  // D[v][u=r]
        int nz = 0;
        if (k < 4) {                           // warp-uniform: luma truncates (ENC:2780), chroma floors (ENC:4642)
#pragma unroll
            for (int v = 0; v < 8; v++) {
                const int L = quant_magic(D[v], st.magic_ac, false);   // the DC slot (v == 0, r == 0) is rewritten by the DC chain kernel
                s_lv[grp][iz_byte(izc, v)] = (int16_t)L;
                if (!(v == 0 && r == 0)) nz |= L;
            }
        } else {
#pragma unroll
            for (int v = 0; v < 8; v++) {
                const int L = quant_magic(D[v], st.magic_ac, true);
                s_lv[grp][iz_byte(izc, v)] = (int16_t)L;
                if (!(v == 0 && r == 0)) nz |= L;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, nz != 0);
        __syncwarp();
        if (valid) {
            *(uint4*)(lvmb + k * 64 + 8 * r) = *(const uint4*)(&s_lv[grp][8 * r]);
            if (r == 0) {
                p.acflag[(f * g.nmb + mb) * 6 + k] = ((bal >> (8 * (grp & 3))) & 0xffu) ? 0 : 1;
                p.dcraw[(size_t)gop * 6 * g.nmb + mb_block(g, mb, mbx, mby, k, 0, 0).dcidx] = D[0];
            }
        }
        __syncwarp();                          // s_lv is rewritten by the next block
    }
}

// =====================================================================================================
// Kernel B: DC-DPCM chains (R6) and MV differentials (R4).  One CTA per frame; warp 0 luma, 1 Cb, 2 Cr walk
// their 8x8 grid in waves w = bx + 2*by (the UR dependency gives slope 2); warp 3 computes mvd = mv - Pm.
// decode == 0: level = quant(Draw - P), rec = level*Q + P ;  decode == 1: rec = level*Q + P.
// =====================================================================================================
__device__ __forceinline__ void mv_predictor(const int16_t* mv, int mbw, int mb, int& px, int& py)
{  // mvPrediction ENC:2353-2425 (the y-median typo `y1>x3` is the reference's and is kept)
    const int mbx = mb % mbw;
    if (mb == 0) { px = 8; py = 8; return; }
    if (mb < mbw) { px = mv[2 * (mb - 1)]; py = mv[2 * (mb - 1) + 1]; return; }
    if (mbx == 0) { px = mv[2 * (mb - mbw)]; py = mv[2 * (mb - mbw) + 1]; return; }
    int a = mb - 1, b, c;
    if (mbx == mbw - 1) { b = mb - mbw - 1; c = mb - mbw; } else { b = mb - mbw; c = mb - mbw + 1; }
    const int x1 = mv[2 * a], x2 = mv[2 * b], x3 = mv[2 * c], y1 = mv[2 * a + 1], y2 = mv[2 * b + 1], y3 = mv[2 * c + 1];
    px = med3(x1, x2, x3);
    if (y1 > y2 && y1 > y3) py = y2 > y3 ? y2 : y3;
    else if (y2 > y1 && y2 > y3) py = (y1 > x3) ? y1 : y3;
    else py = y1 > y2 ? y1 : y2;
}

// `staged` != 0: every block owns one 4-byte shared-memory slot that first holds its input and, once the chain has passed,
// {reconstructed DC (low half), DC level (high half)} as two int16 (|raw DC| <= 2040 for 8-bit video, so |rec| <= 2040 + Q and
// |level| <= 4335 fit); the dependent waves then touch shared memory only (9.5 KB for CIF: 16 frames per SM, the thread limit).
// Input of a slot when decoding: the DC level.  When encoding: the raw DC is a double and the reference's chain is
//     L = (int)trunc|floor((raw - P) + 0.5) / Q            (ENC:2780 / ENC:4642 after DPCM_DC_block, P the integer predictor)
// whose roundings move the value by < 2^-38 while |raw| < 2^13, so with xs = raw + 0.5, m = floor(xs), g = xs - m the result is the INTEGER
// expression floor: m - P, trunc: m - P (+1 if m - P < 0) whenever g is not within `amb_eps` (2^-30) of 0 or 1 — the slot holds
// (m << 1) and the dependent path of the chain has no FP64 and no conversion on it.  A block whose g IS that close to an integer
// (probability ~2^-29; forced for every block by ICSP_DC_EPS=1 in the tests), or whose |raw + 0.5| >= 2^13, carries bit 0 and takes the reference's double
// sequence from the raw value in global memory.  All 128 threads stage the inputs in (loads batched four deep) and the results
// out (coalesced over the macroblock-major global arrays, scattered into the plane-raster slots); only the wave walk itself is
// one warp per plane.  Large frames fall back to staged == 0 (int map only, inputs/outputs in global memory inside the chain).
// slot index of block k of macroblock mb: luma plane-raster [bh][bw] at 0 (+1 unused), Cb [nmb] (+1), Cr [nmb] (+1)
__device__ __forceinline__ int chain_slot(const Geom& g, int mb, int k)
{
    if (k >= 4) return 4 * g.nmb + 1 + (k - 4) * (g.nmb + 1) + mb;
    const int mby = (int)__umulhi((unsigned)mb, g.magic_mbw), mbx = mb - mby * g.mbw;
    return (2 * mby + (k >> 1)) * g.bw + 2 * mbx + (k & 1);
}
// plane-raster block index i of plane pl (0 Y, 1 Cb, 2 Cr) -> macroblock-major index mb*6 + k
__device__ __forceinline__ int chain_mbmajor(const Geom& g, int pl, int i)
{
    if (pl) return i * 6 + 3 + pl;
    const int by = (int)__umulhi((unsigned)i, g.magic_bw), bx = i - by * g.bw;
    return ((by >> 1) * g.mbw + (bx >> 1)) * 6 + (((by & 1) << 1) | (bx & 1));
}
// The wave walk of one plane by one warp.  Each lane owns ONE block row and walks it left to right, one block per step, two
// steps behind the lane above (the UR dependency): at step w lane l codes block (w - 2l) of its row.  Its left neighbour is
// its own previous result (a register); UR is the latest result of the lane above — one warp shuffle per step — and U / UL
// are the two values that shuffle returned before.  Rows are taken 32 at a time, then 31: in every pass after the first, lane 0 does
// not code anything, it REPLAYS the last row of the previous pass from the slots so that lane 1 sees its neighbours the
// same way as every other lane.  Boundary rules of A.5 without branches: a lane that is not yet inside its row forwards
// what it receives, so at bx == 0 its "left neighbour" equals U and med3(U, U, c) = U; the top row uses med3(L, L, c) = L
// with L starting at 1024.  The input of the next block is prefetched one step ahead and results are stored fire-and-forget:
// the dependent path of a step is shuffle -> select -> median -> subtract -> integer divide -> multiply-add.
template <bool CHROMA, bool DECODE>
__device__ __forceinline__ void chain_walk(const Geom& g, int* slot, const double* __restrict__ raw, int pl, int bw, int bh, int lane,
                                           unsigned magic, int q)
{
    const int urm = bw - 1;          // chroma: UR unless bx == bw-1; luma odd rows: UR iff bx is even (bw is even, so bw-1 is odd)
    for (int first = 0;;) {                                      // row of lane 0: row 0, or the replayed last row of the pass before
        const int by = first + lane;
        const bool replay = first > 0 && lane == 0;
        const int rows = min(32, bh - first);
        const int bw_l = lane < rows ? bw : 0;                   // lanes without a row are never active
        const bool top = by == 0;
        const int m = (!CHROMA && (by & 1)) ? 1 : urm;           // ur = (bx & m) != m
        const int steps = bw + 2 * (rows - 1);
        int* sp = slot + by * bw - 2 * lane;                     // sp[w] is the slot of the block coded at step w
        int bx = -2 * lane;
        int h1 = 1024, q1 = 0, q2 = 0;
        int in = bx == 0 && bw_l ? sp[0] : 0;
#pragma unroll 2
        for (int w = 0; w < steps; w++, bx++) {
            const int sent = __shfl_up_sync(0xffffffffu, h1, 1);
            const bool act = (unsigned)bx < (unsigned)bw_l;
            const int in_n = (unsigned)(bx + 1) < (unsigned)bw_l ? sp[w + 1] : 0;
            const int c = (bx & m) != m ? sent : q2;
            const int b = top ? h1 : q1;
            const int P = max(min(h1, b), min(max(h1, b), c));   // med3 is a true median (ENC:3677-3679)
            int L = in;
            if (!DECODE) {                                       // DPCM_DC_block: D -= P, then quantise
                int d = (in >> 1) - P;
                if (!CHROMA) d += (int)((unsigned)d >> 31);      // truncation toward zero of the non-integer (m - P) + g
                L = div_magic(d, magic);
                if (in & 1) {                                    // raw DC + 0.5 too close to an integer: the reference's double sequence
                    if (act && !replay) L = quant_magic(__dsub_rn(raw[chain_mbmajor(g, pl, by * bw + bx)], (double)P), magic, CHROMA);
                }
            }
            int res = L * q + P;                                 // IQuantization + IDPCM_DC_block
            if (act && !replay) sp[w] = (int)__byte_perm((unsigned)res, (unsigned)L, 0x5410);
            if (replay) res = (int)(short)in;                    // stored {rec, level}: the sign-extended low half
            h1 = act ? res : sent;
            q2 = q1; q1 = sent;
            in = in_n;
        }
        __syncwarp();                                            // the next pass replays this pass's last row from shared memory
        if (first + rows >= bh) break;
        first += rows - 1;
    }
}

__global__ void __launch_bounds__(128) dc_chain_kernel(const __grid_constant__ Geom g, const __grid_constant__ FramePtrs p,
                                                        const __grid_constant__ Step st, int decode, int staged, float amb_eps)
{
    extern __shared__ __align__(16) unsigned char s_chain[];
    const int nblk = 6 * g.nmb;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gop = blockIdx.x;
    const size_t f = (size_t)gop * st.gop_len + st.t;
    const double* raw = p.dcraw + (size_t)gop * nblk;
    int32_t* rec = p.dcrec + (size_t)gop * nblk;
    int16_t* lvf = p.levels + f * g.nmb * 384;
    const bool chroma = warp > 0;
    const int bw = chroma ? g.mbw : g.bw, bh = chroma ? g.mbh : g.bh;
    const int nwaves = (bw - 1) + 2 * (bh - 1) + 1;
    const bool walker = warp < 3 && !(warp == 0 && st.intra);   // intra luma DCs are chained inside the wavefront kernel
    if (staged) {
        int* slots = (int*)s_chain;
        const double eps = (double)amb_eps;
        for (int i0 = threadIdx.x; i0 < nblk; i0 += 4 * 128) {
            double rv[4];
            int lv[4], sl[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i6 = i0 + u * 128;
                const int mb = (int)__umulhi((unsigned)i6, 0x2aaaaaabu), k = i6 - 6 * mb;    // i6 / 6 for i6 < 2^31
                sl[u] = (i6 < nblk && !(st.intra && k < 4)) ? chain_slot(g, mb, k) : -1;
                rv[u] = 0.0; lv[u] = 0;
                if (sl[u] >= 0) { if (decode) lv[u] = lvf[i6 * 64]; else rv[u] = raw[i6]; }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (sl[u] < 0) continue;
                if (decode) { slots[sl[u]] = lv[u]; continue; }
                const double xs = __dadd_rn(rv[u], 0.5);
                const int m = __double2int_rd(xs);
                const double gfr = __dsub_rn(xs, (double)m);                                  // exact
                const bool amb = !(gfr >= eps && gfr <= 1.0 - eps) || !(fabs(xs) < 8192.0);   // 8-bit video: |raw| <= 2040
                slots[sl[u]] = (int)((unsigned)m << 1) | (amb ? 1 : 0);
            }
        }
        __syncthreads();
        if (walker) {
            const int base = chroma ? 4 * g.nmb + 1 + (warp - 1) * (g.nmb + 1) : 0;
            if (decode && chroma) chain_walk<true, true>(g, slots + base, raw, warp, bw, bh, lane, st.magic_dc, st.qdc);
            else if (decode) chain_walk<false, true>(g, slots + base, raw, warp, bw, bh, lane, st.magic_dc, st.qdc);
            else if (chroma) chain_walk<true, false>(g, slots + base, raw, warp, bw, bh, lane, st.magic_dc, st.qdc);
            else chain_walk<false, false>(g, slots + base, raw, warp, bw, bh, lane, st.magic_dc, st.qdc);
        } else if (warp == 3 && !st.intra && !decode) {
            const int16_t* mv = p.mv + f * g.nmb * 2;
            int16_t* mvd = p.mvd + f * g.nmb * 2;
            for (int mb = lane; mb < g.nmb; mb += 32) {
                int px, py;
                mv_predictor(mv, g.mbw, mb, px, py);
                mvd[2 * mb] = (int16_t)(mv[2 * mb] - px);
                mvd[2 * mb + 1] = (int16_t)(mv[2 * mb + 1] - py);
            }
        }
        __syncthreads();
        for (int i6 = threadIdx.x; i6 < nblk; i6 += 128) {
            const int mb = (int)__umulhi((unsigned)i6, 0x2aaaaaabu), k = i6 - 6 * mb;
            if (st.intra && k < 4) continue;
            const int v = slots[chain_slot(g, mb, k)];
            rec[i6] = (int)(short)v;
            if (!decode) lvf[i6 * 64] = (int16_t)(v >> 16);
        }
        return;
    }
    if (warp == 3) {
        if (st.intra || decode) return;
        const int16_t* mv = p.mv + f * g.nmb * 2;
        int16_t* mvd = p.mvd + f * g.nmb * 2;
        for (int mb = lane; mb < g.nmb; mb += 32) {
            int px, py;
            mv_predictor(mv, g.mbw, mb, px, py);
            mvd[2 * mb] = (int16_t)(mv[2 * mb] - px);
            mvd[2 * mb + 1] = (int16_t)(mv[2 * mb + 1] - py);
        }
        return;
    }
    if (!walker) return;
    int* dc = (int*)s_chain + (chroma ? 4 * g.nmb + (warp - 1) * g.nmb : 0);   // Y[bh*bw], Cb[nmb], Cr[nmb]
    for (int wv = 0; wv < nwaves; wv++) {
        const int by_lo = max(0, (wv - (bw - 1) + 1) >> 1), by_hi = min(bh - 1, wv >> 1);
        for (int by = by_lo + lane; by <= by_hi; by += 32) {
            const int bx = wv - 2 * by, i = by * bw + bx;
            const int i6 = chain_mbmajor(g, warp, i);
            const int P = chroma ? dc_pred_chroma(dc, bw, bx, by) : dc_pred_luma(dc, bw, bx, by);
            int L;
            if (decode) L = (int)lvf[i6 * 64];
            else {
                L = quant_magic(__dsub_rn(raw[i6], (double)P), st.magic_dc, chroma);
                lvf[i6 * 64] = (int16_t)L;
            }
            const int v = L * st.qdc + P;
            dc[i] = v;
            rec[i6] = v;
        }
        __syncwarp();
    }
}

// =====================================================================================================
// Kernel C: dequantisation + IDCT + reconstruction (R9, R10, R13).  8 lanes per block.
// TAB 0: encoder table (float widened) ; TAB 1: decoder table (binary64).
// =====================================================================================================
template <int TAB>
__global__ void __launch_bounds__(TR_THREADS) idct_recon_kernel(Geom g, FramePtrs p, Step st)
{
    __shared__ double s_tile[TR_THREADS / 8][72];
    __shared__ __align__(16) int16_t s_lv[TR_THREADS / 8][72];   // 144-byte rows: the 4 groups of a warp land on different banks
    __shared__ __align__(8) uint8_t s_px[TR_THREADS / 8][72];
    const int grp = threadIdx.x >> 3, r = threadIdx.x & 7;
    const int mbi = blockIdx.x * (TR_THREADS / 8) + grp;
    const bool valid = mbi < g.nmb;
    const int mb = valid ? mbi : 0;
    const int gop = blockIdx.y;
    const size_t f = (size_t)gop * st.gop_len + st.t;
    const int mby = (int)__umulhi((unsigned)mb, g.magic_mbw), mbx = mb - mby * g.mbw;
    const uint8_t* prevf = p.rec + (f - (st.intra ? 0 : 1)) * g.fb;
    uint8_t* recf = p.rec + f * g.fb;
    const Mags M = load_mags<TAB>();
    const uint2 izr = __ldg(&g_izrow[r]);
    int mx = 0, my = 0;
    if (!st.intra) {
        const int mvw = *(const int*)(p.mv + (f * g.nmb + mb) * 2);
        mx = (int)(int16_t)(mvw & 0xffff); my = mvw >> 16;
    }
    const int16_t* lvmb = p.levels + (f * g.nmb + mb) * 384;
    const int32_t* dcrec = p.dcrec + (size_t)gop * 6 * g.nmb;
    const int k0 = st.intra ? 4 : 0;
    auto fetch = [&](int k, uint4& lv, uint2& pr, int& dc) {
        const MbBlock b = mb_block(g, mb, mbx, mby, k, mx, my);
        lv = __ldg((const uint4*)(lvmb + k * 64 + 8 * r));
        pr = make_uint2(0u, 0u);
        if (!st.intra) pr = ref_row8_packed(prevf + b.poff, b.pw, b.ph, b.pad, b.pad + b.by * 8 + r - b.my, b.pad + b.bx * 8 - b.mx);
        dc = 0;
        if (r == 0) dc = dcrec[b.dcidx];                   // level*QstepDC + P from the DC chain
    };
    uint4 lv; uint2 pr; int dc;
    fetch(k0, lv, pr, dc);
#pragma unroll 1
    for (int k = k0; k < 6; k++) {
        *(uint4*)(&s_lv[grp][8 * r]) = lv;
        *(uint2*)(&s_px[grp][8 * r]) = pr;
        const int dck = dc;
        __syncwarp();
        if (k + 1 < 6) fetch(k + 1, lv, pr, dc);           // next block's loads are in flight during this block's math
        if (k + 2 < 6) asm volatile("prefetch.global.L2 [%0];" ::"l"(lvmb + (k + 2) * 64 + 8 * r));   // and the block after that is on its way to L2
        int q[8];
#pragma unroll
        for (int u = 0; u < 8; u++) q[u] = (int)s_lv[grp][iz_byte(izr, u)] * st.qac;   // IQuantization_block
        if (r == 0) q[0] = dck;
        double t[8], R[8];
        idct_row<TAB>(q, t, M);               // row y=r
        group_transpose(t, s_tile[grp], r);   // lane r holds column x=r: t[v][r]
        idct_col<TAB>(t, R, M);               // R[y][x=r]
        uint8_t out[8];
        if (k < 4) {                             // warp-uniform
#pragma unroll
            for (int y = 0; y < 8; y++)          // mergeBlock ENC:4812 truncates first, interYReconstruct ENC:2343-2346
                out[y] = (uint8_t)clip255((int)s_px[grp][y * 8 + r] + __double2int_rz(R[y]));
        } else if (st.intra) {
#pragma unroll
            for (int y = 0; y < 8; y++) {        // chroma intra, intraImgReconstruct ENC:1964-1971
                const int v = (R[y] > 255.0) ? 255 : __double2int_rz(R[y]);
                out[y] = (uint8_t)max(v, 0);
            }
        } else {
#pragma unroll
            for (int y = 0; y < 8; y++)          // interCbCrReconstruct ENC:2605-2607 truncates the double sum
                out[y] = (uint8_t)clip255(__double2int_rz(__dadd_rn((double)s_px[grp][y * 8 + r], R[y])));
        }
        __syncwarp();
#pragma unroll
        for (int y = 0; y < 8; y++) s_px[grp][y * 8 + r] = out[y];
        __syncwarp();
        if (valid) {
            const MbBlock b = mb_block(g, mb, mbx, mby, k, 0, 0);
            *(uint2*)(recf + b.poff + (unsigned)((b.by * 8 + r) * b.pw + b.bx * 8)) = *(const uint2*)(&s_px[grp][8 * r]);
        }
        __syncwarp();                          // s_lv / s_px are rewritten by the next block
    }
}

// =====================================================================================================
// Intra luma wavefront (R11, R12 + R5-R10 for luma).  One CTA per intra frame; 8 lanes per 8x8 block; the
// blocks of anti-diagonal wave w = bx + 2*by are independent (they need L, U, UL and, for the DC predictor,
// UR).  Bottom rows / right columns of reconstructed blocks, DCs and modes live in shared memory.
// DECODE == 0: mode decision + coding + reconstruction (encoder table);
// DECODE == 1: mode from (MPM, bit) + reconstruction with the decoder's binary64 table.
// =====================================================================================================
#ifndef ICSP_IW_THREADS
#define ICSP_IW_THREADS 96
#endif
constexpr int IW_THREADS = ICSP_IW_THREADS;  // 12 block slots per pass; 8 frames per SM (measured best of 64..192 threads)
#ifndef IW_MIN_CTAS
#define IW_MIN_CTAS 8
#endif
// Edge pixels are kept in place: block (bx, by) reads the bottom row of (bx, by-1) from bot[bx*8..] and the right column
// of (bx-1, by) from right[by*8..], and overwrites exactly those slots with its own bottom row / right column when it is
// done; nobody else needs the old values (the blocks of one wave have pairwise different bx and by, waves are separated
// by __syncthreads).  One row + one column instead of a copy per block row / block column: 0.6 KB instead of 25 KB for
// CIF, which is what lets 8 frames share an SM.
struct IntraSmem {
    uint8_t* bot;    // [w]   bottom row of the block above, per column
    uint8_t* right;  // [h]   right column of the block to the left, per row
    int* dc;         // [bh][bw]
    uint8_t* mode;   // [bh][bw]
};
__host__ __device__ inline size_t intra_smem_bytes(const Geom& g)
{
    return (size_t)g.w + (size_t)g.h + (size_t)g.bh * g.bw * 5 + 16;
}

// `edges` == nullptr: the edge/DC/mode maps live in dynamic shared memory (CIF .. 720x480); otherwise they live in a
// per-GOP global scratch area of intra_smem_bytes(g) bytes (HD frames; only this CTA touches it, __syncthreads orders it).
// THREADS: 96 for batches (8 frames per SM, the throughput configuration); IW_THREADS_WIDE = 192 for small batches, where a
// frame has an SM to itself and the widest waves (22 blocks) should pass in ONE round of the CTA's 24 block slots instead of two.
constexpr int IW_THREADS_WIDE = 192;
template <int DECODE, int THREADS = IW_THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == IW_THREADS ? IW_MIN_CTAS : 2) intra_luma_kernel(Geom g, FramePtrs p, Step st, unsigned char* edges)
{
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ double s_tile[THREADS / 8][72];
    __shared__ __align__(16) int16_t s_lv[THREADS / 8][72];
    unsigned char* s_raw = edges ? edges + (size_t)blockIdx.x * ((intra_smem_bytes(g) + 15) / 16 * 16) : s_dyn;
    IntraSmem sm;
    sm.dc = (int*)s_raw;
    sm.bot = s_raw + (size_t)g.bh * g.bw * 4;
    sm.right = sm.bot + g.w;
    sm.mode = sm.right + g.h;
    constexpr int TAB = DECODE ? 1 : 0;
    const int grp = threadIdx.x >> 3, r = threadIdx.x & 7, ngrp = THREADS / 8;
    const int gop = blockIdx.x;
    const size_t f = (size_t)gop * st.gop_len + st.t;
    const uint8_t* cury = p.cur + f * g.fb;
    uint8_t* recy = p.rec + f * g.fb;
    const int bw = g.bw, bh = g.bh, w = g.w;
    const int nwaves = (bw - 1) + 2 * (bh - 1) + 1;
    const uint2 izc = __ldg(&g_izcol[r]), izr = __ldg(&g_izrow[r]);
    const Mags M0 = load_mags<0>();
    const Mags MI = load_mags<TAB>();

    for (int wv = 0; wv < nwaves; wv++) {
        const int by_lo = max(0, (wv - (bw - 1) + 1) >> 1), by_hi = min(bh - 1, wv >> 1);
        for (int by0 = by_lo; by0 <= by_hi; by0 += ngrp) {  // whole warps iterate together
            const int by = by0 + grp;
            const bool active = by <= by_hi;
            // a warp holds 4 block slots: when none of them has a block in this wave the whole warp skips the body
            // (warp-uniform, so the __syncwarp / shuffles inside stay converged) and only meets the others at the barrier
            if (by0 + (grp & ~3) > by_hi) continue;
            const int bx = active ? wv - 2 * by : 0, byy = active ? by : 0;
            const bool hasL = bx > 0, hasU = byy > 0;
            const int mb = (byy >> 1) * g.mbw + (bx >> 1), k = ((byy & 1) << 1) | (bx & 1);

            // neighbours from the reconstructed plane (shared-memory edges)
            int up[8], left_r, sumU = 0;
#pragma unroll
            for (int x = 0; x < 8; x++) { up[x] = hasU ? sm.bot[bx * 8 + x] : 128; sumU += up[x]; }
            const int up_r = hasU ? sm.bot[bx * 8 + r] : 128;   // column r's own upper neighbour
            left_r = hasL ? sm.right[byy * 8 + r] : 128;
            const int sumL = group_sum(left_r);
            // predVal = (predValLeft + predValUpper) / 16.0, a missing side contributes 128*8 (ENC:711-733)
            const double pd = __ddiv_rn((double)((hasL ? sumL : 1024) + (hasU ? sumU : 1024)), 16.0);

            int mode;
            int e[8];
            if (!DECODE) {
                const uint2 cwd = *(const uint2*)(cury + (size_t)(byy * 8 + r) * w + bx * 8);
                int c[8], e0[8], e1[8], e2[8], sae0 = 0, sae1 = 0, sae2 = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) { c[i] = (cwd.x >> (8 * i)) & 255; c[4 + i] = (cwd.y >> (8 * i)) & 255; }
#pragma unroll
                for (int x = 0; x < 8; x++) {
                    e0[x] = c[x] - up[x];                                       // DPCM_pix_0 (vertical)
                    e1[x] = c[x] - left_r;                                      // DPCM_pix_1 (horizontal)
                    e2[x] = __double2int_rz(__dsub_rn((double)c[x], pd));       // DPCM_pix_2 (DC): (int)(cur - predVal)
                    sae0 += abs(e0[x]); sae1 += abs(e1[x]); sae2 += abs(e2[x]);
                }
                sae0 = group_sum(sae0); sae1 = group_sum(sae1); sae2 = group_sum(sae2);
                if (hasL && hasU) {
                    const int m = min(min(sae0, sae1), sae2);
                    mode = (m == sae0) ? 0 : ((m == sae1) ? 1 : 2);             // ENC:958-976
                } else if (hasL) mode = (sae2 > sae1) ? 1 : 2;                  // ENC:897-908
                else if (hasU) mode = (sae2 > sae0) ? 0 : 2;
                else mode = 2;
#pragma unroll
                for (int x = 0; x < 8; x++) e[x] = mode == 0 ? e0[x] : (mode == 1 ? e1[x] : e2[x]);
            }
            // mode predictor p (ENC:1334-1350 / decoder inverse ENC:1793-1795)
            int pm = 2;
            if (hasL && hasU) pm = med3(sm.mode[byy * bw + bx - 1], sm.mode[(byy - 1) * bw + bx - 1], sm.mode[(byy - 1) * bw + bx]);
            else if (hasL) pm = sm.mode[byy * bw + bx - 1];
            else if (hasU) pm = sm.mode[(byy - 1) * bw + bx];
            const size_t mi = (f * g.nmb + mb) * 4 + k;
            if (!DECODE) {
                int mpmf = 0, bit = 0;
                if (hasL || hasU) {
                    mpmf = (mode == pm);
                    if (!mpmf) bit = (pm == 0) ? (mode == 1 ? 0 : 1) : (mode == 0 ? 0 : 1);
                }
                if (active && r == 0) { p.mpm[mi] = (uint8_t)mpmf; p.ipm[mi] = (uint8_t)bit; }
            } else {
                const int mpmf = p.mpm[mi], bit = p.ipm[mi];
                if (!hasL && !hasU) mode = 2;
                else if (mpmf) mode = pm;
                else if (pm == 0) mode = bit == 0 ? 1 : 2;
                else if (pm == 2) mode = bit == 0 ? 0 : 1;
                else mode = bit == 0 ? 0 : 2;
            }

            int16_t* lv = p.levels + ((f * g.nmb + mb) * 6 + k) * 64;
            int q[8];  // dequantised coefficients of ROW r after the transposition below
            int P = 0;
            if (r == 0) {   // A.5, branch free (the four groups of a warp would otherwise serialise on different cases)
                const int i = byy * bw + bx;
                const bool ur = (bx & 1) == 0 || ((byy & 1) == 0 && bx != bw - 1);
                int a = i - 1, b = i - bw, c2 = ur ? i - bw + 1 : i - bw - 1;
                if (byy == 0) { a = bx == 0 ? 0 : i - 1; b = a; c2 = a; }
                else if (bx == 0) { a = b; c2 = b; }
                const int va = sm.dc[a], vb = sm.dc[b], vc = sm.dc[c2];
                P = (bx == 0 && byy == 0) ? 1024 : max(min(va, vb), min(max(va, vb), vc));
            }
            if (!DECODE) {
                double t[8], D[8];
                fdct_row(e, t, M0);
                group_transpose(t, s_tile[grp], r);
                fdct_col(t, r, D, M0);
                if (p.dct_tap && active) {     // debug tap: forward DCT output before the DC predictor is subtracted
                    double* tap = p.dct_tap + ((f * g.nmb + mb) * 6 + k) * 64 + r;
#pragma unroll
                    for (int v = 0; v < 8; v++) tap[v * 8] = D[v];
                }
                if (r == 0) D[0] = __dsub_rn(D[0], (double)P);
                int nz = 0, L[8];
#pragma unroll
                for (int v = 0; v < 8; v++) {
                    L[v] = quant_magic(D[v], (v == 0 && r == 0) ? st.magic_dc : st.magic_ac, false);
                    s_lv[grp][iz_byte(izc, v)] = (int16_t)L[v];
                    if (!(v == 0 && r == 0)) nz |= L[v];
                }
                const unsigned bal = __ballot_sync(0xffffffffu, nz != 0);
                __syncwarp();
                if (active) {
                    *(uint4*)(lv + 8 * r) = *(const uint4*)(&s_lv[grp][8 * r]);
                    if (r == 0) p.acflag[(f * g.nmb + mb) * 6 + k] = ((bal >> (8 * (grp & 3))) & 0xffu) ? 0 : 1;
                }
            } else {
                *(uint4*)(&s_lv[grp][8 * r]) = *(const uint4*)(lv + 8 * r);
                __syncwarp();
            }
#pragma unroll
            for (int u = 0; u < 8; u++) q[u] = (int)s_lv[grp][iz_byte(izr, u)] * ((r == 0 && u == 0) ? st.qdc : st.qac);
            if (r == 0) { q[0] += P; if (active) sm.dc[byy * bw + bx] = q[0]; }
            double t2[8], R[8];
            idct_row<TAB>(q, t2, MI);
            group_transpose(t2, s_tile[grp], r);
            idct_col<TAB>(t2, R, MI);   // lane r holds column x=r

            // IDPCM_pix_0/1/2 (ENC:744-850): (int)(idct + pred) with the sum formed in double, then clipped
            int lefts[8];
#pragma unroll
            for (int y = 0; y < 8; y++) lefts[y] = __shfl_sync(0xffffffffu, left_r, (threadIdx.x & 24) | y);
            uint8_t out[8];
#pragma unroll
            for (int y = 0; y < 8; y++) {
                const double pv = mode == 0 ? (double)up_r : (mode == 1 ? (double)lefts[y] : pd);
                out[y] = (uint8_t)clip255(__double2int_rz(__dadd_rn(R[y], pv)));
            }
            uint8_t* px = (uint8_t*)s_lv[grp];
            __syncwarp();
#pragma unroll
            for (int y = 0; y < 8; y++) px[y * 8 + r] = out[y];
            __syncwarp();
            if (active) {
                *(uint2*)(recy + (size_t)(byy * 8 + r) * w + bx * 8) = *(const uint2*)(&px[8 * r]);
                sm.bot[bx * 8 + r] = out[7];
                if (r == 7) {
#pragma unroll
                    for (int y = 0; y < 8; y++) sm.right[byy * 8 + y] = out[y];
                }
                if (r == 0) sm.mode[byy * bw + bx] = (uint8_t)mode;
            }
            __syncwarp();
        }
        __syncthreads();
    }
}

// =====================================================================================================
// Motion estimation (R1, R2).  One CTA per macroblock-row segment (<= 22 MBs; a whole CIF row) per frame.
// The 48-row search window of the previous reconstruction (apron = clamp + the reference's zero last
// row/column, never materialised in HBM) and the 16 current rows are staged in shared memory with 16-byte
// loads.  The window is kept in FOUR byte-shifted copies so that every candidate row is word aligned in the
// copy (16+dx)&3: the inner loop is 4 LDS.32 + 1 broadcast LDS.128 + 4 VABSDIFF4.ACC per row, no funnel shifts.
// One warp per macroblock, two candidates per lane.  Row pitch == 8 (mod 32) words and copy offsets {0,1,1,1}
// make the 64 state-0 candidates fall exactly two per bank, and c_slot assigns them to (round, lane) so that
// both rounds are bank-conflict free.  Winner selection (redux.sync min) reproduces visiting order, strict-<
// tie-break and the second-zero early break.  `fixup` re-evaluates MBs whose carried start state is != 0.
// =====================================================================================================
struct MeLayout {
    int seg_mbs;   // macroblocks per segment
    int nseg;      // segments per macroblock row
    int row_w;     // valid words per window row = (seg_mbs*16 + 32)/4
    int pitch_w;   // words per window row in shared memory, == 8 (mod 32)
    int copy_w;    // words per copy = 48*pitch_w
};
// start state, round, lane -> packed candidate handled by that lane: visit index | (dx & 255) << 8 | (dy & 255) << 16.
// Lives in global memory (read through the read-only path, coalesced over lanes): a lane-indexed __constant__
// table would serialise into 32 constant-cache replays per warp.
__device__ uint32_t g_slot[8][2][32];
__host__ __device__ inline int me_copy_off(const MeLayout& L, int s) { return s * L.copy_w + (s ? 1 : 0); }  // +1: see me_stage
__host__ __device__ inline size_t me_smem_bytes(const MeLayout& L) { return (size_t)(4 * L.copy_w + 8) * 4 + (size_t)16 * L.seg_mbs * 16; }

// sum of absolute differences of the four packed bytes, accumulated: one VABSDIFF4.U8.ACC.  (Written as
// __vsadu4(a,b)+c the compiler zeroes the accumulator and adds with a separate IADD3 per pair of SADs.)
__device__ __forceinline__ uint32_t sad4_acc(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ uint4 window_chunk(const uint8_t* __restrict__ plane, int w, int h, int py, int pxc)
{   // 16 bytes of the padded image (pad 16) at padded row py, padded columns [16*pxc, 16*pxc+16)
    const int PH = h + 32, PW = w + 32;
    if (py == PH - 1) return make_uint4(0, 0, 0, 0);
    const int yy = min(max(py - 16, 0), h - 1);
    const uint8_t* row = plane + (unsigned)(yy * w);    // a luma plane is far below 4 GB: 32-bit offsets
    const int x0 = pxc * 16 - 16;
    if (x0 >= 0 && x0 + 16 <= w) return __ldg((const uint4*)(row + (unsigned)x0));
    uint32_t v;
    if (x0 < 0) { v = row[0]; v *= 0x01010101u; return make_uint4(v, v, v, v); }
    v = row[w - 1]; v *= 0x01010101u;
    uint4 o = make_uint4(v, v, v, v);
    if (pxc * 16 + 16 == PW) o.w &= 0x00ffffffu;  // last padded column stays 0
    return o;
}

#ifndef ICSP_ME_RAWCHUNK
#define ICSP_ME_RAWCHUNK 0   // 1: load the apron byte raw and replicate it at store time (measured 2 % slower: more live registers across the search)
#endif
// The same fetch split in two, for the software-pipelined kernel: window_chunk_raw only LOADS (apron chunks: the one edge
// byte, untouched, in .x), window_chunk_finish turns that into the replicated chunk when the data is consumed a whole
// macroblock search later.  (Multiplying the byte right after the load made every staging warp wait for the global
// latency at the top of each macroblock row: 6 % of the kernel's stall samples.)
__device__ __forceinline__ uint4 window_chunk_raw(const uint8_t* __restrict__ plane, int w, int h, int py, int pxc)
{
    const int PH = h + 32;
    if (py == PH - 1) return make_uint4(0, 0, 0, 0);
    const int yy = min(max(py - 16, 0), h - 1);
    const uint8_t* row = plane + (unsigned)(yy * w);
    const int x0 = pxc * 16 - 16;
    if (x0 >= 0 && x0 + 16 <= w) return __ldg((const uint4*)(row + (unsigned)x0));
    return make_uint4((uint32_t)__ldg(row + (x0 < 0 ? 0 : w - 1)), 0, 0, 0);
}
__device__ __forceinline__ uint4 window_chunk_finish(uint4 v, int w, int h, int py, int pxc)
{
    const int x0 = pxc * 16 - 16;
    if (py == h + 31 || (x0 >= 0 && x0 + 16 <= w)) return v;
    const uint32_t r = v.x * 0x01010101u;
    uint4 o = make_uint4(r, r, r, r);
    if (x0 >= 0 && pxc * 16 + 16 == w + 32) o.w &= 0x00ffffffu;  // last padded column stays 0
    return o;
}

// Stage the window and the current rows of segment (mby, m0..m0+nmbs).  One warp per window row, one lane per
// 16-byte chunk: copy 0 is the plain window; copy s (s = 1..3) holds the window shifted left by s bytes and stored
// one word to the right (word i+1 of copy s = window bytes [4i+s, 4i+s+4)), so all four stores are aligned STS.128.
template <bool SHIFTED>
__device__ __forceinline__ void me_stage(const Geom& g, const MeLayout& L, uint32_t* s_win, uint8_t* s_cur,
                                         const uint8_t* __restrict__ cury, const uint8_t* __restrict__ refy, int mby, int m0, int nmbs)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int chunks = nmbs + 2;
    for (int row = warp; row < 48; row += nwarps) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (lane < chunks) v = window_chunk(refy, g.w, g.h, mby * 16 + row, m0 + lane);
        uint32_t* dst = s_win + row * L.pitch_w + lane * 4;
        if (lane < chunks) *(uint4*)dst = v;
        if (SHIFTED) {
            uint32_t prev = __shfl_up_sync(0xffffffffu, v.w, 1);
            if (lane == 0) prev = 0;
            if (lane < chunks) {
#pragma unroll
                for (int sft = 1; sft < 4; sft++) {
                    uint4 o;
                    o.x = __funnelshift_r(prev, v.x, 8 * sft);
                    o.y = __funnelshift_r(v.x, v.y, 8 * sft);
                    o.z = __funnelshift_r(v.y, v.z, 8 * sft);
                    o.w = __funnelshift_r(v.z, v.w, 8 * sft);
                    *(uint4*)(dst + sft * L.copy_w) = o;
                }
            }
        }
    }
    for (int row = warp; row < 16; row += nwarps)
        if (lane < nmbs)
            *(uint4*)(s_cur + (row * L.seg_mbs + lane) * 16) = __ldg((const uint4*)(cury + (size_t)(mby * 16 + row) * g.w + (m0 + lane) * 16));
    __syncthreads();
}

// rows_per_cta macroblock rows per CTA (blockIdx.x = row group * nseg + segment): 1 for the speculative pass of
// non-persistent builds, mbh for the fixup pass, so that the usual "nothing to fix" launch is G*nseg cheap CTAs.
__device__ __forceinline__ void me_search_rows(const Geom& g, const MeLayout& L, const FramePtrs& p, const Step& st, int gop, int seg, int fixup,
                                               int row_begin, int row_end, unsigned char* s_me)
{
    const int m0 = seg * L.seg_mbs, nmbs = min(L.seg_mbs, g.mbw - m0);
    const size_t f = (size_t)gop * st.gop_len + st.t;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* s_win = (uint32_t*)s_me;
    uint8_t* s_cur = s_me + (size_t)(4 * L.copy_w + 8) * 4;
  for (int mby = row_begin; mby < row_end; mby++) {
    const uint8_t* states = p.mestate + (size_t)gop * g.nmb + mby * g.mbw + m0;
    if (fixup) {
        int any = 0;
        for (int i = threadIdx.x; i < nmbs; i += blockDim.x) any |= states[i];
        if (!__syncthreads_or(any)) continue;
    }
    me_stage<true>(g, L, s_win, s_cur, p.cur + f * g.fb, p.rec + (f - 1) * g.fb, mby, m0, nmbs);
    const int mbl = warp;
    const int state = (fixup && warp < nmbs) ? states[mbl] : 0;
    if (warp < nmbs && !(fixup && state == 0)) {
    uint32_t sad[2];
    int vidx[2];
#pragma unroll
    for (int rnd = 0; rnd < 2; rnd++) {
        const uint32_t pk = __ldg(&g_slot[state][rnd][lane]);
        const int idx = pk & 255, dx = (int)(int8_t)(pk >> 8), dy = (int)(int8_t)(pk >> 16);
        const int col = mbl * 16 + 16 + dx;   // window byte column of the candidate's first pixel
        const uint32_t* wrow = s_win + me_copy_off(L, col & 3) + (16 + dy) * L.pitch_w + (col >> 2);
        const uint4* crow = (const uint4*)(s_cur + mbl * 16);
        uint32_t acc = 0;
#pragma unroll 8
        for (int j = 0; j < 16; j++) {
            const uint4 c = *crow;
            acc = __vsadu4(wrow[0], c.x) + acc;
            acc = __vsadu4(wrow[1], c.y) + acc;
            acc = __vsadu4(wrow[2], c.z) + acc;
            acc = __vsadu4(wrow[3], c.w) + acc;
            wrow += L.pitch_w;
            crow += L.seg_mbs;
        }
        sad[rnd] = acc;
        vidx[rnd] = idx;
    }
    // the search breaks at the SECOND zero-SAD visit (ENC:2130-2141); otherwise the first visited minimum wins
    const unsigned zk0 = sad[0] == 0 ? (unsigned)vidx[0] : 64u, zk1 = sad[1] == 0 ? (unsigned)vidx[1] : 64u;
    const unsigned z1 = __reduce_min_sync(0xffffffffu, min(zk0, zk1));
    int win = -1, moves = 64;
    uint32_t best = 0;
    if (z1 < 64u) {
        const unsigned z2 = __reduce_min_sync(0xffffffffu, min(zk0 > z1 ? zk0 : 64u, zk1 > z1 ? zk1 : 64u));
        if (z2 < 64u) { win = (int)z2; moves = win + 1; }
    }
    if (win < 0) {
        const unsigned key = __reduce_min_sync(0xffffffffu, min(sad[0] * 64u + vidx[0], sad[1] * 64u + vidx[1]));
        win = key & 63; best = key >> 6;
    }
    if (lane == 0) {
        const int mb = mby * g.mbw + m0 + mbl;
        // mv = MB origin - best position (ENC:2145-2146)
        const int wdx = c_cand[state][win][0], wdy = c_cand[state][win][1];   // warp-uniform index: broadcast
        *(int*)(p.mv + (f * g.nmb + mb) * 2) = ((-wdx) & 0xffff) | ((-wdy) << 16);
        p.minsad[f * g.nmb + mb] = (int32_t)best;
        if (!fixup) {
            p.memoves[(size_t)gop * g.nmb + mb] = (uint8_t)moves;
            if (moves < 64) atomicAdd(&p.meflag[gop], 1u);
        }
    }
    }
    __syncthreads();   // the next row restages the window
  }
}
__global__ void __launch_bounds__(704) me_sad_kernel(Geom g, MeLayout L, FramePtrs p, Step st, int fixup, int rows_per_cta)
{
    extern __shared__ __align__(16) unsigned char s_me[];
    const int gop = blockIdx.y, rg = blockIdx.x / L.nseg, seg = blockIdx.x - rg * L.nseg;
    if (fixup && p.meflag[gop] == 0) return;
    me_search_rows(g, L, p, st, gop, seg, fixup, rg * rows_per_cta, min(g.mbh, (rg + 1) * rows_per_cta), s_me);
}

// Exact fallback, pass 1: for frames where some search broke early, find for every macroblock and every one of
// the 8 possible start states which visits have SAD == 0 (block identical to the candidate).  The break point
// of a search depends only on these masks, never on non-zero SAD values.
__device__ __forceinline__ void me_zero_rows(const Geom& g, const MeLayout& L, const FramePtrs& p, const Step& st, int gop, int seg, unsigned char* s_me)
{
    const int m0 = seg * L.seg_mbs, nmbs = min(L.seg_mbs, g.mbw - m0);
    const size_t f = (size_t)gop * st.gop_len + st.t;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* s_win = (uint32_t*)s_me;
    uint8_t* s_cur = s_me + (size_t)(4 * L.copy_w + 8) * 4;
  for (int mby = 0; mby < g.mbh; mby++) {
    me_stage<false>(g, L, s_win, s_cur, p.cur + f * g.fb, p.rec + (f - 1) * g.fb, mby, m0, nmbs);
    const int mbl = warp;
    if (warp < nmbs)
    for (int state = 0; state < 8; state++) {
        unsigned long long z = 0;
#pragma unroll
        for (int h2 = 0; h2 < 2; h2++) {
            const int idx = lane + 32 * h2;
            const int dx = c_cand[state][idx][0], dy = c_cand[state][idx][1];
            const int col = mbl * 16 + 16 + dx;
            const uint32_t* wrow = s_win + (16 + dy) * L.pitch_w + (col >> 2);
            const int sh = (col & 3) * 8;
            const uint32_t* crow = (const uint32_t*)(s_cur + mbl * 16);
            uint32_t diff = 0;
            for (int j = 0; j < 16 && diff == 0; j++) {   // early out on the first differing row
                const uint32_t w0 = wrow[0], w1 = wrow[1], w2 = wrow[2], w3 = wrow[3], w4 = wrow[4];
                diff |= __funnelshift_r(w0, w1, sh) ^ crow[0];
                diff |= __funnelshift_r(w1, w2, sh) ^ crow[1];
                diff |= __funnelshift_r(w2, w3, sh) ^ crow[2];
                diff |= __funnelshift_r(w3, w4, sh) ^ crow[3];
                wrow += L.pitch_w;
                crow += L.seg_mbs * 4;
            }
            z |= (unsigned long long)__ballot_sync(0xffffffffu, diff == 0) << (32 * h2);
        }
        if (lane == 0) p.mezero[((size_t)gop * g.nmb + mby * g.mbw + m0 + mbl) * 8 + state] = z;
    }
    __syncthreads();   // the next row restages the window
  }
}
__global__ void __launch_bounds__(704) me_zero_kernel(Geom g, MeLayout L, FramePtrs p, Step st)
{   // grid (nseg, G): one CTA per frame segment walks all macroblock rows (it exits at once for unflagged frames)
    extern __shared__ __align__(16) unsigned char s_me[];
    if (p.meflag[blockIdx.y] == 0) return;
    me_zero_rows(g, L, p, st, blockIdx.y, blockIdx.x, s_me);
}

// Exact fallback, pass 2: resolve the carried spiral state of every macroblock of a flagged frame.  A search
// started in state s makes m = (index of its second zero-SAD visit)+1 moves, or 64; the state handed to the next
// macroblock is c_next[s][m] (ENC:2094-2095: flag/xflag/yflag are initialised once per frame, never per MB).
__device__ __forceinline__ void me_chain_walk(const Geom& g, const FramePtrs& p, int gop)
{   // one thread
    const unsigned long long* zm = p.mezero + (size_t)gop * g.nmb * 8;
    uint8_t* states = p.mestate + (size_t)gop * g.nmb;
    int s = 0;
    for (int mb = 0; mb < g.nmb; mb++) {
        states[mb] = (uint8_t)s;
        const unsigned long long z = zm[mb * 8 + s];
        int m = 64;
        if (__popcll(z) >= 2) m = __ffsll((long long)(z & (z - 1)));
        s = c_next[s][m];
    }
}
// The same walk for a whole CTA that has `bytes` of shared memory to spare: the 8 masks of every macroblock are copied in by
// all threads first (the single-thread walk over global memory is one dependent ~1 us load per macroblock: 0.3 ms per CIF frame).
__device__ __forceinline__ void me_chain_walk_cta(const Geom& g, const FramePtrs& p, int gop, unsigned char* s_buf, size_t bytes)
{
    const unsigned long long* zm = p.mezero + (size_t)gop * g.nmb * 8;
    if ((size_t)g.nmb * 64 > bytes) {
        if (threadIdx.x == 0) me_chain_walk(g, p, gop);
        return;
    }
    unsigned long long* sz = (unsigned long long*)s_buf;
    for (int i = threadIdx.x; i < g.nmb * 8; i += blockDim.x) sz[i] = zm[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        uint8_t* states = p.mestate + (size_t)gop * g.nmb;
        int s = 0;
        for (int mb = 0; mb < g.nmb; mb++) {
            states[mb] = (uint8_t)s;
            const unsigned long long z = sz[mb * 8 + s];
            int m = 64;
            if (__popcll(z) >= 2) m = __ffsll((long long)(z & (z - 1)));
            s = c_next[s][m];
        }
    }
}
__global__ void __launch_bounds__(32) me_chain_kernel(Geom g, FramePtrs p)
{
    const int gop = blockIdx.x;
    if (p.meflag[gop] == 0) return;
    if (threadIdx.x != 0) return;
    me_chain_walk(g, p, gop);
}

// The three fallback passes of one frame in ONE launch (frames that are one segment wide): zero masks, chain walk, re-search.
// Used after the row-parallel search of small batches, where three more launches would cost more than the search itself.
__global__ void __launch_bounds__(704) me_fallback_kernel(Geom g, MeLayout L, FramePtrs p, Step st)
{
    extern __shared__ __align__(16) unsigned char s_me[];
    const int gop = blockIdx.x;
    if (p.meflag[gop] == 0) return;             // natural content: the speculative state-0 search is the answer
    me_zero_rows(g, L, p, st, gop, 0, s_me);
    __syncthreads();
    me_chain_walk_cta(g, p, gop, s_me, me_smem_bytes(L));
    __syncthreads();
    me_search_rows(g, L, p, st, gop, 0, 1, 0, g.mbh, s_me);
}

// ---- persistent variant of the speculative (state 0) search ------------------------------------------------------
// One CTA per (frame, segment) walks DOWN the macroblock rows.  Consecutive rows share 32 of their 48 window rows, so
// the window lives in a 64-row ring (4 bands of 16 rows, + a 15-row mirror of the ring head so that a candidate's 16
// rows never wrap) and only the 16 new rows and the next 16 current rows are fetched per step.  The fetch is software
// pipelined inside every warp: at the top of step r each warp issues the global loads of its share of band r+3 /
// current rows r+1 into registers, then searches its macroblock of row r (hiding the load latency), then writes the
// registers to shared memory (ring slot of band r-1, last read at step r-1; current-row buffer (r+1)&1), and one
// __syncthreads closes the step.
constexpr int ME_RING_ROWS = 64, ME_MIRROR_ROWS = 15;
// words per shifted copy of the ring, rounded to a multiple of 32 so that the copy bases keep the bank residues
// {0,1,1,1} the candidate slot table was built for
__host__ __device__ inline int me_ring_copy_w(const MeLayout& L) { return ((ME_RING_ROWS + ME_MIRROR_ROWS) * L.pitch_w + 31) & ~31; }
__host__ __device__ inline size_t me_frame_smem_bytes(const MeLayout& L)
{
    return (size_t)(4 * me_ring_copy_w(L) + 8) * 4 + (size_t)2 * 16 * L.seg_mbs * 16;
}

// store one 16-byte window chunk `v` (chunk index = lane) of padded row prow into the four shifted copies (+ mirror)
__device__ __forceinline__ void me_store_chunk(const MeLayout& L, uint32_t* s_win, int prow, uint4 v, int lane, int chunks)
{
    const int copy_w = me_ring_copy_w(L);
    const int slot = prow & (ME_RING_ROWS - 1);
    uint32_t prev = __shfl_up_sync(0xffffffffu, v.w, 1);
    if (lane == 0) prev = 0;
    if (lane < chunks) {
        uint4 o[4];
        o[0] = v;
#pragma unroll
        for (int sft = 1; sft < 4; sft++) {
            o[sft].x = __funnelshift_r(prev, v.x, 8 * sft);
            o[sft].y = __funnelshift_r(v.x, v.y, 8 * sft);
            o[sft].z = __funnelshift_r(v.y, v.z, 8 * sft);
            o[sft].w = __funnelshift_r(v.z, v.w, 8 * sft);
        }
        uint32_t* dst = s_win + slot * L.pitch_w + lane * 4;
#pragma unroll
        for (int sft = 0; sft < 4; sft++) *(uint4*)(dst + sft * copy_w) = o[sft];
        if (slot < ME_MIRROR_ROWS) {
            dst += ME_RING_ROWS * L.pitch_w;
#pragma unroll
            for (int sft = 0; sft < 4; sft++) *(uint4*)(dst + sft * copy_w) = o[sft];
        }
    }
}
// fetch task `task` of a step: 0..15 = window row of `band`, 16..31 = current row of macroblock row cur_mby
__device__ __forceinline__ uint4 me_task_load(const Geom& g, const uint8_t* __restrict__ refy, const uint8_t* __restrict__ cury, int task,
                                              int band, int cur_mby, int m0, int nmbs, int lane)
{
    uint4 v = make_uint4(0, 0, 0, 0);
#if ICSP_ME_RAWCHUNK
    if (task < 16) { if (lane < nmbs + 2) v = window_chunk_raw(refy, g.w, g.h, band * 16 + task, m0 + lane); }
#else
    if (task < 16) { if (lane < nmbs + 2) v = window_chunk(refy, g.w, g.h, band * 16 + task, m0 + lane); }
#endif
    else if (task < 32) { if (lane < nmbs) v = __ldg((const uint4*)(cury + (unsigned)((cur_mby * 16 + task - 16) * g.w + (m0 + lane) * 16))); }
    return v;
}
__device__ __forceinline__ void me_task_store(const Geom& g, const MeLayout& L, uint32_t* s_win, uint8_t* s_cur, int task, int band, uint4 v, int m0, int nmbs, int lane)
{
#if ICSP_ME_RAWCHUNK
    if (task < 16) me_store_chunk(L, s_win, band * 16 + task, window_chunk_finish(v, g.w, g.h, band * 16 + task, m0 + lane), lane, nmbs + 2);
#else
    if (task < 16) me_store_chunk(L, s_win, band * 16 + task, v, lane, nmbs + 2);
#endif
    else if (task < 32) { if (lane < nmbs) *(uint4*)(s_cur + ((task - 16) * L.seg_mbs + lane) * 16) = v; }
}

// CPITCH / CSEG: compile-time copies of L.pitch_w / L.seg_mbs (0 = use the runtime values).  With constants every
// shared-memory address of the inner loop is base + immediate, which removes the per-row pointer arithmetic.
// (704, 1): one CTA per SM by shared memory anyway, so let the compiler use up to 93 registers to keep more loads in flight
// `fused` (frames of one segment, nseg == 1): the CTA that searched the frame also knows whether any of its searches broke
// early, so it runs the exact carried-state fallback (zero masks -> chain -> re-search) itself, in the shared memory it
// already owns.  The three fallback kernels then never have to be launched: as separate launches of 704-thread, 142 KB CTAs
// they could only be scheduled once a whole SM had drained, and cost each chunk 0.3-0.5 ms of dead time per step although
// they exit at once on natural content (profiles/README.md, round 2).
template <int CPITCH, int CSEG, bool FUSED>
__global__ void __launch_bounds__(704, 1) me_sad_frame_kernel(Geom g, MeLayout L, FramePtrs p, Step st)
{
    extern __shared__ __align__(16) unsigned char s_me[];
    constexpr bool fused = FUSED;
    unsigned* const s_breaks_p = (unsigned*)(s_me + me_frame_smem_bytes(L));   // FUSED: 16 more bytes of dynamic shared memory
    if (FUSED && threadIdx.x == 0) *s_breaks_p = 0;
    const int pitch_w = CPITCH ? CPITCH : L.pitch_w, seg_mbs = CSEG ? CSEG : L.seg_mbs;
    const int gop = blockIdx.y, seg = blockIdx.x;
    const int m0 = seg * seg_mbs, nmbs = min(seg_mbs, g.mbw - m0);
    const size_t f = (size_t)gop * st.gop_len + st.t;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int copy_w = me_ring_copy_w(L);
    uint32_t* s_win = (uint32_t*)s_me;
    uint8_t* s_cur0 = s_me + (size_t)(4 * copy_w + 8) * 4;
    const size_t cur_bytes = (size_t)16 * seg_mbs * 16;
    const uint8_t* cury = p.cur + f * g.fb;
    const uint8_t* refy = p.rec + (f - 1) * g.fb;

    // prologue: bands 0..2 and current rows of macroblock row 0
    for (int band = 0; band < 3; band++)
        for (int task = warp; task < (band == 0 ? 32 : 16); task += nwarps)
            me_task_store(g, L, s_win, s_cur0, task, band, me_task_load(g, refy, cury, task, band, 0, m0, nmbs, lane), m0, nmbs, lane);
    __syncthreads();

    const int mbl = warp;   // nwarps == seg_mbs >= nmbs
    const uint32_t pk0 = __ldg(&g_slot[0][0][lane]), pk1 = __ldg(&g_slot[0][1][lane]);
    const int idx0 = pk0 & 255, dx0 = (int)(int8_t)(pk0 >> 8), dy0 = (int)(int8_t)(pk0 >> 16);
    const int idx1 = pk1 & 255, dx1 = (int)(int8_t)(pk1 >> 8), dy1 = (int)(int8_t)(pk1 >> 16);
    const int col0 = mbl * 16 + 16 + dx0, col1 = mbl * 16 + 16 + dx1;
    const int off0 = (col0 & 3) * copy_w + ((col0 & 3) ? 1 : 0) + (col0 >> 2);
    const int off1 = (col1 & 3) * copy_w + ((col1 & 3) ? 1 : 0) + (col1 >> 2);
    for (int mby = 0; mby < g.mbh; mby++) {
        const bool more = mby + 1 < g.mbh;
        uint8_t* s_cur_next = s_cur0 + ((mby + 1) & 1) * cur_bytes;
        // (1) issue this warp's share of the next step's loads
        uint4 pre0 = make_uint4(0, 0, 0, 0), pre1 = pre0;
        if (more) {
            pre0 = me_task_load(g, refy, cury, warp, mby + 3, mby + 1, m0, nmbs, lane);
            pre1 = me_task_load(g, refy, cury, warp + nwarps, mby + 3, mby + 1, m0, nmbs, lane);
        }
        // (2) search macroblock (mby, m0 + mbl): both candidates of the lane share the current-row loads
        if (mbl < nmbs) {   // the last segment of a frame can be one macroblock narrower than the CTA
            const uint4* crow = (const uint4*)(s_cur0 + (mby & 1) * cur_bytes + mbl * 16);
            // rows slot..slot+15 are contiguous thanks to the mirror
            const uint32_t* w0 = s_win + off0 + ((mby * 16 + 16 + dy0) & (ME_RING_ROWS - 1)) * pitch_w;
            const uint32_t* w1 = s_win + off1 + ((mby * 16 + 16 + dy1) & (ME_RING_ROWS - 1)) * pitch_w;
            uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const uint4 c = crow[j * seg_mbs];
                a0 = sad4_acc(w0[j * pitch_w + 0], c.x, a0);
                a1 = sad4_acc(w0[j * pitch_w + 1], c.y, a1);
                a2 = sad4_acc(w0[j * pitch_w + 2], c.z, a2);
                a3 = sad4_acc(w0[j * pitch_w + 3], c.w, a3);
                b0 = sad4_acc(w1[j * pitch_w + 0], c.x, b0);
                b1 = sad4_acc(w1[j * pitch_w + 1], c.y, b1);
                b2 = sad4_acc(w1[j * pitch_w + 2], c.z, b2);
                b3 = sad4_acc(w1[j * pitch_w + 3], c.w, b3);
            }
            const uint32_t sad0 = (a0 + a1) + (a2 + a3), sad1 = (b0 + b1) + (b2 + b3);
            const unsigned zk0 = sad0 == 0 ? (unsigned)idx0 : 64u, zk1 = sad1 == 0 ? (unsigned)idx1 : 64u;
            const unsigned z1 = __reduce_min_sync(0xffffffffu, min(zk0, zk1));
            int win = -1, moves = 64;
            uint32_t best = 0;
            if (z1 < 64u) {
                const unsigned z2 = __reduce_min_sync(0xffffffffu, min(zk0 > z1 ? zk0 : 64u, zk1 > z1 ? zk1 : 64u));
                if (z2 < 64u) { win = (int)z2; moves = win + 1; }
            }
            if (win < 0) {
                const unsigned key = __reduce_min_sync(0xffffffffu, min(sad0 * 64u + idx0, sad1 * 64u + idx1));
                win = key & 63; best = key >> 6;
            }
            if (lane == 0) {
                const int mb = mby * g.mbw + m0 + mbl;
                const int wdx = c_cand[0][win][0], wdy = c_cand[0][win][1];
                *(int*)(p.mv + (f * g.nmb + mb) * 2) = ((-wdx) & 0xffff) | ((-wdy) << 16);
                p.minsad[f * g.nmb + mb] = (int32_t)best;
                p.memoves[(size_t)gop * g.nmb + mb] = (uint8_t)moves;
                if (moves < 64) { if (fused) atomicAdd(s_breaks_p, 1u); else atomicAdd(&p.meflag[gop], 1u); }
            }
        }
        // (3) hand the prefetched rows over; narrow segments (fewer than 16 warps) fetch the rest without overlap
        if (more) {
            me_task_store(g, L, s_win, s_cur_next, warp, mby + 3, pre0, m0, nmbs, lane);
            me_task_store(g, L, s_win, s_cur_next, warp + nwarps, mby + 3, pre1, m0, nmbs, lane);
            for (int task = warp + 2 * nwarps; task < 32; task += nwarps)
                me_task_store(g, L, s_win, s_cur_next, task, mby + 3, me_task_load(g, refy, cury, task, mby + 3, mby + 1, m0, nmbs, lane), m0, nmbs, lane);
        }
        __syncthreads();
    }
    if (!fused) return;
    const unsigned breaks = *s_breaks_p;        // the loop's last __syncthreads() ordered every increment before this read
    if (threadIdx.x == 0) p.meflag[gop] = breaks;
    if (breaks == 0) return;                    // natural content: the speculative state-0 search is the answer
    me_zero_rows(g, L, p, st, gop, seg, s_me);  // which visits have SAD 0, for all 8 start states
    __syncthreads();
    me_chain_walk_cta(g, p, gop, s_me, me_smem_bytes(L));   // start state of every macroblock
    __syncthreads();
    me_search_rows(g, L, p, st, gop, seg, 1, 0, g.mbh, s_me);   // re-search the macroblocks whose start state is not 0
}

// ---- unit shims ------------------------------------------------------------------------------------
__global__ void dct8x8_kernel(const int32_t* in, double* out, int n)
{
    __shared__ double s_tile[16][72];
    const int grp = threadIdx.x >> 3, r = threadIdx.x & 7;
    const int blk = blockIdx.x * 16 + grp;
    const int bsafe = min(blk, n - 1);
    int e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) e[i] = in[(size_t)bsafe * 64 + r * 8 + i];
    double t[8], D[8];
    const Mags M = load_mags<0>();
    fdct_row(e, t, M);
    group_transpose(t, s_tile[grp], r);
    fdct_col(t, r, D, M);
    if (blk < n)
#pragma unroll
        for (int v = 0; v < 8; v++) out[(size_t)blk * 64 + v * 8 + r] = D[v];
}
template <int TAB>
__global__ void idct8x8_kernel(const int32_t* in, double* out, int n)
{
    __shared__ double s_tile[16][72];
    const int grp = threadIdx.x >> 3, r = threadIdx.x & 7;
    const int blk = blockIdx.x * 16 + grp;
    const int bsafe = min(blk, n - 1);
    int q[8];
#pragma unroll
    for (int i = 0; i < 8; i++) q[i] = in[(size_t)bsafe * 64 + r * 8 + i];
    double t[8], R[8];
    const Mags M = load_mags<TAB>();
    idct_row<TAB>(q, t, M);
    group_transpose(t, s_tile[grp], r);
    idct_col<TAB>(t, R, M);
    if (blk < n)
#pragma unroll
        for (int y = 0; y < 8; y++) out[(size_t)blk * 64 + y * 8 + r] = R[y];
}

// Quantization_block (ENC:2750-2796) / CQuantization_block (ENC:4610-4656) on n stand-alone blocks: raster order in and out,
// QstepDC at [0][0], QstepAC elsewhere, luma truncates and chroma floors before the integer division; ACflag = all 63 AC == 0
__global__ void __launch_bounds__(256) quant8x8_kernel(const double* in, int32_t* out, uint8_t* acflag, int n, unsigned magic_dc, unsigned magic_ac, int chroma)
{
    __shared__ int s_nz[4];
    const int b = threadIdx.x >> 6, i = threadIdx.x & 63, blk = blockIdx.x * 4 + b;
    if (i == 0) s_nz[b] = 0;
    __syncthreads();
    if (blk < n) {
        const int L = quant_magic(in[(size_t)blk * 64 + i], i == 0 ? magic_dc : magic_ac, chroma != 0);
        out[(size_t)blk * 64 + i] = L;
        if (i != 0 && L != 0) atomicOr(&s_nz[b], 1);
    }
    __syncthreads();
    if (blk < n && i == 0 && acflag) acflag[blk] = s_nz[b] ? 0 : 1;
}

// decoder: motion vector reconstruction mv = mvd + Pm over already reconstructed neighbours (DEC:4301-4370).
// One warp per frame, waves w = mbx + 2*mby (L, U, UL, UR dependencies).  `staged`: the differentials are copied into shared
// memory first and turned into vectors in place, so that a wave costs a shared-memory round trip instead of a global one
// (56 dependent waves per CIF frame: 65 -> ~8 us per launch); frames too large for that work on the global array.
__global__ void __launch_bounds__(32) mv_recon_kernel(Geom g, FramePtrs p, Step st, int staged)
{
    extern __shared__ __align__(16) unsigned char s_mvr[];
    const int gop = blockIdx.x, lane = threadIdx.x;
    const size_t f = (size_t)gop * st.gop_len + st.t;
    const int16_t* mvd = p.mvd + f * g.nmb * 2;
    int16_t* mv = p.mv + f * g.nmb * 2;
    int16_t* w = staged ? (int16_t*)s_mvr : mv;
    if (staged) {
        for (int i = lane; i < g.nmb; i += 32) ((int*)w)[i] = ((const int*)mvd)[i];
        __syncwarp();
    }
    const int nwaves = (g.mbw - 1) + 2 * (g.mbh - 1) + 1;
    for (int wv = 0; wv < nwaves; wv++) {
        const int lo = max(0, (wv - (g.mbw - 1) + 1) >> 1), hi = min(g.mbh - 1, wv >> 1);
        for (int mby = lo + lane; mby <= hi; mby += 32) {
            const int mb = mby * g.mbw + (wv - 2 * mby);
            int px, py;
            mv_predictor(w, g.mbw, mb, px, py);
            const int dx = staged ? w[2 * mb] : mvd[2 * mb], dy = staged ? w[2 * mb + 1] : mvd[2 * mb + 1];
            w[2 * mb] = (int16_t)(dx + px);
            w[2 * mb + 1] = (int16_t)(dy + py);
        }
        __syncwarp();
    }
    if (staged)
        for (int i = lane; i < g.nmb; i += 32) ((int*)mv)[i] = ((const int*)w)[i];
}

// ---- per-plane sum of squared errors between the source frames and the reconstruction (DEC.h:332-346 computes the luma
// MSE in double; every term and partial sum is an integer < 2^53, so the integer sum is the same number).
// grid (3 planes, frames), 256 threads; 16-byte loads, packed |a-b| then dp4a(d, d).  HBM-bound: 2 bytes read per pixel.
__global__ void __launch_bounds__(256) plane_sse_kernel(Geom g, const uint8_t* __restrict__ cur, const uint8_t* __restrict__ rec,
                                                        unsigned long long* __restrict__ sse)
{
    __shared__ unsigned long long s_part[8];
    const int plane = blockIdx.x;
    const size_t f = blockIdx.y;
    const size_t ysz = (size_t)g.w * g.h, csz = ysz / 4;
    const size_t off = f * g.fb + (plane == 0 ? 0 : plane == 1 ? ysz : ysz + csz);
    const int n16 = (int)((plane == 0 ? ysz : csz) / 16);
    const uint4* a = (const uint4*)(cur + off);
    const uint4* b = (const uint4*)(rec + off);
    unsigned long long acc = 0;
    for (int i = threadIdx.x; i < n16; i += 256) {
        const uint4 x = __ldg(a + i), y = __ldg(b + i);
        const uint32_t d0 = __vabsdiffu4(x.x, y.x), d1 = __vabsdiffu4(x.y, y.y), d2 = __vabsdiffu4(x.z, y.z), d3 = __vabsdiffu4(x.w, y.w);
        uint32_t t = __dp4a(d0, d0, 0u);
        t = __dp4a(d1, d1, t); t = __dp4a(d2, d2, t); t = __dp4a(d3, d3, t);   // <= 16 * 255^2 per iteration
        acc += t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; w++) t += s_part[w];
        sse[f * 3 + plane] = t;
    }
}

}  // namespace icsp
