/* TEST INFRASTRUCTURE — see icsp_oracle.h.  Plain-C restatement of the ICSPCodec hot path.
 * Compile WITHOUT FMA contraction (-ffp-contract=off, no -march=native): every product that is not
 * exactly representable must be rounded before the add.
 *
 * Citation key: ENC = /root/reference/source/encoder/ICSP_Codec_Encoder_source.cpp,
 *               ENC.h = .../ICSP_Codec_Encoder.h, DEC = /root/reference/source/decoder/ICSP_Codec_Decoder_source.cpp,
 *               DEC.h = .../ICSP_Codec_Decoder.h
 */
#include "icsp_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <limits.h>

/* ---- constants (ENC.h:190-199, DEC.h:19-28) ------------------------------------------------------ */
static const double LIT[8] = {1.0, 0.980785, 0.92388, 0.83147, 0.707107, 0.55557, 0.382683, 0.19509};
/* costable[u][x] = sign * LIT[idx]; layout of ENC.h:191-198 */
static const signed char TIDX[8][8] = {
    {0, 0, 0, 0, 0, 0, 0, 0}, {1, 3, 5, 7, 7, 5, 3, 1}, {2, 6, 6, 2, 2, 6, 6, 2}, {3, 7, 1, 5, 5, 1, 7, 3},
    {4, 4, 4, 4, 4, 4, 4, 4}, {5, 1, 7, 3, 3, 7, 1, 5}, {6, 2, 2, 6, 6, 2, 2, 6}, {7, 5, 3, 1, 1, 3, 5, 7}};
static const signed char TSGN[8][8] = {
    {1, 1, 1, 1, 1, 1, 1, 1},   {1, 1, 1, 1, -1, -1, -1, -1}, {1, 1, -1, -1, -1, -1, 1, 1}, {1, -1, -1, -1, 1, 1, 1, -1},
    {1, -1, -1, 1, 1, -1, -1, 1}, {1, -1, 1, 1, -1, -1, 1, -1}, {1, -1, 1, -1, -1, 1, -1, 1}, {1, -1, 1, -1, 1, -1, 1, -1}};
static double TAB[2][8][8]; /* [0] encoder: binary32 widened; [1] decoder: binary64 */
static double IRT2;
static int tables_ready = 0;
/* zig-zag: ZZ[k] = raster index (y*8+x) of the k-th coded coefficient (ENC:3031-3094) */
static const unsigned char ZZ[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                     41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                     30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

static void init_tables(void)
{
    if (tables_ready) return;
    for (int u = 0; u < 8; u++)
        for (int x = 0; x < 8; x++) {
            double lit = LIT[TIDX[u][x]] * TSGN[u][x];
            TAB[0][u][x] = (double)(float)lit;
            TAB[1][u][x] = lit;
        }
    IRT2 = 1.0 / sqrt(2.0); /* ENC.h:199 */
    tables_ready = 1;
}

static inline int med3(int a, int b, int c)
{ /* e.g. ENC:3677-3679 */
    if (a > b && a > c) return b > c ? b : c;
    if (b > a && b > c) return a > c ? a : c;
    return a > b ? a : b;
}
static inline int clip255(int v) { return v > 255 ? 255 : (v < 0 ? 0 : v); }

/* ---- transforms -------------------------------------------------------------------------------- */
/* ENC:2685-2749 (DCT_block), ENC:4338-4419 (CDCT_block) */
static void fdct(const int* E, double* D)
{
    const double(*T)[8] = TAB[0];
    double t[8][8];
    for (int v = 0; v < 8; v++)
        for (int u = 0; u < 8; u++) {
            double s = 0;
            for (int x = 0; x < 8; x++) s += (double)E[v * 8 + x] * T[u][x];
            t[v][u] = s;
        }
    for (int u = 0; u < 8; u++)
        for (int v = 0; v < 8; v++) {
            double s = 0;
            for (int y = 0; y < 8; y++) s += t[y][u] * T[v][y];
            D[v * 8 + u] = s;
        }
    for (int i = 0; i < 8; i++) {
        D[0 * 8 + i] *= IRT2;
        D[i * 8 + 0] *= IRT2;
    }
    for (int i = 0; i < 64; i++) D[i] *= (1. / 4.);
}
/* ENC:2825-2893 (IDCT_block), ENC:4687-4768; DEC:3331-3445, DEC:4220-4300 with the binary64 table */
static void idct(const int* Q, double* R, int table)
{
    const double(*T)[8] = TAB[table];
    double C[8], t[8][8];
    C[0] = IRT2;
    for (int i = 1; i < 8; i++) C[i] = 1.;
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) {
            double s = 0;
            for (int u = 0; u < 8; u++) s += C[u] * (double)Q[y * 8 + u] * T[u][x];
            t[y][x] = s;
        }
    for (int x = 0; x < 8; x++)
        for (int y = 0; y < 8; y++) {
            double s = 0;
            for (int v = 0; v < 8; v++) s += C[v] * t[v][x] * T[v][y];
            R[y * 8 + x] = s;
        }
    for (int i = 0; i < 64; i++) R[i] *= (1. / 4.);
}

void icsp_oracle_dct8x8(const int32_t* in, double* out, int nblocks)
{
    init_tables();
    for (int b = 0; b < nblocks; b++) fdct(in + 64 * b, out + 64 * b);
}
void icsp_oracle_idct8x8(const int32_t* in, double* out, int nblocks, int table)
{
    init_tables();
    for (int b = 0; b < nblocks; b++) idct(in + 64 * b, out + 64 * b, table);
}

/* ---- geometry / per-frame state ---------------------------------------------------------------- */
typedef struct {
    int w, h, mbw, mbh, nmb, cw, ch, fb;
} geom_t;
static int geom_init(geom_t* g, int w, int h)
{
    if (w <= 0 || h <= 0 || (w & 15) || (h & 15)) return -1;
    g->w = w; g->h = h; g->mbw = w / 16; g->mbh = h / 16; g->nmb = g->mbw * g->mbh;
    g->cw = w / 2; g->ch = h / 2; g->fb = w * h * 3 / 2;
    return 0;
}

/* R1: getPaddingImage (ENC:2227-2269): clamp, except that the last padded row and column stay 0 */
static inline int ref_px(const uint8_t* P, int w, int h, int pad, int y, int x)
{
    if (y == h + 2 * pad - 1 || x == w + 2 * pad - 1) return 0;
    int yy = y - pad, xx = x - pad;
    yy = yy < 0 ? 0 : (yy > h - 1 ? h - 1 : yy);
    xx = xx < 0 ? 0 : (xx > w - 1 ? w - 1 : xx);
    return P[yy * w + xx];
}

/* A.5 luma DC predictor (ENC:3643-3990 / 3991-4337; DEC:2984-3330) on the 8x8 grid */
static int dc_pred_luma(const int* dc, int bw, int bx, int by)
{
    if (bx == 0 && by == 0) return 1024;
    if (by == 0) return dc[by * bw + bx - 1];
    if (bx == 0) return dc[(by - 1) * bw + bx];
    int L = dc[by * bw + bx - 1], U = dc[(by - 1) * bw + bx], UL = dc[(by - 1) * bw + bx - 1];
    if ((bx & 1) == 0 || ((by & 1) == 0 && bx != bw - 1)) return med3(L, U, dc[(by - 1) * bw + bx + 1]);
    return med3(L, UL, U);
}
/* chroma (ENC:4420-4514 / 4515-4609; DEC:4124-4219) */
static int dc_pred_chroma(const int* dc, int bw, int bx, int by)
{
    if (bx == 0 && by == 0) return 1024;
    if (by == 0) return dc[by * bw + bx - 1];
    if (bx == 0) return dc[(by - 1) * bw + bx];
    int L = dc[by * bw + bx - 1], U = dc[(by - 1) * bw + bx], UL = dc[(by - 1) * bw + bx - 1];
    if (bx == bw - 1) return med3(L, UL, U);
    return med3(L, U, dc[(by - 1) * bw + bx + 1]);
}

/* A.6: code one 8x8 block.  E residual (raster), P DC predictor; writes zig-zag levels (DC differential),
 * acflag, returns reconstructed DC; R = IDCT output (encoder table). */
static int code_block(const int* E, int P, int qdc, int qac, int chroma, int16_t* zz, uint8_t* acflag, double* R,
                      double* dct_tap)
{
    double D[64];
    int L[64], IQ[64];
    fdct(E, D);
    if (dct_tap) memcpy(dct_tap, D, sizeof(D));
    D[0] -= P; /* DPCM_DC_block: double - int, before quantisation */
    for (int i = 0; i < 64; i++) {
        int q = i == 0 ? qdc : qac;
        /* luma (ENC:2780): (int)(D+0.5)/Q ; chroma (ENC:4642): (int)floor(D+0.5)/Q */
        int r = chroma ? (int)floor(D[i] + 0.5) : (int)(D[i] + 0.5);
        L[i] = r / q;
    }
    int flag = 1;
    for (int i = 1; i < 64; i++)
        if (L[i] != 0) { flag = 0; break; }
    *acflag = (uint8_t)flag;
    for (int k = 0; k < 64; k++) zz[k] = (int16_t)L[ZZ[k]];
    for (int i = 0; i < 64; i++) IQ[i] = L[i] * (i == 0 ? qdc : qac);
    IQ[0] += P;
    idct(IQ, R, 0);
    return IQ[0];
}

/* ---- intra frame (A.7; ENC:556-643, 851-1875, 1876-1983) -------------------------------------- */
static void intra_mode_signal(int mode, int hasL, int hasU, const uint8_t* modes, int bw, int bx, int by, uint8_t* mpm,
                              uint8_t* ipm)
{
    *mpm = 0; *ipm = 0;
    if (!hasL && !hasU) return; /* block (0,0): mode 2, MPM=0, bit=0 (ENC:886-891) */
    int p;
    if (hasL && hasU) p = med3(modes[by * bw + bx - 1], modes[(by - 1) * bw + bx - 1], modes[(by - 1) * bw + bx]);
    else if (hasL) p = modes[by * bw + bx - 1];
    else p = modes[(by - 1) * bw + bx];
    if (mode == p) { *mpm = 1; return; }
    if (p == 0) *ipm = (mode == 1) ? 0 : 1;
    else *ipm = (mode == 0) ? 0 : 1; /* p==2 and p==1 (ENC:1342-1350) */
}
static int intra_mode_decode(int mpm, int ipm, int hasL, int hasU, const uint8_t* modes, int bw, int bx, int by)
{ /* ENC:1793-1795 / DEC:3446-3821 */
    if (!hasL && !hasU) return 2;
    int p;
    if (hasL && hasU) p = med3(modes[by * bw + bx - 1], modes[(by - 1) * bw + bx - 1], modes[(by - 1) * bw + bx]);
    else if (hasL) p = modes[by * bw + bx - 1];
    else p = modes[(by - 1) * bw + bx];
    if (mpm) return p;
    if (p == 0) return ipm == 0 ? 1 : 2;
    if (p == 2) return ipm == 0 ? 0 : 1;
    return ipm == 0 ? 0 : 2;
}

/* recon one luma intra block from R (IDCT), given mode; neighbours from the recon plane */
static void intra_recon_block(uint8_t* rec, int w, int bx, int by, int mode, const double* R)
{
    int hasL = bx > 0, hasU = by > 0;
    uint8_t* o = rec + (by * 8) * w + bx * 8;
    if (mode == 0) {
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) {
                int p = hasU ? o[-w + x] : 128;
                int t = (int)(R[y * 8 + x] + p); /* double sum, then truncation (ENC:767-769) */
                o[y * w + x] = (uint8_t)clip255(t);
            }
    } else if (mode == 1) {
        int left[8];
        for (int y = 0; y < 8; y++) left[y] = hasL ? o[y * w - 1] : 128;
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) o[y * w + x] = (uint8_t)clip255((int)(R[y * 8 + x] + left[y]));
    } else {
        double sl = 0, su = 0;
        if (!hasL) sl = 128 * 8;
        else for (int i = 0; i < 8; i++) sl += o[i * w - 1];
        if (!hasU) su = 128 * 8;
        else for (int i = 0; i < 8; i++) su += o[-w + i];
        double pd = (sl + su) / 16;
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) o[y * w + x] = (uint8_t)clip255((int)(R[y * 8 + x] + pd));
    }
}

typedef struct {
    int *dcY, *dcCb, *dcCr;
    uint8_t* modes;
} fstate_t;

static void encode_intra_frame(const geom_t* g, const uint8_t* cur, int qdc, int qac, int16_t* levels, uint8_t* acflag,
                               uint8_t* mpm, uint8_t* ipm, uint8_t* rec, fstate_t* st, double* dct_tap)
{
    const int w = g->w, bw = g->mbw * 2;
    const uint8_t* cy = cur;
    uint8_t* ry = rec;
    for (int mb = 0; mb < g->nmb; mb++) {
        int mbx = mb % g->mbw, mby = mb / g->mbw;
        for (int k = 0; k < 4; k++) {
            int bx = 2 * mbx + (k & 1), by = 2 * mby + (k >> 1);
            int hasL = bx > 0, hasU = by > 0;
            const uint8_t* c = cy + (by * 8) * w + bx * 8;
            const uint8_t* r = ry + (by * 8) * w + bx * 8;
            int E0[64], E1[64], E2[64], sae0 = 0, sae1 = 0, sae2 = 0;
            double sl = 0, su = 0;
            if (!hasL) sl = 128 * 8; else for (int i = 0; i < 8; i++) sl += r[i * w - 1];
            if (!hasU) su = 128 * 8; else for (int i = 0; i < 8; i++) su += r[-w + i];
            double pd = (sl + su) / (double)16;
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++) {
                    int cv = c[y * w + x];
                    E0[y * 8 + x] = cv - (hasU ? r[-w + x] : 128);       /* DPCM_pix_0 ENC:644-674 */
                    E1[y * 8 + x] = cv - (hasL ? r[y * w - 1] : 128);    /* DPCM_pix_1 ENC:675-705 */
                    E2[y * 8 + x] = (int)(cv - pd);                      /* DPCM_pix_2 ENC:706-743 */
                    sae0 += abs(E0[y * 8 + x]); sae1 += abs(E1[y * 8 + x]); sae2 += abs(E2[y * 8 + x]);
                }
            int mode;
            if (hasL && hasU) {
                int m = sae0 < sae1 ? sae0 : sae1; m = m < sae2 ? m : sae2;
                mode = (m == sae0) ? 0 : (m == sae1 ? 1 : 2);     /* ENC:958-976 */
            } else if (hasL) mode = (sae2 > sae1) ? 1 : 2;        /* ENC:897-908 */
            else if (hasU) mode = (sae2 > sae0) ? 0 : 2;
            else mode = 2;
            st->modes[by * bw + bx] = (uint8_t)mode;
            intra_mode_signal(mode, hasL, hasU, st->modes, bw, bx, by, &mpm[mb * 4 + k], &ipm[mb * 4 + k]);
            const int* E = mode == 0 ? E0 : (mode == 1 ? E1 : E2);
            double R[64];
            int P = dc_pred_luma(st->dcY, bw, bx, by);
            st->dcY[by * bw + bx] = code_block(E, P, qdc, qac, 0, levels + (mb * 6 + k) * 64, &acflag[mb * 6 + k], R,
                                               dct_tap ? dct_tap + (mb * 6 + k) * 64 : NULL);
            intra_recon_block(ry, w, bx, by, mode, R);
        }
        /* intraCbCr (ENC:1876-1903): no pixel prediction, DCT of raw pixels */
        for (int c = 0; c < 2; c++) {
            const uint8_t* cp = cur + g->w * g->h + c * g->cw * g->ch;
            uint8_t* rp = rec + g->w * g->h + c * g->cw * g->ch;
            int* dc = c ? st->dcCr : st->dcCb;
            int E[64];
            double R[64];
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++) E[y * 8 + x] = cp[(mby * 8 + y) * g->cw + mbx * 8 + x];
            int P = dc_pred_chroma(dc, g->mbw, mbx, mby);
            dc[mby * g->mbw + mbx] = code_block(E, P, qdc, qac, 1, levels + (mb * 6 + 4 + c) * 64, &acflag[mb * 6 + 4 + c],
                                                R, dct_tap ? dct_tap + (mb * 6 + 4 + c) * 64 : NULL);
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++) { /* intraImgReconstruct ENC:1964-1971 */
                    double d = R[y * 8 + x];
                    int t = (int)((d > 255) ? 255 : d);
                    t = t < 0 ? 0 : t;
                    rp[(mby * 8 + y) * g->cw + mbx * 8 + x] = (uint8_t)t;
                }
        }
    }
}

/* ---- motion estimation (A.8; ENC:2073-2155) ---------------------------------------------------- */
void icsp_oracle_me(const uint8_t* cur_y, const uint8_t* ref_y, int w, int h, int16_t* mv, int32_t* minsad,
                    int32_t* n_sad_evals)
{
    int mbw = w / 16, nmb = mbw * (h / 16);
    int flag = 0, xflag = 1, yflag = -1; /* frame-level state, never reset per MB (ENC:2094-2095) */
    int evals = 0;
    for (int mb = 0; mb < nmb; mb++) {
        int X0 = (mb % mbw) * 16, Y0 = (mb / mbw) * 16;
        int x0 = X0, y0 = Y0, xcnt = 0, ycnt = 0, cnt = 0, min = INT_MAX, tx = 0, ty = 0;
        while (cnt < 64) {
            if (!flag) {
                if (xflag <= 0) x0 += xcnt; else x0 -= xcnt;
                flag = 1; xcnt++; xflag *= -1;
            } else {
                if (yflag < 0) y0 += ycnt; else y0 -= ycnt;
                flag = 0; ycnt++; yflag *= -1;
            }
            int sad = 0;
            for (int j = 0; j < 16; j++)
                for (int i = 0; i < 16; i++)
                    sad += abs((int)cur_y[(Y0 + j) * w + X0 + i] - ref_px(ref_y, w, h, 16, 16 + y0 + j, 16 + x0 + i));
            evals++;
            if (min > sad) { min = sad; tx = x0; ty = y0; }
            else if (sad == 0) { tx = x0; ty = y0; break; }
            cnt++;
        }
        mv[2 * mb] = (int16_t)(X0 - tx);
        mv[2 * mb + 1] = (int16_t)(Y0 - ty);
        if (minsad) minsad[mb] = min;
    }
    if (n_sad_evals) *n_sad_evals = evals;
}

/* MV predictor (ENC:2353-2425 / 2426-2499; DEC:4301-4370) from full MVs of already visited MBs */
static void mv_pred(const int16_t* mv, int mbw, int mb, int* px, int* py)
{
    int mbx = mb % mbw;
    if (mb == 0) { *px = 8; *py = 8; return; }
    if (mb / mbw == 0) { *px = mv[2 * (mb - 1)]; *py = mv[2 * (mb - 1) + 1]; return; }
    if (mbx == 0) { *px = mv[2 * (mb - mbw)]; *py = mv[2 * (mb - mbw) + 1]; return; }
    int a, b, c;
    if (mbx == mbw - 1) { a = mb - 1; b = mb - mbw - 1; c = mb - mbw; }   /* L, UL, U */
    else { a = mb - 1; b = mb - mbw; c = mb - mbw + 1; }                   /* L, U, UR */
    int x1 = mv[2 * a], x2 = mv[2 * b], x3 = mv[2 * c];
    int y1 = mv[2 * a + 1], y2 = mv[2 * b + 1], y3 = mv[2 * c + 1];
    *px = med3(x1, x2, x3);
    if (y1 > y2 && y1 > y3) *py = y2 > y3 ? y2 : y3;
    else if (y2 > y1 && y2 > y3) *py = (y1 > x3) ? y1 : y3; /* reference typo kept (ENC:2399,2418) */
    else *py = y1 > y2 ? y1 : y2;
}

/* ---- inter frame (A.8; ENC:1986-2072, 2156-2226, 2298-2352, 2500-2682) ------------------------- */
static void encode_inter_frame(const geom_t* g, const uint8_t* cur, const uint8_t* prev, int qdc, int qac, int16_t* levels,
                               uint8_t* acflag, int16_t* mvd, int16_t* mv, int32_t* minsad, uint8_t* rec, fstate_t* st,
                               double* dct_tap)
{
    const int w = g->w, h = g->h, bw = g->mbw * 2;
    icsp_oracle_me(cur, prev, w, h, mv, minsad, NULL);
    for (int mb = 0; mb < g->nmb; mb++) {
        int mbx = mb % g->mbw, mby = mb / g->mbw, px, py;
        mv_pred(mv, g->mbw, mb, &px, &py);
        mvd[2 * mb] = (int16_t)(mv[2 * mb] - px);
        mvd[2 * mb + 1] = (int16_t)(mv[2 * mb + 1] - py);
        int X0 = mbx * 16, Y0 = mby * 16, mx = mv[2 * mb], my = mv[2 * mb + 1];
        for (int k = 0; k < 4; k++) {
            int bx = 2 * mbx + (k & 1), by = 2 * mby + (k >> 1), E[64], pred[64];
            double R[64];
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++) {
                    int yy = (k >> 1) * 8 + y, xx = (k & 1) * 8 + x;
                    pred[y * 8 + x] = ref_px(prev, w, h, 16, 16 + Y0 - my + yy, 16 + X0 - mx + xx);
                    E[y * 8 + x] = (int)cur[(Y0 + yy) * w + X0 + xx] - pred[y * 8 + x];
                }
            int P = dc_pred_luma(st->dcY, bw, bx, by);
            st->dcY[by * bw + bx] = code_block(E, P, qdc, qac, 0, levels + (mb * 6 + k) * 64, &acflag[mb * 6 + k], R,
                                               dct_tap ? dct_tap + (mb * 6 + k) * 64 : NULL);
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++) { /* mergeBlock ENC:4812 truncates first; interYReconstruct ENC:2343-2346 */
                    int t = pred[y * 8 + x] + (int)R[y * 8 + x];
                    rec[(by * 8 + y) * w + bx * 8 + x] = (uint8_t)clip255(t);
                }
        }
    }
    for (int mb = 0; mb < g->nmb; mb++) { /* interCbCr ENC:2625-2682 */
        int mbx = mb % g->mbw, mby = mb / g->mbw;
        int cmx = mv[2 * mb] / 2, cmy = mv[2 * mb + 1] / 2; /* C division, toward zero (ENC:2538-2539) */
        for (int c = 0; c < 2; c++) {
            const uint8_t* cp = cur + w * h + c * g->cw * g->ch;
            const uint8_t* pp = prev + w * h + c * g->cw * g->ch;
            uint8_t* rp = rec + w * h + c * g->cw * g->ch;
            int* dc = c ? st->dcCr : st->dcCb;
            int E[64], pred[64];
            double R[64];
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++) {
                    pred[y * 8 + x] = ref_px(pp, g->cw, g->ch, 8, 8 + mby * 8 - cmy + y, 8 + mbx * 8 - cmx + x);
                    E[y * 8 + x] = (int)cp[(mby * 8 + y) * g->cw + mbx * 8 + x] - pred[y * 8 + x];
                }
            int P = dc_pred_chroma(dc, g->mbw, mbx, mby);
            dc[mby * g->mbw + mbx] = code_block(E, P, qdc, qac, 1, levels + (mb * 6 + 4 + c) * 64, &acflag[mb * 6 + 4 + c],
                                                R, dct_tap ? dct_tap + (mb * 6 + 4 + c) * 64 : NULL);
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++) { /* interCbCrReconstruct ENC:2605-2607: truncation of the double sum */
                    int t = (int)(pred[y * 8 + x] + R[y * 8 + x]);
                    rp[(mby * 8 + y) * g->cw + mbx * 8 + x] = (uint8_t)clip255(t);
                }
        }
    }
}

static int fstate_alloc(fstate_t* st, const geom_t* g)
{
    st->dcY = (int*)calloc((size_t)g->nmb * 4, sizeof(int));
    st->dcCb = (int*)calloc((size_t)g->nmb, sizeof(int));
    st->dcCr = (int*)calloc((size_t)g->nmb, sizeof(int));
    st->modes = (uint8_t*)calloc((size_t)g->nmb * 4, 1);
    return (st->dcY && st->dcCb && st->dcCr && st->modes) ? 0 : -1;
}
static void fstate_free(fstate_t* st) { free(st->dcY); free(st->dcCb); free(st->dcCr); free(st->modes); }

int icsp_oracle_encode(const uint8_t* frames, int nframes, int w, int h, int qdc, int qac, int intra_period,
                       int16_t* levels, uint8_t* acflag, uint8_t* mpm, uint8_t* ipm, int16_t* mvd, int16_t* mv,
                       int32_t* minsad, uint8_t* recon, double* dct_tap)
{
    geom_t g;
    fstate_t st;
    if (geom_init(&g, w, h) || qdc <= 0 || qac <= 0 || intra_period < 0) return -1;
    if (fstate_alloc(&st, &g)) return -1;
    init_tables();
    for (int n = 0; n < nframes; n++) {
        const uint8_t* cur = frames + (size_t)n * g.fb;
        uint8_t* rec = recon + (size_t)n * g.fb;
        int16_t* lv = levels + (size_t)n * g.nmb * 384;
        uint8_t* af = acflag + (size_t)n * g.nmb * 6;
        uint8_t* pm = mpm + (size_t)n * g.nmb * 4;
        uint8_t* im = ipm + (size_t)n * g.nmb * 4;
        int16_t* md = mvd + (size_t)n * g.nmb * 2;
        int16_t* mf = mv + (size_t)n * g.nmb * 2;
        int32_t* ms = minsad + (size_t)n * g.nmb;
        double* tap = dct_tap ? dct_tap + (size_t)n * g.nmb * 384 : NULL;
        memset(pm, 0, (size_t)g.nmb * 4); memset(im, 0, (size_t)g.nmb * 4);
        memset(md, 0, (size_t)g.nmb * 4); memset(mf, 0, (size_t)g.nmb * 4); memset(ms, 0, (size_t)g.nmb * 4);
        if (intra_period == 0 || n % intra_period == 0)
            encode_intra_frame(&g, cur, qdc, qac, lv, af, pm, im, rec, &st, tap);
        else
            encode_inter_frame(&g, cur, recon + (size_t)(n - 1) * g.fb, qdc, qac, lv, af, md, mf, ms, rec, &st, tap);
    }
    fstate_free(&st);
    return 0;
}

/* ---- decoder reconstruction (A.9; DEC:2083-2272, 2654-2764, 2929-4475) ------------------------- */
static int decode_block(const int16_t* zz, int P, int qdc, int qac, double* R)
{
    int IQ[64];
    for (int k = 0; k < 64; k++) IQ[ZZ[k]] = zz[k] * (ZZ[k] == 0 ? qdc : qac); /* izigzag DEC:2829, IQuant DEC:2929 */
    IQ[0] += P;                                                               /* IDPCM_DC_block DEC:2984 */
    idct(IQ, R, 1);                                                           /* binary64 table DEC.h:19 */
    return IQ[0];
}

int icsp_oracle_decode(const int16_t* levels, const uint8_t* mpm, const uint8_t* ipm, const int16_t* mvd, int nframes,
                       int w, int h, int qdc, int qac, int intra_period, uint8_t* yuv_out)
{
    geom_t g;
    fstate_t st;
    if (geom_init(&g, w, h) || intra_period < 1) return -1;
    if (fstate_alloc(&st, &g)) return -1;
    init_tables();
    int16_t* mv = (int16_t*)calloc((size_t)g.nmb * 2, sizeof(int16_t));
    const int bw = g.mbw * 2;
    for (int n = 0; n < nframes; n++) {
        const int16_t* lv = levels + (size_t)n * g.nmb * 384;
        uint8_t* rec = yuv_out + (size_t)n * g.fb;
        int is_intra = (intra_period == 1) || (n % intra_period == 0);
        double R[64];
        if (is_intra) {
            for (int mb = 0; mb < g.nmb; mb++) {
                int mbx = mb % g.mbw, mby = mb / g.mbw;
                for (int k = 0; k < 4; k++) {
                    int bx = 2 * mbx + (k & 1), by = 2 * mby + (k >> 1);
                    int P = dc_pred_luma(st.dcY, bw, bx, by);
                    st.dcY[by * bw + bx] = decode_block(lv + (mb * 6 + k) * 64, P, qdc, qac, R);
                    int mode = intra_mode_decode(mpm[((size_t)n * g.nmb + mb) * 4 + k], ipm[((size_t)n * g.nmb + mb) * 4 + k],
                                                 bx > 0, by > 0, st.modes, bw, bx, by);
                    st.modes[by * bw + bx] = (uint8_t)mode;
                    intra_recon_block(rec, w, bx, by, mode, R);
                }
                for (int c = 0; c < 2; c++) {
                    uint8_t* rp = rec + w * h + c * g.cw * g.ch;
                    int* dc = c ? st.dcCr : st.dcCb;
                    int P = dc_pred_chroma(dc, g.mbw, mbx, mby);
                    dc[mby * g.mbw + mbx] = decode_block(lv + (mb * 6 + 4 + c) * 64, P, qdc, qac, R);
                    for (int y = 0; y < 8; y++)
                        for (int x = 0; x < 8; x++) { /* DEC:4000-4079 */
                            double d = R[y * 8 + x];
                            int t = (int)((d > 255) ? 255 : d);
                            rp[(mby * 8 + y) * g.cw + mbx * 8 + x] = (uint8_t)(t < 0 ? 0 : t);
                        }
                }
            }
        } else {
            const uint8_t* prev = yuv_out + (size_t)(n - 1) * g.fb;
            const int16_t* md = mvd + (size_t)n * g.nmb * 2;
            for (int mb = 0; mb < g.nmb; mb++) { /* ImvPrediction DEC:4301-4370 */
                int px, py;
                mv_pred(mv, g.mbw, mb, &px, &py);
                mv[2 * mb] = (int16_t)(md[2 * mb] + px);
                mv[2 * mb + 1] = (int16_t)(md[2 * mb + 1] + py);
            }
            for (int mb = 0; mb < g.nmb; mb++) {
                int mbx = mb % g.mbw, mby = mb / g.mbw, mx = mv[2 * mb], my = mv[2 * mb + 1];
                for (int k = 0; k < 4; k++) {
                    int bx = 2 * mbx + (k & 1), by = 2 * mby + (k >> 1);
                    int P = dc_pred_luma(st.dcY, bw, bx, by);
                    st.dcY[by * bw + bx] = decode_block(lv + (mb * 6 + k) * 64, P, qdc, qac, R);
                    for (int y = 0; y < 8; y++)
                        for (int x = 0; x < 8; x++) { /* DEC:3923 (int)idct, DEC:4371 clip(pred+err) */
                            int pr = ref_px(prev, w, h, 16, 16 + by * 8 - my + y, 16 + bx * 8 - mx + x);
                            rec[(by * 8 + y) * w + bx * 8 + x] = (uint8_t)clip255(pr + (int)R[y * 8 + x]);
                        }
                }
                int cmx = mx / 2, cmy = my / 2;
                for (int c = 0; c < 2; c++) {
                    const uint8_t* pp = prev + w * h + c * g.cw * g.ch;
                    uint8_t* rp = rec + w * h + c * g.cw * g.ch;
                    int* dc = c ? st.dcCr : st.dcCb;
                    int P = dc_pred_chroma(dc, g.mbw, mbx, mby);
                    dc[mby * g.mbw + mbx] = decode_block(lv + (mb * 6 + 4 + c) * 64, P, qdc, qac, R);
                    for (int y = 0; y < 8; y++)
                        for (int x = 0; x < 8; x++) { /* DEC:2699-2764 */
                            int pr = ref_px(pp, g.cw, g.ch, 8, 8 + mby * 8 - cmy + y, 8 + mbx * 8 - cmx + x);
                            rp[(mby * 8 + y) * g.cw + mbx * 8 + x] = (uint8_t)clip255((int)(pr + R[y * 8 + x]));
                        }
                }
            }
        }
    }
    free(mv);
    fstate_free(&st);
    return 0;
}

/* ---- bitstream (A.10 / R14; ENC:4849-6334, DEC:14-2025, 2274-2653) ----------------------------- */
typedef struct {
    uint8_t* buf;
    long cap;      /* bytes */
    long nbits;
    int overflow;
} bitw_t;
static inline void put_bit(bitw_t* b, int bit)
{
    long byte = b->nbits >> 3;
    if (byte >= b->cap) { b->overflow = 1; return; }
    b->buf[byte] = (uint8_t)((b->buf[byte] << 1) | (bit & 1)); /* (frame[cntbits++/8]<<=1) |= bit */
    b->nbits++;
}
static void put_vlc(bitw_t* b, int v)
{ /* DCentropy ENC:5417-5602 == ACentropy ENC:5791-5989 == MVentropy ENC:5990-6334 */
    int s = v >= 0 ? 1 : 0, a = abs(v);
    if (a == 0) { put_bit(b, 0); put_bit(b, 0); return; }
    if (a == 1) { put_bit(b, 0); put_bit(b, 1); put_bit(b, 0); put_bit(b, s); return; }
    int e = 0;
    while (e < 11 && a >= (2 << e)) e++; /* a in [2^e, 2^(e+1)), capped at e = 11 for a >= 2048 */
    if (e <= 4) { /* 011,100,101,110 */
        int code = e + 2;
        put_bit(b, (code >> 2) & 1); put_bit(b, (code >> 1) & 1); put_bit(b, code & 1);
    } else {
        for (int n = 0; n < e - 2; n++) put_bit(b, 1);
        put_bit(b, 0);
    }
    put_bit(b, s);
    int c = a - (1 << e);
    for (int n = e; n > 0; n--) put_bit(b, (c >> (n - 1)) & 1);
}
static void put_block(bitw_t* b, const int16_t* zz, int acflag)
{
    put_vlc(b, zz[0]);
    put_bit(b, acflag);
    if (acflag == 1) for (int n = 0; n < 63; n++) put_bit(b, 0);
    else for (int n = 1; n < 64; n++) put_vlc(b, zz[n]);
}

long icsp_oracle_write_bitstream(const int16_t* levels, const uint8_t* acflag, const uint8_t* mpm, const uint8_t* ipm,
                                 const int16_t* mvd, int nframes, int w, int h, int qdc, int qac, int intra_period,
                                 uint8_t* out, long cap)
{
    geom_t g;
    if (geom_init(&g, w, h) || cap < 15) return -1;
    /* header ENC.h:201-212, ENC:4901-4922 (packed, little endian) */
    memset(out, 0, (size_t)cap);
    out[0] = 0; out[1] = 73; out[2] = 67; out[3] = 83; out[4] = 80;
    out[5] = (uint8_t)(h & 255); out[6] = (uint8_t)(h >> 8);
    out[7] = (uint8_t)(w & 255); out[8] = (uint8_t)(w >> 8);
    out[9] = (uint8_t)qdc; out[10] = (uint8_t)qac; out[11] = 0;
    unsigned outro = ((unsigned)intra_period & 63u) << 7;
    out[12] = (uint8_t)(outro & 255); out[13] = (uint8_t)(outro >> 8);
    bitw_t b = {out + 14, cap - 14, 0, 0};
    for (int n = 0; n < nframes; n++) {
        int is_intra = intra_period == 0 || n % intra_period == 0;
        for (int mb = 0; mb < g.nmb; mb++) {
            size_t m = (size_t)n * g.nmb + mb;
            if (!is_intra) { /* interBody ENC:5132-5236 */
                put_bit(&b, 1);
                put_vlc(&b, mvd[m * 2]);
                put_vlc(&b, mvd[m * 2 + 1]);
            }
            for (int k = 0; k < 6; k++) {
                if (is_intra && k < 4) { put_bit(&b, mpm[m * 4 + k]); put_bit(&b, ipm[m * 4 + k]); } /* ENC:5057-5058 */
                put_block(&b, levels + (m * 6 + k) * 64, acflag[m * 6 + k]);
            }
        }
    }
    if (b.overflow || (b.nbits >> 3) + 1 > b.cap) return -1;
    return 14 + (b.nbits >> 3) + 1; /* fwrite(tempFrame, cntbits/8+1) ENC:4895 */
}

typedef struct {
    const uint8_t* buf;
    long nbits, pos;
} bitr_t;
static inline int get_bit(bitr_t* r)
{
    int v = 0;
    if (r->pos < r->nbits) v = (r->buf[r->pos >> 3] >> (7 - (r->pos & 7))) & 1; /* DEC:68-74 */
    r->pos++;
    return v;
}
static inline int peek_bit(const bitr_t* r, long off)
{
    long p = r->pos + off;
    return p < r->nbits ? (r->buf[p >> 3] >> (7 - (p & 7))) & 1 : 0;
}
static int get_vlc(bitr_t* r)
{ /* DCientropy DEC:407-608 */
    int b0 = peek_bit(r, 0), b1 = peek_bit(r, 1), b2 = peek_bit(r, 2);
    int e, prefix;
    if (b0 == 0 && b1 == 0) { r->pos += 2; return 0; }
    if (b0 == 0 && b1 == 1 && b2 == 0) { int s = peek_bit(r, 3); r->pos += 4; return s ? 1 : -1; }
    if (!(b0 && b1 && b2)) { e = (b0 << 2 | b1 << 1 | b2) - 2; prefix = 3; }
    else {
        int ones = 3;
        while (ones < 10 && peek_bit(r, ones)) ones++;
        if (ones >= 10) return 0; /* no category matches: len 0, val 0 in the reference */
        e = ones + 2; prefix = ones + 1;
    }
    int s = peek_bit(r, prefix), t = 0;
    for (int n = 0; n < e; n++) t = (t << 1) | peek_bit(r, prefix + 1 + n);
    r->pos += prefix + 1 + e;
    int v = (1 << e) + t;
    return s ? v : -v;
}
static void get_block(bitr_t* r, int16_t* zz, uint8_t* acflag)
{
    zz[0] = (int16_t)get_vlc(r);
    int f = get_bit(r);
    *acflag = (uint8_t)f;
    if (f == 1) { r->pos += 63; for (int n = 1; n < 64; n++) zz[n] = 0; }
    else for (int n = 1; n < 64; n++) zz[n] = (int16_t)get_vlc(r);
}

int icsp_oracle_parse_bitstream(const uint8_t* bin, long len, int nframes, int* w, int* h, int* qdc, int* qac,
                                int* intra_period, int16_t* levels, uint8_t* acflag, uint8_t* mpm, uint8_t* ipm,
                                int16_t* mvd)
{
    if (len < 14) return -1;
    int hh = bin[5] | (bin[6] << 8), ww = bin[7] | (bin[8] << 8);
    unsigned outro = bin[12] | (bin[13] << 8);
    int ip = (outro & 0x1F80) >> 7; /* DEC:29 */
    *w = ww; *h = hh; *qdc = bin[9]; *qac = bin[10]; *intra_period = ip;
    geom_t g;
    if (geom_init(&g, ww, hh) || ip < 1) return -1;
    if (!levels) return 0; /* header only */
    bitr_t r = {bin + 14, (len - 14) * 8, 0};
    for (int n = 0; n < nframes; n++) {
        int is_intra = (ip == 1) || (n % ip == 0); /* DEC:98, DEC:201 */
        for (int mb = 0; mb < g.nmb; mb++) {
            size_t m = (size_t)n * g.nmb + mb;
            for (int k = 0; k < 4; k++) mpm[m * 4 + k] = ipm[m * 4 + k] = 0;
            mvd[m * 2] = mvd[m * 2 + 1] = 0;
            if (!is_intra) {
                (void)get_bit(&r); /* MVmodeflag */
                mvd[m * 2] = (int16_t)get_vlc(&r);
                mvd[m * 2 + 1] = (int16_t)get_vlc(&r);
            }
            for (int k = 0; k < 6; k++) {
                if (is_intra && k < 4) { mpm[m * 4 + k] = (uint8_t)get_bit(&r); ipm[m * 4 + k] = (uint8_t)get_bit(&r); }
                get_block(&r, levels + (m * 6 + k) * 64, &acflag[m * 6 + k]);
            }
        }
    }
    return 0;
}
