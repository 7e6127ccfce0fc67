/* TEST INFRASTRUCTURE — CPU restatement ("port") of the ICSPCodec hot path, used ONLY as the parity
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  The product
 * (libicspcuda / icspenc / icspdec) never links, loads or calls anything in this directory.
 *
 * Parity status: PINNED — tests/golden/make_golden.py and make_golden_full.py run the unmodified compiled
 * reference (oracle/_ref, built from /root/reference by oracle/build_ref.sh) on seeded synthetic clips (12-frame
 * cases and the 300-frame BASELINE.json configs) and store its outputs; tests/test_oracle.py requires this port to
 * reproduce them byte for byte: .bin, test_yuv.yuv, decoder YUV, full MVs, DCT/IDCT doubles (0 ulp), and — when
 * oracle/_ref is present — compares against the reference binaries run live (test_oracle_vs_live_reference).
 *
 * SoA layout shared with include/icspcuda.h (nmb = (w/16)*(h/16), fb = w*h*3/2):
 *   levels  int16 [nframes][nmb][6][64]  zig-zag order, blocks Y0..Y3,Cb,Cr, DC already DPCM'd
 *   acflag  uint8 [nframes][nmb][6]      1 iff all 63 AC levels are zero
 *   mpm,ipm uint8 [nframes][nmb][4]      MPMFlag / intraPredMode bit (I-frames; zero on P-frames)
 *   mvd     int16 [nframes][nmb][2]      differential MV (x,y) as coded (P-frames; zero on I-frames)
 *   mv      int16 [nframes][nmb][2]      full MV (= Reconstructedmv)
 *   minsad  int32 [nframes][nmb]         SAD of the selected candidate
 *   recon   uint8 [nframes][fb]          planar I420 reconstruction
 */
#ifndef ICSP_ORACLE_H
#define ICSP_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* frame n is intra iff intra_period==0 || n % intra_period == 0 (ENC:219-241). dct_tap may be NULL;
 * when given it receives the forward DCT output f64 [nframes][nmb][6][64] (raster order inside a block,
 * BEFORE the DC predictor is subtracted). Returns 0, or -1 on bad geometry. */
int icsp_oracle_encode(const uint8_t* frames, int nframes, int w, int h, int qdc, int qac, int intra_period,
                       int16_t* levels, uint8_t* acflag, uint8_t* mpm, uint8_t* ipm, int16_t* mvd, int16_t* mv,
                       int32_t* minsad, uint8_t* recon, double* dct_tap);

/* Decoder reconstruction core (DEC:2083-2272, double cosine table). Frame n is intra iff
 * intra_period==1 (all-intra, DEC:98) || n % intra_period == 0. */
int icsp_oracle_decode(const int16_t* levels, const uint8_t* mpm, const uint8_t* ipm, const int16_t* mvd,
                       int nframes, int w, int h, int qdc, int qac, int intra_period, uint8_t* yuv_out);

/* Bitstream writer (ENC:4849-6334): 14-byte header + MSB-first body, tail right-aligned, body length
 * bits/8+1.  Returns the number of bytes written to out (capacity cap), or -1 if cap is too small. */
long icsp_oracle_write_bitstream(const int16_t* levels, const uint8_t* acflag, const uint8_t* mpm, const uint8_t* ipm,
                                 const int16_t* mvd, int nframes, int w, int h, int qdc, int qac, int intra_period,
                                 uint8_t* out, long cap);

/* Bit reader (DEC:14-404, 407-2025): parses header + body MSB-first, every byte including the last
 * (so a non-empty tail is mis-parsed exactly like the reference decoder does).  Returns 0 or -1. */
int icsp_oracle_parse_bitstream(const uint8_t* bin, long len, int nframes, int* w, int* h, int* qdc, int* qac,
                                int* intra_period, int16_t* levels, uint8_t* acflag, uint8_t* mpm, uint8_t* ipm,
                                int16_t* mvd);

/* Single-block transforms for the 1e-9 DCT check. table: 0 = encoder (float widened), 1 = decoder (double). */
void icsp_oracle_dct8x8(const int32_t* in, double* out, int nblocks);
void icsp_oracle_idct8x8(const int32_t* in, double* out, int nblocks, int table);

/* motionEstimation alone (ENC:2073-2155) for cur luma vs reference luma, with carried state. */
void icsp_oracle_me(const uint8_t* cur_y, const uint8_t* ref_y, int w, int h, int16_t* mv, int32_t* minsad,
                    int32_t* n_sad_evals);

#ifdef __cplusplus
}
#endif
#endif
