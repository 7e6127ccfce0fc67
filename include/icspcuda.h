/* libicspcuda — B200 (sm_100a) implementation of ICSPCodec's per-frame data-parallel core, behind a C ABI.
 *
 * The reference (JawThrow/ICSPCodec) has no plugin/FFI interface; the seam this library replaces is the
 * set of per-frame functions its frame loops call:
 *
 *   encoder  intraPrediction(FrameData&,int,int)              ICSP_Codec_Encoder.h:249   (ENC:556-643)
 *            allintraPrediction(FrameData*,int,int,int)       ICSP_Codec_Encoder.h:250   (ENC:446-555)
 *            interPrediction(FrameData&,FrameData&,int,int)   ICSP_Codec_Encoder.h:265   (ENC:1986-2072)
 *            called from single_thread_encoding (ENC:217-245) and encoding_thread (ENC:186-213)
 *   decoder  intraPredictionDecode / interPredictionDecode / allintraPredictionDecode
 *                                                             ICSP_Codec_Decoder.h:201-202,212 (DEC:2083-2272)
 *            called from IcspCodec::decoding (ICSP_Codec_Decoder.h:290-313)
 *
 * What those functions leave behind for the host — the fields intraBody/interBody (ENC:5032-5236) and the
 * next frame read — is returned here as caller-owned SoA arrays (below) instead of the reference's
 * per-block heap objects.  Plain pointers and sizes only; no C++/torch types.  All entry points return
 * ICSP_OK (0) or a negative error code and never exit the process (the reference exit(-1)s, ENC:64-81).
 *
 * Geometry: width and height multiples of 16.  nmb = (width/16)*(height/16);  fb = width*height*3/2.
 * Frames are planar I420 (Y, Cb, Cr).  A "GOP" is gop_len consecutive frames: 1 intra frame followed by
 * gop_len-1 inter frames, closed (README.md:152, ICSP_thread.cpp:39-77).  gop_len == 1 means all-intra.
 *
 * SoA layout (n = number of frames in the call, frame index f = gop*gop_len + t):
 *   levels  int16 [n][nmb][6][64]  blocks in bitstream order Y0,Y1,Y2,Y3,Cb,Cr; coefficients in zig-zag order
 *                                  (ENC:3031-3094); element 0 is the DC level AFTER DC-DPCM (what DCentropy codes)
 *   acflag  uint8 [n][nmb][6]      1 iff all 63 AC levels are zero (ENC:2784-2792)
 *   mpm     uint8 [n][nmb][4]      MPMFlag        (intra frames; 0 on inter frames)        ENC:5057
 *   ipm     uint8 [n][nmb][4]      intraPredMode  (intra frames; 0 on inter frames)        ENC:5058
 *   mvd     int16 [n][nmb][2]      differential motion vector x,y as coded (inter frames)  ENC:5154
 *   mv      int16 [n][nmb][2]      full motion vector (= Reconstructedmv)                  ENC:2145-2146
 *   minsad  int32 [n][nmb]         SAD of the selected candidate
 *   recon   uint8 [n][fb]          reconstructed I420 frame (reconstructedY/Cb/Cr; test_yuv.yuv ENC:6376-6421)
 *
 * Threading: a context is single-threaded; use one context per GPU per host thread.  Work is issued on the
 * context's own CUDA stream; *_async entry points return before completion, icsp_sync() waits.
 */
#ifndef ICSPCUDA_H
#define ICSPCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICSP_OK 0
#define ICSP_ERR_PARAM (-1)   /* bad argument / geometry */
#define ICSP_ERR_CUDA (-2)    /* CUDA runtime error (see icsp_last_error) */
#define ICSP_ERR_NOMEM (-3)   /* device or pinned-host allocation failed */
#define ICSP_ERR_CAPACITY (-4) /* more frames than the context was created for */

typedef struct icsp_ctx icsp_ctx;

typedef struct icsp_enc_out {     /* any pointer may be NULL: that array is not copied back */
    int16_t* levels;
    uint8_t* acflag;
    uint8_t* mpm;
    uint8_t* ipm;
    int16_t* mvd;
    int16_t* mv;
    int32_t* minsad;
    uint8_t* recon;
    double* dct;       /* optional debug tap, f64 [n][nmb][6][64]: forward DCT output of every block (DCT_block / CDCT_block,
                          ENC:2685-2749, 4338-4419), raster order inside a block, BEFORE the DC predictor is subtracted; the
                          north_star's 1e-9 check of the floating-point intermediates reads it.  Honoured by icsp_encode_gops,
                          icsp_intra_frame and icsp_inter_frame; 3 KB per macroblock, so keep n small. */
} icsp_enc_out;

typedef struct icsp_dec_in {      /* what the host bit reader parsed (DEC:38-404), same SoA layout */
    const int16_t* levels;
    const uint8_t* mpm;
    const uint8_t* ipm;
    const int16_t* mvd;
} icsp_dec_in;

/* ---- lifetime ----------------------------------------------------------------------------------- */
int icsp_create(icsp_ctx** ctx, int device, int width, int height, int max_frames);
void icsp_destroy(icsp_ctx* ctx);
const char* icsp_last_error(const icsp_ctx* ctx);   /* ctx may be NULL: returns the create-time error */
const char* icsp_version(void);
int icsp_sync(icsp_ctx* ctx);

/* Page-locked host memory for the SoA arrays / frames (fast, truly asynchronous H2D/D2H). Plain malloc'd
 * buffers are accepted everywhere too; they are just slower to copy. */
void* icsp_host_alloc(size_t bytes);
/* The same, write-combined: for buffers the CPU only WRITES (frames read from a file, then uploaded).  The device reads
 * them without snooping the CPU caches, which matters when several GPUs upload at once; reading them back on the CPU is
 * very slow, so never use it for outputs.  Freed with icsp_host_free. */
void* icsp_host_alloc_upload(size_t bytes);
void icsp_host_free(void* p);

/* ---- encoder: replaces intraPrediction / interPrediction over a batch of closed GOPs -------------- */
/* One call = H2D of the frames, the whole GOP batch on the GPU, D2H of the requested outputs, sync.
 * Step t of the schedule processes frame t of EVERY GOP in one set of launches. */
int icsp_encode_gops(icsp_ctx* ctx, const uint8_t* i420_frames, int n_gops, int gop_len, int qp_dc, int qp_ac,
                     const icsp_enc_out* out);
/* The same, split so that inputs can stay resident in HBM (bench "value" leg, multi-pass tools). */
int icsp_enc_upload(icsp_ctx* ctx, const uint8_t* i420_frames, int n_frames);                 /* async H2D */
int icsp_enc_run(icsp_ctx* ctx, int n_gops, int gop_len, int qp_dc, int qp_ac);               /* async */
int icsp_enc_download(icsp_ctx* ctx, int n_frames, const icsp_enc_out* out);                  /* async D2H */
/* Quality numbers without reading the reconstruction back (replaces the MSE loop of DEC.h:332-346 and
 * the need for checkResultFrames' dump, ENC:6376-6421, when only PSNR is wanted): sse[f*3 + {0,1,2}] = sum over the
 * Y / Cb / Cr plane of (source - reconstruction)^2 for the resident frames f < n_frames.  Synchronous.
 * PSNR_Y(f) = 20*log10(255/sqrt(sse[f*3]/(width*height))), exactly the reference's double arithmetic. */
int icsp_enc_sse(icsp_ctx* ctx, int n_frames, uint64_t* sse);

/* ---- encoder with entropy coding + bit packing on the GPU ------------------------- */
/* Replaces intraPrediction/interPrediction AND intraBody/interBody + DC/AC/MVentropy (ENC:5032-6334) for a batch of
 * independent streams: stream s = frames [s*gops_per_stream*gop_len, (s+1)*gops_per_stream*gop_len).
 * Only the packed bitstream bodies (and optionally the reconstruction) cross PCIe, not the int16 levels.
 *   bits          caller buffer of cap_bytes; stream s's body starts at bits + stream_offset[s] (16-byte aligned) and is
 *                 an MSB-first bit string of stream_bits[s] bits, zero padded to the next byte.  It is exactly the bit
 *                 string the reference concatenates in makebitstream (ENC:4873-4895); icsp_finish_body() applies the
 *                 reference's file rule (bits/8+1 bytes, tail bits right-aligned in the last byte).
 *   recon         optional [n][fb]; NULL = not copied back.
 * Returns ICSP_ERR_CAPACITY if cap_bytes is too small.  icsp_bits_bound() is always enough (800 bytes per macroblock:
 * the worst case of the VLC over an orthonormal 8x8 DCT of 8-bit residuals); the reference's own buffer, width*height
 * bytes per frame (ENC:4874), is enough for natural content but is overrun by noise at QP 1. */
typedef struct icsp_bits_out {
    uint8_t* bits;
    size_t cap_bytes;
    uint64_t* stream_bits;      /* [n_streams] */
    uint64_t* stream_offset;    /* [n_streams] */
    uint8_t* recon;
} icsp_bits_out;
int icsp_encode_streams(icsp_ctx* ctx, const uint8_t* i420_frames, int n_streams, int gops_per_stream, int gop_len,
                        int qp_dc, int qp_ac, const icsp_bits_out* out);
/* Resident variant: entropy-code what icsp_enc_run left on the device (async); results stay on the device until
 * icsp_bits_download. */
int icsp_entropy_run(icsp_ctx* ctx, int n_streams, int gops_per_stream, int gop_len);
int icsp_bits_download(icsp_ctx* ctx, int n_streams, const icsp_bits_out* out);   /* synchronous */
/* Macroblock-row index of what icsp_encode_streams / icsp_entropy_run just coded: rows[f*(height/16) + y]
 * = bit offset, from the start of its stream's body, of the first macroblock of row y of frame f (f < n_frames, frames in
 * the order of the call).  8 bytes per row (0.4 % of a CIF stream); icspenc --index stores it as "<bin>.idx".  The stream
 * format has no resynchronisation points (DEC:38-404 is one serial VLC chain); with this index every macroblock row is an
 * independent chain and icsp_decode_streams parses on the GPU.  Synchronous. */
int icsp_bits_row_index(icsp_ctx* ctx, int n_frames, uint64_t* rows);
/* In place: turns an MSB-first body of nbits bits (buffer must hold nbits/8+1 bytes) into the reference's file body;
 * returns its length nbits/8+1. */
size_t icsp_finish_body(uint8_t* body, uint64_t nbits);
/* worst-case size in bytes of the packed bodies of n_frames frames (any content, any QP), incl. alignment slack */
size_t icsp_bits_bound(int width, int height, int n_frames);

/* ---- decoder: replaces intraPredictionDecode / interPredictionDecode (double cosine table) -------- */
int icsp_decode_gops(icsp_ctx* ctx, const icsp_dec_in* in, int n_gops, int gop_len, int qp_dc, int qp_ac,
                     uint8_t* i420_out);
int icsp_dec_upload(icsp_ctx* ctx, const icsp_dec_in* in, int n_frames);
int icsp_dec_run(icsp_ctx* ctx, int n_gops, int gop_len, int qp_dc, int qp_ac);
int icsp_dec_download(icsp_ctx* ctx, int n_frames, uint8_t* i420_out);

/* ---- decoder with the bit reader on the GPU ------------------------------------------ */
/* Replaces readBlockData + DC/AC/MVientropy (DEC:38-2025) AND intraPredictionDecode / interPredictionDecode for a batch of
 * independent streams whose macroblock-row index is known: bodies in (file bytes after the 14-byte header, including the
 * reference's right-aligned last byte, ENC:4895), decoded I420 out.  One GPU thread parses one macroblock row.
 *   bits            bodies, stream s at bits + stream_offset[s] (offsets ascending, multiples of 4), stream_bytes[s] long
 *   row_bit_offset  [n_frames][height/16], see icsp_bits_row_index (offsets of a tail GOP decoded by a second call must be
 *                   passed relative to the same body start)
 *   out             [n_frames][fb]
 * A wrong index yields wrong pictures, never out-of-bounds accesses (offsets are validated, reads past a body return 0). */
typedef struct icsp_dec_bits_in {
    const uint8_t* bits;
    const uint64_t* stream_offset;   /* [n_streams] */
    const uint64_t* stream_bytes;    /* [n_streams] */
    const uint64_t* row_bit_offset;  /* [n_streams*gops_per_stream*gop_len][height/16] */
} icsp_dec_bits_in;
int icsp_decode_streams(icsp_ctx* ctx, const icsp_dec_bits_in* in, int n_streams, int gops_per_stream, int gop_len, int qp_dc,
                        int qp_ac, uint8_t* out);

/* ---- kernel-level shims (unit parity + micro-benchmarks); host pointers, synchronous ---------------- */
/* motionEstimation (ENC:2073-2155): n frame pairs, luma planes only [n][w*h]; carried spiral state per frame */
int icsp_me_sad(icsp_ctx* ctx, const uint8_t* cur_y, const uint8_t* ref_y, int n, int16_t* mv, int32_t* minsad);
/* DCT_block (ENC:2685-2749): int32 [n][64] -> f64 [n][64] */
int icsp_dct8x8(icsp_ctx* ctx, const int32_t* blocks, int n, double* out);
/* IDCT_block (ENC:2825-2893 when table==0, DEC:3331-3445 when table==1): int32 [n][64] -> f64 [n][64] */
int icsp_idct8x8(icsp_ctx* ctx, const int32_t* blocks, int n, int table, double* out);
/* Quantization_block (ENC:2750-2824; chroma != 0: CQuantization_block ENC:4610-4656): f64 [n][64] in raster order ->
 * int32 levels [n][64] (raster order, QstepDC at [0][0]) and acflag [n] (1 iff all 63 AC levels are zero; may be NULL) */
int icsp_quant(icsp_ctx* ctx, const double* dct, int n, int qp_dc, int qp_ac, int chroma, int32_t* levels, uint8_t* acflag);
/* intraPrediction (ENC:556-643) on n independent frames: same outputs as icsp_encode_gops with gop_len 1 */
int icsp_intra_frame(icsp_ctx* ctx, const uint8_t* i420_frames, int n, int qp_dc, int qp_ac, const icsp_enc_out* out);
/* interPrediction(cur, prev) (ENC:1986-2072) on n independent frame pairs: cur_i420[i] is coded against prev_recon_i420[i],
 * the RECONSTRUCTION of its predecessor (prev.reconstructedY/Cb/Cr, ENC:2087,2176,2518-2520).  Outputs as in icsp_enc_out
 * for the n current frames; needs a context of capacity >= 2*n frames. */
int icsp_inter_frame(icsp_ctx* ctx, const uint8_t* cur_i420, const uint8_t* prev_recon_i420, int n, int qp_dc, int qp_ac,
                     const icsp_enc_out* out);

/* ---- instrumentation ---------------------------------------------------------------------------- */
/* When enabled every kernel launch is bracketed by CUDA events on the context stream. */
#define ICSP_MAX_KERNELS 32
typedef struct icsp_kernel_stat {
    char name[40];
    uint64_t launches;
    double total_ms;
} icsp_kernel_stat;
int icsp_set_profiling(icsp_ctx* ctx, int enabled);
int icsp_reset_stats(icsp_ctx* ctx);
int icsp_get_stats(icsp_ctx* ctx, icsp_kernel_stat* stats, int cap);   /* returns number of entries, syncs */
uint64_t icsp_launch_count(const icsp_ctx* ctx);                       /* kernels launched since create/reset */
/* Execution shape: number of compute streams the GOP chunks are spread over (1..8, default 4) and GOPs per chunk
 * (0 = automatic).  n_streams = 1 with one chunk per call serialises every kernel on one stream, which is what the
 * per-kernel event timing wants. */
int icsp_configure(icsp_ctx* ctx, int n_compute_streams, int chunk_gops);
/* Event timing on the context stream: record slot a, do work, record slot b, elapsed(a,b) after icsp_sync. */
int icsp_event_record(icsp_ctx* ctx, int slot);                        /* slot in [0,8) */
int icsp_event_elapsed_ms(icsp_ctx* ctx, int slot_a, int slot_b, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* ICSPCUDA_H */
