// Reference-side binding of libicspcuda: the ONE translation unit a maintainer of JawThrow/ICSPCodec adds.
//
// It defines single_thread_encoding() — declared ICSP_Codec_Encoder.h:241, defined ICSP_Codec_Encoder_source.cpp:217-245,
// called from IcspCodec::encoding (ICSPCodec.cpp:41) — so that the reference's OWN main (encoder_main.cpp), option
// parser, loader (IcspCodec::init: YCbCrLoad / splitFrames / splitBlocks, ENC:247-443), bitstream writer
// (makebitstream + intraBody / interBody / allintraBody + DC/AC/MVentropy, ENC:4849-6334) and result dump
// (checkResultFrames, ENC:6376-6421) run unchanged on top of the GPU library.  Only the per-frame calls
// intraPrediction / interPrediction / allintraPrediction (ENC.h:249-250,265) are replaced, by ONE icsp_encode_gops call
// over all closed GOPs of the sequence (+ one for the n % intraPeriod tail).
//
// Build (oracle/build_ref.sh does exactly this; nothing of the reference is copied or edited):
//   g++ -O2 -w -c -Dsingle_thread_encoding=ref_single_thread_encoding  $REF/source/encoder/ICSP_Codec_Encoder_source.cpp
//   g++ -O2 -w -pthread -I$REF/source/encoder -Iinclude  encoder_main.cpp ICSPCodec.cpp ICSP_thread.cpp \
//       ICSP_Codec_Encoder_source.o integration/icsp_ref_shim.cpp -Licspcodec_b200 -licspcuda -o ICSPCodec_gpu
// (the -D renames the reference's own definition inside its translation unit only, so that the two do not collide; a
// maintainer would simply delete the old body).
//
// What the shim hands back is exactly what the writer and the dump read:
//   I frames, per macroblock: MPMFlag[4], intraPredMode[4], intraReorderedblck8[4] (int[64] each), intraACflag[4]
//             (ENC:5057-5078), Cbblocks/Crblocks[mb].intraReorderedblck, .intraACflag (ENC:5088-5128)
//   P frames, per macroblock: mv (the DIFFERENTIAL vector, ENC:5154), interReorderedblck8[4], interACflag[4]
//             (ENC:5167-5184), chroma interReorderedblck / interACflag (ENC:5195-5233)
//   every frame: reconstructedY / Cb / Cr, malloc'd (checkResultFrames frees them, ENC:6414-6419)
#include "ICSP_Codec_Encoder.h"
#include "icspcuda.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <malloc.h>
#include <vector>

static void die(const char* what, icsp_ctx* ctx)
{
    fprintf(stderr, "icsp_ref_shim: %s: %s\n", what, icsp_last_error(ctx));
    exit(-1);                                       // the reference's own error convention (ENC:64-81)
}

static int* zigzag_copy(const int16_t* src)
{
    int* z = (int*)malloc(64 * sizeof(int));        // the writer reads int[64] per block and never frees it
    for (int i = 0; i < 64; i++) z[i] = src[i];
    return z;
}

void single_thread_encoding(FrameData* frames, YCbCr_t* YCbCr, int intra_period, int QstepDC, int QstepAC)
{
    // makebitstream ORs the bits of the last, partial byte into a malloc'd (never cleared) buffer (ENC:4875, 4895):
    // the stock binary gets that buffer from a fresh, zero-filled mmap, which is what the golden files contain.  The CUDA runtime
    // leaves a well-used heap behind, so pin glibc's mmap threshold: every large allocation of the writer is a fresh mapping again.
    mallopt(M_MMAP_THRESHOLD, 64 * 1024);
    const int n = YCbCr->nframe, w = YCbCr->width, h = YCbCr->height;
    const int nmb = (w / 16) * (h / 16), ysz = w * h, csz = ysz / 4, fb = ysz + 2 * csz;
    const bool all_intra = intra_period == ALL_INTRA;
    const int gop = all_intra ? 1 : intra_period, full = n / gop, tail = n - full * gop;

    // input frames into pinned memory.  frames[i].Y is the planar luma (splitFrames, ENC:284-310); frames[i].Cb / Cr were
    // FREED at the end of splitBlocks (ENC:436-441) — the chroma pixels only survive in the per-block copies
    // Cbblocks[mb].originalblck8 (ENC:408-424), so the planes are reassembled from those
    uint8_t* i420 = (uint8_t*)icsp_host_alloc_upload((size_t)n * fb);
    uint8_t* recon = (uint8_t*)icsp_host_alloc((size_t)n * fb);
    int16_t* levels = (int16_t*)icsp_host_alloc((size_t)n * nmb * 384 * sizeof(int16_t));
    if (!i420 || !recon || !levels) die("pinned allocation failed", NULL);
    const int cw = w / 2, cbw = cw / 8;
    for (int i = 0; i < n; i++) {
        uint8_t* dst = i420 + (size_t)i * fb;
        memcpy(dst, frames[i].Y, ysz);
        for (int mb = 0; mb < nmb; mb++) {
            const int x0 = (mb % cbw) * 8, y0 = (mb / cbw) * 8;
            for (int y = 0; y < 8; y++) {
                memcpy(dst + ysz + (size_t)(y0 + y) * cw + x0, frames[i].Cbblocks[mb].originalblck8->block[y], 8);
                memcpy(dst + ysz + csz + (size_t)(y0 + y) * cw + x0, frames[i].Crblocks[mb].originalblck8->block[y], 8);
            }
        }
    }
    std::vector<int16_t> mvd((size_t)n * nmb * 2);
    std::vector<uint8_t> acflag((size_t)n * nmb * 6), mpm((size_t)n * nmb * 4), ipm((size_t)n * nmb * 4);

    icsp_ctx* ctx = NULL;
    if (icsp_create(&ctx, 0, w, h, n) != ICSP_OK) die("icsp_create", NULL);
    auto run = [&](int first, int ngops, int len) {
        icsp_enc_out o;
        memset(&o, 0, sizeof(o));
        o.levels = levels + (size_t)first * nmb * 384;
        o.acflag = &acflag[(size_t)first * nmb * 6];
        o.mpm = &mpm[(size_t)first * nmb * 4];
        o.ipm = &ipm[(size_t)first * nmb * 4];
        o.mvd = &mvd[(size_t)first * nmb * 2];
        o.recon = recon + (size_t)first * fb;
        if (icsp_encode_gops(ctx, i420 + (size_t)first * fb, ngops, len, QstepDC, QstepAC, &o) != ICSP_OK) die("icsp_encode_gops", ctx);
    };
    if (full) run(0, full, gop);
    if (tail) run(full * gop, 1, tail);
    icsp_destroy(ctx);

    // scatter the SoA results into the per-block fields the reference's writer reads
    for (int f = 0; f < n; f++) {
        const bool intra = all_intra || f % intra_period == 0;
        for (int mb = 0; mb < nmb; mb++) {
            const size_t m = (size_t)f * nmb + mb;
            BlockData& bd = frames[f].blocks[mb];
            CBlockData& cb = frames[f].Cbblocks[mb];
            CBlockData& cr = frames[f].Crblocks[mb];
            for (int k = 0; k < 4; k++) {
                int* z = zigzag_copy(levels + (m * 6 + k) * 64);
                if (intra) {
                    bd.intraReorderedblck8[k] = z; bd.intraACflag[k] = acflag[m * 6 + k];
                    bd.MPMFlag[k] = mpm[m * 4 + k]; bd.intraPredMode[k] = ipm[m * 4 + k];
                } else {
                    bd.interReorderedblck8[k] = z; bd.interACflag[k] = acflag[m * 6 + k];
                }
            }
            if (intra) {
                cb.intraReorderedblck = zigzag_copy(levels + (m * 6 + 4) * 64); cb.intraACflag = acflag[m * 6 + 4];
                cr.intraReorderedblck = zigzag_copy(levels + (m * 6 + 5) * 64); cr.intraACflag = acflag[m * 6 + 5];
            } else {
                bd.mv.x = mvd[m * 2]; bd.mv.y = mvd[m * 2 + 1];        // interBody codes bd.mv, the differential vector
                cb.interReorderedblck = zigzag_copy(levels + (m * 6 + 4) * 64); cb.interACflag = acflag[m * 6 + 4];
                cr.interReorderedblck = zigzag_copy(levels + (m * 6 + 5) * 64); cr.interACflag = acflag[m * 6 + 5];
            }
        }
        frames[f].reconstructedY = (unsigned char*)malloc(ysz);          // freed by checkResultFrames
        frames[f].reconstructedCb = (unsigned char*)malloc(csz);
        frames[f].reconstructedCr = (unsigned char*)malloc(csz);
        memcpy(frames[f].reconstructedY, recon + (size_t)f * fb, ysz);
        memcpy(frames[f].reconstructedCb, recon + (size_t)f * fb + ysz, csz);
        memcpy(frames[f].reconstructedCr, recon + (size_t)f * fb + ysz + csz, csz);
        print_frame_end_message(f, intra ? I_FRAME : P_FRAME);
    }
    icsp_host_free(i420); icsp_host_free(recon); icsp_host_free(levels);

    // the reference's own writer and dump, untouched (same calls as ENC:222-223 / 241-242)
    if (all_intra) {
        makebitstream(frames, n, h, w, QstepDC, QstepAC, intra_period, INTRA);
        checkResultFrames(frames, w, h, n, INTRA, SAVE_YUV);
    } else {
        makebitstream(frames, n, h, w, QstepDC, QstepAC, intra_period, INTER);
        checkResultFrames(frames, w, h, n, INTER, SAVE_YUV);
    }
}
