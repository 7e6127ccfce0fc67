// TEST INFRASTRUCTURE — function-level tap harness around the UNMODIFIED reference encoder TUs.
// Built by oracle/build_ref.sh into oracle/_ref/ref_taps by compiling this file together with
// /root/reference/source/encoder/{ICSP_Codec_Encoder_source,ICSPCodec,ICSP_thread}.cpp (sources are
// compiled where they lie; nothing is copied).  All hot functions have external linkage
// (ICSP_Codec_Encoder.h:249-304), so the harness only has to build the FrameData the same way
// IcspCodec::init does and call them.
//
//   ref_taps mv   <in_yuv> <nframes> <qdc> <qac> <ip> <out>   full MVs (Reconstructedmv) of every frame,
//                                                            int32 [nframes][396][2] (I-frames: zeros)
//   ref_taps dct  <in> <out>      in: int32 [n][64] residual blocks -> out: f64 [n][64]   (DCT_block)
//   ref_taps idct <in> <out>      in: int32 [n][64] dequantised blocks -> out: f64 [n][64] (IDCT_block)
//   ref_taps quant <in> <qdc> <qac> <chroma> <out>   in: f64 [n][64] DCT blocks -> out: int32 [n][64] levels + int32 [n] ACflag
//                                                            (Quantization_block / CQuantization_block)
//   ref_taps metime <in_yuv> <nframes> <reps>   time motionEstimation() alone, prints positions/s
#include "ICSP_Codec_Encoder.h"
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <chrono>

static std::vector<int> read_i32(const char* fn)
{
    FILE* f = fopen(fn, "rb");
    if (!f) { perror(fn); exit(2); }
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<int> v(sz / 4);
    if (fread(v.data(), 4, v.size(), f) != v.size()) { perror("fread"); exit(2); }
    fclose(f);
    return v;
}

static int tap_dct(const char* in, const char* out, bool inverse)
{
    std::vector<int> v = read_i32(in);
    size_t n = v.size() / 64;
    FILE* fo = fopen(out, "wb");
    BlockData bd;
    memset(&bd, 0, sizeof(bd));
    bd.blocksize1 = 16; bd.blocksize2 = 8;
    Block8i* errp[4]; Block8d* dctp[4]; Block8i* iqp[4]; Block8d* idctp[4];
    bd.intraErrblck = errp; bd.intraDCTblck = dctp; bd.intraInverseQuanblck = iqp; bd.intraInverseDCTblck = idctp;
    for (size_t i = 0; i < n; i++) {
        if (!inverse) {
            errp[0] = (Block8i*)malloc(sizeof(Block8i));      // DCT_block frees its input
            dctp[0] = (Block8d*)malloc(sizeof(Block8d));
            memcpy(errp[0]->block, &v[i * 64], 64 * sizeof(int));
            DCT_block(bd, 0, 8, INTRA);
            fwrite(dctp[0]->block, sizeof(double), 64, fo);
            free(dctp[0]);
        } else {
            iqp[0] = (Block8i*)malloc(sizeof(Block8i));
            idctp[0] = (Block8d*)malloc(sizeof(Block8d));
            memcpy(iqp[0]->block, &v[i * 64], 64 * sizeof(int));
            IDCT_block(bd, 0, 8, INTRA);
            fwrite(idctp[0]->block, sizeof(double), 64, fo);
            free(iqp[0]); free(idctp[0]);
        }
    }
    fclose(fo);
    return 0;
}

static int tap_quant(const char* in, int qdc, int qac, int chroma, const char* out)
{
    FILE* f = fopen(in, "rb");
    if (!f) { perror(in); return 2; }
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<double> v(sz / 8);
    if (fread(v.data(), 8, v.size(), f) != v.size()) { perror("fread"); return 2; }
    fclose(f);
    const size_t n = v.size() / 64;
    std::vector<int> lv(n * 64), ac(n);
    for (size_t i = 0; i < n; i++) {
        if (!chroma) {
            BlockData bd;
            memset(&bd, 0, sizeof(bd));
            Block8d* dctp[4]; Block8i* qp[4];
            bd.intraDCTblck = dctp; bd.intraQuanblck = qp;
            dctp[0] = (Block8d*)malloc(sizeof(Block8d));     // Quantization_block frees its input (ENC:2795)
            qp[0] = (Block8i*)malloc(sizeof(Block8i));
            memcpy(dctp[0]->block, &v[i * 64], 64 * sizeof(double));
            Quantization_block(bd, 0, 8, qdc, qac, INTRA);
            memcpy(&lv[i * 64], qp[0]->block, 64 * sizeof(int));
            ac[i] = bd.intraACflag[0];
            free(qp[0]);
        } else {
            CBlockData bd;
            memset(&bd, 0, sizeof(bd));
            bd.intraDCTblck = (Block8d*)malloc(sizeof(Block8d));
            bd.intraQuanblck = (Block8i*)malloc(sizeof(Block8i));
            memcpy(bd.intraDCTblck->block, &v[i * 64], 64 * sizeof(double));
            CQuantization_block(bd, 8, qdc, qac, INTRA);
            memcpy(&lv[i * 64], bd.intraQuanblck->block, 64 * sizeof(int));
            ac[i] = bd.intraACflag;
            free(bd.intraQuanblck);
        }
    }
    FILE* fo = fopen(out, "wb");
    fwrite(lv.data(), sizeof(int), lv.size(), fo);
    fwrite(ac.data(), sizeof(int), ac.size(), fo);
    fclose(fo);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: ref_taps mv|dct|idct|quant|metime ...\n"); return 2; }
    if (!strcmp(argv[1], "quant") && argc == 7) return tap_quant(argv[2], atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), argv[6]);
    if (!strcmp(argv[1], "dct") && argc == 4)  return tap_dct(argv[2], argv[3], false);
    if (!strcmp(argv[1], "idct") && argc == 4) return tap_dct(argv[2], argv[3], true);

    if (!strcmp(argv[1], "mv") && argc == 8) {
        int nframes = atoi(argv[3]), qdc = atoi(argv[4]), qac = atoi(argv[5]), ip = atoi(argv[6]);
        IcspCodec codec;
        codec.init(nframes, argv[2], 352, 288, qdc, qac);
        FILE* fo = fopen(argv[7], "wb");
        for (int n = 0; n < nframes; n++) {
            std::vector<int> mv(396 * 2, 0);
            if (n % ip == 0) {
                intraPrediction(codec.frames[n], qdc, qac);
            } else {
                interPrediction(codec.frames[n], codec.frames[n - 1], qdc, qac);
                for (int b = 0; b < 396; b++) {
                    mv[2 * b] = codec.frames[n].blocks[b].Reconstructedmv.x;
                    mv[2 * b + 1] = codec.frames[n].blocks[b].Reconstructedmv.y;
                }
            }
            fwrite(mv.data(), sizeof(int), mv.size(), fo);
        }
        fclose(fo);
        _Exit(0);   // skip the reference's destructor (double frees on this path)
    }

    if (!strcmp(argv[1], "metime") && argc == 5) {
        // motionEstimation() alone: frame 0 is intra-coded once to obtain a reconstructed reference,
        // then ME of frame k against frame 0's reconstruction is timed (396*64 positions per call
        // unless zero-SAD early breaks fire).
        int nframes = atoi(argv[3]), reps = atoi(argv[4]);
        IcspCodec codec;
        codec.init(nframes, argv[2], 352, 288, 8, 8);
        intraPrediction(codec.frames[0], 8, 8);
        auto t0 = std::chrono::steady_clock::now();
        long calls = 0;
        for (int r = 0; r < reps; r++)
            for (int n = 1; n < nframes; n++) { motionEstimation(codec.frames[n], codec.frames[0]); calls++; }
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("{\"calls\": %ld, \"seconds\": %.6f, \"positions_per_s\": %.1f}\n", calls, s, calls * 396.0 * 64.0 / s);
        fflush(stdout);   // _Exit does not flush, and stdout is fully buffered when it is a pipe
        _Exit(0);
    }
    fprintf(stderr, "bad arguments\n");
    return 2;
}
