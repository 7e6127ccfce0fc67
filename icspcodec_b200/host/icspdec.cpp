// icspdec — drop-in decoder front end: the reference's positional CLI (decode.cpp:4-27) and frame loop
// (IcspCodec::decoding, DEC.h:290-313) on the host, the reconstruction core (intraPredictionDecode /
// interPredictionDecode, DEC:2083-2272) on the GPU through libicspcuda (binary64 cosine table, DEC.h:19).
//
//   icspdec <nframes> <stream.bin> <QPDC> <QPAC> <intraPeriod> [<original.yuv>] [--gpus G] [--out file] [--index file | --no-index]
//
// When <stream.bin>.idx exists (icspenc --index: the bit offset of every macroblock row) the bit reader runs on the GPU too
// (icsp_decode_streams); otherwise the stream is parsed serially on the host (bitstream.cpp) as the reference does.
//
// QP / intraPeriod / size are taken from the stream header like the reference does (the positional values are
// accepted for CLI compatibility).  Output: check_test_intra_yuv.yuv when intraPeriod == 1, otherwise
// check_test_inter_yuv.yuv (DEC:4476-4527); with an original YUV the average luma PSNR and the decode time are
// appended to experimental_Result_Decoding.txt (DEC.h:315-355).  The reference opens "output\\<bin>" and
// "data\\<yuv>" with literal backslashes (DEC.h:241,323); both spellings are tried.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/icspcuda.h"
#include "bitstream.h"

static FILE* open_either(const std::string& plain, const std::string& prefixed)
{
    FILE* f = fopen(plain.c_str(), "rb");
    return f ? f : fopen(prefixed.c_str(), "rb");
}

int main(int argc, char** argv)
{
    std::vector<std::string> pos;
    int gpus = 1;
    std::string outname, idxname;
    bool use_index = true;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "--gpus" && i + 1 < argc) gpus = atoi(argv[++i]);
        else if (a == "--index" && i + 1 < argc) idxname = argv[++i];
        else if (a == "--no-index") use_index = false;
        else if (a == "--out" && i + 1 < argc) outname = argv[++i];
        else pos.push_back(a);
    }
    if (pos.size() < 5) { fprintf(stderr, "usage: icspdec <nframes> <stream.bin> <QPDC> <QPAC> <intraPeriod> [<original.yuv>] [--gpus G] [--out file]\n"); return 1; }
    const int nframes = atoi(pos[0].c_str());
    if (nframes <= 0) { fprintf(stderr, "[ERROR] nframes must be positive\n"); return 1; }
    FILE* fb_ = open_either(pos[1], "output\\" + pos[1]);
    if (!fb_) { fprintf(stderr, "[ERROR] cannot open %s\n", pos[1].c_str()); return 1; }
    std::vector<uint8_t> file;
    { uint8_t buf[1 << 16]; size_t r; while ((r = fread(buf, 1, sizeof(buf), fb_)) > 0) file.insert(file.end(), buf, buf + r); }
    fclose(fb_);

    // macroblock-row index side-car: "ICSPIDX1", width, height, frames, macroblock rows (u32 LE), u64 LE bit offsets
    std::vector<uint64_t> index;
    if (use_index) {
        FILE* fx = idxname.empty() ? open_either(pos[1] + ".idx", "output\\" + pos[1] + ".idx") : fopen(idxname.c_str(), "rb");
        if (fx) {
            char magic[8]; uint32_t hdr[4];
            if (fread(magic, 1, 8, fx) == 8 && !memcmp(magic, "ICSPIDX1", 8) && fread(hdr, 4, 4, fx) == 4 && hdr[2] >= (uint32_t)nframes &&
                hdr[3] == hdr[1] / 16) {
                index.resize((size_t)nframes * hdr[3]);
                if (fread(index.data(), 8, index.size(), fx) != index.size()) index.clear();
                if (!index.empty() && file.size() >= 14 &&
                    (hdr[0] != (uint32_t)(file[7] | (file[8] << 8)) || hdr[1] != (uint32_t)(file[5] | (file[6] << 8)))) index.clear();
            }
            fclose(fx);
            if (index.empty()) fprintf(stderr, "icspdec: index does not match the stream, parsing on the host\n");
        } else if (!idxname.empty()) { fprintf(stderr, "[ERROR] cannot open %s\n", idxname.c_str()); return 1; }
    }
    const bool gpu_parse = !index.empty();
    icsp_host::ParsedStream ps;
    try { ps = icsp_host::parse_stream(file, gpu_parse ? 0 : nframes); }     // 0 frames: header only
    catch (const std::exception& e) { fprintf(stderr, "[ERROR] %s\n", e.what()); return 1; }
    const int w = ps.p.width, h = ps.p.height, ip = ps.p.intra_period, nmb = (w / 16) * (h / 16);
    const size_t fbytes = (size_t)w * h * 3 / 2;
    std::vector<uint8_t> yuv((size_t)nframes * fbytes);

    const auto t0 = std::chrono::steady_clock::now();
    const int gop = ip;                              // intra iff ip == 1 or n % ip == 0
    const int full = nframes / gop, tail = nframes - full * gop;
    struct Shard { int device, first, n_gops, gop_len, rc; std::string err; };
    std::vector<Shard> shards;
    const int G = gpus > 0 ? gpus : 1;
    for (int d = 0; d < G; d++) {
        const int g0 = (int)((long long)full * d / G), g1 = (int)((long long)full * (d + 1) / G);
        if (g1 > g0) shards.push_back({d, g0 * gop, g1 - g0, gop, 0, ""});
    }
    if (tail) shards.push_back({G - 1, full * gop, 1, tail, 0, ""});
    auto run = [&](Shard& s) {
        const int per_call = std::max(1, std::min(s.n_gops, 4096 / s.gop_len));
        icsp_ctx* ctx = nullptr;
        s.rc = icsp_create(&ctx, s.device, w, h, per_call * s.gop_len);
        if (s.rc) { s.err = icsp_last_error(nullptr); return; }
        for (int g = 0; g < s.n_gops && !s.rc; g += per_call) {
            const int ng = std::min(per_call, s.n_gops - g);
            const size_t f0 = (size_t)s.first + (size_t)g * s.gop_len;
            if (gpu_parse) {   // only the bytes of these frames travel: [first row of f0, first row of the frame after the call)
                const size_t mbh = (size_t)h / 16, f1 = f0 + (size_t)ng * s.gop_len, body = file.size() - 14;
                const uint64_t start = (index[f0 * mbh] / 8) & ~(uint64_t)3;
                const uint64_t end = f1 < (size_t)nframes ? std::min<uint64_t>(body, index[f1 * mbh] / 8 + 1) : body;
                if (start > end) { s.rc = ICSP_ERR_PARAM; s.err = "row index is not ascending"; break; }
                std::vector<uint64_t> rows(index.begin() + (long)(f0 * mbh), index.begin() + (long)(f1 * mbh));
                for (auto& r : rows) r -= start * 8;
                const uint64_t off = 0, len = end - start;
                icsp_dec_bits_in in{file.data() + 14 + start, &off, &len, rows.data()};
                s.rc = icsp_decode_streams(ctx, &in, 1, ng, s.gop_len, ps.p.qp_dc, ps.p.qp_ac, yuv.data() + f0 * fbytes);
            } else {
                icsp_dec_in in{ps.levels.data() + f0 * nmb * 384, ps.mpm.data() + f0 * nmb * 4, ps.ipm.data() + f0 * nmb * 4, ps.mvd.data() + f0 * nmb * 2};
                s.rc = icsp_decode_gops(ctx, &in, ng, s.gop_len, ps.p.qp_dc, ps.p.qp_ac, yuv.data() + f0 * fbytes);
            }
            if (s.rc) s.err = icsp_last_error(ctx);
        }
        icsp_destroy(ctx);
    };
    {
        std::vector<std::thread> th;
        for (int d = 0; d < G; d++) th.emplace_back([&, d] { for (auto& s : shards) if (s.device == d) run(s); });
        for (auto& t : th) t.join();
    }
    for (auto& s : shards)
        if (s.rc) { fprintf(stderr, "[ERROR] GPU %d: %s (code %d)\n", s.device, s.err.c_str(), s.rc); return 1; }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    if (outname.empty()) outname = ip == 1 ? "check_test_intra_yuv.yuv" : "check_test_inter_yuv.yuv";
    FILE* fo = fopen(outname.c_str(), "wb");
    if (!fo) { fprintf(stderr, "fail to open %s\n", outname.c_str()); return 1; }
    fwrite(yuv.data(), 1, yuv.size(), fo);
    fclose(fo);

    if (pos.size() >= 6) {   // PSNR against the original (DEC.h:315-355): mean over frames of the luma PSNR
        FILE* fy = open_either(pos[5], "data\\" + pos[5]);
        if (fy) {
            std::vector<uint8_t> org(fbytes);
            double acc = 0;
            int cnt = 0;
            for (int n = 0; n < nframes && fread(org.data(), 1, fbytes, fy) == fbytes; n++, cnt++) {
                double mse = 0;
                const uint8_t* d = yuv.data() + (size_t)n * fbytes;
                for (int i = 0; i < w * h; i++) { const double e = (double)org[i] - d[i]; mse += e * e; }
                mse /= (double)w * h;
                acc += mse > 0 ? 10.0 * std::log10(255.0 * 255.0 / mse) : 100.0;
            }
            fclose(fy);
            if (cnt) {
                FILE* fl = fopen("experimental_Result_Decoding.txt", "a");
                if (fl) { fprintf(fl, "%s QPDC %d QPAC %d intraPeriod %d frames %d decode %.4f s PSNR-Y %.4f dB\n", pos[1].c_str(), ps.p.qp_dc, ps.p.qp_ac, ip, cnt, secs, acc / cnt); fclose(fl); }
                fprintf(stderr, "icspdec: PSNR-Y %.4f dB over %d frames\n", acc / cnt, cnt);
            }
        }
    }
    fprintf(stderr, "icspdec: %d frames %dx%d decoded in %.3f s (%.0f fps, bit reader on the %s) -> %s\n", nframes, w, h, secs, nframes / secs,
            gpu_parse ? "GPU" : "host", outname.c_str());
    return 0;
}
