// icspdec — drop-in decoder front end: the reference's positional CLI (decode.cpp:4-27) and frame loop
// (IcspCodec::decoding, DEC.h:290-313) on the host, the reconstruction core (intraPredictionDecode /
// interPredictionDecode, DEC:2083-2272) on the GPU through libicspcuda (binary64 cosine table, DEC.h:19).
//
//   icspdec <nframes> <stream.bin> <QPDC> <QPAC> <intraPeriod> [<original.yuv>] [--gpus G] [--out file] [--index file | --no-index]
//   icspdec --batch <list.txt> <nframes> [--gpus G] [--wave S] [--io-threads T] [--no-index]
//           N streams (one .bin path per line, same geometry / QP / intraPeriod) in ONE process, sharded over the GPUs and
//           decoded in waves of S streams per device call: with <bin>.idx side-cars the bit reader runs on the GPU
//           (icsp_decode_streams over the whole wave); without them the streams of a wave are parsed in parallel on the host,
//           one stream per thread (the stream is one serial VLC chain, DEC:38-404), and reconstructed by one icsp_decode_gops
//           call.  Output per stream: <name>.yuv (the .bin's file name + ".yuv") in the working directory.
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
#include <atomic>
#include <chrono>
#include <fstream>
#include <future>
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

template <typename F>
static void parallel_for(int n, int threads, F fn)
{
    std::atomic<int> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < std::max(1, std::min(threads, n)); t++)
        th.emplace_back([&] { for (int i = next++; i < n; i = next++) fn(i); });
    for (auto& t : th) t.join();
}

static bool read_file(const std::string& path, std::vector<uint8_t>& out)
{
    FILE* f = open_either(path, "output\\" + path);
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(sz > 0 ? (size_t)sz : 0);
    const bool ok = sz >= 0 && fread(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

// macroblock-row index side-car: "ICSPIDX1", width, height, frames, macroblock rows (u32 LE), u64 LE bit offsets
static bool read_index(const std::string& path, int nframes, const std::vector<uint8_t>& file, std::vector<uint64_t>& index)
{
    index.clear();
    FILE* fx = open_either(path, "output\\" + path);
    if (!fx) return false;
    char magic[8]; uint32_t hdr[4];
    if (fread(magic, 1, 8, fx) == 8 && !memcmp(magic, "ICSPIDX1", 8) && fread(hdr, 4, 4, fx) == 4 && hdr[2] >= (uint32_t)nframes && hdr[3] == hdr[1] / 16) {
        index.resize((size_t)nframes * hdr[3]);
        if (fread(index.data(), 8, index.size(), fx) != index.size()) index.clear();
        if (!index.empty() && file.size() >= 14 && (hdr[0] != (uint32_t)(file[7] | (file[8] << 8)) || hdr[1] != (uint32_t)(file[5] | (file[6] << 8)))) index.clear();
    }
    fclose(fx);
    return !index.empty();
}

static int run_batch(const std::string& list, int nframes, int gpus, int wave, int io_threads, bool use_index)
{
    std::vector<std::string> inputs;
    {
        std::ifstream f(list);
        if (!f) { fprintf(stderr, "[ERROR] cannot open %s\n", list.c_str()); return 1; }
        for (std::string line; std::getline(f, line);) {
            while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
            if (!line.empty()) inputs.push_back(line);
        }
    }
    const int S = (int)inputs.size();
    if (S == 0 || nframes <= 0) { fprintf(stderr, "[ERROR] --batch needs a non-empty list and nframes > 0\n"); return 1; }
    const int G = std::max(1, std::min(gpus, S));
    const int io = io_threads > 0 ? io_threads : std::max(2, (int)std::thread::hardware_concurrency() / G);
    // the first stream's header fixes geometry and quantisers of the batch
    std::vector<uint8_t> first;
    if (!read_file(inputs[0], first)) { fprintf(stderr, "[ERROR] cannot open %s\n", inputs[0].c_str()); return 1; }
    icsp_host::ParsedStream hd;
    try { hd = icsp_host::parse_stream(first, 0); } catch (const std::exception& e) { fprintf(stderr, "[ERROR] %s: %s\n", inputs[0].c_str(), e.what()); return 1; }
    const int w = hd.p.width, h = hd.p.height, ip = hd.p.intra_period, qdc = hd.p.qp_dc, qac = hd.p.qp_ac, nmb = (w / 16) * (h / 16), mbh = h / 16;
    const size_t fbytes = (size_t)w * h * 3 / 2;
    if (ip <= 0 || qdc <= 0 || qac <= 0) { fprintf(stderr, "[ERROR] %s: header with intraPeriod 0 / QP 0 (the reference decoder divides by it, DEC:98,201)\n", inputs[0].c_str()); return 1; }
    const int full = nframes / ip, tail = nframes - full * ip;
    std::atomic<int> failed{0};
    std::vector<std::string> errs(G);
    std::atomic<int> on_gpu{0};
    const auto t0 = std::chrono::steady_clock::now();

    auto gpu_thread = [&](int d) {
        const int s_begin = (int)((long long)S * d / G), s_end = (int)((long long)S * (d + 1) / G), mine = s_end - s_begin;
        if (mine <= 0) return;
        const int W = std::max(1, std::min(wave, mine));
        auto bail = [&](const std::string& m) { errs[d] = m; failed++; };
        icsp_ctx* ctx = nullptr;
        if (icsp_create(&ctx, d, w, h, W * std::max(full * ip, tail))) return bail(icsp_last_error(nullptr));
        struct Buf {
            int first = 0, count = 0; bool indexed = false;
            std::vector<std::vector<uint8_t>> files; std::vector<std::vector<uint64_t>> index; std::vector<icsp_host::ParsedStream> parsed;
            uint8_t* yuv = nullptr;
        } buf[2];
        for (auto& b : buf) { b.yuv = (uint8_t*)icsp_host_alloc((size_t)W * nframes * fbytes); if (!b.yuv) return bail("pinned host allocation failed"); }
        auto load = [&](Buf& b, int first_s, int count) {
            b.first = first_s; b.count = count; b.files.assign(count, {}); b.index.assign(count, {}); b.parsed.assign(count, {});
            std::atomic<int> have_idx{0};
            parallel_for(count, io, [&](int i) {
                if (!read_file(inputs[first_s + i], b.files[i])) { errs[d] = "cannot open " + inputs[first_s + i]; failed++; return; }
                if (use_index && read_index(inputs[first_s + i] + ".idx", nframes, b.files[i], b.index[i])) have_idx++;
            });
            b.indexed = have_idx == count;
            if (!b.indexed && !failed)      // no side-car: serial VLC parse on the host, one stream per thread
                parallel_for(count, io, [&](int i) {
                    try { b.parsed[i] = icsp_host::parse_stream(b.files[i], nframes); }
                    catch (const std::exception& e) { errs[d] = inputs[first_s + i] + ": " + e.what(); failed++; }
                });
            for (int i = 0; i < count && !failed; i++) {
                icsp_host::ParsedStream one;
                try { one = b.indexed ? icsp_host::parse_stream(b.files[i], 0) : icsp_host::ParsedStream{b.parsed[i].p, {}, {}, {}, {}, {}}; } catch (...) { errs[d] = "bad header in " + inputs[first_s + i]; failed++; break; }
                if (one.p.width != w || one.p.height != h || one.p.qp_dc != qdc || one.p.qp_ac != qac || one.p.intra_period != ip) {
                    errs[d] = inputs[first_s + i] + ": geometry / QP / intraPeriod differ from the first stream of the batch"; failed++;
                }
            }
        };
        auto decode = [&](Buf& b) -> int {
            int rc = ICSP_OK;
            if (b.indexed) {
                on_gpu += b.count;
                // bodies back to back at 4-byte aligned offsets; the output of a call is [stream][frames of the call]
                std::vector<uint8_t> blob; std::vector<uint64_t> off(b.count), len(b.count);
                for (int i = 0; i < b.count; i++) {
                    while (blob.size() & 3) blob.push_back(0);
                    off[i] = blob.size(); len[i] = b.files[i].size() - 14;
                    blob.insert(blob.end(), b.files[i].begin() + 14, b.files[i].end());
                }
                auto call = [&](int f0, int ngops, int glen, uint8_t* dst) {
                    const int nf = ngops * glen;
                    std::vector<uint64_t> rows((size_t)b.count * nf * mbh);
                    for (int i = 0; i < b.count; i++)
                        std::copy(b.index[i].begin() + (size_t)f0 * mbh, b.index[i].begin() + (size_t)(f0 + nf) * mbh, rows.begin() + (size_t)i * nf * mbh);
                    icsp_dec_bits_in in{blob.data(), off.data(), len.data(), rows.data()};
                    return icsp_decode_streams(ctx, &in, b.count, ngops, glen, qdc, qac, dst);
                };
                if (full) rc = call(0, full, ip, b.yuv);
                if (!rc && tail) rc = call(full * ip, 1, tail, b.yuv + (size_t)W * full * ip * fbytes);
            } else {
                auto call = [&](int f0, int ngops, int glen, uint8_t* dst) {
                    const size_t nf = (size_t)ngops * glen;
                    std::vector<int16_t> lv((size_t)b.count * nf * nmb * 384), mvd((size_t)b.count * nf * nmb * 2);
                    std::vector<uint8_t> mpm((size_t)b.count * nf * nmb * 4), ipm((size_t)b.count * nf * nmb * 4);
                    parallel_for(b.count, io, [&](int i) {
                        const auto& q = b.parsed[i];
                        std::copy(q.levels.begin() + (size_t)f0 * nmb * 384, q.levels.begin() + (size_t)(f0 + nf) * nmb * 384, lv.begin() + (size_t)i * nf * nmb * 384);
                        std::copy(q.mvd.begin() + (size_t)f0 * nmb * 2, q.mvd.begin() + (size_t)(f0 + nf) * nmb * 2, mvd.begin() + (size_t)i * nf * nmb * 2);
                        std::copy(q.mpm.begin() + (size_t)f0 * nmb * 4, q.mpm.begin() + (size_t)(f0 + nf) * nmb * 4, mpm.begin() + (size_t)i * nf * nmb * 4);
                        std::copy(q.ipm.begin() + (size_t)f0 * nmb * 4, q.ipm.begin() + (size_t)(f0 + nf) * nmb * 4, ipm.begin() + (size_t)i * nf * nmb * 4);
                    });
                    icsp_dec_in in{lv.data(), mpm.data(), ipm.data(), mvd.data()};
                    return icsp_decode_gops(ctx, &in, b.count * ngops, glen, qdc, qac, dst);
                };
                if (full) rc = call(0, full, ip, b.yuv);
                if (!rc && tail) rc = call(full * ip, 1, tail, b.yuv + (size_t)W * full * ip * fbytes);
            }
            return rc;
        };
        auto store = [&](Buf& b) {
            parallel_for(b.count, io, [&](int i) {
                const std::string& path = inputs[b.first + i];
                const size_t sl = path.find_last_of('/');
                const std::string name = (sl == std::string::npos ? path : path.substr(sl + 1)) + ".yuv";
                FILE* fo = fopen(name.c_str(), "wb");
                bool ok = fo != nullptr;
                if (ok && full) ok = fwrite(b.yuv + (size_t)i * full * ip * fbytes, fbytes, (size_t)full * ip, fo) == (size_t)full * ip;
                if (ok && tail) ok = fwrite(b.yuv + (size_t)W * full * ip * fbytes + (size_t)i * tail * fbytes, fbytes, tail, fo) == (size_t)tail;
                if (fo) fclose(fo);
                if (!ok) { errs[d] = "cannot write " + name; failed++; }
            });
        };
        const int nwaves = (mine + W - 1) / W;
        std::future<void> loading, storing;
        load(buf[0], s_begin, std::min(W, mine));
        for (int wv = 0; wv < nwaves && !failed; wv++) {
            Buf& cur = buf[wv & 1];
            Buf& nxt = buf[(wv + 1) & 1];
            if (storing.valid()) storing.get();
            if (wv + 1 < nwaves) {
                const int fs = s_begin + (wv + 1) * W, count = std::min(W, s_end - fs);
                loading = std::async(std::launch::async, [&, fs, count] { load(nxt, fs, count); });
            }
            const int rc = decode(cur);
            if (rc) bail(std::string(icsp_last_error(ctx)) + " (code " + std::to_string(rc) + ")");
            if (loading.valid()) loading.get();
            if (!rc) storing = std::async(std::launch::async, [&] { store(cur); });
        }
        if (loading.valid()) loading.get();
        if (storing.valid()) storing.get();
        for (auto& b : buf) icsp_host_free(b.yuv);
        icsp_destroy(ctx);
    };
    {
        std::vector<std::thread> th;
        for (int d = 0; d < G; d++) th.emplace_back(gpu_thread, d);
        for (auto& t : th) t.join();
    }
    if (failed) {
        for (int d = 0; d < G; d++) if (!errs[d].empty()) fprintf(stderr, "[ERROR] GPU %d: %s\n", d, errs[d].c_str());
        return 1;
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "icspdec: batch of %d streams x %d frames %dx%d on %d GPU(s): %.3f s wall incl. file reads and writes (%.0f frames/s; %d streams parsed on the GPU)\n",
            S, nframes, w, h, G, sec, (double)S * nframes / sec, (int)on_gpu);
    return 0;
}

int main(int argc, char** argv)
{
    std::vector<std::string> pos;
    int gpus = 1;
    std::string outname, idxname, batch;
    int wave = 8, io_threads = 0;
    bool use_index = true;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "--gpus" && i + 1 < argc) gpus = atoi(argv[++i]);
        else if (a == "--batch" && i + 1 < argc) batch = argv[++i];
        else if (a == "--wave" && i + 1 < argc) wave = atoi(argv[++i]);
        else if (a == "--io-threads" && i + 1 < argc) io_threads = atoi(argv[++i]);
        else if (a == "--index" && i + 1 < argc) idxname = argv[++i];
        else if (a == "--no-index") use_index = false;
        else if (a == "--out" && i + 1 < argc) outname = argv[++i];
        else pos.push_back(a);
    }
    // CUDA initialises every visible device (≈0.7 s each on a multi-GPU box): expose only the ones this run uses
    if (!getenv("CUDA_VISIBLE_DEVICES") && gpus >= 1 && gpus <= 16) {
        std::string vis;
        for (int d = 0; d < gpus; d++) vis += (d ? "," : "") + std::to_string(d);
        setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 0);
    }
    if (!batch.empty()) {
        if (pos.empty()) { fprintf(stderr, "usage: icspdec --batch <list.txt> <nframes> [--gpus G] [--wave S] [--io-threads T] [--no-index]\n"); return 1; }
        return run_batch(batch, atoi(pos[0].c_str()), gpus, wave, io_threads, use_index);
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
