// icspenc — drop-in encoder front end: the reference's CLI (encoder_main.cpp:4-24, ENC:84-176) and frame loop
// (single_thread_encoding ENC:217-245 / the GOP job model of ICSP_thread.cpp:39-77) on the host, the per-frame
// core (intraPrediction / interPrediction) on the GPU through libicspcuda, then the host bitstream writer.
//
//   icspenc -i <name_cif.yuv> -n <frames> [-q Q | --qpdc D --qpac A] [--intraPeriod P] [-w W -h H]
//           [--EnMultiThread T] [--gpus G] [--no-recon] [--psnr] [--index] [--quiet]
//   icspenc --batch <list.txt> -n <frames> ...      N independent streams (one input path per line) in ONE process: the
//           streams are sharded over the GPUs (--gpus), each GPU encodes waves of --wave streams (default 8) with one
//           icsp_encode_streams call per wave — the batch shape bench.py measures — while the next wave is read and the
//           previous one written by --io-threads file threads (double buffering: read -> pinned -> H2D | kernels | D2H ->
//           write).  Outputs per stream: <prefix>_compCIF_<QDC>_<QAC>_<IP>.bin and <prefix>_test_yuv.yuv.
//
// Outputs, like the reference: <prefix>_compCIF_<QDC>_<QAC>_<IP>.bin (prefix = input name up to the first '_',
// encoder_main.cpp:10-17) and test_yuv.yuv (reconstruction, ENC:6376-6421).  Unlike the reference's
// --EnMultiThread mode, the bitstream is always written and tail frames (n % intraPeriod) are encoded.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <fstream>
#include <future>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/icspcuda.h"
#include "bitstream.h"

namespace {
// ICSPENC_TIMING=1: where the wall time of a run goes (stderr)
struct PhaseClock {
    const bool on = getenv("ICSPENC_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char* what)
    {
        const auto n = std::chrono::steady_clock::now();
        if (on) fprintf(stderr, "[icspenc timing] %-34s %8.3f s\n", what, std::chrono::duration<double>(n - t).count());
        t = n;
    }
};
struct Options {
    std::string input;
    int frames = 1;             // README.md:30 documents 1 (the reference leaves it uninitialised, ENC:84-91)
    int qdc = 0, qac = 0, ip = 0, threads = 0, gpus = 1;
    int width = 352, height = 288;  // encoder_main.cpp:20 hard-wires CIF; -w/-h are accepted here
    bool recon = true, quiet = false;
    bool index = false;         // --index: write <bin>.idx, the macroblock-row index icspdec uses to parse on the GPU
    bool psnr = false;          // --psnr: luma PSNR of the reconstruction, reduced on the GPU (what the reference's decoder logs, DEC.h:332-350)
    bool host_entropy = false;  // --host-entropy: download the syntax arrays and entropy-code on the CPU (round-1 path)
    std::string batch;          // --batch <list>: many streams, one process
    int wave = 8, io_threads = 0;
    bool plan = false;          // --plan: print the GPU shard plan (one line per shard) and exit; touches no device
};

void help()
{
    printf("usage: ./icspenc [option] [values]\n"
           "-i : input yuv sequence (planar I420; the name must contain '_')\n"
           "-w : width (default 352)\n-h : height (default 288; without a value: this help)\n"
           "-n : the number of frames(default is 1)\n-q : QP of DC and AC (16, 8, or 1)\n"
           "--qpdc : QP of DC\n--qpac : QP of AC\n--intraPeriod: period of intra frame(0: All intra)\n"
           "--EnMultiThread: host threads for the entropy coder (0 = all cores)\n"
           "--gpus: GPUs to shard the GOPs over (default 1)\n--no-recon: do not write test_yuv.yuv\n"
           "--plan: print the shard plan for --gpus (GOP ranges, or stream ranges with --batch) and exit\n"
           "--batch <list>: encode every stream listed in <list> (one path per line) in one process, sharded over the GPUs;\n"
           "                --wave S streams per device call (default 8), --io-threads T file threads per GPU\n"
           "--host-entropy: entropy-code on the CPU instead of the GPU\n--psnr: print the average luma PSNR (computed on the GPU; works with --no-recon)\n--help : help message\n");
}

bool is_number(const char* s) { return s && *s && strspn(s, "0123456789") == strlen(s); }

int parse(int argc, char** argv, Options& o)
{
    if (argc < 2) { fprintf(stderr, "[ERROR] unenough parameters in parsing_command\n"); return -1; }
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto val = [&](int& dst) { if (i + 1 >= argc) return false; dst = atoi(argv[++i]); return true; };
        if (a == "--help") { help(); exit(0); }
        else if (a == "-h") { if (i + 1 < argc && is_number(argv[i + 1])) o.height = atoi(argv[++i]); else { help(); exit(0); } }
        else if (a == "-w") { if (!val(o.width)) return -1; }
        else if (a == "-i") { if (i + 1 >= argc) return -1; o.input = argv[++i]; }
        else if (a == "-n") { if (!val(o.frames)) return -1; }
        else if (a == "-q") { if (!val(o.qdc)) return -1; o.qac = o.qdc; }
        else if (a == "--qpdc") { if (!val(o.qdc)) return -1; }
        else if (a == "--qpac") { if (!val(o.qac)) return -1; }
        else if (a == "--intraPeriod") { if (!val(o.ip)) return -1; }
        else if (a == "--EnMultiThread") { if (!val(o.threads)) return -1; }
        else if (a == "--gpus") { if (!val(o.gpus)) return -1; }
        else if (a == "--batch") { if (i + 1 >= argc) return -1; o.batch = argv[++i]; }
        else if (a == "--wave") { if (!val(o.wave)) return -1; }
        else if (a == "--io-threads") { if (!val(o.io_threads)) return -1; }
        else if (a == "--no-recon") o.recon = false;
        else if (a == "--host-entropy") o.host_entropy = true;
        else if (a == "--psnr") o.psnr = true;
        else if (a == "--index") o.index = true;
        else if (a == "--plan") o.plan = true;
        else if (a == "--quiet") o.quiet = true;
        else if (a[0] == '-') { fprintf(stderr, "[ERROR] uncorrect parameters in parsing_command\n"); return -1; }
    }
    return 0;
}

struct Shard {          // one GPU's contiguous range of GOPs
    int device = 0;
    int first_frame = 0, n_gops = 0, gop_len = 0;
    int rc = 0;
    std::string err;
    std::vector<std::pair<std::vector<uint8_t>, uint64_t>> bits;   // GPU entropy path: MSB-first bit strings, in order
    std::vector<uint64_t> sse;  // --psnr: [frames][3]
    std::vector<std::vector<uint64_t>> rows;   // --index: per segment, [frames][mbh] bit offsets inside the segment
};

// The two sharding rules of SURVEY §8e (mirrored for the Python side by icspcodec_b200/sharding.py; tests/test_multi_rank_cpu.py
// holds the two to the same plan through --plan).  Closed GOPs are independent jobs (ICSP_thread.cpp:39-77): full GOPs are split
// contiguously over the GPUs, the tail GOP (n % intraPeriod frames) goes to the last one.
std::vector<Shard> gop_shards(int n, int ip, int gpus)
{
    const int gop = ip == 0 ? 1 : ip, full = n / gop, tail = n - full * gop, G = std::max(1, gpus);
    std::vector<Shard> shards;
    for (int d = 0; d < G; d++) {
        const int g0 = (int)((long long)full * d / G), g1 = (int)((long long)full * (d + 1) / G);
        if (g1 > g0) shards.push_back({d, g0 * gop, g1 - g0, gop, 0, "", {}, {}, {}});
    }
    if (tail) shards.push_back({G - 1, full * gop, 1, tail, 0, "", {}, {}, {}});
    return shards;
}
// whole streams per GPU: device d of G takes streams [S*d/G, S*(d+1)/G)
inline std::pair<int, int> stream_shard(int S, int d, int G)
{
    return {(int)((long long)S * d / G), (int)((long long)S * (d + 1) / G)};
}

// ---- batch mode: N independent streams, one process ---------------------------------------------------------------------
// parallel helper: run fn(i) for i in [0, n) on `threads` threads
template <typename F>
void parallel_for(int n, int threads, F fn)
{
    std::atomic<int> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < std::max(1, std::min(threads, n)); t++)
        th.emplace_back([&] { for (int i = next++; i < n; i = next++) fn(i); });
    for (auto& t : th) t.join();
}

std::string stream_prefix(const std::string& path)
{
    const size_t sl = path.find_last_of('/');
    const std::string base = sl == std::string::npos ? path : path.substr(sl + 1);
    const size_t us = base.find('_');
    return us == std::string::npos ? base : base.substr(0, us);      // encoder_main.cpp:10-17: name up to the first '_'
}

int run_batch(const Options& o)
{
    std::vector<std::string> inputs;
    {
        std::ifstream f(o.batch);
        if (!f) { fprintf(stderr, "[ERROR] cannot open %s\n", o.batch.c_str()); return 1; }
        for (std::string line; std::getline(f, line);) {
            while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
            if (!line.empty()) inputs.push_back(line);
        }
    }
    const int S = (int)inputs.size();
    if (S == 0) { fprintf(stderr, "[ERROR] %s lists no streams\n", o.batch.c_str()); return 1; }
    const int n = o.frames, nmbh = o.height / 16;
    const size_t fb = (size_t)o.width * o.height * 3 / 2;
    const int gop = o.ip == 0 ? 1 : o.ip, full = n / gop, tail = n - full * gop;
    const int G = std::max(1, std::min(o.gpus, S));
    if (o.plan) {
        for (int d = 0; d < G; d++) printf("streams device=%d begin=%d end=%d\n", d, stream_shard(S, d, G).first, stream_shard(S, d, G).second);
        return 0;
    }
    const int hw = std::max(1, (int)std::thread::hardware_concurrency());
    const int io = o.io_threads > 0 ? o.io_threads : std::max(2, hw / G);
    std::atomic<int> failed{0};
    std::vector<std::string> errs(G);
    std::vector<double> steady(G, 0.0);      // per GPU: seconds from the first file read to the last file write (no CUDA init / allocation)
    const auto t0 = std::chrono::steady_clock::now();

    auto gpu_thread = [&](int d) {
        const int s_begin = stream_shard(S, d, G).first, s_end = stream_shard(S, d, G).second;
        const int mine = s_end - s_begin;
        if (mine <= 0) return;
        const int W = std::max(1, std::min(o.wave, mine));
        auto bail = [&](const std::string& m) { errs[d] = m; failed++; };
        PhaseClock pc;
        icsp_ctx* ctx = nullptr;
        if (icsp_create(&ctx, d, o.width, o.height, W * std::max(full * gop, tail))) return bail(icsp_last_error(nullptr));
        pc.lap("batch: CUDA init + icsp_create");
        // double buffers: frames in (write-combined pinned), reconstruction and bits out (pinned)
        struct Buf { uint8_t *in = nullptr, *tail_in = nullptr, *rec = nullptr, *bits = nullptr; size_t bits_cap = 0;
                     std::vector<uint64_t> nbits, off, tnbits, toff, rows, trows; uint8_t* tbits = nullptr; size_t tbits_cap = 0; int count = 0, first = 0; } buf[2];
        const int nwaves = (mine + W - 1) / W;
        for (int bi = 0; bi < (nwaves > 1 ? 2 : 1); bi++) {      // the second set only when there is a wave to overlap with
            Buf& b = buf[bi];
            b.in = (uint8_t*)icsp_host_alloc_upload((size_t)W * full * gop * fb + 64);
            if (tail) b.tail_in = (uint8_t*)icsp_host_alloc_upload((size_t)W * tail * fb);
            if (o.recon) b.rec = (uint8_t*)icsp_host_alloc((size_t)W * n * fb);
            b.bits_cap = (size_t)W * full * gop * ((size_t)o.width * o.height + 32) + 64;
            b.bits = (uint8_t*)icsp_host_alloc(b.bits_cap);
            if (tail) { b.tbits_cap = (size_t)W * tail * ((size_t)o.width * o.height + 32) + 64; b.tbits = (uint8_t*)icsp_host_alloc(b.tbits_cap); }
            if (!b.in || !b.bits || (o.recon && !b.rec) || (tail && (!b.tail_in || !b.tbits))) return bail("pinned host allocation failed");
            b.nbits.resize(W); b.off.resize(W); b.tnbits.resize(W); b.toff.resize(W);
        }
        auto load = [&](Buf& b, int first, int count) {      // read `count` streams starting at stream `first` into b
            b.first = first; b.count = count;
            parallel_for(count, io, [&](int i) {
                FILE* fi = fopen(inputs[first + i].c_str(), "rb");
                bool ok = fi != nullptr;
                if (ok && full) ok = fread(b.in + (size_t)i * full * gop * fb, fb, (size_t)full * gop, fi) == (size_t)full * gop;
                if (ok && tail) ok = fread(b.tail_in + (size_t)i * tail * fb, fb, tail, fi) == (size_t)tail;
                if (fi) fclose(fi);
                if (!ok) { errs[d] = "cannot read " + std::to_string(n) + " frames from " + inputs[first + i]; failed++; }
            });
        };
        auto encode = [&](Buf& b) -> int {
            int rc = ICSP_OK;
            if (full) {
                // reconstruction layout of the call = [stream][full*gop frames]; the per-stream file is assembled when writing
                icsp_bits_out bo{b.bits, b.bits_cap, b.nbits.data(), b.off.data(), o.recon ? b.rec : nullptr};
                rc = icsp_encode_streams(ctx, b.in, b.count, full, gop, o.qdc, o.qac, &bo);
                if (rc == ICSP_ERR_CAPACITY) {       // worst-case bound (noise at QP 1): grow the staging buffer once
                    const size_t cap = icsp_bits_bound(o.width, o.height, b.count * full * gop);
                    icsp_host_free(b.bits); b.bits = (uint8_t*)icsp_host_alloc(cap); b.bits_cap = b.bits ? cap : 0;
                    if (!b.bits) return ICSP_ERR_NOMEM;
                    icsp_bits_out bo2{b.bits, b.bits_cap, b.nbits.data(), b.off.data(), o.recon ? b.rec : nullptr};
                    rc = icsp_encode_streams(ctx, b.in, b.count, full, gop, o.qdc, o.qac, &bo2);
                }
                if (!rc && o.index) { b.rows.resize((size_t)b.count * full * gop * nmbh); rc = icsp_bits_row_index(ctx, b.count * full * gop, b.rows.data()); }
            }
            if (!rc && tail) {
                icsp_bits_out bo{b.tbits, b.tbits_cap, b.tnbits.data(), b.toff.data(), o.recon ? b.rec + (size_t)W * full * gop * fb : nullptr};
                rc = icsp_encode_streams(ctx, b.tail_in, b.count, 1, tail, o.qdc, o.qac, &bo);
                if (rc == ICSP_ERR_CAPACITY) {
                    const size_t cap = icsp_bits_bound(o.width, o.height, b.count * tail);
                    icsp_host_free(b.tbits); b.tbits = (uint8_t*)icsp_host_alloc(cap); b.tbits_cap = b.tbits ? cap : 0;
                    if (!b.tbits) return ICSP_ERR_NOMEM;
                    icsp_bits_out bo2{b.tbits, b.tbits_cap, b.tnbits.data(), b.toff.data(), o.recon ? b.rec + (size_t)W * full * gop * fb : nullptr};
                    rc = icsp_encode_streams(ctx, b.tail_in, b.count, 1, tail, o.qdc, o.qac, &bo2);
                }
                if (!rc && o.index) { b.trows.resize((size_t)b.count * tail * nmbh); rc = icsp_bits_row_index(ctx, b.count * tail, b.trows.data()); }
            }
            return rc;
        };
        auto store = [&](Buf& b) {
            parallel_for(b.count, io, [&](int i) {
                const std::string prefix = stream_prefix(inputs[b.first + i]);
                icsp_host::StreamParams sp;
                sp.width = o.width; sp.height = o.height; sp.qp_dc = o.qdc; sp.qp_ac = o.qac; sp.intra_period = o.ip; sp.nframes = n;
                icsp_host::BitString all;
                if (full) all.append_msb_bytes(b.bits + b.off[i], b.nbits[i]);
                if (tail) all.append_msb_bytes(b.tbits + b.toff[i], b.tnbits[i]);
                std::vector<uint8_t> bin = icsp_host::stream_header(sp);
                const std::vector<uint8_t> body = all.reference_body();
                bin.insert(bin.end(), body.begin(), body.end());
                char name[600];
                snprintf(name, sizeof(name), "%s_compCIF_%d_%d_%d.bin", prefix.c_str(), o.qdc, o.qac, o.ip);
                FILE* fo = fopen(name, "wb");
                if (!fo || fwrite(bin.data(), 1, bin.size(), fo) != bin.size()) { errs[d] = std::string("cannot write ") + name; failed++; }
                if (fo) fclose(fo);
                if (o.index) {
                    snprintf(name, sizeof(name), "%s_compCIF_%d_%d_%d.bin.idx", prefix.c_str(), o.qdc, o.qac, o.ip);
                    FILE* fx = fopen(name, "wb");
                    if (fx) {
                        const uint32_t hdr[4] = {(uint32_t)o.width, (uint32_t)o.height, (uint32_t)n, (uint32_t)nmbh};
                        fwrite("ICSPIDX1", 1, 8, fx); fwrite(hdr, 4, 4, fx);
                        if (full) fwrite(b.rows.data() + (size_t)i * full * gop * nmbh, 8, (size_t)full * gop * nmbh, fx);
                        if (tail) {
                            std::vector<uint64_t> tr(b.trows.begin() + (size_t)i * tail * nmbh, b.trows.begin() + (size_t)(i + 1) * tail * nmbh);
                            for (auto& v : tr) v += full ? b.nbits[i] : 0;
                            fwrite(tr.data(), 8, tr.size(), fx);
                        }
                        fclose(fx);
                    } else { errs[d] = std::string("cannot write ") + name; failed++; }
                }
                if (o.recon) {
                    snprintf(name, sizeof(name), "%s_test_yuv.yuv", prefix.c_str());
                    FILE* fr = fopen(name, "wb");
                    bool ok = fr != nullptr;
                    if (ok && full) ok = fwrite(b.rec + (size_t)i * full * gop * fb, fb, (size_t)full * gop, fr) == (size_t)full * gop;
                    if (ok && tail) ok = fwrite(b.rec + (size_t)W * full * gop * fb + (size_t)i * tail * fb, fb, tail, fr) == (size_t)tail;
                    if (fr) fclose(fr);
                    if (!ok) { errs[d] = std::string("cannot write ") + name; failed++; }
                }
            });
        };
        // pipeline over waves: load(i+1) | encode(i) | store(i-1)
        std::future<void> loading, storing;
        pc.lap("batch: pinned buffers");
        const auto t_work = std::chrono::steady_clock::now();
        load(buf[0], s_begin, std::min(W, mine));
        pc.lap("batch: first wave read");
        for (int wv = 0; wv < nwaves && !failed; wv++) {
            Buf& cur = buf[wv & 1];
            Buf& nxt = buf[(wv + 1) & 1];
            if (storing.valid()) storing.get();                      // nxt's outputs have been written
            if (wv + 1 < nwaves) {
                const int first = s_begin + (wv + 1) * W, count = std::min(W, s_end - first);
                loading = std::async(std::launch::async, [&, first, count] { load(nxt, first, count); });
            }
            const int rc = encode(cur);
            if (rc) { bail(std::string(icsp_last_error(ctx)) + " (code " + std::to_string(rc) + ")"); }
            if (loading.valid()) loading.get();
            if (!rc) storing = std::async(std::launch::async, [&] { store(cur); });
        }
        if (loading.valid()) loading.get();
        pc.lap("batch: waves (encode | read | write)");
        if (storing.valid()) storing.get();
        pc.lap("batch: last wave write");
        steady[d] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_work).count();
        for (auto& b : buf) { icsp_host_free(b.in); icsp_host_free(b.tail_in); icsp_host_free(b.rec); icsp_host_free(b.bits); icsp_host_free(b.tbits); }
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
    if (!o.quiet) {
        const double work = *std::max_element(steady.begin(), steady.end());
        fprintf(stderr, "icspenc: batch of %d streams x %d frames on %d GPU(s): %.3f s wall (%.0f frames/s), of which %.3f s read | encode | write "
                        "(%.0f frames/s once CUDA is initialised and the buffers exist)\n", S, n, G, sec, (double)S * n / sec, work, (double)S * n / work);
    }
    return 0;
}
}  // namespace

int main(int argc, char** argv)
{
    Options o;
    if (parse(argc, argv, o)) return 1;
    // CUDA initialises every visible device (≈0.7 s each on a multi-GPU box): expose only the ones this run uses
    if (!getenv("CUDA_VISIBLE_DEVICES") && o.gpus >= 1 && o.gpus <= 16) {
        std::string vis;
        for (int d = 0; d < o.gpus; d++) vis += (d ? "," : "") + std::to_string(d);
        setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 0);
    }
    if (!o.batch.empty()) {
        if (o.frames <= 0 || o.qdc <= 0 || o.qac <= 0 || o.ip < 0 || o.ip > 63 || (o.width & 15) || (o.height & 15) || o.qdc > 255 || o.qac > 255 ||
            o.host_entropy || o.psnr) {
            fprintf(stderr, "[ERROR] uncorrect parameters (--batch needs -n > 0, QP in 1..255, intraPeriod in 0..63; not with --host-entropy/--psnr)\n");
            return 1;
        }
        return run_batch(o);
    }
    if (o.input.empty() || o.frames <= 0 || o.qdc <= 0 || o.qac <= 0 || o.ip < 0 || o.ip > 63 || (o.width & 15) || (o.height & 15) ||
        o.qdc > 255 || o.qac > 255) {
        fprintf(stderr, "[ERROR] uncorrect parameters (need -i, -n > 0, QP in 1..255, intraPeriod in 0..63, width/height multiples of 16)\n");
        return 1;
    }
    if (o.plan) {
        for (const Shard& s : gop_shards(o.frames, o.ip, o.gpus))
            printf("gops device=%d first_frame=%d n_gops=%d gop_len=%d\n", s.device, s.first_frame, s.n_gops, s.gop_len);
        return 0;
    }
    const size_t us = o.input.find('_');
    if (us == std::string::npos) { fprintf(stderr, "[ERROR] input file name must contain '_' (encoder_main.cpp:13)\n"); return 1; }
    const std::string slash_stripped = o.input.substr(0, us);
    const int nmb = (o.width / 16) * (o.height / 16);
    const size_t fb = (size_t)o.width * o.height * 3 / 2;
    const int n = o.frames;

    PhaseClock pc;
    FILE* fi = fopen(o.input.c_str(), "rb");
    if (!fi) { fprintf(stderr, "[ERROR] cannot open %s\n", o.input.c_str()); return 1; }
    uint8_t* frames = (uint8_t*)icsp_host_alloc_upload((size_t)n * fb);   // the CPU only writes it (fread), the GPUs read it
    pc.lap("CUDA init + pinned input buffer");
    if (!frames) { fprintf(stderr, "[ERROR] fail memory allocation (pinned host, %zu bytes); is a CUDA device present? libicspcuda has no CPU fallback\n", (size_t)n * fb); return 1; }
    if (fread(frames, fb, n, fi) != (size_t)n) { fprintf(stderr, "[ERROR] %s holds fewer than %d frames of %dx%d\n", o.input.c_str(), n, o.width, o.height); return 1; }
    fclose(fi);
    pc.lap("read input file");

    // SoA outputs (pinned)
    const size_t N = (size_t)n * nmb;
    int16_t *levels = nullptr, *mvd = nullptr;
    uint8_t *acflag = nullptr, *mpm = nullptr, *ipm = nullptr;
    if (o.host_entropy) {
        levels = (int16_t*)icsp_host_alloc(N * 384 * 2);
        acflag = (uint8_t*)icsp_host_alloc(N * 6);
        mpm = (uint8_t*)icsp_host_alloc(N * 4);
        ipm = (uint8_t*)icsp_host_alloc(N * 4);
        mvd = (int16_t*)icsp_host_alloc(N * 4);
        if (!levels || !acflag || !mpm || !ipm || !mvd) { fprintf(stderr, "[ERROR] fail memory allocation (pinned host)\n"); return 1; }
    }
    uint8_t* recon = o.recon ? (uint8_t*)icsp_host_alloc((size_t)n * fb) : nullptr;
    if (o.recon && !recon) { fprintf(stderr, "[ERROR] fail memory allocation (pinned host)\n"); return 1; }

    pc.lap("pinned output buffers");
    const auto t0 = std::chrono::steady_clock::now();
    // frame loop: I-frame iff n % intraPeriod == 0 (0 = all intra).  Closed GOPs are independent jobs
    // (ICSP_thread.cpp:39-77): full GOPs are sharded contiguously over the GPUs, the tail GOP goes to the last one.
    const int G = std::max(1, o.gpus);
    std::vector<Shard> shards = gop_shards(n, o.ip, G);
    auto run_shard = [&](Shard& s) {
        const int cnt = s.n_gops * s.gop_len;
        int call_frames = 4096;                                    // bound device memory per call
        if (const char* e = getenv("ICSPENC_CALL_FRAMES")) call_frames = std::max(1, atoi(e));
        const int per_call_gops = std::max(1, std::min(s.n_gops, call_frames / s.gop_len));
        icsp_ctx* ctx = nullptr;
        s.rc = icsp_create(&ctx, s.device, o.width, o.height, per_call_gops * s.gop_len);
        if (s.rc) { s.err = icsp_last_error(nullptr); return; }
        uint8_t* stage = nullptr;
        size_t stage_cap = 0;
        for (int g = 0; g < s.n_gops && !s.rc; g += per_call_gops) {
            const int ng = std::min(per_call_gops, s.n_gops - g);
            const size_t f0 = (size_t)s.first_frame + (size_t)g * s.gop_len;
            if (o.host_entropy) {
                icsp_enc_out out{levels + f0 * nmb * 384, acflag + f0 * nmb * 6, mpm + f0 * nmb * 4, ipm + f0 * nmb * 4,
                                 mvd + f0 * nmb * 2, nullptr, nullptr, recon ? recon + f0 * fb : nullptr, nullptr};
                s.rc = icsp_encode_gops(ctx, frames + f0 * fb, ng, s.gop_len, o.qdc, o.qac, &out);
            } else {   // entropy coding + bit packing on the GPU: only bits (and the reconstruction) cross PCIe
                // pinned staging buffer, reused by the shard's calls.  First sized like the reference's own buffer (width*height
                // bytes per frame, ENC:4874: plenty for natural content); a call that does not fit is repeated with the worst-case bound
                const size_t cnt_frames = (size_t)ng * s.gop_len;
                size_t cap = cnt_frames * ((size_t)o.width * o.height + 32) + 64;
                uint64_t nbits = 0, off = 0;
                for (int attempt = 0; attempt < 2; attempt++) {
                    if (stage_cap < cap) {
                        icsp_host_free(stage);
                        stage = (uint8_t*)icsp_host_alloc(cap);
                        stage_cap = stage ? cap : 0;
                        if (!stage) { s.rc = ICSP_ERR_NOMEM; s.err = "pinned staging buffer for the bitstream"; break; }
                    }
                    icsp_bits_out bo{stage, stage_cap, &nbits, &off, recon ? recon + f0 * fb : nullptr};
                    s.rc = icsp_encode_streams(ctx, frames + f0 * fb, 1, ng, s.gop_len, o.qdc, o.qac, &bo);
                    if (s.rc != ICSP_ERR_CAPACITY) break;
                    cap = icsp_bits_bound(o.width, o.height, (int)cnt_frames);
                    if (stage_cap >= cap) break;             // already at the bound: a real capacity error
                }
                if (s.rc == ICSP_ERR_NOMEM) break;
                std::vector<uint8_t> buf;
                if (!s.rc) buf.assign(stage + off, stage + off + (size_t)((nbits + 7) / 8));
                if (!s.rc && o.index) {
                    std::vector<uint64_t> r((size_t)ng * s.gop_len * (o.height / 16));
                    s.rc = icsp_bits_row_index(ctx, ng * s.gop_len, r.data());
                    s.rows.push_back(std::move(r));
                }
                if (!s.rc) s.bits.emplace_back(std::move(buf), nbits);
            }
            if (!s.rc && o.psnr) {   // the frames and their reconstruction are still resident
                const size_t at = s.sse.size();
                s.sse.resize(at + (size_t)ng * s.gop_len * 3);
                s.rc = icsp_enc_sse(ctx, ng * s.gop_len, s.sse.data() + at);
            }
            if (s.rc) s.err = icsp_last_error(ctx);
        }
        (void)cnt;
        icsp_host_free(stage);
        icsp_destroy(ctx);
    };
    {
        // one host thread per GPU (shards of the same device run back to back on that thread)
        std::vector<std::thread> th;
        for (int d = 0; d < G; d++)
            th.emplace_back([&, d] { for (auto& s : shards) if (s.device == d) run_shard(s); });
        for (auto& t : th) t.join();
    }
    for (auto& s : shards)
        if (s.rc) { fprintf(stderr, "[ERROR] GPU %d: %s (code %d)\n", s.device, s.err.c_str(), s.rc); return 1; }
    const auto t1 = std::chrono::steady_clock::now();
    pc.lap("contexts + device calls");
    if (!o.quiet)
        for (int f = 0; f < n; f++) printf("Encoding FRAME_%03d(%c) done!\n", f, (o.ip == 0 || f % o.ip == 0) ? 'I' : 'P');

    // side-car: "ICSPIDX1", width, height, frames, macroblock rows (u32 LE), then u64 LE bit offsets from the body start
    auto write_index = [&](const std::vector<uint64_t>& index) {
        char iname[520];
        snprintf(iname, sizeof(iname), "%s_compCIF_%d_%d_%d.bin.idx", slash_stripped.c_str(), o.qdc, o.qac, o.ip);
        FILE* fx = fopen(iname, "wb");
        if (!fx) { fprintf(stderr, "fail to open %s\n", iname); return false; }
        const uint32_t hdr[4] = {(uint32_t)o.width, (uint32_t)o.height, (uint32_t)n, (uint32_t)(o.height / 16)};
        fwrite("ICSPIDX1", 1, 8, fx); fwrite(hdr, 4, 4, fx); fwrite(index.data(), 8, index.size(), fx);
        fclose(fx);
        return true;
    };
    icsp_host::StreamParams sp;
    sp.width = o.width; sp.height = o.height; sp.qp_dc = o.qdc; sp.qp_ac = o.qac; sp.intra_period = o.ip; sp.nframes = n;
    std::vector<uint8_t> bin;
    if (o.host_entropy) {
        icsp_host::Syntax syn{levels, acflag, mpm, ipm, mvd};
        const int hw = (int)std::thread::hardware_concurrency();
        std::vector<uint64_t> index;
        bin = icsp_host::write_stream(sp, syn, o.threads > 0 ? o.threads : std::max(1, hw), o.index ? &index : nullptr);
        if (o.index && !write_index(index)) return 1;
    } else {   // concatenate the per-shard bit strings in frame order (frames are not byte aligned in the stream, H7)
        std::sort(shards.begin(), shards.end(), [](const Shard& a, const Shard& b) { return a.first_frame < b.first_frame; });
        icsp_host::BitString all;
        std::vector<uint64_t> index;
        uint64_t base = 0;
        for (auto& s : shards)
            for (size_t i = 0; i < s.bits.size(); i++) {
                all.append_msb_bytes(s.bits[i].first.data(), s.bits[i].second);
                if (o.index) for (uint64_t r : s.rows[i]) index.push_back(base + r);
                base += s.bits[i].second;
            }
        if (o.index && !write_index(index)) return 1;
        bin = icsp_host::stream_header(sp);
        const std::vector<uint8_t> body = all.reference_body();
        bin.insert(bin.end(), body.begin(), body.end());
    }
    const auto t2 = std::chrono::steady_clock::now();

    char name[512];
    snprintf(name, sizeof(name), "%s_compCIF_%d_%d_%d.bin", slash_stripped.c_str(), o.qdc, o.qac, o.ip);
    FILE* fo = fopen(name, "wb");
    if (!fo) { fprintf(stderr, "fail to open %s\n", name); return 1; }
    fwrite(bin.data(), 1, bin.size(), fo);
    fclose(fo);
    if (o.recon) {
        FILE* fr = fopen("test_yuv.yuv", "wb");
        if (!fr) { fprintf(stderr, "fail to open test_yuv.yuv\n"); return 1; }
        fwrite(recon, fb, n, fr);
        fclose(fr);
    }
    pc.lap("bit concatenation + file writes");
    if (o.psnr) {   // mean over frames of 20*log10(255/sqrt(MSE_Y)), the reference decoder's figure (DEC.h:332-348)
        double acc = 0;
        for (auto& s : shards)
            for (size_t f = 0; f * 3 < s.sse.size(); f++) acc += 20. * log10(255. / sqrt((double)s.sse[f * 3] / ((double)o.width * o.height)));
        printf("PSNR: %.4lf QPDC: %d  QPAC: %d Period: %d\n", acc / n, o.qdc, o.qac, o.ip);
    }
    if (!o.quiet) {
        const double core = std::chrono::duration<double>(t1 - t0).count(), ent = std::chrono::duration<double>(t2 - t1).count();
        fprintf(stderr, "icspenc: %d frames, GPU core %.3f s (%.0f fps), entropy+bitstream %.3f s, %zu bytes -> %s\n", n, core, n / core, ent,
                bin.size(), name);
    }
    icsp_host_free(frames); icsp_host_free(levels); icsp_host_free(acflag); icsp_host_free(mpm); icsp_host_free(ipm);
    icsp_host_free(mvd); icsp_host_free(recon);
    return 0;
}
