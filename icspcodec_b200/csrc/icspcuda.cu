// libicspcuda: C ABI (include/icspcuda.h) over the sm_100a kernels in icsp_kernels.cuh.
// Host side only orchestrates: one context = one GPU, one stream, SoA device buffers sized for max_frames.
#include "../../include/icspcuda.h"
#include "icsp_kernels.cuh"
#include "icsp_transform.cuh"
#include "icsp_entropy.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <tuple>
#include <string>
#include <vector>

using namespace icsp;

namespace {

thread_local char g_create_error[256] = "";

enum KernelId {
    K_INTRA_ENC = 0, K_INTRA_DEC, K_FDCT, K_DCCHAIN, K_IDCT_ENC, K_IDCT_DEC, K_ME_SAD, K_ME_ZERO, K_ME_CHAIN, K_ME_FIXUP,
    K_MV_RECON, K_DCT_SHIM, K_IDCT_SHIM, K_EN_SIZE, K_EN_FSCAN, K_EN_SSCAN, K_EN_ZERO, K_EN_PACK, K_FDCT_C, K_IDCT_ENC_C,
    K_IDCT_DEC_C, K_SSE, K_ROWIDX, K_PARSE, K_QUANT_SHIM, K_COUNT
};
const char* const kKernelNames[K_COUNT] = {
    "intra_luma_kernel<enc>", "intra_luma_kernel<dec>", "fdct_quant_kernel", "dc_chain_kernel", "idct_recon_kernel<enc>",
    "idct_recon_kernel<dec>", "me_sad_kernel", "me_zero_kernel", "me_chain_kernel", "me_sad_kernel(fixup)",
    "mv_recon_kernel", "dct8x8_kernel", "idct8x8_kernel", "entropy_size_kernel", "entropy_frame_scan_kernel",
    "entropy_stream_scan_kernels", "entropy_zero_kernel", "entropy_pack_kernel", "fdct_quant_kernel(intra: chroma only)",
    "idct_recon_kernel<enc>(intra)", "idct_recon_kernel<dec>(intra)", "plane_sse_kernel", "row_index_kernel", "parse_rows_kernel",
    "quant8x8_kernel"};

struct Pending { int k; cudaEvent_t a, b; cudaStream_t s; };

}  // namespace

constexpr int MAX_CSTREAMS = 8;
struct icsp_ctx {
    int device = 0;
    Geom g{};
    int cap = 0;  // frames
    cudaStream_t stream = nullptr;          // main stream: everything is ordered with respect to it
    cudaStream_t cstream[MAX_CSTREAMS] = {};           // compute streams (GOP chunks round-robin: latency-bound kernels of one
                                            // chunk overlap throughput-bound kernels of another)
    cudaStream_t s_up = nullptr, s_down = nullptr;   // H2D / D2H copy streams of the pipelined one-shot calls
    cudaEvent_t ev_fork = nullptr, ev_join[MAX_CSTREAMS] = {};
    cudaStream_t hstream[MAX_CSTREAMS] = {};           // high-priority companions of cstream[] for the latency-bound kernels
    cudaEvent_t ev_hi[MAX_CSTREAMS][2] = {};
    bool hi_prio = true;
    std::vector<cudaEvent_t> ev_chunk;      // per-chunk upload / compute-done events
    int n_cstreams = 4;
    int chunk_gops_target = 0;              // 0 = automatic
    // device SoA
    uint8_t *d_cur = nullptr, *d_rec = nullptr, *d_acflag = nullptr, *d_mpm = nullptr, *d_ipm = nullptr;
    int16_t *d_levels = nullptr, *d_mvd = nullptr, *d_mv = nullptr;
    int32_t* d_minsad = nullptr;
    double* d_dcraw = nullptr;
    int32_t* d_dcrec = nullptr;
    uint8_t *d_mestate = nullptr, *d_memoves = nullptr;
    uint32_t* d_meflag = nullptr;
    unsigned long long* d_mezero = nullptr;
    void* d_shim = nullptr;
    size_t shim_bytes = 0;
    double* d_dct_tap = nullptr;              // debug tap of the forward DCT (only while icsp_encode_gops_tap runs)
    bool tr_v1 = false;                       // ICSP_TR_V1=1: first-generation transform kernels (A/B comparisons)
    int skew = 0;                             // ICSP_SKEW: pipeline stage (4*step + stage + 1) after which the next chunk may start; 0 = off
    // entropy coder (allocated on first use)
    uint32_t* d_blkbits = nullptr;
    unsigned long long *d_framebits = nullptr, *d_streambits = nullptr, *d_streamoff = nullptr, *d_chunktotal = nullptr;
    uint32_t* d_overflow = nullptr;
    uint8_t* d_bits = nullptr;
    unsigned long long* d_rows = nullptr;     // [cap][mbh] macroblock-row bit offsets (row index / GPU bit reader)
    unsigned long long* d_sse = nullptr;      // [cap][3] plane SSE (allocated on first use)
    unsigned long long* h_tables = nullptr;   // pinned: [cap] stream bits, [cap] stream offsets, [64] chunk totals, [64] overflow
    struct EnChunk { int s0, ns; size_t f0, nf; unsigned long long region_off, region_cap; };
    std::vector<EnChunk> en_chunks;
    // instrumentation
    bool profiling = false;
    uint64_t launches = 0;
    double total_ms[K_COUNT] = {};
    uint64_t count[K_COUNT] = {};
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> free_events;
    cudaEvent_t slots[8] = {};
    char err[256] = "";
    bool async_err = false;
    // CUDA graphs of the per-chunk kernel sequences, keyed by everything that shapes them (see run_captured)
    struct GraphEntry { cudaGraphExec_t exec = nullptr; uint64_t launches = 0; uint64_t count[32] = {}; };
    std::map<std::tuple<int, int, int, int, int, int, int>, GraphEntry> graphs;
    bool use_graphs = true;                   // ICSP_GRAPHS=0: plain stream launches                   // a CUDA call inside launch bookkeeping failed (event creation, stream join)
    size_t me_smem = 0, me_frame_smem = 0, intra_smem = 0, chain_smem = 0;
    bool me_persistent = true;
    int intra_wide_max_g = 296;               // up to this many I frames per launch the wavefront kernel runs with 192 threads per frame (ICSP_INTRA_WIDE_G)
    int me_rows_max_g = 80;                   // up to this many (frame, segment) pairs per launch the search is row-parallel (ICSP_ME_ROWS_G)
    bool me_fused = true;                     // ICSP_ME_FUSED=0: exact carried-state fallback as three separate launches (A/B)
    int chain_staged = 1;
    bool slice_copies = true;                 // one-chunk host calls: copies sliced by step (ICSP_SLICE_COPIES=0: whole-chunk copies)
    float dc_amb_eps = 9.313225746154785e-10f;  // 2^-30, see dc_chain_kernel; ICSP_DC_EPS=1 sends every block down the double path (tests)
    unsigned char* d_intra_edges = nullptr;   // HD frames: per-GOP edge/DC/mode maps of the intra wavefront in global memory
    size_t intra_edge_stride = 0;
    MeLayout me{};
};

namespace {

int fail(icsp_ctx* c, int code, const char* fmt, ...)
{
    char* dst = c ? c->err : g_create_error;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 256, fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) return fail(c, ICSP_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

int geom_init(Geom& g, int w, int h)
{
    if (w <= 0 || h <= 0 || (w & 15) || (h & 15)) return -1;
    g.w = w; g.h = h; g.mbw = w / 16; g.mbh = h / 16; g.nmb = g.mbw * g.mbh;
    g.cw = w / 2; g.ch = h / 2; g.bw = w / 8; g.bh = h / 8; g.fb = w * h * 3 / 2;
    g.magic_bw = (unsigned)((0x100000000ull + g.bw - 1) / g.bw);
    g.magic_mbw = (unsigned)((0x100000000ull + g.mbw - 1) / g.mbw);
    g.magic_mbw2 = g.mbw > 2 ? (unsigned)((0x100000000ull + g.mbw - 3) / (g.mbw - 2)) : 0u;
    for (int k = 0; k < 4; k++) { g.d_pix[k] = (k >> 1) * 8 * g.w + (k & 1) * 8; g.d_dc[k] = (k >> 1) * g.bw + (k & 1); }
    g.d_pix[4] = 0; g.d_pix[5] = g.cw * g.ch;
    g.d_dc[4] = 4 * g.nmb; g.d_dc[5] = 5 * g.nmb;
    return 0;
}

// ---- constant tables ---------------------------------------------------------------------------------
int upload_tables(icsp_ctx* c)
{
    static const double LIT[8] = {1.0, 0.980785, 0.92388, 0.83147, 0.707107, 0.55557, 0.382683, 0.19509};
    // costable[u][x] = sign * LIT[k] with k = angle index (kTI/kTS in icsp_device.cuh); only the 7 magnitudes are uploaded
    static const unsigned char ZZ[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                         41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                         30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    unsigned char IZ[64];
    for (int k = 0; k < 64; k++) IZ[ZZ[k]] = (unsigned char)k;
    // spiral visiting order for each of the 8 carried start states (motionEstimation ENC:2094-2143)
    signed char cand[8][64][2];
    unsigned char next[8][65];
    for (int s = 0; s < 8; s++) {
        int flag = s & 1, xflag = (s & 2) ? -1 : 1, yflag = (s & 4) ? 1 : -1;
        int x = 0, y = 0, xcnt = 0, ycnt = 0;
        next[s][0] = (unsigned char)s;
        for (int it = 0; it < 64; it++) {
            if (!flag) { if (xflag <= 0) x += xcnt; else x -= xcnt; flag = 1; xcnt++; xflag = -xflag; }
            else { if (yflag < 0) y += ycnt; else y -= ycnt; flag = 0; ycnt++; yflag = -yflag; }
            cand[s][it][0] = (signed char)x;
            cand[s][it][1] = (signed char)y;
            next[s][it + 1] = (unsigned char)(flag | (xflag < 0 ? 2 : 0) | (yflag > 0 ? 4 : 0));
        }
    }
    {
        double mag[2][8];
        for (int t = 0; t < 2; t++) {
            mag[t][0] = 1.0 / std::sqrt(2.0);   // irt2 (ENC.h:199)
            for (int k = 1; k < 8; k++) mag[t][k] = t == 0 ? (double)(float)LIT[k] : LIT[k];
        }
        CU(cudaMemcpyToSymbol(g_mag, mag, sizeof(mag)));
        double mag2[2][16];
        for (int t = 0; t < 2; t++)
            for (int k = 0; k < 8; k++) { mag2[t][k] = mag[t][k]; mag2[t][8 + k] = 0.25 * mag[t][k]; }   // exact scaling
        CU(cudaMemcpyToSymbol(g_mag2, mag2, sizeof(mag2)));
    }
    {
        uint2 izcol[8], izrow[8];
        for (int r = 0; r < 8; r++) {
            unsigned cl = 0, ch = 0, rl = 0, rh = 0;
            for (int i = 0; i < 4; i++) {
                cl |= (unsigned)IZ[i * 8 + r] << (8 * i); ch |= (unsigned)IZ[(4 + i) * 8 + r] << (8 * i);
                rl |= (unsigned)IZ[r * 8 + i] << (8 * i); rh |= (unsigned)IZ[r * 8 + 4 + i] << (8 * i);
            }
            izcol[r] = make_uint2(cl, ch); izrow[r] = make_uint2(rl, rh);
        }
        CU(cudaMemcpyToSymbol(g_izcol, izcol, sizeof(izcol)));
        CU(cudaMemcpyToSymbol(g_izrow, izrow, sizeof(izrow)));
    }
    CU(cudaMemcpyToSymbol(c_cand, cand, sizeof(cand)));
    CU(cudaMemcpyToSymbol(c_next, next, sizeof(next)));
    // (round, lane) slots of the ME kernel: candidate -> shared-memory bank residue of its first word, for a row pitch
    // == 8 (mod 32) words and copy offsets {0,1,1,1} (see me_sad_kernel).  Two candidates per residue -> one per round.
    uint32_t slot[8][2][32];
    for (int s = 0; s < 8; s++) {
        int sl[2][32];
        for (auto& r : sl) for (int& v : r) v = -1;
        std::vector<int> left;
        for (int idx = 0; idx < 64; idx++) {
            const int dx = cand[s][idx][0], dy = cand[s][idx][1];
            const int res = ((((16 + dx) & 3) ? 1 : 0) + (16 + dy) * 8 + ((16 + dx) >> 2)) & 31;
            if (sl[0][res] < 0) sl[0][res] = idx;
            else if (sl[1][res] < 0) sl[1][res] = idx;
            else left.push_back(idx);
        }
        for (int idx : left)
            for (int q = 0; q < 64; q++)
                if (sl[q >> 5][q & 31] < 0) { sl[q >> 5][q & 31] = idx; break; }
        for (int r = 0; r < 2; r++)
            for (int l = 0; l < 32; l++) {
                const int idx = sl[r][l];
                slot[s][r][l] = (uint32_t)idx | ((uint32_t)(uint8_t)cand[s][idx][0] << 8) | ((uint32_t)(uint8_t)cand[s][idx][1] << 16);
            }
    }
    CU(cudaMemcpyToSymbol(g_slot, slot, sizeof(slot)));
    return ICSP_OK;
}

// ---- launch bookkeeping ------------------------------------------------------------------------------
cudaEvent_t get_event(icsp_ctx* c)
{
    if (!c->free_events.empty()) { cudaEvent_t e = c->free_events.back(); c->free_events.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) { c->async_err = true; return nullptr; }   // reported by the next icsp_sync
    return e;
}
void fold_pending(icsp_ctx* c)
{
    if (c->pending.empty()) return;
    cudaDeviceSynchronize();
    static const bool dump = getenv("ICSP_KERNEL_TIMELINE") != nullptr;   // start/end of every launch relative to the first one (tools/kernel_timeline.py)
    for (auto& p : c->pending) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, p.a, p.b);
        if (dump) {
            float t0 = 0.f;
            cudaEventElapsedTime(&t0, c->pending.front().a, p.a);
            int sid = -1;
            for (int i = 0; i < MAX_CSTREAMS; i++) { if (c->cstream[i] == p.s) sid = i; if (c->hstream[i] == p.s) sid = 100 + i; }
            fprintf(stderr, "[icsp kt] %-40s stream %3d start %9.4f end %9.4f\n", kKernelNames[p.k], sid, t0, t0 + ms);
        }
        c->total_ms[p.k] += ms;
        c->free_events.push_back(p.a);
        c->free_events.push_back(p.b);
    }
    c->pending.clear();
}
struct LaunchScope {
    icsp_ctx* c; int k; cudaStream_t s; cudaEvent_t a = nullptr, b = nullptr;
    LaunchScope(icsp_ctx* c_, int k_, cudaStream_t s_ = nullptr) : c(c_), k(k_), s(s_ ? s_ : c_->stream)
    {
        c->launches++; c->count[k]++;
        if (c->profiling) {
            a = get_event(c); b = get_event(c);
            if (a && b) cudaEventRecord(a, s);
        }
    }
    ~LaunchScope()
    {
        if (c->profiling && a && b) { cudaEventRecord(b, s); c->pending.push_back({k, a, b, s}); }
    }
};

Step make_step(int gop_len, int t, int qdc, int qac)
{
    Step st{};
    st.gop_len = gop_len; st.t = t; st.qdc = qdc; st.qac = qac; st.intra = t == 0 ? 1 : 0;
    st.magic_ac = (unsigned)((0x80000000ull + qac - 1) / qac);
    st.magic_dc = (unsigned)((0x80000000ull + qdc - 1) / qdc);
    st.qmode_ac = qac <= 128 ? 0 : 1;                 // div_m19 is exact while 4080*qac < 2^19
    st.m19_ac = (1 << 19) / qac + 1;
    return st;
}

// The kernels of one chunk (every step of its GOPs, the fork / join to the high-priority companion stream included) are
// captured ONCE per shape into a CUDA graph and replayed afterwards: a chunk is 60-130 launches and event operations, i.e.
// ~0.5 ms of host time, which is what a small batch (one 300-frame stream = 30 GOPs per launch) spends most of its time on.
// key = (kind, first GOP / stream, count, gop_len, qp_dc, qp_ac, extra): device pointers follow from the first GOP, so
// equal keys mean equal kernel arguments.  Not used while per-launch profiling events, the DCT tap or experiments are on.
template <typename F>
int run_captured(icsp_ctx* c, std::tuple<int, int, int, int, int, int, int> key, cudaStream_t s, F&& enqueue)
{
    if (!c->use_graphs || c->profiling || c->d_dct_tap || c->skew) return enqueue();
    auto it = c->graphs.find(key);
    if (it == c->graphs.end()) {
        if (c->graphs.size() >= 1024) {                       // a long-lived context fed ever new shapes: start over instead of growing
            CU(cudaDeviceSynchronize());
            for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second.exec);
            c->graphs.clear();
        }
        const uint64_t l0 = c->launches;
        uint64_t c0[K_COUNT];
        for (int k = 0; k < K_COUNT; k++) c0[k] = c->count[k];
        cudaGraph_t graph = nullptr;
        CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue();
        const cudaError_t e = cudaStreamEndCapture(s, &graph);
        if (rc != ICSP_OK || e != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            return rc != ICSP_OK ? rc : fail(c, ICSP_ERR_CUDA, "CUDA graph capture failed: %s", cudaGetErrorString(e));
        }
        icsp_ctx::GraphEntry ent;
        const cudaError_t ei = cudaGraphInstantiate(&ent.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) return fail(c, ICSP_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ei));
        ent.launches = c->launches - l0;
        for (int k = 0; k < K_COUNT; k++) ent.count[k] = c->count[k] - c0[k];
        c->launches = l0;                                     // counted again below, like every replay
        for (int k = 0; k < K_COUNT; k++) c->count[k] = c0[k];
        it = c->graphs.emplace(key, ent).first;
    }
    CU(cudaGraphLaunch(it->second.exec, s));
    c->launches += it->second.launches;
    for (int k = 0; k < K_COUNT; k++) c->count[k] += it->second.count[k];
    return ICSP_OK;
}

// pointers of the chunk that starts at GOP g0 (frame g0*gop_len): kernels index GOPs from 0 inside a chunk
FramePtrs frame_ptrs(icsp_ctx* c, int g0 = 0, int gop_len = 1)
{
    const size_t f0 = (size_t)g0 * gop_len, nmb = (size_t)c->g.nmb, G0 = (size_t)g0;
    FramePtrs p;
    p.cur = c->d_cur + f0 * c->g.fb; p.rec = c->d_rec + f0 * c->g.fb; p.levels = c->d_levels + f0 * nmb * 384;
    p.acflag = c->d_acflag + f0 * nmb * 6; p.mpm = c->d_mpm + f0 * nmb * 4; p.ipm = c->d_ipm + f0 * nmb * 4;
    p.mvd = c->d_mvd + f0 * nmb * 2; p.mv = c->d_mv + f0 * nmb * 2; p.minsad = c->d_minsad + f0 * nmb;
    p.dcraw = c->d_dcraw + G0 * nmb * 6; p.dcrec = c->d_dcrec + G0 * nmb * 6;
    p.mestate = c->d_mestate + G0 * nmb; p.memoves = c->d_memoves + G0 * nmb; p.meflag = c->d_meflag + G0;
    p.mezero = c->d_mezero + G0 * nmb * 8;
    p.dct_tap = c->d_dct_tap ? c->d_dct_tap + f0 * nmb * 384 : nullptr;
    return p;
}

int check_run(icsp_ctx* c, int n_gops, int gop_len, int qdc, int qac)
{
    if (!c) return ICSP_ERR_PARAM;
    if (n_gops <= 0 || gop_len <= 0 || qdc <= 0 || qac <= 0) return fail(c, ICSP_ERR_PARAM, "bad n_gops/gop_len/qp");
    if (n_gops > 65535) return fail(c, ICSP_ERR_PARAM, "n_gops > 65535 per call");
    if ((long long)n_gops * gop_len > c->cap) return fail(c, ICSP_ERR_CAPACITY, "%d frames > capacity %d", n_gops * gop_len, c->cap);
    return ICSP_OK;
}

MeLayout me_layout(const Geom& g)
{
    MeLayout L;
    int max_seg = 22;   // macroblocks per CTA; smaller segments = more resident CTAs per SM (staging overlaps compute)
    if (const char* e = getenv("ICSP_ME_SEG")) { const int v = atoi(e); if (v >= 1 && v <= 22) max_seg = v; }
    L.nseg = (g.mbw + max_seg - 1) / max_seg;
    L.seg_mbs = (g.mbw + L.nseg - 1) / L.nseg;
    L.row_w = (L.seg_mbs * 16 + 32) / 4;
    L.pitch_w = L.row_w + 1;
    while ((L.pitch_w & 31) != 8) L.pitch_w++;
    L.copy_w = 48 * L.pitch_w;
    return L;
}

// The DC chains and the intra wavefront are latency-bound: a few long-lived CTAs per frame, ~115 dependent waves, a few
// percent of the issue slots.  On the chunk's own stream they would queue behind the resident CTAs of the other chunks'
// throughput kernels (which fill the register file) and then run alone.  hi_begin/hi_end move one launch to the chunk's
// high-priority companion stream (fork/join with events): its CTAs are dispatched as soon as slots free up and run
// underneath the transform / ME kernels of the other chunks.
cudaStream_t hi_begin(icsp_ctx* c, cudaStream_t s, int& idx)
{
    idx = -1;
    if (!c->hi_prio) return s;
    for (int i = 0; i < MAX_CSTREAMS; i++) if (c->cstream[i] == s) idx = i;
    if (idx < 0) return s;
    if (cudaEventRecord(c->ev_hi[idx][0], s) != cudaSuccess || cudaStreamWaitEvent(c->hstream[idx], c->ev_hi[idx][0], 0) != cudaSuccess) { idx = -1; return s; }
    return c->hstream[idx];
}
void hi_end(icsp_ctx* c, cudaStream_t s, int idx)
{
    if (idx < 0) return;
    // a failed join would let the chunk's stream run ahead of the companion kernel: remember it, icsp_sync reports it
    if (cudaEventRecord(c->ev_hi[idx][1], c->hstream[idx]) != cudaSuccess || cudaStreamWaitEvent(s, c->ev_hi[idx][1], 0) != cudaSuccess) c->async_err = true;
}

template <bool INTRA, bool QFAST>
void launch_fdct2_t(const Geom& g, const FramePtrs& p, const Step& st, dim3 grid, cudaStream_t s)
{
    if (p.dct_tap) fdct_quant_kernel2<INTRA, QFAST, true><<<grid, TR2_THREADS, 0, s>>>(g, p, st);
    else fdct_quant_kernel2<INTRA, QFAST, false><<<grid, TR2_THREADS, 0, s>>>(g, p, st);
}
void launch_fdct2(const Geom& g, const FramePtrs& p, const Step& st, dim3 grid, cudaStream_t s)
{
    if (st.intra) { if (st.qmode_ac == 0) launch_fdct2_t<true, true>(g, p, st, grid, s); else launch_fdct2_t<true, false>(g, p, st, grid, s); }
    else { if (st.qmode_ac == 0) launch_fdct2_t<false, true>(g, p, st, grid, s); else launch_fdct2_t<false, false>(g, p, st, grid, s); }
}

// motion estimation of step t for every GOP: speculative state-0 search, then the exact carried-state fallback
// (no-ops unless some search of the frame broke early)
int launch_me(icsp_ctx* c, const FramePtrs& p, const Step& st, int G, cudaStream_t s)
{
    const Geom& g = c->g;
    const MeLayout& L = c->me;
    dim3 grid(g.mbh * L.nseg, G);
    const int threads = L.seg_mbs * 32;
    // Big batches: one persistent CTA per frame (a frame is one segment); the frame kernel runs the exact fallback itself, no
    // flags to clear, no further launches.
    // Small batches (one 300-frame stream = 30 GOPs per launch): one persistent CTA per frame would leave most SMs idle for the
    // ~57 us a frame takes, so the rows of a frame are searched by separate CTAs (each restages its own 48-row window) and the
    // exact fallback follows in one more launch.  Measured on one CIF stream: 39 us instead of 57 us per P frame.
    const bool persistent = c->me_persistent && (long long)G * L.nseg > c->me_rows_max_g;
    const bool fused = persistent && L.nseg == 1 && c->me_fused;
    if (!fused) {
        CU(cudaMemsetAsync(p.meflag, 0, sizeof(uint32_t) * G, s));
        CU(cudaMemsetAsync(p.mestate, 0, (size_t)G * g.nmb, s));
    }
    {
        LaunchScope ls(c, K_ME_SAD, s);
        const bool cif = L.pitch_w == 104 && L.seg_mbs == 22;          // CIF: compile-time pitch / segment width
        if (persistent && fused && cif) me_sad_frame_kernel<104, 22, true><<<dim3(L.nseg, G), threads, c->me_frame_smem + 16, s>>>(g, L, p, st);
        else if (persistent && fused) me_sad_frame_kernel<0, 0, true><<<dim3(L.nseg, G), threads, c->me_frame_smem + 16, s>>>(g, L, p, st);
        else if (persistent && cif) me_sad_frame_kernel<104, 22, false><<<dim3(L.nseg, G), threads, c->me_frame_smem, s>>>(g, L, p, st);
        else if (persistent) me_sad_frame_kernel<0, 0, false><<<dim3(L.nseg, G), threads, c->me_frame_smem, s>>>(g, L, p, st);
        else me_sad_kernel<<<grid, threads, c->me_smem, s>>>(g, L, p, st, 0, 1);
    }
    if (fused) return ICSP_OK;
    if (L.nseg == 1 && c->me_fused) {
        LaunchScope ls(c, K_ME_FIXUP, s);
        me_fallback_kernel<<<G, threads, c->me_smem, s>>>(g, L, p, st);
        return ICSP_OK;
    }
    { LaunchScope ls(c, K_ME_ZERO, s); me_zero_kernel<<<dim3(L.nseg, G), threads, c->me_smem, s>>>(g, L, p, st); }
    { LaunchScope ls(c, K_ME_CHAIN, s); me_chain_kernel<<<G, 32, 0, s>>>(g, p); }
    { LaunchScope ls(c, K_ME_FIXUP, s); me_sad_kernel<<<dim3(L.nseg, G), threads, c->me_smem, s>>>(g, L, p, st, 1, g.mbh); }
    return ICSP_OK;
}

// one step (intra-GOP frame index st.t) of GOPs [g0, g0+G) on stream s
int encode_step(icsp_ctx* c, const FramePtrs& p, const Step& st, int g0, int G, cudaStream_t s, cudaEvent_t signal = nullptr, int signal_stage = -1)
{
    auto stage_done = [&](int stage) { if (signal && stage == signal_stage) cudaEventRecord(signal, s); };
    const Geom& g = c->g;
    const int per = TR_THREADS / 8;
    dim3 lgrid((g.nmb + per - 1) / per, G);   // one 8-lane group per macroblock
    const bool v1 = c->tr_v1 && !p.dct_tap;   // the DCT tap exists in the second-generation kernels only
    static const char* exp_skip = getenv("ICSP_EXP_SKIP");   // timing experiments only (results are wrong): "chain", "intra", "me"
    const bool skip_chain = exp_skip && strstr(exp_skip, "chain"), skip_intra = exp_skip && strstr(exp_skip, "intra"), skip_me = exp_skip && strstr(exp_skip, "me");
    if (st.intra && skip_intra) {
    } else if (!st.intra && skip_me) {
    } else if (st.intra) {
        int hi;
        cudaStream_t h = hi_begin(c, s, hi);
        { LaunchScope ls(c, K_INTRA_ENC, h);
          unsigned char* edges = c->d_intra_edges ? c->d_intra_edges + (size_t)g0 * c->intra_edge_stride : nullptr;
          if (G <= c->intra_wide_max_g) intra_luma_kernel<0, IW_THREADS_WIDE><<<G, IW_THREADS_WIDE, c->intra_smem, h>>>(g, p, st, edges);
          else intra_luma_kernel<0><<<G, IW_THREADS, c->intra_smem, h>>>(g, p, st, edges); }
        hi_end(c, s, hi);
    } else {
        const int rc = launch_me(c, p, st, G, s);
        if (rc) return rc;
    }
    stage_done(0);
    {
        LaunchScope ls(c, st.intra ? K_FDCT_C : K_FDCT, s);
        if (v1) fdct_quant_kernel<<<lgrid, TR_THREADS, 0, s>>>(g, p, st);
        else launch_fdct2(g, p, st, lgrid, s);
    }
    stage_done(1);
    if (!skip_chain) {
        int hi;
        cudaStream_t h = hi_begin(c, s, hi);
        { LaunchScope ls(c, K_DCCHAIN, h); dc_chain_kernel<<<G, 128, c->chain_smem, h>>>(g, p, st, 0, c->chain_staged, c->dc_amb_eps); }
        hi_end(c, s, hi);
    }
    stage_done(2);
    {
        LaunchScope ls(c, st.intra ? K_IDCT_ENC_C : K_IDCT_ENC, s);
        if (v1) idct_recon_kernel<0><<<lgrid, TR_THREADS, 0, s>>>(g, p, st);
        else if (st.intra) idct_recon_kernel2<0, true><<<lgrid, TR2_THREADS, 0, s>>>(g, p, st);
        else idct_recon_kernel2<0, false><<<lgrid, TR2_THREADS, 0, s>>>(g, p, st);
    }
    stage_done(3);
    return ICSP_OK;
}

// every step of GOPs [g0, g0+G) on stream s
// Chunks of one call run the same kernel sequence on equal work, so left alone they march in lock step: all in their
// latency-bound kernels (intra wavefront, DC chains) at the same time, then all in ME, ... and nothing overlaps.  `wait`
// (recorded by the previous chunk once it is `skew` stages into its pipeline) holds this chunk back at its start, so that
// the chunks stay out of phase: one chunk's latency-bound kernels then run underneath another's throughput kernels.
int encode_chunk_plain(icsp_ctx* c, int g0, int G, int gop_len, int qdc, int qac, cudaStream_t s, cudaEvent_t wait, cudaEvent_t signal, int skew);
int encode_chunk(icsp_ctx* c, int g0, int G, int gop_len, int qdc, int qac, cudaStream_t s, cudaEvent_t wait = nullptr, cudaEvent_t signal = nullptr,
                 int skew = 0)
{
    return run_captured(c, std::make_tuple(0, g0, G, gop_len, qdc, qac, 0), s, [&] { return encode_chunk_plain(c, g0, G, gop_len, qdc, qac, s, wait, signal, skew); });
}
int encode_chunk_plain(icsp_ctx* c, int g0, int G, int gop_len, int qdc, int qac, cudaStream_t s, cudaEvent_t wait, cudaEvent_t signal, int skew)
{
    const FramePtrs p = frame_ptrs(c, g0, gop_len);
    if (wait) CU(cudaStreamWaitEvent(s, wait, 0));
    bool signalled = false;
    for (int t = 0; t < gop_len; t++) {
        const bool here = signal && skew > 0 && (skew - 1) / 4 == t;
        const int rc = encode_step(c, p, make_step(gop_len, t, qdc, qac), g0, G, s, here ? signal : nullptr, here ? (skew - 1) % 4 : -1);
        if (rc) return rc;
        signalled |= here;
    }
    if (signal && !signalled) CU(cudaEventRecord(signal, s));
    return ICSP_OK;
}

int decode_chunk_plain(icsp_ctx* c, int g0, int G, int gop_len, int qdc, int qac, cudaStream_t s);
int decode_chunk(icsp_ctx* c, int g0, int G, int gop_len, int qdc, int qac, cudaStream_t s)
{
    return run_captured(c, std::make_tuple(3, g0, G, gop_len, qdc, qac, 0), s, [&] { return decode_chunk_plain(c, g0, G, gop_len, qdc, qac, s); });
}
// one decoder step (intra-GOP frame index st.t) of GOPs [g0, g0+G) on stream s
int decode_step(icsp_ctx* c, const FramePtrs& p, const Step& st, int g0, int G, cudaStream_t s)
{
    const Geom& g = c->g;
    {
        const int per = TR_THREADS / 8;
        dim3 lgrid((g.nmb + per - 1) / per, G);   // one 8-lane group per macroblock
        if (st.intra) {
            int hi;
            cudaStream_t h = hi_begin(c, s, hi);
            { LaunchScope ls(c, K_INTRA_DEC, h);
              unsigned char* edges = c->d_intra_edges ? c->d_intra_edges + (size_t)g0 * c->intra_edge_stride : nullptr;
              if (G <= c->intra_wide_max_g) intra_luma_kernel<1, IW_THREADS_WIDE><<<G, IW_THREADS_WIDE, c->intra_smem, h>>>(g, p, st, edges);
              else intra_luma_kernel<1><<<G, IW_THREADS, c->intra_smem, h>>>(g, p, st, edges); }
            hi_end(c, s, hi);
        } else {
            LaunchScope ls(c, K_MV_RECON, s);
            const size_t mv_smem = (size_t)g.nmb * 4 <= 40 * 1024 ? (size_t)g.nmb * 4 : 0;      // vectors of a frame in shared memory (up to 10 240 macroblocks)
            mv_recon_kernel<<<G, 32, mv_smem, s>>>(g, p, st, mv_smem ? 1 : 0);
        }
        {
            int hi;
            cudaStream_t h = hi_begin(c, s, hi);
            { LaunchScope ls(c, K_DCCHAIN, h); dc_chain_kernel<<<G, 128, c->chain_smem, h>>>(g, p, st, 1, c->chain_staged, c->dc_amb_eps); }
            hi_end(c, s, hi);
        }
        {
            LaunchScope ls(c, st.intra ? K_IDCT_DEC_C : K_IDCT_DEC, s);
            if (c->tr_v1) idct_recon_kernel<1><<<lgrid, TR_THREADS, 0, s>>>(g, p, st);
            else if (st.intra) idct_recon_kernel2<1, true><<<lgrid, TR2_THREADS, 0, s>>>(g, p, st);
            else idct_recon_kernel2<1, false><<<lgrid, TR2_THREADS, 0, s>>>(g, p, st);
        }
    }
    return ICSP_OK;
}
int decode_chunk_plain(icsp_ctx* c, int g0, int G, int gop_len, int qdc, int qac, cudaStream_t s)
{
    const FramePtrs p = frame_ptrs(c, g0, gop_len);
    for (int t = 0; t < gop_len; t++) {
        const int rc = decode_step(c, p, make_step(gop_len, t, qdc, qac), g0, G, s);
        if (rc) return rc;
    }
    return ICSP_OK;
}

// ---- entropy coding ------------------------------------------------------------------------------------
constexpr int MAX_EN_CHUNKS = 64;
static inline unsigned long long en_frame_cap(const Geom& g) { return (unsigned long long)g.nmb * 800 + 32; }
int entropy_alloc(icsp_ctx* c)
{
    if (c->d_bits) return ICSP_OK;
    const size_t F = (size_t)c->cap, nmb = (size_t)c->g.nmb;
    CU(cudaMalloc(&c->d_blkbits, F * nmb * 6 * sizeof(uint32_t)));
    CU(cudaMalloc(&c->d_framebits, F * sizeof(unsigned long long)));
    CU(cudaMalloc(&c->d_streambits, F * sizeof(unsigned long long)));
    CU(cudaMalloc(&c->d_streamoff, F * sizeof(unsigned long long)));
    CU(cudaMalloc(&c->d_chunktotal, MAX_EN_CHUNKS * sizeof(unsigned long long)));
    CU(cudaMalloc(&c->d_overflow, MAX_EN_CHUNKS * sizeof(uint32_t)));
    // The reference sizes its buffer as width*height bytes per frame (ENC:4874-4875) and overruns it on noise at QP 1;
    // here the region holds the worst case: the DCT is orthonormal, so sum(level^2) <= 64*255^2 per block and
    // sum(len) <= 64*log2(mean(level^2)+256) < 1040 bits = 130 B per block, 6 blocks + 45 header bits < 800 B per MB.
    CU(cudaMalloc(&c->d_bits, F * en_frame_cap(c->g) + 64));
    CU(cudaHostAlloc(&c->h_tables, (2 * F + 2 * MAX_EN_CHUNKS) * sizeof(unsigned long long), cudaHostAllocDefault));
    return ICSP_OK;
}

// entropy-code streams [s0, s0+ns) (frames [f0, f0+nf)) of the resident batch on stream s; chunk index ci
int entropy_chunk(icsp_ctx* c, int ci, int s0, int ns, int gops_per_stream, int gop_len, cudaStream_t s)
{
    const Geom& g = c->g;
    const int fps = gops_per_stream * gop_len;
    const size_t f0 = (size_t)s0 * fps, nf = (size_t)ns * fps, nmb = (size_t)g.nmb;
    const unsigned long long per_frame = en_frame_cap(g);
    icsp_ctx::EnChunk ch{s0, ns, f0, nf, f0 * per_frame, nf * per_frame};
    if ((int)c->en_chunks.size() <= ci) c->en_chunks.resize(ci + 1);
    c->en_chunks[ci] = ch;
    FramePtrs p = frame_ptrs(c, (int)(f0 / gop_len), gop_len);
    EntropyPtrs e;
    e.blkbits = c->d_blkbits + f0 * nmb * 6; e.framebits = c->d_framebits + f0; e.streambits = c->d_streambits + s0;
    e.streamoff = c->d_streamoff + s0; e.total = c->d_chunktotal + ci; e.overflow = c->d_overflow + ci;
    e.bits = c->d_bits + ch.region_off; e.cap_bytes = ch.region_cap;
    return run_captured(c, std::make_tuple(2, s0, ns, gop_len, gops_per_stream, ci, 0), s, [&]() -> int {
        CU(cudaMemsetAsync(e.overflow, 0, sizeof(uint32_t), s));
        dim3 grid((g.nmb * 6 + EN_THREADS - 1) / EN_THREADS, (unsigned)nf);
        { LaunchScope ls(c, K_EN_SIZE, s); entropy_size_kernel<<<grid, EN_THREADS, 0, s>>>(g, p, e, gop_len); }
        { LaunchScope ls(c, K_EN_FSCAN, s); entropy_frame_scan_kernel<<<(unsigned)nf, 256, 0, s>>>(g, e); }
        {
            LaunchScope ls(c, K_EN_SSCAN, s);
            entropy_stream_scan_kernel<<<ns, 32, 0, s>>>(e, fps);
            entropy_stream_offsets_kernel<<<1, 32, 0, s>>>(e, ns);
        }
        { LaunchScope ls(c, K_EN_ZERO, s); entropy_zero_kernel<<<296, 256, 0, s>>>(e); }
        { LaunchScope ls(c, K_EN_PACK, s); entropy_pack_kernel<<<grid, EN_THREADS, 0, s>>>(g, p, e, gop_len, fps); }
        return ICSP_OK;
    });
}

// GOPs per chunk: enough chunks to overlap (>= 2 per compute stream) but each big enough to fill the GPU
int chunk_gops(const icsp_ctx* c, int n_gops, bool pipelined)
{
    if (c->chunk_gops_target > 0) return std::min(n_gops, c->chunk_gops_target);
    // Measured on B200 (profiles/README.md, chunk sweep): chunks of >= 120 GOPs keep every launch wide enough, and up to four of
    // them in flight overlap the kernels of different pipes (240 GOPs: 2 x 120 is 9 % faster than one chunk, 480: 4 x 120 is 6 %
    // faster than 3 x 160, 960: 4 x 240 beats 8 x 120); between 60 and 119 GOPs two half chunks win by 5 % (both halves run the
    // row-parallel motion search side by side).  Pipelined host calls use up to 8 chunks so that the first upload / last download are short.
    const int min_chunk = 120;
    int chunks = std::min(pipelined ? 8 : std::max(4, c->n_cstreams), std::max(1, n_gops / min_chunk));
    if (n_gops >= 60 && n_gops < min_chunk) chunks = 2;
    return (n_gops + chunks - 1) / chunks;
}

cudaEvent_t chunk_event(icsp_ctx* c, size_t i)
{
    while (c->ev_chunk.size() <= i) { cudaEvent_t e = nullptr; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); c->ev_chunk.push_back(e); }
    return c->ev_chunk[i];
}

// fork the compute streams off the main stream / join them back
int fork_streams(icsp_ctx* c)
{
    CU(cudaEventRecord(c->ev_fork, c->stream));
    for (int i = 0; i < c->n_cstreams; i++) CU(cudaStreamWaitEvent(c->cstream[i], c->ev_fork, 0));
    return ICSP_OK;
}
int join_streams(icsp_ctx* c)
{
    for (int i = 0; i < c->n_cstreams; i++) {
        CU(cudaEventRecord(c->ev_join[i], c->cstream[i]));
        CU(cudaStreamWaitEvent(c->stream, c->ev_join[i], 0));
    }
    return ICSP_OK;
}

}  // namespace

// =========================================================================================================
extern "C" {

const char* icsp_version(void) { return "libicspcuda 0.1 (sm_100a)"; }

const char* icsp_last_error(const icsp_ctx* c) { return c ? c->err : g_create_error; }

int icsp_create(icsp_ctx** out, int device, int width, int height, int max_frames)
{
    icsp_ctx* c = nullptr;
    if (!out) return fail(nullptr, ICSP_ERR_PARAM, "ctx out pointer is NULL");
    *out = nullptr;
    Geom g;
    if (geom_init(g, width, height)) return fail(nullptr, ICSP_ERR_PARAM, "width/height must be positive multiples of 16");
    if (max_frames <= 0) return fail(nullptr, ICSP_ERR_PARAM, "max_frames must be positive");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, ICSP_ERR_CUDA, "no CUDA device: %s (libicspcuda has no CPU fallback)", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, ICSP_ERR_PARAM, "device %d out of range (%d devices)", device, ndev);
    c = new (std::nothrow) icsp_ctx();
    if (!c) return fail(nullptr, ICSP_ERR_NOMEM, "host allocation failed");
    c->device = device; c->g = g; c->cap = max_frames;
    auto bail = [&](int code) { snprintf(g_create_error, sizeof(g_create_error), "%s", c->err); icsp_destroy(c); return code; };
#define CUB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(c, ICSP_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); return bail(e_ == cudaErrorMemoryAllocation ? ICSP_ERR_NOMEM : ICSP_ERR_CUDA); } } while (0)
    CUB(cudaSetDevice(device));
    CUB(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    if (const char* e = getenv("ICSP_STREAMS")) { const int v = atoi(e); if (v >= 1 && v <= MAX_CSTREAMS) c->n_cstreams = v; }
    if (const char* e = getenv("ICSP_CHUNK_GOPS")) { const int v = atoi(e); if (v >= 1) c->chunk_gops_target = v; }
    for (int i = 0; i < MAX_CSTREAMS; i++) CUB(cudaStreamCreateWithFlags(&c->cstream[i], cudaStreamNonBlocking));
    {
        int least = 0, greatest = 0;
        CUB(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        for (int i = 0; i < MAX_CSTREAMS; i++) {
            CUB(cudaStreamCreateWithPriority(&c->hstream[i], cudaStreamNonBlocking, greatest));
            for (auto& e : c->ev_hi[i]) CUB(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        if (const char* e = getenv("ICSP_HI_PRIO")) c->hi_prio = atoi(e) != 0;
    }
    CUB(cudaStreamCreateWithFlags(&c->s_up, cudaStreamNonBlocking));
    CUB(cudaStreamCreateWithFlags(&c->s_down, cudaStreamNonBlocking));
    CUB(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < MAX_CSTREAMS; i++) CUB(cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
    const size_t F = (size_t)max_frames, nmb = (size_t)g.nmb;
    CUB(cudaMalloc(&c->d_cur, F * g.fb + 64));
    CUB(cudaMalloc(&c->d_rec, F * g.fb + 64));
    CUB(cudaMalloc(&c->d_levels, F * nmb * 384 * sizeof(int16_t)));
    CUB(cudaMalloc(&c->d_acflag, F * nmb * 6));
    CUB(cudaMalloc(&c->d_mpm, F * nmb * 4));
    CUB(cudaMalloc(&c->d_ipm, F * nmb * 4));
    CUB(cudaMalloc(&c->d_mvd, F * nmb * 2 * sizeof(int16_t)));
    CUB(cudaMalloc(&c->d_mv, F * nmb * 2 * sizeof(int16_t)));
    CUB(cudaMalloc(&c->d_minsad, F * nmb * sizeof(int32_t)));
    // per-GOP scratch: at most max_frames GOPs (gop_len == 1)
    CUB(cudaMalloc(&c->d_dcraw, F * nmb * 6 * sizeof(double)));
    CUB(cudaMalloc(&c->d_dcrec, F * nmb * 6 * sizeof(int32_t)));
    CUB(cudaMalloc(&c->d_mestate, F * nmb));
    CUB(cudaMalloc(&c->d_memoves, F * nmb));
    CUB(cudaMalloc(&c->d_meflag, F * sizeof(uint32_t)));
    CUB(cudaMalloc(&c->d_mezero, F * nmb * 8 * sizeof(unsigned long long)));
    CUB(cudaMemsetAsync(c->d_mpm, 0, F * nmb * 4, c->stream));
    CUB(cudaMemsetAsync(c->d_ipm, 0, F * nmb * 4, c->stream));
    CUB(cudaMemsetAsync(c->d_mvd, 0, F * nmb * 4, c->stream));
    CUB(cudaMemsetAsync(c->d_mv, 0, F * nmb * 4, c->stream));
    CUB(cudaMemsetAsync(c->d_minsad, 0, F * nmb * 4, c->stream));
    for (auto& s : c->slots) CUB(cudaEventCreate(&s));
    if (upload_tables(c) != ICSP_OK) return bail(ICSP_ERR_CUDA);
    c->me = me_layout(g);
    c->me_smem = me_smem_bytes(c->me);
    c->me_frame_smem = me_frame_smem_bytes(c->me);
    if (const char* e = getenv("ICSP_ME_SMEM_MIN")) c->me_frame_smem = std::max(c->me_frame_smem, (size_t)atoi(e));   // experiments: cap ME CTAs per SM
    if (const char* e = getenv("ICSP_DC_EPS")) c->dc_amb_eps = (float)atof(e);
    if (const char* e = getenv("ICSP_ME_PERSISTENT")) c->me_persistent = atoi(e) != 0;
    if (const char* e = getenv("ICSP_ME_ROWS_G")) c->me_rows_max_g = atoi(e);
    if (const char* e = getenv("ICSP_SLICE_COPIES")) c->slice_copies = atoi(e) != 0;
    if (const char* e = getenv("ICSP_INTRA_WIDE_G")) c->intra_wide_max_g = atoi(e);
    if (const char* e = getenv("ICSP_ME_FUSED")) c->me_fused = atoi(e) != 0;
    if (const char* e = getenv("ICSP_TR_V1")) c->tr_v1 = atoi(e) != 0;
    if (const char* e = getenv("ICSP_SKEW")) c->skew = std::max(0, atoi(e));
    if (const char* e = getenv("ICSP_GRAPHS")) c->use_graphs = atoi(e) != 0;
    c->intra_smem = intra_smem_bytes(g);
    c->chain_smem = (size_t)(6 * g.nmb + 3) * 4 + 32;  // staged: one 4-byte slot per block
    if (c->chain_smem > 100 * 1024) { c->chain_staged = 0; c->chain_smem = (size_t)6 * g.nmb * sizeof(int) + 32; }
    if (c->intra_smem > 150 * 1024) {   // HD: keep the intra wavefront's maps in global memory (one region per GOP in flight)
        c->intra_edge_stride = (c->intra_smem + 15) / 16 * 16;
        CUB(cudaMalloc(&c->d_intra_edges, c->intra_edge_stride * F));
        c->intra_smem = 0;
    }
    if (c->me_smem > 200 * 1024 || c->chain_smem > 200 * 1024) {
        fail(c, ICSP_ERR_PARAM, "frame %dx%d too large for the shared-memory staging of this build", width, height);
        return bail(ICSP_ERR_PARAM);
    }
    {   // the attribute is per function (process wide), not per context: always opt in to the device maximum
        int optin = 0;
        CUB(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        if ((size_t)optin < c->me_smem || (size_t)optin < c->intra_smem + 20 * 1024 || (size_t)optin < c->chain_smem) {
            fail(c, ICSP_ERR_PARAM, "frame %dx%d needs more shared memory than the device offers", width, height);
            return bail(ICSP_ERR_PARAM);
        }
        CUB(cudaFuncSetAttribute(me_sad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        CUB(cudaFuncSetAttribute(me_zero_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        CUB(cudaFuncSetAttribute(me_fallback_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        CUB(cudaFuncSetAttribute(me_sad_frame_kernel<0, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        CUB(cudaFuncSetAttribute(me_sad_frame_kernel<104, 22, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        CUB(cudaFuncSetAttribute(me_sad_frame_kernel<0, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        CUB(cudaFuncSetAttribute(me_sad_frame_kernel<104, 22, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        if ((size_t)optin < c->me_frame_smem + 16) c->me_persistent = false;
        CUB(cudaFuncSetAttribute(intra_luma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 20 * 1024));
        CUB(cudaFuncSetAttribute(intra_luma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 20 * 1024));
        CUB(cudaFuncSetAttribute(intra_luma_kernel<0, IW_THREADS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 40 * 1024));
        CUB(cudaFuncSetAttribute(intra_luma_kernel<1, IW_THREADS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 40 * 1024));
        CUB(cudaFuncSetAttribute(dc_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        // latency-bound kernels want as many resident CTAs as possible: ask for the largest shared-memory carveout
        CUB(cudaFuncSetAttribute(dc_chain_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CUB(cudaFuncSetAttribute(intra_luma_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CUB(cudaFuncSetAttribute(intra_luma_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    CUB(cudaStreamSynchronize(c->stream));
#undef CUB
    *out = c;
    return ICSP_OK;
}

void icsp_destroy(icsp_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto& kv : c->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    for (auto& e : c->ev_chunk) if (e) cudaEventDestroy(e);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    for (auto& e : c->ev_join) if (e) cudaEventDestroy(e);
    for (auto& st : c->cstream) if (st) cudaStreamDestroy(st);
    for (auto& st : c->hstream) if (st) cudaStreamDestroy(st);
    for (auto& ee : c->ev_hi) for (auto& e : ee) if (e) cudaEventDestroy(e);
    if (c->s_up) cudaStreamDestroy(c->s_up);
    if (c->s_down) cudaStreamDestroy(c->s_down);
    for (auto& p : c->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto& e : c->free_events) cudaEventDestroy(e);
    for (auto& s : c->slots) if (s) cudaEventDestroy(s);
    void* bufs[] = {c->d_cur, c->d_rec, c->d_levels, c->d_acflag, c->d_mpm, c->d_ipm, c->d_mvd, c->d_mv, c->d_minsad, c->d_dcraw,
                    c->d_dcrec, c->d_mestate, c->d_memoves, c->d_meflag, c->d_mezero, c->d_shim, c->d_blkbits, c->d_framebits,
                    c->d_streambits, c->d_streamoff, c->d_chunktotal, c->d_overflow, c->d_bits, c->d_intra_edges, c->d_sse, c->d_rows};
    if (c->h_tables) cudaFreeHost(c->h_tables);
    for (void* b : bufs) if (b) cudaFree(b);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

void* icsp_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void* icsp_host_alloc_upload(size_t bytes)
{   // write-combined: not snooped during the transfer (several GPUs uploading at once no longer contend in the host's
    // coherence fabric: 4-GPU e2e 333 k -> measured below in profiles/README.md); CPU reads of it are very slow
    void* p = nullptr;
    if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocWriteCombined) != cudaSuccess) return nullptr;
    return p;
}
void icsp_host_free(void* p) { if (p) cudaFreeHost(p); }

int icsp_sync(icsp_ctx* c)
{
    if (!c) return ICSP_ERR_PARAM;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    if (c->async_err) { c->async_err = false; return fail(c, ICSP_ERR_CUDA, "a CUDA event / stream-join call failed while launching: %s", cudaGetErrorString(cudaGetLastError())); }
    return ICSP_OK;
}

// ---- encoder -------------------------------------------------------------------------------------------
int icsp_enc_upload(icsp_ctx* c, const uint8_t* frames, int n_frames)
{
    if (!c || !frames || n_frames <= 0) return fail(c, ICSP_ERR_PARAM, "icsp_enc_upload: bad arguments");
    if (n_frames > c->cap) return fail(c, ICSP_ERR_CAPACITY, "%d frames > capacity %d", n_frames, c->cap);
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(c->d_cur, frames, (size_t)n_frames * c->g.fb, cudaMemcpyHostToDevice, c->stream));
    return ICSP_OK;
}

int icsp_enc_run(icsp_ctx* c, int n_gops, int gop_len, int qdc, int qac)
{
    int rc = check_run(c, n_gops, gop_len, qdc, qac);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    const int cg = chunk_gops(c, n_gops, false);
    if ((rc = fork_streams(c))) return rc;
    for (int g0 = 0, i = 0; g0 < n_gops; g0 += cg, i++)
        if ((rc = encode_chunk(c, g0, std::min(cg, n_gops - g0), gop_len, qdc, qac, c->cstream[i % c->n_cstreams],
                               i > 0 && c->skew ? chunk_event(c, 200 + i - 1) : nullptr, c->skew ? chunk_event(c, 200 + i) : nullptr, c->skew))) return rc;
    if ((rc = join_streams(c))) return rc;
    CU(cudaGetLastError());
    return ICSP_OK;
}

int icsp_enc_download(icsp_ctx* c, int n, const icsp_enc_out* o)
{
    if (!c || !o || n <= 0) return fail(c, ICSP_ERR_PARAM, "icsp_enc_download: bad arguments");
    if (n > c->cap) return fail(c, ICSP_ERR_CAPACITY, "%d frames > capacity %d", n, c->cap);
    CU(cudaSetDevice(c->device));
    const size_t N = (size_t)n, nmb = (size_t)c->g.nmb;
    if (o->levels) CU(cudaMemcpyAsync(o->levels, c->d_levels, N * nmb * 384 * 2, cudaMemcpyDeviceToHost, c->stream));
    if (o->acflag) CU(cudaMemcpyAsync(o->acflag, c->d_acflag, N * nmb * 6, cudaMemcpyDeviceToHost, c->stream));
    if (o->mpm) CU(cudaMemcpyAsync(o->mpm, c->d_mpm, N * nmb * 4, cudaMemcpyDeviceToHost, c->stream));
    if (o->ipm) CU(cudaMemcpyAsync(o->ipm, c->d_ipm, N * nmb * 4, cudaMemcpyDeviceToHost, c->stream));
    if (o->mvd) CU(cudaMemcpyAsync(o->mvd, c->d_mvd, N * nmb * 4, cudaMemcpyDeviceToHost, c->stream));
    if (o->mv) CU(cudaMemcpyAsync(o->mv, c->d_mv, N * nmb * 4, cudaMemcpyDeviceToHost, c->stream));
    if (o->minsad) CU(cudaMemcpyAsync(o->minsad, c->d_minsad, N * nmb * 4, cudaMemcpyDeviceToHost, c->stream));
    if (o->recon) CU(cudaMemcpyAsync(o->recon, c->d_rec, N * c->g.fb, cudaMemcpyDeviceToHost, c->stream));
    return ICSP_OK;
}

static int icsp_encode_gops_impl(icsp_ctx* c, const uint8_t* frames, int n_gops, int gop_len, int qdc, int qac, const icsp_enc_out* out)
{
    int rc = check_run(c, n_gops, gop_len, qdc, qac);
    if (rc) return rc;
    if (!frames || !out) return fail(c, ICSP_ERR_PARAM, "icsp_encode_gops: NULL frames/out");
    CU(cudaSetDevice(c->device));
    const Geom& g = c->g;
    const size_t nmb = (size_t)g.nmb;
    const int n = n_gops * gop_len;
    if (out->dct) {   // debug tap of the forward DCT: device buffer for this call only
        if (cudaMalloc(&c->d_dct_tap, (size_t)n * nmb * 384 * sizeof(double)) != cudaSuccess) {
            c->d_dct_tap = nullptr;
            return fail(c, ICSP_ERR_NOMEM, "icsp_enc_out.dct: %zu bytes of device memory for the DCT tap", (size_t)n * nmb * 384 * sizeof(double));
        }
    }
    struct TapGuard { icsp_ctx* c; ~TapGuard() { if (c->d_dct_tap) { cudaDeviceSynchronize(); cudaFree(c->d_dct_tap); c->d_dct_tap = nullptr; } } } tap_guard{c};
    // inter frames leave mpm/ipm untouched and intra frames leave mvd/mv/minsad untouched: clear them so the SoA
    // rows of the other frame type read as zero, as documented
    CU(cudaMemsetAsync(c->d_mpm, 0, (size_t)n * nmb * 4, c->stream));
    CU(cudaMemsetAsync(c->d_ipm, 0, (size_t)n * nmb * 4, c->stream));
    CU(cudaMemsetAsync(c->d_mvd, 0, (size_t)n * nmb * 4, c->stream));
    CU(cudaMemsetAsync(c->d_mv, 0, (size_t)n * nmb * 4, c->stream));
    CU(cudaMemsetAsync(c->d_minsad, 0, (size_t)n * nmb * 4, c->stream));
    CU(cudaEventRecord(c->ev_fork, c->stream));
    CU(cudaStreamWaitEvent(c->s_up, c->ev_fork, 0));
    for (int i = 0; i < c->n_cstreams; i++) CU(cudaStreamWaitEvent(c->cstream[i], c->ev_fork, 0));
    CU(cudaStreamWaitEvent(c->s_down, c->ev_fork, 0));
    // software pipeline over GOP chunks: H2D(chunk i+1) | kernels(chunk i) | D2H(chunk i-1) on three kinds of streams
    const int cg = chunk_gops(c, n_gops, true);
    for (int g0 = 0, i = 0; g0 < n_gops; g0 += cg, i++) {
        const int G = std::min(cg, n_gops - g0);
        const size_t f0 = (size_t)g0 * gop_len, cnt = (size_t)G * gop_len;
        cudaStream_t cs = c->cstream[i % c->n_cstreams];
        CU(cudaMemcpyAsync(c->d_cur + f0 * g.fb, frames + f0 * g.fb, cnt * g.fb, cudaMemcpyHostToDevice, c->s_up));
        cudaEvent_t up = chunk_event(c, 2 * i), done = chunk_event(c, 2 * i + 1);
        CU(cudaEventRecord(up, c->s_up));
        CU(cudaStreamWaitEvent(cs, up, 0));
        if ((rc = encode_chunk(c, g0, G, gop_len, qdc, qac, cs))) return rc;
        CU(cudaEventRecord(done, cs));
        CU(cudaStreamWaitEvent(c->s_down, done, 0));
        cudaStream_t d = c->s_down;
        if (out->levels) CU(cudaMemcpyAsync(out->levels + f0 * nmb * 384, c->d_levels + f0 * nmb * 384, cnt * nmb * 384 * 2, cudaMemcpyDeviceToHost, d));
        if (out->acflag) CU(cudaMemcpyAsync(out->acflag + f0 * nmb * 6, c->d_acflag + f0 * nmb * 6, cnt * nmb * 6, cudaMemcpyDeviceToHost, d));
        if (out->mpm) CU(cudaMemcpyAsync(out->mpm + f0 * nmb * 4, c->d_mpm + f0 * nmb * 4, cnt * nmb * 4, cudaMemcpyDeviceToHost, d));
        if (out->ipm) CU(cudaMemcpyAsync(out->ipm + f0 * nmb * 4, c->d_ipm + f0 * nmb * 4, cnt * nmb * 4, cudaMemcpyDeviceToHost, d));
        if (out->mvd) CU(cudaMemcpyAsync(out->mvd + f0 * nmb * 2, c->d_mvd + f0 * nmb * 2, cnt * nmb * 4, cudaMemcpyDeviceToHost, d));
        if (out->mv) CU(cudaMemcpyAsync(out->mv + f0 * nmb * 2, c->d_mv + f0 * nmb * 2, cnt * nmb * 4, cudaMemcpyDeviceToHost, d));
        if (out->minsad) CU(cudaMemcpyAsync(out->minsad + f0 * nmb, c->d_minsad + f0 * nmb, cnt * nmb * 4, cudaMemcpyDeviceToHost, d));
        if (out->recon) CU(cudaMemcpyAsync(out->recon + f0 * g.fb, c->d_rec + f0 * g.fb, cnt * g.fb, cudaMemcpyDeviceToHost, d));
        if (out->dct) CU(cudaMemcpyAsync(out->dct + f0 * nmb * 384, c->d_dct_tap + f0 * nmb * 384, cnt * nmb * 384 * sizeof(double), cudaMemcpyDeviceToHost, d));
    }
    if ((rc = join_streams(c))) return rc;
    CU(cudaEventRecord(c->ev_join[0], c->s_down));
    CU(cudaStreamWaitEvent(c->stream, c->ev_join[0], 0));
    CU(cudaEventRecord(c->ev_join[1], c->s_up));
    CU(cudaStreamWaitEvent(c->stream, c->ev_join[1], 0));
    CU(cudaGetLastError());
    return icsp_sync(c);
}

// intraPrediction (ENC:556-643) on n independent frames
int icsp_intra_frame(icsp_ctx* c, const uint8_t* frames, int n, int qdc, int qac, const icsp_enc_out* out)
{
    return icsp_encode_gops(c, frames, n, 1, qdc, qac, out);
}

// interPrediction(cur, prev) (ENC:1986-2072) on n independent pairs: pair i = device frames 2i (reference: only its
// reconstruction is needed) and 2i+1 (current), step t = 1 of a GOP of length 2
static int icsp_inter_frame_impl(icsp_ctx* c, const uint8_t* cur, const uint8_t* prev_recon, int n, int qdc, int qac, const icsp_enc_out* out)
{
    int rc = check_run(c, n, 2, qdc, qac);
    if (rc) return rc;
    if (!cur || !prev_recon || !out) return fail(c, ICSP_ERR_PARAM, "icsp_inter_frame: NULL frames/out");
    CU(cudaSetDevice(c->device));
    const Geom& g = c->g;
    const size_t nmb = (size_t)g.nmb, fb = (size_t)g.fb;
    if (out->dct) {
        if (cudaMalloc(&c->d_dct_tap, (size_t)2 * n * nmb * 384 * sizeof(double)) != cudaSuccess) {
            c->d_dct_tap = nullptr;
            return fail(c, ICSP_ERR_NOMEM, "icsp_enc_out.dct: device memory for the DCT tap");
        }
    }
    struct TapGuard { icsp_ctx* c; ~TapGuard() { if (c->d_dct_tap) { cudaDeviceSynchronize(); cudaFree(c->d_dct_tap); c->d_dct_tap = nullptr; } } } tap_guard{c};
    cudaStream_t s = c->stream;
    CU(cudaMemcpy2DAsync(c->d_rec, 2 * fb, prev_recon, fb, fb, n, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpy2DAsync(c->d_cur + fb, 2 * fb, cur, fb, fb, n, cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(c->d_mpm, 0, (size_t)2 * n * nmb * 4, s));
    CU(cudaMemsetAsync(c->d_ipm, 0, (size_t)2 * n * nmb * 4, s));
    if ((rc = encode_step(c, frame_ptrs(c, 0, 2), make_step(2, 1, qdc, qac), 0, n, s))) return rc;
    // frame 2i+1 of every array -> row i of the caller's
    auto down = [&](void* dst, const void* src, size_t per_frame) -> int {
        if (!dst) return ICSP_OK;
        CU(cudaMemcpy2DAsync(dst, per_frame, (const char*)src + per_frame, 2 * per_frame, per_frame, n, cudaMemcpyDeviceToHost, s));
        return ICSP_OK;
    };
    if ((rc = down(out->levels, c->d_levels, nmb * 384 * 2)) || (rc = down(out->acflag, c->d_acflag, nmb * 6)) ||
        (rc = down(out->mpm, c->d_mpm, nmb * 4)) || (rc = down(out->ipm, c->d_ipm, nmb * 4)) || (rc = down(out->mvd, c->d_mvd, nmb * 4)) ||
        (rc = down(out->mv, c->d_mv, nmb * 4)) || (rc = down(out->minsad, c->d_minsad, nmb * 4)) || (rc = down(out->recon, c->d_rec, fb)) ||
        (rc = down(out->dct, c->d_dct_tap, nmb * 384 * sizeof(double))))
        return rc;
    CU(cudaGetLastError());
    return icsp_sync(c);
}

// ---- encoder + GPU entropy coding ------------------------------------------------------------------------
size_t icsp_bits_bound(int width, int height, int n_frames)
{
    if (width <= 0 || height <= 0 || n_frames <= 0) return 0;
    return (size_t)n_frames * ((size_t)(width / 16) * (height / 16) * 800 + 32) + 64;
}

size_t icsp_finish_body(uint8_t* body, uint64_t nbits)
{
    const size_t full = (size_t)(nbits / 8);
    const int tail = (int)(nbits % 8);
    body[full] = tail ? (uint8_t)(body[full] >> (8 - tail)) : (uint8_t)0;   // right-align the tail (ENC:4895 writes bits/8+1 bytes)
    return full + 1;
}

static int check_streams(icsp_ctx* c, int n_streams, int gops_per_stream, int gop_len)
{
    if (!c) return ICSP_ERR_PARAM;
    if (n_streams <= 0 || gops_per_stream <= 0 || gop_len <= 0) return fail(c, ICSP_ERR_PARAM, "bad stream geometry");
    if ((long long)n_streams * gops_per_stream * gop_len > c->cap) return fail(c, ICSP_ERR_CAPACITY, "more frames than capacity %d", c->cap);
    if ((long long)n_streams * gops_per_stream > 65535) return fail(c, ICSP_ERR_PARAM, "more than 65535 GOPs per call");
    // chunks are whole streams and the entropy / parse kernels index the frames of a chunk with grid.y
    const long long fps = (long long)gops_per_stream * gop_len;
    if (fps > 65535) return fail(c, ICSP_ERR_PARAM, "more than 65535 frames per stream per call (split the stream at a GOP boundary)");
    if ((n_streams + (65535 / fps) - 1) / (65535 / fps) > 64) return fail(c, ICSP_ERR_PARAM, "too many frames per call for 64 chunks of <= 65535 frames");
    return ICSP_OK;
}

// streams per chunk for the entropy path (chunks are whole streams)
static int chunk_streams(const icsp_ctx* c, int n_streams, int gops_per_stream, int gop_len, bool pipelined)
{
    const int cg = chunk_gops(c, n_streams * gops_per_stream, pipelined);
    const long long fps = (long long)gops_per_stream * gop_len;
    int cs = std::max(1, cg / gops_per_stream);
    while (cs > 1 && cs * fps > 65535) cs--;                 // the entropy kernels put the frame on grid.y
    while ((n_streams + cs - 1) / cs > MAX_EN_CHUNKS) cs++;
    return cs;
}

int icsp_entropy_run(icsp_ctx* c, int n_streams, int gops_per_stream, int gop_len)
{
    int rc = check_streams(c, n_streams, gops_per_stream, gop_len);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    if ((rc = entropy_alloc(c))) return rc;
    const int cs = chunk_streams(c, n_streams, gops_per_stream, gop_len, false);
    c->en_chunks.clear();
    if ((rc = fork_streams(c))) return rc;
    for (int s0 = 0, i = 0; s0 < n_streams; s0 += cs, i++)
        if ((rc = entropy_chunk(c, i, s0, std::min(cs, n_streams - s0), gops_per_stream, gop_len, c->cstream[i % c->n_cstreams]))) return rc;
    if ((rc = join_streams(c))) return rc;
    CU(cudaGetLastError());
    return ICSP_OK;
}

// copy the tables of chunk ci to the pinned staging area (async on stream s)
static int tables_to_host(icsp_ctx* c, int ci, cudaStream_t s)
{
    const auto& ch = c->en_chunks[ci];
    const size_t F = (size_t)c->cap;
    CU(cudaMemcpyAsync(c->h_tables + ch.s0, c->d_streambits + ch.s0, ch.ns * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(c->h_tables + F + ch.s0, c->d_streamoff + ch.s0, ch.ns * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(c->h_tables + 2 * F + ci, c->d_chunktotal + ci, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync((uint32_t*)(c->h_tables + 2 * F + MAX_EN_CHUNKS) + ci, c->d_overflow + ci, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    return ICSP_OK;
}

// after the tables of chunk ci are on the host: enqueue the body copy and fill the caller's tables
static int body_to_host(icsp_ctx* c, int ci, const icsp_bits_out* out, size_t& dst_off, cudaStream_t s)
{
    const auto& ch = c->en_chunks[ci];
    const size_t F = (size_t)c->cap;
    if (((uint32_t*)(c->h_tables + 2 * F + MAX_EN_CHUNKS))[ci]) return fail(c, ICSP_ERR_CAPACITY, "bitstream larger than the 800 bytes per macroblock worst case (corrupt levels?)");
    const size_t total = (size_t)c->h_tables[2 * F + ci];
    if (dst_off + total > out->cap_bytes) return fail(c, ICSP_ERR_CAPACITY, "icsp_bits_out.cap_bytes too small (%zu needed so far)", dst_off + total);
    CU(cudaMemcpyAsync(out->bits + dst_off, c->d_bits + ch.region_off, total, cudaMemcpyDeviceToHost, s));
    for (int i = 0; i < ch.ns; i++) {
        out->stream_bits[ch.s0 + i] = c->h_tables[ch.s0 + i];
        out->stream_offset[ch.s0 + i] = dst_off + c->h_tables[F + ch.s0 + i];
    }
    dst_off += total;
    return ICSP_OK;
}

int icsp_bits_download(icsp_ctx* c, int n_streams, const icsp_bits_out* out)
{
    if (!c || !out || !out->bits || !out->stream_bits || !out->stream_offset) return fail(c, ICSP_ERR_PARAM, "icsp_bits_download: bad arguments");
    if (c->en_chunks.empty()) return fail(c, ICSP_ERR_PARAM, "icsp_bits_download: nothing was entropy coded");
    CU(cudaSetDevice(c->device));
    int rc;
    for (int i = 0; i < (int)c->en_chunks.size(); i++) if ((rc = tables_to_host(c, i, c->stream))) return rc;
    CU(cudaStreamSynchronize(c->stream));
    size_t dst = 0;
    for (int i = 0; i < (int)c->en_chunks.size(); i++) if ((rc = body_to_host(c, i, out, dst, c->stream))) return rc;
    if (out->recon) {
        const auto& last = c->en_chunks.back();
        CU(cudaMemcpyAsync(out->recon, c->d_rec, (last.f0 + last.nf) * c->g.fb, cudaMemcpyDeviceToHost, c->stream));
    }
    (void)n_streams;
    return icsp_sync(c);
}

static int icsp_encode_streams_impl(icsp_ctx* c, const uint8_t* frames, int n_streams, int gops_per_stream, int gop_len, int qdc, int qac,
                                    const icsp_bits_out* out)
{
    int rc = check_streams(c, n_streams, gops_per_stream, gop_len);
    if (rc) return rc;
    if ((rc = check_run(c, n_streams * gops_per_stream, gop_len, qdc, qac))) return rc;
    if (!frames || !out || !out->bits || !out->stream_bits || !out->stream_offset) return fail(c, ICSP_ERR_PARAM, "icsp_encode_streams: bad arguments");
    CU(cudaSetDevice(c->device));
    if ((rc = entropy_alloc(c))) return rc;
    const Geom& g = c->g;
    const int fps = gops_per_stream * gop_len;
    const int cs = chunk_streams(c, n_streams, gops_per_stream, gop_len, true);
    const int nchunks = (n_streams + cs - 1) / cs;
    c->en_chunks.clear();
    // ICSP_TIMELINE=1: print, per chunk, when its upload / kernels / downloads finished (ms from the start of the call);
    // tools/e2e_timeline.py.  Shows the D2H link continuously busy from the end of the first chunk's kernels to the end.
    static const bool timeline = getenv("ICSP_TIMELINE") != nullptr;
    std::vector<cudaEvent_t> tl, tlb;
    cudaEvent_t tl0 = nullptr;
    if (timeline) { cudaEventCreate(&tl0); cudaEventRecord(tl0, c->stream); }
    CU(cudaEventRecord(c->ev_fork, c->stream));
    CU(cudaStreamWaitEvent(c->s_up, c->ev_fork, 0));
    for (int i = 0; i < c->n_cstreams; i++) CU(cudaStreamWaitEvent(c->cstream[i], c->ev_fork, 0));
    CU(cudaStreamWaitEvent(c->s_down, c->ev_fork, 0));
    const bool sliced = nchunks == 1 && gop_len > 1 && gop_len <= 64 && c->slice_copies;   // one graph and two events per step
    // pass 1: enqueue upload | kernels + entropy | table + recon download for every chunk
    for (int s0 = 0, i = 0; s0 < n_streams; s0 += cs, i++) {
        const int ns = std::min(cs, n_streams - s0);
        const size_t f0 = (size_t)s0 * fps, cnt = (size_t)ns * fps;
        cudaStream_t st = c->cstream[i % c->n_cstreams];
        cudaEvent_t up = chunk_event(c, 3 * i), done = chunk_event(c, 3 * i + 1), tbl = chunk_event(c, 3 * i + 2);
        if (sliced) {
            // A call that is ONE chunk (a few streams) has nothing to overlap its copies with chunk-wise: upload, 1 ms of kernels
            // and download would run one after the other.  Its steps are sequential anyway, so the copies are sliced by step
            // instead: frame t of every GOP is one strided 2-D copy (row = one frame, pitch = one GOP), step t waits for
            // slice t only, and the reconstruction of step t goes back while step t+1 runs.  One CUDA graph per step.
            const int G = ns * gops_per_stream, g0 = s0 * gops_per_stream;
            const size_t pitch = (size_t)gop_len * g.fb;
            const FramePtrs p = frame_ptrs(c, g0, gop_len);
            for (int t = 0; t < gop_len; t++) {
                cudaEvent_t upt = chunk_event(c, 3 * MAX_EN_CHUNKS + 2 * t), dnt = chunk_event(c, 3 * MAX_EN_CHUNKS + 2 * t + 1);
                CU(cudaMemcpy2DAsync(c->d_cur + (f0 + t) * g.fb, pitch, frames + (f0 + t) * g.fb, pitch, g.fb, (size_t)G, cudaMemcpyHostToDevice, c->s_up));
                CU(cudaEventRecord(upt, c->s_up));
                CU(cudaStreamWaitEvent(st, upt, 0));
                const Step stp = make_step(gop_len, t, qdc, qac);
                if ((rc = run_captured(c, std::make_tuple(4, g0, G, gop_len, qdc, qac, t), st, [&] { return encode_step(c, p, stp, g0, G, st); }))) return rc;
                if (out->recon) {
                    CU(cudaEventRecord(dnt, st));
                    CU(cudaStreamWaitEvent(c->s_down, dnt, 0));
                    CU(cudaMemcpy2DAsync(out->recon + (f0 + t) * g.fb, pitch, c->d_rec + (f0 + t) * g.fb, pitch, g.fb, (size_t)G, cudaMemcpyDeviceToHost, c->s_down));
                }
            }
            CU(cudaEventRecord(up, c->s_up));
        } else {
            CU(cudaMemcpyAsync(c->d_cur + f0 * g.fb, frames + f0 * g.fb, cnt * g.fb, cudaMemcpyHostToDevice, c->s_up));
            CU(cudaEventRecord(up, c->s_up));
            CU(cudaStreamWaitEvent(st, up, 0));
            if ((rc = encode_chunk(c, s0 * gops_per_stream, ns * gops_per_stream, gop_len, qdc, qac, st))) return rc;
        }
        if ((rc = entropy_chunk(c, i, s0, ns, gops_per_stream, gop_len, st))) return rc;
        if ((rc = tables_to_host(c, i, st))) return rc;
        CU(cudaEventRecord(done, st));
        CU(cudaEventRecord(tbl, st));
        CU(cudaStreamWaitEvent(c->s_down, done, 0));
        if (out->recon && !sliced) CU(cudaMemcpyAsync(out->recon + f0 * g.fb, c->d_rec + f0 * g.fb, cnt * g.fb, cudaMemcpyDeviceToHost, c->s_down));
        if (timeline) {
            cudaEvent_t e[3];
            for (auto& x : e) cudaEventCreate(&x);
            cudaEventRecord(e[0], c->s_up); cudaEventRecord(e[1], st); cudaEventRecord(e[2], c->s_down);
            tl.insert(tl.end(), e, e + 3);
        }
    }
    // pass 2: as each chunk's tables arrive, enqueue its body copy (overlaps the kernels of later chunks)
    size_t dst = 0;
    for (int i = 0; i < nchunks; i++) {
        CU(cudaEventSynchronize(chunk_event(c, 3 * i + 2)));
        if ((rc = body_to_host(c, i, out, dst, c->stream))) { cudaDeviceSynchronize(); return rc; }   // main stream is idle: bodies do not queue behind recon copies
        if (timeline) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, c->stream); tlb.push_back(e); }
    }
    if ((rc = join_streams(c))) return rc;
    CU(cudaEventRecord(c->ev_join[0], c->s_down));
    CU(cudaStreamWaitEvent(c->stream, c->ev_join[0], 0));
    CU(cudaEventRecord(c->ev_join[1], c->s_up));
    CU(cudaStreamWaitEvent(c->stream, c->ev_join[1], 0));
    CU(cudaGetLastError());
    rc = icsp_sync(c);
    if (timeline) {
        cudaEvent_t end; cudaEventCreate(&end); cudaEventRecord(end, c->stream); cudaEventSynchronize(end);
        float t;
        for (int i = 0; i < nchunks; i++) {
            float a, b, d, e;
            cudaEventElapsedTime(&a, tl0, tl[3 * i]); cudaEventElapsedTime(&b, tl0, tl[3 * i + 1]); cudaEventElapsedTime(&d, tl0, tl[3 * i + 2]);
            cudaEventElapsedTime(&e, tl0, tlb[i]);
            fprintf(stderr, "[icsp timeline] chunk %d: upload done %.1f ms, kernels done %.1f ms, recon download done %.1f ms, body download done %.1f ms\n", i, a, b, d, e);
        }
        cudaEventElapsedTime(&t, tl0, end);
        fprintf(stderr, "[icsp timeline] call done %.1f ms\n", t);
        for (auto e : tl) cudaEventDestroy(e);
        for (auto e : tlb) cudaEventDestroy(e);
        cudaEventDestroy(tl0); cudaEventDestroy(end);
    }
    return rc;
}

// macroblock-row index of what icsp_encode_streams / icsp_entropy_run just coded (SURVEY.md §8 f3 side-car)
int icsp_bits_row_index(icsp_ctx* c, int n_frames, uint64_t* rows)
{
    if (!c || !rows || n_frames <= 0) return fail(c, ICSP_ERR_PARAM, "icsp_bits_row_index: bad arguments");
    if (c->en_chunks.empty()) return fail(c, ICSP_ERR_PARAM, "icsp_bits_row_index: nothing was entropy coded");
    const auto& last = c->en_chunks.back();
    if ((size_t)n_frames > last.f0 + last.nf) return fail(c, ICSP_ERR_PARAM, "icsp_bits_row_index: only %zu frames were entropy coded", last.f0 + last.nf);
    CU(cudaSetDevice(c->device));
    const Geom& g = c->g;
    if (!c->d_rows) CU(cudaMalloc(&c->d_rows, (size_t)c->cap * g.mbh * sizeof(unsigned long long)));
    EntropyPtrs e{};
    e.blkbits = c->d_blkbits; e.framebits = c->d_framebits;
    const int n = n_frames * g.mbh;
    { LaunchScope ls(c, K_ROWIDX, c->stream); row_index_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(g, e, c->d_rows, n_frames); }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(rows, c->d_rows, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    return icsp_sync(c);
}

// ---- decoder -------------------------------------------------------------------------------------------
int icsp_dec_upload(icsp_ctx* c, const icsp_dec_in* in, int n)
{
    if (!c || !in || !in->levels || !in->mpm || !in->ipm || !in->mvd || n <= 0)
        return fail(c, ICSP_ERR_PARAM, "icsp_dec_upload: bad arguments");
    if (n > c->cap) return fail(c, ICSP_ERR_CAPACITY, "%d frames > capacity %d", n, c->cap);
    CU(cudaSetDevice(c->device));
    const size_t N = (size_t)n, nmb = (size_t)c->g.nmb;
    CU(cudaMemcpyAsync(c->d_levels, in->levels, N * nmb * 384 * 2, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_mpm, in->mpm, N * nmb * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_ipm, in->ipm, N * nmb * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_mvd, in->mvd, N * nmb * 4, cudaMemcpyHostToDevice, c->stream));
    return ICSP_OK;
}

int icsp_dec_run(icsp_ctx* c, int n_gops, int gop_len, int qdc, int qac)
{
    int rc = check_run(c, n_gops, gop_len, qdc, qac);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    const int cg = chunk_gops(c, n_gops, false);
    if ((rc = fork_streams(c))) return rc;
    for (int g0 = 0, i = 0; g0 < n_gops; g0 += cg, i++)
        if ((rc = decode_chunk(c, g0, std::min(cg, n_gops - g0), gop_len, qdc, qac, c->cstream[i % c->n_cstreams]))) return rc;
    if ((rc = join_streams(c))) return rc;
    CU(cudaGetLastError());
    return ICSP_OK;
}

int icsp_dec_download(icsp_ctx* c, int n, uint8_t* out)
{
    if (!c || !out || n <= 0) return fail(c, ICSP_ERR_PARAM, "icsp_dec_download: bad arguments");
    if (n > c->cap) return fail(c, ICSP_ERR_CAPACITY, "%d frames > capacity %d", n, c->cap);
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(out, c->d_rec, (size_t)n * c->g.fb, cudaMemcpyDeviceToHost, c->stream));
    return ICSP_OK;
}

static int icsp_decode_gops_impl(icsp_ctx* c, const icsp_dec_in* in, int n_gops, int gop_len, int qdc, int qac, uint8_t* out)
{
    int rc = check_run(c, n_gops, gop_len, qdc, qac);
    if (rc) return rc;
    if (!in || !in->levels || !in->mpm || !in->ipm || !in->mvd || !out) return fail(c, ICSP_ERR_PARAM, "icsp_decode_gops: bad arguments");
    CU(cudaSetDevice(c->device));
    const Geom& g = c->g;
    const size_t nmb = (size_t)g.nmb;
    CU(cudaEventRecord(c->ev_fork, c->stream));
    CU(cudaStreamWaitEvent(c->s_up, c->ev_fork, 0));
    for (int i = 0; i < c->n_cstreams; i++) CU(cudaStreamWaitEvent(c->cstream[i], c->ev_fork, 0));
    CU(cudaStreamWaitEvent(c->s_down, c->ev_fork, 0));
    // software pipeline over GOP chunks: H2D of the parsed syntax | reconstruction kernels | D2H of the decoded frames
    const int cg = chunk_gops(c, n_gops, true);
    for (int g0 = 0, i = 0; g0 < n_gops; g0 += cg, i++) {
        const int G = std::min(cg, n_gops - g0);
        const size_t f0 = (size_t)g0 * gop_len, cnt = (size_t)G * gop_len;
        cudaStream_t cs = c->cstream[i % c->n_cstreams];
        CU(cudaMemcpyAsync(c->d_levels + f0 * nmb * 384, in->levels + f0 * nmb * 384, cnt * nmb * 384 * 2, cudaMemcpyHostToDevice, c->s_up));
        CU(cudaMemcpyAsync(c->d_mpm + f0 * nmb * 4, in->mpm + f0 * nmb * 4, cnt * nmb * 4, cudaMemcpyHostToDevice, c->s_up));
        CU(cudaMemcpyAsync(c->d_ipm + f0 * nmb * 4, in->ipm + f0 * nmb * 4, cnt * nmb * 4, cudaMemcpyHostToDevice, c->s_up));
        CU(cudaMemcpyAsync(c->d_mvd + f0 * nmb * 2, in->mvd + f0 * nmb * 2, cnt * nmb * 4, cudaMemcpyHostToDevice, c->s_up));
        cudaEvent_t up = chunk_event(c, 3 * i), done = chunk_event(c, 3 * i + 1);
        CU(cudaEventRecord(up, c->s_up));
        CU(cudaStreamWaitEvent(cs, up, 0));
        if ((rc = decode_chunk(c, g0, G, gop_len, qdc, qac, cs))) return rc;
        CU(cudaEventRecord(done, cs));
        CU(cudaStreamWaitEvent(c->s_down, done, 0));
        CU(cudaMemcpyAsync(out + f0 * g.fb, c->d_rec + f0 * g.fb, cnt * g.fb, cudaMemcpyDeviceToHost, c->s_down));
    }
    if ((rc = join_streams(c))) return rc;
    CU(cudaEventRecord(c->ev_join[0], c->s_down));
    CU(cudaStreamWaitEvent(c->stream, c->ev_join[0], 0));
    CU(cudaEventRecord(c->ev_join[1], c->s_up));
    CU(cudaStreamWaitEvent(c->stream, c->ev_join[1], 0));
    CU(cudaGetLastError());
    return icsp_sync(c);
}

// ---- shims -----------------------------------------------------------------------------------------------
static int shim_reserve(icsp_ctx* c, size_t bytes)
{
    if (c->shim_bytes >= bytes) return ICSP_OK;
    if (c->d_shim) cudaFree(c->d_shim);
    c->d_shim = nullptr; c->shim_bytes = 0;
    if (cudaMalloc(&c->d_shim, bytes) != cudaSuccess) return fail(c, ICSP_ERR_NOMEM, "shim buffer allocation failed");
    c->shim_bytes = bytes;
    return ICSP_OK;
}

int icsp_me_sad(icsp_ctx* c, const uint8_t* cur_y, const uint8_t* ref_y, int n, int16_t* mv, int32_t* minsad)
{
    if (!c || !cur_y || !ref_y || n <= 0 || !mv) return fail(c, ICSP_ERR_PARAM, "icsp_me_sad: bad arguments");
    if (2 * n > c->cap) return fail(c, ICSP_ERR_CAPACITY, "icsp_me_sad needs capacity >= 2*n frames");
    CU(cudaSetDevice(c->device));
    const Geom& g = c->g;
    const size_t ysz = (size_t)g.w * g.h;
    // pair i: reference luma -> rec frame 2i, current luma -> cur frame 2i+1 (gop_len 2, step 1)
    CU(cudaMemcpy2DAsync(c->d_rec, (size_t)2 * g.fb, ref_y, ysz, ysz, n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpy2DAsync(c->d_cur + g.fb, (size_t)2 * g.fb, cur_y, ysz, ysz, n, cudaMemcpyHostToDevice, c->stream));
    const Step st = make_step(2, 1, 1, 1);
    int rc = launch_me(c, frame_ptrs(c), st, n, c->stream);
    if (rc) return rc;
    CU(cudaMemcpy2DAsync(mv, (size_t)g.nmb * 4, c->d_mv + (size_t)g.nmb * 2, (size_t)g.nmb * 8, (size_t)g.nmb * 4, n,
                         cudaMemcpyDeviceToHost, c->stream));
    if (minsad)
        CU(cudaMemcpy2DAsync(minsad, (size_t)g.nmb * 4, c->d_minsad + g.nmb, (size_t)g.nmb * 8, (size_t)g.nmb * 4, n,
                             cudaMemcpyDeviceToHost, c->stream));
    CU(cudaGetLastError());
    return icsp_sync(c);
}

int icsp_dct8x8(icsp_ctx* c, const int32_t* blocks, int n, double* out)
{
    if (!c || !blocks || !out || n <= 0) return fail(c, ICSP_ERR_PARAM, "icsp_dct8x8: bad arguments");
    CU(cudaSetDevice(c->device));
    int rc = shim_reserve(c, (size_t)n * 64 * 12);
    if (rc) return rc;
    int32_t* din = (int32_t*)c->d_shim;
    double* dout = (double*)((char*)c->d_shim + (size_t)n * 64 * 4);
    // keep the double region 8-byte aligned
    dout = (double*)(((uintptr_t)dout + 7) & ~(uintptr_t)7);
    CU(cudaMemcpyAsync(din, blocks, (size_t)n * 64 * 4, cudaMemcpyHostToDevice, c->stream));
    { LaunchScope ls(c, K_DCT_SHIM); dct8x8_kernel<<<(n + 15) / 16, 128, 0, c->stream>>>(din, dout, n); }
    CU(cudaMemcpyAsync(out, dout, (size_t)n * 64 * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaGetLastError());
    return icsp_sync(c);
}

int icsp_idct8x8(icsp_ctx* c, const int32_t* blocks, int n, int table, double* out)
{
    if (!c || !blocks || !out || n <= 0 || (table != 0 && table != 1)) return fail(c, ICSP_ERR_PARAM, "icsp_idct8x8: bad arguments");
    CU(cudaSetDevice(c->device));
    int rc = shim_reserve(c, (size_t)n * 64 * 12 + 16);
    if (rc) return rc;
    int32_t* din = (int32_t*)c->d_shim;
    double* dout = (double*)(((uintptr_t)((char*)c->d_shim + (size_t)n * 64 * 4) + 7) & ~(uintptr_t)7);
    CU(cudaMemcpyAsync(din, blocks, (size_t)n * 64 * 4, cudaMemcpyHostToDevice, c->stream));
    {
        LaunchScope ls(c, K_IDCT_SHIM);
        if (table == 0) idct8x8_kernel<0><<<(n + 15) / 16, 128, 0, c->stream>>>(din, dout, n);
        else idct8x8_kernel<1><<<(n + 15) / 16, 128, 0, c->stream>>>(din, dout, n);
    }
    CU(cudaMemcpyAsync(out, dout, (size_t)n * 64 * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaGetLastError());
    return icsp_sync(c);
}

int icsp_quant(icsp_ctx* c, const double* dct, int n, int qdc, int qac, int chroma, int32_t* levels, uint8_t* acflag)
{
    if (!c || !dct || !levels || n <= 0 || qdc <= 0 || qac <= 0) return fail(c, ICSP_ERR_PARAM, "icsp_quant: bad arguments");
    CU(cudaSetDevice(c->device));
    int rc = shim_reserve(c, (size_t)n * (64 * 12 + 1) + 64);
    if (rc) return rc;
    double* din = (double*)c->d_shim;
    int32_t* dl = (int32_t*)(din + (size_t)n * 64);
    uint8_t* da = (uint8_t*)(dl + (size_t)n * 64);
    CU(cudaMemcpyAsync(din, dct, (size_t)n * 64 * 8, cudaMemcpyHostToDevice, c->stream));
    {
        LaunchScope ls(c, K_QUANT_SHIM);
        quant8x8_kernel<<<(n + 3) / 4, 256, 0, c->stream>>>(din, dl, da, n, (unsigned)((0x80000000ull + qdc - 1) / qdc),
                                                             (unsigned)((0x80000000ull + qac - 1) / qac), chroma);
    }
    CU(cudaMemcpyAsync(levels, dl, (size_t)n * 64 * 4, cudaMemcpyDeviceToHost, c->stream));
    if (acflag) CU(cudaMemcpyAsync(acflag, da, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaGetLastError());
    return icsp_sync(c);
}

// ---- instrumentation -------------------------------------------------------------------------------------
int icsp_set_profiling(icsp_ctx* c, int enabled)
{
    if (!c) return ICSP_ERR_PARAM;
    cudaSetDevice(c->device);
    fold_pending(c);
    c->profiling = enabled != 0;
    return ICSP_OK;
}
int icsp_reset_stats(icsp_ctx* c)
{
    if (!c) return ICSP_ERR_PARAM;
    cudaSetDevice(c->device);
    fold_pending(c);
    for (int k = 0; k < K_COUNT; k++) { c->total_ms[k] = 0; c->count[k] = 0; }
    c->launches = 0;
    return ICSP_OK;
}
int icsp_get_stats(icsp_ctx* c, icsp_kernel_stat* stats, int cap)
{
    if (!c || !stats || cap <= 0) return ICSP_ERR_PARAM;
    cudaSetDevice(c->device);
    fold_pending(c);
    int n = 0;
    for (int k = 0; k < K_COUNT && n < cap; k++) {
        if (!c->count[k]) continue;
        snprintf(stats[n].name, sizeof(stats[n].name), "%s", kKernelNames[k]);
        stats[n].launches = c->count[k];
        stats[n].total_ms = c->total_ms[k];
        n++;
    }
    return n;
}
uint64_t icsp_launch_count(const icsp_ctx* c) { return c ? c->launches : 0; }

int icsp_configure(icsp_ctx* c, int n_compute_streams, int chunk_gops_)
{
    if (!c || n_compute_streams < 1 || n_compute_streams > MAX_CSTREAMS || chunk_gops_ < 0) return fail(c, ICSP_ERR_PARAM, "icsp_configure: bad arguments");
    c->n_cstreams = n_compute_streams;
    c->chunk_gops_target = chunk_gops_;
    return ICSP_OK;
}

// ---- decoder with the bit reader on the GPU ------------------------------------------------
static int icsp_decode_streams_impl(icsp_ctx* c, const icsp_dec_bits_in* in, int n_streams, int gops_per_stream, int gop_len, int qdc, int qac,
                                    uint8_t* out)
{
    int rc = check_streams(c, n_streams, gops_per_stream, gop_len);
    if (rc) return rc;
    if ((rc = check_run(c, n_streams * gops_per_stream, gop_len, qdc, qac))) return rc;
    if (!in || !in->bits || !in->stream_offset || !in->stream_bytes || !in->row_bit_offset || !out)
        return fail(c, ICSP_ERR_PARAM, "icsp_decode_streams: bad arguments");
    CU(cudaSetDevice(c->device));
    if ((rc = entropy_alloc(c))) return rc;
    const Geom& g = c->g;
    const size_t nmb = (size_t)g.nmb;
    if (!c->d_rows) CU(cudaMalloc(&c->d_rows, (size_t)c->cap * g.mbh * sizeof(unsigned long long)));
    const int fps = gops_per_stream * gop_len;
    const size_t cap_bytes = (size_t)c->cap * en_frame_cap(g);
    uint64_t prev_end = 0;
    for (int s = 0; s < n_streams; s++) {      // bodies in stream order, 4-byte aligned starts, inside the device region
        const uint64_t o = in->stream_offset[s], b = in->stream_bytes[s];
        // written so that no sum can wrap: o <= cap, b <= cap - o; the previous stream's end was validated the same way
        bool bad = (o & 3) || o > cap_bytes || b > cap_bytes - o;
        if (!bad && s) bad = o < prev_end;
        if (bad) return fail(c, ICSP_ERR_PARAM, "icsp_decode_streams: stream %d: offsets must ascend, be multiples of 4 and fit the context", s);
        prev_end = o + b;
        for (int r = 0; r < fps * g.mbh; r++)
            if (in->row_bit_offset[(size_t)s * fps * g.mbh + r] > b * 8) return fail(c, ICSP_ERR_PARAM, "icsp_decode_streams: stream %d: row index beyond the body", s);
    }
    // the (validated) tables go up first, on the main stream
    unsigned long long* h_off = c->h_tables;                       // pinned staging: [cap] offsets, [cap] lengths
    unsigned long long* h_len = c->h_tables + c->cap;
    for (int s = 0; s < n_streams; s++) { h_off[s] = in->stream_offset[s]; h_len[s] = in->stream_bytes[s]; }
    CU(cudaMemcpyAsync(c->d_streamoff, h_off, n_streams * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_streambits, h_len, n_streams * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
    CU(cudaEventRecord(c->ev_fork, c->stream));
    CU(cudaStreamWaitEvent(c->s_up, c->ev_fork, 0));
    for (int i = 0; i < c->n_cstreams; i++) CU(cudaStreamWaitEvent(c->cstream[i], c->ev_fork, 0));
    CU(cudaStreamWaitEvent(c->s_down, c->ev_fork, 0));
    // software pipeline over chunks of whole streams: H2D of bodies + row index | parse + reconstruction | D2H of the frames
    const int cs = chunk_streams(c, n_streams, gops_per_stream, gop_len, true);
    for (int s0 = 0, i = 0; s0 < n_streams; s0 += cs, i++) {
        const int ns = std::min(cs, n_streams - s0);
        const size_t f0 = (size_t)s0 * fps, cnt = (size_t)ns * fps;
        cudaStream_t st = c->cstream[i % c->n_cstreams];
        const uint64_t b0 = in->stream_offset[s0], b1 = in->stream_offset[s0 + ns - 1] + in->stream_bytes[s0 + ns - 1];
        CU(cudaMemcpyAsync(c->d_bits + b0, in->bits + b0, b1 - b0, cudaMemcpyHostToDevice, c->s_up));
        CU(cudaMemcpyAsync(c->d_rows + f0 * g.mbh, in->row_bit_offset + f0 * g.mbh, cnt * g.mbh * sizeof(uint64_t), cudaMemcpyHostToDevice, c->s_up));
        cudaEvent_t up = chunk_event(c, 2 * i), done = chunk_event(c, 2 * i + 1);
        CU(cudaEventRecord(up, c->s_up));
        CU(cudaMemsetAsync(c->d_levels + f0 * nmb * 384, 0, cnt * nmb * 384 * sizeof(int16_t), st));   // the parser stores non-zero levels only
        CU(cudaStreamWaitEvent(st, up, 0));
        {
            FramePtrs p = frame_ptrs(c, (int)(f0 / gop_len), gop_len);
            ParsePtrs q{c->d_bits, c->d_streamoff, c->d_streambits, c->d_rows + f0 * g.mbh};
            const int n = (int)cnt * g.mbh;
            LaunchScope ls(c, K_PARSE, st);
            parse_rows_kernel<<<(n + 63) / 64, 64, 0, st>>>(g, p, q, gop_len, fps, s0, (int)cnt);
        }
        if (n_streams <= cs && gop_len > 1 && gop_len <= 64 && c->slice_copies) {
            // one chunk: frame t of every GOP goes back while step t+1 runs (see icsp_encode_streams)
            const int G = ns * gops_per_stream, g0 = s0 * gops_per_stream;
            const size_t pitch = (size_t)gop_len * g.fb;
            const FramePtrs p = frame_ptrs(c, g0, gop_len);
            for (int t = 0; t < gop_len; t++) {
                const Step stp = make_step(gop_len, t, qdc, qac);
                if ((rc = run_captured(c, std::make_tuple(5, g0, G, gop_len, qdc, qac, t), st, [&] { return decode_step(c, p, stp, g0, G, st); }))) return rc;
                cudaEvent_t dnt = chunk_event(c, 2 * MAX_EN_CHUNKS + t);
                CU(cudaEventRecord(dnt, st));
                CU(cudaStreamWaitEvent(c->s_down, dnt, 0));
                CU(cudaMemcpy2DAsync(out + (f0 + t) * g.fb, pitch, c->d_rec + (f0 + t) * g.fb, pitch, g.fb, (size_t)G, cudaMemcpyDeviceToHost, c->s_down));
            }
            continue;
        }
        if ((rc = decode_chunk(c, s0 * gops_per_stream, ns * gops_per_stream, gop_len, qdc, qac, st))) return rc;
        CU(cudaEventRecord(done, st));
        CU(cudaStreamWaitEvent(c->s_down, done, 0));
        CU(cudaMemcpyAsync(out + f0 * g.fb, c->d_rec + f0 * g.fb, cnt * g.fb, cudaMemcpyDeviceToHost, c->s_down));
    }
    if ((rc = join_streams(c))) return rc;
    CU(cudaEventRecord(c->ev_join[0], c->s_down));
    CU(cudaStreamWaitEvent(c->stream, c->ev_join[0], 0));
    CU(cudaEventRecord(c->ev_join[1], c->s_up));
    CU(cudaStreamWaitEvent(c->stream, c->ev_join[1], 0));
    CU(cudaGetLastError());
    return icsp_sync(c);
}

// ---- quality numbers without the reconstruction read-back ----------------------------------
int icsp_enc_sse(icsp_ctx* c, int n, uint64_t* sse)
{
    if (!c || !sse || n <= 0) return fail(c, ICSP_ERR_PARAM, "icsp_enc_sse: bad arguments");
    if (n > c->cap) return fail(c, ICSP_ERR_CAPACITY, "%d frames > capacity %d", n, c->cap);
    CU(cudaSetDevice(c->device));
    if (!c->d_sse) CU(cudaMalloc(&c->d_sse, (size_t)c->cap * 3 * sizeof(unsigned long long)));
    { LaunchScope ls(c, K_SSE, c->stream); plane_sse_kernel<<<dim3(3, n), 256, 0, c->stream>>>(c->g, c->d_cur, c->d_rec, c->d_sse); }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(sse, c->d_sse, (size_t)n * 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    return icsp_sync(c);
}

int icsp_event_record(icsp_ctx* c, int slot)
{
    if (!c || slot < 0 || slot >= 8) return ICSP_ERR_PARAM;
    CU(cudaSetDevice(c->device));
    CU(cudaEventRecord(c->slots[slot], c->stream));
    return ICSP_OK;
}
int icsp_event_elapsed_ms(icsp_ctx* c, int a, int b, float* ms)
{
    if (!c || !ms || a < 0 || a >= 8 || b < 0 || b >= 8) return ICSP_ERR_PARAM;
    CU(cudaSetDevice(c->device));
    CU(cudaEventSynchronize(c->slots[b]));
    CU(cudaEventElapsedTime(ms, c->slots[a], c->slots[b]));
    return ICSP_OK;
}

// The pipelined one-shot calls enqueue asynchronous copies into CALLER buffers on several streams.  Whatever makes one of
// them fail half way (a CUDA error, a capacity check), nothing may still be in flight when the error code is returned: the
// caller is free to release or reuse its buffers at once.
static int settle(icsp_ctx* c, int rc)
{
    if (rc != ICSP_OK && c) { cudaSetDevice(c->device); cudaDeviceSynchronize(); }
    return rc;
}
int icsp_inter_frame(icsp_ctx* c, const uint8_t* cur, const uint8_t* prev_recon, int n, int qdc, int qac, const icsp_enc_out* out)
{
    return settle(c, icsp_inter_frame_impl(c, cur, prev_recon, n, qdc, qac, out));
}
int icsp_encode_gops(icsp_ctx* c, const uint8_t* frames, int n_gops, int gop_len, int qdc, int qac, const icsp_enc_out* out)
{
    return settle(c, icsp_encode_gops_impl(c, frames, n_gops, gop_len, qdc, qac, out));
}
int icsp_decode_gops(icsp_ctx* c, const icsp_dec_in* in, int n_gops, int gop_len, int qdc, int qac, uint8_t* out)
{
    return settle(c, icsp_decode_gops_impl(c, in, n_gops, gop_len, qdc, qac, out));
}
int icsp_encode_streams(icsp_ctx* c, const uint8_t* frames, int n_streams, int gops_per_stream, int gop_len, int qdc, int qac,
                        const icsp_bits_out* out)
{
    return settle(c, icsp_encode_streams_impl(c, frames, n_streams, gops_per_stream, gop_len, qdc, qac, out));
}
int icsp_decode_streams(icsp_ctx* c, const icsp_dec_bits_in* in, int n_streams, int gops_per_stream, int gop_len, int qdc, int qac,
                        uint8_t* out)
{
    return settle(c, icsp_decode_streams_impl(c, in, n_streams, gops_per_stream, gop_len, qdc, qac, out));
}

}  // extern "C"
