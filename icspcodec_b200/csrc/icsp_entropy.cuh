// f1 of SURVEY.md §8: entropy coding + bit packing on the GPU.  Bit-exact restatement of what the reference's
// intraBody / interBody (ENC:5032-5236) and DCentropy / ACentropy / MVentropy (ENC:5417-6334) emit, produced in
// parallel: (1) one thread per 8x8 block computes its bit length, (2) per-frame and per-stream exclusive scans turn
// lengths into absolute bit positions inside each stream's body, (3) one thread per block re-encodes its symbols
// with a 64-bit accumulator and writes whole 32-bit words; only the first and last word of a block can be shared
// with a neighbour and are merged with atomicOr into the pre-zeroed output.
// Output per stream: the body as one MSB-first bit string (byte-aligned start, 16-byte aligned offset).  The host
// adds the 14-byte header and applies the reference's tail rule (last byte right-aligned, ENC:4895).
#pragma once
#include "icsp_kernels.cuh"

namespace icsp {

struct EntropyPtrs {
    uint32_t* blkbits;                  // [F][nmb*6]  bit length, then (after the scan) bit offset inside the frame
    unsigned long long* framebits;      // [F]         frame bit length, then bit offset inside its stream
    unsigned long long* streambits;     // [S]         stream body length in bits
    unsigned long long* streamoff;      // [S]         byte offset of the stream inside the chunk region
    unsigned long long* total;          // [1]         bytes used in the chunk region
    uint32_t* overflow;                 // [1]
    uint8_t* bits;                      // chunk region
    unsigned long long cap_bytes;       // capacity of the chunk region
};

// category VLC (DCentropy ENC:5417-5602 and twins): a = |v|, s = (v >= 0).  a == 0: "00"; e = floor(log2 a) capped at
// 11; e <= 4: (e+2) in 3 bits, s, e bits of a-2^e; e >= 5: (e-2) ones, 0, s, e low bits of a-2^e.
__device__ __forceinline__ int vlc_len(int v)
{
    const unsigned a = (unsigned)abs(v);
    if (a == 0) return 2;
    const int e = min(31 - __clz(a), 11);
    return e <= 4 ? 4 + e : 2 * e;
}
__device__ __forceinline__ uint32_t vlc_code(int v, int& len)
{
    const unsigned a = (unsigned)abs(v), s = v >= 0 ? 1u : 0u;
    if (a == 0) { len = 2; return 0u; }
    const int e = min(31 - __clz(a), 11);
    const uint32_t rem = (a - (1u << e)) & ((1u << e) - 1u);
    uint32_t head;
    if (e <= 4) { head = ((uint32_t)(e + 2) << 1) | s; len = 4 + e; }
    else { head = (((1u << (e - 2)) - 1u) << 2) | s; len = 2 * e; }
    return (head << e) | rem;
}

constexpr int EN_THREADS = 128;

// (1) bit length of every block (the P-frame macroblock header 1 + VLC(mvd.x) + VLC(mvd.y) is charged to block 0)
__global__ void __launch_bounds__(EN_THREADS) entropy_size_kernel(Geom g, FramePtrs p, EntropyPtrs e, int gop_len)
{
    const int nblk = g.nmb * 6;
    const int idx = blockIdx.x * EN_THREADS + threadIdx.x;
    if (idx >= nblk) return;
    const size_t f = blockIdx.y;
    const bool intra = (f % gop_len) == 0;
    const int mb = idx / 6, k = idx - mb * 6;
    const size_t m = f * g.nmb + mb;
    const uint4* lv = (const uint4*)(p.levels + (m * 6 + k) * 64);
    int bits = 1;                                           // ACflag
    if (intra && k < 4) bits += 2;                          // MPMFlag + intraPredMode
    if (!intra && k == 0) {
        const int mvw = *(const int*)(p.mvd + m * 2);
        bits += 1 + vlc_len((int)(int16_t)(mvw & 0xffff)) + vlc_len(mvw >> 16);
    }
    const int acflag = p.acflag[m * 6 + k];
    const uint4 first = __ldg(lv);
    bits += vlc_len((int)(int16_t)(first.x & 0xffffu));     // DC level
    if (acflag) bits += 63;                                 // 63 zero bits (ENC:5069-5073)
    else {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint4 q = i == 0 ? first : __ldg(lv + i);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (!(i == 0 && j == 0)) bits += vlc_len((int)(int16_t)(w[j] & 0xffffu));
                bits += vlc_len((int)(int16_t)(w[j] >> 16));
            }
        }
    }
    e.blkbits[f * nblk + idx] = (uint32_t)bits;
}

// (2a) exclusive scan of the block lengths of one frame (in place) + frame total
__global__ void __launch_bounds__(256) entropy_frame_scan_kernel(Geom g, EntropyPtrs e)
{
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_carry;
    const int nblk = g.nmb * 6;
    const size_t f = blockIdx.x;
    uint32_t* v = e.blkbits + f * nblk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nblk; base += 256) {
        const int i = base + threadIdx.x;
        const uint32_t x = i < nblk ? v[i] : 0u;
        uint32_t incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t woff = 0;
        for (int w = 0; w < warp; w++) woff += s_warp[w];
        const uint32_t carry = s_carry;
        if (i < nblk) v[i] = carry + woff + incl - x;
        __syncthreads();
        if (threadIdx.x == 255) s_carry = carry + woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) e.framebits[f] = s_carry;
}

// (2b) per stream: exclusive scan of its frames' lengths (in place) + stream total.  One warp per stream.
__global__ void __launch_bounds__(32) entropy_stream_scan_kernel(EntropyPtrs e, int frames_per_stream)
{
    const int s = blockIdx.x, lane = threadIdx.x;
    unsigned long long* v = e.framebits + (size_t)s * frames_per_stream;
    unsigned long long carry = 0;
    for (int base = 0; base < frames_per_stream; base += 32) {
        const int i = base + lane;
        const unsigned long long x = i < frames_per_stream ? v[i] : 0ull;
        unsigned long long incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        if (i < frames_per_stream) v[i] = carry + incl - x;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) e.streambits[s] = carry;
}

// (2c) byte offsets of the streams inside the chunk region (16-byte aligned starts), total, overflow check
__global__ void __launch_bounds__(32) entropy_stream_offsets_kernel(EntropyPtrs e, int n_streams)
{
    if (threadIdx.x != 0) return;
    unsigned long long off = 0;
    for (int s = 0; s < n_streams; s++) {
        e.streamoff[s] = off;
        off += ((e.streambits[s] + 7) / 8 + 1 + 15) & ~15ull;   // +1: the reference always writes bits/8 + 1 bytes
    }
    if (off > e.cap_bytes) { *e.overflow = 1; off = 0; }
    *e.total = off;
}

__global__ void __launch_bounds__(256) entropy_zero_kernel(EntropyPtrs e)
{
    const unsigned long long n16 = *e.total / 16;
    uint4* dst = (uint4*)e.bits;
    for (unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x; i < n16; i += (unsigned long long)gridDim.x * 256)
        dst[i] = make_uint4(0, 0, 0, 0);
}

// (3) one thread per block: re-encode and write.  `pos` = absolute bit position inside the chunk region.
struct BitSink {
    uint32_t* words;           // region as big-endian 32-bit words (stored byte-swapped)
    unsigned long long acc;    // pending bits, right aligned
    int nacc;                  // number of pending bits (< 32 between appends)
    size_t w;                  // next word index
    bool first;
    __device__ __forceinline__ void emit(uint32_t be_word, bool shared)
    {
        const uint32_t le = __byte_perm(be_word, 0, 0x0123);
        if (shared) atomicOr(words + w, le); else words[w] = le;
        w++;
    }
    __device__ __forceinline__ void put(uint32_t code, int len)
    {
        acc = (acc << len) | code;
        nacc += len;
        if (nacc >= 32) {
            nacc -= 32;
            emit((uint32_t)(acc >> nacc), first);
            first = false;
        }
    }
    __device__ __forceinline__ void flush()
    {
        if (nacc > 0) emit((uint32_t)(acc << (32 - nacc)), true);   // left-align the tail; shared with the next block
    }
};

__global__ void __launch_bounds__(EN_THREADS) entropy_pack_kernel(Geom g, FramePtrs p, EntropyPtrs e, int gop_len, int frames_per_stream)
{
    // (len << 16 | code) of every level in [-31, 31]: almost all symbols; larger ones take the arithmetic path
    __shared__ uint32_t s_vlc[64];
    if (threadIdx.x < 63) { int len; const uint32_t c = vlc_code((int)threadIdx.x - 31, len); s_vlc[threadIdx.x] = ((uint32_t)len << 16) | c; }
    __syncthreads();
    auto put_level = [&](BitSink& sk, int v) {
        if ((unsigned)(v + 31) < 63u) { const uint32_t t = s_vlc[v + 31]; sk.put(t & 0xffffu, (int)(t >> 16)); }
        else { int len; const uint32_t c = vlc_code(v, len); sk.put(c, len); }
    };
    const int nblk = g.nmb * 6;
    const int idx = blockIdx.x * EN_THREADS + threadIdx.x;
    if (idx >= nblk || *e.overflow) return;
    const size_t f = blockIdx.y;
    const bool intra = (f % gop_len) == 0;
    const int mb = idx / 6, k = idx - mb * 6;
    const size_t m = f * g.nmb + mb;
    const int s = (int)(f / frames_per_stream);
    const unsigned long long pos = e.streamoff[s] * 8ull + e.framebits[f] + e.blkbits[f * nblk + idx];
    BitSink sink;
    sink.words = (uint32_t*)e.bits;
    sink.w = (size_t)(pos >> 5);
    sink.nacc = (int)(pos & 31);       // leading bits belong to the previous block: zeros here, merged by atomicOr
    sink.acc = 0;
    sink.first = true;
    const uint4* lv = (const uint4*)(p.levels + (m * 6 + k) * 64);
    int len;
    uint32_t code;
    if (!intra && k == 0) {
        const int mvw = *(const int*)(p.mvd + m * 2);
        sink.put(1u, 1);                                                     // mv mode flag (ENC:5151)
        code = vlc_code((int)(int16_t)(mvw & 0xffff), len); sink.put(code, len);
        code = vlc_code(mvw >> 16, len); sink.put(code, len);
    }
    if (intra && k < 4) sink.put(((uint32_t)p.mpm[m * 4 + k] << 1) | p.ipm[m * 4 + k], 2);
    const int acflag = p.acflag[m * 6 + k];
    const uint4 first = __ldg(lv);
    code = vlc_code((int)(int16_t)(first.x & 0xffffu), len); sink.put(code, len);
    sink.put((uint32_t)acflag, 1);
    if (acflag) { sink.put(0u, 31); sink.put(0u, 31); sink.put(0u, 1); }   // 63 zero bits (ENC:5069-5073)
    else {
#pragma unroll 1
        for (int i = 0; i < 8; i++) {
            const uint4 q = i == 0 ? first : __ldg(lv + i);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (w[j] == 0u && !(i == 0 && j == 0)) { sink.put(0u, 4); continue; }     // two zero levels: "00" "00"
                if (!(i == 0 && j == 0)) put_level(sink, (int)(int16_t)(w[j] & 0xffffu));
                put_level(sink, (int)(int16_t)(w[j] >> 16));
            }
        }
    }
    sink.flush();
}

// =====================================================================================================
// f3 of SURVEY.md §8: the decoder's bit reader (readBlockData + DC/AC/MVientropy, DEC:38-2025) on the GPU.
// The format has no resynchronisation points (nothing is byte aligned, no start codes), so a bare stream is one serial
// VLC chain.  The encoder, however, knows where everything starts: (2a)/(2b) above are exactly the bit offsets of every
// block.  row_index_kernel exports the offset of every MACROBLOCK ROW (8 bytes per row: 0.4 % of the stream, the
// "<bin>.idx" side-car of icspenc --index); with it parse_rows_kernel runs one independent chain per (frame, MB row):
// 19 200 CIF frames = 345 600 threads instead of 64.  Without an index the C++ parser (host/bitstream.cpp) is used.
// =====================================================================================================
__global__ void __launch_bounds__(128) row_index_kernel(Geom g, EntropyPtrs e, unsigned long long* __restrict__ rows, int n_frames)
{
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n_frames * g.mbh) return;
    const int f = i / g.mbh, y = i - f * g.mbh;
    rows[i] = e.framebits[f] + e.blkbits[(size_t)f * g.nmb * 6 + (size_t)y * g.mbw * 6];
}

// MSB-first bit source over a stream body in global memory.  `buf` holds `nb` (> 32 between calls) valid bits, left
// aligned.  Bits at or after 8*nbytes read as 0, like the host reader.
struct BitSource {
    const uint32_t* w;
    unsigned long long nbytes, widx, buf;
    uint32_t nx0, nx1;      // the next two words, already loaded: the global-memory latency of a refill overlaps ~64 bits of parsing
    int nb;
    __device__ __forceinline__ uint32_t load(unsigned long long i) const
    {
        uint32_t v = 0;
        if (i * 4 < nbytes) {
            v = __byte_perm(__ldg(w + i), 0, 0x0123);
            const unsigned long long rem = nbytes - i * 4;              // valid bytes in this word
            if (rem < 4) v &= 0xffffffffu << (8 * (4 - (int)rem));
        }
        return v;
    }
    __device__ __forceinline__ uint32_t fetch()
    {
        const uint32_t v = nx0;
        nx0 = nx1;
        nx1 = load(widx++);
        return v;
    }
    __device__ __forceinline__ void init(const uint8_t* body, unsigned long long nbytes_, unsigned long long pos)
    {
        w = (const uint32_t*)body; nbytes = nbytes_; widx = pos >> 5;
        const int sh = (int)(pos & 31);
        const uint32_t a = load(widx), b = load(widx + 1);
        nx0 = load(widx + 2); nx1 = load(widx + 3);
        widx += 4;
        buf = (((unsigned long long)a << 32) | b) << sh;
        nb = 64 - sh;
    }
    __device__ __forceinline__ uint32_t top() const { return (uint32_t)(buf >> 32); }
    __device__ __forceinline__ void skip(int k)                          // k <= 32
    {
        buf <<= k; nb -= k;
        if (nb <= 32) { buf |= (unsigned long long)fetch() << (32 - nb); nb += 32; }
    }
    __device__ __forceinline__ int bit() { const int v = (int)(buf >> 63); skip(1); return v; }
    // DCientropy (DEC:407-608) and its AC / MV twins: 3-bit category prefix, unary extension from category 6 up
    __device__ __forceinline__ int vlc()
    {
        const uint32_t t = top();
        const uint32_t top3 = t >> 29;
        if (top3 < 2) { skip(2); return 0; }
        if (top3 == 2) { skip(4); return (t >> 28) & 1 ? 1 : -1; }
        int e, prefix;
        if (top3 != 7) { e = (int)top3 - 2; prefix = 3; }
        else {
            const int ones = __clz((int)~t);
            if (ones >= 10) return 0;              // no category matches: the reference leaves len = 0, val = 0
            e = ones + 2; prefix = ones + 1;
        }
        const int sgn = (int)((t >> (31 - prefix)) & 1u);
        const int rem = (int)((t << (prefix + 1)) >> (32 - e));          // prefix + 1 + e <= 22 bits
        skip(prefix + 1 + e);
        const int v = (1 << e) + rem;
        return sgn ? v : -v;
    }
};

struct ParsePtrs {
    const uint8_t* bits;                    // bodies (file bytes after the 14-byte header)
    const unsigned long long* streamoff;    // [S] byte offset of each body inside bits (multiple of 4)
    const unsigned long long* streambytes;  // [S] body length in bytes
    const unsigned long long* rows;         // [F][mbh] bit offset of every MB row from its body start
};
// one thread per (frame, macroblock row); levels must be pre-zeroed (only non-zero levels are stored)
__global__ void __launch_bounds__(64) parse_rows_kernel(Geom g, FramePtrs p, ParsePtrs q, int gop_len, int frames_per_stream,
                                                        int s_first, int n_frames)
{
    const int i = blockIdx.x * 64 + threadIdx.x;
    if (i >= n_frames * g.mbh) return;
    const int f = i / g.mbh, y = i - f * g.mbh;       // f: frame inside the chunk (p.* are chunk-relative)
    const int s = s_first + f / frames_per_stream;
    const bool intra = (f % gop_len) == 0;
    BitSource b;
    b.init(q.bits + q.streamoff[s], q.streambytes[s], q.rows[i]);
    for (int mbx = 0; mbx < g.mbw; mbx++) {
        const size_t m = (size_t)f * g.nmb + (size_t)y * g.mbw + mbx;
        if (!intra) {
            b.skip(1);                                                  // MVmodeflag (DEC:301)
            const int mx = b.vlc(), my = b.vlc();
            *(uint32_t*)(p.mvd + m * 2) = ((uint32_t)mx & 0xffffu) | ((uint32_t)my << 16);
        }
        for (int k = 0; k < 6; k++) {
            if (intra && k < 4) { p.mpm[m * 4 + k] = (uint8_t)b.bit(); p.ipm[m * 4 + k] = (uint8_t)b.bit(); }
            int16_t* zz = p.levels + (m * 6 + k) * 64;
            const int dc = b.vlc();
            if (dc) zz[0] = (int16_t)dc;
            if (b.bit()) { b.skip(31); b.skip(32); continue; }          // ACflag: 63 zero bits follow
            for (int c = 1; c < 64;) {
                const int z = min(__clz((int)b.top()) >> 1, min(16, 64 - c));   // run of "00" symbols
                if (z) { b.skip(2 * z); c += z; }
                if (c < 64) { const int v = b.vlc(); if (v) zz[c] = (int16_t)v; c++; }
            }
        }
    }
}

}  // namespace icsp
