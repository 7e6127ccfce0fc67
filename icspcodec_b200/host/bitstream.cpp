#include "bitstream.h"

#include <cstring>
#include <algorithm>
#include <cstdlib>
#include <stdexcept>
#include <thread>

namespace icsp_host {

void BitString::flush_word()
{
    words_.push_back(acc_);
    acc_ = 0;
    accbits_ = 0;
}

void BitString::put(uint32_t value, int nbits)
{
    if (nbits <= 0) return;
    const uint64_t v = (nbits == 32) ? value : (value & ((1u << nbits) - 1u));
    const int room = 64 - accbits_;
    if (nbits < room) {
        acc_ = (acc_ << nbits) | v;
        accbits_ += nbits;
    } else {
        const int rest = nbits - room;            // bits that do not fit
        acc_ = (room == 64) ? (v >> rest) : ((acc_ << room) | (v >> rest));
        flush_word();
        if (rest) { acc_ = v & ((1ull << rest) - 1ull); accbits_ = rest; }
    }
    nbits_ += nbits;
}

void BitString::append(const BitString& o)
{
    for (uint64_t w : o.words_) { put((uint32_t)(w >> 32), 32); put((uint32_t)w, 32); }
    if (o.accbits_ > 32) { put((uint32_t)(o.acc_ >> 32), o.accbits_ - 32); put((uint32_t)o.acc_, 32); }
    else put((uint32_t)o.acc_, o.accbits_);
}

void BitString::append_msb_bytes(const uint8_t* p, uint64_t nbits)
{
    uint64_t i = 0;
    for (; i + 32 <= nbits; i += 32, p += 4) put(((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3], 32);
    for (; i + 8 <= nbits; i += 8, p++) put(*p, 8);
    if (i < nbits) put((uint32_t)(*p >> (8 - (nbits - i))), (int)(nbits - i));
}

std::vector<uint8_t> BitString::reference_body() const
{
    std::vector<uint8_t> out;
    out.reserve(nbits_ / 8 + 1);
    for (uint64_t w : words_)
        for (int b = 7; b >= 0; b--) out.push_back((uint8_t)(w >> (8 * b)));
    int left = accbits_;
    while (left >= 8) { out.push_back((uint8_t)(acc_ >> (left - 8))); left -= 8; }
    out.push_back(left ? (uint8_t)(acc_ & ((1u << left) - 1u)) : (uint8_t)0);   // right-aligned tail / extra zero byte
    return out;
}

void put_vlc(BitString& bs, int v)
{
    const unsigned s = v >= 0 ? 1u : 0u;
    const unsigned a = (unsigned)std::abs(v);
    if (a == 0) { bs.put(0, 2); return; }
    if (a == 1) { bs.put(0x4u | s, 4); return; }          // 0 1 0 s
    int e = 31 - __builtin_clz(a);
    if (e > 11) e = 11;                                    // last category: a >= 2048, 11 low bits of a-2048
    const unsigned c = (a - (1u << e)) & ((1u << e) - 1u);
    if (e <= 4) bs.put((((unsigned)(e + 2) << 1) | s), 4);            // 011/100/101/110 then sign
    else bs.put((((1u << (e - 2)) - 1u) << 2) | s, e);                // (e-2) ones, 0, sign
    bs.put(c, e);
}

static inline void put_block(BitString& bs, const int16_t* zz, int acflag)
{
    put_vlc(bs, zz[0]);
    bs.put((unsigned)acflag, 1);
    if (acflag == 1) { bs.put(0, 32); bs.put(0, 31); }                 // 63 zero bits (ENC:5069-5073)
    else for (int k = 1; k < 64; k++) put_vlc(bs, zz[k]);
}

void encode_frame(BitString& bs, const Syntax& s, int frame, int nmb, bool intra, int mbw, uint64_t* row_bits)
{
    for (int mb = 0; mb < nmb; mb++) {
        const size_t m = (size_t)frame * nmb + mb;
        if (row_bits && mbw > 0 && mb % mbw == 0) row_bits[mb / mbw] = bs.size_bits();   // bit offset of this MB row inside the frame
        if (!intra) {
            bs.put(1, 1);                                              // mv mode flag (ENC:5151)
            put_vlc(bs, s.mvd[m * 2]);
            put_vlc(bs, s.mvd[m * 2 + 1]);
        }
        for (int k = 0; k < 6; k++) {
            if (intra && k < 4) bs.put(((unsigned)s.mpm[m * 4 + k] << 1) | s.ipm[m * 4 + k], 2);
            put_block(bs, s.levels + (m * 6 + k) * 64, s.acflag[m * 6 + k]);
        }
    }
}

std::vector<uint8_t> stream_header(const StreamParams& p)
{
    std::vector<uint8_t> out(14);
    // "\0ICSP", u16 height, u16 width, QP_DC, QP_AC, DPCMmode=0, outro (all little endian, packed)
    out[0] = 0; out[1] = 'I'; out[2] = 'C'; out[3] = 'S'; out[4] = 'P';
    out[5] = (uint8_t)(p.height & 255); out[6] = (uint8_t)(p.height >> 8);
    out[7] = (uint8_t)(p.width & 255);  out[8] = (uint8_t)(p.width >> 8);
    out[9] = (uint8_t)p.qp_dc; out[10] = (uint8_t)p.qp_ac; out[11] = 0;
    const unsigned outro = ((unsigned)p.intra_period & 63u) << 7;     // 6 bits of intraPeriod, then 7 zero bits
    out[12] = (uint8_t)(outro & 255); out[13] = (uint8_t)(outro >> 8);
    return out;
}

std::vector<uint8_t> write_stream(const StreamParams& p, const Syntax& s, int n_threads, std::vector<uint64_t>* row_index)
{
    const int nmb = (p.width / 16) * (p.height / 16), mbw = p.width / 16, mbh = p.height / 16;
    std::vector<BitString> per_frame((size_t)p.nframes);
    if (row_index) row_index->assign((size_t)p.nframes * mbh, 0);
    n_threads = std::max(1, std::min(n_threads, p.nframes));
    auto work = [&](int tid) {
        for (int n = tid; n < p.nframes; n += n_threads) {
            const bool intra = p.intra_period == 0 || n % p.intra_period == 0;
            encode_frame(per_frame[n], s, n, nmb, intra, mbw, row_index ? row_index->data() + (size_t)n * mbh : nullptr);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; t++) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    BitString all;
    for (size_t n = 0; n < per_frame.size(); n++) {
        if (row_index) for (int y = 0; y < mbh; y++) (*row_index)[n * mbh + y] += all.size_bits();   // frame offset inside the body
        all.append(per_frame[n]);
    }

    std::vector<uint8_t> out = stream_header(p);
    const std::vector<uint8_t> body = all.reference_body();
    out.insert(out.end(), body.begin(), body.end());
    return out;
}

// ---- reader ------------------------------------------------------------------------------------------
namespace {
struct BitReader {
    // `p` must be readable for 8 bytes past the last whole byte of the stream (parse_stream pads its copy with zeros);
    // bits at or after nbits read as 0, as in the reference's reader on its zero-initialised buffer.
    const uint8_t* p; uint64_t nbits, pos = 0;
    // the next >= 57 bits, left aligned (bit 63 = the bit at pos)
    uint64_t window() const
    {
        if (pos >= nbits) return 0;
        uint64_t w;
        memcpy(&w, p + (pos >> 3), 8);
        return __builtin_bswap64(w) << (pos & 7);
    }
    int get() { const int v = (int)(window() >> 63); pos++; return v; }
    int vlc()
    {   // DCientropy (DEC:407-608) and its AC/MV twins: 3-bit category prefix, then unary extension from category 6 up
        const uint64_t w = window();
        const unsigned top3 = (unsigned)(w >> 61);
        if (top3 < 2) { pos += 2; return 0; }                                   // "00"
        if (top3 == 2) { pos += 4; return (w >> 60) & 1 ? 1 : -1; }             // "010" + sign
        int e, prefix;
        if (top3 != 7) { e = (int)top3 - 2; prefix = 3; }
        else {
            const int ones = ~w ? __builtin_clzll(~w) : 64;                     // >= 3 (clz(0) is undefined)
            if (ones >= 10) return 0;              // no category matches: the reference leaves len = 0, val = 0
            e = ones + 2; prefix = ones + 1;
        }
        // prefix + 1 + e <= 10 + 1 + 11 = 22 bits: inside the window
        const int s = (int)((w >> (63 - prefix)) & 1);
        const int t = (int)((w << (prefix + 1)) >> (64 - e));
        pos += prefix + 1 + e;
        const int v = (1 << e) + t;
        return s ? v : -v;
    }
    // number of leading "00" symbols (at most `limit` <= 28) at pos, consumed
    int zero_run(int limit)
    {
        const uint64_t w = window() | ((1ull << (63 - 2 * limit)) );            // sentinel one after `limit` pairs
        const int z = __builtin_clzll(w) >> 1;
        pos += 2 * (uint64_t)z;
        return z;
    }
};
}  // namespace

ParsedStream parse_stream(const std::vector<uint8_t>& file, int nframes, const std::vector<uint64_t>* row_index)
{
    if (file.size() < 14) throw std::runtime_error("bitstream shorter than its 14-byte header");
    ParsedStream ps;
    ps.p.height = file[5] | (file[6] << 8);
    ps.p.width = file[7] | (file[8] << 8);
    ps.p.qp_dc = file[9]; ps.p.qp_ac = file[10];
    const unsigned outro = file[12] | (file[13] << 8);
    ps.p.intra_period = (outro & 0x1F80) >> 7;                        // DEC:29
    ps.p.nframes = nframes;
    if (ps.p.width <= 0 || ps.p.height <= 0 || (ps.p.width & 15) || (ps.p.height & 15)) throw std::runtime_error("bad geometry in header");
    if (ps.p.intra_period < 1) throw std::runtime_error("intraPeriod 0 in header: the reference decoder divides by zero (DEC:201); encode with --intraPeriod 1 for all-intra");
    if (ps.p.qp_dc == 0 || ps.p.qp_ac == 0) throw std::runtime_error("QP 0 in header");
    const int nmb = (ps.p.width / 16) * (ps.p.height / 16);
    const size_t N = (size_t)nframes * nmb;
    ps.levels.assign(N * 384, 0); ps.acflag.assign(N * 6, 0); ps.mpm.assign(N * 4, 0); ps.ipm.assign(N * 4, 0); ps.mvd.assign(N * 2, 0);
    std::vector<uint8_t> padded(file.size() - 14 + 16, 0);              // the reader loads 8 bytes at a time
    memcpy(padded.data(), file.data() + 14, file.size() - 14);
    BitReader r{padded.data(), (uint64_t)(file.size() - 14) * 8};
    const int mbw = ps.p.width / 16, mbh = ps.p.height / 16;
    if (row_index && row_index->size() < (size_t)nframes * mbh) throw std::runtime_error("row index shorter than nframes x macroblock rows");
    for (int n = 0; n < nframes; n++) {
        const bool intra = ps.p.intra_period == 1 || n % ps.p.intra_period == 0;   // DEC:98, DEC:201
        for (int mb = 0; mb < nmb; mb++) {
            const size_t m = (size_t)n * nmb + mb;
            // with an index every macroblock row starts from its recorded offset (the chain the GPU bit reader follows)
            if (row_index && mb % mbw == 0) r.pos = (*row_index)[(size_t)n * mbh + mb / mbw];
            if (!intra) {
                (void)r.get();                                          // MVmodeflag (DEC:301)
                ps.mvd[m * 2] = (int16_t)r.vlc();
                ps.mvd[m * 2 + 1] = (int16_t)r.vlc();
            }
            for (int k = 0; k < 6; k++) {
                if (intra && k < 4) { ps.mpm[m * 4 + k] = (uint8_t)r.get(); ps.ipm[m * 4 + k] = (uint8_t)r.get(); }
                int16_t* zz = &ps.levels[(m * 6 + k) * 64];
                zz[0] = (int16_t)r.vlc();
                const int f = r.get();
                ps.acflag[m * 6 + k] = (uint8_t)f;
                if (f == 1) r.pos += 63;
                else for (int q = 1; q < 64;) {
                    q += r.zero_run(std::min(28, 64 - q));                      // levels are pre-zeroed: skip runs of "00"
                    if (q < 64) zz[q++] = (int16_t)r.vlc();
                }
            }
        }
    }
    return ps;
}

}  // namespace icsp_host
