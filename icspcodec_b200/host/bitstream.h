// Host-side ICSP bitstream writer / reader (the part of the codec that stays on the CPU).
// Format = what the reference's makebitstream (ENC:4849-4922), intraBody (ENC:5032-5131), interBody
// (ENC:5132-5236) and DCentropy/ACentropy/MVentropy (ENC:5417-6334) emit, and what readHeader/readBlockData
// (DEC:14-404) consume; restated in SURVEY.md R14/A.10.  Input/output is the SoA syntax of include/icspcuda.h.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace icsp_host {

struct StreamParams {
    int width = 352, height = 288;
    int qp_dc = 8, qp_ac = 8;
    int intra_period = 0;   // 0 = all intra (encoder side); the decoder treats 1 as all intra (DEC:98)
    int nframes = 0;
};

struct Syntax {            // views into SoA arrays, layout of include/icspcuda.h
    const int16_t* levels; // [n][nmb][6][64]
    const uint8_t* acflag; // [n][nmb][6]
    const uint8_t* mpm;    // [n][nmb][4]
    const uint8_t* ipm;    // [n][nmb][4]
    const int16_t* mvd;    // [n][nmb][2]
};

// MSB-first bit string with a 64-bit accumulator; segments can be concatenated at arbitrary bit offsets.
class BitString {
public:
    void put(uint32_t value, int nbits);      // nbits in [0,32], value's low nbits, MSB first
    void append(const BitString& other);      // bit-level concatenation
    void append_msb_bytes(const uint8_t* p, uint64_t nbits);   // first nbits bits of an MSB-first byte string
    uint64_t size_bits() const { return nbits_; }
    // bytes of the reference file body: floor(bits/8) full bytes + one last byte holding the tail bits in its LOW
    // bits (the reference shifts bits into each byte from the right and never left-aligns the tail; an extra zero
    // byte appears when bits % 8 == 0; ENC:4895)
    std::vector<uint8_t> reference_body() const;
private:
    void flush_word();
    std::vector<uint64_t> words_;  // full 64-bit words, MSB first
    uint64_t acc_ = 0;             // pending bits, right aligned
    int accbits_ = 0;
    uint64_t nbits_ = 0;
};

// category VLC shared by DC, AC and MV values
void put_vlc(BitString& bs, int v);

// one frame (intra: MPM/mode bits + 6 blocks per MB; inter: mv flag + MVD + 6 blocks per MB)
// row_bits (optional, [height/16]): bit offset of every macroblock row inside the frame's bit string (mbw = width/16)
void encode_frame(BitString& bs, const Syntax& s, int frame, int nmb, bool intra, int mbw = 0, uint64_t* row_bits = nullptr);

// 14-byte header (ENC.h:201-212 packed, ENC:4901-4922)
std::vector<uint8_t> stream_header(const StreamParams& p);

// whole stream; frames are entropy coded in parallel (n_threads) and merged in order
// row_index (optional): [nframes][height/16] bit offsets from the body start, the side-car of icspenc --index / icsp_bits_row_index
std::vector<uint8_t> write_stream(const StreamParams& p, const Syntax& s, int n_threads, std::vector<uint64_t>* row_index = nullptr);

// ---- reader ----------------------------------------------------------------------------------------
struct ParsedStream {
    StreamParams p;
    std::vector<int16_t> levels;
    std::vector<uint8_t> acflag, mpm, ipm;
    std::vector<int16_t> mvd;
};
// Parses header + body MSB-first, every byte including the last one (DEC:60-74).  Throws std::runtime_error.
// row_index (optional): restart every macroblock row at its recorded bit offset — the chains the GPU bit reader
// (parse_rows_kernel) follows, on the host; with a correct index the result equals the serial parse.
ParsedStream parse_stream(const std::vector<uint8_t>& file, int nframes, const std::vector<uint64_t>* row_index = nullptr);

}  // namespace icsp_host
