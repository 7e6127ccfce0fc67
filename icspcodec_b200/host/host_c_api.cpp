// C entry points over the host bitstream writer/reader so that tests (ctypes) can exercise exactly the code
// icspenc/icspdec use.
#include <cstring>
#include <stdexcept>

#include "bitstream.h"

extern "C" {

// returns bytes written, -1 if cap is too small
long icsp_host_write_stream(const int16_t* levels, const uint8_t* acflag, const uint8_t* mpm, const uint8_t* ipm, const int16_t* mvd,
                            int nframes, int w, int h, int qdc, int qac, int ip, int threads, uint8_t* out, long cap)
{
    icsp_host::StreamParams p;
    p.width = w; p.height = h; p.qp_dc = qdc; p.qp_ac = qac; p.intra_period = ip; p.nframes = nframes;
    const icsp_host::Syntax s{levels, acflag, mpm, ipm, mvd};
    const std::vector<uint8_t> v = icsp_host::write_stream(p, s, threads);
    if ((long)v.size() > cap) return -1;
    memcpy(out, v.data(), v.size());
    return (long)v.size();
}

// the same + the macroblock-row index rows[nframes][h/16] (bit offsets from the body start)
long icsp_host_write_stream_indexed(const int16_t* levels, const uint8_t* acflag, const uint8_t* mpm, const uint8_t* ipm, const int16_t* mvd,
                                    int nframes, int w, int h, int qdc, int qac, int ip, int threads, uint8_t* out, long cap, uint64_t* rows)
{
    icsp_host::StreamParams p;
    p.width = w; p.height = h; p.qp_dc = qdc; p.qp_ac = qac; p.intra_period = ip; p.nframes = nframes;
    const icsp_host::Syntax s{levels, acflag, mpm, ipm, mvd};
    std::vector<uint64_t> idx;
    const std::vector<uint8_t> v = icsp_host::write_stream(p, s, threads, &idx);
    if ((long)v.size() > cap) return -1;
    memcpy(out, v.data(), v.size());
    memcpy(rows, idx.data(), idx.size() * sizeof(uint64_t));
    return (long)v.size();
}

// parse with a row index (every macroblock row restarts at its recorded offset); returns 0 / -1
int icsp_host_parse_stream_indexed(const uint8_t* bin, long len, int nframes, const uint64_t* rows, long nrows, int16_t* levels, uint8_t* acflag,
                                   uint8_t* mpm, uint8_t* ipm, int16_t* mvd)
{
    try {
        const std::vector<uint8_t> file(bin, bin + len);
        const std::vector<uint64_t> idx(rows, rows + nrows);
        const icsp_host::ParsedStream ps = icsp_host::parse_stream(file, nframes, &idx);
        memcpy(levels, ps.levels.data(), ps.levels.size() * 2);
        memcpy(acflag, ps.acflag.data(), ps.acflag.size());
        memcpy(mpm, ps.mpm.data(), ps.mpm.size());
        memcpy(ipm, ps.ipm.data(), ps.ipm.size());
        memcpy(mvd, ps.mvd.data(), ps.mvd.size() * 2);
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}

// hdr = {w, h, qdc, qac, ip}; arrays may be NULL for a header-only call. returns 0 / -1
int icsp_host_parse_stream(const uint8_t* bin, long len, int nframes, int* hdr, int16_t* levels, uint8_t* acflag, uint8_t* mpm,
                           uint8_t* ipm, int16_t* mvd)
{
    try {
        const std::vector<uint8_t> file(bin, bin + len);
        const icsp_host::ParsedStream ps = icsp_host::parse_stream(file, levels ? nframes : 0);
        hdr[0] = ps.p.width; hdr[1] = ps.p.height; hdr[2] = ps.p.qp_dc; hdr[3] = ps.p.qp_ac; hdr[4] = ps.p.intra_period;
        if (levels) {
            memcpy(levels, ps.levels.data(), ps.levels.size() * 2);
            memcpy(acflag, ps.acflag.data(), ps.acflag.size());
            memcpy(mpm, ps.mpm.data(), ps.mpm.size());
            memcpy(ipm, ps.ipm.data(), ps.ipm.size());
            memcpy(mvd, ps.mvd.data(), ps.mvd.size() * 2);
        }
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}
}
