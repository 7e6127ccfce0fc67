"""Thin Python binding of the C ABI in include/icspcuda.h (used by tests/ and bench.py).

The reference is compiled C++ with no plugin API, so the *product* host side is C++ (icspcodec_b200/host:
icspenc / icspdec).  This module only marshals numpy arrays into the same extern "C" calls.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib


class IcspError(RuntimeError):
    pass


@dataclass
class EncResult:
    levels: np.ndarray   # int16 [n][nmb][6][64]
    acflag: np.ndarray   # uint8 [n][nmb][6]
    mpm: np.ndarray      # uint8 [n][nmb][4]
    ipm: np.ndarray      # uint8 [n][nmb][4]
    mvd: np.ndarray      # int16 [n][nmb][2]
    mv: np.ndarray       # int16 [n][nmb][2]
    minsad: np.ndarray   # int32 [n][nmb]
    recon: np.ndarray    # uint8 [n][fb]
    dct: np.ndarray | None = None   # f64 [n][nmb][6][64] debug tap (only when asked for)


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class PinnedArray:
    """numpy view over cudaHostAlloc'd memory (kept alive by this object)."""

    def __init__(self, shape, dtype, upload_only: bool = False):
        """upload_only: write-combined memory (icsp_host_alloc_upload) for buffers the CPU only writes."""
        self._lib = _lib.load()
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = (self._lib.icsp_host_alloc_upload if upload_only else self._lib.icsp_host_alloc)(max(nbytes, 1))
        if not self._p:
            raise IcspError("icsp_host_alloc failed")
        buf = (C.c_char * max(nbytes, 1)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            if getattr(self, "_p", None):
                self.array = None
                self._lib.icsp_host_free(self._p)
                self._p = None
        except Exception:
            pass


def stream_header(w: int, h: int, qp_dc: int, qp_ac: int, intra_period: int) -> bytes:
    """14-byte packed header (ENC.h:201-212, ENC:4901-4922)."""
    outro = (intra_period & 63) << 7
    return bytes([0, 73, 67, 83, 80, h & 255, h >> 8, w & 255, w >> 8, qp_dc & 255, qp_ac & 255, 0, outro & 255, outro >> 8])


def finish_stream(body: np.ndarray, nbits: int, w: int, h: int, qp_dc: int, qp_ac: int, intra_period: int) -> bytes:
    """MSB-first body bytes (ceil(nbits/8), as icsp_encode_streams returns them) -> the reference-format file: 14-byte header
    + nbits/8+1 body bytes with the tail bits right-aligned in the last byte (makebitstream ENC:4873-4895)."""
    b = bytearray(bytes(body[: (nbits + 7) // 8]))
    if nbits % 8:
        b[-1] = b[-1] >> (8 - nbits % 8)
    else:
        b += b"\x00"
    return stream_header(w, h, qp_dc, qp_ac, intra_period) + bytes(b)


class IcspCuda:
    """One context = one GPU + one stream + device SoA buffers for up to `max_frames` frames."""

    def __init__(self, width: int = 352, height: int = 288, max_frames: int = 64, device: int = 0):
        self.lib = _lib.load()
        self.w, self.h, self.cap = width, height, max_frames
        self.nmb = (width // 16) * (height // 16)
        self.fb = width * height * 3 // 2
        h = C.c_void_p()
        rc = self.lib.icsp_create(C.byref(h), device, width, height, max_frames)
        if rc != 0:
            raise IcspError(f"icsp_create failed ({rc}): {self.lib.icsp_last_error(None).decode()}")
        self.h_ctx = h

    def close(self):
        if getattr(self, "h_ctx", None):
            self.lib.icsp_destroy(self.h_ctx)
            self.h_ctx = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, rc: int, what: str):
        if rc != 0:
            raise IcspError(f"{what} failed ({rc}): {self.lib.icsp_last_error(self.h_ctx).decode()}")

    # ---- encoder -----------------------------------------------------------------------------------
    def alloc_result(self, n: int, pinned: bool = False, with_dct: bool = False) -> EncResult:
        if with_dct:
            r = self.alloc_result(n, pinned)
            r.dct = np.zeros((n, self.nmb, 6, 64), np.float64)
            return r
        shapes = dict(levels=((n, self.nmb, 6, 64), np.int16), acflag=((n, self.nmb, 6), np.uint8),
                      mpm=((n, self.nmb, 4), np.uint8), ipm=((n, self.nmb, 4), np.uint8),
                      mvd=((n, self.nmb, 2), np.int16), mv=((n, self.nmb, 2), np.int16),
                      minsad=((n, self.nmb), np.int32), recon=((n, self.fb), np.uint8))
        if pinned:
            self._pins = getattr(self, "_pins", [])
            out = {}
            for k, (sh, dt) in shapes.items():
                pa = PinnedArray(sh, dt)
                self._pins.append(pa)
                out[k] = pa.array
            return EncResult(**out)
        return EncResult(**{k: np.zeros(sh, dt) for k, (sh, dt) in shapes.items()})

    @staticmethod
    def _enc_out(res: EncResult, fields=None) -> _lib.EncOut:
        o = _lib.EncOut()
        for name in ("levels", "acflag", "mpm", "ipm", "mvd", "mv", "minsad", "recon"):
            want = fields is None or name in fields
            setattr(o, name, _ptr(getattr(res, name)) if want else None)
        o.dct = _ptr(res.dct) if res.dct is not None and (fields is None or "dct" in fields) else None
        return o

    def encode_gops(self, frames: np.ndarray, n_gops: int, gop_len: int, qp_dc: int, qp_ac: int,
                    out: EncResult | None = None, fields=None, with_dct: bool = False) -> EncResult:
        frames = np.ascontiguousarray(frames, np.uint8)
        n = n_gops * gop_len
        if frames.size != n * self.fb:
            raise IcspError(f"expected {n} frames of {self.fb} bytes")
        res = out or self.alloc_result(n, with_dct=with_dct)
        o = self._enc_out(res, fields)
        self._chk(self.lib.icsp_encode_gops(self.h_ctx, _ptr(frames), n_gops, gop_len, qp_dc, qp_ac, C.byref(o)), "icsp_encode_gops")
        return res

    def intra_frame(self, frames: np.ndarray, qp_dc: int, qp_ac: int, with_dct: bool = False) -> EncResult:
        """intraPrediction (ENC:556-643) on independent frames."""
        frames = np.ascontiguousarray(frames, np.uint8).reshape(-1, self.fb)
        res = self.alloc_result(frames.shape[0], with_dct=with_dct)
        o = self._enc_out(res)
        self._chk(self.lib.icsp_intra_frame(self.h_ctx, _ptr(frames), frames.shape[0], qp_dc, qp_ac, C.byref(o)), "icsp_intra_frame")
        return res

    def inter_frame(self, cur: np.ndarray, prev_recon: np.ndarray, qp_dc: int, qp_ac: int, with_dct: bool = False) -> EncResult:
        """interPrediction(cur, prev) (ENC:1986-2072) on independent pairs: cur[i] coded against the reconstruction prev_recon[i]."""
        cur = np.ascontiguousarray(cur, np.uint8).reshape(-1, self.fb)
        prev_recon = np.ascontiguousarray(prev_recon, np.uint8).reshape(-1, self.fb)
        res = self.alloc_result(cur.shape[0], with_dct=with_dct)
        o = self._enc_out(res)
        self._chk(self.lib.icsp_inter_frame(self.h_ctx, _ptr(cur), _ptr(prev_recon), cur.shape[0], qp_dc, qp_ac, C.byref(o)), "icsp_inter_frame")
        return res

    def quant(self, dct: np.ndarray, qp_dc: int, qp_ac: int, chroma: bool = False):
        """Quantization_block / CQuantization_block (ENC:2750-2824, 4610-4656): (levels int32 [n][64] raster order, acflag uint8 [n])."""
        dct = np.ascontiguousarray(dct, np.float64).reshape(-1, 64)
        lv = np.zeros(dct.shape, np.int32)
        ac = np.zeros(dct.shape[0], np.uint8)
        self._chk(self.lib.icsp_quant(self.h_ctx, _ptr(dct), dct.shape[0], qp_dc, qp_ac, int(chroma), _ptr(lv), _ptr(ac)), "icsp_quant")
        return lv, ac

    def encode_sequence(self, frames: np.ndarray, qp_dc: int, qp_ac: int, intra_period: int) -> EncResult:
        """Frame loop of single_thread_encoding (ENC:217-245): I-frame iff n % intra_period == 0 (0 = all intra).
        Full GOPs go out as one batch, the tail (nframes % intra_period) as a second, shorter GOP."""
        frames = np.ascontiguousarray(frames, np.uint8).reshape(-1, self.fb)
        n = frames.shape[0]
        ip = 1 if intra_period == 0 else intra_period
        res = self.alloc_result(n)
        full = n // ip
        parts = []
        if full:
            parts.append((0, full, ip))
        if n - full * ip:
            parts.append((full * ip, 1, n - full * ip))
        for start, ng, gl in parts:
            cnt = ng * gl
            sub = EncResult(**{k: (None if getattr(res, k) is None else getattr(res, k)[start:start + cnt]) for k in res.__dataclass_fields__})
            self.encode_gops(frames[start:start + cnt], ng, gl, qp_dc, qp_ac, out=sub)
        return res

    def upload(self, frames: np.ndarray):
        n = frames.size // self.fb
        self._chk(self.lib.icsp_enc_upload(self.h_ctx, _ptr(frames), n), "icsp_enc_upload")

    def run(self, n_gops: int, gop_len: int, qp_dc: int, qp_ac: int):
        self._chk(self.lib.icsp_enc_run(self.h_ctx, n_gops, gop_len, qp_dc, qp_ac), "icsp_enc_run")

    def download(self, n: int, res: EncResult, fields=None):
        o = self._enc_out(res, fields)
        self._chk(self.lib.icsp_enc_download(self.h_ctx, n, C.byref(o)), "icsp_enc_download")

    def sync(self):
        self._chk(self.lib.icsp_sync(self.h_ctx), "icsp_sync")

    def enc_sse(self, n: int) -> np.ndarray:
        """[n][3] uint64 plane SSE (Y, Cb, Cr) of the resident frames against their reconstruction (§8 f4)."""
        sse = np.zeros((n, 3), np.uint64)
        self._chk(self.lib.icsp_enc_sse(self.h_ctx, n, _ptr(sse)), "icsp_enc_sse")
        return sse

    def psnr_y(self, n: int) -> float:
        """Average luma PSNR as the reference decoder logs it (DEC.h:332-348)."""
        mse = self.enc_sse(n)[:, 0].astype(np.float64) / float(self.w * self.h)
        with np.errstate(divide="ignore"):
            return float(np.mean(20.0 * np.log10(255.0 / np.sqrt(mse))))

    # ---- encoder with GPU entropy coding ------------------------------------------------
    def encode_streams(self, frames: np.ndarray, n_streams: int, gops_per_stream: int, gop_len: int, qp_dc: int, qp_ac: int,
                       want_recon: bool = False, bits_buf: np.ndarray | None = None, recon_buf: np.ndarray | None = None):
        """Returns (bodies, stream_bits, recon): bodies[s] = MSB-first body bytes of stream s (ceil(bits/8) bytes)."""
        frames = np.ascontiguousarray(frames, np.uint8)
        n = n_streams * gops_per_stream * gop_len
        if frames.size != n * self.fb:
            raise IcspError(f"expected {n} frames of {self.fb} bytes")
        cap = self.lib.icsp_bits_bound(self.w, self.h, n)
        bits = bits_buf if bits_buf is not None else np.zeros(cap, np.uint8)
        sbits = np.zeros(n_streams, np.uint64)
        soff = np.zeros(n_streams, np.uint64)
        recon = (recon_buf if recon_buf is not None else np.zeros((n, self.fb), np.uint8)) if want_recon else None
        o = _lib.BitsOut(_ptr(bits), bits.size, _ptr(sbits), _ptr(soff), _ptr(recon))
        self._chk(self.lib.icsp_encode_streams(self.h_ctx, _ptr(frames), n_streams, gops_per_stream, gop_len, qp_dc, qp_ac, C.byref(o)),
                  "icsp_encode_streams")
        bodies = [bits[int(soff[s]): int(soff[s]) + (int(sbits[s]) + 7) // 8] for s in range(n_streams)]
        return bodies, sbits, recon

    def entropy_run(self, n_streams: int, gops_per_stream: int, gop_len: int):
        self._chk(self.lib.icsp_entropy_run(self.h_ctx, n_streams, gops_per_stream, gop_len), "icsp_entropy_run")

    def bits_download(self, n_streams: int, cap: int):
        bits = np.zeros(cap, np.uint8)
        sbits = np.zeros(n_streams, np.uint64)
        soff = np.zeros(n_streams, np.uint64)
        o = _lib.BitsOut(_ptr(bits), bits.size, _ptr(sbits), _ptr(soff), None)
        self._chk(self.lib.icsp_bits_download(self.h_ctx, n_streams, C.byref(o)), "icsp_bits_download")
        return [bits[int(soff[s]): int(soff[s]) + (int(sbits[s]) + 7) // 8] for s in range(n_streams)], sbits

    def encode_sequence_bitstream(self, frames: np.ndarray, qp_dc: int, qp_ac: int, intra_period: int, want_recon: bool = False,
                                  want_index: bool = False):
        """One stream -> the complete reference-format file (header + body), entropy coded on the GPU.  Full GOPs and
        the tail GOP are two device calls whose bit strings are concatenated here.  want_index: also return the
        macroblock-row index [n][mbh] (bit offsets from the body start) that decode_sequence_bitstream consumes."""
        frames = np.ascontiguousarray(frames, np.uint8).reshape(-1, self.fb)
        n = frames.shape[0]
        ip = 1 if intra_period == 0 else intra_period
        full, tail = divmod(n, ip)
        segs, recs, idx = [], [], []
        if full:
            b, sb, r = self.encode_streams(frames[: full * ip], 1, full, ip, qp_dc, qp_ac, want_recon)
            segs.append((b[0], int(sb[0]))); recs.append(r)
            if want_index:
                idx.append(self.bits_row_index(full * ip))
        if tail:
            b, sb, r = self.encode_streams(frames[full * ip:], 1, 1, tail, qp_dc, qp_ac, want_recon)
            if want_index:
                idx.append(self.bits_row_index(tail) + np.uint64(segs[0][1] if segs else 0))
            segs.append((b[0], int(sb[0]))); recs.append(r)
        bits = np.concatenate([np.unpackbits(b)[:nb] for b, nb in segs])
        nb = bits.size
        body = bytearray(np.packbits(bits).tobytes()) if nb % 8 else bytearray(np.packbits(bits).tobytes() + b"\x00")
        if nb % 8:
            body[-1] = body[-1] >> (8 - nb % 8)              # reference tail rule (ENC:4895): right-aligned last byte
        data = stream_header(self.w, self.h, qp_dc, qp_ac, intra_period) + bytes(body)
        if want_index:
            return data, (np.concatenate(recs) if want_recon else None), np.concatenate(idx)
        return data, (np.concatenate(recs) if want_recon else None)

    def decode_sequence_bitstream(self, data: bytes, rows: np.ndarray, nframes: int) -> np.ndarray:
        """A reference-format file + its macroblock-row index -> decoded frames, parsed on the GPU (§8 f3).  Header fields
        as the reference decoder reads them (DEC:14-37); frame loop of IcspCodec::decoding (DEC.h:290-313)."""
        if len(data) < 14:
            raise IcspError("bitstream shorter than its 14-byte header")
        qdc, qac = data[9], data[10]
        ip = ((data[12] | (data[13] << 8)) & 0x1F80) >> 7
        if ip < 1 or qdc == 0 or qac == 0:
            raise IcspError("bad header (intraPeriod 0 / QP 0)")
        body = np.frombuffer(data, np.uint8)[14:]
        rows = np.ascontiguousarray(rows, np.uint64).reshape(nframes, self.h // 16)
        full, tail = divmod(nframes, ip)
        outs = []

        def part(f0: int, f1: int, ngops: int, glen: int):
            # only the bytes these frames need go up (icspdec does the same): [start, end) of the body, start 4-byte aligned,
            # row offsets rebased to it
            start = (int(rows[f0][0]) // 8) & ~3
            end = body.size if f1 >= nframes else int(rows[f1][0]) // 8 + 1
            sub = np.ascontiguousarray(body[start:end])
            so, sb = np.zeros(1, np.uint64), np.array([sub.size], np.uint64)
            return self.decode_streams(sub, so, sb, rows[f0:f1] - np.uint64(start * 8), 1, ngops, glen, qdc, qac)
        if full:
            outs.append(part(0, full * ip, full, ip))
        if tail:
            outs.append(part(full * ip, nframes, 1, tail))
        return np.concatenate(outs)

    # ---- decoder -----------------------------------------------------------------------------------
    def decode_gops(self, levels, mpm, ipm, mvd, n_gops: int, gop_len: int, qp_dc: int, qp_ac: int,
                    out: np.ndarray | None = None) -> np.ndarray:
        n = n_gops * gop_len
        levels = np.ascontiguousarray(levels, np.int16)
        mpm = np.ascontiguousarray(mpm, np.uint8)
        ipm = np.ascontiguousarray(ipm, np.uint8)
        mvd = np.ascontiguousarray(mvd, np.int16)
        assert levels.size == n * self.nmb * 384
        din = _lib.DecIn(_ptr(levels), _ptr(mpm), _ptr(ipm), _ptr(mvd))
        if out is None:
            out = np.zeros((n, self.fb), np.uint8)
        self._chk(self.lib.icsp_decode_gops(self.h_ctx, C.byref(din), n_gops, gop_len, qp_dc, qp_ac, _ptr(out)), "icsp_decode_gops")
        return out

    # ---- decoder with the bit reader on the GPU ---------------------------------------
    def bits_row_index(self, n_frames: int) -> np.ndarray:
        """[n_frames][mbh] uint64 bit offset of every macroblock row inside its stream's body, for what
        encode_streams / entropy_run just coded."""
        rows = np.zeros((n_frames, self.h // 16), np.uint64)
        self._chk(self.lib.icsp_bits_row_index(self.h_ctx, n_frames, _ptr(rows)), "icsp_bits_row_index")
        return rows

    def decode_streams(self, bits: np.ndarray, stream_offset, stream_bytes, row_bit_offset, n_streams: int, gops_per_stream: int,
                       gop_len: int, qp_dc: int, qp_ac: int, out: np.ndarray | None = None) -> np.ndarray:
        """bodies (file bytes after the 14-byte header) + macroblock-row index in, decoded I420 frames out."""
        n = n_streams * gops_per_stream * gop_len
        bits = np.ascontiguousarray(bits, np.uint8)
        so = np.ascontiguousarray(stream_offset, np.uint64)
        sb = np.ascontiguousarray(stream_bytes, np.uint64)
        ro = np.ascontiguousarray(row_bit_offset, np.uint64)
        if so.size != n_streams or sb.size != n_streams or ro.size != n * (self.h // 16):
            raise IcspError("decode_streams: table sizes do not match the stream geometry")
        if any(int(o) > bits.size or int(b) > bits.size - int(o) for o, b in zip(so, sb)):      # Python ints: no uint64 wrap-around
            raise IcspError("decode_streams: stream table points past the bits buffer")
        if out is None:
            out = np.zeros((n, self.fb), np.uint8)
        din = _lib.DecBitsIn(_ptr(bits), _ptr(so), _ptr(sb), _ptr(ro))
        self._chk(self.lib.icsp_decode_streams(self.h_ctx, C.byref(din), n_streams, gops_per_stream, gop_len, qp_dc, qp_ac, _ptr(out)),
                  "icsp_decode_streams")
        return out

    def decode_sequence(self, levels, mpm, ipm, mvd, qp_dc: int, qp_ac: int, intra_period: int) -> np.ndarray:
        """Frame loop of IcspCodec::decoding (DEC.h:290-313): intra iff intra_period == 1 or n % intra_period == 0."""
        n = levels.shape[0]
        ip = intra_period
        outs = []
        full = n // ip
        if full:
            c = full * ip
            outs.append(self.decode_gops(levels[:c], mpm[:c], ipm[:c], mvd[:c], full, ip, qp_dc, qp_ac))
        if n - full * ip:
            s = full * ip
            outs.append(self.decode_gops(levels[s:], mpm[s:], ipm[s:], mvd[s:], 1, n - s, qp_dc, qp_ac))
        return np.concatenate(outs, axis=0)

    # ---- shims -------------------------------------------------------------------------------------
    def me_sad(self, cur_y: np.ndarray, ref_y: np.ndarray):
        cur_y = np.ascontiguousarray(cur_y, np.uint8).reshape(-1, self.w * self.h)
        ref_y = np.ascontiguousarray(ref_y, np.uint8).reshape(-1, self.w * self.h)
        n = cur_y.shape[0]
        mv = np.zeros((n, self.nmb, 2), np.int16)
        sad = np.zeros((n, self.nmb), np.int32)
        self._chk(self.lib.icsp_me_sad(self.h_ctx, _ptr(cur_y), _ptr(ref_y), n, _ptr(mv), _ptr(sad)), "icsp_me_sad")
        return mv, sad

    def dct8x8(self, blocks: np.ndarray) -> np.ndarray:
        blocks = np.ascontiguousarray(blocks, np.int32).reshape(-1, 64)
        out = np.zeros(blocks.shape, np.float64)
        self._chk(self.lib.icsp_dct8x8(self.h_ctx, _ptr(blocks), blocks.shape[0], _ptr(out)), "icsp_dct8x8")
        return out

    def idct8x8(self, blocks: np.ndarray, table: int = 0) -> np.ndarray:
        blocks = np.ascontiguousarray(blocks, np.int32).reshape(-1, 64)
        out = np.zeros(blocks.shape, np.float64)
        self._chk(self.lib.icsp_idct8x8(self.h_ctx, _ptr(blocks), blocks.shape[0], table, _ptr(out)), "icsp_idct8x8")
        return out

    # ---- instrumentation ---------------------------------------------------------------------------
    def set_profiling(self, on: bool):
        self._chk(self.lib.icsp_set_profiling(self.h_ctx, int(on)), "icsp_set_profiling")

    def reset_stats(self):
        self._chk(self.lib.icsp_reset_stats(self.h_ctx), "icsp_reset_stats")

    def stats(self) -> dict:
        arr = (_lib.KernelStat * _lib.MAX_KERNELS)()
        n = self.lib.icsp_get_stats(self.h_ctx, arr, _lib.MAX_KERNELS)
        return {arr[i].name.decode(): dict(launches=int(arr[i].launches), total_ms=float(arr[i].total_ms)) for i in range(max(n, 0))}

    def configure(self, n_compute_streams: int = 2, chunk_gops: int = 0):
        self._chk(self.lib.icsp_configure(self.h_ctx, n_compute_streams, chunk_gops), "icsp_configure")

    def launch_count(self) -> int:
        return int(self.lib.icsp_launch_count(self.h_ctx))

    def event_record(self, slot: int):
        self._chk(self.lib.icsp_event_record(self.h_ctx, slot), "icsp_event_record")

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        self._chk(self.lib.icsp_event_elapsed_ms(self.h_ctx, a, b, C.byref(ms)), "icsp_event_elapsed_ms")
        return float(ms.value)
