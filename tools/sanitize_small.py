"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icspcodec_b200 import IcspCuda, synth
clips = [synth.make_clip("highmotion", 4, 4242), synth.make_clip("flat", 4, 7)]
frames = np.concatenate(clips, axis=0)
with IcspCuda(352, 288, max_frames=8) as ctx:
    res = ctx.encode_gops(frames, 2, 4, 8, 8)
    bodies, sbits, rec = ctx.encode_streams(frames, 2, 1, 4, 8, 8, want_recon=True)
    out = ctx.decode_gops(res.levels, res.mpm, res.ipm, res.mvd, 2, 4, 8, 8)
    sse = ctx.enc_sse(8)                                          # plane_sse_kernel
    rows = ctx.bits_row_index(8)                                  # row_index_kernel
    blob, offs, lens = bytearray(), [], []
    for s in range(2):                                            # file bodies (reference tail rule), 4-byte aligned
        nb = int(sbits[s]); b = bytearray(bytes(bodies[s]))
        if nb % 8 == 0: b += b"\x00"
        else: b[-1] = b[-1] >> (8 - nb % 8)
        while len(blob) % 4: blob += b"\x00"
        offs.append(len(blob)); lens.append(len(b)); blob += b
    out2 = ctx.decode_streams(np.frombuffer(bytes(blob), np.uint8), offs, lens, rows, 2, 1, 4, 8, 8)   # parse_rows_kernel
    assert np.array_equal(out, out2)
    # round-2 entry points: in-pipeline DCT tap, frame-level shims, stand-alone quantiser, graph replay (second identical call)
    tap = ctx.encode_gops(frames, 2, 4, 8, 8, with_dct=True)
    assert np.array_equal(tap.recon, res.recon) and tap.dct is not None
    ri = ctx.intra_frame(frames[:2], 8, 8, with_dct=True)
    rp = ctx.inter_frame(frames[1:3], ri.recon, 8, 8)
    ql = ctx.quant(tap.dct[0].reshape(-1, 64)[:96], 8, 8)
    assert np.array_equal(ri.recon[0], res.recon[0]) and rp.recon.shape == ri.recon.shape and ql is not None
    print("ok", int(sbits.sum()), out.shape, int(sse.sum()))
with IcspCuda(64, 48, max_frames=4) as ctx:
    f = np.random.default_rng(0).integers(0, 255, size=(4, 64 * 48 * 3 // 2)).astype(np.uint8)
    r = ctx.encode_gops(f, 1, 4, 4, 4)
    print("ok small", r.recon.shape)
