"""Small driver for ncu captures: one encode pass of a reduced batch (defaults: 24 streams x 100 frames, IP 10, q 8)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_batch, FB  # noqa: E402
from icspcodec_b200 import IcspCuda  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=24)
ap.add_argument("--frames", type=int, default=100)
ap.add_argument("--passes", type=int, default=1)
ap.add_argument("--decode", action="store_true")
ap.add_argument("--entropy", action="store_true")
ap.add_argument("--bits-decode", action="store_true", help="encode_streams + row index + decode_streams (GPU bit reader)")
a = ap.parse_args()
batch = make_batch(a.streams, a.frames, 0, 8)
n = batch.shape[0]
ctx = IcspCuda(352, 288, max_frames=n)
ctx.upload(batch)
for _ in range(a.passes):
    ctx.run(n // 10, 10, 8, 8)
ctx.sync()
if a.entropy:
    ctx.entropy_run(a.streams, a.frames // 10, 10)
    ctx.sync()
if a.decode:
    res = ctx.alloc_result(n)
    ctx.download(n, res)
    ctx.sync()
    out = ctx.decode_gops(res.levels, res.mpm, res.ipm, res.mvd, n // 10, 10, 8, 8)
if a.bits_decode:
    gps = a.frames // 10
    bodies, sbits, _ = ctx.encode_streams(batch, a.streams, gps, 10, 8, 8)
    rows = ctx.bits_row_index(n)
    blob, offs, lens = bytearray(), [], []
    for s in range(a.streams):
        nb = int(sbits[s]); b = bytearray(bytes(bodies[s]))
        if nb % 8 == 0: b += b"\x00"
        else: b[-1] = b[-1] >> (8 - nb % 8)
        while len(blob) % 4: blob += b"\x00"
        offs.append(len(blob)); lens.append(len(b)); blob += b
    ctx.decode_streams(np.frombuffer(bytes(blob), np.uint8), offs, lens, rows, a.streams, gps, 10, 8, 8)
    ctx.enc_sse(n)
print("done", n, "frames", ctx.launch_count(), "launches")
