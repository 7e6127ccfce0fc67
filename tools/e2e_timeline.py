import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
from bench import make_batch, FB
from icspcodec_b200 import IcspCuda, PinnedArray
batch = make_batch(64, 300, 0, 8); n = batch.shape[0]
ctx = IcspCuda(352, 288, max_frames=n)
pin_in = PinnedArray((n, FB), np.uint8); pin_in.array[:] = batch
pin_bits = PinnedArray((n * (352 * 288 + 32) + 64,), np.uint8)
pin_rec = PinnedArray((n, FB), np.uint8)
for i in range(3):
    print("--- call", i, file=sys.stderr)
    ctx.encode_streams(pin_in.array, 64, 30, 10, 8, 8, want_recon=True, bits_buf=pin_bits.array, recon_buf=pin_rec.array)
