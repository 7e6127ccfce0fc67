"""Decoder throughput (BASELINE configs[4]): decode the syntax of the benchmark batch; resident and end-to-end."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import make_batch, FB, NMB
from icspcodec_b200 import IcspCuda
streams, frames = 64, 300
batch = make_batch(streams, frames, 0)
n = batch.shape[0]
ctx = IcspCuda(352, 288, max_frames=n)
res = ctx.alloc_result(n, pinned=True)
ctx.encode_gops(batch, n // 10, 10, 8, 8, out=res)
# resident: syntax already on the device from the encode
ctx.lib.icsp_dec_run(ctx.h_ctx, n // 10, 10, 8, 8); ctx.sync()
ctx.event_record(0)
for _ in range(5): ctx.lib.icsp_dec_run(ctx.h_ctx, n // 10, 10, 8, 8)
ctx.event_record(1); ctx.sync()
ms = ctx.event_elapsed_ms(0, 1) / 5
from icspcodec_b200 import PinnedArray
pin_out = PinnedArray((n, FB), np.uint8)
out = ctx.decode_gops(res.levels, res.mpm, res.ipm, res.mvd, n // 10, 10, 8, 8, out=pin_out.array)
t0 = time.perf_counter()
for _ in range(2): ctx.decode_gops(res.levels, res.mpm, res.ipm, res.mvd, n // 10, 10, 8, 8, out=pin_out.array)
e2e = (time.perf_counter() - t0) / 2
diff = np.abs(out.astype(np.int16) - res.recon.astype(np.int16))
print(json.dumps({"decode_resident_fps": n / ms * 1e3, "ms_per_step": ms, "decode_e2e_fps": n / e2e, "e2e_ms": e2e * 1e3,
                  "h2d_bytes": int(res.levels.nbytes + res.mpm.nbytes + res.ipm.nbytes + res.mvd.nbytes), "d2h_bytes": int(out.nbytes),
                  "max_abs_diff_vs_encoder_recon": int(diff.max())}))
