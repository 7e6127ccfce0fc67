"""Decoder throughput (BASELINE configs[4]): decode the syntax of the benchmark batch; resident and end-to-end."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import make_batch, FB, NMB
from icspcodec_b200 import IcspCuda
streams, frames = 64, 300
batch = make_batch(streams, frames, 0, 8)
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
out_syntax_path = out.copy()
t0 = time.perf_counter()
for _ in range(2): ctx.decode_gops(res.levels, res.mpm, res.ipm, res.mvd, n // 10, 10, 8, 8, out=pin_out.array)
e2e = (time.perf_counter() - t0) / 2
diff = np.abs(out.astype(np.int16) - res.recon.astype(np.int16))
print(json.dumps({"decode_resident_fps": n / ms * 1e3, "ms_per_step": ms, "decode_e2e_fps": n / e2e, "e2e_ms": e2e * 1e3,
                  "h2d_bytes": int(res.levels.nbytes + res.mpm.nbytes + res.ipm.nbytes + res.mvd.nbytes), "d2h_bytes": int(out.nbytes),
                  "max_abs_diff_vs_encoder_recon": int(diff.max())}))

# ---- from bitstreams, parsed on the GPU with the macroblock-row index ----
streams_, gps = 64, frames // 10
bodies, sbits, _ = ctx.encode_streams(batch, streams_, gps, 10, 8, 8)
rows = ctx.bits_row_index(n)
total = sum(((int(b) // 8 + 1 + 3) & ~3) for b in sbits)
pin_bits = PinnedArray((total + 64,), np.uint8)
pin_rows = PinnedArray(rows.shape, np.uint64); pin_rows.array[:] = rows
offs, lens, o = [], [], 0
for s in range(streams_):
    nb = int(sbits[s]); ln = nb // 8 + 1
    body = np.zeros(ln, np.uint8); body[: len(bodies[s])] = bodies[s]
    if nb % 8: body[-1] = body[-1] >> (8 - nb % 8)           # the file's right-aligned last byte (ENC:4895)
    pin_bits.array[o: o + ln] = body; offs.append(o); lens.append(ln); o = (o + ln + 3) & ~3
ctx.configure(1, n // 10)                                      # one chunk on one stream: solo kernel times
ctx.set_profiling(True); ctx.reset_stats()
out2 = ctx.decode_streams(pin_bits.array, offs, lens, pin_rows.array, streams_, gps, 10, 8, 8, out=pin_out.array)
st = ctx.stats(); ctx.set_profiling(False); ctx.configure(4, 0)
same = bool(np.array_equal(out2, out_syntax_path))
t0 = time.perf_counter()
for _ in range(3): ctx.decode_streams(pin_bits.array, offs, lens, pin_rows.array, streams_, gps, 10, 8, 8, out=pin_out.array)
e2e2 = (time.perf_counter() - t0) / 3
host_fps = None
try:
    from icspcodec_b200 import hostlib
    from icspcodec_b200.api import stream_header
    f0 = stream_header(352, 288, 8, 8, 10) + bytes(pin_bits.array[offs[0]: offs[0] + lens[0]])
    t0 = time.perf_counter(); hostlib.parse_stream(f0, frames); host_fps = frames / (time.perf_counter() - t0)
except Exception as e:
    host_fps = str(e)
print(json.dumps({"decode_streams_e2e_fps": n / e2e2, "e2e_ms": e2e2 * 1e3, "h2d_bytes": int(o + rows.nbytes), "d2h_bytes": int(out2.nbytes),
                  "frames_equal_syntax_path_decoder": same, "parse_rows_kernel_ms_total": st.get("parse_rows_kernel", {}).get("total_ms"),
                  "parse_rows_kernel_launches": st.get("parse_rows_kernel", {}).get("launches"), "kernels_ms": {k: round(v["total_ms"], 3) for k, v in st.items()},
                  "host_parser_fps_one_core": host_fps}))
