"""Where one 300-frame stream's decode call spends its time: wall clock of icsp_decode_streams (pinned output), per-kernel
CUDA-event times of the same call with profiling on, and the call without the GPU bit reader (syntax in, icsp_decode_gops)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from icspcodec_b200 import IcspCuda, synth, api
from icspcodec_b200.api import PinnedArray
n, ip = 300, 10
clip = synth.make_clip(os.environ.get("KIND", "akiyo"), n, 1)
with IcspCuda(352, 288, max_frames=n) as ctx:
    bodies, sbits, _ = ctx.encode_streams(clip, 1, n // ip, ip, 8, 8)
    rows = ctx.bits_row_index(n)
    body = np.frombuffer(api.finish_stream(bodies[0], int(sbits[0]), 352, 288, 8, 8, ip)[14:], np.uint8).copy()
    out = PinnedArray((n, ctx.fb), np.uint8)
    res = ctx.encode_gops(clip, n // ip, ip, 8, 8)
    def wall(f, reps=5):
        f(); t0 = time.perf_counter()
        for _ in range(reps): f()
        return (time.perf_counter() - t0) / reps * 1e3
    r = {"decode_streams_ms": round(wall(lambda: ctx.decode_streams(body, [0], [body.size], rows, 1, n // ip, ip, 8, 8, out=out.array)), 3),
         "decode_gops_ms": round(wall(lambda: ctx.decode_gops(res.levels, res.mpm, res.ipm, res.mvd, n // ip, ip, 8, 8, out=out.array)), 3)}
    os.environ["X"] = "1"
    ctx.set_profiling(True); ctx.reset_stats()
    ctx.decode_streams(body, [0], [body.size], rows, 1, n // ip, ip, 8, 8, out=out.array)
    st = ctx.stats()
    r["kernels_total_ms"] = {k: round(v["total_ms"], 3) for k, v in st.items() if v["launches"]}
    r["kernels_sum_ms"] = round(sum(v["total_ms"] for v in st.values()), 3)
    r["body_bytes"] = int(body.size)
    print(json.dumps(r))
