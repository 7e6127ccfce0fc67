"""`value` of the benchmark workload plus the per-kernel CUDA-event times (serialised run) for the library currently in
icspcodec_b200/ — the A/B driver of tools/variant_sweep.sh.  The synthetic batch is cached in /tmp between variants.
env: ICSP_KT_STREAMS (64), ICSP_KT_FRAMES (300), ICSP_KT_QP (8)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import make_batch, FB
from icspcodec_b200 import IcspCuda
S = int(os.environ.get("ICSP_KT_STREAMS", "64")); F = int(os.environ.get("ICSP_KT_FRAMES", "300")); Q = int(os.environ.get("ICSP_KT_QP", "8"))
cache = f"/tmp/icsp_batch_{S}_{F}.npy"
if os.path.exists(cache):
    batch = np.load(cache, mmap_mode="r")
else:
    batch = make_batch(S, F, 0, 8)
    np.save(cache, batch)
n = batch.shape[0]
ctx = IcspCuda(352, 288, max_frames=n)
ctx.upload(np.ascontiguousarray(batch)); ctx.sync()
for _ in range(3): ctx.run(n // 10, 10, Q, Q)
ctx.sync()
ctx.event_record(0)
for _ in range(5): ctx.run(n // 10, 10, Q, Q)
ctx.event_record(1); ctx.sync()
ms = ctx.event_elapsed_ms(0, 1) / 5
ctx.configure(1, n // 10)
ctx.run(n // 10, 10, Q, Q); ctx.sync()
ctx.set_profiling(True); ctx.reset_stats()
for _ in range(3): ctx.run(n // 10, 10, Q, Q)
st = ctx.stats()
res = ctx.alloc_result(n)
ctx.download(n, res, fields=("recon", "levels")); ctx.sync()
import hashlib
print(json.dumps({"ms_per_step": round(ms, 3), "value": round(n / ms * 1e3), "recon_md5": hashlib.md5(res.recon.tobytes()).hexdigest()[:12],
                  "levels_md5": hashlib.md5(res.levels.tobytes()).hexdigest()[:12],
                  "kernels_avg_ms": {k: round(v["total_ms"] / v["launches"], 4) for k, v in st.items()}}))
