import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import make_batch
from icspcodec_b200 import IcspCuda
batch = make_batch(64, 300, 0, 8)
n = batch.shape[0]
ctx = IcspCuda(352, 288, max_frames=n)
ctx.upload(batch); ctx.sync()
for _ in range(3): ctx.run(n // 10, 10, 8, 8)
ctx.sync()
ctx.event_record(0)
for _ in range(5): ctx.run(n // 10, 10, 8, 8)
ctx.event_record(1); ctx.sync()
ms = ctx.event_elapsed_ms(0, 1) / 5
print("streams", os.environ.get("ICSP_STREAMS"), "chunk", os.environ.get("ICSP_CHUNK_GOPS"), "ms/step %.2f value %.0f" % (ms, n / ms * 1e3))
