"""One multi-stream step of the benchmark workload with every launch bracketed by events on ITS stream: prints when each kernel
ran (ICSP_KERNEL_TIMELINE=1) so that the overlap between the chunks' pipelines can be read off.  Events add a little serialisation."""
import os, sys
os.environ["ICSP_KERNEL_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import make_batch
from icspcodec_b200 import IcspCuda
batch = make_batch(64, 300, 0, 8)
n = batch.shape[0]
ctx = IcspCuda(352, 288, max_frames=n)
ctx.upload(batch); ctx.sync()
for _ in range(2): ctx.run(n // 10, 10, 8, 8)
ctx.sync()
ctx.set_profiling(True); ctx.reset_stats()
ctx.run(n // 10, 10, 8, 8)
ctx.sync()
ctx.stats()
