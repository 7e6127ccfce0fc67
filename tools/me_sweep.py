"""ME kernel timing sweep over ICSP_ME_SEG (macroblocks per CTA). Prints avg ms per me_sad launch."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    from bench import make_batch
    from icspcodec_b200 import IcspCuda
    batch = make_batch(64, 40, 0, 8)
    n = batch.shape[0]
    ctx = IcspCuda(352, 288, max_frames=n)
    ctx.upload(batch)
    ctx.run(n // 10, 10, 8, 8); ctx.sync()
    ctx.set_profiling(True); ctx.reset_stats()
    for _ in range(3): ctx.run(n // 10, 10, 8, 8)
    st = ctx.stats()
    print(json.dumps({k: round(v["total_ms"] / v["launches"], 4) for k, v in st.items()}))
else:
    for seg in sys.argv[1:] or ["22", "11", "8", "6", "4"]:
        env = dict(os.environ, ICSP_ME_SEG=seg)
        out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print("seg", seg, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:])
