"""Long seeded fuzz of the GPU encoder / bitstream / decoder / GPU bit reader against the oracle (run on a GPU box).
usage: python tools/fuzz_big.py [seconds] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from icspcodec_b200 import IcspCuda
from oracle import oracle_py as O
from test_gpu_parity import _fuzz_clip

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
O.build()
geoms = [(16, 16), (32, 16), (48, 32), (64, 64), (96, 64), (80, 48), (176, 144), (368, 48), (352, 288)]
ctxs = {}
t0 = time.time(); n_ok = 0; bad = []
while time.time() - t0 < budget:
    w, h = geoms[int(rng.integers(0, len(geoms)))]
    if (w, h) not in ctxs: ctxs[(w, h)] = IcspCuda(w, h, max_frames=16)
    ctx = ctxs[(w, h)]
    n = int(rng.integers(1, 10)); ip = int(rng.choice([0, 1, 2, 3, 4, 5, 8, 10]))
    qdc, qac = int(rng.integers(1, 41)), int(rng.integers(1, 41))
    clip = _fuzz_clip(rng, n, w, h)
    tag = f"{w}x{h} n={n} ip={ip} q={qdc}/{qac}"
    try:
        res = ctx.encode_sequence(clip, qdc, qac, ip)
        s = O.encode(clip, w, h, qdc, qac, ip)
        for name in ("levels", "acflag", "mpm", "ipm", "mvd", "recon"):
            assert np.array_equal(np.asarray(getattr(res, name)).reshape(-1), np.asarray(getattr(s, name)).reshape(-1)), name
        data, _, rows = ctx.encode_sequence_bitstream(clip, qdc, qac, ip, want_index=True)
        assert data == O.write_bitstream(s, w, h, qdc, qac, ip), "bitstream"
        if ip > 0:
            want = O.decode(s, w, h, qdc, qac, ip)
            assert np.array_equal(ctx.decode_sequence(res.levels, res.mpm, res.ipm, res.mvd, qdc, qac, ip), want), "decode"
            # the reference decoder parses the FILE (right-aligned last byte): compare the GPU bit reader with the host parser on it
            from icspcodec_b200 import hostlib
            parsed, _ = hostlib.parse_stream(data, n)
            full, tail = divmod(n, ip)
            outs = []
            if full: outs.append(ctx.decode_gops(parsed["levels"][:full * ip], parsed["mpm"][:full * ip], parsed["ipm"][:full * ip], parsed["mvd"][:full * ip], full, ip, qdc, qac))
            if tail: outs.append(ctx.decode_gops(parsed["levels"][full * ip:], parsed["mpm"][full * ip:], parsed["ipm"][full * ip:], parsed["mvd"][full * ip:], 1, tail, qdc, qac))
            assert np.array_equal(ctx.decode_sequence_bitstream(data, rows, n), np.concatenate(outs)), "gpu bit reader"
        n_ok += 1
    except AssertionError as e:
        bad.append((tag, str(e)))
        print("MISMATCH", tag, e, flush=True)
print(f"fuzz: {n_ok} clips ok, {len(bad)} mismatches in {time.time() - t0:.0f} s (seed {seed})")
sys.exit(1 if bad else 0)
