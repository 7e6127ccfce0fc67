"""Generates tests/golden/*.npz|json by RUNNING THE COMPILED, UNMODIFIED REFERENCE (oracle/_ref, built
from /root/reference by oracle/build_ref.sh) on seeded synthetic clips.  Run in the build container
(the reference does not exist on the GPU box):  python tests/golden/make_golden.py

The reference ships no golden vectors of its own (its data/*.yuv and data/output/*.bin are absent
from the mount), so these fixtures are the pinned truth for every parity test:
  * ref_cases.json   md5 of clip / .bin / encoder recon / decoder YUV / full MVs per (clip, qp, ip) case
  * ref_small.npz    the first frames of one case in full (bin bytes, recon, decoder YUV, MVs)
  * ref_dct.npz      random residual / dequantised blocks with the reference's DCT_block / IDCT_block doubles
  * ref_quant.npz    DCT blocks with the reference's Quantization_block / CQuantization_block levels and ACflags
                     (python tests/golden/make_golden.py quant  regenerates only this file)
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from icspcodec_b200 import synth  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [  # (kind, seed, nframes, qdc, qac, ip)
    ("akiyo", 20261017, 12, 8, 8, 10),
    ("akiyo", 20261017, 6, 16, 16, 2),
    ("intra", 1234, 4, 1, 1, 1),
    ("intra", 1234, 4, 8, 8, 1),
    ("intra", 1234, 4, 16, 16, 1),
    ("intra", 1234, 3, 8, 8, 0),
    ("highmotion", 4242, 12, 8, 8, 30),
    ("highmotion", 4242, 6, 1, 16, 3),
    ("flat", 7, 12, 1, 1, 6),
    ("flat", 7, 8, 8, 8, 4),
    ("highmotion", 1000, 10, 8, 8, 10),
]


def md5(b) -> str:
    return hashlib.md5(bytes(b)).hexdigest()


def make_quant():
    rng = np.random.default_rng(2026)
    res = rng.integers(-255, 256, size=(192, 64)).astype(np.int32)
    res[:16] = rng.integers(-6, 7, size=(16, 64))
    d = O.ref_dct(res)                                   # realistic coefficient statistics ...
    d[32:64] = np.rint(d[32:64] * 2) / 2 - 0.5           # ... plus exact .5 ties on both sides of zero
    d[64:72] = rng.integers(-40, 41, size=(8, 64)) * 8 - 0.5
    d[72:80, 1:] = rng.uniform(-0.49, 0.49, size=(8, 63))   # ACflag == 1 blocks
    z = dict(dct=d)
    for qdc, qac in ((1, 1), (8, 8), (16, 16), (8, 3), (100, 255)):
        for chroma in (0, 1):
            lv, ac = O.ref_quant(d, qdc, qac, bool(chroma))
            z[f"lv_{qdc}_{qac}_{chroma}"] = lv
            z[f"ac_{qdc}_{qac}_{chroma}"] = ac.astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "ref_quant.npz"), **z)


def main():
    assert O.have_ref(), "oracle/_ref missing: run oracle/build_ref.sh first"
    if len(sys.argv) > 1 and sys.argv[1] == "quant":
        make_quant()
        return
    make_quant()
    out = []
    for kind, seed, n, qdc, qac, ip in CASES:
        clip = synth.make_clip(kind, n, seed)
        rbin, rrec = O.ref_encode(clip, qdc, qac, ip)
        e = dict(kind=kind, seed=seed, nframes=n, qdc=qdc, qac=qac, ip=ip, clip_md5=md5(clip.tobytes()),
                 bin_md5=md5(rbin), bin_len=len(rbin), recon_md5=md5(rrec.tobytes()))
        if ip > 0:
            e["dec_md5"] = md5(O.ref_decode(rbin, n, qdc, qac, ip).tobytes())
            e["mv_md5"] = md5(O.ref_full_mv(clip, qdc, qac, ip).astype(np.int16).tobytes())
        out.append(e)
        print(e)
    json.dump(out, open(os.path.join(HERE, "ref_cases.json"), "w"), indent=1)

    kind, seed, n, qdc, qac, ip = "highmotion", 4242, 3, 8, 8, 3
    clip = synth.make_clip(kind, n, seed)
    rbin, rrec = O.ref_encode(clip, qdc, qac, ip)
    np.savez_compressed(os.path.join(HERE, "ref_small.npz"), kind=kind, seed=seed, nframes=n, qdc=qdc, qac=qac, ip=ip,
                        bin=np.frombuffer(rbin, np.uint8), recon=rrec, dec=O.ref_decode(rbin, n, qdc, qac, ip),
                        mv=O.ref_full_mv(clip, qdc, qac, ip).astype(np.int16))

    rng = np.random.default_rng(99)
    res = rng.integers(-255, 256, size=(96, 64)).astype(np.int32)
    res[:8] = 0
    res[8:16] = rng.integers(-3, 4, size=(8, 64))
    res[16] = 255
    res[17] = -255
    deq = (rng.integers(-40, 41, size=(96, 64)) * rng.choice([1, 8, 16], size=(96, 1))).astype(np.int32)
    deq[:, 0] = rng.integers(-2100, 2100, size=96)
    deq[:24, 8:] = 0
    np.savez_compressed(os.path.join(HERE, "ref_dct.npz"), res=res, dct=O.ref_dct(res), deq=deq, idct=O.ref_dct(deq, True))


if __name__ == "__main__":
    main()
