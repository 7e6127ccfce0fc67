"""300-frame golden vectors at the sizes BASELINE.json states (configs 0, 1, 2, the first streams of config 3, and
config 4 = the reference decoder on the streams of configs 0-2), produced by RUNNING THE COMPILED, UNMODIFIED REFERENCE
(oracle/_ref, built from /root/reference by oracle/build_ref.sh).  Run in the build container:

    python tests/golden/make_golden_full.py        ->  tests/golden/ref_cases_full.json   (md5s only, ~2 KB)

Every entry: md5 of the clip, of the reference encoder's .bin and test_yuv.yuv (ENC:4895, 6376-6421) and of the reference
decoder's YUV (DEC.h:290-313; not for --intraPeriod 0, which the decoder cannot open: SURVEY.md H9)."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from icspcodec_b200 import synth  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
N = 300
CASES = [  # (config, kind, seed, nframes, qdc, qac, ip)
    ("configs[0] akiyo-shaped, -q 8 --intraPeriod 10", "akiyo", 20261017, N, 8, 8, 10),
    ("configs[1] all-intra QP 1", "intra", 1234, N, 1, 1, 1),
    ("configs[1] all-intra QP 8", "intra", 1234, N, 8, 8, 1),
    ("configs[1] all-intra QP 16", "intra", 1234, N, 16, 16, 1),
    ("configs[2] high motion, intraPeriod 30", "highmotion", 4242, N, 8, 8, 30),
    ("configs[2] flat/static variant (zero-SAD early breaks), intraPeriod 30", "flat", 7, N, 8, 8, 30),
    ("configs[3] stream 0", "highmotion", 1000, N, 8, 8, 10),
    ("configs[3] stream 1", "highmotion", 1001, N, 8, 8, 10),
    ("configs[3] stream 2", "highmotion", 1002, N, 8, 8, 10),
    ("configs[3] stream 3", "highmotion", 1003, N, 8, 8, 10),
]


def md5(b) -> str:
    return hashlib.md5(bytes(b)).hexdigest()


def main():
    assert O.have_ref(), "oracle/_ref missing: run oracle/build_ref.sh first"
    out = []
    for cfg, kind, seed, n, qdc, qac, ip in CASES:
        clip = synth.make_clip(kind, n, seed)
        rbin, rrec = O.ref_encode(clip, qdc, qac, ip)
        e = dict(config=cfg, kind=kind, seed=seed, nframes=n, qdc=qdc, qac=qac, ip=ip, clip_md5=md5(clip.tobytes()),
                 bin_md5=md5(rbin), bin_len=len(rbin), recon_md5=md5(rrec.tobytes()))
        if ip > 0:
            e["dec_md5"] = md5(O.ref_decode(rbin, n, qdc, qac, ip).tobytes())
        out.append(e)
        print(e, flush=True)
    json.dump(out, open(os.path.join(HERE, "ref_cases_full.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
