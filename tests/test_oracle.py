"""CPU tests: the oracle (oracle/icsp_oracle.c) against the fixtures that were produced by running the
compiled, unmodified reference (tests/golden/make_golden.py), and — when oracle/_ref is present —
against the reference binaries themselves."""
import hashlib
import json
import os

import numpy as np
import pytest

from icspcodec_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = json.load(open(os.path.join(GOLD, "ref_cases.json")))
W, H = 352, 288


def md5(b):
    return hashlib.md5(bytes(b)).hexdigest()


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c['kind']}-q{c['qdc']}_{c['qac']}-ip{c['ip']}")
def test_oracle_matches_reference_md5(oracle, case):
    clip = synth.make_clip(case["kind"], case["nframes"], case["seed"])
    assert md5(clip.tobytes()) == case["clip_md5"], "synthetic clip generator is not reproducible"
    s = oracle.encode(clip, W, H, case["qdc"], case["qac"], case["ip"])
    bs = oracle.write_bitstream(s, W, H, case["qdc"], case["qac"], case["ip"])
    assert len(bs) == case["bin_len"]
    assert md5(bs) == case["bin_md5"]                      # every syntax element, bit-exact
    assert md5(s.recon.tobytes()) == case["recon_md5"]      # reconstructed YUV
    if case["ip"] > 0:
        assert md5(s.mv.tobytes()) == case["mv_md5"]        # full motion vectors (Reconstructedmv)
        ps, hdr = oracle.parse_bitstream(bs, case["nframes"])
        assert hdr == dict(w=W, h=H, qdc=case["qdc"], qac=case["qac"], ip=case["ip"])
        dec = oracle.decode(ps, W, H, case["qdc"], case["qac"], case["ip"])
        assert md5(dec.tobytes()) == case["dec_md5"]        # reference decoder YUV (double cosine table)


def test_oracle_small_fixture_full_compare(oracle):
    z = np.load(os.path.join(GOLD, "ref_small.npz"))
    n, qdc, qac, ip = int(z["nframes"]), int(z["qdc"]), int(z["qac"]), int(z["ip"])
    clip = synth.make_clip(str(z["kind"]), n, int(z["seed"]))
    s = oracle.encode(clip, W, H, qdc, qac, ip)
    assert np.array_equal(s.recon, z["recon"])
    assert np.array_equal(s.mv, z["mv"])
    assert oracle.write_bitstream(s, W, H, qdc, qac, ip) == z["bin"].tobytes()
    ps, _ = oracle.parse_bitstream(z["bin"].tobytes(), n)
    assert np.array_equal(ps.levels, s.levels) and np.array_equal(ps.mvd, s.mvd)
    assert np.array_equal(oracle.decode(ps, W, H, qdc, qac, ip), z["dec"])


def test_oracle_dct_bits(oracle):
    z = np.load(os.path.join(GOLD, "ref_dct.npz"))
    assert np.array_equal(oracle.dct8x8(z["res"]), z["dct"])          # 0 ulp, stricter than the 1e-9 bar
    assert np.array_equal(oracle.idct8x8(z["deq"], 0), z["idct"])
    # decoder table differs from the encoder's (DEC.h:19 vs ENC.h:190)
    assert not np.array_equal(oracle.idct8x8(z["deq"], 1), z["idct"])


def test_flat_clip_exercises_early_break(oracle):
    clip = synth.make_clip("flat", 3, 7)
    y0 = clip[0][: W * H]
    _, sad, evals = oracle.me(clip[1][: W * H], y0, W, H)
    assert evals < 396 * 64 // 2            # zero-SAD early breaks fired
    assert (sad == 0).sum() > 100


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(GOLD), "..", "oracle", "_ref", "ICSPCodec_O2")),
                    reason="oracle/_ref not built")
def test_oracle_vs_live_reference(oracle):
    clip = synth.make_clip("highmotion", 5, 31337)
    s = oracle.encode(clip, W, H, 8, 16, 5)
    rbin, rrec = oracle.ref_encode(clip, 8, 16, 5)
    assert oracle.write_bitstream(s, W, H, 8, 16, 5) == rbin
    assert np.array_equal(rrec, s.recon)
    ps, _ = oracle.parse_bitstream(rbin, 5)
    assert np.array_equal(oracle.decode(ps, W, H, 8, 16, 5), oracle.ref_decode(rbin, 5, 8, 16, 5))
    rng = np.random.default_rng(5)
    blocks = rng.integers(-255, 256, size=(32, 64)).astype(np.int32)
    assert np.array_equal(oracle.dct8x8(blocks), oracle.ref_dct(blocks))
    assert np.array_equal(oracle.idct8x8(blocks * 8, 0), oracle.ref_dct(blocks * 8, True))


# ---- BASELINE.json sizes: 300-frame clips (tests/golden/make_golden_full.py) -------------------------------------------
FULL = json.load(open(os.path.join(GOLD, "ref_cases_full.json")))


@pytest.mark.parametrize("case", [FULL[0], FULL[5]], ids=lambda c: c["config"].split(",")[0].replace(" ", "_"))
def test_oracle_matches_reference_at_baseline_size(oracle, case):
    """configs[0] (akiyo-shaped, IP 10) and the flat/static configs[2] variant (thousands of zero-SAD early breaks, carried
    spiral state over 290 inter frames) at the full 300 frames: .bin, reconstruction and decoder YUV md5 of the reference."""
    clip = synth.make_clip(case["kind"], case["nframes"], case["seed"])
    assert md5(clip.tobytes()) == case["clip_md5"]
    s = oracle.encode(clip, W, H, case["qdc"], case["qac"], case["ip"])
    bs = oracle.write_bitstream(s, W, H, case["qdc"], case["qac"], case["ip"])
    assert len(bs) == case["bin_len"] and md5(bs) == case["bin_md5"]
    assert md5(s.recon.tobytes()) == case["recon_md5"]
    ps, _ = oracle.parse_bitstream(bs, case["nframes"])
    assert md5(oracle.decode(ps, W, H, case["qdc"], case["qac"], case["ip"]).tobytes()) == case["dec_md5"]
