"""CPU tests of the PRODUCT host code (icspcodec_b200/host/bitstream.cpp, shared by icspenc/icspdec): fed with
syntax arrays, it must reproduce the reference's .bin byte for byte (golden md5s come from the compiled
reference), and its reader must parse those streams like the reference decoder's bit reader."""
import hashlib
import json
import os

import numpy as np
import pytest

from icspcodec_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = json.load(open(os.path.join(GOLD, "ref_cases.json")))
W, H = 352, 288


@pytest.fixture(scope="module")
def host():
    from icspcodec_b200 import build, hostlib
    build.build()
    return hostlib


@pytest.mark.parametrize("case", CASES[:2] + CASES[2:6:2] + CASES[6:10], ids=lambda c: f"{c['kind']}-q{c['qdc']}_{c['qac']}-ip{c['ip']}")
def test_writer_reproduces_reference_bin(host, oracle, case):
    clip = synth.make_clip(case["kind"], case["nframes"], case["seed"])
    s = oracle.encode(clip, W, H, case["qdc"], case["qac"], case["ip"])      # syntax source (checker side)
    for threads in (1, 3):
        bs = host.write_stream(s.levels, s.acflag, s.mpm, s.ipm, s.mvd, W, H, case["qdc"], case["qac"], case["ip"], threads)
        assert len(bs) == case["bin_len"]
        assert hashlib.md5(bs).hexdigest() == case["bin_md5"]


def test_reader_matches_oracle_reader(host, oracle):
    z = np.load(os.path.join(GOLD, "ref_small.npz"))
    data = z["bin"].tobytes()
    n = int(z["nframes"])
    got, hdr = host.parse_stream(data, n)
    ps, ohdr = oracle.parse_bitstream(data, n)
    assert hdr == ohdr
    for k in ("levels", "acflag", "mpm", "ipm", "mvd"):
        assert np.array_equal(got[k], getattr(ps, k)), k


def test_vlc_all_categories_roundtrip(host):
    """Every VLC category incl. the >= 2048 escape (11 low bits only), and the tail-bit quirk: values survive a
    write -> parse round trip whenever the stream does not end in a partial byte that the MSB-first reader mis-reads."""
    nmb = 1
    vals = [0, 1, -1, 2, 3, -3, 4, 7, -8, 15, 16, -31, 32, 63, -64, 127, 128, -255, 256, 511, -512, 1023, 1024, -2047, 2048, -3000, 4095]
    levels = np.zeros((2, nmb, 6, 64), np.int16)
    levels[0, 0, 0, : len(vals)] = vals
    levels[1, 0, 5, 1: len(vals) + 1] = vals[::-1]
    acflag = (np.abs(levels[..., 1:]).sum(-1) == 0).astype(np.uint8)
    mpm = np.zeros((2, nmb, 4), np.uint8); ipm = np.zeros((2, nmb, 4), np.uint8)
    mpm[0, 0, 1] = 1; ipm[0, 0, 2] = 1
    mvd = np.zeros((2, nmb, 2), np.int16); mvd[1, 0] = (-16, 9)
    bs = host.write_stream(levels, acflag, mpm, ipm, mvd, 16, 16, 8, 8, 2, 1)
    got, hdr = host.parse_stream(bs + b"\x00" * 4, 2)     # padding keeps the right-aligned tail from mattering? no: see below
    assert hdr == dict(w=16, h=16, qdc=8, qac=8, ip=2)
    # all but possibly the very last block parse back exactly (H7: the reference's tail byte is right-aligned)
    assert np.array_equal(got["levels"][0], levels[0])
    assert np.array_equal(got["mvd"], mvd)
    assert np.array_equal(got["mpm"], mpm) and np.array_equal(got["ipm"], ipm)
    assert np.array_equal(got["levels"][1, 0, :5], levels[1, 0, :5])


def test_empty_and_bad_streams(host):
    with pytest.raises(ValueError):
        host.parse_stream(b"\x00ICSP", 1)
    # intraPeriod 0 in the header is rejected (the reference decoder divides by zero, DEC:201)
    hdr0 = bytes([0, 73, 67, 83, 80, 32, 1, 96, 1, 8, 8, 0, 0, 0]) + b"\x00" * 64
    with pytest.raises(ValueError):
        host.parse_stream(hdr0, 1)


def test_row_index_roundtrip(oracle):
    """Macroblock-row index (SURVEY §8 f3 side-car): the host writer's offsets let every macroblock row be parsed on its own
    (parse_stream_indexed = the chains the GPU bit reader follows) with the same result as the serial parse; offsets ascend,
    start at 0, and the stream bytes equal the un-indexed writer's."""
    from icspcodec_b200 import hostlib, synth
    for kind, n, qdc, qac, ip in (("akiyo", 5, 8, 8, 3), ("highmotion", 4, 1, 16, 2), ("intra", 3, 8, 8, 1), ("flat", 4, 16, 16, 4)):
        clip = synth.make_clip(kind, n, 5)
        s = oracle.encode(clip, 352, 288, qdc, qac, ip)
        data, rows = hostlib.write_stream_indexed(s.levels, s.acflag, s.mpm, s.ipm, s.mvd, 352, 288, qdc, qac, ip, threads=3)
        assert data == oracle.write_bitstream(s, 352, 288, qdc, qac, ip)
        flat = rows.reshape(-1).astype(np.int64)
        assert rows.shape == (n, 18) and flat[0] == 0 and np.all(np.diff(flat) > 0) and flat[-1] < (len(data) - 14) * 8
        serial, _ = hostlib.parse_stream(data, n)
        indexed, _ = hostlib.parse_stream_indexed(data, n, rows)
        for k in serial:
            assert np.array_equal(serial[k], indexed[k]), k
        # a shifted index must change the result (the index is really used)
        bad = rows.copy(); bad[1:, :] += np.uint64(1)
        wrong, _ = hostlib.parse_stream_indexed(data, n, bad)
        assert not np.array_equal(wrong["levels"], serial["levels"])
