"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path through the C ABI versus the oracle
(oracle/icsp_oracle.c, pinned to the compiled reference) and versus the golden fixtures produced by the
reference itself.  Everything integer is compared bit-exactly; DCT/IDCT doubles are compared with rel 1e-9
(the north_star tolerance) AND, stricter, for 0-ulp equality."""
import hashlib
import json
import os

import numpy as np
import pytest

from icspcodec_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = json.load(open(os.path.join(GOLD, "ref_cases.json")))
W, H = 352, 288


def md5(b):
    return hashlib.md5(bytes(b)).hexdigest()


@pytest.fixture(scope="module")
def gpu():
    from icspcodec_b200 import IcspCuda
    ctx = IcspCuda(W, H, max_frames=64)
    yield ctx
    ctx.close()


def assert_syntax_equal(res, s, what=""):
    for name in ("levels", "acflag", "mpm", "ipm", "mvd", "mv", "minsad", "recon"):
        a, b = getattr(res, name), getattr(s, name)
        if not np.array_equal(a, b):
            idx = np.argwhere(a != b)
            raise AssertionError(f"{what}{name}: {len(idx)} mismatches, first at {idx[0]} gpu={a[tuple(idx[0])]} oracle={b[tuple(idx[0])]}")


def test_dct_idct_shims(gpu, oracle):
    z = np.load(os.path.join(GOLD, "ref_dct.npz"))
    d = gpu.dct8x8(z["res"])
    np.testing.assert_allclose(d, z["dct"], rtol=1e-9, atol=0)      # the stated tolerance
    assert np.array_equal(d, z["dct"])                              # and in fact bit-identical
    assert np.array_equal(gpu.idct8x8(z["deq"], 0), z["idct"])
    assert np.array_equal(gpu.idct8x8(z["deq"], 1), oracle.idct8x8(z["deq"], 1))
    rng = np.random.default_rng(1)
    blocks = rng.integers(-255, 256, size=(4096, 64)).astype(np.int32)
    assert np.array_equal(gpu.dct8x8(blocks), oracle.dct8x8(blocks))
    deq = (rng.integers(-300, 301, size=(4096, 64))).astype(np.int32) * 8
    deq[:, 0] = rng.integers(-4100, 4100, size=4096)
    assert np.array_equal(gpu.idct8x8(deq, 0), oracle.idct8x8(deq, 0))
    assert np.array_equal(gpu.idct8x8(deq, 1), oracle.idct8x8(deq, 1))


@pytest.mark.parametrize("kind,seed", [("highmotion", 4242), ("flat", 7), ("akiyo", 20261017)])
def test_me_shim(gpu, oracle, kind, seed):
    clip = synth.make_clip(kind, 5, seed)
    ys = clip[:, : W * H]
    mv, sad = gpu.me_sad(ys[1:], ys[:-1])
    for i in range(4):
        omv, osad, _ = oracle.me(ys[i + 1], ys[i], W, H)
        assert np.array_equal(mv[i], omv), f"{kind} pair {i}: mv mismatch at MBs {np.argwhere((mv[i] != omv).any(1))[:8].ravel()}"
        assert np.array_equal(sad[i], osad)


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c['kind']}-q{c['qdc']}_{c['qac']}-ip{c['ip']}")
def test_encode_matches_oracle_and_reference(gpu, oracle, case):
    clip = synth.make_clip(case["kind"], case["nframes"], case["seed"])
    qdc, qac, ip = case["qdc"], case["qac"], case["ip"]
    res = gpu.encode_sequence(clip, qdc, qac, ip)
    s = oracle.encode(clip, W, H, qdc, qac, ip)
    assert_syntax_equal(res, s)
    # golden fixtures produced by the compiled reference: reconstruction, MVs and the bitstream that the host writer
    # derives from exactly these syntax arrays
    assert md5(res.recon.tobytes()) == case["recon_md5"]
    if ip > 0:
        assert md5(res.mv.tobytes()) == case["mv_md5"]
    gs = oracle.Syntax(res.levels, res.acflag, res.mpm, res.ipm, res.mvd)
    assert md5(oracle.write_bitstream(gs, W, H, qdc, qac, ip)) == case["bin_md5"]


@pytest.mark.parametrize("case", [c for c in CASES if c["ip"] > 0], ids=lambda c: f"{c['kind']}-q{c['qdc']}_{c['qac']}-ip{c['ip']}")
def test_decode_matches_reference_decoder(gpu, oracle, case):
    clip = synth.make_clip(case["kind"], case["nframes"], case["seed"])
    qdc, qac, ip = case["qdc"], case["qac"], case["ip"]
    s = oracle.encode(clip, W, H, qdc, qac, ip)
    bs = oracle.write_bitstream(s, W, H, qdc, qac, ip)
    ps, _ = oracle.parse_bitstream(bs, case["nframes"])           # MSB-first reader, like DEC:68-74
    out = gpu.decode_sequence(ps.levels, ps.mpm, ps.ipm, ps.mvd, qdc, qac, ip)
    assert np.array_equal(out, oracle.decode(ps, W, H, qdc, qac, ip))
    assert md5(out.tobytes()) == case["dec_md5"]                    # bit-exact YUV vs the reference decoder


def test_gop_batch_equals_sequential(gpu, oracle):
    """GOPs are independent: a batch of 6 GOPs x 5 frames in one call == the oracle run GOP by GOP."""
    clips = [synth.make_clip("highmotion", 5, 1000 + i) for i in range(4)] + [synth.make_clip("flat", 5, 7), synth.make_clip("akiyo", 5, 3)]
    frames = np.concatenate(clips, axis=0)
    res = gpu.encode_gops(frames, 6, 5, 8, 8)
    for g, c in enumerate(clips):
        s = oracle.encode(c, W, H, 8, 8, 5)
        sub = type(res)(**{k: (None if getattr(res, k) is None else getattr(res, k)[g * 5:(g + 1) * 5]) for k in res.__dataclass_fields__})
        assert_syntax_equal(sub, s, what=f"gop {g}: ")


@pytest.mark.parametrize("w,h", [(16, 16), (64, 48), (176, 144), (368, 48), (720, 480), (1280, 720), (1920, 1088)])   # 368 = 23 MBs: segments of 12 + 11
def test_other_geometries(oracle, w, h):
    from icspcodec_b200 import IcspCuda
    rng = np.random.default_rng(w * 1000 + h)
    n = 4
    base = rng.integers(0, 256, size=(h + 16, w + 16)).astype(np.float64)
    base = (base + np.roll(base, 1, 0) + np.roll(base, 1, 1)) / 3
    frames = []
    for i in range(n):
        y = base[i:i + h, 2 * i:2 * i + w] + rng.normal(0, 1.5, size=(h, w))
        cb = 128 + 20 * np.sin(np.arange(w // 2) / 7.0)[None, :] + np.zeros((h // 2, 1)) + i
        cr = 120 + 15 * np.cos(np.arange(h // 2) / 5.0)[:, None] + np.zeros((1, w // 2)) - i
        frames.append(np.concatenate([np.clip(np.rint(p), 0, 255).astype(np.uint8).ravel() for p in (y, cb, cr)]))
    frames = np.stack(frames)
    with IcspCuda(w, h, max_frames=8) as ctx:
        res = ctx.encode_sequence(frames, 8, 4, 4)
        s = oracle.encode(frames, w, h, 8, 4, 4)
        assert_syntax_equal(res, s, what=f"{w}x{h}: ")
        out = ctx.decode_sequence(res.levels, res.mpm, res.ipm, res.mvd, 8, 4, 4)
        assert np.array_equal(out, oracle.decode(s, w, h, 8, 4, 4))
        data, _ = ctx.encode_sequence_bitstream(frames, 8, 4, 4)            # GPU entropy path at this geometry
        assert data == oracle.write_bitstream(s, w, h, 8, 4, 4)


def test_roundtrip_properties_full_size(gpu, oracle):
    """Size-independent checks on a longer clip (no oracle needed at this size): deterministic re-encode, and
    decoding our own syntax with the ENCODER table would equal the encoder reconstruction; with the decoder table
    the difference stays tiny (|diff| <= 2) — H2 of SURVEY.md."""
    clip = synth.make_clip("highmotion", 60, 99)
    a = gpu.encode_sequence(clip, 8, 8, 10)
    b = gpu.encode_sequence(clip, 8, 8, 10)
    for name in a.__dataclass_fields__:
        assert np.array_equal(getattr(a, name), getattr(b, name))
    dec = gpu.decode_sequence(a.levels, a.mpm, a.ipm, a.mvd, 8, 8, 10)
    diff = np.abs(dec.astype(np.int16) - a.recon.astype(np.int16))
    assert diff.max() <= 3 and (diff != 0).mean() < 0.02
    # PSNR sanity versus the source
    mse = np.mean((a.recon[:, : W * H].astype(np.float64) - clip[:, : W * H]) ** 2)
    assert 10 * np.log10(255 ** 2 / mse) > 30


def test_error_paths(gpu):
    from icspcodec_b200 import IcspError
    clip = synth.make_clip("flat", 2, 7)
    with pytest.raises(IcspError):
        gpu.encode_gops(clip, 1, 2, 0, 8)          # qp must be positive
    with pytest.raises(IcspError):
        gpu.encode_gops(np.zeros((200, gpu.fb), np.uint8), 20, 10, 8, 8)   # over capacity


# ---- f1: entropy coding + bit packing on the GPU ---------------------------------------------------------------
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c['kind']}-q{c['qdc']}_{c['qac']}-ip{c['ip']}")
def test_gpu_bitstream_matches_reference(gpu, case):
    """Bitstream produced entirely on the GPU (icsp_encode_streams) == the reference encoder's .bin, byte for byte."""
    clip = synth.make_clip(case["kind"], case["nframes"], case["seed"])
    data, recon = gpu.encode_sequence_bitstream(clip, case["qdc"], case["qac"], case["ip"], want_recon=True)
    assert len(data) == case["bin_len"]
    assert md5(data) == case["bin_md5"]
    assert md5(recon.tobytes()) == case["recon_md5"]


def test_gpu_bitstream_stream_batch(gpu, oracle):
    """A batch of independent streams in one call: every stream's body equals the oracle writer's body of that stream."""
    clips = [synth.make_clip("highmotion", 10, 500 + i) for i in range(3)] + [synth.make_clip("flat", 10, 7), synth.make_clip("akiyo", 10, 11)]
    frames = np.concatenate(clips, axis=0)
    bodies, sbits, _ = gpu.encode_streams(frames, 5, 2, 5, 8, 8)
    for s, c in enumerate(clips):
        o = oracle.encode(c, W, H, 8, 8, 5)
        ref = oracle.write_bitstream(o, W, H, 8, 8, 5)[14:]
        nb = int(sbits[s])
        assert len(ref) == nb // 8 + 1
        got = bytearray(bodies[s].tobytes()) + (b"" if nb % 8 else b"\x00")
        if nb % 8:
            got[-1] >>= 8 - nb % 8
        assert bytes(got) == ref, f"stream {s}"
    # resident variant: run + entropy_run + bits_download gives the same bodies
    gpu.upload(frames)
    gpu.run(10, 5, 8, 8)
    gpu.entropy_run(5, 2, 5)
    bodies2, sbits2 = gpu.bits_download(5, frames.shape[0] * (W * H + 32) + 64)
    assert np.array_equal(sbits, sbits2)
    for a, b in zip(bodies, bodies2):
        assert np.array_equal(a, b)


def test_multi_chunk_pipelines_match(oracle):
    """Force many small GOP chunks / several compute streams: the chunked, pipelined code paths (H2D | kernels | D2H,
    per-chunk entropy regions and tables) must give exactly the results of the single-chunk run."""
    from icspcodec_b200 import IcspCuda
    clips = [synth.make_clip("highmotion", 6, 900 + i) for i in range(3)] + [synth.make_clip("flat", 6, 7), synth.make_clip("akiyo", 6, 5)]
    frames = np.concatenate(clips, axis=0)
    with IcspCuda(W, H, max_frames=32) as ctx:
        ctx.configure(1, 0)
        ref = ctx.encode_gops(frames, 10, 3, 8, 8)
        rb, rbits, rrec = ctx.encode_streams(frames, 5, 2, 3, 8, 8, want_recon=True)
        for nstreams, cg in ((3, 1), (4, 2), (2, 3)):
            ctx.configure(nstreams, cg)
            got = ctx.encode_gops(frames, 10, 3, 8, 8)
            assert_syntax_equal(got, ref, what=f"streams={nstreams} chunk={cg}: ")
            b, bits, rec = ctx.encode_streams(frames, 5, 2, 3, 8, 8, want_recon=True)
            assert np.array_equal(bits, rbits) and np.array_equal(rec, rrec)
            for x, y in zip(b, rb):
                assert np.array_equal(x, y)
            dec = ctx.decode_gops(ref.levels, ref.mpm, ref.ipm, ref.mvd, 10, 3, 8, 8)
            s0 = oracle.encode(clips[0], W, H, 8, 8, 3)
            assert np.array_equal(dec[:6], oracle.decode(s0, W, H, 8, 8, 3))


@pytest.mark.parametrize("qdc,qac", [(3, 7), (100, 255), (1, 2), (16, 1)])
def test_unusual_quantisers(gpu, oracle, qdc, qac):
    """Any positive QP is accepted by the reference CLI (ENC:94-165); the magic-number division must stay exact."""
    clip = synth.make_clip("highmotion", 4, 77)
    res = gpu.encode_sequence(clip, qdc, qac, 4)
    s = oracle.encode(clip, W, H, qdc, qac, 4)
    assert_syntax_equal(res, s, what=f"q={qdc}/{qac}: ")
    data, _ = gpu.encode_sequence_bitstream(clip, qdc, qac, 4)
    assert data == oracle.write_bitstream(s, W, H, qdc, qac, 4)
    out = gpu.decode_sequence(res.levels, res.mpm, res.ipm, res.mvd, qdc, qac, 4)
    assert np.array_equal(out, oracle.decode(s, W, H, qdc, qac, 4))


def test_extreme_content(gpu, oracle):
    """Saturated / checkerboard / black frames: clipping paths, maximum residuals, all-zero blocks, zero-SAD everywhere."""
    n = 4
    y = np.zeros((n, H, W), np.uint8)
    y[0] = 255
    y[1, ::2, ::2] = 255; y[1, 1::2, 1::2] = 255
    y[2] = 0
    y[3] = (np.indices((H, W)).sum(0) % 256).astype(np.uint8)
    c = np.zeros((n, 2, H // 2, W // 2), np.uint8)
    c[0] = 255; c[1] = 0; c[2] = 128; c[3, 0] = 255; c[3, 1] = 1
    frames = np.concatenate([y.reshape(n, -1), c.reshape(n, -1)], axis=1)
    for q, ip in ((1, 4), (8, 2), (16, 1)):
        res = gpu.encode_sequence(frames, q, q, ip)
        s = oracle.encode(frames, W, H, q, q, ip)
        assert_syntax_equal(res, s, what=f"extreme q={q} ip={ip}: ")
        out = gpu.decode_sequence(res.levels, res.mpm, res.ipm, res.mvd, q, q, ip)
        assert np.array_equal(out, oracle.decode(s, W, H, q, q, ip))


def _fuzz_clip(rng, n, w, h):
    """Random mixture of flat areas, copied/shifted patches (exact matches -> zero-SAD early breaks and carried spiral
    state), noise and gradients, so that rare control paths are hit."""
    y = np.zeros((n, h, w), np.uint8)
    base = rng.integers(0, 256, size=(h, w)).astype(np.uint8)
    if rng.random() < 0.5:
        base[:, : w // 2] = rng.integers(0, 256)
    if rng.random() < 0.5:
        base[h // 2:, :] = rng.integers(0, 256)
    for i in range(n):
        mode = rng.integers(0, 4)
        if i == 0 or mode == 0:
            f = base.copy()
        elif mode == 1:                                   # exact global shift of the previous frame
            f = np.roll(y[i - 1], (int(rng.integers(-3, 4)), int(rng.integers(-3, 4))), axis=(0, 1))
        elif mode == 2:                                   # previous frame + sparse noise
            f = y[i - 1].copy()
            m = rng.random((h, w)) < 0.02
            f[m] = rng.integers(0, 256, size=int(m.sum()))
        else:                                             # identical frame
            f = y[i - 1].copy()
        y[i] = f
    cb = rng.integers(0, 256, size=(n, h // 2, w // 2)).astype(np.uint8) if rng.random() < 0.5 else np.full((n, h // 2, w // 2), 128, np.uint8)
    cr = np.roll(cb, 1, axis=2)
    return np.concatenate([y.reshape(n, -1), cb.reshape(n, -1), cr.reshape(n, -1)], axis=1)


def test_fuzz_small_geometries(oracle):
    """Seeded fuzzing on small frames (16x16 .. 96x64), many QP / intra-period combinations, encoder + GPU bitstream +
    decoder against the oracle."""
    from icspcodec_b200 import IcspCuda
    rng = np.random.default_rng(20261017)
    geoms = [(16, 16), (32, 16), (48, 32), (64, 64), (96, 64), (80, 48)]
    for w, h in geoms:
        with IcspCuda(w, h, max_frames=16) as ctx:
            for trial in range(6):
                n = int(rng.integers(2, 9))
                ip = int(rng.choice([0, 1, 2, 3, 5, 8]))
                qdc, qac = int(rng.choice([1, 2, 5, 8, 16, 31])), int(rng.choice([1, 3, 8, 16, 40]))
                clip = _fuzz_clip(rng, n, w, h)
                res = ctx.encode_sequence(clip, qdc, qac, ip)
                s = oracle.encode(clip, w, h, qdc, qac, ip)
                assert_syntax_equal(res, s, what=f"{w}x{h} n={n} ip={ip} q={qdc}/{qac}: ")
                data, _ = ctx.encode_sequence_bitstream(clip, qdc, qac, ip)
                assert data == oracle.write_bitstream(s, w, h, qdc, qac, ip), f"{w}x{h} bitstream"
                if ip > 0:
                    out = ctx.decode_sequence(res.levels, res.mpm, res.ipm, res.mvd, qdc, qac, ip)
                    assert np.array_equal(out, oracle.decode(s, w, h, qdc, qac, ip))


def test_fuzz_cif_early_breaks(gpu, oracle):
    """CIF clips built from exact shifts / repeats: thousands of zero-SAD early breaks with the spiral state carried
    across macroblocks (ENC:2094-2148), through the exact fallback kernels."""
    rng = np.random.default_rng(7)
    for trial in range(3):
        clip = _fuzz_clip(rng, 6, W, H)
        res = gpu.encode_sequence(clip, 8, 8, 6)
        s = oracle.encode(clip, W, H, 8, 8, 6)
        assert_syntax_equal(res, s, what=f"trial {trial}: ")
        ys = clip[:, : W * H]
        _, _, evals = oracle.me(ys[1], s.recon[0][: W * H], W, H)
        assert evals <= 396 * 64


def test_plane_sse_matches_numpy(gpu):
    """icsp_enc_sse (§8 f4): integer SSE per plane of the resident frames vs their reconstruction, all three planes,
    plus psnr_y against the reference decoder's double formula (DEC.h:332-348)."""
    clip = synth.make_clip("highmotion", 6, 11)
    res = gpu.encode_sequence(clip, 16, 16, 3)
    sse = gpu.enc_sse(6)
    a, b = clip.reshape(6, -1).astype(np.int64), res.recon.reshape(6, -1).astype(np.int64)
    d2 = (a - b) ** 2
    ysz, csz = W * H, W * H // 4
    want = np.stack([d2[:, :ysz].sum(1), d2[:, ysz: ysz + csz].sum(1), d2[:, ysz + csz:].sum(1)], axis=1)
    assert np.array_equal(sse.astype(np.int64), want)
    mse = want[:, 0].astype(np.float64) / (W * H)
    assert gpu.psnr_y(6) == float(np.mean(20.0 * np.log10(255.0 / np.sqrt(mse))))
    with pytest.raises(Exception):
        gpu.enc_sse(10 ** 9)


# ---- bit reader on the GPU (macroblock-row index + parse_rows_kernel) -----------------------------
@pytest.mark.parametrize("case", [c for c in CASES if c["ip"] > 0], ids=lambda c: f"{c['kind']}-q{c['qdc']}_{c['qac']}-ip{c['ip']}")
def test_gpu_bit_reader_matches_reference_decoder(gpu, case):
    """file + row index from the GPU encoder -> decode_sequence_bitstream: the file is the reference encoder's (md5) and
    the decoded YUV is the reference decoder's output on it (md5), incl. tail GOPs and the right-aligned last byte."""
    clip = synth.make_clip(case["kind"], case["nframes"], case["seed"])
    data, _, rows = gpu.encode_sequence_bitstream(clip, case["qdc"], case["qac"], case["ip"], want_index=True)
    assert hashlib.md5(data).hexdigest() == case["bin_md5"]
    assert rows.shape == (case["nframes"], H // 16) and rows[0, 0] == 0 and np.all(np.diff(rows.reshape(-1).astype(np.int64)) > 0)
    out = gpu.decode_sequence_bitstream(data, rows, case["nframes"])
    assert hashlib.md5(out.tobytes()).hexdigest() == case["dec_md5"]
    # the index exported from the GPU's prefix sums is the one the host writer computes for the same syntax
    from icspcodec_b200 import hostlib
    res = gpu.encode_sequence(clip, case["qdc"], case["qac"], case["ip"])
    data_h, rows_h = hostlib.write_stream_indexed(res.levels, res.acflag, res.mpm, res.ipm, res.mvd, W, H, case["qdc"], case["qac"], case["ip"])
    assert data_h == data and np.array_equal(rows_h, rows)


def test_gpu_bit_reader_stream_batch(oracle):
    """decode_streams on a batch (5 streams x 3 GOPs x 4 frames, several chunks) against the host parser + oracle decoder,
    and syntax-level equality of what the GPU parser produced (levels / mpm / ipm / mvd read back through the decoder API)."""
    from icspcodec_b200 import IcspCuda, hostlib
    from icspcodec_b200.api import stream_header
    ns, gps, gl, qdc, qac = 5, 3, 4, 8, 6
    fps = gps * gl
    clips = [synth.make_clip(k, fps, 40 + i) for i, k in enumerate(["highmotion", "akiyo", "flat", "highmotion", "intra"])]
    batch = np.concatenate([c.reshape(fps, -1) for c in clips])
    with IcspCuda(W, H, max_frames=ns * fps) as ctx:
        ctx.configure(4, gps)                                    # one stream per chunk: exercises the chunked pipeline
        bodies, sbits, _ = ctx.encode_streams(batch, ns, gps, gl, qdc, qac)
        rows = ctx.bits_row_index(ns * fps)
        files, offs, lens, blob = [], [], [], bytearray()
        for s in range(ns):
            nb = int(sbits[s])
            b = bytearray(bytes(bodies[s]))
            if nb % 8 == 0:
                b += b"\x00"
            else:
                b[-1] = b[-1] >> (8 - nb % 8)                    # reference tail rule (ENC:4895)
            while len(blob) % 4:
                blob += b"\x00"
            offs.append(len(blob)); lens.append(len(b)); blob += b
            files.append(bytes(stream_header(W, H, qdc, qac, gl)) + bytes(b))
        out = ctx.decode_streams(np.frombuffer(bytes(blob), np.uint8), offs, lens, rows, ns, gps, gl, qdc, qac)
        for s in range(ns):
            syn = oracle.encode(clips[s], W, H, qdc, qac, gl)
            assert files[s] == oracle.write_bitstream(syn, W, H, qdc, qac, gl)
            want = oracle.decode(syn, W, H, qdc, qac, gl)
            assert np.array_equal(out[s * fps:(s + 1) * fps], want), f"stream {s}"
            parsed, _ = hostlib.parse_stream(files[s], fps)      # host reader on the same file -> same pictures
            got = ctx.decode_gops(parsed["levels"], parsed["mpm"], parsed["ipm"], parsed["mvd"], gps, gl, qdc, qac)
            assert np.array_equal(got, want)
        # a corrupt index must fail validation or decode garbage, never crash
        bad = rows.copy(); bad[3, 5] = np.uint64(10 ** 15)
        with pytest.raises(Exception):
            ctx.decode_streams(np.frombuffer(bytes(blob), np.uint8), offs, lens, bad, ns, gps, gl, qdc, qac)
        bad = rows.copy(); bad[:, 1:] += np.uint64(3)
        ctx.decode_streams(np.frombuffer(bytes(blob), np.uint8), offs, lens, bad, ns, gps, gl, qdc, qac)


# ---- BASELINE.json sizes: every config at its stated 300 frames (tests/golden/make_golden_full.py) ----------------------
FULL = json.load(open(os.path.join(GOLD, "ref_cases_full.json")))


@pytest.fixture(scope="module")
def gpu300():
    from icspcodec_b200 import IcspCuda
    ctx = IcspCuda(W, H, max_frames=300)
    yield ctx
    ctx.close()


@pytest.mark.parametrize("case", FULL, ids=lambda c: c["config"].split(",")[0].replace(" ", "_"))
def test_baseline_configs_full_size(gpu300, case):
    """configs[0], [1] (QP 1/8/16), [2], the first streams of [3] and [4] at 300 frames: the bitstream produced on the GPU,
    the reconstruction and the decoded YUV (bit reader on the GPU, then dequant/IDCT/MC/recon) are byte-identical to what
    the unmodified reference encoder / decoder wrote for the same input."""
    n, qdc, qac, ip = case["nframes"], case["qdc"], case["qac"], case["ip"]
    clip = synth.make_clip(case["kind"], n, case["seed"])
    assert md5(clip.tobytes()) == case["clip_md5"]
    data, recon, rows = gpu300.encode_sequence_bitstream(clip, qdc, qac, ip, want_recon=True, want_index=True)
    assert len(data) == case["bin_len"]
    assert md5(data) == case["bin_md5"]
    assert md5(recon.tobytes()) == case["recon_md5"]
    dec = gpu300.decode_sequence_bitstream(data, rows, n)
    assert md5(dec.tobytes()) == case["dec_md5"]


def test_batch_of_streams_full_size_matches_golden():
    """configs[3] shape in one call: 4 streams x 300 frames (120 GOPs per launch, several pipelined chunks) — every stream's
    body and reconstruction equal the reference's for that stream alone."""
    from icspcodec_b200 import IcspCuda, finish_stream
    cases = [c for c in FULL if c["config"].startswith("configs[3]")]
    frames = np.concatenate([synth.make_clip(c["kind"], 300, c["seed"]) for c in cases], axis=0)
    with IcspCuda(W, H, max_frames=frames.shape[0]) as ctx:
        ctx.configure(4, 15)                                     # 8 chunks of 15 GOPs: exercises the chunked pipeline
        bodies, sbits, recon = ctx.encode_streams(frames, len(cases), 30, 10, 8, 8, want_recon=True)
        for s, c in enumerate(cases):
            assert md5(finish_stream(bodies[s], int(sbits[s]), W, H, 8, 8, 10)) == c["bin_md5"], f"stream {s}"
            assert md5(recon[s * 300:(s + 1) * 300].tobytes()) == c["recon_md5"], f"stream {s}"


# ---- SURVEY 8(b) unit shims: icsp_quant, icsp_intra_frame, icsp_inter_frame, the DCT tap -------------------------------
def test_quant_shim_matches_reference(gpu):
    """Quantization_block / CQuantization_block of the compiled reference (tests/golden/ref_quant.npz: realistic coefficients,
    exact .5 ties on both sides of zero, all-zero-AC blocks) for five quantiser pairs, luma (truncate) and chroma (floor)."""
    z = np.load(os.path.join(GOLD, "ref_quant.npz"))
    for qdc, qac in ((1, 1), (8, 8), (16, 16), (8, 3), (100, 255)):
        for chroma in (0, 1):
            lv, ac = gpu.quant(z["dct"], qdc, qac, bool(chroma))
            assert np.array_equal(lv, z[f"lv_{qdc}_{qac}_{chroma}"]), (qdc, qac, chroma)
            assert np.array_equal(ac, z[f"ac_{qdc}_{qac}_{chroma}"]), (qdc, qac, chroma)


def test_intra_and_inter_frame_shims(gpu, oracle):
    """icsp_intra_frame == intraPrediction on independent frames; icsp_inter_frame(cur, prev reconstruction) ==
    interPrediction(cur, prev): every SoA output of the frame, against the oracle run on the same pair."""
    clips = [synth.make_clip("highmotion", 2, 77), synth.make_clip("flat", 2, 7), synth.make_clip("akiyo", 2, 5)]
    first = np.stack([c[0] for c in clips])
    ri = gpu.intra_frame(first, 8, 16)
    prev_recon, cur = [], []
    for i, c in enumerate(clips):
        s = oracle.encode(c, W, H, 8, 16, 2)
        one = type(ri)(**{k: (None if getattr(ri, k) is None else getattr(ri, k)[i:i + 1]) for k in ri.__dataclass_fields__})
        so = type(s)(**{k: (None if getattr(s, k) is None else getattr(s, k)[0:1]) for k in s.__dataclass_fields__})
        assert_syntax_equal(one, so, what=f"intra {i}: ")
        prev_recon.append(s.recon[0]); cur.append(c[1])
    rp = gpu.inter_frame(np.stack(cur), np.stack(prev_recon), 8, 16)
    for i, c in enumerate(clips):
        s = oracle.encode(c, W, H, 8, 16, 2)
        one = type(rp)(**{k: (None if getattr(rp, k) is None else getattr(rp, k)[i:i + 1]) for k in rp.__dataclass_fields__})
        so = type(s)(**{k: (None if getattr(s, k) is None else getattr(s, k)[1:2]) for k in s.__dataclass_fields__})
        assert_syntax_equal(one, so, what=f"inter {i}: ")


def test_dct_tap_in_pipeline(gpu, oracle):
    """The forward-DCT doubles INSIDE the encode pipeline (intra wavefront, inter luma, chroma of both frame types), through
    the optional icsp_enc_out.dct tap, against the oracle's tap: within the north_star tolerance 1e-9 — and in fact 0 ulp."""
    clip = synth.make_clip("highmotion", 4, 2024)
    res = gpu.encode_gops(clip, 2, 2, 8, 8, with_dct=True)
    for g in range(2):
        s = oracle.encode(clip[2 * g:2 * g + 2], W, H, 8, 8, 2, with_dct=True)
        got = res.dct[2 * g:2 * g + 2]
        np.testing.assert_allclose(got, s.dct, rtol=1e-9, atol=1e-9)
        assert np.array_equal(got, s.dct)
    assert np.array_equal(res.recon[:2], oracle.encode(clip[:2], W, H, 8, 8, 2).recon)      # the tap changes nothing else
    rp = gpu.inter_frame(clip[1:2], res.recon[0:1], 8, 8, with_dct=True)
    assert np.array_equal(rp.dct[0], oracle.encode(clip[:2], W, H, 8, 8, 2, with_dct=True).dct[1])


def test_cuda_graph_replay_matches_plain_launches(oracle, monkeypatch):
    """The per-chunk kernel sequences are captured into CUDA graphs on first use and replayed afterwards: first call (capture +
    launch), replays, a different shape in between, and a context with ICSP_GRAPHS=0 must all give the oracle's bits."""
    from icspcodec_b200 import IcspCuda
    clips = [synth.make_clip("highmotion", 6, 40 + i) for i in range(3)] + [synth.make_clip("flat", 6, 7)]
    frames = np.concatenate(clips, axis=0)
    want = [oracle.encode(c, W, H, 8, 8, 3) for c in clips]

    def check(ctx):
        for rep in range(3):                                   # capture, then two replays
            res = ctx.encode_gops(frames, 8, 3, 8, 8)
            for i, s in enumerate(want):
                sub = type(res)(**{k: (None if getattr(res, k) is None else getattr(res, k)[6 * i:6 * i + 6]) for k in res.__dataclass_fields__})
                assert_syntax_equal(sub, s, what=f"rep {rep} clip {i}: ")
            if rep == 0:                                       # another shape in between: its own graph, must not disturb the first
                other = ctx.encode_gops(frames[:6], 1, 6, 16, 16)
                assert np.array_equal(other.recon, oracle.encode(clips[0], W, H, 16, 16, 6).recon)
        bodies, sbits, _ = ctx.encode_streams(frames, 4, 2, 3, 8, 8)
        bodies2, sbits2, _ = ctx.encode_streams(frames, 4, 2, 3, 8, 8)
        assert np.array_equal(sbits, sbits2) and all(np.array_equal(a, b) for a, b in zip(bodies, bodies2))
        return ctx.launch_count()

    with IcspCuda(W, H, max_frames=32) as ctx:
        with_graphs = check(ctx)
    monkeypatch.setenv("ICSP_GRAPHS", "0")
    with IcspCuda(W, H, max_frames=32) as ctx:
        plain = check(ctx)
    assert with_graphs == plain                                # replays are counted like the launches they contain


@pytest.mark.parametrize("geom", [(W, H), (64, 48), (1280, 720)])
def test_dc_chain_integer_path_and_double_path_agree(oracle, monkeypatch, geom):
    """The DC-DPCM chain codes a block with integer arithmetic unless raw DC + 0.5 is within 2^-30 of an integer, where it falls
    back to the reference's double sequence (dc_chain_kernel).  Both paths must give the oracle's bits: the default context
    (integer path for practically every block) and one with ICSP_DC_EPS=1, which sends EVERY block down the double path.  Flat
    and saturated content puts raw DCs exactly on integers and on k + 0.5 (the ambiguous cases proper)."""
    from icspcodec_b200 import IcspCuda
    w, h = geom
    kinds = ["highmotion", "flat", "akiyo", "saturated"] if (w, h) == (W, H) else ["noise"]
    for kind in kinds:
        if kind == "noise":
            clip = np.random.default_rng(5).integers(0, 256, size=(4, w * h * 3 // 2), dtype=np.uint8)
        elif kind == "saturated":                   # 8x8 tiles of 0 / 255 that move by one tile per frame: DCs of +-2040, +-1020, 0
            yy, xx = np.mgrid[0:h * 3 // 2, 0:w]
            clip = np.stack([(((yy // 8 + xx // 8 + t) & 1) * 255).astype(np.uint8).reshape(-1) for t in range(4)])
        else:
            clip = synth.make_clip(kind, 4, 11)
        for qdc, qac, ip in [(8, 8, 2), (1, 3, 4), (255, 100, 0)]:
            want = oracle.encode(clip, w, h, qdc, qac, ip)
            for eps in (None, "1"):
                if eps is None:
                    monkeypatch.delenv("ICSP_DC_EPS", raising=False)
                else:
                    monkeypatch.setenv("ICSP_DC_EPS", eps)
                with IcspCuda(w, h, max_frames=8) as ctx:
                    gop = 1 if ip == 0 else ip
                    got = ctx.encode_gops(clip, 4 // gop, gop, qdc, qac)
                    assert_syntax_equal(got, want, what=f"{w}x{h} {kind} qp {qdc}/{qac} ip {ip} eps {eps}: ")
                    dec = ctx.decode_gops(got.levels, got.mpm, got.ipm, got.mvd, 4 // gop, gop, qdc, qac)
                    assert np.array_equal(dec, oracle.decode(want, w, h, qdc, qac, max(ip, 1))), f"decode {w}x{h} {kind} qp {qdc}/{qac} ip {ip} eps {eps}"


@pytest.mark.parametrize("env", [{"ICSP_ME_ROWS_G": "0"}, {"ICSP_ME_ROWS_G": "1000000"}, {"ICSP_ME_ROWS_G": "0", "ICSP_ME_FUSED": "0"},
                                 {"ICSP_ME_ROWS_G": "1000000", "ICSP_ME_FUSED": "0"}],
                         ids=["persistent+fused", "rows+one-fallback-launch", "persistent+3-launches", "rows+3-launches"])
def test_me_launch_shapes_agree(oracle, monkeypatch, env):
    """The motion search has two launch shapes — one persistent CTA per frame (big batches) and one CTA per macroblock row
    (small batches, ICSP_ME_ROWS_G) — and the exact carried-state fallback runs inside the frame kernel, as one more launch, or
    as three.  Every combination must give the oracle's vectors on clips full of zero-SAD early breaks, also on a frame that is two
    segments wide (368 = 23 macroblocks) where only the three-launch fallback applies."""
    from icspcodec_b200 import IcspCuda
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(99)
    for w, h in ((W, H), (368, 48)):
        with IcspCuda(w, h, max_frames=12) as ctx:
            for trial in range(2):
                clip = _fuzz_clip(rng, 6, w, h)
                res = ctx.encode_gops(np.concatenate([clip, clip[::-1]]), 2, 6, 8, 8)
                for i, c in enumerate((clip, clip[::-1])):
                    s = oracle.encode(np.ascontiguousarray(c), w, h, 8, 8, 6)
                    sub = type(res)(**{k: (None if getattr(res, k) is None else getattr(res, k)[6 * i:6 * i + 6]) for k in res.__dataclass_fields__})
                    assert_syntax_equal(sub, s, what=f"{env} {w}x{h} trial {trial} gop {i}: ")


@pytest.mark.parametrize("wide_g", ["0", "100000"], ids=["96-thread-frames", "192-thread-frames"])
def test_intra_wavefront_cta_widths_agree(oracle, monkeypatch, wide_g):
    """The intra wavefront runs with 96 threads per frame in big batches and with 192 in small ones (ICSP_INTRA_WIDE_G): both
    widths, encoder and decoder, against the oracle."""
    from icspcodec_b200 import IcspCuda
    monkeypatch.setenv("ICSP_INTRA_WIDE_G", wide_g)
    for (w, h), kind in (((W, H), "highmotion"), ((W, H), "flat"), ((64, 48), None), ((720, 480), None)):
        clip = synth.make_clip(kind, 4, 3) if kind else np.random.default_rng(8).integers(0, 256, size=(4, w * h * 3 // 2), dtype=np.uint8)
        with IcspCuda(w, h, max_frames=8) as ctx:
            for qdc, qac, ip in ((8, 8, 0), (3, 16, 2)):
                want = oracle.encode(clip, w, h, qdc, qac, ip)
                gop = 1 if ip == 0 else ip
                got = ctx.encode_gops(clip, 4 // gop, gop, qdc, qac)
                assert_syntax_equal(got, want, what=f"{w}x{h} {kind} qp {qdc}/{qac} ip {ip} wide_g {wide_g}: ")
                dec = ctx.decode_gops(got.levels, got.mpm, got.ipm, got.mvd, 4 // gop, gop, qdc, qac)
                assert np.array_equal(dec, oracle.decode(want, w, h, qdc, qac, max(ip, 1)))     # the decoder's all-intra is period 1


@pytest.mark.parametrize("slice_copies", ["1", "0"], ids=["copies-sliced-by-step", "whole-chunk-copies"])
def test_one_chunk_host_call_copy_modes_agree(oracle, monkeypatch, slice_copies):
    """A host call that is a single chunk uploads / downloads frame t of every GOP as its own strided copy around step t
    (ICSP_SLICE_COPIES); with the switch off the whole chunk is copied before / after.  Same bits either way, also when the call is
    repeated (graph replay per step) and for GOP lengths from 2 to 30."""
    from icspcodec_b200 import IcspCuda, api
    monkeypatch.setenv("ICSP_SLICE_COPIES", slice_copies)
    with IcspCuda(W, H, max_frames=64) as ctx:
        for gop_len, gps, ns in ((3, 2, 3), (30, 1, 2), (2, 5, 1)):
            n = gop_len * gps
            clips = [synth.make_clip("highmotion", n, 70 + i) for i in range(ns)]
            frames = np.concatenate(clips, axis=0)
            want = [oracle.encode(cl, W, H, 8, 6, gop_len) for cl in clips]
            for rep in range(2):
                bodies, sbits, rec = ctx.encode_streams(frames, ns, gps, gop_len, 8, 6, want_recon=True)
                for i, s in enumerate(want):
                    assert np.array_equal(rec[i * n:(i + 1) * n], s.recon), f"recon stream {i} gop_len {gop_len} rep {rep}"
                    ref = oracle.write_bitstream(s, W, H, 8, 6, gop_len)
                    got = api.finish_stream(bodies[i], int(sbits[i]), W, H, 8, 6, gop_len)
                    assert got == ref, f"bitstream stream {i} gop_len {gop_len} rep {rep}"
                # and back: the GPU bit reader + decoder on the same one-chunk shape (the decoded frames return slice by slice)
                rows = ctx.bits_row_index(ns * n)
                blob, offs, lens = bytearray(), [], []
                for i in range(ns):
                    body = api.finish_stream(bodies[i], int(sbits[i]), W, H, 8, 6, gop_len)[14:]      # file body without the header
                    while len(blob) % 4:
                        blob += b"\x00"
                    offs.append(len(blob)); lens.append(len(body)); blob += body
                dec = ctx.decode_streams(np.frombuffer(bytes(blob), np.uint8), offs, lens, rows, ns, gps, gop_len, 8, 6)
                for i, s in enumerate(want):
                    assert np.array_equal(dec[i * n:(i + 1) * n], oracle.decode(s, W, H, 8, 6, gop_len)), f"decode stream {i} gop_len {gop_len} rep {rep}"
