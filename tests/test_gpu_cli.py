"""GPU tests of the drop-in CLIs (icspcodec_b200/host/icspenc, icspdec): same options as the reference's
encoder/decoder, outputs compared byte for byte with what the compiled reference produced (golden md5s), and —
when oracle/_ref travelled to this box — with the reference binaries run live on the same input."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from icspcodec_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CASES = json.load(open(os.path.join(GOLD, "ref_cases.json")))
ENC = os.path.join(ROOT, "icspcodec_b200", "host", "icspenc")
DEC = os.path.join(ROOT, "icspcodec_b200", "host", "icspdec")


def md5f(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


@pytest.fixture(scope="module", autouse=True)
def built():
    from icspcodec_b200 import build
    build.build()
    assert os.path.exists(ENC) and os.path.exists(DEC)


@pytest.mark.parametrize("case", [CASES[0], CASES[3], CASES[5], CASES[6], CASES[8]], ids=lambda c: f"{c['kind']}-q{c['qdc']}_{c['qac']}-ip{c['ip']}")
def test_icspenc_icspdec_match_reference(tmp_path, case):
    clip = synth.make_clip(case["kind"], case["nframes"], case["seed"])
    clip.tofile(tmp_path / "clip_cif.yuv")
    n, qdc, qac, ip = case["nframes"], case["qdc"], case["qac"], case["ip"]
    subprocess.run([ENC, "-i", "clip_cif.yuv", "-n", str(n), "--qpdc", str(qdc), "--qpac", str(qac), "--intraPeriod", str(ip)],
                   cwd=tmp_path, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    binp = tmp_path / f"clip_compCIF_{qdc}_{qac}_{ip}.bin"
    assert md5f(binp) == case["bin_md5"]                       # final bitstream, bit-exact vs the reference encoder
    assert md5f(tmp_path / "test_yuv.yuv") == case["recon_md5"]  # reconstructed YUV
    if ip > 0:
        subprocess.run([DEC, str(n), binp.name, str(qdc), str(qac), str(ip), "clip_cif.yuv"], cwd=tmp_path, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        out = "check_test_intra_yuv.yuv" if ip == 1 else "check_test_inter_yuv.yuv"
        assert md5f(tmp_path / out) == case["dec_md5"]         # bit-exact YUV vs the reference decoder
        assert (tmp_path / "experimental_Result_Decoding.txt").exists()


def test_cli_options_and_tail_frames(tmp_path, oracle):
    """-q, -w/-h, a frame count that is not a multiple of intraPeriod (the reference's MT mode drops the tail,
    ICSP_thread.cpp:43; single-thread mode and this tool encode it), and --gpus 1."""
    w, h, n = 176, 144, 7
    rng = np.random.default_rng(3)
    base = rng.integers(0, 200, size=(h + 8, w + 8)).astype(np.float64)
    base = (base + np.roll(base, 1, 0) + np.roll(base, 1, 1)) / 3
    frames = []
    for i in range(n):
        y = np.clip(np.rint(base[i:i + h, i:i + w]), 0, 255).astype(np.uint8)
        frames.append(np.concatenate([y.ravel(), np.full(w * h // 4, 120 + i, np.uint8), np.full(w * h // 4, 130 - i, np.uint8)]))
    frames = np.stack(frames)
    frames.tofile(tmp_path / "q_cif.yuv")
    subprocess.run([ENC, "-i", "q_cif.yuv", "-n", str(n), "-q", "16", "--intraPeriod", "3", "-w", str(w), "-h", str(h), "--gpus", "1", "--quiet"],
                   cwd=tmp_path, check=True)
    s = oracle.encode(frames, w, h, 16, 16, 3)
    assert open(tmp_path / "q_compCIF_16_16_3.bin", "rb").read() == oracle.write_bitstream(s, w, h, 16, 16, 3)
    assert np.array_equal(np.fromfile(tmp_path / "test_yuv.yuv", np.uint8).reshape(n, -1), s.recon)
    subprocess.run([DEC, str(n), "q_compCIF_16_16_3.bin", "16", "16", "3", "--out", "dec.yuv"], cwd=tmp_path, check=True, stderr=subprocess.DEVNULL)
    ps, _ = oracle.parse_bitstream(open(tmp_path / "q_compCIF_16_16_3.bin", "rb").read(), n)
    assert np.array_equal(np.fromfile(tmp_path / "dec.yuv", np.uint8).reshape(n, -1), oracle.decode(ps, w, h, 16, 16, 3))
    # help / bad arguments never crash the process with a signal
    assert subprocess.run([ENC, "--help"], capture_output=True).returncode == 0
    assert subprocess.run([ENC, "-i", "q_cif.yuv", "-n", "2", "-q", "0"], cwd=tmp_path, capture_output=True).returncode == 1


def test_many_device_calls_and_host_entropy(tmp_path):
    """A sequence larger than one device call: per-call bit strings are concatenated at bit granularity on the host
    (frames are not byte aligned in the stream); and the --host-entropy path gives the same file."""
    case = CASES[6]   # highmotion 12 frames ip 30 -> one GOP of 12; use ip 3 -> 4 GOPs, 1 GOP per call
    clip = synth.make_clip(case["kind"], 12, case["seed"])
    clip.tofile(tmp_path / "m_cif.yuv")
    env = dict(os.environ, ICSPENC_CALL_FRAMES="3")
    subprocess.run([ENC, "-i", "m_cif.yuv", "-n", "12", "-q", "8", "--intraPeriod", "3", "--quiet"], cwd=tmp_path, check=True, env=env)
    a = open(tmp_path / "m_compCIF_8_8_3.bin", "rb").read()
    ya = open(tmp_path / "test_yuv.yuv", "rb").read()
    subprocess.run([ENC, "-i", "m_cif.yuv", "-n", "12", "-q", "8", "--intraPeriod", "3", "--quiet", "--host-entropy"], cwd=tmp_path, check=True)
    assert open(tmp_path / "m_compCIF_8_8_3.bin", "rb").read() == a
    assert open(tmp_path / "test_yuv.yuv", "rb").read() == ya
    subprocess.run([ENC, "-i", "m_cif.yuv", "-n", "12", "-q", "8", "--intraPeriod", "3", "--quiet"], cwd=tmp_path, check=True)
    assert open(tmp_path / "m_compCIF_8_8_3.bin", "rb").read() == a


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ICSPCodec_O2")), reason="oracle/_ref not on this box")
def test_live_reference_binaries(tmp_path, oracle):
    clip = synth.make_clip("highmotion", 8, 2025)
    clip.tofile(tmp_path / "live_cif.yuv")
    subprocess.run([ENC, "-i", "live_cif.yuv", "-n", "8", "-q", "8", "--intraPeriod", "4", "--quiet"], cwd=tmp_path, check=True)
    rbin, rrec = oracle.ref_encode(clip, 8, 8, 4)
    assert open(tmp_path / "live_compCIF_8_8_4.bin", "rb").read() == rbin
    assert np.array_equal(np.fromfile(tmp_path / "test_yuv.yuv", np.uint8).reshape(8, -1), rrec)
    subprocess.run([DEC, "8", "live_compCIF_8_8_4.bin", "8", "8", "4", "--out", "d.yuv"], cwd=tmp_path, check=True, stderr=subprocess.DEVNULL)
    assert np.array_equal(np.fromfile(tmp_path / "d.yuv", np.uint8).reshape(8, -1), oracle.ref_decode(rbin, 8, 8, 8, 4))


def test_icspenc_psnr_on_gpu(tmp_path):
    """--psnr: the figure the reference's decoder logs (mean luma PSNR, DEC.h:332-350), reduced on the GPU
    from the resident frames, also with --no-recon; checked against the same formula on the written reconstruction."""
    case = CASES[0]
    n, qdc, qac, ip = case["nframes"], case["qdc"], case["qac"], case["ip"]
    clip = synth.make_clip(case["kind"], n, case["seed"])
    clip.tofile(tmp_path / "clip_cif.yuv")
    r = subprocess.run([ENC, "-i", "clip_cif.yuv", "-n", str(n), "--qpdc", str(qdc), "--qpac", str(qac), "--intraPeriod", str(ip), "--psnr", "--quiet"],
                       cwd=tmp_path, check=True, capture_output=True, text=True)
    assert md5f(tmp_path / "test_yuv.yuv") == case["recon_md5"]
    rec = np.fromfile(tmp_path / "test_yuv.yuv", np.uint8).reshape(n, -1)[:, : 352 * 288].astype(np.float64)
    src = clip.reshape(n, -1)[:, : 352 * 288].astype(np.float64)
    mse = ((src - rec) ** 2).sum(axis=1) / (352 * 288)
    want = float(np.mean(20.0 * np.log10(255.0 / np.sqrt(mse))))
    line = [l for l in r.stdout.splitlines() if l.startswith("PSNR:")][0]
    assert line == f"PSNR: {want:.4f} QPDC: {qdc}  QPAC: {qac} Period: {ip}"
    os.remove(tmp_path / "test_yuv.yuv")
    r2 = subprocess.run([ENC, "-i", "clip_cif.yuv", "-n", str(n), "--qpdc", str(qdc), "--qpac", str(qac), "--intraPeriod", str(ip), "--psnr", "--quiet", "--no-recon"],
                        cwd=tmp_path, check=True, capture_output=True, text=True)
    assert [l for l in r2.stdout.splitlines() if l.startswith("PSNR:")][0] == line
    assert not (tmp_path / "test_yuv.yuv").exists()


@pytest.mark.parametrize("case", [CASES[0], CASES[3], CASES[6], CASES[7]], ids=lambda c: f"{c['kind']}-q{c['qdc']}_{c['qac']}-ip{c['ip']}")
def test_index_sidecar_gpu_bit_reader(tmp_path, case):
    """icspenc --index writes <bin>.idx; icspdec finds it and parses on the GPU: same YUV as the reference
    decoder (md5), same as the host-parser path (--no-index), also when the sequence is split into several device calls."""
    n, qdc, qac, ip = case["nframes"], case["qdc"], case["qac"], case["ip"]
    synth.make_clip(case["kind"], n, case["seed"]).tofile(tmp_path / "clip_cif.yuv")
    env = dict(os.environ, ICSPENC_CALL_FRAMES=str(ip))          # one device call per GOP: the per-call indices are rebased
    subprocess.run([ENC, "-i", "clip_cif.yuv", "-n", str(n), "--qpdc", str(qdc), "--qpac", str(qac), "--intraPeriod", str(ip), "--index", "--quiet"],
                   cwd=tmp_path, check=True, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    binp = tmp_path / f"clip_compCIF_{qdc}_{qac}_{ip}.bin"
    assert md5f(binp) == case["bin_md5"]
    idx = np.fromfile(str(binp) + ".idx", np.uint8)
    assert idx[:8].tobytes() == b"ICSPIDX1" and idx.size == 24 + 8 * n * 18
    out = "check_test_intra_yuv.yuv" if ip == 1 else "check_test_inter_yuv.yuv"
    r = subprocess.run([DEC, str(n), binp.name, str(qdc), str(qac), str(ip)], cwd=tmp_path, check=True, capture_output=True, text=True)
    assert "bit reader on the GPU" in r.stderr
    assert md5f(tmp_path / out) == case["dec_md5"]
    os.remove(tmp_path / out)
    r = subprocess.run([DEC, str(n), binp.name, str(qdc), str(qac), str(ip), "--no-index"], cwd=tmp_path, check=True, capture_output=True, text=True)
    assert "bit reader on the host" in r.stderr
    assert md5f(tmp_path / out) == case["dec_md5"]
    # the host entropy coder writes the same side-car
    idx_gpu = md5f(str(binp) + ".idx")
    os.remove(str(binp) + ".idx")
    subprocess.run([ENC, "-i", "clip_cif.yuv", "-n", str(n), "--qpdc", str(qdc), "--qpac", str(qac), "--intraPeriod", str(ip), "--index", "--host-entropy", "--quiet"],
                   cwd=tmp_path, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert md5f(binp) == case["bin_md5"] and md5f(str(binp) + ".idx") == idx_gpu


def test_icspenc_noise_at_qp1_exceeds_reference_buffer(tmp_path, oracle):
    """Random noise at QP 1 codes to more than width*height bytes per frame — the reference overruns its own buffer there
    (ENC:4874); icspenc retries the call with the worst-case staging buffer and still matches the oracle bit for bit."""
    w, h, n = 64, 48, 3
    rng = np.random.default_rng(99)
    clip = rng.integers(0, 256, size=(n, w * h * 3 // 2)).astype(np.uint8)
    clip.tofile(tmp_path / "noise_x.yuv")
    subprocess.run([ENC, "-i", "noise_x.yuv", "-w", str(w), "-h", str(h), "-n", str(n), "-q", "1", "--intraPeriod", "2", "--quiet"],
                   cwd=tmp_path, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    s = oracle.encode(clip, w, h, 1, 1, 2)
    want = oracle.write_bitstream(s, w, h, 1, 1, 2)
    assert len(want) > n * w * h                                   # the case really is beyond the reference's bound
    assert open(tmp_path / "noise_compCIF_1_1_2.bin", "rb").read() == want
    assert np.array_equal(np.fromfile(tmp_path / "test_yuv.yuv", np.uint8).reshape(n, -1), s.recon.reshape(n, -1))


REF_GPU = os.path.join(ROOT, "oracle", "_ref", "ICSPCodec_gpu")


@pytest.mark.skipif(not os.path.exists(REF_GPU), reason="oracle/_ref/ICSPCodec_gpu not built (oracle/build_ref.sh needs /root/reference)")
@pytest.mark.parametrize("case", [CASES[0], CASES[5], CASES[6], CASES[8], CASES[9]], ids=lambda c: f"{c['kind']}-q{c['qdc']}_{c['qac']}-ip{c['ip']}")
def test_reference_front_end_on_libicspcuda(tmp_path, case):
    """The reference-side binding, compiled (integration/icsp_ref_shim.cpp): the reference's OWN main, option parser, loader,
    makebitstream and checkResultFrames linked against libicspcuda — only single_thread_encoding's per-frame calls are
    replaced.  The reference's own writer must then produce the reference's own .bin and test_yuv.yuv, byte for byte."""
    clip = synth.make_clip(case["kind"], case["nframes"], case["seed"])
    clip.tofile(tmp_path / "clip_cif.yuv")
    n, qdc, qac, ip = case["nframes"], case["qdc"], case["qac"], case["ip"]
    subprocess.run([REF_GPU, "-i", "clip_cif.yuv", "-n", str(n), "--qpdc", str(qdc), "--qpac", str(qac), "--intraPeriod", str(ip)],
                   cwd=tmp_path, check=True, stdout=subprocess.DEVNULL)
    assert md5f(tmp_path / f"clip_compCIF_{qdc}_{qac}_{ip}.bin") == case["bin_md5"]
    assert md5f(tmp_path / "test_yuv.yuv") == case["recon_md5"]


def _gpu_count():
    try:
        return int(subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.count("GPU "))
    except Exception:
        return 0


@pytest.mark.parametrize("gpus", [1, 2])
def test_icspenc_batch(tmp_path, oracle, gpus):
    """icspenc --batch: several streams in one process, one icsp_encode_streams call per wave (the shape bench.py measures),
    tail frames (n % intraPeriod) included, double-buffered file IO; with --gpus 2 the streams are sharded over two devices.
    Every stream's .bin / test_yuv.yuv / .idx equal the single-stream results of the oracle."""
    if gpus > _gpu_count():
        pytest.skip(f"needs {gpus} GPUs")
    n, ip = 12, 5
    kinds = [("highmotion", 301), ("flat", 7), ("akiyo", 9), ("highmotion", 302), ("intra", 12)]
    names = []
    for i, (k, seed) in enumerate(kinds):
        synth.make_clip(k, n, seed).tofile(tmp_path / f"s{i}_cif.yuv")
        names.append(f"s{i}_cif.yuv")
    (tmp_path / "list.txt").write_text("\n".join(names) + "\n")
    subprocess.run([ENC, "--batch", "list.txt", "-n", str(n), "-q", "8", "--intraPeriod", str(ip), "--gpus", str(gpus), "--wave", "2", "--index",
                    "--quiet"], cwd=tmp_path, check=True)
    for i, (k, seed) in enumerate(kinds):
        clip = synth.make_clip(k, n, seed)
        s = oracle.encode(clip, 352, 288, 8, 8, ip)
        assert open(tmp_path / f"s{i}_compCIF_8_8_{ip}.bin", "rb").read() == oracle.write_bitstream(s, 352, 288, 8, 8, ip), f"stream {i}"
        assert np.array_equal(np.fromfile(tmp_path / f"s{i}_test_yuv.yuv", np.uint8).reshape(n, -1), s.recon), f"stream {i}"
        # the row index of the batch path drives the GPU bit reader of icspdec
        subprocess.run([DEC, str(n), f"s{i}_compCIF_8_8_{ip}.bin", "8", "8", str(ip), "--out", f"dec{i}.yuv"], cwd=tmp_path, check=True,
                       stderr=subprocess.DEVNULL)
        ps, _ = oracle.parse_bitstream(open(tmp_path / f"s{i}_compCIF_8_8_{ip}.bin", "rb").read(), n)
        assert np.array_equal(np.fromfile(tmp_path / f"dec{i}.yuv", np.uint8).reshape(n, -1), oracle.decode(ps, 352, 288, 8, 8, ip)), f"stream {i}"
    # icspdec --batch: the same streams in one process, bit reader on the GPU (side-cars present), then on the host (--no-index)
    (tmp_path / "bins.txt").write_text("\n".join(f"s{i}_compCIF_8_8_{ip}.bin" for i in range(len(kinds))) + "\n")
    for extra in ([], ["--no-index"]):
        subprocess.run([DEC, "--batch", "bins.txt", str(n), "--gpus", str(gpus), "--wave", "2"] + extra, cwd=tmp_path, check=True, stderr=subprocess.DEVNULL)
        for i in range(len(kinds)):
            assert md5f(tmp_path / f"s{i}_compCIF_8_8_{ip}.bin.yuv") == md5f(tmp_path / f"dec{i}.yuv"), f"stream {i} {extra}"
            os.remove(tmp_path / f"s{i}_compCIF_8_8_{ip}.bin.yuv")


def test_icspenc_two_gpus_single_stream(tmp_path):
    """icspenc --gpus 2 on one sequence: GOPs sharded over two devices, bit strings concatenated on the host (needs 2 GPUs)."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    case = CASES[0]
    clip = synth.make_clip(case["kind"], case["nframes"], case["seed"])
    clip.tofile(tmp_path / "clip_cif.yuv")
    n, qdc, qac = case["nframes"], case["qdc"], case["qac"]
    subprocess.run([ENC, "-i", "clip_cif.yuv", "-n", str(n), "--qpdc", str(qdc), "--qpac", str(qac), "--intraPeriod", "3", "--gpus", "2", "--quiet"],
                   cwd=tmp_path, check=True)
    one = tmp_path / "one"
    one.mkdir()
    clip.tofile(one / "clip_cif.yuv")
    subprocess.run([ENC, "-i", "clip_cif.yuv", "-n", str(n), "--qpdc", str(qdc), "--qpac", str(qac), "--intraPeriod", "3", "--gpus", "1", "--quiet"],
                   cwd=one, check=True)
    assert md5f(tmp_path / f"clip_compCIF_{qdc}_{qac}_3.bin") == md5f(one / f"clip_compCIF_{qdc}_{qac}_3.bin")
    assert md5f(tmp_path / "test_yuv.yuv") == md5f(one / "test_yuv.yuv")
