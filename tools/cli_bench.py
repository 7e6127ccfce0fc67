"""End-to-end throughput of the PRODUCT front ends (icspenc / icspdec), files in and files out, on the benchmark shape:
64 CIF streams x 300 frames (BASELINE configs[3]) through `icspenc --batch`, the same through `icspdec --batch`, and
configs[0] (one 300-frame stream) through plain `icspenc -i`.  Files live in /dev/shm so that the numbers are the
tools' own (page-cache reads / writes, no disk).   python tools/cli_bench.py [streams] [gpus]  -> JSON on stdout"""
import json, os, shutil, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiprocessing as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENC = os.path.join(ROOT, "icspcodec_b200", "host", "icspenc")
DEC = os.path.join(ROOT, "icspcodec_b200", "host", "icspdec")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
G = int(sys.argv[2]) if len(sys.argv) > 2 else 1
N = 300
work = "/dev/shm/icsp_cli"


def gen(i):
    from icspcodec_b200 import synth
    synth.make_clip("highmotion", N, 1000 + i).tofile(os.path.join(work, f"s{i:02d}_cif.yuv"))


def timed(cmd, reps=2):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        subprocess.run(cmd, cwd=work, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        best = min(best, time.perf_counter() - t0)
    return best


if __name__ == "__main__":
    shutil.rmtree(work, ignore_errors=True)
    os.makedirs(work)
    with mp.Pool(min(S, os.cpu_count() or 4)) as pool:
        pool.map(gen, range(S))
    from icspcodec_b200 import synth
    synth.make_clip("akiyo", N, 20261017).tofile(os.path.join(work, "akiyo_cif.yuv"))
    open(os.path.join(work, "list.txt"), "w").write("".join(f"s{i:02d}_cif.yuv\n" for i in range(S)))
    open(os.path.join(work, "bins.txt"), "w").write("".join(f"s{i:02d}_compCIF_8_8_10.bin\n" for i in range(S)))
    out = {"streams": S, "frames_per_stream": N, "gpus": G, "files": "/dev/shm (page cache)"}
    base = [ENC, "--batch", "list.txt", "-n", str(N), "-q", "8", "--intraPeriod", "10", "--gpus", str(G), "--quiet"]
    t = timed(base + ["--index"])
    out["icspenc_batch_fps"] = round(S * N / t, 1)
    out["icspenc_batch_s"] = round(t, 3)
    t = timed(base + ["--no-recon"])
    out["icspenc_batch_no_recon_fps"] = round(S * N / t, 1)
    t = timed([DEC, "--batch", "bins.txt", str(N), "--gpus", str(G)])
    out["icspdec_batch_gpu_reader_fps"] = round(S * N / t, 1)
    t = timed([DEC, "--batch", "bins.txt", str(N), "--gpus", str(G), "--no-index"], reps=1)
    out["icspdec_batch_host_parser_fps"] = round(S * N / t, 1)
    t = timed([ENC, "-i", "akiyo_cif.yuv", "-n", str(N), "-q", "8", "--intraPeriod", "10", "--quiet"], reps=3)
    out["icspenc_config0_single_stream_fps"] = round(N / t, 1)
    out["icspenc_config0_single_stream_s"] = round(t, 3)
    t = timed([ENC, "-i", "akiyo_cif.yuv", "-n", str(N), "-q", "8", "--intraPeriod", "10", "--quiet", "--index"], reps=1)
    t = timed([DEC, str(N), "akiyo_compCIF_8_8_10.bin", "8", "8", "10"], reps=3)
    out["icspdec_config4_single_stream_fps"] = round(N / t, 1)
    print(json.dumps(out))
    shutil.rmtree(work, ignore_errors=True)
