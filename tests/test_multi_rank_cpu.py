"""world_size-2 gloo test (CPU) of the N>1 host logic: each rank takes its contiguous GOP shard, produces the SoA
syntax for it (here with the oracle standing in for the GPU core — this is the checker side, the product path is
exercised by the -m gpu tests), rank 0 gathers the shards in order and the PRODUCT bitstream writer must reproduce the
single-process reference bitstream byte for byte.  Also covers max-over-ranks timing aggregation used by bench.py."""
import hashlib
import json
import os
import socket

import numpy as np
import pytest

from icspcodec_b200.sharding import gop_shards, stream_shards

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_shard_rules():
    s = gop_shards(300, 10, 8)
    assert sum(x.n_frames for x in s) == 300 and [x.device for x in s] == list(range(8))
    assert all(a.first_frame + a.n_frames == b.first_frame for a, b in zip(s, s[1:]))
    s = gop_shards(37, 10, 2)            # tail GOP of 7 frames on the last device
    assert [(x.device, x.first_frame, x.n_gops, x.gop_len) for x in s] == [(0, 0, 1, 10), (1, 10, 2, 10), (1, 30, 1, 7)]
    s = gop_shards(5, 0, 4)              # all intra: every frame its own GOP
    assert sum(x.n_frames for x in s) == 5 and all(x.gop_len == 1 for x in s)
    assert gop_shards(3, 10, 4) == [type(s[0])(3, 0, 1, 3)]
    assert [len(r) for r in stream_shards(64, 8)] == [8] * 8 and [len(r) for r in stream_shards(3, 2)] == [1, 2]


ICSPENC = os.path.join(os.path.dirname(GOLD), "..", "icspcodec_b200", "host", "icspenc")


@pytest.mark.parametrize("n,ip,gpus", [(300, 10, 8), (300, 10, 1), (37, 10, 2), (5, 0, 4), (3, 10, 4), (299, 7, 3), (64, 63, 8)])
def test_icspenc_plan_is_the_same_gop_rule(n, ip, gpus):
    """The product CLI (C++) and the Python mirror must shard identically: `icspenc --plan` prints its plan without a device."""
    import re
    import subprocess
    out = subprocess.run([ICSPENC, "-i", "x_cif.yuv", "-n", str(n), "-q", "8", "--intraPeriod", str(ip), "--gpus", str(gpus), "--plan"],
                         capture_output=True, text=True, check=True).stdout
    got = [tuple(int(v) for v in re.findall(r"=(\d+)", line)) for line in out.splitlines() if line.startswith("gops ")]
    assert got == [(x.device, x.first_frame, x.n_gops, x.gop_len) for x in gop_shards(n, ip, gpus)]


@pytest.mark.parametrize("streams,gpus", [(64, 8), (3, 2), (5, 4), (2, 8)])
def test_icspenc_plan_is_the_same_stream_rule(tmp_path, streams, gpus):
    import re
    import subprocess
    lst = tmp_path / "list.txt"
    lst.write_text("".join(f"s{i}_cif.yuv\n" for i in range(streams)))
    out = subprocess.run([ICSPENC, "--batch", str(lst), "-n", "30", "-q", "8", "--intraPeriod", "10", "--gpus", str(gpus), "--plan"],
                         capture_output=True, text=True, check=True).stdout
    got = [tuple(int(v) for v in re.findall(r"=(\d+)", line)) for line in out.splitlines() if line.startswith("streams ")]
    g = min(gpus, streams)      # the CLI never opens more devices than it has streams
    assert got == [(d, r.start, r.stop) for d, r in enumerate(stream_shards(streams, g))]


def _worker(rank, world, port, case, out_q):
    import torch
    import torch.distributed as dist
    from icspcodec_b200 import hostlib, synth
    from oracle import oracle_py as O
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n, qdc, qac, ip = case["nframes"], case["qdc"], case["qac"], case["ip"]
    clip = synth.make_clip(case["kind"], n, case["seed"])
    parts = {}
    for sh in gop_shards(n, ip, world):
        if sh.device != rank:
            continue
        s = O.encode(clip[sh.first_frame: sh.first_frame + sh.n_frames], 352, 288, qdc, qac, sh.gop_len if sh.gop_len > 1 else ip)
        parts[sh.first_frame] = {k: getattr(s, k) for k in ("levels", "acflag", "mpm", "ipm", "mvd")}
    gathered = [None] * world
    dist.all_gather_object(gathered, parts)
    # max-over-ranks timing aggregation as in bench.py
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        merged = {}
        for g in gathered:
            merged.update(g)
        cat = {k: np.concatenate([merged[f][k] for f in sorted(merged)], axis=0) for k in ("levels", "acflag", "mpm", "ipm", "mvd")}
        bs = hostlib.write_stream(cat["levels"], cat["acflag"], cat["mpm"], cat["ipm"], cat["mvd"], 352, 288, qdc, qac, ip, 2)
        out_q.put((hashlib.md5(bs).hexdigest(), float(t[0])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("idx", [0, 8])
def test_two_rank_gop_sharding_reproduces_reference_bitstream(idx):
    import torch.multiprocessing as mp
    from icspcodec_b200 import build
    from oracle import oracle_py as O
    build.build()
    O.build()
    case = json.load(open(os.path.join(GOLD, "ref_cases.json")))[idx]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    md5, tmax = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert md5 == case["bin_md5"]
    assert tmax == 11.0
