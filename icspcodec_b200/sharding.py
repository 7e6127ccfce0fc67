"""GOP sharding across GPUs: closed GOPs share no state, so a sequence (or a batch of streams) is
split into contiguous GOP ranges, one per device, with no data-path collective; the host concatenates the per-GOP
results in order.  Python mirror of the two rules in host/icspenc.cpp (`gop_shards`, `stream_shard`); used by bench.py's
strong-scaling leg and the 2-rank gloo test, and held to the CLI's own plan (`icspenc --plan`) by tests/test_multi_rank_cpu.py."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Shard:
    device: int
    first_frame: int
    n_gops: int
    gop_len: int

    @property
    def n_frames(self) -> int:
        return self.n_gops * self.gop_len


def gop_shards(n_frames: int, intra_period: int, n_devices: int) -> list[Shard]:
    """Full GOPs are split contiguously (device d gets GOPs [full*d/N, full*(d+1)/N)); the tail GOP
    (n_frames % intra_period frames, dropped by the reference's MT mode, ICSP_thread.cpp:43) goes to the last device.
    intra_period 0 = all intra (every frame is its own GOP)."""
    gop = 1 if intra_period == 0 else intra_period
    full, tail = divmod(n_frames, gop)
    n_devices = max(1, n_devices)
    out = []
    for d in range(n_devices):
        g0, g1 = full * d // n_devices, full * (d + 1) // n_devices
        if g1 > g0:
            out.append(Shard(d, g0 * gop, g1 - g0, gop))
    if tail:
        out.append(Shard(n_devices - 1, full * gop, 1, tail))
    return out


def stream_shards(n_streams: int, n_devices: int) -> list[range]:
    """Whole independent streams per device (the 64-stream batch of BASELINE configs[3])."""
    return [range(n_streams * d // n_devices, n_streams * (d + 1) // n_devices) for d in range(n_devices)]
