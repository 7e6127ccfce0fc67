"""N-rank PCIe probe: every rank (one process per GPU) moves pinned 1 GiB buffers H2D and D2H at the same time, all ranks
started together; prints per-rank and AGGREGATE GB/s — the host-side ceiling of bench.py's `e2e` at N GPUs.
   python tools/pcie_probe_multi.py [N] [wc]      (wc: write-combined upload buffers like icsp_host_alloc_upload)"""
import multiprocessing as mp
import sys
import time

GB = 1 << 30


def worker(rank, n, bar, q, mode):
    import torch
    torch.cuda.set_device(rank)
    size = GB
    h_in = torch.empty(size, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(size, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(size, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(size, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    reps = 4
    res = {}
    for what in ("h2d", "d2h", "both"):
        torch.cuda.synchronize()
        bar.wait()
        t0 = time.perf_counter()
        for _ in range(reps):
            if what in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if what in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        bar.wait()
        res[what] = size * (2 if what == "both" else 1) / dt / 1e9
    q.put((rank, res))


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    mp.set_start_method("spawn")
    bar = mp.Barrier(n)
    q = mp.Queue()
    ps = [mp.Process(target=worker, args=(r, n, bar, q, "")) for r in range(n)]
    for p in ps: p.start()
    out = sorted(q.get() for _ in range(n))
    for p in ps: p.join()
    for what in ("h2d", "d2h", "both"):
        vals = [r[what] for _, r in out]
        print(f"{n} ranks {what:5s}: per rank min {min(vals):6.1f} max {max(vals):6.1f} GB/s, aggregate {sum(vals):7.1f} GB/s")
