"""PCIe probe: pinned H2D / D2H bandwidth with 1 or 2 concurrent copy streams per direction, alone and bidirectional."""
import time
import torch
GB = 1 << 30
n = 2 * GB
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
streams = [torch.cuda.Stream() for _ in range(4)]

def run(h2d_parts, d2h_parts, reps=3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in range(h2d_parts):
            a, b = n * i // h2d_parts, n * (i + 1) // h2d_parts
            with torch.cuda.stream(streams[i]):
                d_in[a:b].copy_(h_in[a:b], non_blocking=True)
        for i in range(d2h_parts):
            a, b = n * i // d2h_parts, n * (i + 1) // d2h_parts
            with torch.cuda.stream(streams[2 + i]):
                h_out[a:b].copy_(d_out[a:b], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return dt

run(1, 1, 1)
for h, d in ((1, 0), (2, 0), (0, 1), (0, 2), (1, 1), (2, 2)):
    dt = run(h, d)
    print(f"h2d streams {h} d2h streams {d}: {dt*1e3:7.1f} ms  h2d {n*(h>0)/dt/1e9:6.1f} GB/s  d2h {n*(d>0)/dt/1e9:6.1f} GB/s  total {n*((h>0)+(d>0))/dt/1e9:6.1f} GB/s")
