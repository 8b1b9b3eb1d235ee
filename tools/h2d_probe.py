"""PCIe probe: rate of pitched (column-chunk) host->device copies from a pinned row-major matrix, alone and with a
device->host copy running next to it: python tools/h2d_probe.py"""
import time, torch
m = n = 32768
host = torch.empty((m, n), dtype=torch.float32).pin_memory()
dev = torch.empty((m, n), dtype=torch.float32, device="cuda")
back = torch.empty((m, n), dtype=torch.float32).pin_memory()
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
def run(width, with_d2h):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(s_in):
        for c in range(0, n, width):
            dev[:, c:c + width].copy_(host[:, c:c + width], non_blocking=True)
    if with_d2h:
        with torch.cuda.stream(s_out):
            for c in range(0, n, 1024):
                back[:, c:c + 1024].copy_(dev[:, c:c + 1024], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"H2D width {width:6d} cols ({width * 4 // 1024:4d} KB rows){' + D2H 1024-col blocks' if with_d2h else '':24s}: {dt * 1e3:7.1f} ms  {m * n * 4 / dt / 1e9:5.1f} GB/s", flush=True)
def run_d2h(width):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(s_out):
        for c in range(0, n, width):
            back[:, c:c + width].copy_(dev[:, c:c + width], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"D2H width {width:6d} cols: {dt * 1e3:7.1f} ms  {m * n * 4 / dt / 1e9:5.1f} GB/s", flush=True)
for w in (n, 8192, 4096, 2048, 1024, 512):
    run(w, False)
for w in (n, 4096, 1024):
    run_d2h(w)
for w in (n, 4096, 1024):
    run(w, True)
