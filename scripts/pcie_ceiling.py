"""Aggregate PCIe ceiling of the box for the e2e path: the bytes config 2 moves per step and GPU
(340 MB host-to-device + 177 MB device-to-host, page-locked, both directions at once), on
1 / 2 / 4 / 8 GPUs concurrently (one process, one stream pair per device).  Prints JSON lines."""
import json
import sys
import time

import torch

H2D, D2H = 340_000_000, 177_000_000
n_all = torch.cuda.device_count()
bufs = []
for d in range(n_all):
    with torch.cuda.device(d):
        bufs.append((torch.empty(H2D, dtype=torch.uint8).pin_memory(), torch.empty(H2D, dtype=torch.uint8, device="cuda"),
                     torch.empty(D2H, dtype=torch.uint8).pin_memory(), torch.empty(D2H, dtype=torch.uint8, device="cuda"),
                     torch.cuda.Stream(), torch.cuda.Stream()))
for n in (1, 2, 4, 8):
    if n > n_all:
        break
    def step():
        for d in range(n):
            hx, dx, hy, dy, s1, s2 = bufs[d]
            with torch.cuda.device(d):
                with torch.cuda.stream(s1):
                    dx.copy_(hx, non_blocking=True)
                with torch.cuda.stream(s2):
                    hy.copy_(dy, non_blocking=True)
    def sync():
        for d in range(n):
            torch.cuda.synchronize(d)
    for _ in range(2):
        step()
    sync()
    t = time.perf_counter()
    K = 5
    for _ in range(K):
        step()
    sync()
    ms = (time.perf_counter() - t) / K * 1e3
    print(json.dumps({"gpus": n, "ms_per_step": ms, "h2d_gbs_per_gpu": H2D / ms / 1e6, "d2h_gbs_per_gpu": D2H / ms / 1e6,
                      "aggregate_gbs": n * (H2D + D2H) / ms / 1e6, "pairs_per_s_ceiling": n * 1e6 / (ms / 1e3)}))
    sys.stdout.flush()
