import torch, time
x = torch.empty(320_000_000, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(x, non_blocking=True)), ("D2H", lambda: x.copy_(d, non_blocking=True))):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 5
    print("%s pinned 320MB: %.2f ms = %.1f GB/s" % (name, dt * 1e3, 0.32 / dt))
# both directions at once on two streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
y = torch.empty(177_000_000, dtype=torch.uint8).pin_memory(); e = torch.empty_like(y, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): y.copy_(e, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print("H2D 320MB + D2H 177MB concurrently: %.2f ms" % (dt * 1e3))
