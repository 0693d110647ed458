import torch, time
torch.cuda.init()
for mb, n in ((3.1104, 32), (99.5, 1), (400, 1)):
    nbytes = int(mb * 1e6)
    h = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(n)]
    d = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in range(n)]
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            for a, b in zip(h, d): b.copy_(a, non_blocking=True)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(10):
            for a, b in zip(h, d): b.copy_(a, non_blocking=True)
        e1.record(s); s.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("H2D %d x %.1f MB: %.3f ms  %.1f GB/s" % (n, mb, ms, n * nbytes / ms / 1e6))
    with torch.cuda.stream(s):
        e0.record(s)
        for _ in range(10):
            for a, b in zip(h, d): a.copy_(b, non_blocking=True)
        e1.record(s); s.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("D2H %d x %.1f MB: %.3f ms  %.1f GB/s" % (n, mb, ms, n * nbytes / ms / 1e6))
