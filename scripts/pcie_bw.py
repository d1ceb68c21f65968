import torch
n = 192_000_000
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n // 2, dtype=torch.uint8, pin_memory=True)
d2 = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, both in (("H2D alone", False), ("H2D with concurrent D2H", True)):
    for _ in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s1):
            e0.record()
            for k in range(8):
                d[k * n // 8:(k + 1) * n // 8].copy_(h[k * n // 8:(k + 1) * n // 8], non_blocking=True)
            e1.record()
        if both:
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
    print("%s: %.2f ms for %d MB -> %.1f GB/s" % (name, e0.elapsed_time(e1), n // 1000000, n / e0.elapsed_time(e1) / 1e6))
