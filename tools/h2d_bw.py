import torch, time
x = torch.empty(267_000_000, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
for n in (2_000_000, 16_000_000, 33_000_000, 267_000_000):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        d[:n].copy_(x[:n], non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print("H2D %d MB: %.1f GB/s" % (n // 1000000, 5 * n / e0.elapsed_time(e1) / 1e6))
