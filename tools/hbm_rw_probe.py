import torch
x = torch.empty(4 * 1024**3 // 4, device="cuda"); y = torch.empty_like(x)
def t(f, n=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
ms = t(lambda: x.zero_()); print(f"write-only 4 GiB: {ms:.3f} ms {4*1024**3/ms/1e6:.0f} GB/s")
ms = t(lambda: y.copy_(x)); print(f"copy 4 GiB: {ms:.3f} ms {2*4*1024**3/ms/1e6:.0f} GB/s")
ms = t(lambda: x.sum()); print(f"read-only 4 GiB: {ms:.3f} ms {4*1024**3/ms/1e6:.0f} GB/s")
