import torch, sys
sys.path.insert(0, "/root/repo")
x = torch.empty(1865 * 250000, dtype=torch.float32, device="cuda")   # 1.865 GB
y = torch.empty_like(x)
def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: x.fill_(1.5)); print("fill_  %.4f ms  %.0f GB/s written" % (ms, x.numel() * 4 / ms / 1e6))
ms = t(lambda: x.zero_()); print("zero_  %.4f ms  %.0f GB/s written" % (ms, x.numel() * 4 / ms / 1e6))
ms = t(lambda: y.copy_(x)); print("copy_  %.4f ms  %.0f GB/s read+written" % (ms, 2 * x.numel() * 4 / ms / 1e6))
