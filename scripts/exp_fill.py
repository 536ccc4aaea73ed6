"""GPU experiment: pure-write HBM bandwidth (torch fill_ / zero_) vs the read+write copy peak of MEASURED_PEAKS.json."""
import torch
x = torch.empty(2 * 1024 ** 3, dtype=torch.float32, device="cuda")   # 8 GiB >> L2
for name, fn in [("fill_", lambda: x.fill_(1.0)), ("zero_", lambda: x.zero_())]:
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("%s: %.3f ms -> %.0f GB/s written" % (name, ms, x.numel() * 4 / ms / 1e6))
y = torch.empty_like(x)
for _ in range(2): y.copy_(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): y.copy_(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("copy_: %.3f ms -> %.0f GB/s read+written" % (ms, 2 * x.numel() * 4 / ms / 1e6))
