"""HBM write / read / copy bandwidth with plain torch ops (context for the store-bound conv layers)."""
import torch
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: x.fill_(3)); print(f"fill  1 GiB: {ms:.3f} ms  {1.0737/ms*1e3:.0f} GB/s write")
ms = t(lambda: y.copy_(x)); print(f"copy  1 GiB: {ms:.3f} ms  {2*1.0737/ms*1e3:.0f} GB/s read+write")
xf = x.view(torch.float32)
ms = t(lambda: xf.sum()); print(f"sum   1 GiB: {ms:.3f} ms  {1.0737/ms*1e3:.0f} GB/s read")
