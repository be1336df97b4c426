#!/usr/bin/env python
"""Read-only HBM bandwidth probe (what a pure streaming read of the 8 GB capture can reach on this GPU),
to put the fused kernel's achieved GB/s next to both the driver's copy peak and a read-only ceiling."""
import torch
x = torch.empty(10 ** 9, 2, dtype=torch.float32, device="cuda").normal_()
for name, fn in (("torch.sum(f32)", lambda: x.sum()), ("torch.max(f32)", lambda: x.max()),
                 ("torch.sum(view f64)", lambda: x.view(torch.float64).sum()),
                 ("copy_ (read+write)", None)):
    if fn is None:
        y = torch.empty_like(x)
        fn = lambda: y.copy_(x)
        nbytes = 2 * x.numel() * 4
    else:
        nbytes = x.numel() * 4
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("%-22s %.3f ms  %.1f GB/s" % (name, best, nbytes / best / 1e6))
