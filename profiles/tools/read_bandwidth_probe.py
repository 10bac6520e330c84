"""Read-only HBM bandwidth of PyTorch's own reductions over the K1 input size (1.78 GB), as a
yardstick for the K1 streaming pass (which reads the same bytes and also finds the peaks)."""
import torch, time
x = torch.empty(64*17*640*640, dtype=torch.float32, device='cuda').uniform_()
for name, fn in (('sum', lambda: x.sum()), ('max', lambda: x.max()), ('amax_f4', lambda: x.view(-1, 4).amax()), ('count_nonzero', lambda: torch.count_nonzero(x > 2.0))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(name, round(ms, 4), 'ms', round(x.numel() * 4 / ms / 1e6, 1), 'GB/s')
