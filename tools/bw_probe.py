"""HBM bandwidth probe: read-only, write-only and copy streams (torch kernels, CUDA events)."""
import torch
dev = torch.device("cuda:0")
n = 1 << 28                      # 1 GiB of fp32
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
gb = n * 4 / 1e9
t = timeit(lambda: a.zero_());            print("write-only (fill)   %.0f GB/s" % (gb / t))
t = timeit(lambda: a.fill_(1.5));         print("write-only (fill v) %.0f GB/s" % (gb / t))
t = timeit(lambda: torch.sum(a));         print("read-only (sum)     %.0f GB/s" % (gb / t))
t = timeit(lambda: b.copy_(a));           print("copy (r+w)          %.0f GB/s" % (2 * gb / t))
t = timeit(lambda: torch.add(a, 1.0, out=b)); print("add (r+w)           %.0f GB/s" % (2 * gb / t))
