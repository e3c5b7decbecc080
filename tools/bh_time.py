import torch, numpy as np, sys
sys.path.insert(0, '/root/repo')
from diffxpy_b200.testing import correction
import sys as _s
p = torch.rand(int(_s.argv[1]) if len(_s.argv) > 1 else 20000, dtype=torch.float64, device='cuda')
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("kernel path ms", t(lambda: correction.bh_device(p)))
correction.BH_KERNEL_MAX = 0
print("sort path ms", t(lambda: correction.bh_device(p)))
