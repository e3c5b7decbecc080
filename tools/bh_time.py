import sys
import numpy as np
import torch
sys.path.insert(0, '/root/repo')
from diffxpy_b200.testing import correction

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000


def t(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


cases = {"uniform": torch.rand(n, dtype=torch.float64, device='cuda'),
         "10% signal": torch.cat([torch.rand(n - n // 10, dtype=torch.float64, device='cuda'),
                                  10.0 ** (-30 * torch.rand(n // 10, dtype=torch.float64, device='cuda'))]),
         "all equal": torch.full((n,), 0.5, dtype=torch.float64, device='cuda')}
for name, p in cases.items():
    correction.BH_KERNEL_MAX = 1 << 24
    a = t(lambda: correction.bh_device(p))
    correction.BH_KERNEL_MAX = 0
    b = t(lambda: correction.bh_device(p))
    print("%-12s n %d: bucket kernels %.4f ms, sort path %.4f ms" % (name, n, a, b))
