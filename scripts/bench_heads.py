import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moss_ttsd_b200 import ops
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
N, K = 159928, 2048
w = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
for M in (1, 16, 64, 128, 256):
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    ms = timeit(lambda: ops.gemm(x, w, out=out))
    print(f"heads M={M}: {ms*1e3:.1f} us  {N*K*2/ms/1e6:.0f} GB/s")
