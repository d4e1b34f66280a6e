"""Large-M (prefill-shaped) bf16 GEMM throughput of mtts_gemm next to torch.matmul (cuBLAS), CUDA events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moss_ttsd_b200 import ops

def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters

M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
for (N, K) in ((4096, 2048), (2048, 2048), (12288, 2048), (2048, 6144)):
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    ms = timeit(lambda: ops.gemm(x, w, out=out))
    ms_t = timeit(lambda: torch.matmul(x, w.t(), out=out))
    print(f"M={M} N={N:6d} K={K:5d}  mtts {2.0*M*N*K/ms/1e9:7.1f} TF/s ({ms:.3f} ms)   cuBLAS {2.0*M*N*K/ms_t/1e9:7.1f} TF/s", flush=True)

torch.backends.cuda.matmul.allow_tf32 = True
for (M2, N, K) in ((192000, 4096, 512), (192000, 512, 4096), (96000, 3072, 768), (96000, 768, 3072), (96000, 2304, 768)):
    x = torch.randn(M2, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * 0.02
    out = torch.empty(M2, N, device="cuda")
    ms = timeit(lambda: ops.gemm(x, w, out=out))
    ms_t = timeit(lambda: torch.matmul(x, w.t(), out=out))
    print(f"TF32 M={M2} N={N:6d} K={K:5d}  mtts {2.0*M2*N*K/ms/1e9:7.1f} TF/s ({ms:.3f} ms)   cuBLAS(tf32) {2.0*M2*N*K/ms_t/1e9:7.1f} TF/s", flush=True)
