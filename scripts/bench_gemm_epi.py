"""Epilogue cost of the large-M GEMM on a ConvNeXt pw1-shaped problem (M x 4096 x 512) and pw2 (M x 512 x 4096)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moss_ttsd_b200 import ops

M = int(sys.argv[1]) if len(sys.argv) > 1 else 192000
torch.manual_seed(0)

def t(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for (N, K) in ((4096, 512), (512, 4096), (3072, 768), (768, 3072)):
    x32 = torch.randn(M, K, device="cuda")
    w32 = torch.randn(N, K, device="cuda") * K ** -0.5
    b = torch.randn(N, device="cuda") * 0.1
    x16, w16 = x32.half(), w32.half()
    xb, wb = x32.bfloat16(), w32.bfloat16()
    o32 = torch.empty(M, N, device="cuda")
    o16 = torch.empty(M, N, device="cuda", dtype=torch.float16)
    ob = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    fl = 2.0 * M * N * K / 1e9
    rows = {
        "f16 in, f16 out, bias+gelu": lambda: ops.gemm(x16, w16, out=o16, bias=b, gelu=True),
        "f16 in, f32 out, bias+gelu": lambda: ops.gemm(x16, w16, out=o32, bias=b, gelu=True),
        "f16 in, f32 out, bias": lambda: ops.gemm(x16, w16, out=o32, bias=b),
        "f16 in, f16 out, bias": lambda: ops.gemm(x16, w16, out=o16, bias=b),
        "f16 in, f32 out, plain": lambda: ops.gemm(x16, w16, out=o32),
        "bf16 in, bf16 out, plain": lambda: ops.gemm(xb, wb, out=ob),
        "tf32 in, f32 out, bias+gelu": lambda: ops.gemm(x32, w32, out=o32, bias=b, gelu=True),
        "tf32 in, f32 out, plain": lambda: ops.gemm(x32, w32, out=o32),
    }
    for name, fn in rows.items():
        ms = t(fn)
        print(f"M={M} N={N} K={K}  {name:30s} {ms * 1e3:8.1f} us  {fl / ms:8.1f} TFLOP/s", flush=True)
