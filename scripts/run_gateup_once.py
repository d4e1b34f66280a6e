"""A few gate/up projections (stream-K pair GEMM + SwiGLU epilogue) at batch M over distinct weights: for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moss_ttsd_b200 import _lib, ops
L = _lib.load()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H, I = 2048, 6144
g = torch.Generator(device="cuda").manual_seed(0)
ws_ = [(torch.randn(2 * I, H, device="cuda", generator=g) * 0.02).to(torch.bfloat16) for _ in range(4)]
x = torch.randn(M, H, device="cuda").to(torch.bfloat16)
o_i = torch.empty(M, I, device="cuda", dtype=torch.bfloat16)
gws = torch.zeros(L.mtts_gemm_workspace_bytes(M, 2 * I, H, 0), dtype=torch.uint8, device="cuda")
for _ in range(2):
    for w in ws_:
        ops.gemm(x, w, out=o_i, swiglu=True, workspace=gws)
torch.cuda.synchronize()
print("ok")
