"""One ConvNeXt pw1-shaped GEMM (M x 4096 x 512, fp16 operands, bias + GELU -> fp16) a few times: for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moss_ttsd_b200 import ops
M = int(sys.argv[1]) if len(sys.argv) > 1 else 192000
x = torch.randn(M, 512, device="cuda").half()
w = (torch.randn(4096, 512, device="cuda") * 512 ** -0.5).half()
b = torch.randn(4096, device="cuda") * 0.1
o = torch.empty(M, 4096, device="cuda", dtype=torch.float16)
for _ in range(4):
    ops.gemm(x, w, out=o, bias=b, gelu=True)
torch.cuda.synchronize()
print("ok")
