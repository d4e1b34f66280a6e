"""Per-launch time of weight-streaming GEMMs inside a CUDA graph: 28 launches over 28 distinct weight matrices
(so nothing is L2-resident), for each decode shape and batch size. Prints us/launch and GB/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moss_ttsd_b200 import ops

ops.ensure_init()
Ms = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "1,16,64".split(","))]
shapes = [(4096, 2048), (2048, 2048), (12288, 2048), (2048, 6144)]
L = 28
for M in Ms:
    for (N, K) in shapes:
        ws = [(torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16) for _ in range(L)]
        x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
        gws = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
        def chain():
            for w in ws:
                ops.gemm(x, w, out=out, workspace=gws)
        chain(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            chain()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (10 * L)
        gb = (N * K * 2 + M * K * 2 + M * N * 2) / us / 1e3
        print(f"M={M:3d} N={N:6d} K={K:5d}  {us:7.2f} us/launch  {gb:7.1f} GB/s  ({gb/6541.8*100:4.1f}% of measured HBM peak)", flush=True)
