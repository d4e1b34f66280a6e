"""Decode attention kernel alone: 28 launches over 28 layers' KV in a CUDA graph, sweep of nsplit."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moss_ttsd_b200 import _lib, ops
from moss_ttsd_b200.lm_engine import KVCache, LMShape
ops.ensure_init()
L = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctxs = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["460", "800"])]
shape = LMShape()
cache = KVCache(shape, B, max(1024, max(ctxs) + 8), "cuda")
cache.k.normal_(); cache.v.normal_()
q = torch.randn(B, 16 * 128, device="cuda").to(torch.bfloat16)
out = torch.empty_like(q)
for ctx in ctxs:
    pos = torch.full((B,), ctx - 1, dtype=torch.int32, device="cuda")
    for nsplit in (1, 2, 3, 4, 6, 8):
        n = L.mtts_gqa_attention_workspace_bytes(B, 8, 2, 1, nsplit)
        ws = torch.zeros(n, dtype=torch.uint8, device="cuda")
        def run():
            for l in range(28):
                _lib.check(L.mtts_gqa_attention(q.data_ptr(), cache.k[l].data_ptr(), cache.v[l].data_ptr(), None, cache.max_pages,
                                                cache.page_size, None, None, None, pos.data_ptr(), out.data_ptr(), B, 1, 16, 8, 128,
                                                nsplit, ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
        run(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run()
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10): g.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 280
        gb = B * ctx * 4096 / us / 1e3
        print(f"B={B} ctx={ctx} nsplit={nsplit}: {us:6.2f} us  {gb:7.1f} GB/s ({gb/65.418:4.1f}%)", flush=True)
