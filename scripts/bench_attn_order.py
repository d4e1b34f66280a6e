"""Decode attention at batch 256 with the bench's ragged contexts: does the ORDER of the rows matter (tail of the last
wave of CTAs)? Same contexts in arrival order, longest first, shortest first, and all equal to the mean (no raggedness)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from moss_ttsd_b200 import _lib, ops
from moss_ttsd_b200.lm_engine import KVCache, LMShape
ops.ensure_init()
L = _lib.load()
B = 256
rng = np.random.default_rng(1000)
for lo, hi, tag in ((64, 512, "whole job (text 64..512)"), (363, 512, "longest third"), (64, 213, "shortest third")):
    ctx = rng.integers(lo, hi + 1, B) + 250 + 7 + 187
    shape = LMShape()
    cache = KVCache(shape, B, 1032, "cuda")
    cache.k.normal_(); cache.v.normal_()
    q = torch.randn(B, 16 * 128, device="cuda").to(torch.bfloat16)
    out = torch.empty_like(q)
    for name, c in (("arrival order", ctx), ("longest first", np.sort(ctx)[::-1].copy()), ("shortest first", np.sort(ctx)),
                    ("all = mean", np.full(B, int(ctx.mean())))):
        pos = torch.from_numpy((c - 1).astype(np.int32)).cuda()
        def run():
            for l in range(28):
                _lib.check(L.mtts_gqa_attention(q.data_ptr(), cache.k[l].data_ptr(), cache.v[l].data_ptr(), None, cache.max_pages,
                                                cache.page_size, None, None, None, pos.data_ptr(), out.data_ptr(), B, 1, 16, 8, 128,
                                                1, None, 0, _lib.stream_ptr()))
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
        gb = float(c.sum()) * 4096 / us / 1e3
        print(f"{tag:28s} {name:15s}: {us:7.2f} us  {gb:7.1f} GB/s", flush=True)
    del cache
