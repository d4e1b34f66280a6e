"""Prefill attention alone: tcgen05 kernel (128-row tiles) against the mma.sync flash kernel (64-row tiles) on a packed
batch of B sequences of T rows, v0.5 head layout (16 q heads, 8 kv heads, head_dim 128). us per launch, TFLOP/s (causal)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from moss_ttsd_b200 import _lib, ops
ops.ensure_init()
L = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 545
Hq, Hkv, D, page = 16, 8, 128, 64
max_pages = (T + page - 1) // page
num_pages = B * max_pages
k_pool = torch.randn((num_pages, Hkv, page, D), device="cuda").to(torch.bfloat16)
v_pool = torch.randn((num_pages, Hkv, page, D), device="cuda").to(torch.bfloat16)
R = B * T
q = torch.randn((R, Hq * D), device="cuda").to(torch.bfloat16)
out = torch.empty_like(q)
pos = torch.arange(T, dtype=torch.int32).repeat(B).cuda()
seq = torch.arange(B, dtype=torch.int32).repeat_interleave(T).cuda()


def tiles(n):
    r0, nr = [], []
    for b in range(B):
        for t0 in range(0, T, n):
            r0.append(b * T + t0); nr.append(min(n, T - t0))
    return torch.tensor(r0, dtype=torch.int32).cuda(), torch.tensor(nr, dtype=torch.int32).cuda()


def t(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


r128, n128 = tiles(128)
r64, n64 = tiles(64)
new = lambda: _lib.check(L.mtts_gqa_prefill_tc(q.data_ptr(), R, k_pool.data_ptr(), v_pool.data_ptr(), None, max_pages, page, num_pages,
                                               r128.data_ptr(), n128.data_ptr(), seq.data_ptr(), pos.data_ptr(), out.data_ptr(),
                                               r128.numel(), Hq, Hkv, D, _lib.stream_ptr()))
old = lambda: _lib.check(L.mtts_gqa_attention(q.data_ptr(), k_pool.data_ptr(), v_pool.data_ptr(), None, max_pages, page, r64.data_ptr(),
                                              n64.data_ptr(), seq.data_ptr(), pos.data_ptr(), out.data_ptr(), r64.numel(), 64, Hq, Hkv, D,
                                              1, None, 0, _lib.stream_ptr()))
flops = 4.0 * B * Hq * D * (T * (T + 1) / 2)
for name, fn in (("tcgen05 (128-row tiles)", new), ("mma.sync (64-row tiles)", old)):
    us = t(fn)
    print(f"B={B} T={T}  {name:26s} {us:9.1f} us  {flops / us / 1e6:7.1f} TFLOP/s (causal flops)", flush=True)
