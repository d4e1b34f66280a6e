"""Codec attention alone (B items x T rows, 12 heads x 64): tcgen05 kernel (fp16 in/out) against the mma.sync fp16 kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moss_ttsd_b200 import _lib, ops
ops.ensure_init()
L = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
H, D = 12, 64
E = H * D
qkv32 = torch.randn(B * T, 3 * E, device="cuda")
qkv16 = qkv32.half()
o32 = torch.empty(B * T, E, device="cuda")
o16 = torch.empty(B * T, E, device="cuda", dtype=torch.float16)
lens = torch.full((B,), T, dtype=torch.int32, device="cuda")


def t(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


flops = 4.0 * B * H * D * T * T
for name, fn in (("tcgen05 fp16", lambda: _lib.check(L.mtts_mha_varlen_tc(qkv16.data_ptr(), o16.data_ptr(), lens.data_ptr(), B, T, H, D, _lib.stream_ptr()))),
                 ("mma.sync fp16", lambda: _lib.check(L.mtts_mha_varlen_f16(qkv32.data_ptr(), o32.data_ptr(), lens.data_ptr(), B, T, H, D, _lib.stream_ptr())))):
    us = t(fn)
    print(f"B={B} T={T}  {name:14s} {us:9.1f} us  {flops / us / 1e6:7.1f} TFLOP/s", flush=True)
