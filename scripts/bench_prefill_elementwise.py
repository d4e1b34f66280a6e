"""The two memory-bound kernels of a bench-sized prefill (135 k packed rows): mtts_rmsnorm and mtts_qknorm_rope_kvappend,
timed with CUDA events over buffers far larger than L2; GB/s = algorithmic bytes / time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from moss_ttsd_b200 import _lib, ops
ops.ensure_init()
L = _lib.load()
sp = _lib.stream_ptr
dev = "cuda"
R = int(sys.argv[1]) if len(sys.argv) > 1 else 134983
B, per = 256, 1152            # the engine's layout: pages of 64 rows, 18 pages per sequence, contiguous
x = torch.randn(R, 2048, device=dev).to(torch.bfloat16)
w = torch.randn(2048, device=dev).to(torch.bfloat16)
out = torch.empty_like(x)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


us = timeit(lambda: _lib.check(L.mtts_rmsnorm(x.data_ptr(), x.stride(0), w.data_ptr(), out.data_ptr(), out.stride(0), R, 2048, 1e-6, sp())))
print(f"mtts_rmsnorm               rows {R}: {us:7.1f} us  {2 * x.numel() * 2 / us / 1e3:7.1f} GB/s")
qkv = torch.randn(R, 4096, device=dev).to(torch.bfloat16)
q = torch.empty(R, 2048, device=dev, dtype=torch.bfloat16)
kp = torch.empty(B * (per // 64), 8, 64, 128, device=dev, dtype=torch.bfloat16)
vp = torch.empty_like(kp)
qn = torch.randn(128, device=dev).to(torch.bfloat16)
kn = torch.randn(128, device=dev).to(torch.bfloat16)
inv_freq = (1.0 / (1e6 ** (torch.arange(0, 128, 2, dtype=torch.float32) / 128))).to(dev)
lens = np.full(B, R // B)
lens[: R - lens.sum()] += 1
pos = torch.from_numpy(np.concatenate([np.arange(n) for n in lens]).astype(np.int32)).to(dev)
seq = torch.from_numpy(np.repeat(np.arange(B), lens).astype(np.int32)).to(dev)
us = timeit(lambda: _lib.check(L.mtts_qknorm_rope_kvappend(qkv.data_ptr(), qkv.stride(0), qn.data_ptr(), kn.data_ptr(), inv_freq.data_ptr(),
                                                           pos.data_ptr(), seq.data_ptr(), q.data_ptr(), kp.data_ptr(), vp.data_ptr(), None, per // 64, 64,
                                                           B * (per // 64), R, 16, 8, 128, 1e-6, None, sp())))
print(f"mtts_qknorm_rope_kvappend  rows {R}: {us:7.1f} us  {2 * qkv.numel() * 2 / us / 1e3:7.1f} GB/s")
