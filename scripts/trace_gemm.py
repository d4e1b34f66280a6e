"""Phase timeline of the weight-streaming GEMM (gemm_tc_kernel) inside a decode-like launch sequence. Needs the debug
build: MTTS_NVCC_FLAGS=-DMTTS_GEMM_TRACE python moss-ttsd_b200/build.py --force. Prints, per phase, the median and the
maximum over CTAs of the time since the first CTA of the traced launch started (microseconds)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from moss_ttsd_b200 import _lib, ops

L = _lib.load()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H, I, QKV = 2048, 6144, 4096
NL = 6
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda n, k: (torch.randn(n, k, device="cuda", generator=g) * 0.02).to(torch.bfloat16)
layers = [dict(wqkv=mk(QKV, H), wo=mk(H, 2048), wgu=mk(2 * I, H), wd=mk(H, I)) for _ in range(NL)]
x = torch.randn(M, H, device="cuda").to(torch.bfloat16)
hq = torch.randn(M, 2048, device="cuda").to(torch.bfloat16)
hi = torch.randn(M, I, device="cuda").to(torch.bfloat16)
o_qkv = torch.empty(M, QKV, device="cuda", dtype=torch.bfloat16)
o_h = torch.zeros(M, H, device="cuda", dtype=torch.bfloat16)
o_i = torch.empty(M, I, device="cuda", dtype=torch.bfloat16)
NAMES = ["entry", "tmem ready", "producer: pdl_wait done", "mma: first stage landed", "mma: last commit issued",
         "epi: accumulator complete", "epi: staged", "cluster sync 1 done", "epi: stored", "cluster sync 2 done", "dealloc done",
         "epi: pdl_wait done"]


def seq(stop):
    for li, lw in enumerate(layers):
        last = li == NL - 1
        ops.gemm(x, lw["wqkv"], out=o_qkv)
        if last and stop == "qkv": return
        ops.gemm(hq, lw["wo"], out=o_h, residual=o_h)
        if last and stop == "wo": return
        ops.gemm(x, lw["wgu"], out=o_i, swiglu=True)
        if last and stop == "gu": return
        ops.gemm(hi, lw["wd"], out=o_h, residual=o_h)


grids = {"qkv": QKV // 128, "wo": H // 128, "gu": 2 * I // 128, "wd": H // 128}
fn = L.mtts_debug_gemm_trace
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
for stop in ("qkv", "wo", "gu", "wd"):
    for _ in range(2):
        seq(stop)
        torch.cuda.synchronize()
    buf = np.zeros(4096 * 16, dtype=np.uint64)
    assert fn(buf.ctypes.data, buf.size) == 0
    t = buf.reshape(4096, 16).astype(np.int64)
    live = t[:, 0] > 0
    # CTAs of the last launch: those whose entry stamp is within 200 us of the latest entry
    latest = t[live, 0].max()
    sel = live & (t[:, 0] > latest - 200_000)
    tt = t[sel]
    t0 = tt[:, 0].min()
    print(f"== {stop}: M={M}, {sel.sum()} CTAs traced, kernel span {(tt[:, 10].max() - t0) / 1e3:.2f} us")
    for slot in (0, 1, 2, 11, 3, 4, 5, 6, 7, 8, 9, 10):
        v = (tt[:, slot] - t0) / 1e3
        print(f"   {NAMES[slot]:28s} median {np.median(v):7.2f}  min {v.min():7.2f}  max {v.max():7.2f} us")
