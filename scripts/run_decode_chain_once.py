"""The dense-projection chain of a batch-M decode step (3 layers, eager launches, distinct weights): for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moss_ttsd_b200 import ops
from moss_ttsd_b200._lib import check, ptr, stream_ptr
from moss_ttsd_b200.lm_engine import LMShape, LMWeights, DecoderEngine

M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
shape = LMShape(num_hidden_layers=3)
w = LMWeights(shape, "cuda").init_random_(0)
eng = DecoderEngine(w)
a = eng._alloc_acts(M)
for t in a.values():
    t.normal_(0, 1)
gws, pws = eng._gemm_ws(M), eng._splitk_ws(M)
x, xn, ao, h = a["x"], a["xn"], a["ao"], a["h"]
for lw in w.layers:
    eng._splitk(xn, lw["wqkv"], pws)
    S = eng._splitk(ao, lw["wo"], pws)
    check(eng.L.mtts_splitk_reduce_rmsnorm(ptr(pws), S, M, 2048, ptr(x), x.stride(0), ptr(lw["ln2"]), ptr(xn), xn.stride(0), 1e-6, stream_ptr()))
    ops.gemm(xn, lw["wgu"], out=h, swiglu=True, workspace=gws)
    S = eng._splitk(h, lw["wd"], pws)
    check(eng.L.mtts_splitk_reduce_rmsnorm(ptr(pws), S, M, 2048, ptr(x), x.stride(0), ptr(lw["ln1"]), ptr(xn), xn.stride(0), 1e-6, stream_ptr()))
torch.cuda.synchronize()
print("ok")
