"""In-graph time of each dense-projection phase of a decode step at batch M (default 256), 28 layers of distinct
weights per sweep so nothing is L2-resident: q/k/v partials, o_proj partials + reducer, gate/up + SwiGLU, down partials +
reducer, and the whole projection chain of a layer (no attention). us per layer, graph replay, CUDA events."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moss_ttsd_b200 import _lib, ops
from moss_ttsd_b200._lib import check, ptr, stream_ptr
from moss_ttsd_b200.lm_engine import LMShape, LMWeights, DecoderEngine

M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
shape = LMShape()
w = LMWeights(shape, "cuda").init_random_(0)
eng = DecoderEngine(w)
if os.environ.get("MTTS_BENCH_TILED", "0") == "1":   # weights in the tile-contiguous layout
    for lw in w.layers:
        for k in ("wqkv", "wo", "wgu", "wd"):
            lw[k] = ops.TiledWeight(lw[k])
L = eng.L
a = eng._alloc_acts(M)
for t in a.values():
    t.normal_(0, 1)
gws = eng._gemm_ws(M)
pws = eng._splitk_ws(M)
x, xn, ao, h = a["x"], a["xn"], a["ao"], a["h"]
H = shape.hidden_size
eps = shape.rms_norm_eps


def reduce_norm(S, lw_norm):
    check(L.mtts_splitk_reduce_rmsnorm(ptr(pws), S, M, H, ptr(x), x.stride(0), ptr(lw_norm), ptr(xn), xn.stride(0), eps, stream_ptr()))


def ph_qkv():
    for lw in w.layers:
        eng._splitk(xn, lw["wqkv"], pws)


def ph_o():
    for lw in w.layers:
        S = eng._splitk(ao, lw["wo"], pws)
        reduce_norm(S, lw["ln2"])


def ph_o_gemm():
    for lw in w.layers:
        eng._splitk(ao, lw["wo"], pws)


def ph_gu():
    for lw in w.layers:
        ops.gemm(xn, lw["wgu"], out=h, swiglu=True, workspace=gws)


def ph_down():
    for lw in w.layers:
        S = eng._splitk(h, lw["wd"], pws)
        reduce_norm(S, lw["ln1"])


def ph_down_gemm():
    for lw in w.layers:
        eng._splitk(h, lw["wd"], pws)


def ph_chain():
    for lw in w.layers:
        eng._splitk(xn, lw["wqkv"], pws)
        S = eng._splitk(ao, lw["wo"], pws)
        reduce_norm(S, lw["ln2"])
        ops.gemm(xn, lw["wgu"], out=h, swiglu=True, workspace=gws)
        S = eng._splitk(h, lw["wd"], pws)
        reduce_norm(S, lw["ln1"])


def replay_us(fn, reps=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * len(w.layers))


only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
for name, fn in (("qkv", ph_qkv), ("o_gemm", ph_o_gemm), ("o+reduce", ph_o), ("gate/up", ph_gu), ("down_gemm", ph_down_gemm),
                 ("down+reduce", ph_down), ("chain", ph_chain)):
    if only and name not in only:
        continue
    print(f"M={M} {name:12s} {replay_us(fn):8.2f} us/layer", flush=True)
