"""Every dense GEMM of one batched XY_Tokenizer.decode (64 x 30 s): shape, epilogue, time and TFLOP/s (events around each call)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, yaml
from moss_ttsd_b200 import ops
from moss_ttsd_b200.xy_tokenizer.model import XY_Tokenizer

items = int(sys.argv[1]) if len(sys.argv) > 1 else 64
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(root, "moss-ttsd_b200", "xy_tokenizer", "xy_tokenizer_config.yaml")) as f:
    spt = XY_Tokenizer(yaml.safe_load(f)["generator_params"])
spt.init_random_weights(seed=5, device="cuda")
codes = [torch.randint(0, 1024, (8, 375), device="cuda") for _ in range(items)]
spt.decode(codes)
torch.cuda.synchronize()
log = []
orig = ops.gemm
def hooked(x, w, *a, **kw):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    r = orig(x, w, *a, **kw)
    e1.record()
    log.append((tuple(x.shape), tuple(w.shape), str(x.dtype)[6:], tuple(sorted(k for k, v in kw.items() if v is not None and v is not False and k not in ("out", "workspace"))), e0, e1))
    return r
ops.gemm = hooked
import moss_ttsd_b200.xy_tokenizer.model as mm
if hasattr(mm, "ops"):
    mm.ops.gemm = hooked
t0, t1 = torch.cuda.Event(True), torch.cuda.Event(True)
t0.record()
spt.decode(codes)
t1.record()
torch.cuda.synchronize()
agg = collections.OrderedDict()
for xs, ws, dt, fl, e0, e1 in log:
    k = (xs, ws, dt, fl)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
tot_ms = sum(v[1] for v in agg.values())
tot_fl = 0
print(f"decode total {t0.elapsed_time(t1):.1f} ms; {len(log)} GEMM calls, {tot_ms:.1f} ms inside them")
for (xs, ws, dt, fl), (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    M = 1
    for d in xs[:-1]:
        M *= d
    fl_ = 2.0 * M * ws[0] * ws[1] * c
    tot_fl += fl_
    print(f"  {ms:8.2f} ms x{c:3d}  M={M:7d} N={ws[0]:5d} K={ws[1]:5d} {dt} {','.join(fl):28s} {fl_/ms/1e9:7.1f} TFLOP/s")
print(f"total {tot_fl/1e12:.1f} TFLOP -> {tot_fl/tot_ms/1e9:.1f} TFLOP/s inside GEMMs")
