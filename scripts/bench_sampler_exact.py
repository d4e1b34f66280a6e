"""The sampling step of all 8 channels (scan + finish [+ exact wide-vocabulary kernel]) for a few configurations:
the candidate-list fast path next to the exact kernel that handles what the list cannot represent."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moss_ttsd_b200 import _lib
from moss_ttsd_b200.lm_engine import LMShape, SamplerSetup

L = _lib.load()
shape = LMShape(num_hidden_layers=1)
CFGS = [("top_k 50 + top_p 0.95 (candidate list)", dict(repetition_penalty=1.1, temperature=0.9, top_k=50, top_p=0.95)),
        ("top_k 1000 (exact)", dict(temperature=0.9, top_k=1000)),
        ("temperature only (exact)", dict(temperature=0.9)),
        ("top_k 3000 + top_p 0.9 (exact, 32-pass bisection)", dict(temperature=1.3, top_k=3000, top_p=0.9))]
for B in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,64,256").split(",")]:
    logits = (torch.randn((B, shape.vpad), device="cuda") * 2.5).to(torch.bfloat16)
    for name, cfg in CFGS:
        sm = SamplerSetup(shape, [True] * 8, [dict(cfg) for _ in range(8)])
        seen = torch.zeros((B, sm.words_per_row), dtype=torch.int32, device="cuda")
        step = torch.full((1,), 9, dtype=torch.int32, device="cuda")
        toks = torch.zeros((B, 8), dtype=torch.int64, device="cuda")
        err = torch.zeros(8, dtype=torch.int32, device="cuda")
        sws = torch.zeros(L.mtts_sample8_workspace_bytes(B, 8), dtype=torch.uint8, device="cuda")
        seed = torch.zeros(1, dtype=torch.int64, device="cuda")

        def run():
            _lib.check(L.mtts_sample8(logits.data_ptr(), logits.stride(0), B, ctypes.byref(sm.cfg), seen.data_ptr(),
                                      step.data_ptr(), seed.data_ptr(), toks.data_ptr(), err.data_ptr(), sws.data_ptr(),
                                      sws.numel(), _lib.stream_ptr()))
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        print(f"B={B:3d} {name:52s} {e0.elapsed_time(e1) / 20 * 1e3:9.1f} us per step, err flags {err.cpu().tolist()[:4]}", flush=True)
