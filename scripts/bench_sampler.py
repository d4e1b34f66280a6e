"""Greedy sampler scan alone at a few batch sizes (logits resident in L2 as they are right after the LM-heads GEMM)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moss_ttsd_b200 import _lib
from moss_ttsd_b200.lm_engine import LMShape, SamplerSetup

L = _lib.load()
shape = LMShape(num_hidden_layers=1)
for B in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,64,256").split(",")]:
    logits = torch.randn((B, shape.vpad), device="cuda").to(torch.bfloat16)
    sm = SamplerSetup(shape, [False] * 8, None)
    seen = torch.zeros((B, sm.words_per_row), dtype=torch.int32, device="cuda")
    step = torch.full((1,), 9, dtype=torch.int32, device="cuda")
    toks = torch.zeros((B, 8), dtype=torch.int64, device="cuda")
    err = torch.zeros(4, dtype=torch.int32, device="cuda")
    sws = torch.zeros(L.mtts_sample8_workspace_bytes(B, 8), dtype=torch.uint8, device="cuda")
    seed = torch.zeros(1, dtype=torch.int64, device="cuda")

    def run():
        _lib.check(L.mtts_sample8(logits.data_ptr(), logits.stride(0), B, ctypes.byref(sm.cfg), seen.data_ptr(), step.data_ptr(),
                                  seed.data_ptr(), toks.data_ptr(), err.data_ptr(), sws.data_ptr(), sws.numel(), _lib.stream_ptr()))
    for _ in range(5):
        run()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(50):
        run()
    e1.record()
    torch.cuda.synchronize()
    ref = logits[:, :shape.vocabs[0]].float().argmax(-1)
    print(f"B={B} chunk={os.environ.get('MTTS_SAMPLE_CHUNK', 'auto')}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per scan, "
          f"argmax ok={bool(torch.equal(ref, toks[:, 0]))}", flush=True)
