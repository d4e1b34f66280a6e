"""Soak of the persistent decode kernel: many teacher-synchronised steps against the kernel chain (3-layer full-width
model), long enough for the tag counters / ring phases to wrap many times and the context to cross many pages."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import test_mega_gpu as T

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
for B, paged in ((1, False), (4, True), (2, False)):
    shape, w, chain, mega = T._engine_pair(3, seed=11 + B)
    ids, mask = T._prompt(shape, B, 90, seed=B)
    st_c, _ = T._session(chain, shape, B, ids, mask, paged, rows=steps + 200)
    st_m, _ = T._session(mega, shape, B, ids, mask, paged, rows=steps + 200)
    assert st_m["mega"] is not None
    worst, agree, total = 0.0, 0, 0
    for s in range(steps):
        chain.decode_step(st_c)
        mega.decode_step(st_m)
        if s % 25 == 0 or s > steps - 5:
            lc, lm = st_c["logits"].float(), st_m["logits"].float()
            assert torch.isfinite(lm).all(), s
            worst = max(worst, (lc - lm).abs().max().item() / max(1.0, lc.abs().max().item()))
        agree += int((st_c["tokens"] == st_m["tokens"]).sum()); total += st_c["tokens"].numel()
        st_m["tokens"].copy_(st_c["tokens"])
    torch.cuda.synchronize()
    assert int(mega.err.abs().sum()) == 0
    print(f"B={B} paged={paged}: {steps} steps, worst relative |dlogit| {worst:.4f}, token agreement {agree/total:.4f}", flush=True)
    assert worst < 0.05 and agree / total > 0.97
print("soak ok")
