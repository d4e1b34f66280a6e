"""Per-phase cycle breakdown of the persistent small-batch decode kernel (CTA 0's view).
Needs a profiling build: MTTS_NVCC_FLAGS="-DMTTS_MEGA_PROFILE [-DMTTS_MEGA_TRACE]" python moss-ttsd_b200/build.py --force"""
import os, sys
os.environ["MTTS_MEGA_PROFILE"] = "1"
os.environ.setdefault("MTTS_MEGA_MAX_B", "4")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scripts.bench_lm import SHAPE, make_prompt
from moss_ttsd_b200.modeling_asteroid import AsteroidTTSConfig, AsteroidTTSInstruct

NAMES = ["ln1+frag", "qkv", "bar", "attention", "bar", "attn_out+frag", "wo", "bar", "ln2+frag", "gate_up", "bar",
         "h frag", "down", "bar", "final norm", "heads"]

def main():
    cfg = AsteroidTTSConfig(**SHAPE, eos_token_id=152694, pad_token_id=151643, tie_word_embeddings=True)
    m = AsteroidTTSInstruct(cfg, device="cuda")
    m.init_random_weights(0)
    m._w.heads[:151665].zero_(); m._w.heads[152689:152704].zero_()
    m.generation_config.eos_token_id = 152694
    rng = np.random.default_rng(0)
    for B in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,8").split(",")]:
        ids, mask = make_prompt(rng, B, int(os.environ.get("MTTS_PROFILE_TEXT_ROWS", "200")), 250)
        ids, mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
        m.generate(input_ids=ids, attention_mask=mask, max_new_tokens=8)
        st = m._last_state
        prof = st["mega"]["prof"]
        prof.zero_()
        n = 20
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        for _ in range(n):
            st["graph"].replay()
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / n
        c = prof.cpu().numpy()[:16] / n
        tot = c.sum()
        print(f"B={B} step {ms:.3f} ms; CTA0 cycles/step {tot:.0f} (={tot/ms/1e3:.0f} MHz equiv)")
        tr = prof.cpu().numpy()[32:32 + 148 * 16].reshape(148, 16).astype(np.float64)
        if tr.max() == 0:
            tr = None
        t0 = tr[:, 13].min() if tr is not None else 0
        cols = [13, 0, 1, 3, 5, 6, 8, 9, 11, 12]
        lab = ["start", "ln1", "qkv", "attn", "attn_out", "wo", "ln2", "gu", "hfrag", "down"]
        print("   layer-5 timeline, us since first CTA entered the layer: min / median / max over CTAs")
        for cc, nm in zip(cols, lab):
            if tr is None:
                break
            v = (tr[:, cc] - t0) / 1e3
            print(f"      {nm:9s} {v.min():7.2f} {np.median(v):7.2f} {v.max():7.2f}   argmax CTA {int(v.argmax())}")
        ca = prof.cpu().numpy()[16:22] / n
        print("   attention (CTA 0) cycles/layer:", dict(zip(["kv issue", "qkv poll", "norm+rope", "sync", "kv loop", "merge+publish"], (ca / 28).round(0).tolist())))
        pc = prof.cpu().numpy() / n / 28
        print("   gate_up inner (cycles/layer): frag loads %.0f, ldmatrix+mma+arrive %.0f, reduce store+sync %.0f, epilogue %.0f" % (pc[10], pc[2], pc[4], pc[7]))
        cw = prof.cpu().numpy()[24:29] / n
        print("   cycles/layer waiting for weight chunks:", dict(zip(["qkv", "wo", "gate_up", "down"], (cw[:4] / 28).round(0).tolist())), "heads total", round(cw[4]))
        for nm, v in zip(NAMES, c):
            print(f"   {nm:14s} {v:10.0f} cyc  {100*v/tot:5.1f}%   {v/tot*ms*1e3/ (1 if nm in ('final norm','heads') else 28):7.2f} us/layer")

if __name__ == "__main__":
    main()
