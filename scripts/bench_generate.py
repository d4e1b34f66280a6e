"""LM-only timing of generate() at the bench shape: prefill ms and decode ms/step over a few repeats."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scripts.bench_lm import SHAPE, make_prompt
from moss_ttsd_b200.modeling_asteroid import AsteroidTTSConfig, AsteroidTTSInstruct

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
new = int(sys.argv[2]) if len(sys.argv) > 2 else 375
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
cfg = AsteroidTTSConfig(**SHAPE, eos_token_id=152694, pad_token_id=151643, tie_word_embeddings=False)
m = AsteroidTTSInstruct(cfg, device="cuda")
m.init_random_weights(0, tied=False, speech_only_head0=True)
m.generation_config.eos_token_id = 152694
ids, mask = make_prompt(np.random.default_rng(0), B, 200, 250)
ids, mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
for r in range(reps):
    out = m.generate(input_ids=ids, attention_mask=mask, max_new_tokens=new)
    torch.cuda.synchronize()
    ev, steps = m._last_timing
    pre, dec = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    print(f"rep {r}: out {tuple(out.shape)} prefill {pre:.1f} ms, decode {dec:.1f} ms over {steps - 1} steps = {dec / (steps - 1):.3f} ms/step", flush=True)
