"""A few eager decode steps of the persistent small-batch kernel (for ncu captures)."""
import os, sys
os.environ.setdefault("MTTS_NO_GRAPH", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scripts.bench_lm import SHAPE, make_prompt
from moss_ttsd_b200.modeling_asteroid import AsteroidTTSConfig, AsteroidTTSInstruct

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = AsteroidTTSConfig(**SHAPE, eos_token_id=152694, pad_token_id=151643, tie_word_embeddings=True)
m = AsteroidTTSInstruct(cfg, device="cuda")
m.init_random_weights(0)
m._w.heads[:151665].zero_(); m._w.heads[152689:152704].zero_()
m.generation_config.eos_token_id = 152694
ids, mask = make_prompt(np.random.default_rng(0), B, 200, 250)
out = m.generate(input_ids=torch.from_numpy(ids).cuda(), attention_mask=torch.from_numpy(mask).cuda(), max_new_tokens=8)
torch.cuda.synchronize()
print("ok", tuple(out.shape))
