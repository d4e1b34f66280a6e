"""A few eager decode steps with every channel sampled (rep. penalty 1.1, temperature 0.9, top-k 50, top-p 0.95): for ncu launch lists."""
import os, sys
os.environ["MTTS_NO_GRAPH"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scripts.bench_lm import SHAPE, make_prompt
from moss_ttsd_b200.modeling_asteroid import AsteroidTTSConfig, AsteroidTTSInstruct

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
shape = dict(SHAPE, num_hidden_layers=2)
cfg = AsteroidTTSConfig(**shape, eos_token_id=152694, pad_token_id=151643, tie_word_embeddings=True)
m = AsteroidTTSInstruct(cfg, device="cuda")
m.init_random_weights(0)
m.generation_config.eos_token_id = 152694
m.generation_config.do_samples = [True] * 8
m.generation_config.layers = [dict(repetition_penalty=1.1, temperature=0.9, top_k=50, top_p=0.95)] * 8
ids, mask = make_prompt(np.random.default_rng(0), B, 200, 250)
out = m.generate(input_ids=torch.from_numpy(ids).cuda(), attention_mask=torch.from_numpy(mask).cuda(), max_new_tokens=3)
torch.cuda.synchronize()
print("ok", out.shape)
