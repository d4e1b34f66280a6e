"""Prefill (B=64 x 457 rows, few layers) + one batched codec decode window, eager, for ncu launch lists."""
import os, sys
os.environ["MTTS_NO_GRAPH"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, yaml
from scripts.bench_lm import SHAPE, make_prompt
from moss_ttsd_b200.modeling_asteroid import AsteroidTTSConfig, AsteroidTTSInstruct
from moss_ttsd_b200.xy_tokenizer.model import XY_Tokenizer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 4
items = int(sys.argv[3]) if len(sys.argv) > 3 else 16
shape = dict(SHAPE, num_hidden_layers=layers)
cfg = AsteroidTTSConfig(**shape, eos_token_id=152694, pad_token_id=151643, tie_word_embeddings=True)
m = AsteroidTTSInstruct(cfg, device="cuda")
m.init_random_weights(0, speech_only_head0=True)
m.generation_config.eos_token_id = 152694
ids, mask = make_prompt(np.random.default_rng(0), B, 200, 250)
out = m.generate(input_ids=torch.from_numpy(ids).cuda(), attention_mask=torch.from_numpy(mask).cuda(), max_new_tokens=1)
torch.cuda.synchronize()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(root, "moss-ttsd_b200", "xy_tokenizer", "xy_tokenizer_config.yaml")) as f:
    spt = XY_Tokenizer(yaml.safe_load(f)["generator_params"])
spt.init_random_weights(seed=5, device="cuda")
codes = [torch.randint(0, 1024, (8, 375), device="cuda") for _ in range(items)]
torch.cuda.synchronize()
print("CODEC_BEGIN", flush=True)
w = spt.decode(codes)["syn_wav_list"]
torch.cuda.synchronize()
print("ok", out.shape, w[0].shape)
