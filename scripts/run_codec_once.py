"""One batched XY_Tokenizer.decode (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, yaml
from moss_ttsd_b200.xy_tokenizer.model import XY_Tokenizer
items = int(sys.argv[1]) if len(sys.argv) > 1 else 64
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(root, "moss-ttsd_b200", "xy_tokenizer", "xy_tokenizer_config.yaml")) as f:
    spt = XY_Tokenizer(yaml.safe_load(f)["generator_params"])
spt.init_random_weights(seed=5, device="cuda")
codes = [torch.randint(0, 1024, (8, 375), device="cuda") for _ in range(items)]
torch.cuda.synchronize()
print("CODEC_BEGIN", flush=True)
w = spt.decode(codes)["syn_wav_list"]
torch.cuda.synchronize()
print("ok", w[0].shape)
