"""One batched XY_Tokenizer.encode (B items x S seconds of 16 kHz audio, shipped config, exact mode): for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, yaml
from moss_ttsd_b200.xy_tokenizer.model import XY_Tokenizer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S = int(sys.argv[2]) if len(sys.argv) > 2 else 30
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(root, "moss-ttsd_b200", "xy_tokenizer", "xy_tokenizer_config.yaml")) as f:
    spt = XY_Tokenizer(yaml.safe_load(f)["generator_params"])
spt.init_random_weights(seed=5, device="cuda", encoder=True)
wavs = [torch.randn(16000 * S, device="cuda") * 0.1 for _ in range(B)]
for _ in range(2):
    out = spt.encode(wavs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
out = spt.encode(wavs)
e1.record(); torch.cuda.synchronize()
print("encode ms", e0.elapsed_time(e1), out["codes_list"][0].shape)
