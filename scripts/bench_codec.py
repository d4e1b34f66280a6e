"""Codec decode timing at the bench shape (B items x 375 frames, shipped config, random init): fp16-operand vs TF32 GEMMs."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, yaml
from moss_ttsd_b200.xy_tokenizer.model import XY_Tokenizer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
with open(os.path.join(ROOT, "moss-ttsd_b200", "xy_tokenizer", "xy_tokenizer_config.yaml")) as f:
    spt = XY_Tokenizer(yaml.safe_load(f)["generator_params"])
spt.init_random_weights(seed=5, device="cuda")
codes = [torch.randint(0, 1024, (8, 375), device="cuda") for _ in range(B)]
out = {}
ref = None
for mode in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("tf32", "f16")):
    spt.decode_gemm = mode
    for _ in range(2):
        w = spt.decode(codes)["syn_wav_list"]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    if os.environ.get("MTTS_PROFILE_RANGE") == mode:      # ncu --profile-from-start off: exactly one decode call
        torch.cuda.profiler.start()
        w = spt.decode(codes)["syn_wav_list"]
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    e0.record()
    for _ in range(3):
        w = spt.decode(codes)["syn_wav_list"]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    w0 = torch.stack(w[:8]).double()
    if ref is None:
        ref = w0
        snr = None
    else:
        snr = float(10 * torch.log10((ref ** 2).sum() / ((ref - w0) ** 2).sum()))
    out[mode] = dict(ms=ms, audio_s_per_s=B * 30 / (ms / 1e3), tflops_effective=B * 1139.2e9 / (ms * 1e-3) / 1e12, snr_vs_tf32_db=snr)
    print(mode, out[mode], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"bench_codec_b{B}.json"), "w"), indent=1)
