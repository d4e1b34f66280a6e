"""Timings of the BASELINE.json configurations that are not the bench.py line (C2 codec round trip, C4 long-form
KV-bound decode, C5 batch-256 decode + codec), on one GPU, synthetic data. Writes gpurun_out/configs.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, yaml
from scripts.bench_lm import SHAPE, make_prompt
from moss_ttsd_b200.lm_engine import KVCache, SamplerSetup
from moss_ttsd_b200.modeling_asteroid import AsteroidTTSConfig, AsteroidTTSInstruct
from moss_ttsd_b200.xy_tokenizer.model import XY_Tokenizer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HBM = 6541.8
res = {}


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def codec():
    with open(os.path.join(ROOT, "moss-ttsd_b200", "xy_tokenizer", "xy_tokenizer_config.yaml")) as f:
        spt = XY_Tokenizer(yaml.safe_load(f)["generator_params"])
    spt.init_random_weights(seed=5, device="cuda", encoder=True)
    return spt


def c2(spt):
    g = torch.Generator(device="cuda").manual_seed(0)
    wavs = [0.1 * torch.randn(960000, device="cuda", generator=g) for _ in range(32)]
    enc = lambda: spt.encode(wavs)["codes_list"]
    codes = enc()
    ms_e = timed(enc, 2)
    dec = lambda: spt.decode(codes)["syn_wav_list"]
    out = dec()
    ms_d = timed(dec, 2)
    res["C2_codec_roundtrip_b32_60s"] = dict(encode_ms=ms_e, decode_ms=ms_d, codes_per_item=int(codes[0].shape[1]),
                                             samples_per_item=int(out[0].numel()),
                                             encode_audio_s_per_s=32 * 60 / (ms_e / 1e3), decode_audio_s_per_s=32 * 60 / (ms_d / 1e3))
    print("C2", res["C2_codec_roundtrip_b32_60s"], flush=True)


def decode_step_at(model, B, ctx, paged, label):
    eng, shape = model.engine, model.shape
    cache = KVCache(shape, B, ctx + 64, "cuda", paged=paged, shuffle_pages=paged)
    cache.k.normal_(); cache.v.normal_()
    sm = SamplerSetup(shape, [False] * 8, None)
    st = eng.make_decode_state(B, cache, sm, ctx + 64, (151665, 152689), 152694, False)
    eng.reset_decode_state(st, 0, ctx - 8, ctx + 48)
    st["positions"].fill_(ctx - 1)
    st["tokens"][:, 0] = 151700
    st["tokens"][:, 1:] = 5
    for _ in range(3):
        eng.decode_step(st)
    ms = timed(lambda: eng.decode_step(st), 12)
    w = model._w
    streamed = w.heads.numel() * 2 + sum(lw[k].numel() * 2 for lw in w.layers for k in ("wqkv", "wo", "wgu", "wd")) + B * ctx * 114688
    res[label] = dict(batch=B, context_rows=ctx, kv="paged" if paged else "contiguous", ms_per_step=ms,
                      audio_s_per_s=B * 0.08 / (ms / 1e3), streamed_gb=streamed / 1e9, hbm_frac=streamed / (ms * 1e-3) / 1e9 / HBM)
    print(label, res[label], flush=True)
    del st, cache
    torch.cuda.empty_cache()


def main():
    spt = codec()
    c2(spt)
    cfg = AsteroidTTSConfig(**SHAPE, eos_token_id=152694, pad_token_id=151643, tie_word_embeddings=True)
    model = AsteroidTTSInstruct(cfg, device="cuda")
    model.init_random_weights(0)
    model._w.heads[:151665].zero_(); model._w.heads[152689:152704].zero_()
    model.generation_config.eos_token_id = 152694
    for ctx in (4000, 12000):
        for paged in (False, True):
            decode_step_at(model, 16, ctx, paged, f"C4_longform_b16_ctx{ctx}_{'paged' if paged else 'contig'}")
    decode_step_at(model, 256, 700, False, "C5_decode_b256_ctx700")
    # C5 end to end on one GPU: 256 scripts, 375 frames each, LM + codec
    ids, mask = make_prompt(np.random.default_rng(0), 256, 200, 250)
    ids, mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    from moss_ttsd_b200.generation_utils import undelay
    def e2e():
        out = model.generate(input_ids=ids, attention_mask=mask, max_new_tokens=375, do_sample=False)
        speech = undelay(out[:, ids.shape[1] - 7:])
        speech[..., 0] = (speech[..., 0] - 151665).clamp(0, 1023)
        return spt.decode([speech[i].clamp(0, 1023).permute(1, 0) for i in range(256)])["syn_wav_list"]
    e2e()
    torch.cuda.synchronize(); t0 = time.time()
    w = e2e()
    torch.cuda.synchronize(); dt = time.time() - t0
    res["C5_e2e_b256_375frames_1gpu"] = dict(seconds=dt, audio_s_per_s=256 * 30 / dt, samples_per_item=int(w[0].numel()))
    print("C5 e2e", res["C5_e2e_b256_375frames_1gpu"], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/configs.json", "w"), indent=1)


if __name__ == "__main__":
    main()
