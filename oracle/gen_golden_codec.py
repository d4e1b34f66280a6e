"""TEST INFRASTRUCTURE. Codec-decode goldens from the UNMODIFIED reference XY_Tokenizer (CPU, build container only)."""
import os

import numpy as np
import torch

from oracle import ref_shims
from oracle.codec_weights import TINY_CODEC, full_codec_params, make_codec_weights

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def build_ref_codec(gp, seed):
    XY, _, _ = ref_shims.import_codec()
    torch.manual_seed(0)
    model = XY(gp).eval()
    sd = make_codec_weights(gp, seed)
    tsd = {k: torch.from_numpy(v) for k, v in sd.items()}
    missing, unexpected = model.load_state_dict(tsd, strict=False)
    assert not unexpected, unexpected
    dec_prefixes = ("quantizer.", "post_rvq_adapter.", "upsample.", "acoustic_decoder.", "enhanced_vocos.")
    bad = [m for m in missing if m.startswith(dec_prefixes) and not any(
        s in m for s in ("positional_embedding", "inited", "cluster_size", "embed_avg", "istft.window"))]
    assert not bad, bad
    for q in model.quantizer.quantizers:
        q.inited.fill_(True)
    return model


def build_ref_codec_with_encoder(gp, seed):
    from oracle.codec_weights import make_encoder_weights
    XY, _, _ = ref_shims.import_codec()
    torch.manual_seed(0)
    model = XY(gp).eval()
    sd = make_codec_weights(gp, seed)
    sd.update(make_encoder_weights(gp, seed + 7))
    missing, unexpected = model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not unexpected, unexpected
    bad = [m for m in missing if not any(s in m for s in ("positional_embedding", "inited", "cluster_size", "embed_avg", "istft.window"))]
    assert not bad, bad
    for q in model.quantizer.quantizers:
        q.inited.fill_(True)
    return model


def encode_signal(rng, n):
    """A smooth-ish signal so that the log-mel has structure: tones under a slow envelope + noise."""
    t = np.arange(n) / 16000.0
    x = 0.3 * np.sin(2 * np.pi * 220 * t) + 0.2 * np.sin(2 * np.pi * 1330 * t + 1.0) + 0.05 * rng.standard_normal(n)
    return (x * (0.5 + 0.5 * np.sin(2 * np.pi * 0.7 * t))).astype(np.float32)


class _QuantizerTap:
    """Forward pre-hook on the reference's ResidualVQ: records the pre-RVQ features it is handed and their projection
    (quantizer.input_proj), which the parity test needs to adjudicate code flips in fp64."""

    def __init__(self, model):
        self.pre, self.zin = [], []
        self.model = model
        self.h = model.quantizer.register_forward_pre_hook(self._hook)

    def _hook(self, mod, args):
        z = args[0]
        self.pre.append(z.detach().clone())
        with torch.no_grad():
            self.zin.append(mod.input_proj(z).detach().clone())

    def close(self):
        self.h.remove()


def gen_encode_full():
    """Reference XY_Tokenizer.encode at the SHIPPED config (xy_tokenizer_config.yaml), one 9.6 s item: codes, the
    projected pre-RVQ vectors (quantizer.input_proj output) and a decimated copy of the 3072-d pre-RVQ features."""
    gp, seed = full_codec_params(), 41
    model = build_ref_codec_with_encoder(gp, seed)
    rng = np.random.default_rng(19)
    n = int(9.6 * 16000)
    wav = encode_signal(rng, n)
    tap = _QuantizerTap(model)
    with torch.no_grad():
        codes = model.encode([torch.from_numpy(wav)], overlap_seconds=10, device=torch.device("cpu"))["codes_list"][0]
    tap.close()
    T = codes.shape[1]
    out = {"seed": np.int64(seed), "sig_seed": np.int64(19), "n": np.int64(n), "codes": codes.numpy().astype(np.int16),
           "zin": tap.zin[0][0, :, :T].numpy().astype(np.float32),                 # (512, T)
           "pre_sub": tap.pre[0][0, ::16, :T].numpy().astype(np.float32)}          # (192, T)
    print("encode full", codes.shape, out["zin"].shape, float(np.abs(out["zin"]).max()))
    np.savez_compressed(os.path.join(GOLD, "codec_encode_full.npz"), **out)


def gen_encode():
    """Reference XY_Tokenizer.encode on CPU (tiny config): codes, and the pre-RVQ features / mel of one chunk."""
    gp, seed = TINY_CODEC, 33
    model = build_ref_codec_with_encoder(gp, seed)
    rng = np.random.default_rng(9)
    out = {"seed": np.int64(seed)}
    sig = lambda n: encode_signal(rng, n)
    wavs = [sig(35 * 16000), sig(3 * 16000)]
    with torch.no_grad():
        codes = model.encode([torch.from_numpy(w) for w in wavs], overlap_seconds=10, device=torch.device("cpu"))["codes_list"]
        # one-chunk internals for the oracle pin
        x = torch.zeros(2, 1, 480000)
        x[0, 0] = torch.from_numpy(wavs[0][:480000])
        x[1, 0, :48000] = torch.from_numpy(wavs[1])
        lens = torch.tensor([480000, 48000])
        feats = model.feature_extractor([x[0, 0].numpy(), x[1, 0, :48000].numpy()], sampling_rate=16000, return_tensors="pt",
                                        return_attention_mask=True)
        tap = _QuantizerTap(model)
        tok = model.inference_tokenize(x, lens)
        tap.close()
    out["chunk_zin"] = tap.zin[0].numpy().astype(np.float32)                            # (2, 64, 375)
    out["chunk_pre_sub"] = tap.pre[0][:, ::8].numpy().astype(np.float32)                # (2, 64, 375)
    for i, (w, c) in enumerate(zip(wavs, codes)):
        out[f"len{i}"] = np.int64(len(w))
        out[f"codes{i}"] = c.numpy().astype(np.int16)
        print("encode", i, len(w), c.shape)
    out["mel_sub"] = feats["input_features"].numpy()[:, :, ::25].astype(np.float32)   # decimated in time
    out["mel_frames"] = feats["attention_mask"].sum(-1).numpy()
    out["chunk_codes"] = tok["codes"].numpy().astype(np.int16)
    out["chunk_code_lens"] = tok["codes_lengths"].numpy()
    np.savez_compressed(os.path.join(GOLD, "codec_encode.npz"), **out)


def main():
    gen_encode()
    gen_encode_full()
    out = {}
    cases = {
        "tiny": (TINY_CODEC, 21, [30, 11]),           # one window, ragged batch
        "tiny_long": (TINY_CODEC, 21, [400, 120]),    # two windows (375/250 chunking), SURVEY §4 sizes
        "full": (full_codec_params(), 5, [16]),       # the shipped config, one short item
    }
    for name, (gp, seed, lens) in cases.items():
        model = build_ref_codec(gp, seed)
        rng = np.random.default_rng(100 + len(name))
        K = gp["quantizer_kwargs"]["codebook_size"]
        codes = [torch.from_numpy(rng.integers(0, K, (8, n)).astype(np.int64)) for n in lens]
        with torch.no_grad():
            res = model.decode(codes, overlap_seconds=10, device=torch.device("cpu"))["syn_wav_list"]
        for i, (c, w) in enumerate(zip(codes, res)):
            out[f"{name}_codes{i}"] = c.numpy().astype(np.int16)
            wv = w.numpy().astype(np.float32)
            # the two-window case is stored decimated (every 8th sample) to keep the fixture small; the chunk
            # seam at 250 codes = sample 480000 is covered because 480000 % 8 == 0
            out[f"{name}_wav{i}"] = wv[::8] if name == "tiny_long" else wv
            print(name, i, c.shape, w.shape, float(w.abs().max()), float(w.std()))
        out[f"{name}_seed"] = np.int64(seed)
    np.savez_compressed(os.path.join(GOLD, "codec_decode.npz"), **out)


if __name__ == "__main__":
    main()
