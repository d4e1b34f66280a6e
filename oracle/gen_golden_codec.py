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


def main():
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
