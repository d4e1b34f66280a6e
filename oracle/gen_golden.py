"""TEST INFRASTRUCTURE. Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only:   python -m oracle.gen_golden [lm] [sampler] [rvq] [codec] [utils]
The fixtures hold inputs and the reference's outputs; weights are regenerated from a seed by
oracle.lm_oracle.make_weights / oracle.codec_oracle.make_codec_weights (numpy PCG64 -> identical on any host).
"""
import os
import sys

import numpy as np
import torch

from oracle import ref_shims
from oracle import lm_oracle

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TINY = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
            head_dim=128, rms_norm_eps=1e-6, rope_theta=1e6, vocab_size=152697, speech_vocab_size=1025, channels=8,
            speech_token_range=[151665, 152689])
TINY_SEED = 1234


def ref_cfg_kwargs(shape):
    kw = {k: shape[k] for k in ("hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
                                "num_key_value_heads", "head_dim", "rms_norm_eps", "vocab_size", "speech_vocab_size",
                                "channels", "speech_token_range")}
    kw.update(speech_pad_token=1024, eos_token_id=152694, pad_token_id=151643, max_position_embeddings=32768,
              rope_parameters={"rope_type": "default", "rope_theta": shape["rope_theta"]}, attention_bias=False)
    return kw


def make_prompt(rng, B, text_rows, audio_rows, shape, pad_token_id=151643):
    """Delay-shifted, left-padded prompt grids like generation_utils.process_inputs/shifting_inputs/rpadding build."""
    lo, hi = shape["speech_token_range"]
    C = shape["channels"]
    grids = []
    for b in range(B):
        nt, na = text_rows[b], audio_rows[b]
        g = np.full((nt + na, C), 1024, dtype=np.int64)
        g[:nt, 0] = rng.integers(0, 151000, nt)
        g[nt:, 0] = rng.integers(lo, hi, na)
        g[nt:, 1:] = rng.integers(0, 1024, (na, C - 1))
        n = g.shape[0]
        sh = np.full((n + C - 1, C), 1024, dtype=np.int64)
        sh[:, 0] = pad_token_id
        for i in range(C):
            sh[i:n + i, i] = g[:, i]
        grids.append(sh)
    T = max(g.shape[0] for g in grids)
    ids = np.full((B, T, C), 1024, dtype=np.int64)
    ids[:, :, 0] = pad_token_id
    mask = np.zeros((B, T), dtype=np.float64)
    for b, g in enumerate(grids):
        ids[b, T - g.shape[0]:] = g
        mask[b, T - g.shape[0]:] = 1
    return ids, mask


def build_ref_lm(shape, seed, dtype, tied=False):
    ma = ref_shims.import_lm()
    model = ref_shims.make_lm(ma, ref_cfg_kwargs(shape), torch.float32)
    sd = lm_oracle.make_weights(shape, seed, tied=tied)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("embed_tokens" in m for m in missing), missing  # only the dead table is absent
    rot = model.model.language_model.rotary_emb
    inv32 = rot.inv_freq.clone()
    model = model.to(dtype).eval()
    # `from_pretrained(torch_dtype=bf16)` (generation_utils.py:18) keeps the non-persistent inv_freq buffer in fp32
    # (it is created with an explicit float dtype); a blanket .to(bf16) would round it, which is a harness artefact.
    rot.inv_freq = inv32
    rot.original_inv_freq = inv32.clone()
    ref_shims.bind_generation_helpers(model)
    inv = model.model.language_model.rotary_emb.inv_freq
    want = 1.0 / (shape["rope_theta"] ** (torch.arange(0, shape["head_dim"], 2, dtype=torch.int64).float() / shape["head_dim"]))
    assert torch.equal(inv.float(), want), "rope theta not honoured by the installed transformers"
    return model, sd


def gen_lm():
    shape = TINY
    rng = np.random.default_rng(7)
    ids, mask = make_prompt(rng, 2, [9, 5], [6, 4], shape)
    out = dict(ids=ids, mask=mask)
    lo, hi = shape["speech_token_range"]
    for name, dtype in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        model, _ = build_ref_lm(shape, TINY_SEED, dtype)
        with torch.no_grad():
            o = model(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask), return_dict=True)
        la = [l.float().numpy() for l in o.logits_all]
        # keep fixtures small: ch0 speech-range slice + full ch1-7, at the last 4 positions
        out[f"logits0_speech_{name}"] = la[0][:, -4:, lo:hi]
        out[f"logits0_eos_{name}"] = la[0][:, -4:, 152694]
        out[f"logits17_{name}"] = np.stack([l[:, -4:] for l in la[1:]], axis=0)
        # greedy generation: 24 new rows
        T = ids.shape[1]
        seq = ref_shims.run_sample(model, torch.from_numpy(ids), torch.from_numpy(mask), max_length=T + 24)
        out[f"greedy_{name}"] = seq.numpy()
        print(name, "greedy rows", seq.shape, "ch0 in speech range:",
              bool(((seq[:, T - 7:, 0] >= lo) & (seq[:, T - 7:, 0] < hi)).all()))
    np.savez_compressed(os.path.join(GOLD, "lm_tiny.npz"), **out)


MARGIN, MARGIN_SEED, MARGIN_GAIN, MARGIN_NEW = (lm_oracle.MARGIN_SHAPE, lm_oracle.MARGIN_SEED, lm_oracle.MARGIN_GAIN,
                                                 lm_oracle.MARGIN_NEW)


def build_ref_lm_sd(shape, sd, dtype):
    """The reference model loaded with an explicit state dict (same harness care as build_ref_lm)."""
    ma = ref_shims.import_lm()
    model = ref_shims.make_lm(ma, ref_cfg_kwargs(shape), torch.float32)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("embed_tokens" in m for m in missing), missing
    rot = model.model.language_model.rotary_emb
    inv32 = rot.inv_freq.clone()
    model = model.to(dtype).eval()
    rot.inv_freq = inv32
    rot.original_inv_freq = inv32.clone()
    ref_shims.bind_generation_helpers(model)
    return model


def gen_lm_margin():
    """Free-running greedy horizon with a planted decision margin (lm_oracle.make_planted_weights): the reference's
    bf16 and fp32 `_sample` runs must agree with each other over all MARGIN_NEW rows, and so must any correct
    implementation."""
    shape = MARGIN
    rng = np.random.default_rng(8)
    ids, mask = make_prompt(rng, 2, [7, 4], [14, 11], shape)
    sd = lm_oracle.make_planted_weights(shape, MARGIN_SEED, emb_gain=MARGIN_GAIN)
    out = dict(ids=ids, mask=mask, seed=np.int64(MARGIN_SEED), gain=np.float64(MARGIN_GAIN))
    T = ids.shape[1]
    lo, hi = shape["speech_token_range"]
    for name, dtype in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        model = build_ref_lm_sd(shape, sd, dtype)
        seq = ref_shims.run_sample(model, torch.from_numpy(ids), torch.from_numpy(mask), max_length=T + MARGIN_NEW)
        out[f"greedy_{name}"] = seq.numpy()
        with torch.no_grad():
            o = model(input_ids=seq, attention_mask=torch.cat([torch.from_numpy(mask)[:, :T - 7],
                                                               torch.ones(2, seq.shape[1] - (T - 7), dtype=torch.float64)], 1),
                      return_dict=True)
        la = [l.float() for l in o.logits_all]
        gaps = []
        for c in range(8):
            l = la[c][:, T - 7 - 1 + 7:-1]                      # positions that predict rows where every channel is free
            l = l[..., lo:hi] if c == 0 else l[..., :1024]
            t2 = l.topk(2, -1).values
            gaps.append((t2[..., 0] - t2[..., 1]).min().item())
        out[f"min_gap_{name}"] = np.float64(min(gaps))
        print("margin", name, seq.shape, "min top-2 gap", min(gaps))
    assert np.array_equal(out["greedy_f32"], out["greedy_bf16"]), "the reference's own bf16 and fp32 runs disagree"
    np.savez_compressed(os.path.join(GOLD, "lm_margin.npz"), **out)


LONG_ROWS = 12100


def gen_lm_longctx():
    """RoPE / attention range: reference bf16 + fp32 last-position logits of the TINY model on a 12.1 k-row prompt."""
    shape = TINY
    rng = np.random.default_rng(21)
    ids, mask = make_prompt(rng, 1, [100], [LONG_ROWS - 100 - 7], shape)
    assert ids.shape[1] == LONG_ROWS
    out = dict(seed=np.int64(21), rows=np.int64(LONG_ROWS))
    lo, hi = shape["speech_token_range"]
    for name, dtype in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        model, _ = build_ref_lm(shape, TINY_SEED, dtype)
        with torch.no_grad():
            h = model.model(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask)).last_hidden_state[:, -4:]
            la = [head(h).float().numpy() for head in model.lm_heads]
        out[f"logits0_speech_{name}"] = la[0][..., lo:hi]
        out[f"logits17_{name}"] = np.stack(la[1:], 0)
        print("longctx", name, la[1].shape)
    np.savez_compressed(os.path.join(GOLD, "lm_longctx.npz"), **out)


class _Scripted(torch.nn.Module):
    """Stands in for the network inside the reference's `_sample`: returns scripted last-position logits."""


def gen_sampler():
    """Wind-down / finished-row behaviour of the reference `_sample` with scripted logits, plus HF processor KATs."""
    ma = ref_shims.import_lm()
    shape = TINY
    model, _ = build_ref_lm(shape, TINY_SEED, torch.float32)
    rng = np.random.default_rng(11)
    ids, mask = make_prompt(rng, 3, [6, 4, 5], [3, 2, 0], shape)
    B, T, C = ids.shape
    lo, hi = shape["speech_token_range"]
    # script: row 0 emits a non-speech ch0 token (EOS) at step 9, row 1 emits a TEXT token at step 3 (inside the
    # teacher-forced window), row 2 never stops (max_length cuts it)
    n_steps = 40
    script = rng.integers(0, 1024, (n_steps, B, C)).astype(np.int64)
    script[:, :, 0] += lo
    script[9, 0, 0] = 152694
    script[3, 1, 0] = 777
    vocabs = [shape["vocab_size"]] + [shape["speech_vocab_size"]] * (C - 1)
    state = dict(step=0)
    from transformers.cache_utils import DynamicCache

    def fake_forward(**kw):
        s = state["step"]
        state["step"] += 1
        S = kw["input_ids"].shape[1]
        logits_all = []
        for c in range(C):
            l = torch.zeros(B, S, vocabs[c])
            l[torch.arange(B), -1, torch.from_numpy(script[s, :, c])] = 10.0
            # a competing pad / EOS logit that the masks must suppress when the rule says so
            if c > 0:
                l[:, -1, 1024] = 11.0 if s % 2 == 0 else 0.0
            logits_all.append(l)
        cache = kw.get("past_key_values") or DynamicCache(config=model.config)
        return ma.AsteroidTTSOutputWithPast(logits=logits_all[0], logits_all=logits_all, past_key_values=cache)

    model.forward = fake_forward
    model.__class__.__call__ = lambda self, **kw: fake_forward(**kw)
    try:
        seq = ref_shims.run_sample(model, torch.from_numpy(ids), torch.from_numpy(mask), max_length=T + 20)
    finally:
        del model.__class__.__call__
    out = dict(ids=ids, mask=mask, script=script, seq=seq.numpy(), max_length=np.int64(T + 20))
    print("scripted wind-down trace:", seq.shape)

    # HF processor known-answer vectors
    from transformers.generation.logits_process import (RepetitionPenaltyLogitsProcessor, TemperatureLogitsWarper,
                                                        TopKLogitsWarper, TopPLogitsWarper)
    g = torch.Generator().manual_seed(5)
    scores = torch.randn(4, 1025, generator=g) * 3
    hist = torch.randint(0, 1025, (4, 37), generator=g)
    out["proc_scores"] = scores.numpy()
    out["proc_hist"] = hist.numpy()
    out["proc_rep"] = RepetitionPenaltyLogitsProcessor(penalty=1.3)(hist, scores.clone()).numpy()
    out["proc_temp"] = TemperatureLogitsWarper(temperature=0.8)(hist, scores.clone()).numpy()
    out["proc_topk"] = TopKLogitsWarper(top_k=50)(hist, scores.clone()).numpy()
    out["proc_topp"] = TopPLogitsWarper(top_p=0.9)(hist, scores.clone()).numpy()
    chain = scores.clone()
    for pr in (RepetitionPenaltyLogitsProcessor(penalty=1.1), TemperatureLogitsWarper(temperature=0.9),
               TopKLogitsWarper(top_k=40), TopPLogitsWarper(top_p=0.85)):
        chain = pr(hist, chain)
    out["proc_chain"] = chain.numpy()
    np.savez_compressed(os.path.join(GOLD, "sampler_trace.npz"), **out)


def gen_rvq():
    from oracle.codec_weights import make_rvq_weights
    XY, qmod, _ = ref_shims.import_codec()
    out = {}
    for name, (B, T, din, D, K, nq) in {"small": (2, 19, 96, 64, 128, 3), "full": (2, 12, 3072, 512, 1024, 8)}.items():
        rvq = qmod.ResidualVQ(input_dim=din, rvq_dim=D, output_dim=din, num_quantizers=nq, codebook_size=K,
                              codebook_dim=D, quantizer_dropout=0.0).eval()
        w = make_rvq_weights(din, D, din, nq, K, seed=3)
        sd = rvq.state_dict()
        for k, v in w.items():
            sd[k] = torch.from_numpy(v)
        rvq.load_state_dict(sd)
        for q in rvq.quantizers:  # a random-init codebook is all zero and "not inited"
            q.inited.fill_(True)
        rng = np.random.default_rng(17)
        z = torch.from_numpy(rng.standard_normal((B, din, T)).astype(np.float32))
        lengths = torch.tensor([T, T - 5])
        with torch.no_grad():
            zq, codes, _, allq, _ = rvq(z, lengths)
            z_in = rvq.input_proj(z)
            dec = rvq.decode_codes(codes)
        out.update({f"{name}_dims": np.array([B, T, din, D, K, nq]), f"{name}_z": z.numpy(),
                    f"{name}_lengths": lengths.numpy(), f"{name}_codes": codes.numpy(), f"{name}_zq_out": zq.numpy(),
                    f"{name}_z_in": z_in.numpy(), f"{name}_decode": dec.numpy()})
        print("rvq", name, codes.shape, zq.shape)
    np.savez_compressed(os.path.join(GOLD, "rvq.npz"), **out)


def main(argv):
    os.makedirs(GOLD, exist_ok=True)
    what = set(argv) or {"lm", "sampler", "rvq", "codec", "utils"}
    if "lm" in what:
        gen_lm()
    if "lm_margin" in what:
        gen_lm_margin()
    if "lm_longctx" in what:
        gen_lm_longctx()
    if "sampler" in what:
        gen_sampler()
    if "rvq" in what:
        gen_rvq()
    if "codec" in what:
        from oracle import gen_golden_codec
        gen_golden_codec.main()
    if "utils" in what:
        from oracle import gen_golden_utils
        gen_golden_utils.main()


if __name__ == "__main__":
    assert ref_shims.available(), "the reference is not mounted; fixtures can only be generated in the build container"
    main(sys.argv[1:])
