"""Oracle (TEST INFRASTRUCTURE, never shipped): CPU restatement of the LM side of the hot path in plain torch.

Follows, op for op and with the same dtype at every intermediate:
  embed_sum            <- /root/reference/modeling_asteroid.py:235-250
  rmsnorm, mlp, rope,
  eager attention,
  decoder layer        <- transformers Qwen3 (third-party, pinned 4.53.2 by the reference's requirements.txt:3; the
                          arithmetic of the installed 5.5.0 modeling_qwen3.py:50-66,70-83,150-219,236-290,300-335
                          is identical) as invoked from modeling_asteroid.py:226,273-284
  heads                <- modeling_asteroid.py:412
  processors           <- HF RepetitionPenaltyLogitsProcessor / TemperatureLogitsWarper / TopKLogitsWarper /
                          TopPLogitsWarper (constructed at modeling_asteroid.py:99-106)
  sample_loop          <- CustomMixin._sample, modeling_asteroid.py:83-169
Parity status: PINNED against outputs of the reference itself (tests/golden/lm_*.npz from oracle/gen_golden.py,
checked in tests/test_oracle_pin.py).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F


def make_weights(shape: dict, seed: int, std: float = 0.02, speech_only_head0: bool = True, tied: bool = False,
                 norm_jitter: float = 0.1) -> Dict[str, torch.Tensor]:
    """Deterministic fp32 weights under the reference's state-dict key names (numpy PCG64, so the GPU box
    regenerates bit-identical values). `speech_only_head0` zeroes the non-speech rows of lm_heads.0
    so that greedy decoding stays inside the speech range for a fixed horizon."""
    rng = np.random.default_rng(seed)
    H, I, L = shape["hidden_size"], shape["intermediate_size"], shape["num_hidden_layers"]
    Hq, Hkv, D = shape["num_attention_heads"], shape["num_key_value_heads"], shape["head_dim"]
    V, Vs, C = shape["vocab_size"], shape["speech_vocab_size"], shape["channels"]

    def n(*s):
        return torch.from_numpy((rng.standard_normal(s, dtype=np.float32) * np.float32(std)))

    def nw(k):
        return torch.from_numpy((1.0 + norm_jitter * rng.standard_normal(k, dtype=np.float32)).astype(np.float32))

    sd = {}
    for c in range(C):
        sd[f"lm_heads.{c}.weight"] = n(V if c == 0 else Vs, H)
    if speech_only_head0:
        lo, hi = shape["speech_token_range"]
        w = sd["lm_heads.0.weight"]
        w[:lo] = 0
        w[hi:] = 0
    for c in range(C):
        sd[f"model.embedding_list.{c}.weight"] = sd[f"lm_heads.{c}.weight"] if tied else n(V if c == 0 else Vs, H)
    p = "model.language_model."
    for l in range(L):
        b = f"{p}layers.{l}."
        sd[b + "input_layernorm.weight"] = nw(H)
        sd[b + "post_attention_layernorm.weight"] = nw(H)
        sd[b + "self_attn.q_norm.weight"] = nw(D)
        sd[b + "self_attn.k_norm.weight"] = nw(D)
        sd[b + "self_attn.q_proj.weight"] = n(Hq * D, H)
        sd[b + "self_attn.k_proj.weight"] = n(Hkv * D, H)
        sd[b + "self_attn.v_proj.weight"] = n(Hkv * D, H)
        sd[b + "self_attn.o_proj.weight"] = n(H, Hq * D)
        sd[b + "mlp.gate_proj.weight"] = n(I, H)
        sd[b + "mlp.up_proj.weight"] = n(I, H)
        sd[b + "mlp.down_proj.weight"] = n(H, I)
    sd[p + "norm.weight"] = nw(H)
    return sd


# Shape / seed / gain of the planted-margin model behind tests/golden/lm_margin.npz (oracle/gen_golden.py gen_lm_margin)
MARGIN_SHAPE = dict(hidden_size=1024, intermediate_size=2048, num_hidden_layers=2, num_attention_heads=8,
                    num_key_value_heads=4, head_dim=128, rms_norm_eps=1e-6, rope_theta=1e6, vocab_size=152697,
                    speech_vocab_size=1025, channels=8, speech_token_range=[151665, 152689])
MARGIN_SEED, MARGIN_GAIN, MARGIN_NEW = 4321, 10.0, 40


def make_planted_weights(shape: dict, seed: int, std: float = 0.02, emb_gain: float = 5.0) -> Dict[str, torch.Tensor]:
    """`make_weights` plus a planted decision margin, for the free-running greedy horizon test.

    A random-init model's top-2 logit gap is below bf16 rounding noise in ~1 decision out of 10, so two correct bf16
    implementations (and the reference's own bf16 vs fp32 runs) part after a handful of rows. Here every channel's
    head copies the channel's embedding rows through a seeded permutation of the 1024 codes
    (lm_heads.c[base + perm_c[j]] = embedding_list.c[base + j]; base = speech_token_range[0] for channel 0), so that
    the winning logit is ~|E|^2 / rms(x) — about sqrt(hidden/8) standard deviations of the competing logits — while
    the decoder layers still perturb every logit (embedding rows are scaled by `emb_gain` so that the embedding sum is
    not drowned by the layers' residual contributions). With hidden >= 1024 the gap is > 5 sigma: the next token of channel c
    is perm_c[previous token of channel c] for the reference and for any correct implementation, and a wrong token
    fed back, a wrong position or a corrupted cache row changes the chain immediately."""
    sd = make_weights(shape, seed, std=std, speech_only_head0=True, tied=False)
    rng = np.random.default_rng(seed + 99991)
    lo, _ = shape["speech_token_range"]
    n_codes = shape["speech_vocab_size"] - 1
    for c in range(shape["channels"]):
        base = lo if c == 0 else 0
        perm = torch.from_numpy(rng.permutation(n_codes))
        emb = sd[f"model.embedding_list.{c}.weight"]
        emb *= emb_gain
        head = sd[f"lm_heads.{c}.weight"]
        head[base + perm] = emb[base:base + n_codes].clone()
    return sd


def rmsnorm(x, w, eps):
    dt = x.dtype
    h = x.to(torch.float32)
    var = h.pow(2).mean(-1, keepdim=True)
    h = h * torch.rsqrt(var + eps)
    return w * h.to(dt)


def rotate_half(x):
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


class OracleLM:
    def __init__(self, shape: dict, sd: Dict[str, torch.Tensor], dtype=torch.float32, device="cpu"):
        """`device="cuda"` runs the same eager torch ops on the GPU (bench.py's `gpu_eager_baseline` leg: the reference's
        own PyTorch-eager path on the same B200); parity tests and the CPU baseline use the default."""
        self.s = shape
        self.dtype = dtype
        self.dev = torch.device(device)
        moved = {}
        self.sd = {}
        for k, v in sd.items():                                   # tied tables stay one tensor
            if id(v) not in moved:
                moved[id(v)] = v.to(self.dev, dtype)
            self.sd[k] = moved[id(v)]
        D = shape["head_dim"]
        self.inv_freq = (1.0 / (shape["rope_theta"] ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))).to(self.dev)

    def embed_sum(self, ids):
        B, S, C = ids.shape
        if C != self.s["channels"]:
            raise ValueError(f"Expected {self.s['channels']} channels, got {C}")
        out = torch.zeros(B, S, self.s["hidden_size"], dtype=self.dtype, device=self.dev)
        for c in range(C):
            out += F.embedding(ids[..., c], self.sd[f"model.embedding_list.{c}.weight"])
        return out

    def hidden(self, ids, attention_mask=None):
        """Full-sequence (teacher-forced) pass with HF eager attention and the additive causal+padding mask."""
        s, sd = self.s, self.sd
        B, S, _ = ids.shape
        if attention_mask is None:
            attention_mask = torch.ones(B, S, device=self.dev)
        am = attention_mask.long()
        pos = am.cumsum(-1) - 1
        pos = pos.masked_fill(am == 0, 1)
        freqs = pos[:, :, None].float() * self.inv_freq[None, None, :].float()
        emb = torch.cat((freqs, freqs), dim=-1)
        cos, sin = emb.cos().to(self.dtype)[:, None], emb.sin().to(self.dtype)[:, None]
        minv = torch.finfo(self.dtype).min
        causal = torch.tril(torch.ones(S, S, dtype=torch.bool, device=self.dev))
        allowed = causal[None, None] & (am[:, None, None, :] != 0)
        add_mask = torch.zeros(B, 1, S, S, dtype=self.dtype, device=self.dev).masked_fill(~allowed, minv)
        x = self.embed_sum(ids)
        Hq, Hkv, D = s["num_attention_heads"], s["num_key_value_heads"], s["head_dim"]
        eps = s["rms_norm_eps"]
        p = "model.language_model."
        for l in range(s["num_hidden_layers"]):
            b = f"{p}layers.{l}."
            res = x
            h = rmsnorm(x, sd[b + "input_layernorm.weight"], eps)
            q = rmsnorm(F.linear(h, sd[b + "self_attn.q_proj.weight"]).view(B, S, Hq, D), sd[b + "self_attn.q_norm.weight"], eps).transpose(1, 2)
            k = rmsnorm(F.linear(h, sd[b + "self_attn.k_proj.weight"]).view(B, S, Hkv, D), sd[b + "self_attn.k_norm.weight"], eps).transpose(1, 2)
            v = F.linear(h, sd[b + "self_attn.v_proj.weight"]).view(B, S, Hkv, D).transpose(1, 2)
            q = (q * cos) + (rotate_half(q) * sin)
            k = (k * cos) + (rotate_half(k) * sin)
            g = Hq // Hkv
            kr = k[:, :, None].expand(B, Hkv, g, S, D).reshape(B, Hq, S, D)
            vr = v[:, :, None].expand(B, Hkv, g, S, D).reshape(B, Hq, S, D)
            w = torch.matmul(q, kr.transpose(2, 3)) * (D ** -0.5)
            w = w + add_mask
            w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
            ao = torch.matmul(w, vr).transpose(1, 2).contiguous().reshape(B, S, Hq * D)
            x = res + F.linear(ao, sd[b + "self_attn.o_proj.weight"])
            res = x
            h = rmsnorm(x, sd[b + "post_attention_layernorm.weight"], eps)
            h = F.linear(F.silu(F.linear(h, sd[b + "mlp.gate_proj.weight"])) * F.linear(h, sd[b + "mlp.up_proj.weight"]),
                         sd[b + "mlp.down_proj.weight"])
            x = res + h
        return rmsnorm(x, sd[p + "norm.weight"], eps)

    def logits_all(self, ids, attention_mask=None, last_only=False):
        h = self.hidden(ids, attention_mask)
        if last_only:
            h = h[:, -1:]
        return [F.linear(h, self.sd[f"lm_heads.{c}.weight"]) for c in range(self.s["channels"])]


# ------------------------------------------------------------------------------------------------ processors
def repetition_penalty(input_ids, scores, penalty):
    score = torch.gather(scores, 1, input_ids)
    score = torch.where(score < 0, score * penalty, score / penalty)
    return scores.scatter(1, input_ids, score)


def temperature(scores, t):
    return scores / t


def top_k(scores, k, filter_value=-float("inf")):
    k = min(k, scores.size(-1))
    remove = scores < torch.topk(scores, k)[0][..., -1, None]
    return scores.masked_fill(remove, filter_value)


def top_p(scores, p, filter_value=-float("inf"), min_tokens_to_keep=1):
    sorted_logits, sorted_indices = torch.sort(scores, descending=False)
    cum = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
    remove_sorted = cum <= (1 - p)
    remove_sorted[..., -min_tokens_to_keep:] = 0
    remove = remove_sorted.scatter(1, sorted_indices, remove_sorted)
    return scores.masked_fill(remove, filter_value)


def apply_processors(history, scores, layer_cfg):
    if layer_cfg.get("repetition_penalty") is not None:
        scores = repetition_penalty(history, scores, layer_cfg["repetition_penalty"])
    if layer_cfg.get("temperature") is not None:
        scores = temperature(scores, layer_cfg["temperature"])
    if layer_cfg.get("top_k") is not None:
        scores = top_k(scores, layer_cfg["top_k"])
    if layer_cfg.get("top_p") is not None:
        scores = top_p(scores, layer_cfg["top_p"])
    return scores


# ------------------------------------------------------------------------------------------------ sampler loop
def sample_loop(logits_fn, input_ids, max_length, speech_range, eos_token_id=152694, has_eos_criteria=True,
                layers=None, do_samples=None, draw_fn=None, channels=8, speech_pad=1024, eos_mask_idx=152694,
                record=None):
    """CustomMixin._sample restated (modeling_asteroid.py:83-169). `logits_fn(ids (B,L,8)) -> list of 8 (B,V_c)`
    last-position logits; `draw_fn(step, channel, probs)` replaces torch.multinomial for sampled channels."""
    B, T, C = input_ids.shape
    dev = input_ids.device
    unfinished = torch.ones(B, dtype=torch.long, device=dev)
    needs = -1 * torch.ones(B, dtype=torch.long, device=dev)
    tf = input_ids
    ids = input_ids[:, :-(C - 1)]
    base = ids.shape[1]
    layers = layers or [{} for _ in range(C)]
    do_samples = do_samples or [False] * C
    lo, hi = speech_range
    step = 0
    while True:
        logits = [l.clone().float() for l in logits_fn(ids)]
        for i, cl in enumerate(logits):
            if i != 0 and ids.shape[1] + 1 > tf.shape[1] - 7 + i:
                cl[:, speech_pad] = -torch.inf
            if i == 0 and ids.shape[1] + 1 <= tf.shape[1]:
                cl[:, eos_mask_idx] = -torch.inf
        scores = [apply_processors(ids[..., i], l, layers[i]) for i, l in enumerate(logits)]
        if record is not None:
            record.append([s.clone() for s in scores])
        toks = []
        for i, sc in enumerate(scores):
            if do_samples[i]:
                probs = F.softmax(sc, dim=-1)
                toks.append(draw_fn(step, i, probs) if draw_fn else torch.multinomial(probs, 1).squeeze(1))
            else:
                toks.append(torch.argmax(sc, dim=-1))
        nt = torch.stack(toks, dim=-1)
        idx = (~((nt[:, 0] >= lo) & (nt[:, 0] < hi))) & (needs < 0)
        needs[idx] = C - 1
        if ids.shape[1] + 1 <= tf.shape[1]:
            i = ids.shape[1] + 1 - base
            nt[:, i:] = tf[:, ids.shape[1], i:]
        mask = (needs > 0) & (needs < 7)
        if mask.any():
            nt[mask, 0] = eos_token_id
            for i in range(1, C):
                mi = mask & (needs < C - i)
                nt[mi, i] = speech_pad
        if has_eos_criteria:
            for i in range(C):
                pd = eos_token_id if i == 0 else speech_pad
                nt[:, i] = nt[:, i] * unfinished + pd * (1 - unfinished)
        ids = torch.cat([ids, nt[:, None, :]], dim=1)
        needs = torch.where(needs > 0, needs - 1, needs)
        stop = torch.full((B,), ids.shape[1] >= max_length, dtype=torch.bool, device=dev)
        if has_eos_criteria:
            stop = stop | (ids[:, -1, 0] == eos_token_id)
        stop = stop | (needs == 0)
        unfinished = unfinished & ~stop
        unfinished = unfinished | (needs > 0)
        step += 1
        if unfinished.max() == 0:
            break
    return ids


# ------------------------------------------------------------------------------------------------ cached decode
class OracleCachedLM(OracleLM):
    """The same arithmetic with a per-layer K/V cache (what the reference does through HF DynamicCache): used as the
    timed CPU baseline in bench.py and pinned against the full-recompute path in tests/test_oracle_pin.py."""

    def __init__(self, shape, sd, dtype=torch.float32, device="cpu"):
        super().__init__(shape, sd, dtype, device)
        self.k, self.v, self.mask = None, None, None

    def reset(self):
        self.k, self.v, self.mask = None, None, None

    def step(self, ids, attention_mask):
        """ids (B, S, 8): the new rows; attention_mask (B, past + S) over everything so far -> last-position logits."""
        s, sd = self.s, self.sd
        B, S, _ = ids.shape
        am = attention_mask.long()
        total = am.shape[1]
        past = total - S
        pos = (am.cumsum(-1) - 1).masked_fill(am == 0, 1)[:, past:]
        freqs = pos[:, :, None].float() * self.inv_freq[None, None, :].float()
        emb = torch.cat((freqs, freqs), dim=-1)
        cos, sin = emb.cos().to(self.dtype)[:, None], emb.sin().to(self.dtype)[:, None]
        minv = torch.finfo(self.dtype).min
        qi = torch.arange(past, total, device=self.dev)[:, None]
        ki = torch.arange(total, device=self.dev)[None, :]
        allowed = (ki <= qi)[None, None] & (am[:, None, None, :] != 0)
        add_mask = torch.zeros(B, 1, S, total, dtype=self.dtype, device=self.dev).masked_fill(~allowed, minv)
        x = self.embed_sum(ids)
        Hq, Hkv, D = s["num_attention_heads"], s["num_key_value_heads"], s["head_dim"]
        eps = s["rms_norm_eps"]
        p = "model.language_model."
        L = s["num_hidden_layers"]
        if self.k is None:
            self.k, self.v = [None] * L, [None] * L
        for l in range(L):
            b = f"{p}layers.{l}."
            res = x
            h = rmsnorm(x, sd[b + "input_layernorm.weight"], eps)
            q = rmsnorm(F.linear(h, sd[b + "self_attn.q_proj.weight"]).view(B, S, Hq, D), sd[b + "self_attn.q_norm.weight"], eps).transpose(1, 2)
            k = rmsnorm(F.linear(h, sd[b + "self_attn.k_proj.weight"]).view(B, S, Hkv, D), sd[b + "self_attn.k_norm.weight"], eps).transpose(1, 2)
            v = F.linear(h, sd[b + "self_attn.v_proj.weight"]).view(B, S, Hkv, D).transpose(1, 2)
            q = (q * cos) + (rotate_half(q) * sin)
            k = (k * cos) + (rotate_half(k) * sin)
            self.k[l] = k if self.k[l] is None else torch.cat([self.k[l], k], dim=2)
            self.v[l] = v if self.v[l] is None else torch.cat([self.v[l], v], dim=2)
            g = Hq // Hkv
            kr = self.k[l][:, :, None].expand(B, Hkv, g, total, D).reshape(B, Hq, total, D)
            vr = self.v[l][:, :, None].expand(B, Hkv, g, total, D).reshape(B, Hq, total, D)
            w = torch.matmul(q, kr.transpose(2, 3)) * (D ** -0.5) + add_mask
            w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
            ao = torch.matmul(w, vr).transpose(1, 2).contiguous().reshape(B, S, Hq * D)
            x = res + F.linear(ao, sd[b + "self_attn.o_proj.weight"])
            res = x
            h = rmsnorm(x, sd[b + "post_attention_layernorm.weight"], eps)
            h = F.linear(F.silu(F.linear(h, sd[b + "mlp.gate_proj.weight"])) * F.linear(h, sd[b + "mlp.up_proj.weight"]),
                         sd[b + "mlp.down_proj.weight"])
            x = res + h
        hl = rmsnorm(x, sd[p + "norm.weight"], eps)[:, -1]
        return [F.linear(hl, sd[f"lm_heads.{c}.weight"]) for c in range(s["channels"])]

    def generate(self, input_ids, attention_mask, max_length, speech_range, **kw):
        """Greedy/sampled generation with the cache, driving `sample_loop` (the restated `_sample`)."""
        self.reset()
        P = input_ids.shape[1] - (self.s["channels"] - 1)
        state = dict(fed=0)

        def logits_fn(cur):
            am = torch.cat([attention_mask[:, :P], torch.ones(cur.shape[0], cur.shape[1] - P, dtype=attention_mask.dtype,
                                                              device=attention_mask.device)], 1)
            new = cur[:, state["fed"]:]
            state["fed"] = cur.shape[1]
            with torch.no_grad():
                return self.step(new, am)

        return sample_loop(logits_fn, input_ids, max_length, speech_range, **kw)


def random_weights_fast(shape: dict, seed: int = 0, std: float = 0.02, dtype=torch.float32):
    """Full-size random weights for the timed CPU baseline (torch RNG, no parity role)."""
    g = torch.Generator().manual_seed(seed)
    H, I, L = shape["hidden_size"], shape["intermediate_size"], shape["num_hidden_layers"]
    Hq, Hkv, D = shape["num_attention_heads"], shape["num_key_value_heads"], shape["head_dim"]
    V, Vs, C = shape["vocab_size"], shape["speech_vocab_size"], shape["channels"]
    n = lambda *s: (torch.randn(*s, generator=g, dtype=torch.float32) * std).to(dtype)
    sd = {}
    for c in range(C):
        sd[f"lm_heads.{c}.weight"] = n(V if c == 0 else Vs, H)
        sd[f"model.embedding_list.{c}.weight"] = sd[f"lm_heads.{c}.weight"]
    lo, hi = shape["speech_token_range"]
    sd["lm_heads.0.weight"][:lo] = 0
    sd["lm_heads.0.weight"][hi:] = 0
    p = "model.language_model."
    for l in range(L):
        b = f"{p}layers.{l}."
        sd[b + "input_layernorm.weight"] = torch.ones(H, dtype=dtype)
        sd[b + "post_attention_layernorm.weight"] = torch.ones(H, dtype=dtype)
        sd[b + "self_attn.q_norm.weight"] = torch.ones(D, dtype=dtype)
        sd[b + "self_attn.k_norm.weight"] = torch.ones(D, dtype=dtype)
        sd[b + "self_attn.q_proj.weight"] = n(Hq * D, H)
        sd[b + "self_attn.k_proj.weight"] = n(Hkv * D, H)
        sd[b + "self_attn.v_proj.weight"] = n(Hkv * D, H)
        sd[b + "self_attn.o_proj.weight"] = n(H, Hq * D)
        sd[b + "mlp.gate_proj.weight"] = n(I, H)
        sd[b + "mlp.up_proj.weight"] = n(I, H)
        sd[b + "mlp.down_proj.weight"] = n(H, I)
    sd[p + "norm.weight"] = torch.ones(H, dtype=dtype)
    return sd
