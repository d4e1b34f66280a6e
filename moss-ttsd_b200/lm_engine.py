"""Decoder engine: HBM layout of the LM weights, the (paged or contiguous) KV cache, and the launch sequence of
one prefill / one decode step. All arithmetic is in libmtts kernels; torch only owns memory, streams and the
CUDA graph that replays the ~230 launches of a decode step.

Reference behaviour being replaced: AsteroidTTSInstruct.forward + HF Qwen3Model (modeling_asteroid.py:252-285,
337-426) and the per-step part of CustomMixin._sample (modeling_asteroid.py:110-169).

HBM layout (one replica per GPU):
  heads   [Vpad, H] bf16   the 8 LM heads stacked (each head's rows padded to a multiple of 32); when the
                           checkpoint ties heads and embedding tables (tie_weights, modeling_asteroid.py:315-317)
                           the 8 embedding tables are views into this buffer
  layer l : wqkv [(Hq+2Hkv)*D, H]   q|k|v rows stacked -> one GEMM
            wo   [H, Hq*D]
            wgu  [2I, H]            gate/up rows interleaved (2j = gate_j, 2j+1 = up_j) -> SwiGLU in the epilogue
            wd   [H, I]
            ln1, ln2 [H], q_norm, k_norm [D]
  KV pool per layer: k, v [num_pages, Hkv, page_size, D] bf16; left-pad rows are never stored.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from . import _lib, ops
from ._lib import check, ptr, stream_ptr


def _pad_rows(n: int) -> int:
    """Rows each LM head is padded to in the stacked head matrix: a multiple of 32, so that no 32-row quarter of a
    weight tile straddles two channels (the fused heads + greedy pick reports candidates per quarter) and every
    channel's logits start 16-byte aligned."""
    return (n + 31) // 32 * 32


@dataclass
class LMShape:
    hidden_size: int = 2048
    intermediate_size: int = 6144
    num_hidden_layers: int = 28
    num_attention_heads: int = 16
    num_key_value_heads: int = 8
    head_dim: int = 128
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1e6
    vocab_size: int = 152697
    speech_vocab_size: int = 1025
    channels: int = 8

    @classmethod
    def from_config(cls, cfg):
        g = lambda k, d=None: getattr(cfg, k, d)
        hd = g("head_dim") or g("hidden_size") // g("num_attention_heads")
        rope_theta = g("rope_theta")
        if rope_theta is None:
            rp = g("rope_parameters") or {}
            rope_theta = rp.get("rope_theta", 1e6) if isinstance(rp, dict) else 1e6
        return cls(g("hidden_size"), g("intermediate_size"), g("num_hidden_layers"), g("num_attention_heads"),
                   g("num_key_value_heads"), hd, g("rms_norm_eps", 1e-6), float(rope_theta), g("vocab_size"),
                   g("speech_vocab_size", 1025), g("channels", 8))

    @property
    def vocabs(self) -> List[int]:
        return [self.vocab_size] + [self.speech_vocab_size] * (self.channels - 1)

    @property
    def head_offsets(self) -> List[int]:
        offs, o = [], 0
        for v in self.vocabs:
            offs.append(o)
            o += _pad_rows(v)
        return offs

    @property
    def vpad(self) -> int:
        return sum(_pad_rows(v) for v in self.vocabs)


class LMWeights:
    """Weights in the kernel layout, built from a state dict that uses the reference's key names."""

    def __init__(self, shape: LMShape, device):
        self.shape = shape
        self.device = torch.device(device)
        s = shape
        bf = dict(dtype=torch.bfloat16, device=self.device)
        self.heads = torch.zeros((s.vpad, s.hidden_size), **bf)
        self.embeds: Optional[torch.Tensor] = None  # separate tables when not tied
        self.tied = True
        self.layers = []
        for _ in range(s.num_hidden_layers):
            self.layers.append(dict(
                ln1=torch.ones(s.hidden_size, **bf), ln2=torch.ones(s.hidden_size, **bf),
                q_norm=torch.ones(s.head_dim, **bf), k_norm=torch.ones(s.head_dim, **bf),
                wqkv=torch.empty(((s.num_attention_heads + 2 * s.num_key_value_heads) * s.head_dim, s.hidden_size), **bf),
                wo=torch.empty((s.hidden_size, s.num_attention_heads * s.head_dim), **bf),
                wgu=torch.empty((2 * s.intermediate_size, s.hidden_size), **bf),
                wd=torch.empty((s.hidden_size, s.intermediate_size), **bf)))
        self.final_norm = torch.ones(s.hidden_size, **bf)
        # inv_freq exactly as HF computes it (fp32 pow on the host), modeling_qwen3.py rotary embedding init
        inv = 1.0 / (s.rope_theta ** (torch.arange(0, s.head_dim, 2, dtype=torch.int64).float() / s.head_dim))
        self.inv_freq = inv.to(self.device, torch.float32).contiguous()

    # ---- views
    def head_view(self, c: int) -> torch.Tensor:
        o, v = self.shape.head_offsets[c], self.shape.vocabs[c]
        return self.heads[o:o + v]

    def embed_view(self, c: int) -> torch.Tensor:
        if self.embeds is None:
            return self.head_view(c)
        o, v = self.shape.head_offsets[c], self.shape.vocabs[c]
        return self.embeds[o:o + v]

    def nbytes(self) -> int:
        n = self.heads.numel() + (self.embeds.numel() if self.embeds is not None else 0) + self.final_norm.numel()
        for L in self.layers:
            n += sum(t.numel() for t in L.values())
        return n * 2

    def load_state_dict(self, sd: dict, tie_word_embeddings: Optional[bool] = None):
        """Accepts the reference's keys: model.embedding_list.{i}.weight, model.language_model.layers.{l}.*,
        model.language_model.norm.weight, lm_heads.{i}.weight (SURVEY.md §5 checkpoint row)."""
        s = self.shape
        dev = self.device

        def get(k):
            t = sd[k]
            return t.to(dev, torch.bfloat16) if isinstance(t, torch.Tensor) else torch.as_tensor(t).to(dev, torch.bfloat16)

        have_heads = all(f"lm_heads.{c}.weight" in sd for c in range(s.channels))
        have_emb = all(f"model.embedding_list.{c}.weight" in sd for c in range(s.channels))
        if not (have_heads or have_emb):
            raise KeyError("state dict has neither lm_heads.* nor model.embedding_list.*")
        if tie_word_embeddings is None:
            tie_word_embeddings = not (have_heads and have_emb) or all(
                sd[f"lm_heads.{c}.weight"].data_ptr() == sd[f"model.embedding_list.{c}.weight"].data_ptr()
                for c in range(s.channels))
        self.tied = bool(tie_word_embeddings)
        for c in range(s.channels):
            hk = f"lm_heads.{c}.weight" if have_heads else f"model.embedding_list.{c}.weight"
            self.head_view(c).copy_(get(hk))
        if not self.tied:
            self.embeds = torch.zeros_like(self.heads)
            for c in range(s.channels):
                self.embed_view(c).copy_(get(f"model.embedding_list.{c}.weight"))
        else:
            self.embeds = None
        pre = "model.language_model."
        for l, L in enumerate(self.layers):
            b = f"{pre}layers.{l}."
            L["ln1"].copy_(get(b + "input_layernorm.weight"))
            L["ln2"].copy_(get(b + "post_attention_layernorm.weight"))
            L["q_norm"].copy_(get(b + "self_attn.q_norm.weight"))
            L["k_norm"].copy_(get(b + "self_attn.k_norm.weight"))
            nq = s.num_attention_heads * s.head_dim
            nk = s.num_key_value_heads * s.head_dim
            L["wqkv"][:nq].copy_(get(b + "self_attn.q_proj.weight"))
            L["wqkv"][nq:nq + nk].copy_(get(b + "self_attn.k_proj.weight"))
            L["wqkv"][nq + nk:].copy_(get(b + "self_attn.v_proj.weight"))
            L["wo"].copy_(get(b + "self_attn.o_proj.weight"))
            L["wgu"][0::2].copy_(get(b + "mlp.gate_proj.weight"))
            L["wgu"][1::2].copy_(get(b + "mlp.up_proj.weight"))
            L["wd"].copy_(get(b + "mlp.down_proj.weight"))
        self.final_norm.copy_(get(pre + "norm.weight"))
        return self

    def init_random_(self, seed: int = 0, std: float = 0.02, tied: bool = True, speech_only_head0=None):
        """Seeded random init directly on the device (bench / smoke only; parity tests load reference weights).
        `speech_only_head0=(lo, hi)` zeroes the non-speech rows of head 0 so that greedy decoding of a random-init
        model stays inside the speech range and never emits EOS."""
        g = torch.Generator(device=self.device).manual_seed(seed)
        self.heads.normal_(0.0, std, generator=g)
        self.tied = tied
        if not tied:
            self.embeds = torch.empty_like(self.heads).normal_(0.0, std, generator=g)
        else:
            self.embeds = None
        for c in range(self.shape.channels):  # padding rows between heads stay zero
            o, v = self.shape.head_offsets[c], self.shape.vocabs[c]
            self.heads[o + v:o + _pad_rows(v)].zero_()
        if speech_only_head0 is not None:
            lo, hi = speech_only_head0
            self.heads[:lo].zero_()
            self.heads[hi:self.shape.vocabs[0]].zero_()
        for L in self.layers:
            for k in ("wqkv", "wo", "wgu", "wd"):
                L[k].normal_(0.0, std, generator=g)
        return self


class KVCache:
    """bf16 K/V pools for all layers plus the page mapping of one batch."""

    def __init__(self, shape: LMShape, batch: int, max_tokens: int, device, paged: bool = False, page_size: int = 64,
                 shuffle_pages: bool = False):
        s = shape
        self.page_size = page_size
        self.max_pages = (max_tokens + page_size - 1) // page_size
        self.num_pages = batch * self.max_pages
        self.paged = paged
        shp = (s.num_hidden_layers, self.num_pages, s.num_key_value_heads, page_size, s.head_dim)
        self.k = torch.empty(shp, dtype=torch.bfloat16, device=device)
        self.v = torch.empty(shp, dtype=torch.bfloat16, device=device)
        self.block_table = None
        if paged:
            # every sequence gets its pages up front from one pool; `shuffle_pages` scatters them to exercise
            # the indirection (an on-demand allocator only changes which integers are written here)
            ids = np.arange(self.num_pages, dtype=np.int32)
            if shuffle_pages:
                np.random.default_rng(0).shuffle(ids)
            self.block_table = torch.from_numpy(ids.reshape(batch, self.max_pages)).to(device)

    def nbytes(self):
        return self.k.numel() * 4


class MttsCache:
    """`past_key_values` of the drop-in model: the bf16 K/V pools of one batch plus, per sequence, how many (real, i.e.
    unpadded) tokens they hold. Returned by `forward(use_cache=True)`, accepted by `forward(past_key_values=...)`."""

    def __init__(self, cache: KVCache, lengths, capacity: int):
        self.cache = cache
        self.lengths = np.asarray(lengths, dtype=np.int64).copy()
        self.capacity = capacity

    def get_seq_length(self, layer_idx: int = 0) -> int:
        return int(self.lengths.max()) if self.lengths.size else 0


class SamplerSetup:
    """Host-side description of the per-channel processors -> `mtts_sampler_config`."""

    def __init__(self, shape: LMShape, do_samples, layers, pad_token=1024, eos_mask_token=152694):
        cfg = _lib.SamplerConfig()
        cfg.channels = shape.channels
        off_words = 0
        for c in range(shape.channels):
            cfg.vocab[c] = shape.vocabs[c]
            cfg.logit_offset[c] = shape.head_offsets[c]
            cfg.do_sample[c] = 1 if do_samples[c] else 0
            lc = layers[c] if layers is not None and c < len(layers) else {}
            rp, tp, tk, pp = lc.get("repetition_penalty"), lc.get("temperature"), lc.get("top_k"), lc.get("top_p")
            cfg.has_rep[c] = 0 if rp is None else 1
            cfg.rep_penalty[c] = 1.0 if rp is None else float(rp)
            cfg.has_temp[c] = 0 if tp is None else 1
            cfg.temperature[c] = 1.0 if tp is None else float(tp)
            cfg.top_k[c] = 0 if tk is None else int(tk)
            cfg.has_top_p[c] = 0 if pp is None else 1
            cfg.top_p[c] = 1.0 if pp is None else float(pp)
            cfg.seen_offset_words[c] = off_words
            off_words += (shape.vocabs[c] + 31) // 32
        cfg.seen_words_per_row = off_words
        cfg.pad_token = pad_token
        cfg.eos_mask_token = eos_mask_token
        self.cfg = cfg
        self.words_per_row = off_words


class DecoderEngine:
    """Runs prefill and decode steps for one batch on one GPU."""

    def __init__(self, weights: LMWeights):
        self.w = weights
        self.s = weights.shape
        self.dev = weights.device
        self.L = _lib.load()
        check(self.L.mtts_init())
        s = self.s
        self._tables = (ctypes.c_void_p * 8)(*[weights.embed_view(c).data_ptr() for c in range(s.channels)],
                                             *([None] * (8 - s.channels)))
        self._vocabs = (ctypes.c_int * 8)(*s.vocabs, *([0] * (8 - s.channels)))
        self.err = torch.zeros(4, dtype=torch.int32, device=self.dev)
        self.use_graph = os.environ.get("MTTS_NO_GRAPH", "0") != "1"
        # batch <= mega_max_b: the whole step (layer stack + LM heads) is ONE persistent kernel. Measured on B200
        # (ms/step, ctx 460): batch 1: 0.97 vs 1.44 for the kernel chain, batch 2: 1.06 vs 1.46, batch 3: 1.16 vs 1.33,
        # batch 4: 1.22 vs 1.36.
        self.use_mega = os.environ.get("MTTS_NO_MEGA", "0") != "1"
        # decode steps of the kernel chain: q/k norm + RoPE + cache append run inside the attention kernel
        self.fused_decode_attn = (os.environ.get("MTTS_ATTN_FUSED", "1") != "0" and os.environ.get("MTTS_ATTN_SIMT", "0") != "1"
                                  and s.head_dim == 128 and s.num_attention_heads // s.num_key_value_heads in (1, 2, 4))
        self.mega_max_b = int(os.environ.get("MTTS_MEGA_MAX_B", "4"))
        # 128: tcgen05 flash attention (mtts_gqa_prefill_tc), 64: mma.sync flash kernel, 4: CUDA cores
        self.prefill_tile_rows = int(os.environ.get("MTTS_PREFILL_TILE", "128"))
        if s.head_dim != 128 and self.prefill_tile_rows == 128:
            self.prefill_tile_rows = 64
        # decode steps of the kernel chain with splitk_min_rows <= batch <= 256: q/k/v, o_proj and down_proj store fp32
        # split-K partial tiles and the consumer (bf16 cast / residual + RMSNorm) sums them (mtts_gemm_splitk*)
        # (measured ms per decode step, cluster split-K with fused epilogues -> this path: batch 64 2.32 -> 2.21,
        # batch 128 3.42 -> 3.20, batch 256 4.91 -> 4.49; batch 16 is slower this way, 1.46 -> 1.57)
        self.splitk_min_rows = int(os.environ.get("MTTS_SPLITK_MIN_ROWS", "64"))
        # decode steps: LM heads fused with the sampler (mtts_heads8_sample); False keeps the [B, vpad] logits in
        # st["logits"] (tests that inspect them)
        self.fuse_heads = os.environ.get("MTTS_FUSE_HEADS", "1") != "0"
        self.use_splitk = os.environ.get("MTTS_SPLITK", "1") != "0"
        self.graph_replayed_launches = 0  # kernels executed through graph replays (not seen by mtts_launch_count)

    # ------------------------------------------------------------------ primitive launches
    def _embed(self, ids, out):
        rows = ids.shape[0]
        check(self.L.mtts_embed_sum8(ptr(ids), rows, self.s.channels, self._tables, self._vocabs, self.s.hidden_size,
                                     ptr(out), ptr(self.err), stream_ptr()))

    def _rmsnorm(self, x, w, out):
        check(self.L.mtts_rmsnorm(ptr(x), x.stride(0), ptr(w), ptr(out), out.stride(0), x.shape[0], x.shape[1],
                                  self.s.rms_norm_eps, stream_ptr()))

    def _rope_kv(self, qkv, lw, positions, row_seq, q_out, cache: KVCache, layer: int):
        s = self.s
        check(self.L.mtts_qknorm_rope_kvappend(
            ptr(qkv), qkv.stride(0), ptr(lw["q_norm"]), ptr(lw["k_norm"]), ptr(self.w.inv_freq), ptr(positions),
            ptr(row_seq), ptr(q_out), ptr(cache.k[layer]), ptr(cache.v[layer]), ptr(cache.block_table), cache.max_pages,
            cache.page_size, cache.num_pages, qkv.shape[0], s.num_attention_heads, s.num_key_value_heads, s.head_dim,
            s.rms_norm_eps, ptr(self.err), stream_ptr()))

    def _attention(self, q, cache: KVCache, layer: int, positions, row_seq, out, tiles, rows_per_tile, tile_row0,
                   tile_nrows, nsplit, ws):
        s = self.s
        if rows_per_tile == 128:
            check(self.L.mtts_gqa_prefill_tc(
                ptr(q), q.shape[0], ptr(cache.k[layer]), ptr(cache.v[layer]), ptr(cache.block_table), cache.max_pages,
                cache.page_size, cache.num_pages, ptr(tile_row0), ptr(tile_nrows), ptr(row_seq), ptr(positions), ptr(out), tiles,
                s.num_attention_heads, s.num_key_value_heads, s.head_dim, stream_ptr()))
            return
        check(self.L.mtts_gqa_attention(
            ptr(q), ptr(cache.k[layer]), ptr(cache.v[layer]), ptr(cache.block_table), cache.max_pages, cache.page_size,
            ptr(tile_row0), ptr(tile_nrows), ptr(row_seq), ptr(positions), ptr(out), tiles, rows_per_tile,
            s.num_attention_heads, s.num_key_value_heads, s.head_dim, nsplit, ptr(ws), ws.numel() if ws is not None else 0,
            stream_ptr()))

    def _attn_workspace(self, tiles, rows_per_tile, nsplit):
        s = self.s
        n = self.L.mtts_gqa_attention_workspace_bytes(tiles, s.num_key_value_heads,
                                                      s.num_attention_heads // s.num_key_value_heads, rows_per_tile,
                                                      nsplit)
        return torch.zeros(n, dtype=torch.uint8, device=self.dev)

    # ------------------------------------------------------------------ layer stack over R packed rows
    def _alloc_acts(self, R):
        s = self.s
        bf = dict(dtype=torch.bfloat16, device=self.dev)
        nqkv = (s.num_attention_heads + 2 * s.num_key_value_heads) * s.head_dim
        return dict(x=torch.empty((R, s.hidden_size), **bf), xn=torch.empty((R, s.hidden_size), **bf),
                    qkv=torch.empty((R, nqkv), **bf), q=torch.empty((R, s.num_attention_heads * s.head_dim), **bf),
                    ao=torch.empty((R, s.num_attention_heads * s.head_dim), **bf),
                    h=torch.empty((R, s.intermediate_size), **bf))

    def _gemm_ws(self, R):
        s = self.s
        need = 0
        nqkv = (s.num_attention_heads + 2 * s.num_key_value_heads) * s.head_dim
        for (n, k) in ((nqkv, s.hidden_size), (s.hidden_size, s.num_attention_heads * s.head_dim),
                       (2 * s.intermediate_size, s.hidden_size), (s.hidden_size, s.intermediate_size),
                       (s.vpad, s.hidden_size)):
            need = max(need, self.L.mtts_gemm_workspace_bytes(R, n, k, ops.BF16))
        return torch.zeros(need, dtype=torch.uint8, device=self.dev)

    def _splitk(self, x, w, ws):
        splits = ctypes.c_int(0)
        check(self.L.mtts_gemm_splitk(ptr(x), x.stride(0), ptr(w), w.stride(0), ptr(ws), ws.numel() * 4, x.shape[0], w.shape[0],
                                      x.shape[1], ctypes.byref(splits), stream_ptr()))
        return splits.value

    def _splitk_ws(self, R):
        s = self.s
        nqkv = (s.num_attention_heads + 2 * s.num_key_value_heads) * s.head_dim
        need = max(self.L.mtts_gemm_splitk_workspace_bytes(R, n, k) for (n, k) in
                   ((nqkv, s.hidden_size), (s.hidden_size, s.num_attention_heads * s.head_dim), (s.hidden_size, s.intermediate_size)))
        return torch.empty(need // 4 + 64, dtype=torch.float32, device=self.dev)

    def _layers_splitk(self, a, cache, positions, attn_kw, gws, pws):
        """Decode step (one row per sequence) with consumer-side split-K reduction: per layer
        q/k/v partials -> bf16 cast -> fused norm/RoPE/append/attention -> o_proj partials -> (+residual, RMSNorm) ->
        gate/up + SwiGLU -> down_proj partials -> (+residual, next RMSNorm)."""
        x, xn, qkv, ao, h = a["x"], a["xn"], a["qkv"], a["ao"], a["h"]
        s, L = self.s, self.L
        R, H = x.shape
        eps = s.rms_norm_eps
        ws = attn_kw["ws"]
        layers = self.w.layers
        self._rmsnorm(x, layers[0]["ln1"], xn)
        for l, lw in enumerate(layers):
            S = self._splitk(xn, lw["wqkv"], pws)
            check(L.mtts_gqa_decode_fused_splitk(
                ptr(pws), S, ptr(lw["q_norm"]), ptr(lw["k_norm"]), ptr(self.w.inv_freq), eps,
                ptr(cache.k[l]), ptr(cache.v[l]), ptr(cache.block_table), cache.max_pages, cache.page_size,
                cache.num_pages, ptr(positions), ptr(ao), R, s.num_attention_heads, s.num_key_value_heads, s.head_dim,
                attn_kw["nsplit"], ptr(ws), ws.numel() if ws is not None else 0, ptr(self.err), stream_ptr()))
            S = self._splitk(ao, lw["wo"], pws)
            check(L.mtts_splitk_reduce_rmsnorm(ptr(pws), S, R, H, ptr(x), x.stride(0), ptr(lw["ln2"]), ptr(xn), xn.stride(0), eps,
                                               stream_ptr()))
            ops.gemm(xn, lw["wgu"], out=h, swiglu=True, workspace=gws)
            S = self._splitk(h, lw["wd"], pws)
            nw = layers[l + 1]["ln1"] if l + 1 < len(layers) else self.w.final_norm
            check(L.mtts_splitk_reduce_rmsnorm(ptr(pws), S, R, H, ptr(x), x.stride(0), ptr(nw), ptr(xn), xn.stride(0), eps,
                                               stream_ptr()))
        return xn

    def _layers(self, a, cache, positions, row_seq, attn_kw, gws, want_output=True):
        x, xn, qkv, q, ao, h = a["x"], a["xn"], a["qkv"], a["q"], a["ao"], a["h"]
        last = len(self.w.layers) - 1
        for l, lw in enumerate(self.w.layers):
            self._rmsnorm(x, lw["ln1"], xn)
            ops.gemm(xn, lw["wqkv"], out=qkv, workspace=gws)
            if l == last and not want_output:          # only this layer's K/V are still needed
                self._rope_kv(qkv, lw, positions, row_seq, q, cache, l)
                return None
            if self.fused_decode_attn and attn_kw["rows_per_tile"] == 1 and row_seq is None:
                ws = attn_kw["ws"]
                check(self.L.mtts_gqa_decode_fused(
                    ptr(qkv), qkv.stride(0), ptr(lw["q_norm"]), ptr(lw["k_norm"]), ptr(self.w.inv_freq), self.s.rms_norm_eps,
                    ptr(cache.k[l]), ptr(cache.v[l]), ptr(cache.block_table), cache.max_pages, cache.page_size,
                    cache.num_pages, ptr(positions), ptr(ao), qkv.shape[0], self.s.num_attention_heads,
                    self.s.num_key_value_heads, self.s.head_dim, attn_kw["nsplit"], ptr(ws),
                    ws.numel() if ws is not None else 0, ptr(self.err), stream_ptr()))
            else:
                self._rope_kv(qkv, lw, positions, row_seq, q, cache, l)
                self._attention(q, cache, l, positions, row_seq, ao, **attn_kw)
            ops.gemm(ao, lw["wo"], out=x, residual=x, workspace=gws)
            self._rmsnorm(x, lw["ln2"], xn)
            ops.gemm(xn, lw["wgu"], out=h, swiglu=True, workspace=gws)
            ops.gemm(h, lw["wd"], out=x, residual=x, workspace=gws)
        self._rmsnorm(x, self.w.final_norm, xn)
        return xn

    # ------------------------------------------------------------------ prefill
    def prefill(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, cache: KVCache, all_logits: bool = False):
        """input_ids (B, P, C) int64, attention_mask (B, P). Real tokens (mask != 0) are packed; their positions
        are cumsum(mask) - 1. Returns (logits bf16 [B, Vpad] of each sequence's last real token, lengths [B]) or,
        with all_logits, the logits of every packed row plus the packing index."""
        B, P, C = input_ids.shape
        mask = (attention_mask != 0)
        lens = mask.sum(1).to(torch.int64)
        lens_h = lens.cpu().numpy()
        flat_idx = mask.reshape(-1).nonzero(as_tuple=False).squeeze(1)          # packing is pure index plumbing
        ids = input_ids.reshape(B * P, C).index_select(0, flat_idx).contiguous()
        out = self.prefill_packed(ids, lens_h, np.arange(B, dtype=np.int32), cache, "all" if all_logits else "last")
        if all_logits:
            return out, flat_idx, lens
        return out, lens

    def prefill_packed(self, ids: Optional[torch.Tensor], lens_h, slots_h, cache: KVCache, logits: Optional[str] = "last",
                       pos0_h=None, embeds: Optional[torch.Tensor] = None):
        """Packed prefill: `ids` (R, C) int64 holds the rows of len(lens_h) sequences back to back; sequence i has
        lens_h[i] rows at positions 0 .. lens_h[i]-1 and its K/V go to cache slot slots_h[i] (a row of the page table:
        with continuous batching the slot is wherever a finished request made room). logits: "last" -> [n, Vpad] of each
        sequence's last row, "all" -> [R, Vpad], None -> only the K/V side effect (admission of a queued request).
        pos0_h: first position of each sequence's rows (continuing a cache that already holds pos0 tokens, i.e.
        `forward(past_key_values=...)`); embeds: (R, H) bf16 rows used instead of the embedding sum (`inputs_embeds=`)."""
        lens_h = np.asarray(lens_h, dtype=np.int64)
        n = len(lens_h)
        R = int(lens_h.sum())
        cu = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens_h, out=cu[1:])
        p0 = np.zeros(n, dtype=np.int64) if pos0_h is None else np.asarray(pos0_h, dtype=np.int64)
        pos_h = (np.concatenate([np.arange(a, a + k, dtype=np.int32) for a, k in zip(p0, lens_h)]) if R
                 else np.zeros(0, np.int32)).astype(np.int32)
        seq_h = np.repeat(np.asarray(slots_h, dtype=np.int32), lens_h)
        row0_h, nrows_h = [], []
        tile_rows = self.prefill_tile_rows
        if tile_rows == 128 and cache.page_size < 64:   # the TMA key tiles of the tcgen05 kernel are 64 keys of ONE page
            tile_rows = 64
        for b in range(n):
            for t in range(0, int(lens_h[b]), tile_rows):
                row0_h.append(cu[b] + t)
                nrows_h.append(min(tile_rows, int(lens_h[b]) - t))
        positions = torch.from_numpy(pos_h).to(self.dev)
        row_seq = torch.from_numpy(seq_h).to(self.dev)
        tile_row0 = torch.tensor(row0_h, dtype=torch.int32, device=self.dev)
        tile_nrows = torch.tensor(nrows_h, dtype=torch.int32, device=self.dev)
        a = self._alloc_acts(R)
        gws = self._gemm_ws(R)
        if embeds is not None:
            a["x"].copy_(embeds.to(self.dev, torch.bfloat16))
        else:
            self._embed(ids, a["x"])
        attn_kw = dict(tiles=len(row0_h), rows_per_tile=tile_rows, tile_row0=tile_row0, tile_nrows=tile_nrows, nsplit=1, ws=None)
        xn = self._layers(a, cache, positions, row_seq, attn_kw, gws, want_output=logits is not None)
        if logits is None:
            return None
        if logits == "all":
            out = torch.empty((R, self.s.vpad), dtype=torch.bfloat16, device=self.dev)
            ops.gemm(xn, self.w.heads, out=out, workspace=gws)
            return out
        last = torch.from_numpy(cu[1:] - 1).to(self.dev)
        hl = xn.index_select(0, last).contiguous()
        out = torch.empty((n, self.s.vpad), dtype=torch.bfloat16, device=self.dev)
        ops.gemm(hl, self.w.heads, out=out, workspace=self._gemm_ws(n))
        return out

    # ------------------------------------------------------------------ decode
    def make_decode_state(self, B: int, cache: KVCache, sampler: SamplerSetup, max_len_rows: int, speech_range,
                          eos_token: int, has_eos_criteria: bool):
        """Buffers (and, after the first step, the captured CUDA graph) of one decode session. Everything that
        changes between generate() calls — seed, prompt rows, max_length, the prompt tail, the per-row state — lives
        in device memory and is refreshed by `reset_decode_state`, so the session (incl. its graph) is reusable."""
        st = dict(B=B, cache=cache, sampler=sampler, max_len_rows=max_len_rows, speech=speech_range, eos=eos_token,
                  has_eos=has_eos_criteria)
        i32 = dict(dtype=torch.int32, device=self.dev)
        C = self.s.channels
        st["tokens"] = torch.zeros((B, C), dtype=torch.int64, device=self.dev)
        st["tf_tail"] = torch.zeros((B, C - 1, C), dtype=torch.int64, device=self.dev)
        st["positions"] = torch.zeros(B, **i32)
        st["unfinished"] = torch.ones(B, **i32)
        st["needs"] = torch.full((B,), -1, **i32)
        st["finish_len"] = torch.zeros(B, **i32)
        st["step"] = torch.zeros(1, **i32)
        st["dyn"] = torch.zeros(4, **i32)                       # [prompt rows P, max_length, -, -]
        st["seed_dev"] = torch.zeros(1, dtype=torch.int64, device=self.dev)
        st["hist"] = torch.full((max(max_len_rows, 8) + 16,), -1, **i32)
        st["seen"] = torch.zeros((B, sampler.words_per_row), dtype=torch.int32, device=self.dev)
        st["sequences"] = torch.zeros((B, max_len_rows, C), dtype=torch.int64, device=self.dev)
        st["logits"] = torch.empty((B, self.s.vpad), dtype=torch.bfloat16, device=self.dev)
        st["acts"] = self._alloc_acts(B)
        st["gws"] = self._gemm_ws(B)
        st["pws"] = self._splitk_ws(B) if (self.use_splitk and self.fused_decode_attn and self.splitk_min_rows <= B <= 256) else None
        s = self.s
        # split-KV only while (rows x kv heads) cannot fill the machine by itself: measured at B=64 the combine costs
        # more than it gains (29 us unsplit vs 36 us with 2 splits), at B=1 eight splits are 2.3x faster than none
        nsplit = max(1, min(8, -(-(2 * 148) // max(1, B * s.num_key_value_heads))))
        if os.environ.get("MTTS_ATTN_NSPLIT"):
            nsplit = max(1, min(8, int(os.environ["MTTS_ATTN_NSPLIT"])))
        st["nsplit"] = nsplit
        st["attn_ws"] = self._attn_workspace(B, 1, nsplit)
        st["sample_ws"] = torch.zeros(self.L.mtts_heads8_sample_workspace_bytes(B, self.s.vpad, self.s.channels),
                                      dtype=torch.uint8, device=self.dev)
        st["mega"] = self._make_mega(st) if self.mega_supported(B) else None
        st["graph"] = None
        return st

    # ------------------------------------------------------------------ small-batch persistent decode kernel
    def mega_supported(self, B: int) -> bool:
        s = self.s
        return self.use_mega and B <= self.mega_max_b and bool(self.L.mtts_decode_mega_supported(
            s.hidden_size, s.intermediate_size, s.num_attention_heads, s.num_key_value_heads, s.head_dim, B))

    def _make_mega(self, st):
        """Argument block of mtts_decode_mega for this session: a device table of per-layer pointers (weights + this
        cache's K/V pools), the barrier words and the split-KV workspace."""
        s, cache, B = self.s, st["cache"], st["B"]
        tab = np.zeros((s.num_hidden_layers, 10), dtype=np.int64)
        for l, lw in enumerate(self.w.layers):
            tab[l] = [lw[k].data_ptr() for k in ("wqkv", "wo", "wgu", "wd", "ln1", "ln2", "q_norm", "k_norm")] + \
                     [cache.k[l].data_ptr(), cache.v[l].data_ptr()]
        layers = torch.from_numpy(tab).to(self.dev)
        # one attention unit (row, kv head, key range) per CTA. Measured (round 2, after the stack-frame work): 4 key ranges
        # are best at every batch size at ctx 460 (batch 4: 1.058 ms against 1.071 with 3); from 700 rows of context on,
        # batch 1 wants 8 (ctx 880: 1.000 ms with 4, 0.975 with 6, 0.965 with 8) and batch 2 wants 6 (1.038 -> 1.010):
        # every 32-key pass of a unit is a serial ~0.7 us step on the critical path of all 148 CTAs, while more ranges
        # cost more partial reads in every CTA's o_proj phase. `ctx_hint` = prompt + new rows of the call that opens the
        # session.
        nsplit = 4
        if getattr(self, "ctx_hint", 0) >= 700:
            nsplit = 8 if B == 1 else (6 if B == 2 else 4)
        if os.environ.get("MTTS_MEGA_NSPLIT"):
            nsplit = int(os.environ["MTTS_MEGA_NSPLIT"])
        ws = torch.zeros(self.L.mtts_decode_mega_workspace_bytes(B, nsplit) + 256, dtype=torch.uint8, device=self.dev)
        off = (-ws.data_ptr()) % 256
        a = st["acts"]
        args = _lib.DecodeMegaArgs(
            layers=layers.data_ptr(), num_layers=s.num_hidden_layers, hidden=s.hidden_size,
            intermediate=s.intermediate_size, num_q_heads=s.num_attention_heads, num_kv_heads=s.num_key_value_heads,
            head_dim=s.head_dim, heads=self.w.heads.data_ptr(), vpad=s.vpad, final_norm=self.w.final_norm.data_ptr(),
            inv_freq=self.w.inv_freq.data_ptr(), positions=st["positions"].data_ptr(),
            block_table=ptr(cache.block_table), max_pages=cache.max_pages, page_size=cache.page_size,
            num_pages=cache.num_pages, x=a["x"].data_ptr(),
            logits=st["logits"].data_ptr(), ld_logits=st["logits"].stride(0), B=B, nsplit=nsplit, eps=s.rms_norm_eps,
            workspace=ws.data_ptr() + off, workspace_bytes=ws.numel() - 256, err_flag=self.err.data_ptr())
        prof = None
        if os.environ.get("MTTS_MEGA_PROFILE", "0") == "1":   # per-phase cycle counters of CTA 0 (scripts/profile_mega.py)
            prof = torch.zeros(32 + 16 * 160, dtype=torch.int64, device=self.dev)
            args.profile_cycles = prof.data_ptr()
        return dict(args=args, keep=(layers, ws), prof=prof, nsplit=nsplit)

    def reset_decode_state(self, st, seed: int, prompt_rows: int, max_length: int):
        st["unfinished"].fill_(1)
        st["needs"].fill_(-1)
        st["finish_len"].zero_()
        st["step"].zero_()
        st["hist"].fill_(-1)
        st["seen"].zero_()
        st["dyn"].copy_(torch.tensor([prompt_rows, max_length, 0, 0], dtype=torch.int32))
        st["seed_dev"].copy_(torch.tensor([seed & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64))
        st["P"] = prompt_rows
        st["max_length"] = max_length

    def sample_and_advance(self, st, logits):
        """Draw 8 tokens per row from `logits` and run the delay-pattern state machine (one step)."""
        sm = st["sampler"]
        rc = st.get("row_ctl")   # per-row step origin / prompt length / limits (continuous batching), else None
        check(self.L.mtts_sample8_rows(ptr(logits), logits.stride(0), st["B"], ctypes.byref(sm.cfg), ptr(st["seen"]),
                                       ptr(st["step"]), ptr(rc), ptr(st["seed_dev"]), ptr(st["tokens"]), ptr(self.err),
                                       ptr(st["sample_ws"]), st["sample_ws"].numel(), stream_ptr()))
        check(self.L.mtts_delay_step_rows(ptr(st["tokens"]), ptr(st["tf_tail"]), ptr(st["sequences"]), st["max_len_rows"],
                                          ptr(st["unfinished"]), ptr(st["needs"]), ptr(st["positions"]), ptr(st["seen"]),
                                          ptr(st["step"]), ptr(st["hist"]), st.get("hist_len", 0), ptr(st["finish_len"]),
                                          st["B"], ptr(st["dyn"]), ptr(rc), st["speech"][0], st["speech"][1], st["eos"],
                                          1 if st["has_eos"] else 0, ctypes.byref(sm.cfg), stream_ptr()))

    def heads_sample_and_advance(self, st, xn):
        """LM heads + sampler + delay-pattern state machine from the final-norm output `xn` [B, H]. Greedy rows without a
        repetition penalty at batch > 64 never materialise the logits (mtts_heads8_sample)."""
        if not self.fuse_heads or st.get("keep_logits"):
            ops.gemm(xn, self.w.heads, out=st["logits"], workspace=st["gws"])
            return self.sample_and_advance(st, st["logits"])
        sm = st["sampler"]
        rc = st.get("row_ctl")
        s = self.s
        check(self.L.mtts_heads8_sample(ptr(xn), xn.stride(0), ptr(self.w.heads), self.w.heads.stride(0), st["B"], s.hidden_size,
                                        s.vpad, ctypes.byref(sm.cfg), ptr(st["seen"]), ptr(st["step"]), ptr(rc),
                                        ptr(st["seed_dev"]), ptr(st["logits"]), st["logits"].stride(0), ptr(st["tokens"]),
                                        ptr(self.err), ptr(st["sample_ws"]), st["sample_ws"].numel(), stream_ptr()))
        check(self.L.mtts_delay_step_rows(ptr(st["tokens"]), ptr(st["tf_tail"]), ptr(st["sequences"]), st["max_len_rows"],
                                          ptr(st["unfinished"]), ptr(st["needs"]), ptr(st["positions"]), ptr(st["seen"]),
                                          ptr(st["step"]), ptr(st["hist"]), st.get("hist_len", 0), ptr(st["finish_len"]),
                                          st["B"], ptr(st["dyn"]), ptr(rc), st["speech"][0], st["speech"][1], st["eos"],
                                          1 if st["has_eos"] else 0, ctypes.byref(sm.cfg), stream_ptr()))

    def _decode_body(self, st):
        a = st["acts"]
        self._embed(st["tokens"], a["x"])
        if st.get("mega") is not None:
            check(self.L.mtts_decode_mega(ctypes.byref(st["mega"]["args"]), stream_ptr()))
            self.sample_and_advance(st, st["logits"])
            return
        attn_kw = dict(tiles=st["B"], rows_per_tile=1, tile_row0=None, tile_nrows=None, nsplit=st["nsplit"],
                       ws=st["attn_ws"] if st["nsplit"] > 1 else None)
        if st.get("pws") is not None:
            xn = self._layers_splitk(a, st["cache"], st["positions"], attn_kw, st["gws"], st["pws"])
        else:
            xn = self._layers(a, st["cache"], st["positions"], None, attn_kw, st["gws"])
        self.heads_sample_and_advance(st, xn)

    def decode_step(self, st):
        """Feed the row appended by the previous step, sample the next one. Replays a CUDA graph after the
        first call (launch-bound otherwise: ~230 kernels per step)."""
        if not self.use_graph:
            self._decode_body(st)
            return
        if st["graph"] is None:
            # warm-up launch outside capture would advance the state, so capture directly; all buffers are
            # preallocated and every kernel argument that changes per step lives in device memory.
            g = torch.cuda.CUDAGraph()
            n0 = self.L.mtts_launch_count()
            cap_stream = torch.cuda.Stream(device=self.dev)
            cap_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cap_stream):
                with torch.cuda.graph(g, stream=cap_stream):
                    self._decode_body(st)
            torch.cuda.current_stream().wait_stream(cap_stream)
            st["graph"] = g
            st["graph_nodes"] = self.L.mtts_launch_count() - n0
        st["graph"].replay()
        self.graph_replayed_launches += st["graph_nodes"]
