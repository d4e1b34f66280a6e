"""Drop-in mirror of the reference's `modeling_asteroid.py` for the inference path.

Same names, argument meaning and error behaviour as /root/reference/modeling_asteroid.py:
  AsteroidTTSConfig (:17-28), AsteroidTTSOutputWithPast (:31-39), GenerateDecoderOnlyOutput (:42-49),
  AsteroidTTSInstruct (:288-426) with .generate (HF GenerationMixin.generate -> CustomMixin._sample :52-197),
  .forward (:337-426, inference branch), .is_speech_token (:311-312), .set_weights (:334-335).
What differs is what runs underneath: libmtts CUDA kernels driven by `lm_engine.DecoderEngine`, one CUDA graph
replay per generated frame, no host sync inside the step. Training (`labels=`) is out of scope and raises.
"""
from __future__ import annotations

import copy
import json
import os
from dataclasses import dataclass
from typing import Optional, Tuple, Union

import torch

from .lm_engine import DecoderEngine, KVCache, LMShape, LMWeights, MttsCache, SamplerSetup


class AsteroidTTSConfig:
    """Qwen3-style decoder config plus the four TTS fields of the reference (modeling_asteroid.py:17-28)."""
    model_type = "asteroid_tts"

    def __init__(self, channels=8, speech_pad_token=1024, speech_vocab_size=1025, speech_token_range=None, **kwargs):
        defaults = dict(vocab_size=151936, hidden_size=4096, intermediate_size=22016, num_hidden_layers=32,
                        num_attention_heads=32, num_key_value_heads=32, head_dim=128, hidden_act="silu",
                        max_position_embeddings=32768, rms_norm_eps=1e-6, rope_theta=10000.0, attention_bias=False,
                        tie_word_embeddings=False, pad_token_id=None, bos_token_id=None, eos_token_id=None,
                        output_attentions=False, output_hidden_states=False, use_return_dict=True)
        defaults.update(kwargs)
        for k, v in defaults.items():
            setattr(self, k, v)
        self.channels = channels
        self.speech_pad_token = speech_pad_token
        self.speech_vocab_size = speech_vocab_size
        self.speech_token_range = list(speech_token_range) if speech_token_range is not None else []

    @classmethod
    def from_pretrained(cls, path, **kwargs):
        with open(os.path.join(path, "config.json")) as f:
            d = json.load(f)
        d.update(kwargs)
        return cls(**d)

    def to_dict(self):
        return dict(self.__dict__)


class GenerationConfig:
    """The subset of HF GenerationConfig the sampler reads, with HF's defaults, plus the reference's custom
    fields `do_samples` and `layers` (modeling_asteroid.py:95-106)."""

    def __init__(self, **kwargs):
        self.max_length = 20
        self.max_new_tokens = None
        self.do_sample = False
        self.temperature = 1.0
        self.top_k = 50
        self.top_p = 1.0
        self.repetition_penalty = 1.0
        self.eos_token_id = None
        self.pad_token_id = None
        self.return_dict_in_generate = False
        self.output_scores = False
        self.output_logits = False
        self.output_attentions = False
        self.output_hidden_states = False
        self.do_samples = None
        self.layers = None
        for k, v in kwargs.items():
            setattr(self, k, v)

    @classmethod
    def from_pretrained(cls, path):
        p = os.path.join(path, "generation_config.json")
        if not os.path.exists(p):
            return cls()
        with open(p) as f:
            return cls(**json.load(f))

    def update(self, **kwargs):
        unused = {}
        for k, v in kwargs.items():
            if hasattr(self, k):
                setattr(self, k, v)
            else:
                unused[k] = v
        return unused


@dataclass
class AsteroidTTSOutputWithPast:
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    loss_all: Optional[Tuple[torch.Tensor]] = None
    logits_all: Optional[Tuple[torch.Tensor]] = None
    past_key_values: Optional[object] = None
    hidden_states: Optional[Tuple[torch.Tensor, ...]] = None
    attentions: Optional[Tuple[torch.Tensor, ...]] = None


@dataclass
class GenerateDecoderOnlyOutput:
    sequences: Optional[torch.Tensor] = None
    scores: Optional[tuple] = None
    logits: Optional[tuple] = None
    attentions: Optional[tuple] = None
    hidden_states: Optional[tuple] = None
    past_key_values: Optional[object] = None


class AsteroidTTSInstruct:
    """Inference-only drop-in for the reference class of the same name."""

    def __init__(self, config: AsteroidTTSConfig, device: Union[str, torch.device, None] = None):
        self.config = config
        self.channels = config.channels
        self.weights = [1 for _ in range(self.channels)]
        self.vocab_size = config.vocab_size
        self.generation_config = GenerationConfig(eos_token_id=config.eos_token_id, pad_token_id=config.pad_token_id)
        self.shape = LMShape.from_config(config)
        self.device = torch.device(device) if device is not None else None
        self._w: Optional[LMWeights] = None
        self._engine: Optional[DecoderEngine] = None
        self.training = False
        self.dtype = torch.bfloat16
        # knobs of the B200 engine (not in the reference)
        self.kv_paged = False
        self.kv_page_size = 64
        self.sync_every = 16
        if self.device is not None and self.device.type == "cuda":
            self._materialize()

    # ------------------------------------------------------------------ construction / loading
    def _materialize(self):
        if self._w is None:
            if self.device is None or self.device.type != "cuda":
                raise RuntimeError("AsteroidTTSInstruct runs on CUDA only (no CPU path); call .to('cuda') first")
            self._w = LMWeights(self.shape, self.device)
            self._engine = None

    @property
    def engine(self) -> DecoderEngine:
        self._materialize()
        if self._engine is None:
            self._engine = DecoderEngine(self._w)
        return self._engine

    def to(self, device=None, dtype=None, **kw):
        if dtype is not None and dtype != torch.bfloat16:
            raise ValueError("the B200 decoder path computes in bf16 (north_star); got dtype %s" % dtype)
        if device is not None:
            device = torch.device(device)
            if self._w is not None and device != self.device:
                raise RuntimeError("moving an already materialised model between devices is not supported")
            self.device = device
            if device.type == "cuda":
                self._pending = getattr(self, "_pending", None)
                self._materialize()
                if self._pending is not None:
                    self._w.load_state_dict(self._pending, self.config.tie_word_embeddings or None)
                    self._pending = None
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", device if device is not None else torch.cuda.current_device()))

    def eval(self):
        self.training = False
        return self

    def load_state_dict(self, sd, strict=True, tie_word_embeddings=None):
        if self._w is None:
            if self.device is not None and self.device.type == "cuda":
                self._materialize()
            else:
                self._pending = sd
                return self
        self._w.load_state_dict(sd, tie_word_embeddings)
        self._engine = None
        return self

    def init_random_weights(self, seed=0, tied=True, speech_only_head0=False):
        self._materialize()
        rng = tuple(self.config.speech_token_range) if speech_only_head0 else None
        self._w.init_random_(seed, tied=tied, speech_only_head0=rng)
        self._engine = None
        return self

    @classmethod
    def from_pretrained(cls, model_path, torch_dtype=torch.bfloat16, attn_implementation=None, device=None, **kwargs):
        """Local directory with config.json (+ generation_config.json) and *.safetensors / pytorch_model*.bin.
        `attn_implementation` is accepted for signature compatibility and ignored: attention always runs in the
        libmtts kernel."""
        if torch_dtype not in (torch.bfloat16, None):
            raise ValueError("the B200 decoder path computes in bf16 only")
        config = AsteroidTTSConfig.from_pretrained(model_path)
        model = cls(config, device=device)
        model.generation_config = GenerationConfig.from_pretrained(model_path)
        if model.generation_config.eos_token_id is None:
            model.generation_config.eos_token_id = config.eos_token_id
        sd = {}
        files = sorted(os.listdir(model_path))
        st_files = [f for f in files if f.endswith(".safetensors")]
        if st_files:
            from safetensors.torch import load_file
            for f in st_files:
                sd.update(load_file(os.path.join(model_path, f)))
        else:
            for f in files:
                if f.endswith(".bin") or f.endswith(".pt"):
                    sd.update(torch.load(os.path.join(model_path, f), map_location="cpu"))
        if not sd:
            raise FileNotFoundError(f"no weights found under {model_path}")
        model.load_state_dict(sd, tie_word_embeddings=config.tie_word_embeddings or None)
        return model

    # ------------------------------------------------------------------ small API of the reference class
    def can_generate(self):
        return True

    def is_speech_token(self, tokens):
        return (tokens >= self.config.speech_token_range[0]) & (tokens < self.config.speech_token_range[1])

    def set_weights(self, weights):
        self.weights = weights

    # ------------------------------------------------------------------ forward (teacher-forced logits)
    @torch.no_grad()
    def forward(self, input_ids: torch.LongTensor = None, attention_mask: Optional[torch.Tensor] = None,
                position_ids=None, past_key_values=None, inputs_embeds=None, labels=None, use_cache=None,
                output_attentions=None, output_hidden_states=None, return_dict=None, cache_position=None,
                skip_logits=None, **kwargs):
        """Teacher-forced logits of every position for (B, S, 8) ids (or (B, S, H) `inputs_embeds`), optionally continuing
        a `past_key_values` cache returned by an earlier call with `use_cache=True` (modeling_asteroid.py:337-426,
        inference branch; `attention_mask` then covers past + new positions, as HF passes it). Positions are
        cumsum(mask) - 1 (`position_ids` / `cache_position` are accepted and must be consistent with that)."""
        if (input_ids is None) ^ (inputs_embeds is not None):
            raise ValueError("You must specify exactly one of input_ids or inputs_embeds")
        if labels is not None:
            raise NotImplementedError("training loss is out of scope of the B200 inference path")
        if output_attentions or output_hidden_states:
            raise NotImplementedError("attention / hidden-state capture is not supported by the fused kernels")
        eng = self.engine
        if input_ids is not None:
            B, S, C = input_ids.shape
            if C != self.config.channels:
                raise ValueError(f"Expected {self.config.channels} channels, got {C}")
            input_ids = input_ids.to(self.device)
        else:
            B, S, Hd = inputs_embeds.shape
            if Hd != self.shape.hidden_size:
                raise ValueError(f"inputs_embeds must have hidden size {self.shape.hidden_size}, got {Hd}")
        past = past_key_values
        if past is not None and not isinstance(past, MttsCache):
            raise TypeError("past_key_values must be the object returned by forward(use_cache=True) of this model")
        if attention_mask is None:
            attention_mask = torch.ones((B, S + (past.get_seq_length() if past is not None else 0)), device=self.device)
        attention_mask = attention_mask.to(self.device)
        if attention_mask.shape[1] < S:
            raise ValueError("attention_mask must cover at least the new positions")
        new_mask = (attention_mask[:, -S:] != 0)
        lens_h = new_mask.sum(1).cpu().numpy().astype("int64")
        if past is None:
            cap = S + (int(getattr(self, "forward_cache_rows", 512)) if use_cache else 0)
            cache = KVCache(self.shape, B, cap, self.device, paged=self.kv_paged, page_size=self.kv_page_size)
            pos0 = lens_h * 0
        else:
            cache, cap, pos0 = past.cache, past.capacity, past.lengths
            if (pos0 + lens_h).max() > cap:
                raise RuntimeError(f"past_key_values holds {int(pos0.max())} rows of a {cap}-row cache; set "
                                   f"model.forward_cache_rows before the first call to reserve more")
        flat_idx = new_mask.reshape(-1).nonzero(as_tuple=False).squeeze(1)
        import numpy as np
        slots = np.arange(B, dtype=np.int32)
        if input_ids is not None:
            ids = input_ids.reshape(B * S, C).index_select(0, flat_idx).contiguous()
            logits = eng.prefill_packed(ids, lens_h, slots, cache, "all", pos0_h=pos0)
        else:
            emb = inputs_embeds.to(self.device).reshape(B * S, -1).index_select(0, flat_idx).contiguous()
            logits = eng.prefill_packed(None, lens_h, slots, cache, "all", pos0_h=pos0, embeds=emb)
        self._check_err()
        full = torch.zeros((B * S, self.shape.vpad), dtype=torch.bfloat16, device=self.device)
        full.index_copy_(0, flat_idx, logits)
        full = full.view(B, S, -1)
        logits_all = tuple(full[..., o:o + v] for o, v in zip(self.shape.head_offsets, self.shape.vocabs))
        pkv = MttsCache(cache, pos0 + lens_h, cap) if use_cache else None
        return AsteroidTTSOutputWithPast(loss=None, logits=logits_all[0], loss_all=None, logits_all=logits_all,
                                         past_key_values=pkv)

    __call__ = forward

    def _check_err(self):
        e = self.engine.err.cpu().tolist()
        if any(e):
            self.engine.err.zero_()
            raise RuntimeError(f"libmtts device-side error flags {e} (1: token id out of range, 2: KV page out of range, "
                               f"3: sampler candidate overflow, 4: unused, "
                               f"5: the persistent decode kernel gave up waiting)")

    # ------------------------------------------------------------------ continuous batching (not in the reference)
    @torch.no_grad()
    def generate_continuous(self, prompts, max_new_tokens=None, max_length=None, max_batch: int = 256, eos_at=None,
                            generation_config: Optional[GenerationConfig] = None, seed: Optional[int] = None,
                            pool_pages: Optional[int] = None, sync_every: int = 8, queue_order: str = "fifo", **kwargs):
        """Decode a QUEUE of scripts through `max_batch` slots, refilling a slot as soon as its row has finished
        (continuous.ContinuousDecoder) instead of padding every batch to its longest row (modeling_asteroid.py:155-169).
        prompts: list of (T_i, 8) int64 delay-shifted grids WITHOUT left padding (last 7 rows = teacher-forced tail);
        max_new_tokens / max_length: scalar or per-prompt list with generate()'s meaning; eos_at: optional per-prompt
        length budgets (sequence row from which channel 0 is forced to EOS). Returns a list of (L_i, 8) LongTensors: for
        every prompt the rows a solo generate() call returns. queue_order: "fifo" admits in arrival order;
        "longest_first" admits the scripts with the largest row budget first (the makespan of a finite job is bounded by
        the longest script that starts late) — results are returned in prompt order either way."""
        from .continuous import ContinuousDecoder, Request
        if queue_order not in ("fifo", "longest_first"):
            raise ValueError(f"queue_order must be 'fifo' or 'longest_first', got {queue_order!r}")
        gc = copy.deepcopy(generation_config if generation_config is not None else self.generation_config)
        gc.update(**kwargs)
        n = len(prompts)
        C = self.config.channels
        per = lambda v, i: (v[i] if isinstance(v, (list, tuple)) else v)
        reqs = []
        for i, g in enumerate(prompts):
            g = torch.as_tensor(g)
            if g.dim() != 2 or g.shape[1] != C:
                raise ValueError(f"Expected (T, {C}) prompt grids, got {tuple(g.shape)}")
            if g.shape[0] < C:
                raise ValueError("prompt grid must include the (channels-1) delay rows")
            mnt = per(max_new_tokens, i) if max_new_tokens is not None else gc.max_new_tokens
            ml = g.shape[0] + mnt if mnt is not None else (per(max_length, i) if max_length is not None else gc.max_length)
            reqs.append(Request(i, g.to(self.device).contiguous(), int(ml), int(per(eos_at, i)) if eos_at is not None else 0))
        if not reqs:
            return []
        if queue_order == "longest_first":
            budget = lambda r: (r.eos_at if r.eos_at > 0 else r.max_length) - r.grid.shape[0]
            reqs.sort(key=lambda r: (-budget(r), r.index))
        eos = gc.eos_token_id
        if isinstance(eos, (list, tuple)):
            eos = eos[0] if len(eos) else None
        eos_fill = self.config.eos_token_id if self.config.eos_token_id is not None else (eos if eos is not None else 0)
        sampler = self._sampler_setup(gc)
        slots = min(max_batch, n)
        max_rows = max(r.max_length for r in reqs) + 2 * C
        key = (slots, bytes(sampler.cfg), pool_pages, sync_every)
        cd = getattr(self, "_continuous", None)
        if cd is None or cd[0] != key or cd[1].max_rows < max_rows:
            self._continuous = None
            cd = (key, ContinuousDecoder(self.engine, slots, max(max_rows, 1024), sampler, tuple(self.config.speech_token_range),
                                         int(eos_fill), pool_pages=pool_pages, page_size=self.kv_page_size,
                                         sync_every=sync_every))
            self._continuous = cd
        if seed is None:
            seed = int(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF)
        res = cd[1].run(reqs, seed=seed)
        return [res[i] for i in range(n)]

    # ------------------------------------------------------------------ generate
    def _sampler_setup(self, gc: GenerationConfig) -> SamplerSetup:
        C = self.channels
        if gc.do_samples is not None:
            do_samples = list(gc.do_samples)
            layers = [dict(l) for l in (gc.layers or [])] + [{} for _ in range(C - len(gc.layers or []))]
        else:
            # shared processor list as HF builds it (modeling_asteroid.py:107-109): repetition penalty always (if != 1),
            # warpers only when sampling
            do_samples = [bool(gc.do_sample)] * C
            lc = {}
            if gc.repetition_penalty is not None and gc.repetition_penalty != 1.0:
                lc["repetition_penalty"] = gc.repetition_penalty
            if gc.do_sample:
                if gc.temperature is not None and gc.temperature != 1.0:
                    lc["temperature"] = gc.temperature
                if gc.top_k is not None and gc.top_k != 0:
                    lc["top_k"] = gc.top_k
                if gc.top_p is not None and gc.top_p < 1.0:
                    lc["top_p"] = gc.top_p
            layers = [dict(lc) for _ in range(C)]
        return SamplerSetup(self.shape, do_samples, layers, pad_token=1024, eos_mask_token=152694)

    @torch.no_grad()
    def generate(self, input_ids: torch.LongTensor = None, attention_mask: Optional[torch.Tensor] = None,
                 generation_config: Optional[GenerationConfig] = None, streamer=None, seed: Optional[int] = None,
                 eos_at=None, **kwargs):
        """input_ids (B, T, 8) int64 — the delay-shifted, left-padded prompt grid whose last 7 rows are the
        teacher-forced tail; attention_mask (B, T). Returns LongTensor (B, T - 7 + G, 8) (or
        GenerateDecoderOnlyOutput with return_dict_in_generate), exactly the rows CustomMixin._sample returns.
        `eos_at` (not in the reference): optional per-row length budgets — sequence row (in this padded grid) from which
        channel 0 of row b is forced to EOS, after which the row winds down as after a sampled EOS; 0 = no budget."""
        gc = copy.deepcopy(generation_config if generation_config is not None else self.generation_config)
        gc.update(**kwargs)
        if gc.output_attentions or gc.output_hidden_states:
            raise NotImplementedError("attention / hidden-state capture is not supported by the fused decode step")
        capture = bool(gc.return_dict_in_generate and (gc.output_scores or gc.output_logits))
        B, T, C = input_ids.shape
        if C != self.config.channels:
            raise ValueError(f"Expected {self.config.channels} channels, got {C}")
        if T < C:
            raise ValueError("prompt grid must include the (channels-1) delay rows")
        dev = self.device
        eng = self.engine
        input_ids = input_ids.to(dev).contiguous()
        if attention_mask is None:
            attention_mask = torch.ones((B, T), device=dev)
        attention_mask = attention_mask.to(dev)
        P = T - (C - 1)
        max_length = T + gc.max_new_tokens if gc.max_new_tokens is not None else gc.max_length
        eos = gc.eos_token_id
        if isinstance(eos, (list, tuple)):
            eos = eos[0] if len(eos) else None
        has_eos = eos is not None
        eos_fill = self.config.eos_token_id if self.config.eos_token_id is not None else (eos if eos is not None else 0)
        if isinstance(eos_fill, (list, tuple)):
            eos_fill = eos_fill[0]
        # A row in wind-down ignores max_length for up to C-2 extra rows; allocate for it.
        max_rows = max(max_length, P + 1) + C + 2 * self.sync_every
        sampler = self._sampler_setup(gc)
        if capture and gc.output_scores and any(sampler.cfg.has_rep[c] or sampler.cfg.has_temp[c] or sampler.cfg.top_k[c] or
                                                sampler.cfg.has_top_p[c] for c in range(C)):
            raise NotImplementedError("output_scores with logits processors: the processed scores never leave the sampler "
                                      "kernel; use output_logits (masked raw logits, modeling_asteroid.py:123-128,176)")
        if seed is None:
            seed = int(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF)  # follows torch.manual_seed / accelerate set_seed
        # One decode session (KV pool, state buffers, captured graph) is kept and re-used while the batch size, the
        # sampler configuration and the KV geometry allow it: repeated generate() calls (serving, bench steps) then
        # neither re-allocate ~10 GB nor re-capture the 230-kernel graph.
        cfg_key = bytes(sampler.cfg)
        key = (B, self.kv_paged, self.kv_page_size, cfg_key, tuple(self.config.speech_token_range), int(eos_fill), has_eos,
               eng.use_graph, eos_at is not None, capture)
        sess = getattr(self, "_session", None)
        if sess is None or sess["key"] != key or sess["rows"] < max_rows:
            self._session = None
            sess = None
            rows_cap = max_rows if not getattr(self, "session_headroom", True) else max(max_rows, 1024)
            cache = KVCache(self.shape, B, rows_cap + 1, dev, paged=self.kv_paged, page_size=self.kv_page_size,
                            shuffle_pages=self.kv_paged)
            eng.ctx_hint = max_rows   # rows this call will reach (the session's KV capacity may be larger)
            st = eng.make_decode_state(B, cache, sampler, rows_cap, tuple(self.config.speech_token_range), int(eos_fill),
                                       has_eos)
            if eos_at is not None:
                st["row_ctl"] = torch.zeros((B, 4), dtype=torch.int32, device=dev)
                st["hist_len"] = st["hist"].numel()
                st["mega"] = None
            st["keep_logits"] = capture
            sess = dict(key=key, rows=rows_cap, st=st, cache=cache)
            self._session = sess
        st, cache = sess["st"], sess["cache"]
        max_rows = sess["rows"]
        eng.reset_decode_state(st, seed, P, max_length)
        if eos_at is not None:
            ctl = torch.zeros((B, 4), dtype=torch.int32)
            ctl[:, 1], ctl[:, 2] = P, max_length
            ctl[:, 3] = torch.as_tensor(eos_at, dtype=torch.int32)
            st["row_ctl"].copy_(ctl)
        st["sequences"][:, :P].copy_(input_ids[:, :P])
        st["tf_tail"].copy_(input_ids[:, P:])
        from . import _lib
        import ctypes
        _lib.check(eng.L.mtts_sampler_init_history(input_ids.data_ptr(), B, P, input_ids.stride(0), ctypes.byref(sampler.cfg),
                                                   st["seen"].data_ptr(), _lib.stream_ptr()))
        # ---- step 0: prefill the prompt, sample from its last position
        ev_t = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev_t[0].record()
        logits, lens = eng.prefill(input_ids[:, :P], attention_mask[:, :P], cache)
        st["positions"].copy_((lens - 1).to(torch.int32))
        eng.sample_and_advance(st, logits)
        ev_t[1].record()
        step_logits = [logits.clone()] if capture else None
        if streamer is not None:
            streamer.put(st["tokens"][:, 0].cpu())
        # ---- steps 1..: one graph replay per frame. The host never waits for the GPU inside the loop: after every
        # block of `sync_every` replays it enqueues an async copy of the per-step "unfinished rows" counters into pinned
        # memory and inspects only copies that have already landed, so the stop is noticed at most two blocks late
        # (the surplus rows are dropped below; the reference syncs twice per step, modeling_asteroid.py:149,169).
        steps_done = 1
        final_len = None
        max_steps = max_rows - P - 1
        hard_stop = max(1, max_length - P)  # rows that are not winding down all stop here: check synchronously
        block = 1 if (streamer is not None or capture) else max(1, int(self.sync_every))
        pinned = torch.empty(st["hist"].numel(), dtype=torch.int32).pin_memory()
        pinned_err = torch.zeros(4, dtype=torch.int32).pin_memory()
        pending = []

        def scan(upto):
            # device-side error flags travel with every block: a bad token id / page / sampler overflow surfaces at the
            # next check (<= two blocks of `sync_every` steps later), not after the whole loop
            if int(pinned_err.max()) != 0:
                self._check_err()
            zero = (pinned[:upto] == 0).nonzero()
            return P + int(zero[0]) + 1 if zero.numel() else None

        def post_check():
            pinned[:steps_done].copy_(st["hist"][:steps_done], non_blocking=True)
            pinned_err.copy_(eng.err, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            pending.append((ev, steps_done))

        post_check()
        if streamer is not None:
            pending[-1][0].synchronize()
        while final_len is None:
            if len(pending) >= 2:
                pending[0][0].synchronize()  # back-pressure: at most two blocks of replays are queued ahead of the GPU
            while pending and pending[0][0].query():
                _, upto = pending.pop(0)
                final_len = scan(upto)
                if final_len is not None:
                    break
            if final_len is not None or steps_done >= max_steps:
                break
            n = min(block, max_steps - steps_done)
            if steps_done < hard_stop:
                n = min(n, hard_stop - steps_done)
            else:
                n = 1
            for _ in range(n):
                eng.decode_step(st)
                steps_done += 1
                if capture:
                    step_logits.append(st["logits"].clone())
                if streamer is not None:
                    streamer.put(st["tokens"][:, 0].cpu())
            post_check()
            if streamer is not None or steps_done >= hard_stop:
                pending[-1][0].synchronize()
        if final_len is None:
            torch.cuda.synchronize()
            final_len = scan(steps_done) or (P + steps_done)
        ev_t[2].record()
        self._last_timing = (ev_t, steps_done)
        self._check_err()
        if streamer is not None:
            streamer.end()
        out = st["sequences"][:, :final_len].clone()
        self._last_state = st
        if gc.return_dict_in_generate:
            raw = None
            if capture:
                # one entry per generated row: the list of 8 fp32 (B, V_c) last-position logits with the step's pad / EOS
                # mask applied in place, exactly what the reference appends (modeling_asteroid.py:123-128,176)
                raw = []
                offs, vocabs = self.shape.head_offsets, self.shape.vocabs
                for s_i, lg in enumerate(step_logits[:final_len - P]):
                    per = [lg[:, o:o + v].float() for o, v in zip(offs, vocabs)]
                    for c in range(1, C):
                        if s_i >= c:
                            per[c][:, 1024] = float("-inf")
                    if s_i <= C - 2:
                        per[0][:, 152694] = float("-inf")
                    raw.append(per)
                raw = tuple(raw)
            return GenerateDecoderOnlyOutput(sequences=out, past_key_values=None, logits=raw if gc.output_logits else None,
                                             scores=raw if gc.output_scores else None)
        return out
