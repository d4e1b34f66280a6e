"""TEST INFRASTRUCTURE ONLY. Imports the UNMODIFIED reference from /root/reference on CPU (shims S1-S4).

Only `oracle/gen_golden.py` uses this, and only in the build container: /root/reference does not exist on the
GPU box, so nothing that runs there may import this module. The reference's own code is executed as is; the
shims only (S1) satisfy an unused `import librosa`, (S2) neutralise transformers-5.x weight tying that the
reference's 4.53-era class does not understand, and (S3) supply the three GenerationMixin helpers that
`CustomMixin._sample` calls and that changed signature between transformers 4.53.2 (pinned by the reference,
requirements.txt:3) and the 5.5.0 installed here.
"""
import importlib.machinery
import os
import sys
import types

REF = os.environ.get("MTTS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(REF)


def _install_stubs():
    import transformers  # noqa: F401  (must be imported before the librosa stub, S1)
    if "librosa" not in sys.modules:
        m = types.ModuleType("librosa")
        m.__spec__ = importlib.machinery.ModuleSpec("librosa", None)
        sys.modules["librosa"] = m
    if REF not in sys.path:
        sys.path.insert(0, REF)
    xy = os.path.join(REF, "XY_Tokenizer")
    if xy not in sys.path:
        sys.path.insert(0, xy)


def import_lm():
    """-> the reference's modeling_asteroid module."""
    _install_stubs()
    import modeling_asteroid as ma
    ma.AsteroidTTSInstruct.tie_weights = lambda self, *a, **k: None  # S2
    return ma


def import_codec():
    """-> (XY_Tokenizer class, quantizer module, modules module)."""
    _install_stubs()
    from xy_tokenizer.model import XY_Tokenizer
    from xy_tokenizer.nn import quantizer, modules
    return XY_Tokenizer, quantizer, modules


def import_generation_utils():
    _install_stubs()
    for name in ("torchaudio",):
        try:
            __import__(name)
        except Exception:  # pragma: no cover
            sys.modules[name] = types.ModuleType(name)
    import generation_utils as gu
    return gu


def make_lm(ma, cfg_kwargs, dtype):
    import torch
    cfg = ma.AsteroidTTSConfig(**cfg_kwargs, tie_word_embeddings=False)
    cfg._attn_implementation = "eager"
    model = ma.AsteroidTTSInstruct(cfg).eval().to(dtype)
    return model


def bind_generation_helpers(model):
    """S3: 4.53.2-equivalent helpers bound on the instance so that the reference's `_sample` body runs verbatim."""
    import torch
    from transformers.cache_utils import DynamicCache

    def _get_initial_cache_position(seq_length, device, model_kwargs):
        model_kwargs["cache_position"] = torch.arange(seq_length, device=device)
        return model_kwargs

    def prepare_inputs_for_generation(input_ids, past_key_values=None, attention_mask=None, cache_position=None, **kw):
        if past_key_values is None:
            past_key_values = DynamicCache(config=model.config)
        if past_key_values.get_seq_length() > 0:
            input_ids = input_ids[:, cache_position]
        position_ids = attention_mask.long().cumsum(-1) - 1
        position_ids.masked_fill_(attention_mask == 0, 1)
        position_ids = position_ids[:, -input_ids.shape[1]:]
        return dict(input_ids=input_ids, past_key_values=past_key_values, attention_mask=attention_mask,
                    position_ids=position_ids, cache_position=cache_position, use_cache=True)

    def _update_model_kwargs_for_generation(outputs, model_kwargs, **kw):
        model_kwargs["past_key_values"] = outputs.past_key_values
        am = model_kwargs["attention_mask"]
        model_kwargs["attention_mask"] = torch.cat([am, am.new_ones((am.shape[0], 1))], dim=-1)
        model_kwargs["cache_position"] = model_kwargs["cache_position"][-1:] + 1
        return model_kwargs

    model._get_initial_cache_position = _get_initial_cache_position
    model.prepare_inputs_for_generation = prepare_inputs_for_generation
    model._update_model_kwargs_for_generation = _update_model_kwargs_for_generation
    return model


def run_sample(model, input_ids, attention_mask, max_length, gen_cfg_overrides=None, eos_token_id=152694):
    """Calls the reference's CustomMixin._sample exactly as HF generate() would."""
    import torch
    from transformers.generation.configuration_utils import GenerationConfig
    from transformers.generation.logits_process import LogitsProcessorList
    from transformers.generation.stopping_criteria import (EosTokenCriteria, MaxLengthCriteria, StoppingCriteriaList)
    gc = GenerationConfig()
    gc.eos_token_id = eos_token_id
    gc.max_length = max_length
    gc.do_sample = False
    gc.output_attentions = False
    gc.output_hidden_states = False
    gc.output_scores = False
    gc.output_logits = False
    gc.return_dict_in_generate = False
    gc.do_samples = None
    for k, v in (gen_cfg_overrides or {}).items():
        setattr(gc, k, v)
    crit = [MaxLengthCriteria(max_length)]
    if eos_token_id is not None:
        crit.append(EosTokenCriteria(eos_token_id))
    with torch.no_grad():
        return model._sample(input_ids, LogitsProcessorList(), StoppingCriteriaList(crit), gc, False, None,
                             attention_mask=attention_mask)
