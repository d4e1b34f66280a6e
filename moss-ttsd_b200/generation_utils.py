"""Drop-in mirror of the reference's batch pipeline `generation_utils.py` (same function names, argument meaning,
return values and error behaviour; /root/reference/generation_utils.py):

  load_model (:15-24) · process_jsonl_item (:27-87) · load_audio_data (:90-127) · process_inputs (:180-208) ·
  shifting_inputs (:211-218) · rpadding (:221-237) · find_max_valid_positions (:240-249) · normalize_text (:252-338) ·
  process_batch (:341-477)

The host-side logic is restated here; `model.generate`, `spt.encode` and `spt.decode` are the B200 implementations
of this package. Differences that do not change results: the un-delay step is one gather instead of eight strided
copies, and finished samples of EQUAL length are decoded by the codec in one batched call (items of equal length
in one batch see exactly what they would see alone; see SURVEY.md §7 "codec batch-composition leak").
"""
from __future__ import annotations

import os
import re
from collections import defaultdict

import numpy as np
import torch

MAX_CHANNELS = 8
SILENCE_DURATION = 0.0
SPEECH_OFFSET = 151665  # channel-0 offset of codec codes (generation_utils.py:202,425)


def load_model(model_path, spt_config_path, spt_checkpoint_path, torch_dtype=torch.bfloat16,
               attn_implementation="flash_attention_2"):
    """-> (tokenizer, model, spt). `attn_implementation` is accepted and ignored (attention is a libmtts kernel)."""
    from transformers import AutoTokenizer
    from .modeling_asteroid import AsteroidTTSInstruct
    from .xy_tokenizer.model import XY_Tokenizer
    tokenizer = AutoTokenizer.from_pretrained(model_path)
    model = AsteroidTTSInstruct.from_pretrained(model_path, torch_dtype=torch_dtype, attn_implementation=attn_implementation)
    spt = XY_Tokenizer.load_from_checkpoint(config_path=spt_config_path, ckpt_path=spt_checkpoint_path)
    model.eval()
    spt.eval()
    return tokenizer, model, spt


# ---------------------------------------------------------------------------------------------- JSONL items
def _join_base(base_path, p):
    return os.path.join(base_path, p) if (isinstance(p, str) and base_path and p) else p


def process_jsonl_item(item):
    """Extracts {"text", "prompt_text", "prompt_audio"} from the three JSONL schemas of examples/*.jsonl."""
    base_path = item.get("base_path", "")
    text = item.get("text", "")
    prompt_audio, prompt_text = None, ""
    if "prompt_audio" in item and "prompt_text" in item:
        print("Using prompt_audio and prompt_text directly from item.")
        if item["prompt_audio"]:
            prompt_audio = _join_base(base_path, item["prompt_audio"])
            prompt_text = item["prompt_text"]
    else:
        a1, t1 = item.get("prompt_audio_speaker1", ""), item.get("prompt_text_speaker1", "")
        a2, t2 = item.get("prompt_audio_speaker2", ""), item.get("prompt_text_speaker2", "")
        has = lambda a: (isinstance(a, str) and bool(a)) or isinstance(a, tuple)
        if has(a1) or has(a2):
            print("Using speaker1 and speaker2 information for prompt audio and text.")
            prompt_audio = {"speaker1": _join_base(base_path, a1), "speaker2": _join_base(base_path, a2)}
        merged = (f"[S1]{t1}" if t1 else "") + (f"[S2]{t2}" if t2 else "")
        prompt_text = merged.strip()
    return {"text": text, "prompt_text": prompt_text, "prompt_audio": prompt_audio}


def _load_single_audio(audio_input):
    if isinstance(audio_input, tuple) and len(audio_input) == 2:
        return audio_input
    if isinstance(audio_input, str):
        import torchaudio
        return torchaudio.load(audio_input)
    raise ValueError(f"Unsupported audio input format: {type(audio_input)}")


def _to_mono_16k(wav, sr, target_sample_rate):
    if sr != target_sample_rate:
        import torchaudio
        wav = torchaudio.functional.resample(wav, sr, target_sample_rate)
    if wav.shape[0] > 1:
        wav = wav.mean(dim=0, keepdim=True)
    if wav.dim() == 1:
        wav = wav.unsqueeze(0)
    return wav


def merge_speaker_audios(wav1, sr1, wav2, sr2, target_sample_rate=16000):
    try:
        return torch.cat([_to_mono_16k(wav1, sr1, target_sample_rate), _to_mono_16k(wav2, sr2, target_sample_rate)], dim=1)
    except Exception as e:
        print(f"Error merging audio: {e}")
        raise


def load_audio_data(prompt_audio, target_sample_rate=16000):
    """str path | (wav, sr) | {"speaker1": ..., "speaker2": ...} -> (1, N) mono tensor at 16 kHz (speakers concatenated)."""
    if prompt_audio is None:
        return None
    try:
        if isinstance(prompt_audio, dict) and "speaker1" in prompt_audio and "speaker2" in prompt_audio:
            w1, s1 = _load_single_audio(prompt_audio["speaker1"])
            w2, s2 = _load_single_audio(prompt_audio["speaker2"])
            return merge_speaker_audios(w1, s1, w2, s2, target_sample_rate)
        wav, sr = _load_single_audio(prompt_audio)
        return _to_mono_16k(wav, sr, target_sample_rate)
    except Exception as e:
        print(f"Error loading audio data: {e}")
        raise


# ---------------------------------------------------------------------------------------------- token grids
def process_inputs(tokenizer, spt, prompt, text, device, audio_data=None, max_channels=8, pad_token=1024):
    seq = f"<|begin_of_style|>{prompt}<|end_of_style|>\n<|begin_of_text|>{text}<|end_of_text|>\n<|begin_of_speech|>"
    text_ids = np.array(tokenizer.encode(seq))
    input_ids = np.full((text_ids.shape[0], max_channels), pad_token)
    input_ids[:, 0] = text_ids
    if audio_data is not None:
        try:
            wav = audio_data
            silence = torch.zeros(wav.shape[0], int(SILENCE_DURATION * 16000))
            wav = torch.cat([wav, silence], dim=1)
            with torch.no_grad():
                enc = spt.encode([wav.squeeze().to(device)])
                audio_token = enc["codes_list"][0].permute(1, 0).cpu().numpy()
            audio_token[:, 0] = audio_token[:, 0] + SPEECH_OFFSET
            input_ids = np.concatenate([input_ids, audio_token])
        except Exception as e:
            print(f"Error processing audio data: {e}")
            raise
    return input_ids


def shifting_inputs(input_ids, tokenizer, pad_token=1024, max_channels=8):
    """Delay pattern: channel j is pushed down by j rows; the 7 extra rows hold pad_token_id (ch0) / 1024 (others)."""
    n = input_ids.shape[0]
    out = np.full((n + max_channels - 1, max_channels), pad_token, dtype=np.int64)
    out[:, 0] = tokenizer.pad_token_id
    for j in range(max_channels):
        out[j:n + j, j] = input_ids[:, j]
    return out


def rpadding(input_ids, channels, tokenizer):
    """LEFT-pads a list of (n_i, channels) grids to the longest; returns (int64 ids (B,T,C), float64 mask (B,T))."""
    T = max(g.shape[0] for g in input_ids)
    ids = np.full((len(input_ids), T, channels), 1024, dtype=np.int64)
    ids[:, :, 0] = tokenizer.pad_token_id
    mask = np.zeros((len(input_ids), T), dtype=np.float64)
    for b, g in enumerate(input_ids):
        if g.shape[0]:
            ids[b, T - g.shape[0]:] = g
            mask[b, T - g.shape[0]:] = 1.0
    return torch.tensor(ids), torch.tensor(mask)


def find_max_valid_positions(C: torch.Tensor, invalid_value=1024) -> torch.Tensor:
    """Last row t with C[b, t, 1] != invalid_value, or -1."""
    ok = C[:, :, 1] != invalid_value
    idx = torch.arange(C.size(1), device=C.device).expand_as(ok)
    return torch.where(ok, idx, torch.full_like(idx, -1)).max(dim=1).values


# ---------------------------------------------------------------------------------------------- text normalisation
_DECOR = "【】《》（）『』「」\"-“”～~"
_DECOR_RE = re.compile("[" + re.escape(_DECOR) + "]")
_PUNCT = str.maketrans({"！": "，", "!": ",", "；": "，", ";": ",", "：": "，", ":": ",", "、": "，", "？": "，", "?": ","})


def _clean_segment(content: str) -> str:
    content = _DECOR_RE.sub("", content)
    content = re.sub(r"哈{2,}", "(笑)", content)
    content = re.sub(r"\b(ha(\s*ha)+)\b", "(laughs)", content, flags=re.IGNORECASE)
    content = content.replace("——", "，").replace("……", "，")
    content = content.translate(_PUNCT).strip()
    if len(content) > 1:  # only the final sentence mark survives
        tail = {"，": "。", ",": "."}.get(content[-1], content[-1])
        content = content[:-1].replace("。", "，") + tail
    return content


def normalize_text(text: str) -> str:
    """Same rewrite rules as the reference (generation_utils.py:252-338): [n] -> [Sn], non-speaker brackets dropped,
    decorative symbols removed, laughter normalised, inner punctuation folded to commas, one final period per
    segment, adjacent segments of one speaker merged."""
    text = re.sub(r"\[(\d+)\]", r"[S\1]", text)
    text = re.sub(r"\[(?!S\d+\])([^\]]*)\]", r"\1", text)
    parts = []
    for seg in re.split(r"(?=\[S\d+\])", text.replace("\n", " ")):
        seg = seg.strip()
        if not seg:
            continue
        m = re.match(r"^(\[S\d+\])\s*(.*)", seg)
        tag, content = m.groups() if m else ("", seg)
        parts.append([tag, _clean_segment(content)])
    if not parts:
        return ""
    merged = [parts[0]]
    for tag, content in parts[1:]:
        if tag == merged[-1][0] and tag:
            merged[-1][1] += content
        else:
            merged.append([tag, content])
    out = "".join(f"{t}{c}".strip() for t, c in merged)
    return out.replace("‘", "'").replace("’", "'")


# ---------------------------------------------------------------------------------------------- batch pipeline
def undelay(outputs: torch.Tensor, channels: int = MAX_CHANNELS) -> torch.Tensor:
    """speech[b, t, j] = outputs[b, t + j, j]; channel 0 minus 151665 (generation_utils.py:416-425)."""
    B, G, C = outputs.shape
    seq_len = G - channels + 1
    t = torch.arange(seq_len, device=outputs.device)[:, None] + torch.arange(C, device=outputs.device)[None, :]
    speech = outputs.gather(1, t[None].expand(B, seq_len, C))
    speech[..., 0] -= SPEECH_OFFSET
    return speech


def _prepare_item(item, start_idx, i, use_normalize):
    """Host-only part of one item (generation_utils.py:362-392): schema, text normalisation, prompt-audio load/resample."""
    p = process_jsonl_item(item)
    full_text = p["prompt_text"] + p["text"] if p["prompt_text"] else p["text"]
    original_full_text = full_text
    if use_normalize:
        full_text = normalize_text(full_text)
    final_text = full_text.replace("[S1]", "<speaker1>").replace("[S2]", "<speaker2>")
    meta = {
        "index": start_idx + i,
        "original_text": original_full_text,
        "normalized_text": normalize_text(original_full_text) if use_normalize else None,
        "final_text": final_text,
        "use_normalize": use_normalize,
    }
    audio = load_audio_data(p["prompt_audio"]) if p["prompt_audio"] else None
    return meta, final_text, audio


def encode_prompt_audios(spt, audios, device, group: int = 32):
    """All prompt audios of a batch through `spt.encode` in groups of `group` items instead of one call per item
    (generation_utils.py:198): the codec treats every item of a call independently (own log-mel, own length mask), so
    the codes equal the per-item calls'. -> list aligned with `audios` of (n_i, 8) numpy grids with channel 0 offset by
    151665, or None."""
    out = [None] * len(audios)
    todo = [(i, a) for i, a in enumerate(audios) if a is not None]
    for g0 in range(0, len(todo), group):
        part = todo[g0:g0 + group]
        wavs = []
        for _, wav in part:
            silence = torch.zeros(wav.shape[0], int(SILENCE_DURATION * 16000))
            wavs.append(torch.cat([wav, silence], dim=1).squeeze().to(device))
        with torch.no_grad():
            codes = spt.encode(wavs)["codes_list"]
        for (i, _), c in zip(part, codes):
            tok = c.permute(1, 0).cpu().numpy()
            tok[:, 0] = tok[:, 0] + SPEECH_OFFSET
            out[i] = tok
    return out


def prepare_batch(batch_items, tokenizer, spt, device, system_prompt, start_idx, use_normalize=False, pool=None):
    """Prompt side of `process_batch` (generation_utils.py:362-404): -> (actual_texts_data, input_ids (B,T,8) int64,
    attention_mask (B,T) float64). Host work per item (schema, normalisation, audio file load + resample) runs on `pool`
    (a concurrent.futures executor) when given; all prompt audios are encoded in batched `spt.encode` calls."""
    n = len(batch_items)
    if pool is not None:
        prepared = list(pool.map(lambda t: _prepare_item(t[1], start_idx, t[0], use_normalize), enumerate(batch_items)))
    else:
        prepared = [_prepare_item(item, start_idx, i, use_normalize) for i, item in enumerate(batch_items)]
    metas = [p[0] for p in prepared]
    try:
        audio_tokens = encode_prompt_audios(spt, [p[2] for p in prepared], device)
    except Exception as e:
        print(f"Error processing audio data: {e}")
        raise
    grids = []
    for (meta, text, _), tok in zip(prepared, audio_tokens):
        g = process_inputs(tokenizer, spt, system_prompt, text, device, None)
        if tok is not None:
            g = np.concatenate([g, tok])
        grids.append(shifting_inputs(g, tokenizer))
    input_ids, attention_mask = rpadding(grids, MAX_CHANNELS, tokenizer)
    return metas, input_ids, attention_mask


def finish_batch(speech_ids, ends, wavs, spt, start_idx):
    """(speech ids, per-row end, decoded waveforms or None) -> the reference's audio_results list (:456-467)."""
    audio_results = [None] * len(ends)
    for i, (e, w) in enumerate(zip(ends, wavs)):
        if e <= 0:
            print(f"Sample {start_idx + i} has no valid speech tokens")
            continue
        if w is None:
            continue
        w = w.cpu().detach()
        if w.ndim == 1:
            w = w.unsqueeze(0)
        audio_results[i] = {"audio_data": w, "sample_rate": spt.output_sample_rate, "index": start_idx + i}
        print(f"Audio generation completed: sample {start_idx + i}")
    return audio_results


def _decode_rows(spt, speech_ids, ends, start_idx):
    """Codec decode of every row with speech; rows of EQUAL length share one call (they see exactly what they would see
    alone). A failing group is retried row by row, so that — like the reference, which decodes one sample per call and
    skips the one that raised (generation_utils.py:448-467) — only the failing sample yields None."""
    wavs = [None] * len(ends)
    by_len = defaultdict(list)
    for i, e in enumerate(ends):
        if e > 0:
            by_len[e].append(i)
    for e, idxs in by_len.items():
        try:
            got = spt.decode([speech_ids[i, :e].permute(1, 0) for i in idxs], overlap_seconds=10)["syn_wav_list"]
            for i, w in zip(idxs, got):
                wavs[i] = w
        except Exception:
            for i in idxs:
                try:
                    wavs[i] = spt.decode([speech_ids[i, :e].permute(1, 0)], overlap_seconds=10)["syn_wav_list"][0]
                except Exception as ex:
                    print(f"Error processing sample {start_idx + i}: {str(ex)}, skipping...")
                    import traceback
                    traceback.print_exc()
    return wavs


def process_batch(batch_items, tokenizer, model, spt, device, system_prompt, start_idx, use_normalize=False):
    """-> (actual_texts_data, audio_results); see the reference for the dict layouts (generation_utils.py:374-380,456-460)."""
    try:
        batch_size = len(batch_items)
        print(f"Processing {batch_size} samples starting from index {start_idx}...")
        actual_texts_data, input_ids, attention_mask = prepare_batch(batch_items, tokenizer, spt, device, system_prompt,
                                                                     start_idx, use_normalize)
        print("Starting batch audio generation...")
        start = input_ids.shape[1] - MAX_CHANNELS + 1
        outputs = model.generate(input_ids=input_ids.to(device), attention_mask=attention_mask.to(device))
        print(f"Original outputs shape: {outputs.shape}")
        speech_ids = undelay(outputs[:, start:])
        ends = (find_max_valid_positions(speech_ids) + 1).cpu().tolist()
        wavs = _decode_rows(spt, speech_ids, ends, start_idx)
        audio_results = finish_batch(speech_ids, ends, wavs, spt, start_idx)
        torch.cuda.empty_cache()
        return actual_texts_data, audio_results
    except Exception as e:
        print(f"Error during batch processing: {str(e)}")
        raise


def process_batches(items, tokenizer, model, spt, device, system_prompt, batch_size, start_idx=0, use_normalize=False,
                    bucket_by_length=True, overlap=True, workers=4):
    """Pipelined form of the reference's driver loop (inference.py:73-101 calls `process_batch` once per batch): yields
    `(actual_texts_data, audio_results)` per batch, with three stages in flight —
      * prompt side of batch i+1 (JSONL schema, text normalisation, audio load/resample) on a host thread pool (§8f-3),
      * LM decode of batch i on the main stream,
      * codec decode of batch i-1 on a second stream (§8f-4, `pipeline.CodecStage`).
    bucket_by_length: batches are formed from scripts of similar text length (`scheduler.length_bucketed_batches`; a batch
    decodes until its longest row is done) instead of arrival order; every result carries its item's `index`."""
    from concurrent.futures import ThreadPoolExecutor
    from .pipeline import CodecStage
    from . import scheduler
    n = len(items)
    if n == 0:
        return
    if bucket_by_length:
        est = [len(process_jsonl_item(it)["text"]) for it in items]
        groups = scheduler.length_bucketed_batches(list(range(n)), est, batch_size)
    else:
        groups = scheduler.batches(list(range(n)), batch_size)
    stage = CodecStage(spt, device, overlap=overlap)
    host_pool = ThreadPoolExecutor(max_workers=max(1, workers))

    def prep_host(idxs):
        return [_prepare_item(items[j], start_idx + j, 0, use_normalize) for j in idxs]

    def collect(job, metas, idxs):
        job.wait()
        res = [None] * len(idxs)
        for k, (e, w) in enumerate(zip(job.ends, job.wavs)):
            if e <= 0:
                print(f"Sample {metas[k]['index']} has no valid speech tokens")
            elif w is not None:
                w = w.cpu().detach()
                res[k] = {"audio_data": w.unsqueeze(0) if w.ndim == 1 else w, "sample_rate": spt.output_sample_rate,
                          "index": metas[k]["index"]}
        return metas, res

    try:
        fut = host_pool.submit(prep_host, groups[0])
        prev = None
        for gi, idxs in enumerate(groups):
            prepared = fut.result()
            if gi + 1 < len(groups):
                fut = host_pool.submit(prep_host, groups[gi + 1])
            metas = [p[0] for p in prepared]
            audio_tokens = encode_prompt_audios(spt, [p[2] for p in prepared], device)
            grids = []
            for (meta, text, _), tok in zip(prepared, audio_tokens):
                g = process_inputs(tokenizer, spt, system_prompt, text, device, None)
                grids.append(shifting_inputs(np.concatenate([g, tok]) if tok is not None else g, tokenizer))
            input_ids, attention_mask = rpadding(grids, MAX_CHANNELS, tokenizer)
            start = input_ids.shape[1] - MAX_CHANNELS + 1
            outputs = model.generate(input_ids=input_ids.to(device), attention_mask=attention_mask.to(device))
            job = stage.submit(outputs, start, index=gi)
            if prev is not None:
                yield collect(*prev)
            prev = (job, metas, idxs)
        if prev is not None:
            yield collect(*prev)
    finally:
        host_pool.shutdown(wait=False)
