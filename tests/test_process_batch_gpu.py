"""process_batch end to end on the GPU with tiny random-init models and a stand-in tokenizer (no tokenizer/weights are
available offline): return layout, text-only and prompt-audio items, and agreement with the staged
pipeline run by hand."""
import copy

import numpy as np
import pytest
import torch

from tests.common import TINY, tiny_model

pytestmark = pytest.mark.gpu


class Tok:
    pad_token_id = 151643

    def encode(self, s):
        return [(ord(c) * 7 + i) % 151000 for i, c in enumerate(s)]


def _codec():
    from moss_ttsd_b200.xy_tokenizer.model import XY_Tokenizer
    from oracle.codec_weights import TINY_CODEC, make_codec_weights, make_encoder_weights
    gp = copy.deepcopy(TINY_CODEC)
    gp["quantizer_kwargs"]["codebook_size"] = 1024   # LM speech channels emit ids in [0, 1024)
    sd = make_codec_weights(gp, 3)
    sd.update(make_encoder_weights(gp, 4))
    spt = XY_Tokenizer(gp)
    spt.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return spt.to("cuda")


def test_process_batch_text_and_prompt_audio():
    from moss_ttsd_b200 import generation_utils as gu
    model, _ = tiny_model()
    model.generation_config.eos_token_id = 152694
    model.generation_config.max_new_tokens = 12
    spt = _codec()
    rng = np.random.default_rng(0)
    wav = torch.from_numpy((0.1 * rng.standard_normal(16000 * 2)).astype(np.float32))[None]
    items = [
        {"text": "[S1]Hello there.[S2]Hi!"},
        {"text": "[S1]Second item, a bit longer than the first one.", "prompt_audio": (wav, 16000), "prompt_text": "[S1]ref"},
        {"text": "[S1]a", "prompt_audio_speaker1": (wav, 16000), "prompt_text_speaker1": "one",
         "prompt_audio_speaker2": (wav[:, :16000], 16000), "prompt_text_speaker2": "two"},
    ]
    texts, audios = gu.process_batch(items, Tok(), model, spt, "cuda", "You are a speech synthesizer.", start_idx=5,
                                     use_normalize=True)
    assert [t["index"] for t in texts] == [5, 6, 7]
    assert texts[0]["final_text"].startswith("<speaker1>") and texts[0]["use_normalize"] is True
    assert texts[2]["original_text"] == "[S1]one[S2]two[S1]a"
    assert len(audios) == 3
    for i, a in enumerate(audios):
        assert a is not None and a["sample_rate"] == 24000 and a["index"] == 5 + i
        w = a["audio_data"]
        assert w.device.type == "cpu" and w.dim() == 2 and w.shape[0] == 1 and w.dtype == torch.float32
        assert w.shape[1] % 1920 == 0 and 0 < w.shape[1] <= 12 * 1920 and torch.isfinite(w).all()
    # the same thing staged by hand gives the same waveform for item 0
    g0 = gu.shifting_inputs(gu.process_inputs(Tok(), spt, "You are a speech synthesizer.", texts[0]["final_text"], "cuda"), Tok())
    g1 = gu.shifting_inputs(gu.process_inputs(Tok(), spt, "You are a speech synthesizer.", texts[1]["final_text"], "cuda",
                                              gu.load_audio_data((wav, 16000))), Tok())
    g2 = gu.shifting_inputs(gu.process_inputs(Tok(), spt, "You are a speech synthesizer.", texts[2]["final_text"], "cuda",
                                              gu.load_audio_data({"speaker1": (wav, 16000), "speaker2": (wav[:, :16000], 16000)})), Tok())
    assert g1.shape[0] - (len(Tok().encode("x")) - 1) > g0.shape[0] - 100  # prompt audio rows were appended
    ids, mask = gu.rpadding([g0, g1, g2], 8, Tok())
    out = model.generate(input_ids=ids.cuda(), attention_mask=mask.cuda())
    speech = gu.undelay(out[:, ids.shape[1] - 7:])
    end = int(gu.find_max_valid_positions(speech)[0]) + 1
    w0 = spt.decode([speech[0, :end].permute(1, 0)])["syn_wav_list"][0].cpu()
    # decoded alone vs decoded inside the equal-length group: same arithmetic up to the TF32 GEMMs' tile / split-K
    # configuration, which depends on the row count
    ref = audios[0]["audio_data"][0].double()
    snr = 10 * torch.log10((ref ** 2).sum() / ((ref - w0.double()) ** 2).sum().clamp_min(1e-30))
    assert w0.shape == ref.shape and snr >= 40.0, snr


def test_process_batch_sample_without_audio_returns_none():
    """A row whose channel 1 never leaves the pad value has no audio: None, like the reference (generation_utils.py:437-440)."""
    from moss_ttsd_b200 import generation_utils as gu
    speech = torch.full((2, 6, 8), 1024)
    speech[1, :3, 1] = 7
    assert gu.find_max_valid_positions(speech).tolist() == [-1, 2]


def _items(rng, n):
    wav = torch.from_numpy((0.1 * rng.standard_normal(16000 * 2)).astype(np.float32))[None]
    items = []
    for i in range(n):
        it = {"text": "[S1]" + "word " * int(rng.integers(2, 14)) + f"item {i}."}
        if i % 2:
            it.update(prompt_audio=(wav * (0.5 + 0.1 * i), 16000), prompt_text="[S1]ref")
        items.append(it)
    return items


def test_process_batches_pipelined_equals_process_batch():
    """The pipelined driver (host worker pool -> LM on the main stream -> codec on a second stream) yields, batch by
    batch, what `process_batch` returns for the same batches (arrival-order batches so that batch composition is equal)."""
    from moss_ttsd_b200 import generation_utils as gu
    model, _ = tiny_model()
    model.generation_config.eos_token_id = 152694
    model.generation_config.max_new_tokens = 10
    spt = _codec()
    items = _items(np.random.default_rng(3), 7)
    sys_prompt = "You are a speech synthesizer."
    want = []
    for b0 in range(0, len(items), 3):
        want.append(gu.process_batch(items[b0:b0 + 3], Tok(), model, spt, "cuda", sys_prompt, start_idx=b0))
    for overlap in (True, False):
        got = list(gu.process_batches(items, Tok(), model, spt, "cuda", sys_prompt, batch_size=3, bucket_by_length=False,
                                      overlap=overlap, workers=2))
        assert len(got) == len(want)
        for (tw, aw), (tg, ag) in zip(want, got):
            assert tw == tg
            for a, b in zip(aw, ag):
                assert (a is None) == (b is None)
                if a is not None:
                    assert a["index"] == b["index"] and torch.equal(a["audio_data"], b["audio_data"])
    # length-bucketed batches: every item comes back exactly once, under its own index
    got = list(gu.process_batches(items, Tok(), model, spt, "cuda", sys_prompt, batch_size=3, bucket_by_length=True))
    idx = sorted(t["index"] for texts, _ in got for t in texts)
    assert idx == list(range(len(items)))
    lens = [[len(items[t["index"]]["text"]) for t in texts] for texts, _ in got]
    assert all(min(a) >= max(b) for a, b in zip(lens, lens[1:]))           # longest scripts first, similar lengths together


def test_batched_prompt_encode_equals_per_item_encode():
    """All prompt audios of a batch in one `spt.encode` call give the codes of the reference's one-call-per-item loop
    (generation_utils.py:198)."""
    from moss_ttsd_b200 import generation_utils as gu
    spt = _codec()
    rng = np.random.default_rng(8)
    audios = [torch.from_numpy((0.1 * rng.standard_normal(n)).astype(np.float32))[None] for n in (16000 * 2, 16000 * 5, 9000)]
    audios.insert(1, None)
    batched = gu.encode_prompt_audios(spt, audios, "cuda")
    assert batched[1] is None
    for a, tok in zip(audios, batched):
        if a is None:
            continue
        alone = spt.encode([a.squeeze().cuda()])["codes_list"][0].permute(1, 0).cpu().numpy()
        alone[:, 0] += 151665
        assert tok.shape == alone.shape and (tok == alone).all()


def test_decode_failure_nulls_only_the_failing_sample(monkeypatch):
    """A codec failure on one sample yields None for that sample only (generation_utils.py:463-467), also when equal-length
    samples were grouped into one decode call."""
    from moss_ttsd_b200 import generation_utils as gu
    spt = _codec()
    speech = torch.randint(0, 1024, (3, 6, 8), device="cuda")
    real = spt.decode
    poisoned = int(speech[1, 0, 0])

    def flaky(codes_list, **kw):
        if any(int(c[0, 0]) == poisoned and c.shape[-1] == 6 and torch.equal(c, speech[1].permute(1, 0)) for c in codes_list):
            raise RuntimeError("injected codec failure")
        return real(codes_list, **kw)

    monkeypatch.setattr(spt, "decode", flaky)
    wavs = gu._decode_rows(spt, speech, [6, 6, 6], 0)
    assert wavs[1] is None and wavs[0] is not None and wavs[2] is not None
