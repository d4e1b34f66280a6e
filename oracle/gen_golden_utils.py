"""TEST INFRASTRUCTURE. Known-answer vectors for the host-side helpers, produced by the reference's own
generation_utils.py (build container only)."""
import json
import os

import numpy as np
import torch

from oracle import ref_shims

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TEXTS = [
    "[S1]Hello!! How are you?[S2]haha, fine... [S2]ok; bye.",
    "[1]你好！哈哈哈，今天——天气不错。好的。[2]嗯……[笑] 是的：对",
    "no tags at all: just text; with stuff!",
    "[S1]  leading spaces\n and a newline.[S1]same speaker again?[S2]【书名】《标题》（括号）『a』「b」~～",
    "[S2]Ha ha ha that's funny -- really \"quoted\" ‘single’ ’apos’",
    "",
    "[S1]。",
    "[S1]一。二。三，[S2]a.b.c,",
    "[S3]third speaker[S10]tenth[x]bracket[S1]",
    "[S1]哈哈[S1]哈[S2]HAHA ha",
    "[S1]end with comma，[S2]end with comma,",
]


class Tok:
    pad_token_id = 151643

    def encode(self, s):
        return [(ord(c) * 7 + i) % 151000 for i, c in enumerate(s)]


def main():
    gu = ref_shims.import_generation_utils()
    out = {"normalize": [[t, gu.normalize_text(t)] for t in TEXTS]}
    tok = Tok()
    g = np.full((8, 8), 1024, dtype=np.int64)
    g[:5, 0] = [11, 12, 13, 14, 15]
    a = (8 * np.arange(3)[:, None] + np.arange(8)[None, :]).astype(np.int64)
    a[:, 0] += 151665
    g[5:] = a
    sh = gu.shifting_inputs(g, tok)
    out["shift_in"] = g.tolist()
    out["shift_out"] = sh.tolist()
    ids, mask = gu.rpadding([sh, sh[2:]], 8, tok)
    out["rpad_ids"] = ids.tolist()
    out["rpad_mask"] = mask.tolist()
    out["rpad_dtypes"] = [str(ids.dtype), str(mask.dtype)]
    C = torch.full((3, 6, 8), 7)
    C[0, :, 1] = torch.tensor([5, 5, 5, 5, 1024, 1024])
    C[1, :, 1] = 1024
    C[2, :, 1] = torch.tensor([1024, 3, 1024, 3, 3, 3])
    out["fmv_in"] = C.tolist()
    out["fmv_out"] = gu.find_max_valid_positions(C).tolist()
    items = [
        {"text": "[S1]hi", "prompt_audio": "a.wav", "prompt_text": "[S1]p", "base_path": "/data"},
        {"text": "[S1]hi", "prompt_audio": "", "prompt_text": "[S1]p"},
        {"text": "t", "prompt_audio_speaker1": "s1.wav", "prompt_text_speaker1": "one", "prompt_audio_speaker2": "s2.wav",
         "prompt_text_speaker2": "two", "base_path": "/b"},
        {"text": "t", "prompt_text_speaker1": "one"},
        {"text": "only text"},
    ]
    out["items"] = [[it, gu.process_jsonl_item(it)] for it in items]
    out["process_inputs_text_only"] = gu.process_inputs(tok, None, "sys", "<speaker1>hello", "cpu").tolist()
    with open(os.path.join(GOLD, "utils.json"), "w") as f:
        json.dump(out, f, ensure_ascii=False, indent=1)
    print("utils golden written:", len(out["normalize"]), "normalize cases")


if __name__ == "__main__":
    main()
