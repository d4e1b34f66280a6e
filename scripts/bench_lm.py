"""Decode-step timing of the v0.5-shaped LM (random init) at a few batch sizes: ms/step and HBM GB/s."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from moss_ttsd_b200.modeling_asteroid import AsteroidTTSConfig, AsteroidTTSInstruct

SHAPE = dict(hidden_size=2048, intermediate_size=6144, num_hidden_layers=28, num_attention_heads=16,
             num_key_value_heads=8, head_dim=128, rms_norm_eps=1e-6, rope_theta=1e6, vocab_size=152697,
             speech_vocab_size=1025, channels=8, speech_token_range=[151665, 152689])


def make_prompt(rng, B, text_rows, audio_rows):
    lo, hi, C = 151665, 152689, 8
    n = text_rows + audio_rows
    g = np.full((B, n, C), 1024, dtype=np.int64)
    g[:, :text_rows, 0] = rng.integers(0, 151000, (B, text_rows))
    g[:, text_rows:, 0] = rng.integers(lo, hi, (B, audio_rows))
    g[:, text_rows:, 1:] = rng.integers(0, 1024, (B, audio_rows, C - 1))
    sh = np.full((B, n + C - 1, C), 1024, dtype=np.int64)
    sh[:, :, 0] = 151643
    for i in range(C):
        sh[:, i:n + i, i] = g[:, :, i]
    return sh, np.ones((B, n + C - 1))


def main():
    cfg = AsteroidTTSConfig(**SHAPE, eos_token_id=152694, pad_token_id=151643, tie_word_embeddings=True)
    m = AsteroidTTSInstruct(cfg, device="cuda")
    m.init_random_weights(0)
    m._w.heads[:151665].zero_(); m._w.heads[152689:152704].zero_()  # H3: greedy stays in the speech range
    m.generation_config.eos_token_id = 152694
    rng = np.random.default_rng(0)
    rows = []
    batches = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 16, 64]
    for B, new in [(b, 64) for b in batches]:
        ids, mask = make_prompt(rng, B, int(sys.argv[2]) if len(sys.argv) > 2 else 200, int(sys.argv[3]) if len(sys.argv) > 3 else 250)
        ids, mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
        T = ids.shape[1]
        for it in range(2):
            torch.cuda.synchronize(); t0 = time.time()
            out = m.generate(input_ids=ids, attention_mask=mask, max_new_tokens=new)
            torch.cuda.synchronize(); dt = time.time() - t0
        st = m._last_state
        # time the graph replay alone
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        for _ in range(20):
            st["graph"].replay()
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        ctx = T + new
        bytes_step = m._w.nbytes() + B * ctx * 114688
        rows.append(dict(B=B, T=T, new=new, out=list(out.shape), gen_s=round(dt, 3), ms_per_step=round(ms, 4),
                         GBs=round(bytes_step / ms / 1e6, 1), rtf=round(B * 0.08 / (ms / 1e3), 1)))
        print(rows[-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/bench_lm.json", "w"), indent=1)


if __name__ == "__main__":
    main()
