"""Experiment: one batch-B decode graph against two batch-B/2 decode graphs replayed concurrently on two streams (the GEMM
chain of one half under the HBM-bound attention of the other). ms per frame of B rows, v0.5 shape, ctx ~460+."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scripts.bench_lm import SHAPE, make_prompt
from moss_ttsd_b200.modeling_asteroid import AsteroidTTSConfig, AsteroidTTSInstruct

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
text = int(sys.argv[2]) if len(sys.argv) > 2 else 200


def mk():
    cfg = AsteroidTTSConfig(**SHAPE, eos_token_id=152694, pad_token_id=151643, tie_word_embeddings=True)
    m = AsteroidTTSInstruct(cfg, device="cuda")
    m.init_random_weights(0)
    m._w.heads[:151665].zero_(); m._w.heads[152689:152704].zero_()
    m.generation_config.eos_token_id = 152694
    return m


def session(m, b, seed):
    ids, mask = make_prompt(np.random.default_rng(seed), b, text, 250)
    m.generate(input_ids=torch.from_numpy(ids).cuda(), attention_mask=torch.from_numpy(mask).cuda(), max_new_tokens=16)
    torch.cuda.synchronize()
    return m._last_state


def time_replays(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


m = mk()
st_full = session(m, B, 0)
t_full = time_replays(lambda: st_full["graph"].replay())
print(f"one graph, batch {B}: {t_full:.3f} ms per frame", flush=True)
m._session = None   # st_full keeps its own buffers alive; the halves get new sessions on the SAME weights
st_a = session(m, B // 2, 1)
m._session = None
st_b = session(m, B // 2, 2)
t_half = time_replays(lambda: st_a["graph"].replay())
print(f"one graph, batch {B // 2}: {t_half:.3f} ms per frame", flush=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
main = torch.cuda.current_stream()


def both():
    s1.wait_stream(main); s2.wait_stream(main)
    with torch.cuda.stream(s1):
        st_a["graph"].replay()
    with torch.cuda.stream(s2):
        st_b["graph"].replay()
    main.wait_stream(s1); main.wait_stream(s2)


t_two = time_replays(both)
print(f"two graphs of batch {B // 2} on two streams: {t_two:.3f} ms per frame of {B} rows ({t_full / t_two:.2f}x the single graph)", flush=True)
# offset start: the second graph launched half a layer later cannot be expressed with whole-graph replays; this is the lower bound
