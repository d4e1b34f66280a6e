"""Where does a bench-shaped prefill spend its wall time? For each of the three length-bucketed batches of the C5 job:
device time between events around `engine.prefill`, host time until the call returns (launch side), the GPU idle time in
front of the first kernel (host preparation behind a synchronising `.cpu()`), and the caching allocator's cudaMalloc count.
Not a bench value."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    from moss_ttsd_b200 import scheduler
    from moss_ttsd_b200.modeling_asteroid import AsteroidTTSConfig, AsteroidTTSInstruct
    cfg = AsteroidTTSConfig(**bench.SHAPE, eos_token_id=152694, pad_token_id=151643, tie_word_embeddings=False)
    model = AsteroidTTSInstruct(cfg, device=dev)
    model.init_random_weights(seed=0, tied=False, speech_only_head0=True)
    model.generation_config.eos_token_id = 152694
    scripts = bench.make_script_list(3 * bench.BATCH)
    est = [s["text_rows"] + s["audio_rows"] for s in scripts]
    groups = scheduler.length_bucketed_batches(list(range(len(scripts))), est, bench.BATCH)
    batches = []
    for g in groups:
        ids_np, mask_np = bench.build_batch([scripts[i] for i in g])
        batches.append((torch.from_numpy(ids_np).to(dev), torch.from_numpy(mask_np).to(dev)))
    for rep in range(3):
        for k, (ids, mask) in enumerate(batches):
            out = model.generate(input_ids=ids, attention_mask=mask, max_new_tokens=4, do_sample=False)
            ev, steps = model._last_timing
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1])
            rows = int(mask.sum())
            st = torch.cuda.memory_stats()
            print(f"rep {rep} batch {k}: rows {rows}  prefill (events, as bench reports) {ms:8.1f} ms   "
                  f"cudaMalloc calls so far {st['num_device_alloc']}  reserved {st['reserved_bytes.all.current'] / 2**30:.1f} GiB", flush=True)
            del out
    # the engine call alone, host-timed in pieces
    eng = model.engine
    sess = model._session
    cache = sess["cache"]
    for k, (ids, mask) in enumerate(batches):
        P = ids.shape[1] - 0
        for rep in range(2):
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            t0 = time.perf_counter()
            e0.record()
            lens = (mask != 0).sum(1)
            lens_h = lens.cpu()
            t1 = time.perf_counter()
            e1.record()
            logits, _ = eng.prefill(ids, mask, cache)
            t2 = time.perf_counter()
            e2.record()
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            print(f"batch {k} rep {rep}: engine.prefill device {e1.elapsed_time(e2):8.1f} ms; host: launch side returns after "
                  f"{(t2 - t1) * 1e3:7.1f} ms, drained after {(t3 - t1) * 1e3:7.1f} ms", flush=True)


if __name__ == "__main__":
    main()
