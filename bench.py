#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native MOSS-TTSD hot path (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[4], "C5": the end-to-end LM + codec throughput configuration, batch 256 per GPU):
  v0.5-shaped LM (Qwen3-1.7B dims, 8 codebooks), random init, bf16; 256 dialogue scripts per GPU per step; prompts of
  a ragged text part (uniform 64..512 rows) + 250 audio rows (two 10 s speaker prompts), delay-shifted and
  left-padded as process_inputs does; greedy; 375 new frames (30 s) per script; then XY_Tokenizer.decode (shipped
  config, random init) to 24 kHz waveforms. One "step" = one such batch end to end (2048 scripts = 8 steps on one GPU,
  one step each on 8). metric = audio-seconds generated per wall-second, whole job (all ranks).

One rank per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); requests are sharded, no data-path collective; the
timed region is bracketed by barrier + synchronize and the MAX over ranks is reported.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

SHAPE = dict(hidden_size=2048, intermediate_size=6144, num_hidden_layers=28, num_attention_heads=16,
             num_key_value_heads=8, head_dim=128, rms_norm_eps=1e-6, rope_theta=1e6, vocab_size=152697,
             speech_vocab_size=1025, channels=8, speech_token_range=[151665, 152689])
BATCH = int(os.environ.get("MTTS_BENCH_BATCH", 256))
TEXT_ROWS, AUDIO_ROWS, NEW_FRAMES = 200, 250, int(os.environ.get("MTTS_BENCH_FRAMES", 375))
FRAME_S = 0.08
METRIC = "audio-sec generated/sec (RTF), LM decode + codec decode, whole job"
UNIT = "audio_s/s"
TEXT_MIN, TEXT_MAX = 64, 512
WORKLOAD = (f"C5: end-to-end LM + codec, v0.5-shaped LM bf16 greedy decode, batch {BATCH} scripts/GPU, prompts of "
            f"{TEXT_MIN}..{TEXT_MAX} text rows (ragged, left-padded) + {AUDIO_ROWS} audio rows (2x10 s), {NEW_FRAMES} new frames "
            f"({NEW_FRAMES * FRAME_S:.0f} s) per script, then XY_Tokenizer.decode to 24 kHz")


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
    return sh, np.ones((B, n + C - 1), dtype=np.float64)


def make_ragged_prompts(rng, B, audio_rows):
    """C5 prompts: per script a text part of TEXT_MIN..TEXT_MAX rows + `audio_rows` prompt-audio rows, delay-shifted
    (+7 rows) and LEFT-padded to the longest script with mask 0 (generation_utils.process_inputs / rpadding)."""
    lo, hi, C = 151665, 152689, 8
    lens = rng.integers(TEXT_MIN, TEXT_MAX + 1, B)
    T = int(lens.max()) + audio_rows + C - 1
    ids = np.full((B, T, C), 1024, dtype=np.int64)
    ids[:, :, 0] = 151643
    mask = np.zeros((B, T), dtype=np.float64)
    for b in range(B):
        one, _ = make_prompt(rng, 1, int(lens[b]), audio_rows)
        n = one.shape[1]
        ids[b, T - n:] = one[0]
        mask[b, T - n:] = 1.0
    return ids, mask, lens


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_baseline_sample(threads):
    """The reference's algorithm (pinned oracle port: KV-cached fp32 eager LM + `_sample` + codec decode) on the host
    cores, on a bounded sample of the same workload: batch 1, 24 prompt rows, 12 new frames, full-size models."""
    from oracle import lm_oracle
    from oracle.codec_oracle import CodecOracle
    from oracle.codec_weights import full_codec_params, make_codec_weights
    torch.set_num_threads(threads)
    sample = "batch 1, 24 prompt rows (8 text + 16 audio), 12 new frames (0.96 s), full-size fp32 LM + full codec decode"
    sd = lm_oracle.random_weights_fast(SHAPE, 0)
    lm = lm_oracle.OracleCachedLM(SHAPE, sd, torch.float32)
    codec = CodecOracle(full_codec_params(), make_codec_weights(full_codec_params(), 5))
    ids, mask = make_prompt(np.random.default_rng(1), 1, 8, 16)
    ids, mask = torch.from_numpy(ids), torch.from_numpy(mask)
    new = 12

    def one():
        t0 = time.perf_counter()
        seq = lm.generate(ids, mask, max_length=ids.shape[1] + new, speech_range=SHAPE["speech_token_range"])
        out = seq[:, ids.shape[1] - 7:]
        n = out.shape[1] - 7
        speech = torch.stack([out[:, j:n + j, j] for j in range(8)], -1)
        speech[..., 0] -= 151665
        with torch.no_grad():
            wav = codec.decode([speech[0].clamp(0, 1023).permute(1, 0)])
        dt = time.perf_counter() - t0
        return n * FRAME_S, dt, wav[0].shape[0]

    return one, sample


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    one, sample = cpu_baseline_sample(threads)
    for _ in range(max(0, min(args.warmup, 1))):
        one()
    audio = secs = 0.0
    steps = max(1, min(args.steps, 5))
    for _ in range(steps):
        a, t, _ = one()
        audio += a
        secs += t
    v = audio / secs
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * secs / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (random-init weights, random prompt grids)",
        "config": {"workload": WORKLOAD, "reference_arm": "CPU port of the reference algorithm (oracle/, pinned to the "
                   "reference's outputs); the reference itself is Python/PyTorch and is not installable on the GPU box"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=dev)

    from moss_ttsd_b200 import _lib
    from moss_ttsd_b200.generation_utils import undelay, find_max_valid_positions
    from moss_ttsd_b200.modeling_asteroid import AsteroidTTSConfig, AsteroidTTSInstruct
    from moss_ttsd_b200.xy_tokenizer.model import XY_Tokenizer

    L = _lib.load()
    cfg = AsteroidTTSConfig(**SHAPE, eos_token_id=152694, pad_token_id=151643, tie_word_embeddings=False)
    model = AsteroidTTSInstruct(cfg, device=dev)
    model.init_random_weights(seed=0, tied=False, speech_only_head0=True)
    model.generation_config.eos_token_id = 152694
    import yaml
    with open(os.path.join(ROOT, "moss-ttsd_b200", "xy_tokenizer", "xy_tokenizer_config.yaml")) as f:
        spt = XY_Tokenizer(yaml.safe_load(f)["generator_params"])
    spt.init_random_weights(seed=5, device=dev)
    rng = np.random.default_rng(1000 + rank)
    ids_np, mask_np, text_lens = make_ragged_prompts(rng, BATCH, AUDIO_ROWS)
    ids_host = torch.from_numpy(ids_np).pin_memory()
    mask_host = torch.from_numpy(mask_np).pin_memory()
    ids_dev, mask_dev = ids_host.to(dev), mask_host.to(dev)
    T = ids_np.shape[1]
    start = T - 7
    wav_host = torch.empty((BATCH, NEW_FRAMES * 1920), dtype=torch.float32).pin_memory()

    phase_ms = {"generate": 0.0, "codec": 0.0}

    def hot_path(ids, mask):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        out = model.generate(input_ids=ids, attention_mask=mask, max_new_tokens=NEW_FRAMES, do_sample=False)
        ev[1].record()
        speech = undelay(out[:, start:])
        ends = (find_max_valid_positions(speech) + 1)
        n = speech.shape[1]
        wavs = spt.decode([speech[i].permute(1, 0) for i in range(BATCH)], overlap_seconds=10)["syn_wav_list"]
        ev[2].record()
        hot_path.events.append(ev)
        return wavs, n, ends

    hot_path.events = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(k, fn):
        barrier()
        l0, g0 = L.mtts_launch_count(), model.engine.graph_replayed_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        frames = 0
        for _ in range(k):
            frames += fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = (L.mtts_launch_count() - l0) + (model.engine.graph_replayed_launches - g0)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            f = torch.tensor([float(frames)], device=dev)
            dist.all_reduce(f, op=dist.ReduceOp.SUM)
            frames = float(f.item())
        return ms, frames, launches

    def step_resident():
        wavs, n, ends = hot_path(ids_dev, mask_dev)
        return BATCH * n

    def step_e2e():
        ids = ids_host.to(dev, non_blocking=True)
        mask = mask_host.to(dev, non_blocking=True)
        wavs, n, ends = hot_path(ids, mask)
        for i, w in enumerate(wavs):
            wav_host[i, :w.numel()].copy_(w, non_blocking=True)
        torch.cuda.synchronize()
        return BATCH * n

    if os.environ.get("MTTS_BENCH_LAUNCHLIST", "0") == "1":
        # launch-list mode for `ncu --metrics gpu__time_duration.sum`: the hot path exactly as the timed region runs it
        # (prefill, every decode step, un-delay, codec decode), --steps times, and nothing else (no sweeps, no CPU leg)
        for _ in range(args.warmup):
            step_resident()
        ms, frames, launches = timed(args.steps, step_resident)
        if rank == 0:
            print(json.dumps({"mode": "launch list (not a bench value)", "steps": args.steps, "warmup": args.warmup,
                              "batch_per_gpu": BATCH, "new_frames": NEW_FRAMES, "gpu_launches": int(launches)}))
        if world > 1:
            dist.destroy_process_group()
        return
    for _ in range(args.warmup):
        step_resident()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    hot_path.events = []
    ms, frames, launches = timed(args.steps, step_resident)
    gen_ms = sum(e[0].elapsed_time(e[1]) for e in hot_path.events) / max(1, len(hot_path.events))
    codec_ms = sum(e[1].elapsed_time(e[2]) for e in hot_path.events) / max(1, len(hot_path.events))
    clk = clocks.stop() if rank == 0 else None
    value = frames * FRAME_S / (ms / 1e3)
    step_e2e()
    ms_e, frames_e, _ = timed(args.steps, step_e2e)
    e2e = frames_e * FRAME_S / (ms_e / 1e3)

    # ---- roofline leg: the dominant kernel of the step (dense-projection GEMM of the decode step, M = batch)
    roof = roof_other = None
    if rank == 0:
        from moss_ttsd_b200 import ops
        w = model._w
        x = torch.randn(BATCH, SHAPE["hidden_size"], device=dev).to(torch.bfloat16)
        hq = torch.randn(BATCH, SHAPE["num_attention_heads"] * SHAPE["head_dim"], device=dev).to(torch.bfloat16)
        hi = torch.randn(BATCH, SHAPE["intermediate_size"], device=dev).to(torch.bfloat16)
        o_qkv = torch.empty(BATCH, w.layers[0]["wqkv"].shape[0], device=dev, dtype=torch.bfloat16)
        o_h = torch.empty(BATCH, SHAPE["hidden_size"], device=dev, dtype=torch.bfloat16)
        o_i = torch.empty(BATCH, SHAPE["intermediate_size"], device=dev, dtype=torch.bfloat16)
        gws = model.engine._gemm_ws(BATCH)

        def gemm_sweep():
            for lw in w.layers:
                ops.gemm(x, lw["wqkv"], out=o_qkv, workspace=gws)
                ops.gemm(hq, lw["wo"], out=o_h, residual=o_h, workspace=gws)
                ops.gemm(x, lw["wgu"], out=o_i, swiglu=True, workspace=gws)
                ops.gemm(hi, lw["wd"], out=o_h, residual=o_h, workspace=gws)

        nl = 4 * len(w.layers)
        algo = sum(lw[k].numel() * 2 for lw in w.layers for k in ("wqkv", "wo", "wgu", "wd"))
        algo += len(w.layers) * BATCH * 2 * (2 * x.shape[1] + hq.shape[1] + hi.shape[1] + o_qkv.shape[1] + 2 * o_h.shape[1] + o_i.shape[1])
        gemm_sweep()
        torch.cuda.synchronize()
        # replayed from a CUDA graph, exactly as the decode step issues these launches (PDL edges included)
        sweep_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(sweep_graph):
            gemm_sweep()
        for _ in range(3):
            sweep_graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            sweep_graph.replay()  # 3.1 GB of weights per sweep >> 126 MB L2: every launch streams from HBM
        e1.record()
        torch.cuda.synchronize()
        per_launch_ms = e0.elapsed_time(e1) / (reps * nl)
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        ach = (algo / nl) / (per_launch_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "gemm_traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        gemm_roof = {"kernel": f"gemm_tc_kernel<bf16> (decode-step dense projections, M = {BATCH} rows, cluster split-K)", "bound": "hbm",
                     "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                     "algorithmic_bytes_per_launch": algo / nl, "avg_launch_us": per_launch_ms * 1e3, "traffic": traffic,
                     "launches_per_decode_step": nl + 1, "us_per_decode_step": per_launch_ms * 1e3 * nl}
        # ---- decode attention over the KV cache the timed batches left behind (K+V of every cached token of every row)
        sess = model._session
        cache = sess["cache"]
        ctx_rows = torch.from_numpy((text_lens + AUDIO_ROWS + 7 + NEW_FRAMES // 2).astype(np.int32)).to(dev)
        q = torch.randn(BATCH, SHAPE["num_attention_heads"] * SHAPE["head_dim"], device=dev).to(torch.bfloat16)
        ao = torch.empty_like(q)

        def attn_sweep():
            for l in range(len(w.layers)):
                _lib.check(L.mtts_gqa_attention(q.data_ptr(), cache.k[l].data_ptr(), cache.v[l].data_ptr(), _lib.ptr(cache.block_table),
                                                cache.max_pages, cache.page_size, None, None, None, ctx_rows.data_ptr(), ao.data_ptr(),
                                                BATCH, 1, SHAPE["num_attention_heads"], SHAPE["num_key_value_heads"],
                                                SHAPE["head_dim"], 1, None, 0, _lib.stream_ptr()))

        attn_sweep()
        torch.cuda.synchronize()
        ag = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ag):
            attn_sweep()
        for _ in range(3):
            ag.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            ag.replay()  # 28 layers x (K + V of ~700 rows x 256 sequences) = 21 GB per sweep >> L2
        e1.record()
        torch.cuda.synchronize()
        a_ms = e0.elapsed_time(e1) / (reps * len(w.layers))
        a_bytes = float((ctx_rows.sum().item() + BATCH)) * 2 * SHAPE["num_key_value_heads"] * SHAPE["head_dim"] * 2 + 2 * q.numel() * 2
        a_ach = a_bytes / (a_ms * 1e-3) / 1e9
        a_traffic = None
        try:  # dram__bytes_read + dram__bytes_write of one launch at batch 256 / context 720 (uniform), ncu --set full
            with open(os.path.join(ROOT, "profiles", "attn_traffic.json")) as f:
                a_traffic = json.load(f).get("dram_bytes_per_launch") if BATCH == 256 else None
        except Exception:
            pass
        attn_roof = {"kernel": "gqa_decode_tc_kernel<G=2> (decode attention over the KV cache, one query row per sequence)",
                     "bound": "hbm", "achieved": a_ach, "peak": peak, "unit": "GB/s", "frac": a_ach / peak,
                     "peak_source": gemm_roof["peak_source"], "algorithmic_bytes_per_launch": a_bytes,
                     "avg_launch_us": a_ms * 1e3, "traffic": a_traffic, "launches_per_decode_step": len(w.layers),
                     "us_per_decode_step": a_ms * 1e3 * len(w.layers), "mean_context_rows": float(ctx_rows.float().mean().item()),
                     "note": "peak is the measured COPY bandwidth (read + write); a read-only stream can exceed it "
                             "(the weight stream of mtts_decode_mega reads at 7.1-7.7 TB/s)"}
        # the dominant kernel of the step is the one with the larger share of a decode step
        roof, roof_other = (attn_roof, gemm_roof) if attn_roof["us_per_decode_step"] >= gemm_roof["us_per_decode_step"] else (gemm_roof, attn_roof)

    cpu = None
    if rank == 0 and world == 1 and os.environ.get("MTTS_BENCH_SKIP_CPU", "0") != "1":
        threads = os.cpu_count() or 1
        one, sample = cpu_baseline_sample(threads)
        a, t, _ = one()
        cpu = {"value": a / t, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}

    # ---- p50 decode-step latency at batch 1 (the second half of BASELINE.json's metric): one graph replay = one frame
    lat = None
    if rank == 0:
        t_dec = model._last_timing
        b64_ms = t_dec[0][1].elapsed_time(t_dec[0][2]) / max(1, t_dec[1] - 1)
        model.generate(input_ids=ids_dev[:1].contiguous(), attention_mask=mask_dev[:1].contiguous(), max_new_tokens=32,
                       do_sample=False)
        st = model._last_state
        n_rep = 128
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_rep + 1)]
        for _ in range(8):
            st["graph"].replay()
        evs[0].record()
        for i in range(n_rep):
            st["graph"].replay()
            evs[i + 1].record()
        torch.cuda.synchronize()
        t = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(n_rep))
        w = model._w
        ctx = T + 32 + 8 + n_rep // 2
        streamed = w.heads.numel() * 2 + sum(lw[k].numel() * 2 for lw in w.layers for k in ("wqkv", "wo", "wgu", "wd")) + \
            ctx * 2 * len(w.layers) * SHAPE["num_key_value_heads"] * SHAPE["head_dim"] * 2
        peak = float((roof or {}).get("peak", 6650.0))
        lat = {"batch1_p50_ms": t[n_rep // 2], "batch1_p90_ms": t[int(n_rep * 0.9)], "batch1_context_rows": ctx,
               "batch1_bytes_per_step": streamed, "batch1_hbm_frac": streamed / (t[n_rep // 2] * 1e-3) / 1e9 / peak,
               "batch1_path": "persistent single-kernel step (mtts_decode_mega)" if st.get("mega") else "kernel chain",
               f"batch{BATCH}_avg_ms": b64_ms}

    if rank == 0:
        h2d = ids_host.numel() * 8 + mask_host.numel() * 8
        d2h = BATCH * NEW_FRAMES * 1920 * 4
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic (random-init weights, random prompt grids)",
            "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "prompt_rows_padded": T,
                       "prompt_rows_mean": float(text_lens.mean()) + AUDIO_ROWS + 7, "new_frames": NEW_FRAMES,
                       "kv_cache": "contiguous", "sampling": "greedy", "codec": "fp32 storage, TF32 tensor-core GEMMs",
                       "l2": "no flush needed: 3.5 GB weights + ~20 GB KV per decode step exceed the 126 MB L2"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e / args.steps},
            "phases_ms_per_step": {"lm_generate": gen_ms, "codec_decode": codec_ms},
            "decode_step_latency": lat,
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "roofline_other": roof_other, "cpu_baseline": cpu,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
