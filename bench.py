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


def run_reference(args, rank, world, emit):
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
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * secs / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (random-init weights, random prompt grids)",
        "config": {"workload": WORKLOAD, "reference_arm": "CPU port of the reference algorithm (oracle/, pinned to the "
                   "reference's outputs); the reference itself is Python/PyTorch and is not installable on the GPU box"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ------------------------------------------------------------------------------------------------ script list
def make_script_list(n_scripts, seed=1000):
    """The job: `n_scripts` dialogue scripts, identical on every rank (seeded). A script is a light descriptor — index,
    text rows, prompt-audio rows, RNG seed of its token ids — whose prompt grid is built by the rank that owns it."""
    rng = np.random.default_rng(seed)
    text = rng.integers(TEXT_MIN, TEXT_MAX + 1, n_scripts)
    return [dict(index=i, text_rows=int(text[i]), audio_rows=AUDIO_ROWS, seed=int(seed * 1000003 + i)) for i in range(n_scripts)]


def build_batch(scripts):
    """Prompt grids of one batch: delay-shifted (+7 rows), LEFT-padded to the longest script with mask 0
    (generation_utils.process_inputs / shifting_inputs / rpadding)."""
    C = 8
    B = len(scripts)
    T = max(s["text_rows"] + s["audio_rows"] for s in scripts) + C - 1
    ids = np.full((B, T, C), 1024, dtype=np.int64)
    ids[:, :, 0] = 151643
    mask = np.zeros((B, T), dtype=np.float64)
    for b, s in enumerate(scripts):
        one, _ = make_prompt(np.random.default_rng(s["seed"]), 1, s["text_rows"], s["audio_rows"])
        n = one.shape[1]
        ids[b, T - n:] = one[0]
        mask[b, T - n:] = 1.0
    return ids, mask


def nccl_env():
    """NCCL's communicator lines are kept (the driver checks the rank count from them): INFO/INIT goes to a per-process
    file unless the caller already asked for INFO or more itself; nccl_summary() echoes the lines to stderr and puts them
    into the JSON line; stdout stays one JSON line."""
    # absent, or preset to a level that carries no communicator lines (the GPU pool presets VERSION): raise it to INFO/INIT
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
        os.environ["NCCL_DEBUG"] = "INFO"
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        d = os.path.join(ROOT, "gpurun_out")
        try:
            os.makedirs(d, exist_ok=True)
        except OSError:
            d = "/tmp"
        os.environ.setdefault("NCCL_DEBUG_FILE", os.path.join(d, "nccl_%h_%p.log"))


def nccl_summary():
    """-> {"nranks": N seen in this process's NCCL log, "lines": the communicator-init lines}; lines are echoed to stderr."""
    path = os.environ.get("NCCL_DEBUG_FILE")
    if not path:
        return None
    import re
    import socket
    path = path.replace("%h", socket.gethostname()).replace("%p", str(os.getpid()))
    try:
        with open(path, errors="replace") as f:
            lines = [ln.strip() for ln in f if "nranks" in ln and ("comm" in ln or "Init" in ln)]
        seen = sorted({int(m.group(1)) for ln in lines for m in [re.search(r"nranks (\d+)", ln)] if m})
        for ln in lines[-4:]:
            print(ln, file=sys.stderr)
        return {"nranks_seen": seen, "lines": lines[-4:], "file": path}
    except Exception:   # a log that cannot be read must never cost the bench line
        return None


# ------------------------------------------------------------------------------------------------ main arm
def _claim_stdout():
    """stdout of this process carries the JSON line and nothing else: file descriptor 1 is pointed at stderr for the
    whole run (NCCL prints its version banner to fd 1 even with NCCL_DEBUG_FILE set; any library chatter follows it to
    stderr) and the line is written to the saved descriptor by emit()."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(saved, (json.dumps(obj) + "\n").encode())
    return emit


def main():
    emit = _claim_stdout()
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
        return run_reference(args, rank, world, emit)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ranks_seen = 1
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        nccl_env()
        dist.init_process_group("nccl", device_id=dev)
        one = torch.ones(1, device=dev)
        dist.all_reduce(one)            # every rank of the NCCL communicator contributes 1
        ranks_seen = one.item()

    from moss_ttsd_b200 import _lib, scheduler
    from moss_ttsd_b200.modeling_asteroid import AsteroidTTSConfig, AsteroidTTSInstruct
    from moss_ttsd_b200.pipeline import CodecStage
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
    overlap = os.environ.get("MTTS_BENCH_OVERLAP", "1") != "0"
    stage = CodecStage(spt, dev, overlap=overlap)

    # ---- the job: steps x BATCH x world scripts (2048 at the quoted configuration: 8 batches on one GPU, one batch each
    # on 8), built ONCE as a list identical on every rank, request-sharded with scheduler.shard_requests, batched per
    # rank with scheduler.length_bucketed_batches; per-script results are merged with scheduler.gather_results.
    scripts = make_script_list(args.steps * BATCH * world)
    est = [s["text_rows"] + s["audio_rows"] for s in scripts]
    mine = scheduler.shard_requests(est, world, rank, policy=os.environ.get("MTTS_BENCH_SHARD", "round_robin"))
    groups = scheduler.length_bucketed_batches(mine, est, BATCH)
    assert len(groups) == args.steps and all(len(g) == BATCH for g in groups), (len(groups), [len(g) for g in groups])
    batches = []
    for g in groups:
        ids_np, mask_np = build_batch([scripts[i] for i in g])
        ih, mh = torch.from_numpy(ids_np).pin_memory(), torch.from_numpy(mask_np).pin_memory()
        batches.append(dict(idx=g, ids_host=ih, mask_host=mh, ids_dev=ih.to(dev), mask_dev=mh.to(dev), T=ids_np.shape[1]))
    text_lens = np.array([scripts[i]["text_rows"] for i in mine])
    wav_host = torch.empty((BATCH, NEW_FRAMES * 1920), dtype=torch.float32).pin_memory()
    wav_host2 = torch.empty((BATCH, NEW_FRAMES * 1920), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lm_timings = []   # (events, steps) of every generate call: prefill / decode split of the timed job

    def run_job(bs, resident):
        """The hot path over a list of batches: generate (main stream) -> un-delay + codec decode (codec stream, overlapping
        the next batch's generate) [-> waveforms to pinned host memory], then the host-side merge of per-script results."""
        jobs, local_res, frames = [], {}, 0
        for k, b in enumerate(bs):
            if resident:
                ids, mask = b["ids_dev"], b["mask_dev"]
            else:
                ids = b["ids_host"].to(dev, non_blocking=True)
                mask = b["mask_host"].to(dev, non_blocking=True)
            out = model.generate(input_ids=ids, attention_mask=mask, max_new_tokens=NEW_FRAMES, do_sample=False)
            lm_timings.append(model._last_timing)
            host = None if resident else (wav_host if k % 2 == 0 else wav_host2)
            if len(jobs) >= 2:
                jobs[-2][0].wait()          # a pinned buffer / a batch of waveforms is reused two batches later
            jobs.append((stage.submit(out, b["T"] - 7, host_out=host), b))
        for job, b in jobs:
            job.wait()
            frames += job.frames
            for i, e in zip(b["idx"], job.ends):
                local_res[i] = (rank, int(e) * 1920)
            job.wavs = None
        merged = scheduler.gather_results(local_res, world)
        assert len(merged) == len({id(b["ids_host"]) for b in bs}) * BATCH * world, (len(merged), len(bs))
        return frames

    def timed(fn):
        barrier()
        l0, g0 = L.mtts_launch_count(), model.engine.graph_replayed_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        frames = fn()
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

    # warm-up: W batches (the first group holds the longest prompts, so the decode session is sized once)
    warm = [batches[i % len(batches)] for i in range(args.warmup)]
    if os.environ.get("MTTS_BENCH_LAUNCHLIST", "0") == "1":
        # launch-list mode for `ncu --metrics gpu__time_duration.sum`: the hot path exactly as the timed region runs it
        run_job(warm, True)
        ms, frames, launches = timed(lambda: run_job(batches, True))
        if rank == 0:
            emit({"mode": "launch list (not a bench value)", "steps": args.steps, "warmup": args.warmup,
                  "batch_per_gpu": BATCH, "new_frames": NEW_FRAMES, "gpu_launches": int(launches)})
        if world > 1:
            dist.destroy_process_group()
        return
    run_job(warm, True)
    # one serial batch with per-phase events (explains the overlapped number; not part of the timed region)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    b0 = batches[0]
    torch.cuda.synchronize()
    ev[0].record()
    out0 = model.generate(input_ids=b0["ids_dev"], attention_mask=b0["mask_dev"], max_new_tokens=NEW_FRAMES, do_sample=False)
    ev[1].record()
    serial_stage = CodecStage(spt, dev, overlap=False)
    serial_stage.submit(out0, b0["T"] - 7).wait()
    ev[2].record()
    torch.cuda.synchronize()
    gen_ms, codec_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    t_dec = model._last_timing
    prefill_ms = t_dec[0][0].elapsed_time(t_dec[0][1])
    bN_ms = t_dec[0][1].elapsed_time(t_dec[0][2]) / max(1, t_dec[1] - 1)
    del out0

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    del lm_timings[:]
    ms, frames, launches = timed(lambda: run_job(batches, True))
    job_prefill_ms = float(np.mean([t[0][0].elapsed_time(t[0][1]) for t in lm_timings]))
    job_decode_step_ms = float(np.mean([t[0][1].elapsed_time(t[0][2]) / max(1, t[1] - 1) for t in lm_timings]))
    clk = clocks.stop() if rank == 0 else None
    value = frames * FRAME_S / (ms / 1e3)
    e2e_steps = min(args.steps, int(os.environ.get("MTTS_BENCH_E2E_STEPS", 10)))
    run_job(batches[:1], False)
    ms_e, frames_e, _ = timed(lambda: run_job(batches[:e2e_steps], False))
    e2e = frames_e * FRAME_S / (ms_e / 1e3)

    extras = {}
    roof = roof_other = cpu = lat = None
    if rank == 0:
        roof, roof_other = roofline_legs(model, L, dev, text_lens)
        lat = batch1_latency(model, batches[0], roof, bN_ms)
        if os.environ.get("MTTS_BENCH_EXTRAS", "1") != "0":
            extras["rvq_encode"] = rvq_leg(spt, dev)
            if world == 1:
                extras["sampling"] = sampling_leg(model, batches[0], bN_ms)
                extras["ragged"] = ragged_leg(model, dev)
                free_sessions(model)
                extras["gpu_eager_baseline"] = gpu_eager_leg(dev, value)
        if world == 1 and os.environ.get("MTTS_BENCH_SKIP_CPU", "0") != "1":
            threads = os.cpu_count() or 1
            one, sample = cpu_baseline_sample(threads)
            a, t, _ = one()
            cpu = {"value": a / t, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}

    if rank == 0:
        h2d = batches[0]["ids_host"].numel() * 8 + batches[0]["mask_host"].numel() * 8
        d2h = BATCH * NEW_FRAMES * 1920 * 4
        emit({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic (random-init weights, random prompt grids)",
            "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "scripts_in_job": len(scripts),
                       "sharding": "scheduler.shard_requests (round robin) -> length_bucketed_batches -> gather_results",
                       "prompt_rows_padded_max": max(b["T"] for b in batches),
                       "prompt_rows_mean": float(text_lens.mean()) + AUDIO_ROWS + 7, "new_frames": NEW_FRAMES,
                       "kv_cache": "contiguous", "sampling": "greedy",
                       "codec": "fp32 residual stream, fp16-operand tcgen05 GEMMs + tcgen05 fp16 attention (XY_Tokenizer.decode_gemm = "
                                "'f16', the default: 49-51 dB against the TF32 path, which is 48 dB against the reference golden)",
                       "stage_overlap": "codec decode of batch i on a second stream under the LM decode of batch i+1" if overlap else "off",
                       "l2": "no flush needed: 3.5 GB weights + ~20 GB KV per decode step exceed the 126 MB L2"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e / e2e_steps, "steps": e2e_steps},
            "phases_ms_per_step": {"serial_lm_generate": gen_ms, "serial_prefill": prefill_ms, "serial_codec_decode": codec_ms,
                                   "serial_total": gen_ms + codec_ms, "overlapped_total": ms / args.steps,
                                   "note": "serial_* = batch 0 alone (the group with the longest prompts); job_* = mean over "
                                           "the batches of the timed job (codec of the previous batch running underneath)",
                                   "job_prefill_mean": job_prefill_ms, "job_decode_step_mean": job_decode_step_ms},
            "decode_step_latency": lat,
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "roofline_other": roof_other, "cpu_baseline": cpu,
            "extras": extras,
            "nccl": dict(nccl_summary() or {"note": "NCCL_DEBUG was preset by the caller: NCCL's own init lines are wherever that "
                                                    "setting sends them"},
                         NCCL_DEBUG=os.environ.get("NCCL_DEBUG"), world_size=world, ranks_counted_by_allreduce=int(ranks_seen),
                         nccl_version=".".join(str(v) for v in torch.cuda.nccl.version()))
            if world > 1 else None,
        })
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ roofline legs
def free_sessions(model):
    """Drop the decode sessions (KV pools, state, graphs) a previous leg left behind."""
    model._session = None
    model._last_state = None
    model._continuous = None
    torch.cuda.synchronize()
    torch.cuda.empty_cache()


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def _traffic(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return {}


def roofline_legs(model, L, dev, text_lens):
    """The two kernels that make up a decode step, each timed as a graph-replayed sweep over all 28 layers with CUDA
    events: the dense projections at M = BATCH rows, and the decode attention over a KV cache of the bench's contexts.
    -> (dominant kernel by share of the step, the other)."""
    from moss_ttsd_b200 import _lib, ops
    w = model._w
    peaks = _peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    psrc = "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"
    x = torch.randn(BATCH, SHAPE["hidden_size"], device=dev).to(torch.bfloat16)
    hq = torch.randn(BATCH, SHAPE["num_attention_heads"] * SHAPE["head_dim"], device=dev).to(torch.bfloat16)
    hi = torch.randn(BATCH, SHAPE["intermediate_size"], device=dev).to(torch.bfloat16)
    o_qkv = torch.empty(BATCH, w.layers[0]["wqkv"].shape[0], device=dev, dtype=torch.bfloat16)
    o_h = torch.empty(BATCH, SHAPE["hidden_size"], device=dev, dtype=torch.bfloat16)
    o_i = torch.empty(BATCH, SHAPE["intermediate_size"], device=dev, dtype=torch.bfloat16)
    gws = model.engine._gemm_ws(BATCH)

    eng = model.engine
    acts = eng._alloc_acts(BATCH)
    for t in acts.values():
        t.normal_(0, 1)
    pws = eng._splitk_ws(BATCH)
    H, eps = SHAPE["hidden_size"], SHAPE["rms_norm_eps"]

    def gemm_sweep():
        """The dense-projection chain of a decode step exactly as DecoderEngine._layers_splitk launches it at this batch
        (q/k/v, o_proj and down_proj as split-K partial GEMMs whose fp32 slices are summed by the residual + RMSNorm
        kernels; gate/up + SwiGLU on CTA-pair tiles) — everything but the attention kernel."""
        xx, xn, ao, h = acts["x"], acts["xn"], acts["ao"], acts["h"]
        for lw in w.layers:
            eng._splitk(xn, lw["wqkv"], pws)
            S = eng._splitk(ao, lw["wo"], pws)
            _lib.check(L.mtts_splitk_reduce_rmsnorm(pws.data_ptr(), S, BATCH, H, xx.data_ptr(), xx.stride(0), lw["ln2"].data_ptr(),
                                                    xn.data_ptr(), xn.stride(0), eps, _lib.stream_ptr()))
            ops.gemm(xn, lw["wgu"], out=h, swiglu=True, workspace=gws)
            S = eng._splitk(h, lw["wd"], pws)
            _lib.check(L.mtts_splitk_reduce_rmsnorm(pws.data_ptr(), S, BATCH, H, xx.data_ptr(), xx.stride(0), lw["ln1"].data_ptr(),
                                                    xn.data_ptr(), xn.stride(0), eps, _lib.stream_ptr()))

    def replay_ms(fn, n_launch, reps=10):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()          # > 3 GB per sweep >> 126 MB L2: every launch streams from HBM
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (reps * n_launch)

    nl = 4 * len(w.layers)
    algo = sum(lw[k].numel() * 2 for lw in w.layers for k in ("wqkv", "wo", "wgu", "wd"))
    flops = 2.0 * BATCH * algo / 2
    algo += len(w.layers) * BATCH * 2 * (2 * x.shape[1] + hq.shape[1] + hi.shape[1] + o_qkv.shape[1] + 2 * o_h.shape[1] + o_i.shape[1])
    per_launch_ms = replay_ms(gemm_sweep, nl)
    ach = (algo / nl) / (per_launch_ms * 1e-3) / 1e9
    tr = _traffic("gemm_traffic.json")
    tpeak = float(peaks.get("bf16_tflops", 1624.2))
    gemm_roof = {"kernel": f"decode-step dense projections at M = {BATCH} rows: 4 tcgen05 GEMMs per layer (q/k/v, o_proj, down_proj "
                           "as split-K partial tiles, gate/up + SwiGLU on cta_group::2 tiles) with the 2 split-K reducers "
                           "(residual + RMSNorm) counted in; avg_launch_us = chain time / 4 GEMMs", "bound": "hbm",
                 "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": psrc,
                 "algorithmic_bytes_per_launch": algo / nl, "avg_launch_us": per_launch_ms * 1e3,
                 "traffic": tr.get("dram_bytes_per_launch"), "xbar_l2_to_sm_bytes_per_launch": tr.get("xbar_bytes_per_launch"),
                 "traffic_source": tr.get("source"),
                 "tensor_tflops": flops / nl / (per_launch_ms * 1e-3) / 1e12, "tensor_frac_of_burst": flops / nl / (per_launch_ms * 1e-3) / 1e12 / tpeak,
                 "launches_per_decode_step": 6 * len(w.layers), "us_per_decode_step": per_launch_ms * 1e3 * nl}
    sess = model._session
    cache = sess["cache"]
    ctx_rows = torch.from_numpy((text_lens[:BATCH] + AUDIO_ROWS + 7 + NEW_FRAMES // 2).astype(np.int32)).to(dev)
    q = torch.randn(BATCH, SHAPE["num_attention_heads"] * SHAPE["head_dim"], device=dev).to(torch.bfloat16)
    ao = torch.empty_like(q)

    def attn_sweep():
        for l in range(len(w.layers)):
            _lib.check(L.mtts_gqa_attention(q.data_ptr(), cache.k[l].data_ptr(), cache.v[l].data_ptr(), _lib.ptr(cache.block_table),
                                            cache.max_pages, cache.page_size, None, None, None, ctx_rows.data_ptr(), ao.data_ptr(),
                                            BATCH, 1, SHAPE["num_attention_heads"], SHAPE["num_key_value_heads"],
                                            SHAPE["head_dim"], 1, None, 0, _lib.stream_ptr()))

    a_ms = replay_ms(attn_sweep, len(w.layers))
    a_bytes = float((ctx_rows.sum().item() + BATCH)) * 2 * SHAPE["num_key_value_heads"] * SHAPE["head_dim"] * 2 + 2 * q.numel() * 2
    a_ach = a_bytes / (a_ms * 1e-3) / 1e9
    atr = _traffic("attn_traffic.json")
    attn_roof = {"kernel": "gqa_decode_tc_kernel<G=2> (decode attention over the KV cache, one query row per sequence)",
                 "bound": "hbm", "achieved": a_ach, "peak": peak, "unit": "GB/s", "frac": a_ach / peak, "peak_source": psrc,
                 "algorithmic_bytes_per_launch": a_bytes, "avg_launch_us": a_ms * 1e3,
                 "traffic": atr.get("dram_bytes_per_launch") if BATCH == 256 else None,
                 "launches_per_decode_step": len(w.layers), "us_per_decode_step": a_ms * 1e3 * len(w.layers),
                 "mean_context_rows": float(ctx_rows.float().mean().item()),
                 "note": "peak is the measured COPY bandwidth (read + write); a read-only stream can exceed it "
                         "(the weight stream of mtts_decode_mega reads at 7.1-7.7 TB/s)"}
    return (attn_roof, gemm_roof) if attn_roof["us_per_decode_step"] >= gemm_roof["us_per_decode_step"] else (gemm_roof, attn_roof)


def batch1_latency(model, b0, roof, bN_ms):
    """p50 decode-step latency at batch 1 (the second half of BASELINE.json's metric): one graph replay = one frame."""
    model.generate(input_ids=b0["ids_dev"][:1].contiguous(), attention_mask=b0["mask_dev"][:1].contiguous(), max_new_tokens=32,
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
    model.engine.err.zero_()
    t = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(n_rep))
    w = model._w
    ctx = int(b0["mask_dev"][0].sum().item()) + 32 + 8 + n_rep // 2
    streamed = w.heads.numel() * 2 + sum(lw[k].numel() * 2 for lw in w.layers for k in ("wqkv", "wo", "wgu", "wd")) + \
        ctx * 2 * len(w.layers) * SHAPE["num_key_value_heads"] * SHAPE["head_dim"] * 2
    peak = float((roof or {}).get("peak", 6650.0))
    return {"batch1_p50_ms": t[n_rep // 2], "batch1_p90_ms": t[int(n_rep * 0.9)], "batch1_context_rows": ctx,
            "batch1_bytes_per_step": streamed, "batch1_hbm_frac": streamed / (t[n_rep // 2] * 1e-3) / 1e9 / peak,
            "batch1_path": "persistent single-kernel step (mtts_decode_mega)" if st.get("mega") else "kernel chain",
            f"batch{BATCH}_avg_ms": bN_ms}


def rvq_leg(spt, dev):
    """ResidualVQ search alone (BASELINE configs[1], one 30 s chunk of batch 32: 12 000 vectors x 512 dims, 8 x 1024
    codes): exact fp32 FMA arithmetic, so the bound is the fp32 CUDA-core peak (148 SMs x 128 FMA x 2 x SM clock)."""
    from moss_ttsd_b200 import ops
    N = 12000
    z = torch.randn(N, 512, device=dev) * 0.3
    cb, norms = spt.quantizer.codebooks, spt.quantizer.norms
    for _ in range(2):
        ops.rvq_encode(z, cb, norms, want_zq=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        ops.rvq_encode(z, cb, norms, want_zq=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = N * 8 * 2 * 512 * 1024
    peak = 148 * 128 * 2 * float(_peaks().get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
    return {"kernel": "rvq_encode_kernel (8 sequential nearest-code searches per vector, fp32 FMA, bit-exact formula)",
            "vectors": N, "ms": ms, "vectors_per_s": N / (ms * 1e-3), "bound": "fp32 FMA", "achieved": flops / (ms * 1e-3) / 1e12,
            "peak": peak, "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / peak,
            "algorithmic_flops": flops, "audio_s_per_s": N * 0.08 / (ms * 1e-3)}


def sampling_leg(model, b0, greedy_ms):
    """The same batch with per-channel sampling: temperature / top-k / top-p / repetition penalty on every
    channel, the delay-pattern state machine unchanged. LM only (the codec does not depend on how tokens were drawn)."""
    layers = [dict(repetition_penalty=1.1, temperature=0.9, top_k=50, top_p=0.95) for _ in range(8)]
    frames = 96
    kw = dict(input_ids=b0["ids_dev"], attention_mask=b0["mask_dev"], max_new_tokens=frames, do_samples=[True] * 8, layers=layers)
    model.generate(**kw)
    torch.cuda.synchronize()
    model.generate(**kw)
    t = model._last_timing
    torch.cuda.synchronize()
    step_ms = t[0][1].elapsed_time(t[0][2]) / max(1, t[1] - 1)
    return {"config": "do_samples all true; repetition_penalty 1.1, temperature 0.9, top_k 50, top_p 0.95 on every channel",
            "decode_step_ms": step_ms, "greedy_decode_step_ms": greedy_ms, "frames": frames,
            "decode_audio_s_per_s": BATCH * FRAME_S / (step_ms * 1e-3)}


def ragged_leg(model, dev):
    """The ragged variant of the workload: 2 x BATCH scripts whose target length is uniform in 10..60 s
    (a per-request length budget: channel 0 is forced to EOS there and the row winds down as after a sampled EOS).
    LM only, three schedules: arrival-order static batches that run to their longest row (the reference's scheme,
    modeling_asteroid.py:166-169), length-bucketed static batches, and continuous batching (finished rows' slots are
    refilled from the queue). audio seconds = frames actually generated before each row's EOS."""
    from moss_ttsd_b200 import scheduler
    rng = np.random.default_rng(77)
    n = 2 * BATCH
    scripts = make_script_list(n, seed=555)
    target = rng.integers(125, 751, n)                       # frames: 10 .. 60 s
    grids = []
    for s in scripts:
        one, _ = make_prompt(np.random.default_rng(s["seed"]), 1, s["text_rows"], s["audio_rows"])
        grids.append(torch.from_numpy(one[0]))
    audio_s = float(target.sum()) * FRAME_S
    max_new = 760

    def pack(idx, tgt):
        T = max(grids[i].shape[0] for i in idx)
        ids = torch.full((len(idx), T, 8), 1024, dtype=torch.int64)
        ids[:, :, 0] = 151643
        mask = torch.zeros(len(idx), T, dtype=torch.float64)
        for b, i in enumerate(idx):
            g = grids[i]
            ids[b, T - g.shape[0]:] = g
            mask[b, T - g.shape[0]:] = 1
        return ids.to(dev), mask.to(dev), [T - 7 + int(t) for t in tgt]

    def static(order):
        packed = [pack(order[k:k + BATCH], [target[i] for i in order[k:k + BATCH]]) for k in range(0, len(order), BATCH)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        steps = 0
        for ids, mask, eos_at in packed:
            model.generate(input_ids=ids, attention_mask=mask, max_new_tokens=max_new, do_sample=False, eos_at=eos_at)
            steps += model._last_timing[1]
        torch.cuda.synchronize()
        return time.perf_counter() - t0, steps

    res = {}
    free_sessions(model)
    ids, mask, _ = pack(list(range(BATCH)), [0] * BATCH)                # warm-up: session + graph of the budgeted variant
    model.generate(input_ids=ids, attention_mask=mask, max_new_tokens=max_new, do_sample=False, eos_at=[ids.shape[1] - 7 + 12] * BATCH)
    dt, steps = static(list(range(n)))
    res["static_arrival_order"] = {"s": dt, "decode_steps": steps, "audio_s_per_s": audio_s / dt}
    order = [i for g in scheduler.length_bucketed_batches(list(range(n)), [int(t) for t in target], BATCH) for i in g]
    dt, steps = static(order)
    res["static_length_bucketed"] = {"s": dt, "decode_steps": steps, "audio_s_per_s": audio_s / dt}
    free_sessions(model)
    kw = dict(max_new_tokens=max_new, max_batch=BATCH, do_sample=False,
              eos_at=[grids[i].shape[0] - 7 + int(target[i]) for i in range(n)])
    model.generate_continuous(grids[:BATCH + 8], max_new_tokens=max_new, max_batch=BATCH, do_sample=False,
                              eos_at=[grids[i].shape[0] - 7 + 20 for i in range(BATCH + 8)])     # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    outs = model.generate_continuous(grids, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    cd = model._continuous[1]
    got = sum(o.shape[0] - (g.shape[0] - 7) - 7 for o, g in zip(outs, grids)) * FRAME_S
    res["continuous"] = {"s": dt, "decode_steps": cd.steps_done, "audio_s_per_s": audio_s / dt,
                         "idle_slot_step_frac": cd.idle_slot_steps / max(1, cd.steps_done * BATCH), "audio_s_check": got}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model.generate_continuous(grids, queue_order="longest_first", **kw)
    torch.cuda.synchronize()
    dt2 = time.perf_counter() - t0
    cd = model._continuous[1]
    res["continuous_longest_first"] = {"s": dt2, "decode_steps": cd.steps_done, "audio_s_per_s": audio_s / dt2,
                                       "idle_slot_step_frac": cd.idle_slot_steps / max(1, cd.steps_done * BATCH)}
    res["continuous_vs_arrival_order"] = res["static_arrival_order"]["s"] / dt
    res["continuous_longest_first_vs_arrival_order"] = res["static_arrival_order"]["s"] / dt2
    res["scripts"] = n
    res["target"] = "uniform 125..750 frames (10..60 s) per script; LM only"
    free_sessions(model)
    return res


def gpu_eager_leg(dev, ours_value):
    """SURVEY §8d's "bar to beat": the reference's own PyTorch-eager arithmetic on the same B200 and the same C5 batch —
    the pinned oracle restatement (KV-cached bf16 LM + `_sample` loop + fp32 codec decode, `oracle/`) run with
    device='cuda': stock torch ops (cuBLAS GEMMs, eager softmax attention, torch.cat KV cache, ~60 sampler launches and
    two host syncs per step). One batch, fewer frames (decode cost per step grows only with context); reported beside
    the product's number, never used by it."""
    from oracle import lm_oracle
    from oracle.codec_oracle import CodecOracle
    from oracle.codec_weights import full_codec_params
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    torch.backends.cuda.matmul.allow_tf32 = False           # the reference's setting
    frames = int(os.environ.get("MTTS_BENCH_EAGER_FRAMES", 64))
    B = int(os.environ.get("MTTS_BENCH_EAGER_BATCH", BATCH))
    try:
        sd = lm_oracle.random_weights_fast(SHAPE, 0, dtype=torch.bfloat16)
        lm = lm_oracle.OracleCachedLM(SHAPE, sd, torch.bfloat16, device=dev)
        del sd
        scripts = make_script_list(B, seed=4242)
        ids_np, mask_np = build_batch(scripts)
        ids, mask = torch.from_numpy(ids_np).to(dev), torch.from_numpy(mask_np).to(dev)
        T = ids.shape[1]

        def gen(n_frames):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            seq = lm.generate(ids, mask, max_length=T + n_frames, speech_range=SHAPE["speech_token_range"])
            torch.cuda.synchronize()
            return seq, time.perf_counter() - t0

        gen(2)                                                   # cuBLAS / allocator warm-up
        _, t_short = gen(8)
        seq, t_long = gen(frames)
        step_s = (t_long - t_short) / (frames - 8)               # steady-state decode step incl. its host syncs
        prefill_s = max(t_short - (8 + 7) * step_s, 0.0)
        lm_s = prefill_s + (NEW_FRAMES + 7) * step_s             # extrapolated to the bench's 375 frames (+7 delay steps)
        del lm
        torch.cuda.empty_cache()
        # codec: the oracle's eager fp32 decode (allow_tf32 False, as the reference leaves it) on a slice of the batch
        from moss_ttsd_b200.xy_tokenizer.model import XY_Tokenizer
        gp = full_codec_params()
        gen_t = torch.Generator(device="cpu").manual_seed(5)
        spt = XY_Tokenizer(gp)
        sdc = {}
        for name, shape, fan in spt.param_shapes():
            t = torch.empty(shape, dtype=torch.float32).normal_(0.0, 1.0, generator=gen_t)
            t = t * fan ** -0.5 if fan > 0 else (1.0 + 0.1 * t if fan == 0 else (0.02 * t if fan == -1 else (0.1 * t if fan == -2 else 0.1 + 0.02 * t)))
            sdc[name] = t
        codec = CodecOracle(gp, sdc, device=dev)
        nb = 16
        codes = [torch.randint(0, 1024, (8, NEW_FRAMES), device=dev) for _ in range(nb)]
        with torch.no_grad():
            codec.decode(codes[:2])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            codec.decode(codes)
            torch.cuda.synchronize()
            codec_s = (time.perf_counter() - t0) * (B / nb)
        del codec
        torch.cuda.empty_cache()
        audio = B * NEW_FRAMES * FRAME_S
        v = audio / (lm_s + codec_s)
        return {"value": v, "unit": UNIT, "kind": "port of the reference's PyTorch-eager path on the same GPU (oracle/, device=cuda)",
                "batch": B, "prefill_s": prefill_s, "decode_step_ms": step_s * 1e3, "lm_s_per_batch": lm_s, "codec_s_per_batch": codec_s,
                "measured": f"prefill + {frames} of {NEW_FRAMES} frames timed, decode extrapolated per step; codec on {nb} of {B} items, scaled",
                "ours_over_eager": ours_value / v}
    except Exception as e:  # an out-of-memory eager path is a result too
        torch.cuda.empty_cache()
        return {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}


if __name__ == "__main__":
    main()
