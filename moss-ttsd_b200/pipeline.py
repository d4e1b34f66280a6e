"""Stage overlap for batched generation: the codec decode of batch i runs on a second CUDA stream
while the LM decodes batch i+1.

The reference runs the two stages back to back (generation_utils.py:406-467: `model.generate`, then un-delay, then one
`spt.decode` per sample); its only streaming hook is `streamer.put(ch0)` (modeling_asteroid.py:161-162). The stages
use complementary resources — the decode steps are bound by HBM (KV cache + weights once per step), the codec by the
tensor pipe — so on one GPU they overlap instead of queueing. `CodecStage.submit` is called right after `generate`
returned for a batch (the LM stream is drained at that point, the host knows every row's length); it enqueues the
un-delay + `XY_Tokenizer.decode` launches on the codec stream and returns a job handle; the caller goes on to the next
batch's `generate` and collects the job later.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .generation_utils import find_max_valid_positions, undelay


class CodecJob:
    def __init__(self, stream, index=None):
        self.stream = stream
        self.index = index
        self.wavs: List[Optional[torch.Tensor]] = []
        self.ends: List[int] = []
        self.frames = 0
        self.done = torch.cuda.Event()
        self.error: Optional[BaseException] = None
        self.host: Optional[torch.Tensor] = None

    def wait(self):
        """Block the host until the job's kernels (and its device->host copy, if any) have finished."""
        self.done.synchronize()
        return self


class CodecStage:
    def __init__(self, spt, device=None, overlap: bool = True):
        self.spt = spt
        self.device = torch.device(device) if device is not None else spt.device
        self.stream = torch.cuda.Stream(device=self.device) if overlap else None

    def submit(self, outputs: torch.Tensor, start: int, host_out: Optional[torch.Tensor] = None, index=None,
               group_equal_lengths: bool = True) -> CodecJob:
        """outputs (B, L, 8): what `generate` returned; rows from `start` on are the generated part (start = T - 7).
        Un-delays, finds each row's last valid frame and decodes the rows with the codec on the codec stream.
        host_out: optional pinned (B, n) fp32 buffer; row b receives waveform b (device->host inside the job)."""
        cur = torch.cuda.current_stream(self.device)
        s = self.stream if self.stream is not None else cur
        job = CodecJob(s, index)
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(s):
            s.wait_event(ready)
            outputs.record_stream(s)
            speech = undelay(outputs[:, start:])
            ends = (find_max_valid_positions(speech) + 1).cpu().tolist()      # the LM stream is already drained
            job.ends = ends
            B = speech.shape[0]
            job.wavs = [None] * B
            groups = {}
            for i, e in enumerate(ends):
                if e > 0:
                    groups.setdefault(e if group_equal_lengths else (e, i), []).append(i)
            for key, idxs in groups.items():
                e = key if group_equal_lengths else key[0]
                wavs = self.spt.decode([speech[i, :e].permute(1, 0) for i in idxs], overlap_seconds=10)["syn_wav_list"]
                for i, w in zip(idxs, wavs):
                    job.wavs[i] = w
                    job.frames += e
                    if host_out is not None:
                        host_out[i, :w.numel()].copy_(w, non_blocking=True)
            job.host = host_out
            job.done.record(s)
        return job
