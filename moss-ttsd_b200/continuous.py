"""Continuous batching for the delay-pattern decoder.

The reference decodes a fixed batch until its LONGEST row has finished and feeds every finished row `[EOS, pad x7]`
meanwhile (modeling_asteroid.py:155-158,166-169). Here the batch is a set of SLOTS: when a row finishes, its sequence
is handed back, its KV pages return to the pool, and the next queued request is admitted into the slot while the other
rows keep decoding. Per request the result is what a solo `generate` call returns for it (same prompt rows, same
teacher-forced tail, same per-channel history for the repetition penalty — there are no left-pad rows at all, every
slot has its own prompt length).

How a request enters a running batch without a special first step: its first P-1 prompt rows are prefilled into the
slot's KV pages (no LM heads), the LAST prompt row is put in the slot's `tokens`, and the slot's control block
`row_ctl = {step0, P, max_length, eos_at}` tells the sampler / state-machine kernels to run this row at step
`step - step0`. The next decode step of the whole batch then computes that row's prompt-final logits exactly where the
reference's prefill does, and the row's step 0 (EOS mask, teacher-forced tail, ...) happens there.

KV memory: one pool of pages for all slots (`PagePool`, a host free list behind `block_table`); a request owns the
pages its current length needs plus a look-ahead, and is only admitted when the pool can also cover its worst-case
length (`max_length` rows), so a running row can never starve. Page 0 is a scratch page that idle slots point at.

The host never waits for the GPU inside the loop: every `sync_every` steps it enqueues an async copy of the per-row
`unfinished` / `finish_len` words and inspects only copies that have landed (at most two blocks late).
"""
from __future__ import annotations

import ctypes
from collections import deque
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from ._lib import check, stream_ptr
from .lm_engine import DecoderEngine, KVCache, SamplerSetup


class PagePool:
    """Host-side free list of KV pages. Page 0 is never handed out (scratch page of idle slots)."""

    def __init__(self, num_pages: int):
        if num_pages < 2:
            raise ValueError("the KV pool needs at least 2 pages (page 0 is the scratch page)")
        self.num_pages = num_pages
        self._free = list(range(num_pages - 1, 0, -1))
        self.reserved = 0          # pages promised to admitted requests but not yet handed out

    @property
    def available(self) -> int:
        return len(self._free) - self.reserved

    def reserve(self, n: int) -> bool:
        if n > self.available:
            return False
        self.reserved += n
        return True

    def take_reserved(self, n: int) -> List[int]:
        assert n <= self.reserved and n <= len(self._free)
        self.reserved -= n
        return [self._free.pop() for _ in range(n)]

    def release(self, pages: Sequence[int], unreserve: int = 0):
        for p in pages:
            assert 0 < p < self.num_pages
        self._free.extend(pages)
        self.reserved -= unreserve
        assert self.reserved >= 0 and len(self._free) <= self.num_pages - 1


@dataclass
class Request:
    """One script: the delay-shifted prompt grid (T, C) int64 whose last C-1 rows are the teacher-forced tail, the
    reference's `max_length` for it (rows incl. the prompt; generate(max_new_tokens=N) == T + N) and an optional length
    budget `eos_at` (sequence row from which channel 0 is forced to EOS)."""
    index: int
    grid: torch.Tensor
    max_length: int
    eos_at: int = 0


class ContinuousDecoder:
    def __init__(self, engine: DecoderEngine, slots: int, max_rows: int, sampler: SamplerSetup, speech_range, eos_token: int,
                 pool_pages: Optional[int] = None, page_size: int = 64, sync_every: int = 8, lookahead_pages: int = 1,
                 max_admit_rows: int = 262144):
        self.eng, self.B, self.max_rows = engine, slots, max_rows
        s = engine.s
        dev = engine.dev
        self.page_size = page_size
        self.max_pages = (max_rows + page_size - 1) // page_size + 1
        if pool_pages is None:
            pool_pages = slots * self.max_pages + 1
        self.pool = PagePool(pool_pages)
        cache = KVCache.__new__(KVCache)
        cache.page_size, cache.max_pages, cache.num_pages, cache.paged = page_size, self.max_pages, pool_pages, True
        shp = (s.num_hidden_layers, pool_pages, s.num_key_value_heads, page_size, s.head_dim)
        cache.k = torch.empty(shp, dtype=torch.bfloat16, device=dev)
        cache.v = torch.empty(shp, dtype=torch.bfloat16, device=dev)
        cache.k[:, 0].zero_()                  # the scratch page idle slots read: finite values
        cache.v[:, 0].zero_()
        cache.block_table = torch.zeros((slots, self.max_pages), dtype=torch.int32, device=dev)
        self.cache = cache
        self.bt = np.zeros((slots, self.max_pages), dtype=np.int32)
        self.sampler = sampler
        self.sync_every = max(1, int(sync_every))
        self.lookahead = max(1, int(lookahead_pages))
        self.max_admit_rows = max_admit_rows
        self.hist_len = 4096
        st = engine.make_decode_state(slots, cache, sampler, max_rows, tuple(speech_range), int(eos_token), True)
        st["row_ctl"] = torch.zeros((slots, 4), dtype=torch.int32, device=dev)
        st["hist"] = torch.zeros(self.hist_len, dtype=torch.int32, device=dev)
        st["hist_len"] = self.hist_len
        st["mega"] = None                      # the persistent small-batch kernel has no per-row control block
        st["unfinished"].zero_()
        self.st = st
        self.C = s.channels
        self.steps_done = 0                    # == the device step counter
        self.decode_steps = 0
        self.admitted = 0
        self.idle_slot_steps = 0               # slot-steps spent on rows that had nothing to do (occupancy diagnostics)

    # ------------------------------------------------------------------ slots
    def _reset(self):
        st = self.st
        st["unfinished"].zero_()
        st["needs"].fill_(-1)
        st["finish_len"].zero_()
        st["step"].zero_()
        st["positions"].zero_()
        st["row_ctl"].zero_()
        st["tokens"].zero_()
        self.steps_done = 0
        self.bt[:] = 0
        self.cache.block_table.zero_()

    def _pages_for(self, rows: int) -> int:
        return (rows + self.page_size - 1) // self.page_size

    def _admit(self, pairs):
        """pairs: [(slot, Request)]: pages, packed prefill of the first P-1 rows, per-slot state."""
        st, eng, C, dev = self.st, self.eng, self.C, self.eng.dev
        L = eng.L
        ids_parts, lens, slots = [], [], []
        for slot, rq in pairs:
            g = rq.grid
            T = g.shape[0]
            P = T - (C - 1)
            assert P >= 1 and rq.max_length + C <= self.max_rows, (P, rq.max_length, self.max_rows)
            info = self.slot[slot]
            have = self._pages_for(P + self.sync_every * 2 + 1) + self.lookahead - 1
            have = min(have, info["reserved"])
            pages = self.pool.take_reserved(have)
            info["reserved"] -= have
            info["pages"] = pages
            self.bt[slot, :] = 0
            self.bt[slot, :len(pages)] = pages
            if P > 1:
                ids_parts.append(g[:P - 1])
                lens.append(P - 1)
                slots.append(slot)
        self.cache.block_table.copy_(torch.from_numpy(self.bt))
        if lens:
            ids = torch.cat(ids_parts, 0).to(dev).contiguous()
            eng.prefill_packed(ids, np.asarray(lens), np.asarray(slots, dtype=np.int32), self.cache, logits=None)
        sl = torch.tensor([p[0] for p in pairs], dtype=torch.int64, device=dev)
        grids = [p[1].grid.to(dev) for p in pairs]
        Ps = [g.shape[0] - (C - 1) for g in grids]
        st["tokens"].index_copy_(0, sl, torch.stack([g[P - 1] for g, P in zip(grids, Ps)]))
        st["tf_tail"].index_copy_(0, sl, torch.stack([g[P:] for g, P in zip(grids, Ps)]))
        st["positions"].index_copy_(0, sl, torch.tensor([P - 1 for P in Ps], dtype=torch.int32, device=dev))
        st["unfinished"].index_fill_(0, sl, 1)
        st["needs"].index_fill_(0, sl, -1)
        st["finish_len"].index_fill_(0, sl, 0)
        ctl = torch.tensor([[self.steps_done, P, rq.max_length, rq.eos_at] for (_, rq), P in zip(pairs, Ps)],
                           dtype=torch.int32, device=dev)
        st["row_ctl"].index_copy_(0, sl, ctl)
        st["seen"].index_fill_(0, sl, 0)
        words = self.sampler.words_per_row
        for (slot, rq), g, P in zip(pairs, grids, Ps):
            st["sequences"][slot, :P].copy_(g[:P])
            check(L.mtts_sampler_init_history(g.data_ptr(), 1, P, g.stride(0) * g.shape[0], ctypes.byref(self.sampler.cfg),
                                              st["seen"].data_ptr() + slot * words * 4, stream_ptr()))
            info = self.slot[slot]
            info.update(req=rq, P=P, step0=self.steps_done, grid=g)
        self.admitted += len(pairs)

    def _grow(self, upto_step: int):
        """Every live slot owns pages for all rows it can reach by global step `upto_step`."""
        changed = False
        for slot, info in enumerate(self.slot):
            if info["req"] is None:
                continue
            rows = min(info["P"] + (upto_step - info["step0"]) + 1, info["req"].max_length + self.C)
            need = self._pages_for(rows) + self.lookahead - 1 - len(info["pages"])
            need = min(need, info["reserved"])
            if need > 0:
                new = self.pool.take_reserved(need)
                info["reserved"] -= need
                n0 = len(info["pages"])
                info["pages"].extend(new)
                self.bt[slot, n0:n0 + need] = new
                changed = True
        if changed:
            self.cache.block_table.copy_(torch.from_numpy(self.bt))

    # ------------------------------------------------------------------ the loop
    @torch.no_grad()
    def run(self, requests: Sequence[Request], seed: int = 0) -> dict:
        """-> {request.index: LongTensor (L_i, C)}: prompt rows [0, P_i) followed by the generated rows, exactly the rows
        a solo generate() returns for the request."""
        st, eng, B = self.st, self.eng, self.B
        dev = eng.dev
        self._reset()
        st["seed_dev"].copy_(torch.tensor([seed & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64))
        self.slot = [dict(req=None, pages=[], reserved=0, P=0, step0=0, grid=None) for _ in range(B)]
        queue = deque(requests)
        results = {}
        pending = []
        live = 0
        pin_u = [torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(3)]
        pin_f = [torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(3)]
        ring = 0

        def try_admit():
            nonlocal live
            pairs, rows = [], 0
            for slot in range(B):
                if not queue:
                    break
                if self.slot[slot]["req"] is not None or self.slot[slot].get("busy"):
                    continue
                rq = queue[0]
                worst = self._pages_for(rq.max_length + self.C) + self.lookahead - 1
                if rows + rq.grid.shape[0] > self.max_admit_rows and pairs:
                    break
                if not self.pool.reserve(worst):
                    break                                   # the pool cannot cover this request's worst case yet
                queue.popleft()
                self.slot[slot]["reserved"] = worst
                self.slot[slot]["busy"] = True
                pairs.append((slot, rq))
                rows += rq.grid.shape[0]
            if pairs:
                self._admit(pairs)
                live += len(pairs)
            return len(pairs)

        def retire(u_host, f_host, snap_step):
            nonlocal live
            gone = []
            for slot, info in enumerate(self.slot):
                rq = info["req"]
                if rq is None or info["step0"] >= snap_step or int(u_host[slot]) != 0:
                    continue
                n = int(f_host[slot])
                results[rq.index] = st["sequences"][slot, :n].clone()
                self.pool.release(info["pages"], unreserve=info["reserved"])
                info.update(req=None, pages=[], reserved=0, busy=False, grid=None)
                self.bt[slot, :] = 0
                gone.append(slot)
                live -= 1
            if gone:
                # the idle slot keeps being stepped (its [EOS, pad x7] row is written at its frozen position): point it at
                # the scratch page BEFORE its pages can be handed to another request
                st["positions"].index_fill_(0, torch.tensor(gone, dtype=torch.int64, device=dev), 0)
                self.cache.block_table.copy_(torch.from_numpy(self.bt))

        while queue or live:
            if len(pending) >= 2:
                pending[0][0].synchronize()                 # at most two blocks of replays queued ahead of the GPU
            while pending and pending[0][0].query():
                _, k, snap = pending.pop(0)
                retire(pin_u[k], pin_f[k], snap)
            if queue:
                try_admit()
            if not live:
                if queue and not pending:
                    raise MemoryError("KV pool too small for the next request even with every slot idle")
                if not queue and not pending:
                    break
                if pending:
                    pending[0][0].synchronize()
                continue
            n = self.sync_every
            self._grow(self.steps_done + n)
            for _ in range(n):
                eng.decode_step(st)
            self.steps_done += n
            self.decode_steps += n
            self.idle_slot_steps += n * (B - live)
            k = ring
            ring = (ring + 1) % 3
            pin_u[k].copy_(st["unfinished"], non_blocking=True)
            pin_f[k].copy_(st["finish_len"], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            pending.append((ev, k, self.steps_done))
        e = eng.err.cpu().tolist()
        if any(e):
            eng.err.zero_()
            raise RuntimeError(f"libmtts device-side error flags {e}")
        return results
