"""Request sharding across the GPUs of one box. The path is embarrassingly data-parallel: every
dialogue script is an independent unit from prompt build to waveform, so ranks own disjoint sets of requests, keep a
full weight replica and never exchange tensors. The only cross-rank traffic is the host-side gather of results
(`gather_results`, an object all-gather over whatever backend `torch.distributed` was initialised with)."""
from __future__ import annotations

from typing import List, Sequence


def shard_requests(lengths: Sequence[int], world_size: int, rank: int, policy: str = "lpt") -> List[int]:
    """Indices of the requests rank `rank` owns.

    policy "round_robin": i % world_size == rank (the reference has no sharding; this is the obvious baseline).
    policy "lpt": longest-processing-time-first greedy bin packing on `lengths` (expected frames per script) so that the
    batch that runs until its LONGEST row finishes (modeling_asteroid.py:169) is balanced across ranks. Deterministic
    and identical on every rank (ties broken by index)."""
    n = len(lengths)
    if world_size <= 1:
        return list(range(n))
    if policy == "round_robin":
        return [i for i in range(n) if i % world_size == rank]
    order = sorted(range(n), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * world_size
    counts = [0] * world_size
    owner = [0] * n
    cap = (n + world_size - 1) // world_size  # keep per-rank batch sizes equal (weak scaling, fixed batch per GPU)
    for i in order:
        r = min((r for r in range(world_size) if counts[r] < cap), key=lambda r: (loads[r], r))
        owner[i] = r
        loads[r] += int(lengths[i])
        counts[r] += 1
    return [i for i in range(n) if owner[i] == rank]


def batches(indices: Sequence[int], batch_size: int) -> List[List[int]]:
    return [list(indices[i:i + batch_size]) for i in range(0, len(indices), batch_size)]


def length_bucketed_batches(indices: Sequence[int], lengths: Sequence[int], batch_size: int) -> List[List[int]]:
    """Batches of at most `batch_size` requests with similar expected lengths (the longest-row tail): a
    batch decodes until its LONGEST row has finished (modeling_asteroid.py:166-169), so its cost is max(length), not the
    mean. Sorting by decreasing expected length and cutting consecutive runs minimises the sum of the batch maxima over
    all partitions into batches of this size; ties are broken by index, so every rank computes the same batches."""
    order = sorted(indices, key=lambda i: (-int(lengths[i]), i))
    return batches(order, batch_size)


def decode_steps(lengths: Sequence[int], batch_list: Sequence[Sequence[int]]) -> int:
    """Decode steps a rank spends on `batch_list` when every batch runs until its longest row is done."""
    return sum(max(int(lengths[i]) for i in b) for b in batch_list if len(b))


def gather_results(local: dict, world_size: int) -> dict:
    """Merge {request index: result} dicts from all ranks on every rank (host-side; no GPU collective)."""
    if world_size <= 1:
        return dict(local)
    import torch.distributed as dist
    parts = [None] * world_size
    dist.all_gather_object(parts, local)
    merged = {}
    for p in parts:
        dup = set(merged) & set(p)
        if dup:
            raise RuntimeError(f"requests {sorted(dup)} were processed by more than one rank")
        merged.update(p)
    return merged
