"""Request sharding (host logic) incl. a world_size-2 gloo run on CPU."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from moss_ttsd_b200 import scheduler

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shards_partition_the_requests():
    lengths = [375, 100, 250, 375, 30, 60, 90, 375, 10, 200, 128]
    for policy in ("round_robin", "lpt"):
        for ws in (1, 2, 4, 8):
            owned = [scheduler.shard_requests(lengths, ws, r, policy) for r in range(ws)]
            flat = sorted(i for o in owned for i in o)
            assert flat == list(range(len(lengths)))
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
    # LPT balances the longest rows across ranks
    owned = [scheduler.shard_requests(lengths, 2, r, "lpt") for r in range(2)]
    loads = [sum(lengths[i] for i in o) for o in owned]
    assert abs(loads[0] - loads[1]) <= max(lengths)
    assert scheduler.batches(list(range(5)), 2) == [[0, 1], [2, 3], [4]]
    assert scheduler.shard_requests([], 4, 1) == []


def test_length_bucketed_batches_minimise_the_longest_row_tail():
    import random
    rng = random.Random(3)
    lengths = [rng.randint(125, 750) for _ in range(37)]  # 10..60 s scripts, in frames
    idx = list(range(len(lengths)))
    got = scheduler.length_bucketed_batches(idx, lengths, 8)
    assert sorted(i for b in got for i in b) == idx and all(1 <= len(b) <= 8 for b in got) and len(got) == 5
    arrival = scheduler.batches(idx, 8)
    assert scheduler.decode_steps(lengths, got) <= scheduler.decode_steps(lengths, arrival)
    # optimal among ALL partitions into batches of <= 2 (brute force over pairings of 6 requests)
    small = [300, 120, 710, 125, 690, 310]

    def pairings(items):
        if not items:
            yield []
            return
        a = items[0]
        for j in range(1, len(items)):
            for rest in pairings(items[1:j] + items[j + 1:]):
                yield [[a, items[j]]] + rest
    best = min(scheduler.decode_steps(small, p) for p in pairings(list(range(6))))
    assert scheduler.decode_steps(small, scheduler.length_bucketed_batches(range(6), small, 2)) == best
    assert scheduler.length_bucketed_batches([], [], 4) == []


def test_two_rank_gloo_gather(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import torch.distributed as dist
        from moss_ttsd_b200 import scheduler
        dist.init_process_group("gloo")
        r, ws = dist.get_rank(), dist.get_world_size()
        lengths = [5, 9, 2, 7, 7, 1, 3]
        mine = scheduler.shard_requests(lengths, ws, r)
        merged = scheduler.gather_results({{i: (r, lengths[i] * 1920) for i in mine}}, ws)
        assert sorted(merged) == list(range(len(lengths))), merged
        assert all(v[1] == lengths[i] * 1920 for i, v in merged.items())
        assert {{v[0] for v in merged.values()}} == {{0, 1}}
        dist.barrier()
        dist.destroy_process_group()
        print("rank", r, "ok")
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29653", str(script)],
                         capture_output=True, text=True, env=env, timeout=180)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_page_pool_never_loses_or_duplicates_a_page():
    """PagePool (continuous.py, the free list behind `block_table`): under a random admit / grow / retire schedule page 0
    is never handed out, no page is owned twice, a reservation can always be drawn, and everything returns to the pool."""
    from moss_ttsd_b200.continuous import PagePool
    with pytest.raises(ValueError):
        PagePool(1)
    rng = np.random.default_rng(3)
    pool = PagePool(97)
    owners = []            # per live request: [pages owned, pages still reserved]
    for _ in range(4000):
        op = rng.integers(0, 3)
        if op == 0:        # admit: reserve the worst case, take the pages the prompt needs now
            need = int(rng.integers(1, 12))
            before = pool.available
            if pool.reserve(need):
                first = int(rng.integers(0, need + 1))
                owners.append([pool.take_reserved(first), need - first])
                assert pool.available == before - need
            else:
                assert need > before and pool.available == before
        elif op == 1 and owners:   # grow a running request out of its reservation
            o = owners[int(rng.integers(0, len(owners)))]
            if o[1]:
                k = int(rng.integers(1, o[1] + 1))
                o[0] += pool.take_reserved(k)
                o[1] -= k
        elif op == 2 and owners:   # retire: pages and the unused reservation go back
            o = owners.pop(int(rng.integers(0, len(owners))))
            pool.release(o[0], unreserve=o[1])
        held = [p for o in owners for p in o[0]]
        assert 0 not in held and len(held) == len(set(held)) and all(0 < p < 97 for p in held)
        assert pool.reserved == sum(o[1] for o in owners)
        assert pool.available == 96 - len(held) - pool.reserved >= 0
    for o in owners:
        pool.release(o[0], unreserve=o[1])
    assert pool.available == 96 and pool.reserved == 0
