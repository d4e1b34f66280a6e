"""Request sharding (host logic) incl. a world_size-2 gloo run on CPU."""
import os
import subprocess
import sys
import textwrap

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
