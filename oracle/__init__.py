"""TEST INFRASTRUCTURE ONLY — CPU restatements of the reference's algorithms for the hot path.

Nothing under `moss-ttsd_b200/` may import this package. Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` use it, and only as the checker or the timed CPU
baseline, never as the product path.
"""
