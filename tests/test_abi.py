"""The C-ABI library loads and exports every symbol include/mtts.h declares (no GPU needed, no compute calls)."""
import ctypes
import os

from moss_ttsd_b200 import _lib


def test_header_symbols_are_bound_and_exported():
    syms = _lib.header_symbols()
    assert len(syms) >= 15
    assert set(syms) == set(_lib.SIGNATURES), (set(syms) ^ set(_lib.SIGNATURES))
    lib = _lib.load()
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.mtts_version() == 100
    assert isinstance(lib.mtts_last_error(), bytes)


def test_no_torch_types_in_the_abi():
    import re
    with open(_lib.HEADER_PATH) as f:
        code = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)  # declarations only, comments stripped
    assert "torch" not in code.lower() and "at::" not in code and "c10" not in code and "Tensor" not in code


def test_workspace_queries_work_without_a_device():
    lib = _lib.load()
    assert lib.mtts_gemm_workspace_bytes(1, 4096, 2048, 0) > 0
    assert lib.mtts_sample8_workspace_bytes(4, 8) > 4 * 8 * 2048 * 8
    assert lib.mtts_gqa_attention_workspace_bytes(4, 8, 2, 1, 8) > 65536


def test_product_never_imports_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "moss-ttsd_b200")
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dp, f)) as fh:
                    src = fh.read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dp, f)


def test_sampler_config_struct_size_matches_header():
    # 1 + 8*11 arrays of 4-byte fields + 3 trailing ints = 92 ints
    assert ctypes.sizeof(_lib.SamplerConfig) == 4 * (1 + 8 * 11 + 1 + 2)


def test_bench_launch_list_summary_tool_reads_the_committed_list():
    """profiles/r01_launches_bench.csv.gz (ncu launch list of the bench command) parses into complete decode steps whose
    dominant kernels are the ones bench.py reports as roofline / roofline_other."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "scripts"))
    import summarize_bench_launches as sbl
    rows = list(sbl.load(os.path.join(root, "profiles", "r01_launches_bench.csv.gz")))
    dec, nsteps = sbl.decode_steps(rows)
    assert nsteps >= 40 and len(dec) % nsteps == 0
    per_step = {}
    for name, us, _ in dec:
        per_step[sbl.short(name)] = per_step.get(sbl.short(name), 0.0) + us / nsteps
    top2 = sorted(per_step, key=per_step.get, reverse=True)[:2]
    assert any("gqa_decode_tc_kernel" in k for k in top2) and any("gemm_tc_kernel" in k for k in top2)


def test_bench_stdout_is_one_json_line_even_when_a_library_writes_to_fd1():
    """bench.py's contract: rank 0 prints ONE JSON line. Anything a native library writes to file descriptor 1 during the
    run (NCCL's version banner does, under torchrun) must land on stderr, the line itself on the real stdout."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import os, sys; sys.path.insert(0, %r); import bench; emit = bench._claim_stdout(); "
            "os.write(1, b'NCCL version 0.0.0\\n'); print('chatter'); emit({'metric': 'm', 'value': 1.5})" % root)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert json.loads(out.stdout) == {"metric": "m", "value": 1.5}
    assert out.stdout.count("\n") == 1
    assert "NCCL version 0.0.0" in out.stderr and "chatter" in out.stderr


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    """`bench.py --impl reference` launched the way the driver launches N > 1 (torchrun, one process per GPU): rank 0 alone
    times the CPU port and prints the line, the other ranks exit 0 without work; the line carries the contract's keys."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29533", os.path.join(root, "bench.py"), "--impl", "reference",
                          "--gpus", "2", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=900,
                         env=env, cwd=root)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["metric"].startswith("audio-sec") and d["unit"] == "audio_s/s"
