"""Quick bandwidth / throughput probe of mtts_gemm on the decode and prefill shapes (CUDA events)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moss_ttsd_b200 import ops

def timeit(fn, iters=20, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rows = []
for M in (1, 16, 64, 256, 4096):
    for (N, K) in ((4096, 2048), (2048, 2048), (12288, 2048), (2048, 6144), (159872, 2048)):
        x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
        out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
        ms = timeit(lambda: ops.gemm(x, w, out=out), flush=flush)
        ms_t = timeit(lambda: torch.matmul(x, w.t(), out=out), flush=flush)
        gbs = (N * K * 2 + M * K * 2 + M * N * 2) / ms / 1e6
        tf = 2.0 * M * N * K / ms / 1e9
        rows.append(dict(M=M, N=N, K=K, ms=round(ms, 4), torch_ms=round(ms_t, 4), GBs=round(gbs, 1), TFs=round(tf, 1)))
        print(rows[-1], flush=True)
for M in (1500 * 4, 3000 * 4):
    for (N, K) in ((3072, 768), (768, 3072), (4096, 512), (512, 4096)):
        x = torch.randn(M, K, device="cuda")
        w = torch.randn(N, K, device="cuda") * 0.02
        out = torch.empty(M, N, device="cuda")
        ms = timeit(lambda: ops.gemm(x, w, out=out), flush=flush)
        torch.backends.cuda.matmul.allow_tf32 = True
        ms_t = timeit(lambda: torch.matmul(x, w.t(), out=out), flush=flush)
        rows.append(dict(M=M, N=N, K=K, dtype="tf32", ms=round(ms, 4), torch_tf32_ms=round(ms_t, 4),
                         TFs=round(2.0 * M * N * K / ms / 1e9, 1)))
        print(rows[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/bench_gemm.json", "w"), indent=1)
