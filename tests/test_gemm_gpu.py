"""tcgen05 GEMM (mtts_gemm) and the exact-fp32 CUDA-core GEMM (mtts_gemm_simt) against a torch fp64 product."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(x, w, bias=None, gelu=False, gamma=None, residual=None, out_dtype=torch.float32):
    v = x.double() @ w.double().t()
    if bias is not None:
        v = v + bias.double()
    if gelu:
        v = torch.nn.functional.gelu(v)
    if out_dtype == torch.bfloat16:
        v = v.to(torch.bfloat16).double()
    if gamma is not None:
        v = v * gamma.double()
    if residual is not None:
        v = residual.double() + v
    return v


SHAPES = [
    (1, 256, 128), (1, 2048, 2048), (3, 1000, 512), (16, 4096, 2048), (17, 384, 6144), (64, 2048, 6144),
    (100, 130, 200), (128, 768, 3072), (129, 2048, 2048), (256, 4096, 2048), (300, 1025, 2048), (257, 962, 512),
    (1000, 512, 560),
    # >= 592 weight tiles at 65..256 rows: the LM-heads route (CTA-pair kernel with one, partly out-of-bounds, row tile)
    (200, 76032, 128), (77, 75900, 192),
    # short K, many rows (the codec's ConvNeXt pw1 regime): many tiles per pair, last tiles partly out of bounds
    (19000, 1100, 448), (40000, 2048, 512),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_tc_bf16(M, N, K):
    from moss_ttsd_b200 import ops
    K = (K + 7) // 8 * 8
    g = torch.Generator(device="cuda").manual_seed(M * 7919 + N * 31 + K)
    x = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    out = ops.gemm(x, w, out_dtype=torch.float32)
    ref = _ref(x, w)
    err = (out.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * max(scale, 1.0), (err, scale)
    out2 = ops.gemm(x, w, out_dtype=torch.float32)
    assert torch.equal(out, out2), "split-K reduction must be deterministic"


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_tc_tf32(M, N, K):
    from moss_ttsd_b200 import ops
    K = (K + 3) // 4 * 4
    g = torch.Generator(device="cuda").manual_seed(M * 7919 + N * 31 + K + 1)
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.05
    out = ops.gemm(x, w)
    ref = _ref(x, w)
    err = (out.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 4e-3 * max(scale, 1.0), (err, scale)  # tf32: 10-bit mantissa operands


@pytest.mark.parametrize("M,N,K", [(5, 256, 512), (64, 2048, 2048), (200, 700, 768)])
def test_gemm_tc_epilogues_bf16(M, N, K):
    from moss_ttsd_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1234 + M)
    x = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.03).to(torch.bfloat16)
    res = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16)
    out = ops.gemm(x, w, residual=res)
    ref = _ref(x, w, residual=res, out_dtype=torch.bfloat16)
    err = (out.double() - ref).abs().max().item()
    assert err <= 0.05, err
    # SwiGLU: interleaved gate/up rows
    gate, up = w[: N // 2], w[N // 2:]
    wi = torch.stack([gate, up], dim=1).reshape(N, K).contiguous()
    h = ops.gemm(x, wi, swiglu=True)
    gq = (x.double() @ gate.double().t()).to(torch.bfloat16).float()
    uq = (x.double() @ up.double().t()).to(torch.bfloat16).float()
    href = (torch.nn.functional.silu(gq).to(torch.bfloat16).float() * uq).to(torch.bfloat16)
    assert h.shape == (M, N // 2)
    bad = ((h.float() - href.float()).abs() > 0.02 * (1 + href.float().abs())).float().mean().item()
    assert bad < 1e-3, bad


@pytest.mark.parametrize("M,N,K", [(7, 300, 512), (1500, 768, 3072)])
def test_gemm_tc_epilogues_f32(M, N, K):
    from moss_ttsd_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(99 + M)
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.03
    b = torch.randn(N, device="cuda", generator=g)
    gm = torch.rand(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    out = ops.gemm(x, w, bias=b, gelu=True)
    ref = _ref(x, w, bias=b, gelu=True)
    assert (out.double() - ref).abs().max().item() <= 5e-3 * max(ref.abs().max().item(), 1)
    out = ops.gemm(x, w, bias=b, gamma=gm, residual=res)
    ref = _ref(x, w, bias=b, gamma=gm, residual=res)
    assert (out.double() - ref).abs().max().item() <= 5e-3 * max(ref.abs().max().item(), 1)


@pytest.mark.parametrize("M,N,K", [(1, 64, 64), (33, 130, 77), (375, 512, 3072)])
def test_gemm_simt_exact_fp32(M, N, K):
    from moss_ttsd_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5 + M)
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.05
    b = torch.randn(N, device="cuda", generator=g)
    out = ops.gemm_simt(x, w, bias=b)
    ref = _ref(x, w, bias=b)
    assert (out.double() - ref).abs().max().item() <= 2e-5 * max(ref.abs().max().item(), 1)
    # channel-major activations (B, K, T) read in place
    B, T = 3, 50
    xc = torch.randn(B, K, T, device="cuda", generator=g)
    out = ops.gemm_simt(xc, w, x_layout=(T, K * T, 1, T), M=B * T)
    ref = _ref(xc.permute(0, 2, 1).reshape(B * T, K), w)
    assert (out.double() - ref).abs().max().item() <= 2e-5 * max(ref.abs().max().item(), 1)


def test_gemm_rejects_cpu_tensors():
    from moss_ttsd_b200 import ops, _lib
    with pytest.raises(_lib.MttsError):
        ops.gemm(torch.zeros(4, 64, dtype=torch.bfloat16), torch.zeros(8, 64, dtype=torch.bfloat16))


def test_3xtf32_gemm_is_fp32_accurate():
    """ops.gemm_exact (hi/lo TF32 split, K' = 3K, fp32 TMEM accumulators) against an fp64 product. The operand error is
    gone (2^-22); what remains is the tensor core's accumulation, which TRUNCATES each partial sum to fp32 (a bias of
    ~2^-25 |acc| per k-block of 8, growing linearly with K instead of with sqrt(K)): measured 1.05e-5 of the largest
    output at K = 768 on B200, against 9.8e-7 for the round-to-nearest fp32 FMA chain (gemm_simt) and 7.5e-4 for the
    plain TF32 product. The encode parity tests (tests/test_codec_gpu.py) show this is enough for bit-identical codes."""
    from moss_ttsd_b200 import ops
    torch.manual_seed(1)
    M, N, K = 300, 520, 768
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * K ** -0.5
    b = torch.randn(N, device="cuda") * 0.1
    ref = (x.double() @ w.double().t() + b.double())
    scale = ref.abs().max().item()
    e_exact = (ops.gemm_exact(x, ops.ExactWeight(w), bias=b).double() - ref).abs().max().item() / scale
    e_simt = (ops.gemm_simt(x, w, bias=b).double() - ref).abs().max().item() / scale
    e_tf32 = (ops.gemm(x, w, bias=b).double() - ref).abs().max().item() / scale
    print(f"relative to max |out|: 3xTF32 {e_exact:.2e}, fp32 SIMT {e_simt:.2e}, TF32 {e_tf32:.2e}")
    assert e_exact <= 3e-5 and e_simt <= 3e-6
    assert e_tf32 >= 30 * e_exact
    # erf GELU epilogue of the exact path
    g_ref = torch.nn.functional.gelu(ref)
    g = ops.gemm_exact(x, ops.ExactWeight(w), bias=b, gelu=True).double()
    assert (g - g_ref).abs().max().item() / g_ref.abs().max().item() <= 3e-5


@pytest.mark.parametrize("M,N,K", [(256, 4096, 2048), (200, 2048, 6144), (129, 2048, 2048), (64, 4096, 2048), (17, 2048, 6144),
                                   (1, 2048, 2048), (256, 1000, 2048 + 64)])
def test_splitk_partials_and_consumer_side_reduction(M, N, K):
    """mtts_gemm_splitk + mtts_splitk_reduce / mtts_splitk_reduce_rmsnorm against torch fp32 (same bf16 rounding points as
    the fused-epilogue GEMM followed by mtts_rmsnorm: those two paths must agree bit for bit up to fp32 summation order)."""
    import ctypes
    from moss_ttsd_b200 import _lib, ops
    L = _lib.load()
    ops.ensure_init()
    torch.manual_seed(M + N)
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    S = L.mtts_gemm_splitk_splits(M, N, K)
    ws = torch.empty(L.mtts_gemm_splitk_workspace_bytes(M, N, K) // 4, dtype=torch.float32, device="cuda")
    got_s = ctypes.c_int(0)
    _lib.check(L.mtts_gemm_splitk(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), ws.data_ptr(), ws.numel() * 4, M, N, K,
                                  ctypes.byref(got_s), _lib.stream_ptr()))
    assert got_s.value == S >= 1
    ref = x.float() @ w.float().t()
    part = ws[:S * M * N].view(S, M, N)
    assert (part.sum(0) - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()
    if S > 1:
        assert part[0].abs().max().item() > 0 and (part[0] - ref).abs().max().item() > 1e-3      # really partial sums
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    _lib.check(L.mtts_splitk_reduce(ws.data_ptr(), S, M, N, out.data_ptr(), out.stride(0), _lib.stream_ptr()))
    want = part.sum(0) if S == 1 else None
    acc = part[0].clone()
    for s in range(1, S):
        acc += part[s]
    assert torch.equal(out, acc.to(torch.bfloat16))                      # ascending-order fp32 sum, one bf16 rounding
    fused = ops.gemm(x, w)
    assert (out.float() - fused.float()).abs().max().item() <= 0.02 * ref.abs().max().item()
    if N % 8 == 0 and N <= 8192:
        res = torch.randn(M, N, device="cuda").to(torch.bfloat16)
        nw = (1 + 0.1 * torch.randn(N, device="cuda")).to(torch.bfloat16)
        xr = res.clone()
        xn = torch.empty_like(xr)
        _lib.check(L.mtts_splitk_reduce_rmsnorm(ws.data_ptr(), S, M, N, xr.data_ptr(), xr.stride(0), nw.data_ptr(), xn.data_ptr(),
                                                xn.stride(0), 1e-6, _lib.stream_ptr()))
        x_want = (res.float() + acc.to(torch.bfloat16).float()).to(torch.bfloat16)
        assert torch.equal(xr, x_want)
        xn_want = torch.empty_like(xr)
        _lib.check(L.mtts_rmsnorm(x_want.data_ptr(), x_want.stride(0), nw.data_ptr(), xn_want.data_ptr(), xn_want.stride(0), M, N,
                                  1e-6, _lib.stream_ptr()))
        d = (xn.float() - xn_want.float()).abs()
        assert d.max().item() <= 2.0 ** -6 * xn_want.float().abs().max().item()         # block- vs warp-order sum of squares
        assert (d > 0).float().mean().item() <= 0.01
