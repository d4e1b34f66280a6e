"""Tensor-level wrappers over the C-ABI (torch is used for device memory and streams only).

Every function validates that its tensors are CUDA tensors, passes raw pointers/strides to libmtts and
raises `MttsError` on a non-zero return. Nothing here computes with torch ops.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

BF16, F32, F16 = 0, 1, 2
EPI_BIAS, EPI_GELU, EPI_GAMMA, EPI_RESIDUAL, EPI_SWIGLU, EPI_EXACT_ACT = 1, 2, 4, 8, 16, 32


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float16:
        return F16
    raise TypeError(f"unsupported dtype {t.dtype}")


_inited = set()


def ensure_init():
    """mtts_init() once per device (sets kernel attributes; must happen outside CUDA-graph capture)."""
    d = torch.cuda.current_device()
    if d not in _inited:
        check(_lib.load().mtts_init())
        _inited.add(d)


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.MttsError("libmtts ops need CUDA tensors; there is no CPU path")
    ensure_init()


_workspaces: dict = {}


def gemm_workspace(device, nbytes: int) -> torch.Tensor:
    """Per (device, stream) scratch for split-K partials; grown on demand, counter area kept zero."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream_ptr())
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 22), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def split_tf32x3(x: torch.Tensor, weights_order: bool = False) -> torch.Tensor:
    """fp32 [rows, K] -> [rows, 3K] hi/lo TF32 split (mtts_split_tf32x3): [hi|lo|hi] for activations, [hi|hi|lo] for
    weights, so that a TF32 tensor-core GEMM over 3K computes the product to fp32 accuracy."""
    _cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    rows, K = x.shape
    out = torch.empty((rows, 3 * K), dtype=torch.float32, device=x.device)
    check(_lib.load().mtts_split_tf32x3(ptr(x), x.stride(0), ptr(out), out.stride(0), rows, K, 1 if weights_order else 0,
                                        stream_ptr()))
    return out


class ExactWeight:
    """An fp32 weight [N, K] together with its [hi|hi|lo] split [N, 3K], built once at load time (exact-mode GEMMs)."""

    def __init__(self, w: torch.Tensor):
        self.w = w
        self.split = split_tf32x3(w.contiguous(), weights_order=True)
        self.shape = w.shape


def gemm_exact(x, w: ExactWeight, out=None, **kw):
    """fp32-accurate `x @ w^T` on the tcgen05 TF32 path: 3xTF32 (hi*hi + lo*hi + hi*lo, fp32 accumulate), erff GELU."""
    return gemm(split_tf32x3(x), w.split, out=out, _exact_act=True, **kw)


def gemm(x, w, out=None, *, out_dtype=None, bias=None, gelu=False, gamma=None, residual=None, swiglu=False,
         workspace=None, _exact_act=False):
    """out[M,N] = epi(x[M,K] @ w[N,K]^T) on the tcgen05 path (mtts_gemm)."""
    _cuda(x, w, out, bias, gamma, residual)
    assert x.dim() == 2 and w.dim() == 2 and x.shape[1] == w.shape[1], (x.shape, w.shape)
    assert x.stride(1) == 1 and w.stride(1) == 1
    M, K = x.shape
    N = w.shape[0]
    out_dtype = out_dtype or (out.dtype if out is not None else x.dtype)
    n_out = N // 2 if swiglu else N
    if out is None:
        out = torch.empty((M, n_out), dtype=out_dtype, device=x.device)
    assert out.shape == (M, n_out) and out.stride(1) == 1
    flags = 0
    if bias is not None:
        flags |= EPI_BIAS
        assert bias.dtype == torch.float32 and bias.numel() == N
    if gelu:
        flags |= EPI_GELU
    if gamma is not None:
        flags |= EPI_GAMMA
        assert gamma.dtype == torch.float32 and gamma.numel() == N
    if residual is not None:
        flags |= EPI_RESIDUAL
        assert residual.dtype == out.dtype and residual.shape == out.shape and residual.stride(1) == 1
    if swiglu:
        flags |= EPI_SWIGLU
    if _exact_act:
        flags |= EPI_EXACT_ACT
    L = _lib.load()
    need = L.mtts_gemm_workspace_bytes(M, N, K, _dt(x))
    ws = workspace if workspace is not None else gemm_workspace(x.device, need)
    check(L.mtts_gemm(ptr(x), x.stride(0), ptr(w), w.stride(0), ptr(out), out.stride(0), M, N, K, _dt(x), _dt(out),
                      flags, ptr(bias), ptr(gamma), ptr(residual), residual.stride(0) if residual is not None else 0,
                      ptr(ws), ws.numel(), stream_ptr()))
    return out


def gemm_simt(x, w, out=None, *, out_dtype=None, bias=None, gelu=False, gamma=None, residual=None,
              x_layout=None, M=None):
    """Exact-fp32 CUDA-core GEMM (mtts_gemm_simt). `x_layout=(rows_per_batch, batch_stride, row_stride, k_stride)`
    lets x be read in place from e.g. a channel-major (B, C, T) tensor."""
    _cuda(x, w, out, bias, gamma, residual)
    N, K = w.shape
    if x_layout is None:
        assert x.dim() == 2 and x.shape[1] == K
        M = x.shape[0]
        x_layout = (max(M, 1), 0, x.stride(0), x.stride(1))
    assert M is not None
    out_dtype = out_dtype or (out.dtype if out is not None else x.dtype)
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=x.device)
    flags = 0
    if bias is not None:
        flags |= EPI_BIAS
    if gelu:
        flags |= EPI_GELU
    if gamma is not None:
        flags |= EPI_GAMMA
    if residual is not None:
        flags |= EPI_RESIDUAL
    L = _lib.load()
    check(L.mtts_gemm_simt(ptr(x), x_layout[0], x_layout[1], x_layout[2], x_layout[3], ptr(w), w.stride(0), ptr(out),
                           out.stride(0), M, N, K, _dt(x), _dt(out), flags, ptr(bias), ptr(gamma), ptr(residual),
                           residual.stride(0) if residual is not None else 0, stream_ptr()))
    return out


# ------------------------------------------------------------------------------------------------ RVQ
def rvq_codebook_norms(codebooks: torch.Tensor) -> torch.Tensor:
    _cuda(codebooks)
    nq, K, D = codebooks.shape
    assert codebooks.dtype == torch.float32 and codebooks.is_contiguous()
    norms = torch.empty((nq, K), dtype=torch.float32, device=codebooks.device)
    check(_lib.load().mtts_rvq_codebook_norms(ptr(codebooks), nq, K, D, ptr(norms), stream_ptr()))
    return norms


def rvq_encode(z, codebooks, norms, valid=None, want_zq=True, want_residual=False):
    """z [N, D] fp32 token-major -> (codes [nq, N] int64, zq [N, D] | None, residual [N, D] | None)."""
    _cuda(z, codebooks, norms, valid)
    assert z.dtype == torch.float32 and z.is_contiguous() and z.dim() == 2
    nq, K, D = codebooks.shape
    N = z.shape[0]
    assert z.shape[1] == D
    codes = torch.empty((nq, N), dtype=torch.int64, device=z.device)
    zq = torch.empty((N, D), dtype=torch.float32, device=z.device) if want_zq else None
    res = torch.empty((N, D), dtype=torch.float32, device=z.device) if want_residual else None
    if valid is not None:
        assert valid.numel() == N and valid.dtype in (torch.bool, torch.uint8) and valid.is_contiguous()
    check(_lib.load().mtts_rvq_encode(ptr(z), ptr(valid), ptr(codebooks), ptr(norms), N, nq, K, D, ptr(codes), ptr(zq),
                                      ptr(res), stream_ptr()))
    return codes, zq, res


def rvq_decode(codes, codebooks, err_flag=None):
    """codes [nq, N] int64 -> [N, D] fp32 token-major sum of code vectors."""
    _cuda(codes, codebooks)
    assert codes.dtype == torch.int64 and codes.dim() == 2 and codes.stride(1) == 1
    nq_all, K, D = codebooks.shape
    nq, N = codes.shape
    assert nq <= nq_all
    out = torch.empty((N, D), dtype=torch.float32, device=codes.device)
    check(_lib.load().mtts_rvq_decode(ptr(codes), codes.stride(0), ptr(codebooks), N, nq, K, D, ptr(out), ptr(err_flag),
                                      stream_ptr()))
    return out
