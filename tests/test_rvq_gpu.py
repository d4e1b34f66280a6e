"""ResidualVQ CUDA kernels vs the numpy oracle (oracle/rvq_np.py), near-ties adjudicated in fp64."""
import numpy as np
import pytest
import torch

from oracle import rvq_np

pytestmark = pytest.mark.gpu

TIE_RTOL = 1e-6  # north_star: code indices identical except distance ties within 1e-6 (relative, fp64-adjudicated)


def _mk(N, nq, K, D, seed):
    rng = np.random.default_rng(seed)
    cb = (0.1 * rng.standard_normal((nq, K, D))).astype(np.float32)
    # a residual-like input: sum of one code per layer plus noise, so every layer has structure to find
    z = np.zeros((N, D), dtype=np.float32)
    for i in range(nq):
        z += cb[i][rng.integers(0, K, N)]
    z += (0.05 * rng.standard_normal((N, D))).astype(np.float32)
    return z.astype(np.float32), cb


def _adjudicate(e, cb_layer, got, want):
    """Every mismatch must be a near-tie of the fp64 distances."""
    bad = np.nonzero(got != want)[0]
    if bad.size == 0:
        return 0
    d = rvq_np.vq_dist64(e[bad], cb_layer)
    dg = d[np.arange(bad.size), got[bad]]
    dw = d[np.arange(bad.size), want[bad]]
    tol = TIE_RTOL * np.maximum(1.0, np.abs(dw))
    assert np.all(np.abs(dg - dw) <= tol), (bad[:8], dg[:8], dw[:8])
    return bad.size


@pytest.mark.parametrize("N,nq,K,D", [(1, 8, 1024, 512), (37, 8, 1024, 512), (1000, 8, 1024, 512), (70, 3, 128, 64)])
def test_rvq_encode_teacher_forced(N, nq, K, D):
    """Per layer, with the ORACLE's residual fed in (SURVEY §7 hard parts): one flip cannot cascade."""
    from moss_ttsd_b200 import ops
    z, cb = _mk(N, nq, K, D, seed=N + K)
    codes_o, zq_o, res_o, layer_in = rvq_np.rvq_forward(z, cb)
    cbt = torch.from_numpy(cb).cuda()
    norms = ops.rvq_codebook_norms(cbt)
    ties = 0
    for i in range(nq):
        e = torch.from_numpy(layer_in[i]).cuda()
        codes, _, _ = ops.rvq_encode(e, cbt[i:i + 1].contiguous(), norms[i:i + 1].contiguous(), want_zq=False)
        ties += _adjudicate(layer_in[i], cb[i], codes[0].cpu().numpy(), codes_o[i])
    assert ties <= max(1, N * nq // 1000)


@pytest.mark.parametrize("N", [5, 333, 4000])
def test_rvq_encode_free_running_and_mask(N):
    from moss_ttsd_b200 import ops
    nq, K, D = 8, 1024, 512
    z, cb = _mk(N, nq, K, D, seed=11 * N)
    rng = np.random.default_rng(N)
    valid = rng.random(N) > 0.2
    codes_o, zq_o, res_o, _ = rvq_np.rvq_forward(z, cb, valid)
    cbt = torch.from_numpy(cb).cuda()
    norms = ops.rvq_codebook_norms(cbt)
    codes, zq, res = ops.rvq_encode(torch.from_numpy(z).cuda(), cbt, norms, valid=torch.from_numpy(valid).cuda(),
                                    want_residual=True)
    codes = codes.cpu().numpy()
    rows_equal = (codes == codes_o).all(0)
    # free-running mismatch rate is reported separately from the teacher-forced gate; it must stay tiny
    assert (~rows_equal).mean() <= 2e-3, (~rows_equal).mean()
    ok = rows_equal
    np.testing.assert_array_equal(zq.cpu().numpy()[ok], zq_o[ok])       # bit-exact where the codes agree
    np.testing.assert_array_equal(res.cpu().numpy()[ok], res_o[ok])
    # masked rows: quantise the zero vector, contribute nothing
    inv = ~valid
    if inv.any():
        zero_codes = rvq_np.rvq_forward(np.zeros((1, D), np.float32), cb)[0][:, 0]
        # the reference searches the zero vector at EVERY layer for masked rows (residual is not updated)
        for i in range(nq):
            want = rvq_np.vq_search(np.zeros((1, D), np.float32), cb[i])[0]
            assert (codes[i][inv] == want).all()
        assert (zq.cpu().numpy()[inv] == 0).all()
        np.testing.assert_array_equal(res.cpu().numpy()[inv], z[inv])


@pytest.mark.parametrize("N,nq", [(1, 8), (375, 8), (1000, 5)])
def test_rvq_decode_bit_exact(N, nq):
    from moss_ttsd_b200 import ops
    K, D = 1024, 512
    rng = np.random.default_rng(N)
    cb = (0.1 * rng.standard_normal((8, K, D))).astype(np.float32)
    codes = rng.integers(0, K, (nq, N)).astype(np.int64)
    out = ops.rvq_decode(torch.from_numpy(codes).cuda(), torch.from_numpy(cb).cuda())
    np.testing.assert_array_equal(out.cpu().numpy(), rvq_np.rvq_decode(codes, cb))


def test_rvq_empty_and_bad_codes():
    from moss_ttsd_b200 import ops
    cb = torch.randn(8, 1024, 512, device="cuda")
    norms = ops.rvq_codebook_norms(cb)
    codes, zq, _ = ops.rvq_encode(torch.zeros(0, 512, device="cuda"), cb, norms)
    assert codes.shape == (8, 0) and zq.shape == (0, 512)
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    bad = torch.full((8, 4), 5000, dtype=torch.int64, device="cuda")
    out = ops.rvq_decode(bad, cb, err_flag=flag)
    assert flag.item() == 1 and (out == 0).all()
