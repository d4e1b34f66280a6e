"""Decode attention kernels (tensor-core path, mtts_gqa_attention with one query row per sequence) against an fp32
torch restatement of HF eager attention (installed modeling_qwen3.py:184-219): ragged contexts, split-KV, paged cache
with shuffled pages, grouped / multi-head layouts."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(q, k, v, lens, Hq, Hkv):
    # q [B, Hq, D] ; k, v [B, T, Hkv, D] (fp32) ; row b attends keys 0..lens[b]-1
    B, _, D = q.shape
    G = Hq // Hkv
    out = torch.zeros_like(q)
    for b in range(B):
        n = int(lens[b])
        kk = k[b, :n].repeat_interleave(G, dim=1)  # [n, Hq, D]
        vv = v[b, :n].repeat_interleave(G, dim=1)
        s = torch.einsum("hd,nhd->hn", q[b], kk) / np.sqrt(D)
        p = torch.softmax(s, dim=-1)
        out[b] = torch.einsum("hn,nhd->hd", p, vv)
    return out


@pytest.mark.parametrize("B,Hq,Hkv,max_ctx,nsplit,paged", [
    (3, 16, 8, 70, 1, False), (3, 16, 8, 333, 3, True), (20, 16, 8, 1000, 1, False), (20, 16, 8, 1000, 8, True),
    (5, 16, 4, 257, 2, True), (4, 8, 8, 129, 1, False), (70, 16, 8, 460, 1, False), (2, 16, 8, 3000, 5, True)])
def test_decode_attention_matches_fp32_reference(B, Hq, Hkv, max_ctx, nsplit, paged):
    from moss_ttsd_b200 import _lib, ops
    ops.ensure_init()
    L = _lib.load()
    D, page = 128, 64
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + max_ctx)
    rng = np.random.default_rng(B + max_ctx)
    lens = rng.integers(1, max_ctx + 1, B)
    lens[0] = max_ctx
    if B > 1:
        lens[1] = 1
    max_pages = (max_ctx + page - 1) // page
    num_pages = B * max_pages
    k_pool = torch.randn((num_pages, Hkv, page, D), device="cuda", generator=g).to(torch.bfloat16)
    v_pool = torch.randn((num_pages, Hkv, page, D), device="cuda", generator=g).to(torch.bfloat16)
    q = torch.randn((B, Hq, D), device="cuda", generator=g).to(torch.bfloat16)
    ids = np.arange(num_pages, dtype=np.int32)
    if paged:
        rng.shuffle(ids)
    table = torch.from_numpy(ids.reshape(B, max_pages)).cuda()
    # gather the logical [B, T, Hkv, D] view for the reference
    kl = k_pool[table.long()].permute(0, 1, 3, 2, 4).reshape(B, max_pages * page, Hkv, D).float()
    vl = v_pool[table.long()].permute(0, 1, 3, 2, 4).reshape(B, max_pages * page, Hkv, D).float()
    ref = _reference(q.float(), kl, vl, lens, Hq, Hkv)
    pos = torch.from_numpy((lens - 1).astype(np.int32)).cuda()
    out = torch.empty((B, Hq * D), dtype=torch.bfloat16, device="cuda")
    ws = torch.zeros(L.mtts_gqa_attention_workspace_bytes(B, Hkv, Hq // Hkv, 1, nsplit), dtype=torch.uint8, device="cuda")
    for _ in range(2):  # second call: the split counters must have been left clean
        _lib.check(L.mtts_gqa_attention(q.data_ptr(), k_pool.data_ptr(), v_pool.data_ptr(), table.data_ptr() if paged else None,
                                        max_pages, page, None, None, None, pos.data_ptr(), out.data_ptr(), B, 1, Hq, Hkv, D,
                                        nsplit, ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    got = out.float().view(B, Hq, D)
    err = (got - ref).abs().max().item()
    # bf16 probabilities and bf16 output rounding: |O| <= ~3, one bf16 ulp there is 0.016
    assert err <= 0.03, err
    assert torch.isfinite(got).all()


@pytest.mark.parametrize("B,Hq,Hkv,max_ctx,nsplit,paged", [(3, 16, 8, 200, 1, False), (9, 16, 8, 333, 3, True),
                                                           (70, 16, 8, 130, 1, False), (5, 16, 4, 257, 2, True)])
def test_fused_decode_attention_is_bit_identical_to_the_two_kernel_path(B, Hq, Hkv, max_ctx, nsplit, paged):
    """mtts_gqa_decode_fused == mtts_qknorm_rope_kvappend + mtts_gqa_attention: same output bits, same cache rows."""
    from moss_ttsd_b200 import _lib, ops
    ops.ensure_init()
    L = _lib.load()
    D, page = 128, 64
    g = torch.Generator(device="cuda").manual_seed(7 * B + max_ctx)
    rng = np.random.default_rng(3 * B + max_ctx)
    lens = rng.integers(1, max_ctx + 1, B)
    lens[0] = max_ctx
    lens[-1] = 1
    if B > 2:
        lens[1] = 64 * 2 + 1  # the new row opens a page / a key tile
    max_pages = (max_ctx + page - 1) // page
    num_pages = B * max_pages
    ids = np.arange(num_pages, dtype=np.int32)
    if paged:
        rng.shuffle(ids)
    table = torch.from_numpy(ids.reshape(B, max_pages)).cuda()
    pos = torch.from_numpy((lens - 1).astype(np.int32)).cuda()
    qkv = torch.randn((B, (Hq + 2 * Hkv) * D), device="cuda", generator=g).to(torch.bfloat16)
    qn = (1 + 0.2 * torch.randn(D, device="cuda", generator=g)).to(torch.bfloat16)
    kn = (1 + 0.2 * torch.randn(D, device="cuda", generator=g)).to(torch.bfloat16)
    inv_freq = (1.0 / (1e6 ** (torch.arange(0, D, 2).float() / D))).cuda()
    pools = [torch.randn((num_pages, Hkv, page, D), device="cuda", generator=g).to(torch.bfloat16) for _ in range(2)]
    # everything from the new row on is stale memory in real life: poison it (0 x NaN must not reach the output)
    for b in range(B):
        for t_ in range(int(lens[b]) - 1, max_pages * page):
            pg = int(table[b, t_ // page])
            pools[0][pg, :, t_ % page] = float("nan")
            pools[1][pg, :, t_ % page] = float("nan")
    err = torch.zeros(4, dtype=torch.int32, device="cuda")
    tb = table.data_ptr() if paged else None
    outs, caches = [], []
    for fused in (False, True):
        k_pool, v_pool = pools[0].clone(), pools[1].clone()
        out = torch.empty((B, Hq * D), dtype=torch.bfloat16, device="cuda")
        ws = torch.zeros(L.mtts_gqa_attention_workspace_bytes(B, Hkv, Hq // Hkv, 1, nsplit), dtype=torch.uint8, device="cuda")
        if fused:
            _lib.check(L.mtts_gqa_decode_fused(qkv.data_ptr(), qkv.stride(0), qn.data_ptr(), kn.data_ptr(), inv_freq.data_ptr(), 1e-6,
                                               k_pool.data_ptr(), v_pool.data_ptr(), tb, max_pages, page, num_pages, pos.data_ptr(),
                                               out.data_ptr(), B, Hq, Hkv, D, nsplit, ws.data_ptr(), ws.numel(), err.data_ptr(),
                                               _lib.stream_ptr()))
        else:
            q = torch.empty((B, Hq * D), dtype=torch.bfloat16, device="cuda")
            _lib.check(L.mtts_qknorm_rope_kvappend(qkv.data_ptr(), qkv.stride(0), qn.data_ptr(), kn.data_ptr(), inv_freq.data_ptr(),
                                                   pos.data_ptr(), None, q.data_ptr(), k_pool.data_ptr(), v_pool.data_ptr(), tb,
                                                   max_pages, page, num_pages, B, Hq, Hkv, D, 1e-6, err.data_ptr(),
                                                   _lib.stream_ptr()))
            _lib.check(L.mtts_gqa_attention(q.data_ptr(), k_pool.data_ptr(), v_pool.data_ptr(), tb, max_pages, page, None, None,
                                            None, pos.data_ptr(), out.data_ptr(), B, 1, Hq, Hkv, D, nsplit, ws.data_ptr(),
                                            ws.numel(), _lib.stream_ptr()))
        torch.cuda.synchronize()
        outs.append(out)
        caches.append((k_pool, v_pool))
    assert int(err.abs().sum()) == 0
    assert torch.isfinite(outs[0].float()).all() and torch.isfinite(outs[1].float()).all()
    for a, b_ in zip(caches[0], caches[1]):
        assert torch.equal(torch.nan_to_num(a.float(), nan=123.0), torch.nan_to_num(b_.float(), nan=123.0))
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("lens,Hq,Hkv,paged", [([70, 1, 64, 65, 130], 16, 8, False), ([457, 300], 16, 8, True),
                                                ([33, 200, 17], 8, 8, True), ([129, 64], 16, 4, False)])
def test_prefill_attention_matches_fp32_reference(lens, Hq, Hkv, paged):
    """64-row causal tiles (flash-style tensor-core kernel) against fp32 eager attention, packed ragged sequences."""
    from moss_ttsd_b200 import _lib, ops
    ops.ensure_init()
    L = _lib.load()
    D, page = 128, 64
    B = len(lens)
    g = torch.Generator(device="cuda").manual_seed(sum(lens))
    rng = np.random.default_rng(sum(lens))
    max_ctx = max(lens)
    max_pages = (max_ctx + page - 1) // page
    num_pages = B * max_pages
    ids = np.arange(num_pages, dtype=np.int32)
    if paged:
        rng.shuffle(ids)
    table = torch.from_numpy(ids.reshape(B, max_pages)).cuda()
    k_pool = torch.randn((num_pages, Hkv, page, D), device="cuda", generator=g).to(torch.bfloat16)
    v_pool = torch.randn((num_pages, Hkv, page, D), device="cuda", generator=g).to(torch.bfloat16)
    R = sum(lens)
    q = torch.randn((R, Hq, D), device="cuda", generator=g).to(torch.bfloat16)
    pos_h = np.concatenate([np.arange(n, dtype=np.int32) for n in lens])
    seq_h = np.repeat(np.arange(B, dtype=np.int32), lens)
    cu = np.concatenate([[0], np.cumsum(lens)])
    row0_h, nrows_h = [], []
    for b in range(B):
        for t0 in range(0, lens[b], 64):
            row0_h.append(cu[b] + t0)
            nrows_h.append(min(64, lens[b] - t0))
    dev = lambda a: torch.from_numpy(np.asarray(a, dtype=np.int32)).cuda()
    positions, row_seq, tile_row0, tile_nrows = dev(pos_h), dev(seq_h), dev(row0_h), dev(nrows_h)
    out = torch.empty((R, Hq * D), dtype=torch.bfloat16, device="cuda")
    _lib.check(L.mtts_gqa_attention(q.data_ptr(), k_pool.data_ptr(), v_pool.data_ptr(), table.data_ptr() if paged else None,
                                    max_pages, page, tile_row0.data_ptr(), tile_nrows.data_ptr(), row_seq.data_ptr(),
                                    positions.data_ptr(), out.data_ptr(), len(row0_h), 64, Hq, Hkv, D, 1, None, 0,
                                    _lib.stream_ptr()))
    torch.cuda.synchronize()
    kl = k_pool[table.long()].permute(0, 1, 3, 2, 4).reshape(B, max_pages * page, Hkv, D).float()
    vl = v_pool[table.long()].permute(0, 1, 3, 2, 4).reshape(B, max_pages * page, Hkv, D).float()
    got = out.float().view(R, Hq, D)
    worst = 0.0
    for b in range(B):
        for r in range(lens[b]):
            ref = _reference(q[cu[b] + r:cu[b] + r + 1].float(), kl[b:b + 1], vl[b:b + 1], [r + 1], Hq, Hkv)[0]
            worst = max(worst, (got[cu[b] + r] - ref).abs().max().item())
    assert worst <= 0.03, worst
    assert torch.isfinite(got).all()


@pytest.mark.parametrize("lens,Hq,Hkv,paged,pos0", [([70, 1, 64, 65, 130], 16, 8, False, 0), ([457, 300], 16, 8, True, 0),
                                                     ([33, 200, 17], 8, 8, True, 0), ([129, 64, 128, 127], 16, 4, False, 0),
                                                     ([100, 31], 16, 8, True, 77)])
def test_prefill_attention_tcgen05_matches_fp32_reference(lens, Hq, Hkv, paged, pos0):
    """128-row causal tiles on tcgen05 (S and O in TMEM, K/V tiles of the paged pool by TMA) against fp32 eager attention:
    packed ragged sequences, shuffled pages, rows that continue a cache holding pos0 tokens; pool rows past each sequence's
    last key hold NaN bit patterns (an uninitialised pool) and must not reach the output."""
    from moss_ttsd_b200 import _lib, ops
    ops.ensure_init()
    L = _lib.load()
    D, page = 128, 64
    B = len(lens)
    g = torch.Generator(device="cuda").manual_seed(sum(lens) + pos0)
    rng = np.random.default_rng(sum(lens))
    max_ctx = max(lens) + pos0
    max_pages = (max_ctx + page - 1) // page
    num_pages = B * max_pages
    ids = np.arange(num_pages, dtype=np.int32)
    if paged:
        rng.shuffle(ids)
    table = torch.from_numpy(ids.reshape(B, max_pages)).cuda()
    k_log = torch.randn((B, max_pages * page, Hkv, D), device="cuda", generator=g).to(torch.bfloat16)
    v_log = torch.randn((B, max_pages * page, Hkv, D), device="cuda", generator=g).to(torch.bfloat16)
    for b in range(B):  # keys past the sequence: NaN
        k_log[b, pos0 + lens[b]:] = float("nan")
        v_log[b, pos0 + lens[b]:] = float("nan")
    k_pool = torch.empty((num_pages, Hkv, page, D), device="cuda", dtype=torch.bfloat16)
    v_pool = torch.empty_like(k_pool)
    k_pool[table.long().reshape(-1)] = k_log.view(B * max_pages, page, Hkv, D).permute(0, 2, 1, 3)
    v_pool[table.long().reshape(-1)] = v_log.view(B * max_pages, page, Hkv, D).permute(0, 2, 1, 3)
    R = sum(lens)
    q = torch.randn((R, Hq, D), device="cuda", generator=g).to(torch.bfloat16)
    pos_h = np.concatenate([np.arange(pos0, pos0 + n, dtype=np.int32) for n in lens])
    seq_h = np.repeat(np.arange(B, dtype=np.int32), lens)
    cu = np.concatenate([[0], np.cumsum(lens)])
    row0_h, nrows_h = [], []
    for b in range(B):
        for t0 in range(0, lens[b], 128):
            row0_h.append(cu[b] + t0)
            nrows_h.append(min(128, lens[b] - t0))
    dev = lambda a: torch.from_numpy(np.asarray(a, dtype=np.int32)).cuda()
    positions, row_seq, tile_row0, tile_nrows = dev(pos_h), dev(seq_h), dev(row0_h), dev(nrows_h)
    out = torch.full((R, Hq * D), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(L.mtts_gqa_prefill_tc(q.data_ptr(), R, k_pool.data_ptr(), v_pool.data_ptr(), table.data_ptr() if paged else None,
                                     max_pages, page, num_pages, tile_row0.data_ptr(), tile_nrows.data_ptr(), row_seq.data_ptr(),
                                     positions.data_ptr(), out.data_ptr(), len(row0_h), Hq, Hkv, D, _lib.stream_ptr()))
    torch.cuda.synchronize()
    got = out.float().view(R, Hq, D)
    assert torch.isfinite(got).all()
    kl, vl = k_log.float(), v_log.float()
    worst = 0.0
    for b in range(B):
        for r in range(lens[b]):
            ref = _reference(q[cu[b] + r:cu[b] + r + 1].float(), kl[b:b + 1], vl[b:b + 1], [pos0 + r + 1], Hq, Hkv)[0]
            worst = max(worst, (got[cu[b] + r] - ref).abs().max().item())
    assert worst <= 0.03, worst
