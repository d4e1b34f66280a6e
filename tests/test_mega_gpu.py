"""The persistent small-batch decode kernel (mtts_decode_mega) against the kernel chain it replaces
(mtts_rmsnorm / mtts_gemm / mtts_qknorm_rope_kvappend / mtts_gqa_attention), on a full-width model with a few layers.
The chain itself is pinned to the oracle / reference goldens in test_lm_gpu.py; a full-width oracle comparison is at the
end of this file."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _engine_pair(num_layers, seed=1):
    from moss_ttsd_b200.lm_engine import DecoderEngine, LMShape, LMWeights
    shape = LMShape(num_hidden_layers=num_layers)
    w = LMWeights(shape, "cuda").init_random_(seed=seed, std=0.02)
    g = torch.Generator(device="cuda").manual_seed(seed + 100)
    for L in w.layers:  # norm weights away from 1 so that every scale factor matters
        for k in ("ln1", "ln2", "q_norm", "k_norm"):
            L[k].copy_((1.0 + 0.2 * torch.randn(L[k].shape, device="cuda", generator=g)).to(torch.bfloat16))
    w.final_norm.copy_((1.0 + 0.2 * torch.randn(w.final_norm.shape, device="cuda", generator=g)).to(torch.bfloat16))
    chain = DecoderEngine(w)
    chain.use_mega = False
    chain.use_graph = False
    mega = DecoderEngine(w)
    mega.use_graph = False
    mega.mega_max_b = 4  # the kernel supports batch 1..4; the engine only routes batch <= 2 to it by default
    return shape, w, chain, mega


def _prompt(shape, B, P, seed):
    rng = np.random.default_rng(seed)
    C = shape.channels
    ids = np.zeros((B, P, C), dtype=np.int64)
    ids[:, :, 0] = rng.integers(0, shape.vocab_size, (B, P))
    ids[:, :, 1:] = rng.integers(0, 1025, (B, P, C - 1))
    mask = np.ones((B, P), dtype=np.int64)
    for b in range(B):  # ragged, left-padded
        mask[b, :(b * 5) % 23] = 0
    return torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()


def _session(eng, shape, B, ids, mask, paged, rows=256):
    from moss_ttsd_b200.lm_engine import KVCache, SamplerSetup
    cache = KVCache(shape, B, rows, "cuda", paged=paged, shuffle_pages=paged)
    sm = SamplerSetup(shape, [False] * shape.channels, None)
    st = eng.make_decode_state(B, cache, sm, rows, (151665, 152689), 152694, False)
    eng.reset_decode_state(st, 0, ids.shape[1], rows - 16)
    logits, lens = eng.prefill(ids, mask, cache)
    st["positions"].copy_((lens - 1).to(torch.int32))
    eng.sample_and_advance(st, logits)
    return st, cache


@pytest.mark.parametrize("B,paged,P,hint", [(1, False, 70, 0), (3, True, 61, 0), (4, False, 130, 0), (2, True, 200, 0),
                                            (1, True, 150, 900), (2, False, 150, 900)])  # the 8 / 6 key-range layouts of long contexts
def test_mega_step_matches_kernel_chain(B, paged, P, hint):
    shape, w, chain, mega = _engine_pair(3)
    ids, mask = _prompt(shape, B, P, seed=B)
    st_c, cache_c = _session(chain, shape, B, ids, mask, paged)
    mega.ctx_hint = hint
    st_m, cache_m = _session(mega, shape, B, ids, mask, paged)
    mega.ctx_hint = 0
    assert st_m["mega"] is not None and st_c["mega"] is None
    assert st_m["mega"]["nsplit"] == ({1: 8, 2: 6}.get(B, 4) if hint >= 700 else 4)
    assert torch.equal(st_c["tokens"], st_m["tokens"])
    worst, agree, total = 0.0, 0, 0
    for step in range(12):
        chain.decode_step(st_c)
        mega.decode_step(st_m)
        lc, lm = st_c["logits"].float(), st_m["logits"].float()
        assert torch.isfinite(lm).all()
        d = (lc - lm).abs().max().item()
        worst = max(worst, d)
        # different fp32 summation order (16-way K split vs UMMA split-K) -> bf16 noise only
        assert d <= 0.04 * max(1.0, lc.abs().max().item()), f"step {step}: max|dlogit| {d}"
        agree += int((st_c["tokens"] == st_m["tokens"]).sum())
        total += st_c["tokens"].numel()
        # a differing pick must be a near-tie of the chain's own logits (random-init weights: top-2 gaps below the
        # bf16 noise between the two summation orders do occur), never a different decision on a clear margin
        for b, c in (st_c["tokens"] != st_m["tokens"]).nonzero().tolist():
            off = shape.head_offsets[c]
            tc, tm = int(st_c["tokens"][b, c]), int(st_m["tokens"][b, c])
            gap = (lc[b, off + tc] - lc[b, off + tm]).abs().item()
            assert gap <= 2 * d + 1e-6, f"step {step} row {b} channel {c}: picks {tc}/{tm} differ on a gap of {gap} (noise {d})"
        assert torch.equal(st_c["positions"], st_m["positions"])
        st_m["tokens"].copy_(st_c["tokens"])  # keep both paths on the same token stream
    assert agree >= 0.9 * total, (agree, total)
    # the K/V rows appended by both paths agree to bf16 rounding of near-identical inputs
    for l in range(shape.num_hidden_layers):
        kc, km = cache_c.k[l].float(), cache_m.k[l].float()
        pos = st_c["positions"].cpu().numpy()
        for b in range(B):
            for t in range(int(pos[b]) - 10, int(pos[b])):
                page = int(cache_c.block_table[b, t // 64]) if paged else b * cache_c.max_pages + t // 64
                a, m_ = kc[page, :, t % 64], km[page, :, t % 64]
                assert (a - m_).abs().max().item() <= 0.05 * max(1.0, a.abs().max().item())
    assert int(mega.err.abs().sum()) == 0


def test_mega_graph_replay_matches_eager():
    """Captured in a CUDA graph (cooperative kernel node) the step produces the same tokens as eager launches."""
    shape, w, chain, mega = _engine_pair(2, seed=5)
    B, P = 2, 40
    ids, mask = _prompt(shape, B, P, seed=9)
    st_a, _ = _session(mega, shape, B, ids, mask, False)
    toks_a = []
    for _ in range(10):
        mega.decode_step(st_a)
        toks_a.append(st_a["tokens"].clone())
    mega.use_graph = True
    st_b, _ = _session(mega, shape, B, ids, mask, False)
    toks_b = []
    for _ in range(10):
        mega.decode_step(st_b)
        toks_b.append(st_b["tokens"].clone())
    assert st_b["graph"] is not None
    assert torch.equal(torch.stack(toks_a), torch.stack(toks_b))


def test_mega_rejects_unsupported_shapes():
    from moss_ttsd_b200 import _lib
    L = _lib.load()
    assert L.mtts_decode_mega_supported(2048, 6144, 16, 8, 128, 4) == 1
    assert L.mtts_decode_mega_supported(2048, 6144, 16, 8, 128, 5) == 0
    assert L.mtts_decode_mega_supported(1024, 3072, 16, 8, 128, 1) == 0
    args = _lib.DecodeMegaArgs(hidden=1024, intermediate=3072, num_q_heads=16, num_kv_heads=8, head_dim=128, B=1)
    import ctypes
    rc = L.mtts_decode_mega(ctypes.byref(args), None)
    assert rc != 0 and b"unsupported shape" in L.mtts_last_error()


def test_full_width_decode_logits_match_the_fp32_oracle():
    """Teacher-forced decode-step logits of BOTH decode paths (kernel chain, persistent kernel) at the real model width
    (2 layers, tied 152697-row heads) against the fp32 oracle restatement of the reference (oracle/lm_oracle.py)."""
    from moss_ttsd_b200.lm_engine import DecoderEngine, KVCache, LMShape, LMWeights, SamplerSetup
    from oracle import lm_oracle
    shape_d = dict(hidden_size=2048, intermediate_size=6144, num_hidden_layers=2, num_attention_heads=16,
                   num_key_value_heads=8, head_dim=128, rms_norm_eps=1e-6, rope_theta=1e6, vocab_size=152697,
                   speech_vocab_size=1025, channels=8, speech_token_range=[151665, 152689])
    sd = lm_oracle.make_weights(shape_d, 77, tied=True)
    shape = LMShape(num_hidden_layers=2)
    w = LMWeights(shape, "cuda").load_state_dict(sd, tie_word_embeddings=True)
    rng = np.random.default_rng(5)
    B, P, n = 2, 24, 6
    full = np.zeros((B, P + n, 8), dtype=np.int64)
    full[:, :, 0] = rng.integers(151665, 152689, (B, P + n))
    full[:, :, 1:] = rng.integers(0, 1024, (B, P + n, 7))
    mask = np.ones((B, P + n), dtype=np.int64)
    mask[1, :3] = 0
    ref = lm_oracle.OracleLM(shape_d, sd, torch.float32).logits_all(torch.from_numpy(full), torch.from_numpy(mask))
    ref = torch.cat([r for r in ref], dim=-1)  # (B, P+n, sum vocab), unpadded heads
    offs, vocabs = shape.head_offsets, shape.vocabs
    ids, m = torch.from_numpy(full).cuda(), torch.from_numpy(mask).cuda()
    worst = {}
    for name, use_mega in (("chain", False), ("mega", True)):
        eng = DecoderEngine(w)
        eng.use_mega, eng.use_graph, eng.mega_max_b = use_mega, False, 4
        cache = KVCache(shape, B, 128, "cuda")
        st = eng.make_decode_state(B, cache, SamplerSetup(shape, [False] * 8, None), 128, (151665, 152689), 152694, False)
        eng.reset_decode_state(st, 0, P, 100)
        lg, lens = eng.prefill(ids[:, :P], m[:, :P], cache)
        st["positions"].copy_((lens - 1).to(torch.int32))
        eng.sample_and_advance(st, lg)
        assert (st["mega"] is not None) == use_mega
        err = 0.0
        for k in range(n):
            st["tokens"].copy_(ids[:, P + k])
            eng.decode_step(st)
            got = st["logits"].float().cpu()
            got = torch.cat([got[:, offs[c]:offs[c] + vocabs[c]] for c in range(8)], dim=-1)
            err = max(err, (got - ref[:, P + k]).abs().max().item())
        worst[name] = err
    scale = ref.abs().max().item()
    # bf16 activations / weights against an fp32 reference: a few bf16 ulps of the largest logit
    assert worst["chain"] <= 0.03 * max(1.0, scale), (worst, scale)
    assert worst["mega"] <= 0.03 * max(1.0, scale), (worst, scale)


@pytest.mark.parametrize("B", [70, 130, 256])
def test_splitk_decode_chain_matches_the_fused_epilogue_chain(B):
    """Decode steps at batch 65..256 through the consumer-side split-K path (mtts_gemm_splitk -> attention prologue /
    residual+RMSNorm reducers; gate/up on the CTA-pair kernel above batch 128) against the cluster split-K chain with
    fused epilogues: same bf16 rounding points, different fp32 summation order -> logits within a few bf16 ulps and
    the same greedy tokens wherever the top-2 gap is not a near-tie; and against the fp32 oracle for a few rows."""
    from moss_ttsd_b200.lm_engine import DecoderEngine, KVCache, LMShape, LMWeights, SamplerSetup
    from oracle import lm_oracle
    shape_d = dict(hidden_size=2048, intermediate_size=6144, num_hidden_layers=2, num_attention_heads=16,
                   num_key_value_heads=8, head_dim=128, rms_norm_eps=1e-6, rope_theta=1e6, vocab_size=152697,
                   speech_vocab_size=1025, channels=8, speech_token_range=[151665, 152689])
    sd = lm_oracle.make_weights(shape_d, 78, tied=True)
    shape = LMShape(num_hidden_layers=2)
    w = LMWeights(shape, "cuda").load_state_dict(sd, tie_word_embeddings=True)
    rng = np.random.default_rng(B)
    P, n = 12, 3
    full = np.zeros((B, P + n, 8), dtype=np.int64)
    full[:, :, 0] = rng.integers(151665, 152689, (B, P + n))
    full[:, :, 1:] = rng.integers(0, 1024, (B, P + n, 7))
    mask = np.ones((B, P + n), dtype=np.int64)
    mask[1::3, :2] = 0
    ids, m = torch.from_numpy(full).cuda(), torch.from_numpy(mask).cuda()
    got = {}
    for name, use in (("fused", False), ("splitk", True)):
        eng = DecoderEngine(w)
        eng.use_splitk, eng.use_graph, eng.fuse_heads = use, False, False
        cache = KVCache(shape, B, 64, "cuda")
        st = eng.make_decode_state(B, cache, SamplerSetup(shape, [False] * 8, None), 64, (151665, 152689), 152694, False)
        assert (st["pws"] is not None) == use
        eng.reset_decode_state(st, 0, P, 100)
        lg, lens = eng.prefill(ids[:, :P], m[:, :P], cache)
        st["positions"].copy_((lens - 1).to(torch.int32))
        eng.sample_and_advance(st, lg)
        steps = []
        for k in range(n):
            st["tokens"].copy_(ids[:, P + k])
            eng.decode_step(st)
            steps.append(st["logits"].float().clone())
        got[name] = torch.stack(steps, 1)            # (B, n, vpad)
        assert not eng.err.cpu().any()
    a, b = got["fused"], got["splitk"]
    scale = a.abs().max().item()
    assert (a - b).abs().max().item() <= 0.02 * max(1.0, scale), ((a - b).abs().max().item(), scale)
    offs = shape.head_offsets
    sl = a[..., offs[1]:offs[1] + 1024], b[..., offs[1]:offs[1] + 1024]
    flips = sl[0].argmax(-1) != sl[1].argmax(-1)
    top2 = sl[0].topk(2, -1).values
    assert ((top2[..., 0] - top2[..., 1])[flips] <= 0.05 * max(1.0, scale)).all()
    # fp32 oracle on the first 3 sequences
    ref = lm_oracle.OracleLM(shape_d, sd, torch.float32).logits_all(torch.from_numpy(full[:3]), torch.from_numpy(mask[:3]))
    ref = torch.cat([r for r in ref], dim=-1)[:, P:]
    ours = torch.cat([b[:3, :, offs[c]:offs[c] + shape.vocabs[c]].cpu() for c in range(8)], dim=-1)
    assert (ours - ref).abs().max().item() <= 0.03 * max(1.0, ref.abs().max().item())
