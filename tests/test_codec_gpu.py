"""CUDA codec-decode path vs goldens from the reference XY_Tokenizer.decode and vs the pinned CPU oracle."""
import numpy as np
import pytest
import torch

from tests.common import gold

pytestmark = pytest.mark.gpu

# Stated waveform tolerance (north_star "within a stated SNR"): the dense layers run as TF32 tensor-core GEMMs
# (10-bit mantissa operands, fp32 accumulate); everything else is fp32.
SNR_DB = 30.0


def snr_db(ref, got):
    ref, got = np.asarray(ref, np.float64), np.asarray(got, np.float64)
    return 10 * np.log10((ref ** 2).sum() / max(((ref - got) ** 2).sum(), 1e-30))


def _spt(gp, seed):
    from moss_ttsd_b200.xy_tokenizer.model import XY_Tokenizer
    from oracle.codec_weights import make_codec_weights
    spt = XY_Tokenizer(gp)
    spt.load_state_dict({k: torch.from_numpy(v) for k, v in make_codec_weights(gp, seed).items()})
    return spt.to("cuda")


@pytest.mark.parametrize("name,lens", [("tiny", [30, 11]), ("tiny_long", [400, 120])])
def test_decode_matches_reference_golden(name, lens):
    from oracle.codec_weights import TINY_CODEC
    g = gold("codec_decode.npz")
    spt = _spt(TINY_CODEC, int(g[f"{name}_seed"]))
    codes = [torch.from_numpy(g[f"{name}_codes{i}"].astype(np.int64)).cuda() for i in range(len(lens))]
    wavs = spt.decode(codes, overlap_seconds=10)["syn_wav_list"]
    for i, n in enumerate(lens):
        w = wavs[i].cpu().numpy()
        assert w.shape == (n * 1920,) and w.dtype == np.float32
        ref = g[f"{name}_wav{i}"]
        if name == "tiny_long":
            w = w[::8]
        s = snr_db(ref, w)
        print(f"{name}[{i}] SNR {s:.1f} dB")
        assert s >= SNR_DB, s


def test_decode_full_config_matches_reference_golden():
    from oracle.codec_weights import full_codec_params
    g = gold("codec_decode.npz")
    spt = _spt(full_codec_params(), int(g["full_seed"]))
    codes = [torch.from_numpy(g["full_codes0"].astype(np.int64)).cuda()]
    w = spt.decode(codes)["syn_wav_list"][0].cpu().numpy()
    s = snr_db(g["full_wav0"], w)
    print(f"full config SNR {s:.1f} dB")
    assert w.shape == g["full_wav0"].shape
    assert s >= SNR_DB, s


@pytest.mark.parametrize("mode", ["f16", "tf32"])
def test_decode_gemm_modes_keep_the_snr(mode):
    """Both operand modes of the decoder's large GEMMs — fp16 (default) and TF32 — against the reference waveforms (tiny
    two-window ragged golden and the shipped-config golden): stated gate 40 dB for either (both carry a 10-bit mantissa)."""
    from oracle.codec_weights import TINY_CODEC, full_codec_params
    g = gold("codec_decode.npz")
    worst = 1e9
    for name, gp, n_items in (("tiny_long", TINY_CODEC, 2), ("full", full_codec_params(), 1)):
        spt = _spt(gp, int(g[f"{name}_seed"]))
        assert spt.decode_gemm == "f16"
        spt.decode_gemm = mode
        codes = [torch.from_numpy(g[f"{name}_codes{i}"].astype(np.int64)).cuda() for i in range(n_items)]
        wavs = spt.decode(codes, overlap_seconds=10)["syn_wav_list"]
        for i in range(n_items):
            w = wavs[i].cpu().numpy()
            ref = g[f"{name}_wav{i}"]
            s = snr_db(ref, w[::8] if name == "tiny_long" else w)
            print(f"decode_gemm={mode} {name}[{i}] SNR {s:.1f} dB")
            worst = min(worst, s)
    assert worst >= 40.0, worst


def test_inference_detokenize_shapes_and_empty():
    from oracle.codec_weights import TINY_CODEC
    spt = _spt(TINY_CODEC, 21)
    codes = torch.randint(0, 128, (8, 3, 20), device="cuda")
    out = spt.inference_detokenize(codes, torch.tensor([20, 7, 13]))
    assert out["y"].shape == (3, 1, 20 * 1920)
    assert out["output_length"].tolist() == [20 * 1920, 7 * 1920, 13 * 1920]
    assert spt.decode([])["syn_wav_list"] == []
    z = spt.decode([torch.zeros(8, 0, dtype=torch.long)])["syn_wav_list"]
    assert z[0].numel() == 0
    assert spt.input_sample_rate == 16000 and spt.output_sample_rate == 24000 and spt.nq == 8
    assert spt.encoder_downsample_rate == 1280 and spt.decoder_upsample_rate == 1920


def test_batched_decode_equals_oracle_batched():
    """A ragged batch goes through the same padded computation as the reference's batched decode (the Vocos tail of
    the shorter item depends on the padding) — compare against the oracle run on the same batch."""
    from oracle.codec_oracle import CodecOracle
    from oracle.codec_weights import TINY_CODEC, make_codec_weights
    rng = np.random.default_rng(5)
    codes = [torch.from_numpy(rng.integers(0, 128, (8, n)).astype(np.int64)) for n in (60, 25, 44)]
    with torch.no_grad():
        ref = CodecOracle(TINY_CODEC, make_codec_weights(TINY_CODEC, 21)).decode(codes)
    spt = _spt(TINY_CODEC, 21)
    got = spt.decode([c.cuda() for c in codes])["syn_wav_list"]
    for r, w in zip(ref, got):
        assert snr_db(r.numpy(), w.cpu().numpy()) >= SNR_DB


@pytest.mark.parametrize("name", ["small", "full"])
def test_residual_vq_forward_matches_reference_golden(name):
    """ResidualVQ.forward through the drop-in class: codes bit-exact vs the reference's (near-ties adjudicated)."""
    from moss_ttsd_b200.xy_tokenizer.model import ResidualVQ
    from oracle import rvq_np
    from oracle.codec_weights import make_rvq_weights
    g = gold("rvq.npz")
    B, T, din, D, K, nq = [int(v) for v in g[f"{name}_dims"]]
    w = make_rvq_weights(din, D, din, nq, K, seed=3)
    rvq = ResidualVQ(input_dim=din, rvq_dim=D, output_dim=din, num_quantizers=nq, codebook_size=K, codebook_dim=D)
    rvq.load({k: torch.from_numpy(v) for k, v in w.items()}, "", torch.device("cuda"))
    z = torch.from_numpy(g[f"{name}_z"]).cuda()
    zq, codes, losses, _, out_len = rvq(z, torch.from_numpy(g[f"{name}_lengths"]))
    assert codes.shape == (nq, B, T) and codes.dtype == torch.int64 and zq.shape == (B, din, T)
    got, want = codes.cpu().numpy().reshape(nq, -1), g[f"{name}_codes"].reshape(nq, -1)
    cbs = np.stack([w[f"quantizers.{i}.codebook"] for i in range(nq)])
    rows_ok = (got == want).all(0)
    if not rows_ok.all():
        # first differing layer of each row must be a near-tie of the fp64 distances on the reference's own residual
        lengths = g[f"{name}_lengths"]
        valid = (np.arange(T)[None, :] < lengths[:, None]).reshape(-1)
        tok = g[f"{name}_z_in"].transpose(0, 2, 1).reshape(B * T, D)
        _, _, _, layer_in = rvq_np.rvq_forward(tok, cbs, valid)
        for r in np.nonzero(~rows_ok)[0]:
            i = int(np.nonzero(got[:, r] != want[:, r])[0][0])
            d = rvq_np.vq_dist64(layer_in[i][r:r + 1], cbs[i])[0]
            assert abs(d[got[i, r]] - d[want[i, r]]) <= 1e-6 * max(1.0, abs(d[want[i, r]]))
    assert rows_ok.mean() >= 0.95
    # quantized_out where the codes agree: the output projection runs in TF32 -> tolerance, not bit-exactness
    ref = g[f"{name}_zq_out"]
    err = np.abs(zq.cpu().numpy() - ref).transpose(0, 2, 1).reshape(B * T, -1)[rows_ok].max()
    assert err <= 5e-3 * max(1.0, np.abs(ref).max()), err
    dec = rvq.decode_codes(torch.from_numpy(g[f"{name}_codes"]).cuda())
    assert np.abs(dec.cpu().numpy() - g[f"{name}_decode"]).max() <= 5e-3 * max(1.0, np.abs(g[f"{name}_decode"]).max())


def test_codec_kernels_against_torch_functional():
    import torch.nn.functional as F
    from moss_ttsd_b200 import _lib, ops
    ops.ensure_init()
    L = _lib.load()
    sp = _lib.stream_ptr
    g = torch.Generator(device="cuda").manual_seed(0)
    # layernorm with length masking
    x = torch.randn(2 * 50, 768, device="cuda", generator=g)
    w, b = torch.randn(768, device="cuda", generator=g), torch.randn(768, device="cuda", generator=g)
    out = torch.empty_like(x)
    lens = torch.tensor([50, 20], dtype=torch.int32, device="cuda")
    _lib.check(L.mtts_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), 100, 768, 1e-5, lens.data_ptr(), 50, sp()))
    ref = F.layer_norm(x, (768,), w, b, 1e-5)
    ref[70:] = 0
    assert (out - ref).abs().max().item() <= 2e-5
    # attention
    B, T, H = 2, 150, 3
    qkv = torch.randn(B * T, 3 * H * 64, device="cuda", generator=g)
    ao = torch.empty(B * T, H * 64, device="cuda")
    lens = torch.tensor([150, 77], dtype=torch.int32, device="cuda")
    _lib.check(L.mtts_mha_varlen(qkv.data_ptr(), ao.data_ptr(), lens.data_ptr(), B, T, H, 64, sp()))
    q, k, v = (t.view(B, T, H, 64).transpose(1, 2).double() for t in qkv.view(B, T, 3, H * 64).unbind(2))
    s = (q * 64 ** -0.5) @ k.transpose(-1, -2)
    km = torch.arange(T, device="cuda")[None, :] < lens[:, None]
    s = s.masked_fill(~km[:, None, None, :], float("-inf"))
    ref = (s.softmax(-1) @ v).transpose(1, 2).reshape(B, T, H * 64)
    got = ao.view(B, T, H * 64).double()
    # TF32 tensor-core attention (10-bit mantissa operands, fp32 accumulate / softmax)
    assert (got[0] - ref[0]).abs().max().item() <= 4e-3 * ref.abs().max().item()
    assert (got[1, :77] - ref[1, :77]).abs().max().item() <= 4e-3 * ref.abs().max().item()
    # fp16-operand kernel (same mantissa width) and the fp32 CUDA-core kernel of the exact encode mode
    for fn, tol in ((L.mtts_mha_varlen_f16, 4e-3), (L.mtts_mha_varlen_fp32, 2e-6)):
        ao2 = torch.full_like(ao, float("nan"))
        _lib.check(fn(qkv.data_ptr(), ao2.data_ptr(), lens.data_ptr(), B, T, H, 64, sp()))
        got2 = ao2.view(B, T, H * 64).double()
        assert (got2[0] - ref[0]).abs().max().item() <= tol * ref.abs().max().item()
        assert (got2[1, :77] - ref[1, :77]).abs().max().item() <= tol * ref.abs().max().item()
    # LayerNorm with fp16 output
    o16 = torch.empty(100, 768, dtype=torch.float16, device="cuda")
    x_ln = torch.randn(100, 768, device="cuda", generator=g)
    _lib.check(L.mtts_layernorm_f16(x_ln.data_ptr(), w.data_ptr(), b.data_ptr(), o16.data_ptr(), 100, 768, 1e-5, None, 0, sp()))
    assert torch.equal(o16, F.layer_norm(x_ln, (768,), w, b, 1e-5).half()) or \
        (o16.float() - F.layer_norm(x_ln, (768,), w, b, 1e-5)).abs().max().item() <= 4e-3
    # dwconv7 + LN
    Bc, Tc, C = 2, 40, 512
    x = torch.randn(Bc, Tc, C, device="cuda", generator=g)
    cw, cb = torch.randn(C, 7, device="cuda", generator=g) * 0.3, torch.randn(C, device="cuda", generator=g) * 0.1
    out = torch.empty_like(x)
    _lib.check(L.mtts_dwconv7_ln(x.data_ptr(), cw.data_ptr(), cb.data_ptr(), w[:C].contiguous().data_ptr(),
                                 b[:C].contiguous().data_ptr(), out.data_ptr(), Bc, Tc, C, 1e-6, sp()))
    y = F.conv1d(x.transpose(1, 2), cw[:, None, :], cb, padding=3, groups=C).transpose(1, 2)
    ref = F.layer_norm(y, (C,), w[:C], b[:C], 1e-6)
    assert (out - ref).abs().max().item() <= 1e-4
    # istft (spec -> basis GEMM -> overlap-add) against torch.fft.irfft + fold
    from oracle.codec_weights import TINY_CODEC
    spt = _spt(TINY_CODEC, 21)
    Bi, Ti, Fb = 2, 12, 481
    hx = torch.randn(Bi * Ti, 2 * Fb, device="cuda", generator=g)
    spec = torch.empty(Bi * Ti, spt.head_ld, device="cuda")
    _lib.check(L.mtts_istft_spec(hx.data_ptr(), hx.stride(0), spec.data_ptr(), spt.head_ld, Bi * Ti, Fb, sp()))
    frames = ops.gemm_simt(spec, spt.basis)
    wav = torch.empty(Bi, Ti * 240, device="cuda")
    _lib.check(L.mtts_istft_ola(frames.data_ptr(), spt.window.data_ptr(), wav.data_ptr(), Bi, Ti, 960, 240, sp()))
    o = hx.view(Bi, Ti, 2 * Fb).transpose(1, 2)
    mag, ph = o.chunk(2, dim=1)
    S = torch.clip(torch.exp(mag), max=1e2) * (torch.cos(ph) + 1j * torch.sin(ph))
    win = torch.hann_window(960, device="cuda")
    ifft = torch.fft.irfft(S, 960, dim=1, norm="backward") * win[None, :, None]
    size = (Ti - 1) * 240 + 960
    yy = F.fold(ifft, output_size=(1, size), kernel_size=(1, 960), stride=(1, 240))[:, 0, 0, 360:-360]
    env = F.fold(win.square().expand(1, Ti, -1).transpose(1, 2), output_size=(1, size), kernel_size=(1, 960), stride=(1, 240)).squeeze()[360:-360]
    ref = yy / env
    assert (wav - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()
    # the same head as ONE C-ABI call (mtts_istft_head: projection + the three stages above, TF32 GEMMs) against
    # torch: fp32 linear -> irfft -> fold
    Cv = spt.head_w.shape[1]
    xh = torch.randn(Bi * Ti, Cv, device="cuda", generator=g) * 0.5
    ws = torch.empty(L.mtts_istft_head_workspace_bytes(Bi, Ti, 960, spt.head_ld), dtype=torch.uint8, device="cuda")
    wav1 = torch.full((Bi, Ti * 240), float("nan"), device="cuda")
    _lib.check(L.mtts_istft_head(xh.data_ptr(), xh.stride(0), Cv, spt.head_w.data_ptr(), spt.head_w.stride(0),
                                 spt.head_b.data_ptr(), spt.basis.data_ptr(), spt.head_ld, spt.window.data_ptr(),
                                 wav1.data_ptr(), Bi, Ti, 960, 240, ws.data_ptr(), ws.numel(), sp()))
    o = (xh.double() @ spt.head_w.double().t() + spt.head_b.double()).float().view(Bi, Ti, 2 * Fb).transpose(1, 2)
    mag, ph = o.chunk(2, dim=1)
    S = torch.clip(torch.exp(mag), max=1e2) * (torch.cos(ph) + 1j * torch.sin(ph))
    ifft = torch.fft.irfft(S, 960, dim=1, norm="backward") * win[None, :, None]
    yy = F.fold(ifft, output_size=(1, size), kernel_size=(1, 960), stride=(1, 240))[:, 0, 0, 360:-360]
    ref1 = yy / env
    assert torch.isfinite(wav1).all()
    assert (wav1 - ref1).abs().max().item() <= 5e-3 * ref1.abs().max().item()
    # too-small workspace and an empty batch: an error code / a no-op, never a launch
    assert L.mtts_istft_head(xh.data_ptr(), xh.stride(0), Cv, spt.head_w.data_ptr(), spt.head_w.stride(0),
                             spt.head_b.data_ptr(), spt.basis.data_ptr(), spt.head_ld, spt.window.data_ptr(),
                             wav1.data_ptr(), Bi, Ti, 960, 240, ws.data_ptr(), 16, sp()) != 0
    assert L.mtts_istft_head(None, 0, Cv, None, 0, None, None, spt.head_ld, None, None, 0, Ti, 960, 240, None, 0, sp()) == 0


def _spt_with_encoder(gp, seed):
    from moss_ttsd_b200.xy_tokenizer.model import XY_Tokenizer
    from oracle.codec_weights import make_codec_weights, make_encoder_weights
    sd = make_codec_weights(gp, seed)
    sd.update(make_encoder_weights(gp, seed + 7))
    spt = XY_Tokenizer(gp)
    spt.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return spt.to("cuda")


def _adjudicate(codes_ours, codes_ref, zin_ours, zin_ref, codebooks):
    """Every frame whose codes differ from the reference's must be EXPLAINED: at the first RVQ layer that differs (all
    earlier layers equal, so both sides quantise the same residual up to the feature difference delta = zin_ours -
    zin_ref), the fp64 distances evaluated on the REFERENCE's vector may favour the reference's code by at most
    2 |delta . (C[ours] - C[ref])| (+ fp32 rounding of the distance): that is exactly how far a perturbation of the
    input by delta can move the comparison. Returns (frames, frames that differ, flipped decisions with a relative gap
    <= 1e-6 [north_star's tie], max relative gap)."""
    nq, N = codes_ref.shape
    cb = codebooks.astype(np.float64)
    differ = ties = 0
    worst = 0.0
    for t in range(N):
        if (codes_ours[:, t] == codes_ref[:, t]).all():
            continue
        differ += 1
        i = int(np.argmax(codes_ours[:, t] != codes_ref[:, t]))
        e = zin_ref[t].astype(np.float64)
        for j in range(i):
            e = e - cb[j, codes_ref[j, t]]
        co, cr = cb[i, codes_ours[i, t]], cb[i, codes_ref[i, t]]
        d_o, d_r = ((e - co) ** 2).sum(), ((e - cr) ** 2).sum()
        delta = zin_ours[t].astype(np.float64) - zin_ref[t].astype(np.float64)
        slack = 2.0 * abs(np.dot(delta, co - cr)) + 4.0 * np.finfo(np.float32).eps * max(1.0, d_r)
        gap = d_o - d_r
        assert gap <= slack, (t, i, gap, slack)
        rel = gap / max(1.0, abs(d_r))
        worst = max(worst, rel)
        ties += rel <= 1e-6
    return N, differ, ties, worst


def test_encode_matches_reference_golden():
    """XY_Tokenizer.encode (log-mel, two encoders, adapters, gated down-conv, RVQ), tiny config, vs the reference's encode.
    Exact mode (default): 3xTF32 GEMMs / fp32 attention -> the projected pre-RVQ vectors agree with the reference's to
    fp32 summation-order noise, and every code that differs is an fp64-adjudicated near-tie explained by that noise.
    TF32 mode (`encode_exact = False`): agreement RATE only, reported."""
    from oracle.codec_weights import TINY_CODEC
    from tests.test_codec_encode_common import make_signals
    g = gold("codec_encode.npz")
    spt = _spt_with_encoder(TINY_CODEC, int(g["seed"]))
    assert spt.encode_exact
    wavs = [torch.from_numpy(w).cuda() for w in make_signals()]
    two = torch.stack([wavs[0][:480000], torch.nn.functional.pad(wavs[1], (0, 480000 - 48000))])
    mel = spt.log_mel(two)
    mel = mel.view(2, 3000, 80).permute(0, 2, 1).cpu().numpy()
    assert np.abs(mel[:, :, ::25] - g["mel_sub"]).max() <= 2e-3
    # ---- one chunk, exact mode: features, then codes with adjudication
    tok = spt.inference_tokenize(two[:, None], torch.tensor([480000, 48000]))
    assert tok["codes"].shape == (8, 2, 375) and tok["codes_lengths"].tolist() == g["chunk_code_lens"].tolist()
    zin_ref = np.transpose(g["chunk_zin"], (0, 2, 1)).reshape(2 * 375, -1)              # (B*T, 64)
    zin = spt.quantizer._last_zin.cpu().numpy()
    pre = spt._last_pre_rvq.view(2, 375, -1)[:, :, ::8].permute(0, 2, 1).cpu().numpy()
    lens = g["chunk_code_lens"].tolist()
    valid = np.concatenate([np.arange(375) < n for n in lens])
    feat_err = np.abs(pre - g["chunk_pre_sub"])[:, :, :lens[1]].max() / np.abs(g["chunk_pre_sub"]).max()
    zin_err = np.abs(zin - zin_ref)[valid].max() / np.abs(zin_ref).max()
    print(f"exact mode: pre-RVQ feature error {feat_err:.2e}, projected-vector error {zin_err:.2e} (relative to max)")
    assert feat_err <= 2e-5 and zin_err <= 2e-5
    cb = np.stack([spt.quantizer.codebooks[i].cpu().numpy() for i in range(8)])
    ours = tok["codes"].reshape(8, -1).cpu().numpy()[:, valid]
    ref = g["chunk_codes"].astype(np.int64).reshape(8, -1)[:, valid]
    N, differ, ties, worst = _adjudicate(ours, ref, zin[valid], zin_ref[valid], cb)
    print(f"exact mode: {N - differ}/{N} frames bit-identical, {differ} adjudicated near-ties (max relative gap {worst:.2e}, "
          f"{ties} within 1e-6)")
    assert differ <= 0.01 * N
    # ---- the public API with chunking (35 s and 3 s items), exact mode
    out = spt.encode(wavs, overlap_seconds=10)["codes_list"]
    same = tot = 0
    for i, c in enumerate(out):
        want = g[f"codes{i}"].astype(np.int64)
        assert tuple(c.shape) == want.shape and c.dtype == torch.int64
        eq = (c.cpu().numpy() == want).all(0)
        same += int(eq.sum())
        tot += eq.size
    print(f"exact mode, encode(): {same}/{tot} frames bit-identical over both items")
    assert same >= 0.99 * tot
    # ---- TF32 mode: reported agreement rate (first layer >= 95 %)
    spt.encode_exact = False
    out = spt.encode(wavs, overlap_seconds=10)["codes_list"]
    rates = np.mean([(c.cpu().numpy() == g[f"codes{i}"].astype(np.int64)).mean(1) for i, c in enumerate(out)], 0)
    print("TF32 mode: code agreement per RVQ layer:", np.round(rates, 3))
    assert rates[0] >= 0.95 and rates.mean() >= 0.80


def test_encode_shipped_config_bit_exact_modulo_adjudicated_ties():
    """The shipped xy_tokenizer_config.yaml (12-layer encoders, 4-layer adapters, 3072 -> 512 projection, 8 x 1024 x 512
    codebooks), one 9.6 s item, against the reference's own encode (tests/golden/codec_encode_full.npz)."""
    from oracle.codec_weights import full_codec_params
    from oracle.gen_golden_codec import encode_signal
    g = gold("codec_encode_full.npz")
    spt = _spt_with_encoder(full_codec_params(), int(g["seed"]))
    wav = torch.from_numpy(encode_signal(np.random.default_rng(int(g["sig_seed"])), int(g["n"]))).cuda()
    out = spt.encode([wav], overlap_seconds=10)["codes_list"][0]
    ref = g["codes"].astype(np.int64)
    assert tuple(out.shape) == ref.shape
    T = ref.shape[1]
    zin_ref = g["zin"].T                                                            # (T, 512)
    zin = spt.quantizer._last_zin.cpu().numpy()[:T]
    pre = spt._last_pre_rvq[:T, ::16].t().cpu().numpy()
    feat_err = np.abs(pre - g["pre_sub"]).max() / np.abs(g["pre_sub"]).max()
    zin_err = np.abs(zin - zin_ref).max() / np.abs(zin_ref).max()
    cb = np.stack([spt.quantizer.codebooks[i].cpu().numpy() for i in range(8)])
    N, differ, ties, worst = _adjudicate(out.cpu().numpy(), ref, zin, zin_ref, cb)
    print(f"shipped config, exact mode: feature error {feat_err:.2e}, projected {zin_err:.2e}; {N - differ}/{N} frames "
          f"bit-identical, {differ} adjudicated (max relative gap {worst:.2e}, {ties} within 1e-6)")
    assert feat_err <= 2e-4 and zin_err <= 2e-4  # 20 transformer layers of fp32 summation-order noise
    assert differ <= 0.03 * N
    spt.encode_exact = False
    fast = spt.encode([wav], overlap_seconds=10)["codes_list"][0].cpu().numpy()
    rates = (fast == ref).mean(1)
    print("shipped config, TF32 mode: code agreement per RVQ layer:", np.round(rates, 3))


def test_encode_then_decode_round_trip_shapes():
    from oracle.codec_weights import TINY_CODEC
    spt = _spt_with_encoder(TINY_CODEC, 33)
    wav = torch.randn(16000 * 4, device="cuda") * 0.1
    codes = spt.encode([wav])["codes_list"]
    assert codes[0].shape == (8, 50)
    rec = spt.decode(codes)["syn_wav_list"][0]
    assert rec.shape == (50 * 1920,) and torch.isfinite(rec).all()
    assert spt.encode([])["codes_list"] == []


@pytest.mark.parametrize("B,T,H,lens", [(3, 300, 12, [300, 117, 0]), (2, 1500, 12, [1500, 1337]), (4, 64, 2, [64, 1, 33, 64]),
                                        (1, 129, 12, [129])])
def test_mha_varlen_tcgen05_matches_fp32_softmax(B, T, H, lens):
    """tcgen05 attention (S and O in TMEM, fp16 operands, fp16 in/out) against an fp64 softmax(QK^T/8)V of the same fp16
    inputs: padded batch, ragged lengths incl. an all-masked item (uniform attention, as the reference computes it), query
    tiles and key tiles that end inside an item."""
    from moss_ttsd_b200 import _lib
    L = _lib.load()
    _lib.check(L.mtts_init())
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + T)
    E = H * 64
    qkv = (torch.randn(B * T, 3 * E, device="cuda", generator=g) * 1.5).half()
    out = torch.full((B * T, E), float("nan"), device="cuda", dtype=torch.float16)
    lengths = torch.tensor(lens, dtype=torch.int32, device="cuda")
    _lib.check(L.mtts_mha_varlen_tc(qkv.data_ptr(), out.data_ptr(), lengths.data_ptr(), B, T, H, 64, _lib.stream_ptr()))
    torch.cuda.synchronize()
    q, k, v = qkv.double().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)          # (B, H, T, 64)
    s = q @ k.transpose(-1, -2) / 8.0
    for b, n in enumerate(lens):
        if 0 < n < T:
            s[b, :, :, n:] = float("-inf")
    ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * T, E)
    got = out.double()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    assert err <= 6e-3, err      # fp16 probabilities (2^-11 relative) and fp16 output rounding on values of magnitude <= ~4
