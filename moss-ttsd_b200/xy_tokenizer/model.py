"""Drop-in mirror of the reference codec class `XY_Tokenizer` (XY_Tokenizer/xy_tokenizer/model.py) for the
generation hot path: `decode` / `inference_detokenize` (model.py:103-128,194-256) and the ResidualVQ
(`quantizer.forward`, `quantizer.decode_codes`, nn/quantizer.py:244-364) run on libmtts CUDA kernels.

Same constructor (`generator_params` dict from xy_tokenizer_config.yaml), same attributes
(input_sample_rate, output_sample_rate, encoder_downsample_rate, decoder_upsample_rate, nq), same
`load_from_checkpoint(config_path, ckpt_path)` and state-dict key names (incl. old-style weight_norm
`weight_g` / `weight_v` of quantizer.input_proj / output_proj).

HBM layout: every activation is token-major [batch*frames, channels] fp32; ConvTranspose1d / Conv1d weights are
re-packed once at load time so that each of them is one mtts_gemm (TF32 tensor cores, fp32 accumulate).
"""
from __future__ import annotations

import os
import math
from typing import Dict, Optional

import torch
import yaml

from .. import _lib, ops
from .._lib import check, ptr, stream_ptr


def sinusoids(length, channels, max_timescale=10000):
    """Positional table exactly as the reference builds it (modules.py:25-31; numpy log, torch exp/sin/cos on host)."""
    import numpy as np
    assert channels % 2 == 0
    log_timescale_increment = np.log(max_timescale) / (channels // 2 - 1)
    inv_timescales = torch.exp(-log_timescale_increment * torch.arange(channels // 2))
    scaled_time = torch.arange(length)[:, np.newaxis] * inv_timescales[np.newaxis, :]
    return torch.cat([torch.sin(scaled_time), torch.cos(scaled_time)], dim=1)


def mel_filter_bank_slaney(num_frequency_bins: int, num_mel_filters: int, min_frequency: float, max_frequency: float,
                           sampling_rate: int):
    """Slaney-scale, Slaney-normalised triangular mel filters [num_frequency_bins, num_mel_filters] (float64), the table
    MelFeatureExtractor builds with transformers.audio_utils.mel_filter_bank(norm="slaney", mel_scale="slaney")
    (nn/feature_extractor.py:41-49)."""
    import numpy as np

    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        lin = 3.0 * f / 200.0
        logstep = 27.0 / np.log(6.4)
        return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) * logstep, lin)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        lin = 200.0 * m / 3.0
        logstep = np.log(6.4) / 27.0
        return np.where(m >= 15.0, 1000.0 * np.exp(logstep * (m - 15.0)), lin)

    mel_freqs = np.linspace(hz_to_mel(min_frequency), hz_to_mel(max_frequency), num_mel_filters + 2)
    filter_freqs = mel_to_hz(mel_freqs)
    fft_freqs = np.linspace(0, sampling_rate // 2, num_frequency_bins)
    diff = np.diff(filter_freqs)
    slopes = filter_freqs[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / diff[:-1]
    up = slopes[:, 2:] / diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    fb *= (2.0 / (filter_freqs[2:num_mel_filters + 2] - filter_freqs[:num_mel_filters]))[None, :]
    return fb


def _wn(v: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """torch.nn.utils.weight_norm (dim=0) weight: g * v / ||v|| (the reference recomputes this every forward)."""
    n = v.float().reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
    return v.float() * (g.float() / n)


class _TransformerStack:
    """Weights of `n` OmniWhisperTransformerLayer (modules.py:163-205) in GEMM layout."""

    def __init__(self, sd, prefix, n_layers, dev):
        f = lambda k: sd[prefix + k].to(dev, torch.float32).contiguous()
        self.layers = []
        for l in range(n_layers):
            p = f"layers.{l}."
            qw, kw, vw = f(p + "self_attn.q_proj.weight"), f(p + "self_attn.k_proj.weight"), f(p + "self_attn.v_proj.weight")
            qb, vb = f(p + "self_attn.q_proj.bias"), f(p + "self_attn.v_proj.bias")
            self.layers.append(dict(
                ln1_w=f(p + "self_attn_layer_norm.weight"), ln1_b=f(p + "self_attn_layer_norm.bias"),
                wqkv=torch.cat([qw, kw, vw], 0).contiguous(), bqkv=torch.cat([qb, torch.zeros_like(qb), vb]).contiguous(),
                wo=f(p + "self_attn.out_proj.weight"), bo=f(p + "self_attn.out_proj.bias"),
                ln2_w=f(p + "final_layer_norm.weight"), ln2_b=f(p + "final_layer_norm.bias"),
                fc1_w=f(p + "fc1.weight"), fc1_b=f(p + "fc1.bias"), fc2_w=f(p + "fc2.weight"), fc2_b=f(p + "fc2.bias")))
        self.ln_w, self.ln_b = f("layer_norm.weight"), f("layer_norm.bias")


class ResidualVQ:
    """Inference-side mirror of nn/quantizer.py ResidualVQ: `forward(z, input_length, n_quantizers=None)` returns the
    reference's 5-tuple (commit losses are zeros: training-only bookkeeping), `decode_codes(codes)`."""

    def __init__(self, input_dim, rvq_dim, output_dim, num_quantizers, codebook_size, codebook_dim, **_):
        assert rvq_dim == codebook_dim, "in/out_project of VectorQuantize are Identity in the shipped config"
        self.input_dim, self.rvq_dim, self.output_dim = input_dim, rvq_dim, output_dim
        self.num_quantizers, self.codebook_size = num_quantizers, codebook_size
        self.codebooks = None

    def load(self, sd: Dict[str, torch.Tensor], prefix: str, dev):
        f = lambda k: sd[prefix + k].to(dev, torch.float32)
        self.codebooks = torch.stack([f(f"quantizers.{i}.codebook") for i in range(self.num_quantizers)]).contiguous()
        self.norms = ops.rvq_codebook_norms(self.codebooks)
        if self.input_dim != self.rvq_dim:
            self.in_w = _wn(f("input_proj.weight_v"), f("input_proj.weight_g"))[:, :, 0].contiguous()
            self.in_b = f("input_proj.bias").contiguous()
        else:
            self.in_w = None
        if self.rvq_dim != self.output_dim:
            self.out_w = _wn(f("output_proj.weight_v"), f("output_proj.weight_g"))[:, :, 0].contiguous()
            self.out_b = f("output_proj.bias").contiguous()
        else:
            self.out_w = None
        self.dev = dev

    def decode_tokens(self, codes: torch.Tensor) -> torch.Tensor:
        """codes (nq, B, T) -> token-major (B*T, output_dim) fp32."""
        nq, B, T = codes.shape
        flat = codes.reshape(nq, B * T).contiguous()
        err = torch.zeros(1, dtype=torch.int32, device=codes.device)
        emb = ops.rvq_decode(flat, self.codebooks, err_flag=err)
        if self.out_w is not None:
            emb = ops.gemm(emb, self.out_w, bias=self.out_b)
        self._err = err
        return emb

    def decode_codes(self, codes: torch.Tensor) -> torch.Tensor:
        """(nq, B, T) int64 -> (B, D, T) fp32, as the reference returns it."""
        nq, B, T = codes.shape
        out = self.decode_tokens(codes.to(self.dev))
        if int(self._err.item()):
            raise IndexError("code index out of range in decode_codes")
        return out.view(B, T, -1).permute(0, 2, 1)

    def encode_tokens(self, zt: torch.Tensor, valid: torch.Tensor, n_quantizers: Optional[int] = None):
        """Token-major entry used by XY_Tokenizer.encode: zt (N, input_dim) fp32, valid (N,) bool -> codes (nq, N)."""
        nq = n_quantizers or self.num_quantizers
        if self.in_w is not None:
            zt = ops.gemm_simt(zt, self.in_w, bias=self.in_b)  # exact fp32: a TF32 product here would move codes
        self._last_zin = zt                                    # (N, rvq_dim) projected vectors: parity tests read them
        codes, _, _ = ops.rvq_encode(zt.contiguous(), self.codebooks[:nq].contiguous(), self.norms[:nq].contiguous(),
                                     valid=valid.contiguous(), want_zq=False)
        return codes

    def forward(self, z: torch.Tensor, input_length: torch.Tensor, n_quantizers: Optional[int] = None):
        """z (B, input_dim, T) fp32 channel-major, input_length (B,). Returns (quantized_out (B, output_dim, T),
        all_indices (nq, B, T) int64, all_commit_losses (nq,), all_quantized (nq, B, D, T) — not materialised, None —,
        output_length)."""
        z = z.to(self.dev, torch.float32)
        B, Cin, T = z.shape
        nq = n_quantizers or self.num_quantizers
        if self.in_w is not None:
            # exact-fp32 projection read in place from the channel-major tensor: a TF32 product here would move codes
            zt = ops.gemm_simt(z, self.in_w, bias=self.in_b, x_layout=(T, z.stride(0), z.stride(2), z.stride(1)), M=B * T)
        else:
            zt = z.permute(0, 2, 1).reshape(B * T, Cin).contiguous()
        lengths = input_length.to(self.dev)
        valid = (torch.arange(T, device=self.dev)[None, :] < lengths[:, None]).reshape(-1).contiguous()
        codes, zq, _ = ops.rvq_encode(zt, self.codebooks[:nq].contiguous(), self.norms[:nq].contiguous(), valid=valid)
        if self.out_w is not None:
            zq = ops.gemm(zq, self.out_w, bias=self.out_b)
        quantized_out = zq.view(B, T, -1).permute(0, 2, 1)
        return (quantized_out, codes.view(nq, B, T), torch.zeros(nq, device=self.dev), None, input_length)

    __call__ = forward


class XY_Tokenizer:
    def __init__(self, generator_params: dict):
        gp = generator_params
        self.params = gp
        self.input_sample_rate = gp["input_sample_rate"]
        self.output_sample_rate = gp["output_sample_rate"]
        self.encoder_downsample_rate = 1280
        self.decoder_upsample_rate = 1920
        self.code_dim = gp["quantizer_kwargs"]["input_dim"]
        self.nq = gp["quantizer_kwargs"]["num_quantizers"]
        self.quantizer = ResidualVQ(**gp["quantizer_kwargs"])
        self.device = None
        self._sd = None
        self._ready = False
        self.training = False
        # encode() produces INTEGER codes that must equal the reference's, whose matmuls are true fp32
        # (`matmul.allow_tf32 = False`): by default the encoder front end runs every GEMM as 3xTF32
        # (fp32-accurate products, fp32 accumulate), attention on fp32 CUDA cores and the exact erf GELU. Setting
        # `encode_exact = False` selects plain TF32 tensor-core GEMMs (2-3x faster; ~1e-3 feature noise flips the codes
        # whose two best candidates are closer than that — the rate is reported by tests/test_codec_gpu.py).
        self.encode_exact = True
        self._exact_w = {}
        # decode(): the large GEMMs of the decoder (transformer q/k/v + MLP, ConvNeXt point-wise convs: ~90 % of its
        # FLOPs) take fp16 operands with fp32 accumulation — the same 10-bit mantissa as the TF32 path they replace
        # ("tf32"), at twice the tensor rate and half the operand bytes; LayerNorm / GELU outputs and the weights are
        # well inside the fp16 range. Waveform parity is gated by SNR against the reference (tests/test_codec_gpu.py).
        self.decode_gemm = "f16"
        self._half_w = {}

    # ------------------------------------------------------------------ loading
    @classmethod
    def load_from_checkpoint(cls, config_path: str, ckpt_path: str):
        with open(config_path, "r") as f:
            config = yaml.safe_load(f)
        model = cls(config["generator_params"])
        checkpoint = torch.load(ckpt_path, map_location="cpu")
        model.load_state_dict(checkpoint["generator"] if "generator" in checkpoint else checkpoint)
        return model

    def load_state_dict(self, sd, strict=True):
        self._sd = {k: (v if isinstance(v, torch.Tensor) else torch.as_tensor(v)) for k, v in sd.items()}
        self._ready = False
        if self.device is not None:
            self._prepare()
        return self

    def param_shapes(self):
        """(name, shape, fan_in) of every decode-side tensor under the reference's state-dict names."""
        gp = self.params
        out = []

        def layer(p, d, ffn):
            for nm in ("self_attn_layer_norm", "final_layer_norm"):
                out.extend([(p + nm + ".weight", (d,), 0), (p + nm + ".bias", (d,), -1)])
            for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
                out.append((p + f"self_attn.{nm}.weight", (d, d), d))
                if nm != "k_proj":
                    out.append((p + f"self_attn.{nm}.bias", (d,), -1))
            out.extend([(p + "fc1.weight", (ffn, d), d), (p + "fc1.bias", (ffn,), -1),
                        (p + "fc2.weight", (d, ffn), ffn), (p + "fc2.bias", (d,), -1)])

        qk = gp["quantizer_kwargs"]
        for i in range(qk["num_quantizers"]):
            out.append((f"quantizer.quantizers.{i}.codebook", (qk["codebook_size"], qk["codebook_dim"]), -2))
        for nm, (o, i) in (("input_proj", (qk["rvq_dim"], qk["input_dim"])), ("output_proj", (qk["output_dim"], qk["rvq_dim"]))):
            out.extend([(f"quantizer.{nm}.weight_v", (o, i, 1), i), (f"quantizer.{nm}.weight_g", (o, 1, 1), 0),
                        (f"quantizer.{nm}.bias", (o,), -1)])
        pk = gp["post_rvq_adapter_kwargs"]
        d = pk["d_model"]
        out.extend([("post_rvq_adapter.proj.weight", (d, pk["input_dim"]), pk["input_dim"]), ("post_rvq_adapter.proj.bias", (d,), -1)])
        for l in range(pk["encoder_layers"]):
            layer(f"post_rvq_adapter.layers.{l}.", d, pk["encoder_ffn_dim"])
        out.extend([("post_rvq_adapter.layer_norm.weight", (d,), 0), ("post_rvq_adapter.layer_norm.bias", (d,), -1),
                    ("post_rvq_adapter.out_proj.weight", (pk["output_dim"], d), d), ("post_rvq_adapter.out_proj.bias", (pk["output_dim"],), -1)])
        uk = gp["upsample_kwargs"]
        out.append(("upsample.up_conv.weight", (uk["stride"] * uk["d_model"], uk["d_model"], uk["stride"]), uk["stride"] * uk["d_model"]))
        ak = gp["acoustic_decoder_kwargs"]
        d = ak["d_model"]
        out.extend([("acoustic_decoder.deconv1.weight", (d, d, ak["kernel_size"]), d), ("acoustic_decoder.deconv1.bias", (d,), -1),
                    ("acoustic_decoder.deconv2.weight", (d, ak["num_mel_bins"], ak["kernel_size"]), d),
                    ("acoustic_decoder.deconv2.bias", (ak["num_mel_bins"],), -1)])
        for l in range(ak["decoder_layers"]):
            layer(f"acoustic_decoder.layers.{l}.", d, ak["decoder_ffn_dim"])
        out.extend([("acoustic_decoder.layer_norm.weight", (d,), 0), ("acoustic_decoder.layer_norm.bias", (d,), -1)])
        vk = gp["vocos_kwargs"]
        dim, inter, p = vk["dim"], vk["intermediate_dim"], "enhanced_vocos.backbone."
        out.extend([(p + "embed.weight", (dim, vk["input_channels"], 7), 7 * vk["input_channels"]), (p + "embed.bias", (dim,), -1),
                    (p + "norm.weight", (dim,), 0), (p + "norm.bias", (dim,), -1)])
        for i in range(vk["num_layers"]):
            q = f"{p}convnext.{i}."
            out.extend([(q + "gamma", (dim,), -3), (q + "dwconv.weight", (dim, 1, 7), 7), (q + "dwconv.bias", (dim,), -1),
                        (q + "norm.weight", (dim,), 0), (q + "norm.bias", (dim,), -1),
                        (q + "pwconv1.weight", (inter, dim), dim), (q + "pwconv1.bias", (inter,), -1),
                        (q + "pwconv2.weight", (dim, inter), inter), (q + "pwconv2.bias", (dim,), -1)])
        out.extend([(p + "final_layer_norm.weight", (dim,), 0), (p + "final_layer_norm.bias", (dim,), -1),
                    ("enhanced_vocos.head.out.weight", (vk["n_fft"] + 2, dim), dim), ("enhanced_vocos.head.out.bias", (vk["n_fft"] + 2,), -1)])
        return out

    def encoder_param_shapes(self):
        """(name, shape, fan_in) of every encode-side tensor under the reference's state-dict names (model.py:26-38:
        semantic / acoustic OmniAudioEncoder, the two adapter Transformers, ResidualDownConv)."""
        gp = self.params
        out = []

        def layer(p, d, ffn):
            for nm in ("self_attn_layer_norm", "final_layer_norm"):
                out.extend([(p + nm + ".weight", (d,), 0), (p + nm + ".bias", (d,), -1)])
            for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
                out.append((p + f"self_attn.{nm}.weight", (d, d), d))
                if nm != "k_proj":
                    out.append((p + f"self_attn.{nm}.bias", (d,), -1))
            out.extend([(p + "fc1.weight", (ffn, d), d), (p + "fc1.bias", (ffn,), -1),
                        (p + "fc2.weight", (d, ffn), ffn), (p + "fc2.bias", (d,), -1)])

        for name in ("semantic_encoder", "acoustic_encoder"):
            kw = gp[f"{name}_kwargs"]
            d, k = kw["d_model"], kw["kernel_size"]
            out.extend([(f"{name}.conv1.weight", (d, kw["num_mel_bins"], k), kw["num_mel_bins"] * k), (f"{name}.conv1.bias", (d,), -1),
                        (f"{name}.conv2.weight", (d, d, k), d * k), (f"{name}.conv2.bias", (d,), -1)])
            for l in range(kw["encoder_layers"]):
                layer(f"{name}.layers.{l}.", d, kw["encoder_ffn_dim"])
            out.extend([(f"{name}.layer_norm.weight", (d,), 0), (f"{name}.layer_norm.bias", (d,), -1)])
        for name in ("semantic_encoder_adapter", "pre_rvq_adapter"):
            kw = gp[f"{name}_kwargs"]
            d = kw["d_model"]
            if kw["input_dim"] != d:
                out.extend([(f"{name}.proj.weight", (d, kw["input_dim"]), kw["input_dim"]), (f"{name}.proj.bias", (d,), -1)])
            for l in range(kw["encoder_layers"]):
                layer(f"{name}.layers.{l}.", d, kw["encoder_ffn_dim"])
            out.extend([(f"{name}.layer_norm.weight", (d,), 0), (f"{name}.layer_norm.bias", (d,), -1)])
            if kw["output_dim"] != d:
                out.extend([(f"{name}.out_proj.weight", (kw["output_dim"], d), d), (f"{name}.out_proj.bias", (kw["output_dim"],), -1)])
        dk = gp["downsample_kwargs"]
        d, pl = dk["d_model"], dk["avg_pooler"]
        out.extend([("downsample.gate_proj.weight", (d * pl, d, pl), d * pl), ("downsample.up_proj.weight", (d * pl, d, pl), d * pl),
                    ("downsample.down_proj.weight", (d * pl, d * pl), d * pl),
                    ("downsample.layer_norm.weight", (d * pl,), 0), ("downsample.layer_norm.bias", (d * pl,), -1)])
        return out

    def init_random_weights(self, seed=0, device="cuda", encoder=False):
        """Seeded random weights generated on the device (bench / smoke; no parity role): the decode side, plus the
        encode side with `encoder=True`."""
        dev = torch.device(device)
        g = torch.Generator(device=dev).manual_seed(seed)
        sd = {}
        for name, shape, fan in self.param_shapes() + (self.encoder_param_shapes() if encoder else []):
            t = torch.empty(shape, dtype=torch.float32, device=dev).normal_(0.0, 1.0, generator=g)
            if fan > 0:
                t *= fan ** -0.5
            elif fan == 0:
                t = 1.0 + 0.1 * t
            elif fan == -1:
                t *= 0.02
            elif fan == -2:
                t *= 0.1
            else:
                t = 0.1 + 0.02 * t
            sd[name] = t
        self.device = dev
        return self.load_state_dict(sd)

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("XY_Tokenizer (B200 path) runs on CUDA only; there is no CPU path")
        self.device = device
        if self._sd is not None:
            self._prepare()
        return self

    def cuda(self):
        return self.to(torch.device("cuda", torch.cuda.current_device()))

    def eval(self):
        return self

    def _prepare(self):
        sd, dev, gp = self._sd, self.device, self.params
        f = lambda k: sd[k].to(dev, torch.float32).contiguous()
        self._exact_w = {}
        self._half_w = {}
        self.quantizer.load(sd, "quantizer.", dev)
        # post-RVQ adapter (Transformer, modules.py:519-640)
        pk = gp["post_rvq_adapter_kwargs"]
        self.pr_in_w, self.pr_in_b = f("post_rvq_adapter.proj.weight"), f("post_rvq_adapter.proj.bias")
        self.pr_out_w, self.pr_out_b = f("post_rvq_adapter.out_proj.weight"), f("post_rvq_adapter.out_proj.bias")
        self.pr_stack = _TransformerStack(sd, "post_rvq_adapter.", pk["encoder_layers"], dev)
        self.pr_heads = pk["encoder_attention_heads"]
        self.pr_pos = sinusoids(pk["max_source_positions"], pk["d_model"]).to(dev, torch.float32).contiguous()
        # UpConv (modules.py:480-515): ConvTranspose1d(k = stride) == GEMM; rows ordered (tap, out_channel)
        w = f("upsample.up_conv.weight")  # [Cin, Cout, k]
        self.up_stride = gp["upsample_kwargs"]["stride"]
        self.up_w = w.permute(2, 1, 0).reshape(-1, w.shape[0]).contiguous()
        # acoustic decoder (modules.py:329-423)
        ak = gp["acoustic_decoder_kwargs"]
        self.ad_stack = _TransformerStack(sd, "acoustic_decoder.", ak["decoder_layers"], dev)
        self.ad_heads = ak["decoder_attention_heads"]
        max_pos = (ak["max_audio_seconds"] * ak["sampling_rate"] // ak["hop_length"]) // ak["stride_size"]
        self.ad_pos = sinusoids(max_pos, ak["d_model"]).to(dev, torch.float32).contiguous()
        w1, w2 = f("acoustic_decoder.deconv1.weight"), f("acoustic_decoder.deconv2.weight")
        self.dc1_w = w1.permute(2, 1, 0).reshape(-1, w1.shape[0]).contiguous()
        self.dc1_b = f("acoustic_decoder.deconv1.bias")
        self.dc2_w = w2.permute(2, 1, 0).reshape(-1, w2.shape[0]).contiguous()
        self.dc2_b = f("acoustic_decoder.deconv2.bias")
        self.dc_k, self.dc1_stride, self.mel_bins = ak["kernel_size"], ak["stride_size"], ak["num_mel_bins"]
        # Vocos (modules.py:1347-1479)
        vk = gp["vocos_kwargs"]
        we = f("enhanced_vocos.backbone.embed.weight")  # [dim, Cin, 7]
        self.v_k = we.shape[2]
        self.v_embed_ld = (self.v_k * we.shape[1] + 3) // 4 * 4
        wcol = torch.zeros((we.shape[0], self.v_embed_ld), dtype=torch.float32, device=dev)
        wcol[:, :self.v_k * we.shape[1]] = we.permute(0, 2, 1).reshape(we.shape[0], -1)
        self.v_embed_w, self.v_embed_b = wcol, f("enhanced_vocos.backbone.embed.bias")
        self.v_norm_w, self.v_norm_b = f("enhanced_vocos.backbone.norm.weight"), f("enhanced_vocos.backbone.norm.bias")
        self.v_blocks = []
        for i in range(vk["num_layers"]):
            p = f"enhanced_vocos.backbone.convnext.{i}."
            self.v_blocks.append(dict(dw_w=f(p + "dwconv.weight").reshape(-1, self.v_k).contiguous(), dw_b=f(p + "dwconv.bias"),
                                      ln_w=f(p + "norm.weight"), ln_b=f(p + "norm.bias"),
                                      pw1_w=f(p + "pwconv1.weight"), pw1_b=f(p + "pwconv1.bias"),
                                      pw2_w=f(p + "pwconv2.weight"), pw2_b=f(p + "pwconv2.bias"), gamma=f(p + "gamma")))
        self.v_fln_w = f("enhanced_vocos.backbone.final_layer_norm.weight")
        self.v_fln_b = f("enhanced_vocos.backbone.final_layer_norm.bias")
        self.v_dim = vk["dim"]
        self.n_fft, self.hop = vk["n_fft"], vk["hop_size"]
        F = self.n_fft // 2 + 1
        self.head_ld = (2 * F + 3) // 4 * 4
        self.head_w, self.head_b = f("enhanced_vocos.head.out.weight"), f("enhanced_vocos.head.out.bias")
        self.window = torch.hann_window(self.n_fft).to(dev, torch.float32).contiguous()
        # windowed inverse real DFT as a [n_fft, 2F(+pad)] matrix (irfft(n=n_fft, norm="backward") x hann, modules.py:765-766)
        n = torch.arange(self.n_fft, dtype=torch.float64)[:, None]
        k = torch.arange(F, dtype=torch.float64)[None, :]
        ck = torch.full((1, F), 2.0, dtype=torch.float64)
        ck[0, 0] = 1.0
        if self.n_fft % 2 == 0:
            ck[0, -1] = 1.0
        ang = 2.0 * math.pi * n * k / self.n_fft
        win = torch.hann_window(self.n_fft, dtype=torch.float64)[:, None]
        basis = torch.zeros((self.n_fft, self.head_ld), dtype=torch.float64)
        basis[:, :F] = win * ck * torch.cos(ang) / self.n_fft
        basis[:, F:2 * F] = -win * ck * torch.sin(ang) / self.n_fft
        self.basis = basis.to(dev, torch.float32).contiguous()
        ops.ensure_init()
        self.L = _lib.load()
        self._ready = True
        self._prepare_encoder()

    def _prepare_encoder(self):
        """Encode-side weights (semantic / acoustic encoders, adapters, gated down-conv); absent keys -> encode() unavailable."""
        sd, dev, gp = self._sd, self.device, self.params
        self._enc_ready = False
        if "semantic_encoder.conv1.weight" not in sd:
            return
        f = lambda k: sd[k].to(dev, torch.float32).contiguous()

        def conv_as_gemm(w):  # Conv1d [Cout, Cin, k] -> [Cout, k*Cin (padded to 4)]
            co, ci, k = w.shape
            ld = (k * ci + 3) // 4 * 4
            out = torch.zeros((co, ld), dtype=torch.float32, device=dev)
            out[:, :k * ci] = w.permute(0, 2, 1).reshape(co, -1)
            return out

        self.enc = {}
        for name in ("semantic_encoder", "acoustic_encoder"):
            kw = gp[f"{name}_kwargs"]
            max_pos = (kw["max_audio_seconds"] * kw["sampling_rate"] // kw["hop_length"]) // kw["stride_size"]
            self.enc[name] = dict(
                c1_w=conv_as_gemm(f(f"{name}.conv1.weight")), c1_b=f(f"{name}.conv1.bias"),
                c2_w=conv_as_gemm(f(f"{name}.conv2.weight")), c2_b=f(f"{name}.conv2.bias"),
                k=kw["kernel_size"], stride=kw["stride_size"], heads=kw["encoder_attention_heads"],
                stack=_TransformerStack(sd, f"{name}.", kw["encoder_layers"], dev),
                pos=sinusoids(max_pos, kw["d_model"]).to(dev, torch.float32).contiguous())
        for name in ("semantic_encoder_adapter", "pre_rvq_adapter"):
            kw = gp[f"{name}_kwargs"]
            self.enc[name] = dict(
                proj_w=f(f"{name}.proj.weight") if kw["input_dim"] != kw["d_model"] else None,
                proj_b=f(f"{name}.proj.bias") if kw["input_dim"] != kw["d_model"] else None,
                out_w=f(f"{name}.out_proj.weight") if kw["output_dim"] != kw["d_model"] else None,
                out_b=f(f"{name}.out_proj.bias") if kw["output_dim"] != kw["d_model"] else None,
                heads=kw["encoder_attention_heads"], stack=_TransformerStack(sd, f"{name}.", kw["encoder_layers"], dev),
                pos=sinusoids(kw["max_source_positions"], kw["d_model"]).to(dev, torch.float32).contiguous())
        dk = gp["downsample_kwargs"]
        gw, uw = f("downsample.gate_proj.weight"), f("downsample.up_proj.weight")  # [4D, D, 4]
        g2 = gw.permute(0, 2, 1).reshape(gw.shape[0], -1)
        u2 = uw.permute(0, 2, 1).reshape(uw.shape[0], -1)
        self.ds_gu_w = torch.stack([g2, u2], dim=1).reshape(2 * g2.shape[0], g2.shape[1]).contiguous()  # interleaved
        self.ds_down_w = f("downsample.down_proj.weight")
        self.ds_ln_w, self.ds_ln_b = f("downsample.layer_norm.weight"), f("downsample.layer_norm.bias")
        self.ds_pool = dk["avg_pooler"]
        # log-mel front end: Hann window, real-DFT matrix (cos | -sin), Slaney mel filters
        fk = gp["feature_extractor_kwargs"]
        self.fe_nfft, self.fe_hop, self.fe_mels = fk["n_fft"], fk["hop_length"], fk["feature_size"]
        self.fe_nsamples = fk["chunk_length"] * fk["sampling_rate"]
        F = self.fe_nfft // 2 + 1
        n = torch.arange(self.fe_nfft, dtype=torch.float64)[None, :]
        k = torch.arange(F, dtype=torch.float64)[:, None]
        ang = 2.0 * math.pi * k * n / self.fe_nfft
        self.fe_dft = torch.cat([torch.cos(ang), -torch.sin(ang)], 0).to(dev, torch.float32).contiguous()  # [2F, n_fft]
        self.fe_window = torch.hann_window(self.fe_nfft).to(dev, torch.float32).contiguous()
        self.fe_pld = (F + 3) // 4 * 4
        fb = torch.from_numpy(mel_filter_bank_slaney(F, self.fe_mels, 0.0, fk["sampling_rate"] / 2, fk["sampling_rate"]))
        melw = torch.zeros((self.fe_mels, self.fe_pld), dtype=torch.float32)
        melw[:, :F] = fb.t().to(torch.float32)
        self.fe_melw = melw.to(dev).contiguous()
        self._enc_ready = True

    # ------------------------------------------------------------------ kernels
    def _ln(self, x, w, b, eps=1e-5, lengths=None, rows_per_item=0, out=None):
        out = torch.empty_like(x) if out is None else out
        check(self.L.mtts_layernorm(ptr(x), ptr(w), ptr(b), ptr(out), x.shape[0], x.shape[1], eps, ptr(lengths),
                                    rows_per_item, stream_ptr()))
        return out

    def _mm(self, x, w, exact=False, **kw):
        """Dense layer: TF32 tensor-core GEMM, or (exact) the 3xTF32 product against the weight's cached hi/lo split."""
        if not exact:
            return ops.gemm(x, w, **kw)
        ew = self._exact_w.get(w.data_ptr())
        if ew is None or ew.w is not w:
            ew = ops.ExactWeight(w)
            self._exact_w[w.data_ptr()] = ew
        return ops.gemm_exact(x, ew, **kw)

    def _half(self, w):
        hw = self._half_w.get(w.data_ptr())
        if hw is None or hw[0] is not w:
            hw = (w, w.to(torch.float16).contiguous())
            self._half_w[w.data_ptr()] = hw
        return hw[1]

    def _ln16(self, x, w, b, eps=1e-5):
        out = torch.empty(x.shape, dtype=torch.float16, device=x.device)
        check(self.L.mtts_layernorm_f16(ptr(x), ptr(w), ptr(b), ptr(out), x.shape[0], x.shape[1], eps, None, 0, stream_ptr()))
        return out

    def _stack_f16(self, h, st: _TransformerStack, heads, lengths, B, T):
        """_stack with fp16 operands for every projection (q/k/v, attention output, MLP) and the tcgen05 attention kernel
        (fp16 q/k/v in, fp16 out); the residual stream stays fp32."""
        E = h.shape[1]
        tc_attn = E // heads == 64 and os.environ.get("MTTS_CODEC_ATTN_TC", "1") != "0"
        for lw in st.layers:
            xn = self._ln16(h, lw["ln1_w"], lw["ln1_b"])
            if tc_attn:
                qkv = ops.gemm(xn, self._half(lw["wqkv"]), bias=lw["bqkv"], out_dtype=torch.float16)
                ao = torch.empty((B * T, E), dtype=torch.float16, device=h.device)
                check(self.L.mtts_mha_varlen_tc(ptr(qkv), ptr(ao), ptr(lengths), B, T, heads, 64, stream_ptr()))
                ops.gemm(ao, self._half(lw["wo"]), bias=lw["bo"], residual=h, out=h)
            else:
                qkv = ops.gemm(xn, self._half(lw["wqkv"]), bias=lw["bqkv"], out_dtype=torch.float32)
                ao = torch.empty((B * T, E), dtype=torch.float32, device=h.device)
                check(self.L.mtts_mha_varlen_f16(ptr(qkv), ptr(ao), ptr(lengths), B, T, heads, E // heads, stream_ptr()))
                ops.gemm(ao, lw["wo"], bias=lw["bo"], residual=h, out=h)
            xn = self._ln16(h, lw["ln2_w"], lw["ln2_b"])
            ff = ops.gemm(xn, self._half(lw["fc1_w"]), bias=lw["fc1_b"], gelu=True, out_dtype=torch.float16)
            ops.gemm(ff, self._half(lw["fc2_w"]), bias=lw["fc2_b"], residual=h, out=h)
        return self._ln(h, st.ln_w, st.ln_b, lengths=lengths, rows_per_item=T)

    def _stack(self, h, st: _TransformerStack, heads, lengths, B, T, exact=False, half=False):
        if half and not exact:
            return self._stack_f16(h, st, heads, lengths, B, T)
        E = h.shape[1]
        mha = self.L.mtts_mha_varlen_fp32 if exact else self.L.mtts_mha_varlen
        for lw in st.layers:
            xn = self._ln(h, lw["ln1_w"], lw["ln1_b"])
            qkv = self._mm(xn, lw["wqkv"], exact, bias=lw["bqkv"])
            ao = torch.empty((B * T, E), dtype=torch.float32, device=h.device)
            check(mha(ptr(qkv), ptr(ao), ptr(lengths), B, T, heads, E // heads, stream_ptr()))
            self._mm(ao, lw["wo"], exact, bias=lw["bo"], residual=h, out=h)
            xn = self._ln(h, lw["ln2_w"], lw["ln2_b"], out=xn)
            ff = self._mm(xn, lw["fc1_w"], exact, bias=lw["fc1_b"], gelu=True)
            self._mm(ff, lw["fc2_w"], exact, bias=lw["fc2_b"], residual=h, out=h)
        return self._ln(h, st.ln_w, st.ln_b, lengths=lengths, rows_per_item=T)  # + zero rows beyond each length

    @torch.no_grad()
    def detokenize_tokens(self, codes: torch.Tensor, codes_lengths: torch.Tensor):
        """codes (nq, B, T) int64 on device, lengths (B,) -> waveform (B, T*1920) fp32."""
        if not self._ready:
            raise RuntimeError("XY_Tokenizer: weights not loaded / not moved to CUDA")
        nq, B, T = codes.shape
        dev = self.device
        L = self.L
        len1 = codes_lengths.to(dev, torch.int32).contiguous()
        z = self.quantizer.decode_tokens(codes)                                 # (B*T, 3072)
        # post-RVQ adapter
        h = ops.gemm(z, self.pr_in_w, bias=self.pr_in_b)                        # (B*T, 768)
        check(L.mtts_add_rows_mod(ptr(h), ptr(self.pr_pos), B * T, h.shape[1], T, stream_ptr()))
        half = self.decode_gemm == "f16"
        h = self._stack(h, self.pr_stack, self.pr_heads, len1, B, T, half=half)
        z = ops.gemm(h, self.pr_out_w, bias=self.pr_out_b)                      # (B*T, 3072)
        # upsample x4 (tap-major rows -> a plain reshape gives (B*4T, 768))
        s = self.up_stride
        h = ops.gemm(z, self.up_w).view(B * T * s, -1)
        T2 = T * s
        len2 = (len1 * s).contiguous()
        check(L.mtts_add_rows_mod(ptr(h), ptr(self.ad_pos), B * T2, h.shape[1], T2, stream_ptr()))
        h = self._stack(h, self.ad_stack, self.ad_heads, len2, B, T2, half=half)
        # deconv1 (k3 s2) + GELU, deconv2 (k3 s1) + GELU, trim to 2*T2
        K, st = self.dc_k, self.dc1_stride
        T3 = (T2 - 1) * st + K
        y = ops.gemm(h, self.dc1_w)                                             # (B*T2, K*768)
        h3 = torch.empty((B * T3, self.dc1_b.numel()), dtype=torch.float32, device=dev)
        check(L.mtts_convt_gather(ptr(y), ptr(self.dc1_b), ptr(h3), B, T2, self.dc1_b.numel(), K, st, T3, 1, stream_ptr()))
        y = ops.gemm(h3, self.dc2_w)                                            # (B*T3, K*80)
        T4 = min(T3 + K - 1, T2 * st)
        mel = torch.empty((B * T4, self.mel_bins), dtype=torch.float32, device=dev)
        check(L.mtts_convt_gather(ptr(y), ptr(self.dc2_b), ptr(mel), B, T3, self.mel_bins, K, 1, T4, 1, stream_ptr()))
        # Vocos backbone
        col = torch.empty((B * T4, self.v_embed_ld), dtype=torch.float32, device=dev)
        if self.v_embed_ld != self.v_k * self.mel_bins:
            col.zero_()
        check(L.mtts_im2col(ptr(mel), ptr(col), B, T4, self.mel_bins, self.v_k, self.v_embed_ld, 1, stream_ptr()))
        x = ops.gemm(col, self.v_embed_w, bias=self.v_embed_b)                  # (B*T4, 512)
        x = self._ln(x, self.v_norm_w, self.v_norm_b, eps=1e-6)
        t = torch.empty(x.shape, dtype=torch.float16 if half else torch.float32, device=dev)
        ff = None
        for bl in self.v_blocks:
            if half:
                check(L.mtts_dwconv7_ln_f16(ptr(x), ptr(bl["dw_w"]), ptr(bl["dw_b"]), ptr(bl["ln_w"]), ptr(bl["ln_b"]), ptr(t), B,
                                            T4, self.v_dim, 1e-6, stream_ptr()))
                ff = ops.gemm(t, self._half(bl["pw1_w"]), out=ff, bias=bl["pw1_b"], gelu=True, out_dtype=torch.float16)
                ops.gemm(ff, self._half(bl["pw2_w"]), bias=bl["pw2_b"], gamma=bl["gamma"], residual=x, out=x)
            else:
                check(L.mtts_dwconv7_ln(ptr(x), ptr(bl["dw_w"]), ptr(bl["dw_b"]), ptr(bl["ln_w"]), ptr(bl["ln_b"]), ptr(t), B, T4,
                                        self.v_dim, 1e-6, stream_ptr()))
                ff = ops.gemm(t, bl["pw1_w"], out=ff, bias=bl["pw1_b"], gelu=True)
                ops.gemm(ff, bl["pw2_w"], bias=bl["pw2_b"], gamma=bl["gamma"], residual=x, out=x)
        x = self._ln(x, self.v_fln_w, self.v_fln_b, eps=1e-6)
        # ISTFT head
        # head projection -> (log-mag | phase) -> (Re | Im) -> windowed inverse DFT (GEMM) -> overlap-add, one entry point
        ws = torch.empty(L.mtts_istft_head_workspace_bytes(B, T4, self.n_fft, self.head_ld), dtype=torch.uint8, device=dev)
        wav = torch.empty((B, T4 * self.hop), dtype=torch.float32, device=dev)
        check(L.mtts_istft_head(ptr(x), x.stride(0), x.shape[1], ptr(self.head_w), self.head_w.stride(0), ptr(self.head_b),
                                ptr(self.basis), self.head_ld, ptr(self.window), ptr(wav), B, T4, self.n_fft, self.hop,
                                ptr(ws), ws.numel(), stream_ptr()))
        return wav

    # ------------------------------------------------------------------ reference API
    @torch.inference_mode()
    def inference_detokenize(self, codes, codes_lengths):
        wav = self.detokenize_tokens(codes.to(self.device), codes_lengths)
        up = self.decoder_upsample_rate
        return {"y": wav[:, None, :], "output_length": codes_lengths.to(self.device) * up}

    @torch.inference_mode()
    def decode(self, codes_list, overlap_seconds=10, device=None):
        """B x (nq, T) int64 -> {"syn_wav_list": B x (T*1920,) fp32}; 30 s windows, (30 - overlap) s hop
        (model.py:194-256). Chunk bookkeeping is done on the host from the known lengths (no device syncs)."""
        device = torch.device(device) if device is not None else self.device
        if device.type != "cuda":
            raise RuntimeError("XY_Tokenizer.decode runs on CUDA only (no CPU path)")
        duration_seconds = 30 - overlap_seconds
        chunk_code_length = int(30 * self.input_sample_rate // self.encoder_downsample_rate)
        duration_code_length = int(duration_seconds * self.input_sample_rate // self.encoder_downsample_rate)
        duration_wav_length = duration_code_length * self.decoder_upsample_rate
        batch_size = len(codes_list)
        lens = [int(c.shape[-1]) for c in codes_list]
        max_code_length = max(lens) if lens else 0
        if batch_size and min(lens) == max_code_length:
            codes_tensor = torch.stack([c.to(device) for c in codes_list], dim=1).contiguous()   # (nq, B, T), one launch
        else:
            codes_tensor = torch.zeros(self.nq, batch_size, max_code_length, device=device, dtype=torch.long)
            for i, c in enumerate(codes_list):
                codes_tensor[:, i, :c.shape[-1]] = c.to(device)
        max_chunks = (max_code_length + duration_code_length - 1) // duration_code_length if duration_code_length > 0 else 0
        n_windows = sum(1 for k in range(max_chunks) if max(lens) - k * duration_code_length > 0)
        wav_tensor = torch.empty(batch_size, max(n_windows, 1) * duration_wav_length, device=device) if batch_size else None
        w_idx = 0
        for chunk_idx in range(max_chunks):
            start = chunk_idx * duration_code_length
            end = min(start + chunk_code_length, max_code_length)
            chunk_lens = [min(max(l - start, 0), end - start) for l in lens]
            if max(chunk_lens) == 0:
                continue
            chunk_lens_dev = torch.tensor(chunk_lens, dtype=torch.int32, device=device)
            wav = self.detokenize_tokens(codes_tensor[:, :, start:end].contiguous(), chunk_lens_dev)
            # keep each item's own samples of the hop (zeros beyond its length): one launch for the whole window
            keep = torch.clamp(chunk_lens_dev * self.decoder_upsample_rate, max=duration_wav_length).contiguous()
            dst = wav_tensor[:, w_idx * duration_wav_length:(w_idx + 1) * duration_wav_length]
            n = min(duration_wav_length, wav.shape[1])
            if n < duration_wav_length:
                dst.zero_()
            check(self.L.mtts_rows_prefix_copy(ptr(wav), wav.stride(0), ptr(dst), dst.stride(0), ptr(keep), batch_size, n, 4,
                                               stream_ptr()))
            w_idx += 1
        if w_idx:
            syn = [wav_tensor[i, :lens[i] * self.decoder_upsample_rate] for i in range(batch_size)]
        else:
            syn = [torch.zeros(0, device=device) for _ in range(batch_size)]
        return {"syn_wav_list": syn}

    # ------------------------------------------------------------------ encode side
    def _conv_gelu(self, x, w, b, B, T, cin, k, stride, exact=False):
        """Conv1d(k, pad=(k-1)/2, stride) + exact GELU on token-major x (B*T, cin) as im2col + GEMM."""
        Tout = (T + 2 * ((k - 1) // 2) - k) // stride + 1
        col = torch.empty((B * Tout, w.shape[1]), dtype=torch.float32, device=x.device)
        if w.shape[1] != k * cin:
            col.zero_()
        check(self.L.mtts_im2col(ptr(x), ptr(col), B, T, cin, k, w.shape[1], stride, stream_ptr()))
        return self._mm(col, w, exact, bias=b, gelu=True), Tout

    def _encoder(self, mel, mel_len, B, T, e, exact=False):
        h, T1 = self._conv_gelu(mel, e["c1_w"], e["c1_b"], B, T, mel.shape[1], e["k"], 1, exact)
        h, T2 = self._conv_gelu(h, e["c2_w"], e["c2_b"], B, T1, h.shape[1], e["k"], e["stride"], exact)
        lens = (mel_len // e["stride"]).to(torch.int32).contiguous()
        check(self.L.mtts_add_rows_mod(ptr(h), ptr(e["pos"]), B * T2, h.shape[1], T2, stream_ptr()))
        return self._stack(h, e["stack"], e["heads"], lens, B, T2, exact), lens, T2

    def _adapter(self, x, lens, B, T, a, exact=False):
        h = self._mm(x, a["proj_w"], exact, bias=a["proj_b"]) if a["proj_w"] is not None else x.clone()
        check(self.L.mtts_add_rows_mod(ptr(h), ptr(a["pos"]), B * T, h.shape[1], T, stream_ptr()))
        h = self._stack(h, a["stack"], a["heads"], lens, B, T, exact)
        return self._mm(h, a["out_w"], exact, bias=a["out_b"]) if a["out_w"] is not None else h

    @torch.no_grad()
    def log_mel(self, wav: torch.Tensor) -> torch.Tensor:
        """wav (B, n_samples) fp32 on device (zero-padded to the 30 s chunk) -> log-mel (B*T, n_mels) token-major."""
        B, Ls = wav.shape
        T = Ls // self.fe_hop
        L = self.L
        frames = torch.empty((B * T, self.fe_nfft), dtype=torch.float32, device=wav.device)
        check(L.mtts_stft_frames(ptr(wav), wav.stride(0), ptr(self.fe_window), ptr(frames), B, T, Ls, self.fe_nfft, self.fe_hop,
                                 stream_ptr()))
        spec = ops.gemm_simt(frames, self.fe_dft)                       # exact fp32 DFT: (B*T, 2F)
        F = self.fe_nfft // 2 + 1
        power = torch.empty((B * T, self.fe_pld), dtype=torch.float32, device=wav.device)
        check(L.mtts_power_spectrum(ptr(spec), spec.stride(0), ptr(power), self.fe_pld, B * T, F, stream_ptr()))
        mel = ops.gemm_simt(power, self.fe_melw)                        # (B*T, n_mels)
        check(L.mtts_logmel_finish(ptr(mel), B, T * self.fe_mels, stream_ptr()))
        return mel

    @torch.no_grad()
    def tokenize_tokens(self, x: torch.Tensor, input_lengths: torch.Tensor):
        """x (B, 1, T<=30 s) on device, lengths (B,) -> (codes (nq, B, T_code) int64, code lengths (B,))."""
        if not getattr(self, "_enc_ready", False):
            raise RuntimeError("XY_Tokenizer.encode: encoder-side weights are not loaded")
        dev = self.device
        B = x.shape[0]
        wav = torch.zeros((B, self.fe_nsamples), dtype=torch.float32, device=dev)
        Tn = min(x.shape[-1], self.fe_nsamples)
        wav[:, :Tn] = x[:, 0, :Tn]
        lens = torch.clamp(input_lengths.to(dev), max=self.fe_nsamples)
        # samples beyond each item's length are padding for the extractor (it is handed xi[:, :x_len], model.py:66)
        wav *= (torch.arange(self.fe_nsamples, device=dev)[None, :] < lens[:, None])
        mel = self.log_mel(wav)
        T = self.fe_nsamples // self.fe_hop
        mel_len = (lens + self.fe_hop - 1) // self.fe_hop                 # attention_mask[:, ::hop].sum()
        ex = bool(self.encode_exact)
        sem, l2, T2 = self._encoder(mel, mel_len, B, T, self.enc["semantic_encoder"], ex)
        sem = self._adapter(sem, l2, B, T2, self.enc["semantic_encoder_adapter"], ex)
        aco, _, _ = self._encoder(mel, mel_len, B, T, self.enc["acoustic_encoder"], ex)
        cat = torch.cat([sem, aco], dim=1).contiguous()                   # channel concat, token-major
        h = self._adapter(cat, l2, B, T2, self.enc["pre_rvq_adapter"], ex)
        # ResidualDownConv (modules.py:426-477): 4 frames -> 1; gate/up convs (k = stride = 4) are one interleaved GEMM
        # with the SwiGLU epilogue; down_proj + residual; LayerNorm
        pool = self.ds_pool
        T3 = T2 // pool
        x4 = h.view(B * T3, pool * h.shape[1])
        gu = self._mm(x4, self.ds_gu_w, ex, swiglu=True)
        c = self._mm(gu, self.ds_down_w, ex, residual=x4)
        z = self._ln(c, self.ds_ln_w, self.ds_ln_b)
        l3 = (l2 // pool).to(torch.int64)
        valid = (torch.arange(T3, device=dev)[None, :] < l3[:, None]).reshape(-1)
        self._last_pre_rvq = z                                            # (B*T3, input_dim): parity tests read it
        codes = self.quantizer.encode_tokens(z, valid)
        return codes.view(self.nq, B, T3), l3

    @torch.inference_mode()
    def inference_tokenize(self, x, input_lengths):
        codes, l3 = self.tokenize_tokens(x.to(self.device), input_lengths)
        return {"zq": None, "codes": codes, "codes_lengths": l3}

    @torch.inference_mode()
    def encode(self, wav_list, overlap_seconds=10, device=None):
        """B x (T,) waveforms at 16 kHz -> {"codes_list": B x (nq, T // 1280) int64}; 30 s windows with
        (30 - overlap) s hop (model.py:130-192). Chunk bookkeeping is host arithmetic on the known lengths."""
        device = torch.device(device) if device is not None else self.device
        if device.type != "cuda":
            raise RuntimeError("XY_Tokenizer.encode runs on CUDA only (no CPU path)")
        duration_seconds = 30 - overlap_seconds
        chunk_size = int(30 * self.input_sample_rate)
        duration_size = int(duration_seconds * self.input_sample_rate)
        code_duration_length = duration_size // self.encoder_downsample_rate
        lens = [int(len(w)) for w in wav_list]
        batch_size = len(wav_list)
        max_length = max(lens) if lens else 0
        wav_tensor = torch.zeros(batch_size, 1, max_length, device=device)
        for i, w in enumerate(wav_list):
            wav_tensor[i, 0, :lens[i]] = w.to(device)
        max_chunks = (max_length + duration_size - 1) // duration_size if duration_size > 0 else 0
        chunks = []
        for chunk_idx in range(max_chunks):
            start = chunk_idx * duration_size
            end = min(start + chunk_size, max_length)
            chunk_lens = [min(max(l - start, 0), end - start) for l in lens]
            if max(chunk_lens) == 0:
                continue
            codes, code_lens = self.tokenize_tokens(wav_tensor[:, :, start:end],
                                                    torch.tensor(chunk_lens, dtype=torch.int64, device=device))
            code_lens = code_lens.cpu().tolist()
            valid = torch.zeros(self.nq, batch_size, code_duration_length, device=device, dtype=torch.long)
            for b in range(batch_size):
                n = min(code_lens[b], code_duration_length)
                if n > 0:
                    valid[:, b, :n] = codes[:, b, :n]
            chunks.append(valid)
        if chunks:
            codes_tensor = torch.cat(chunks, dim=-1)
            codes_list = [codes_tensor[:, i, :lens[i] // self.encoder_downsample_rate] for i in range(batch_size)]
        else:
            codes_list = [torch.zeros(self.nq, 0, device=device, dtype=torch.long) for _ in range(batch_size)]
        return {"codes_list": codes_list}
