"""TEST INFRASTRUCTURE. Deterministic (numpy PCG64) codec weights under the reference's state-dict key names, so that
the build container (which runs the reference) and the GPU box (which does not have it) construct identical models
without shipping weights as fixtures."""
import numpy as np


def make_rvq_weights(input_dim, rvq_dim, output_dim, num_quantizers, codebook_size, seed):
    """Keys of XY_Tokenizer/xy_tokenizer/nn/quantizer.py ResidualVQ (old-style weight_norm: weight_g / weight_v)."""
    rng = np.random.default_rng(seed)
    f = np.float32
    sd = {}
    cbs = (0.1 * rng.standard_normal((num_quantizers, codebook_size, rvq_dim))).astype(f)
    for i in range(num_quantizers):
        sd[f"quantizers.{i}.codebook"] = cbs[i]
    if input_dim != rvq_dim:
        sd["input_proj.weight_v"] = (rng.standard_normal((rvq_dim, input_dim, 1)) / np.sqrt(input_dim)).astype(f)
        sd["input_proj.weight_g"] = (1.0 + 0.1 * rng.standard_normal((rvq_dim, 1, 1))).astype(f)
        sd["input_proj.bias"] = (0.01 * rng.standard_normal(rvq_dim)).astype(f)
    if rvq_dim != output_dim:
        sd["output_proj.weight_v"] = (rng.standard_normal((output_dim, rvq_dim, 1)) / np.sqrt(rvq_dim)).astype(f)
        sd["output_proj.weight_g"] = (1.0 + 0.1 * rng.standard_normal((output_dim, 1, 1))).astype(f)
        sd["output_proj.bias"] = (0.01 * rng.standard_normal(output_dim)).astype(f)
    return sd


def weight_norm_weight(v: np.ndarray, g: np.ndarray) -> np.ndarray:
    """torch.nn.utils.weight_norm (dim=0): w = g * v / ||v|| with the norm over all dims but 0, fp32."""
    n = np.sqrt((v.astype(np.float32) ** 2).reshape(v.shape[0], -1).sum(1, dtype=np.float32)).reshape(-1, 1, 1)
    return (v * (g / n)).astype(np.float32)


# ------------------------------------------------------------------------------------------------ full decoder side
def _transformer_layer_specs(prefix, d, ffn):
    s = []
    for nm in ("self_attn_layer_norm", "final_layer_norm"):
        s += [(f"{prefix}{nm}.weight", (d,), "ln_w"), (f"{prefix}{nm}.bias", (d,), "bias")]
    s += [(f"{prefix}self_attn.k_proj.weight", (d, d), "lin"),
          (f"{prefix}self_attn.v_proj.weight", (d, d), "lin"), (f"{prefix}self_attn.v_proj.bias", (d,), "bias"),
          (f"{prefix}self_attn.q_proj.weight", (d, d), "lin"), (f"{prefix}self_attn.q_proj.bias", (d,), "bias"),
          (f"{prefix}self_attn.out_proj.weight", (d, d), "lin"), (f"{prefix}self_attn.out_proj.bias", (d,), "bias"),
          (f"{prefix}fc1.weight", (ffn, d), "lin"), (f"{prefix}fc1.bias", (ffn,), "bias"),
          (f"{prefix}fc2.weight", (d, ffn), "lin"), (f"{prefix}fc2.bias", (d,), "bias")]
    return s


def decoder_param_specs(gp):
    """(name, shape, kind) of every decode-side parameter of XY_Tokenizer, in a fixed order, under the reference's
    state-dict names (XY_Tokenizer/xy_tokenizer/model.py:40-49 and nn/modules.py)."""
    specs = []
    pk = gp["post_rvq_adapter_kwargs"]
    d = pk["d_model"]
    specs += [("post_rvq_adapter.proj.weight", (d, pk["input_dim"]), "lin"), ("post_rvq_adapter.proj.bias", (d,), "bias")]
    for l in range(pk["encoder_layers"]):
        specs += _transformer_layer_specs(f"post_rvq_adapter.layers.{l}.", d, pk["encoder_ffn_dim"])
    specs += [("post_rvq_adapter.layer_norm.weight", (d,), "ln_w"), ("post_rvq_adapter.layer_norm.bias", (d,), "bias"),
              ("post_rvq_adapter.out_proj.weight", (pk["output_dim"], d), "lin"),
              ("post_rvq_adapter.out_proj.bias", (pk["output_dim"],), "bias")]
    uk = gp["upsample_kwargs"]
    specs += [("upsample.up_conv.weight", (uk["stride"] * uk["d_model"], uk["d_model"], uk["stride"]), "conv_in")]
    ak = gp["acoustic_decoder_kwargs"]
    d = ak["d_model"]
    specs += [("acoustic_decoder.deconv1.weight", (d, d, ak["kernel_size"]), "conv_in"), ("acoustic_decoder.deconv1.bias", (d,), "bias"),
              ("acoustic_decoder.deconv2.weight", (d, ak["num_mel_bins"], ak["kernel_size"]), "conv_in"),
              ("acoustic_decoder.deconv2.bias", (ak["num_mel_bins"],), "bias")]
    for l in range(ak["decoder_layers"]):
        specs += _transformer_layer_specs(f"acoustic_decoder.layers.{l}.", d, ak["decoder_ffn_dim"])
    specs += [("acoustic_decoder.layer_norm.weight", (d,), "ln_w"), ("acoustic_decoder.layer_norm.bias", (d,), "bias")]
    vk = gp["vocos_kwargs"]
    dim, inter = vk["dim"], vk["intermediate_dim"]
    p = "enhanced_vocos.backbone."
    specs += [(p + "embed.weight", (dim, vk["input_channels"], 7), "conv_out"), (p + "embed.bias", (dim,), "bias"),
              (p + "norm.weight", (dim,), "ln_w"), (p + "norm.bias", (dim,), "bias")]
    for i in range(vk["num_layers"]):
        q = f"{p}convnext.{i}."
        specs += [(q + "gamma", (dim,), "gamma"), (q + "dwconv.weight", (dim, 1, 7), "dw"), (q + "dwconv.bias", (dim,), "bias"),
                  (q + "norm.weight", (dim,), "ln_w"), (q + "norm.bias", (dim,), "bias"),
                  (q + "pwconv1.weight", (inter, dim), "lin"), (q + "pwconv1.bias", (inter,), "bias"),
                  (q + "pwconv2.weight", (dim, inter), "lin"), (q + "pwconv2.bias", (dim,), "bias")]
    specs += [(p + "final_layer_norm.weight", (dim,), "ln_w"), (p + "final_layer_norm.bias", (dim,), "bias"),
              ("enhanced_vocos.head.out.weight", (vk["n_fft"] + 2, dim), "lin"),
              ("enhanced_vocos.head.out.bias", (vk["n_fft"] + 2,), "bias")]
    return specs


def make_codec_weights(gp, seed):
    """All decode-side weights (incl. quantizer.*) as float32 numpy arrays keyed by the reference's names."""
    rng = np.random.default_rng(seed)
    f = np.float32
    sd = {}
    qk = gp["quantizer_kwargs"]
    for k, v in make_rvq_weights(qk["input_dim"], qk["rvq_dim"], qk["output_dim"], qk["num_quantizers"],
                                 qk["codebook_size"], seed + 1).items():
        sd["quantizer." + k] = v
    for name, shape, kind in decoder_param_specs(gp):
        if kind == "lin":
            w = rng.standard_normal(shape, dtype=f) * f(1.0 / np.sqrt(shape[1]))
        elif kind == "conv_in":      # ConvTranspose1d [Cin, Cout, k]
            w = rng.standard_normal(shape, dtype=f) * f(1.0 / np.sqrt(shape[0]))
        elif kind == "conv_out":     # Conv1d [Cout, Cin, k]
            w = rng.standard_normal(shape, dtype=f) * f(1.0 / np.sqrt(shape[1] * shape[2]))
        elif kind == "dw":
            w = rng.standard_normal(shape, dtype=f) * f(0.4)
        elif kind == "ln_w":
            w = (1.0 + 0.1 * rng.standard_normal(shape)).astype(f)
        elif kind == "gamma":
            w = (0.1 + 0.02 * rng.standard_normal(shape)).astype(f)
        else:
            w = (0.02 * rng.standard_normal(shape)).astype(f)
        sd[name] = w.astype(f)
    return sd


TINY_CODEC = dict(
    input_sample_rate=16000, output_sample_rate=24000,
    feature_extractor_kwargs=dict(chunk_length=30, feature_size=80, hop_length=160, n_fft=400, n_samples=480000,
                                  nb_max_frames=3000, padding_side="right", padding_value=0.0,
                                  return_attention_mask=False, sampling_rate=16000),
    semantic_encoder_kwargs=dict(num_mel_bins=80, sampling_rate=16000, hop_length=160, stride_size=2, kernel_size=3,
                                 d_model=128, scale_embedding=False, max_audio_seconds=30, encoder_layers=1,
                                 encoder_attention_heads=2, encoder_ffn_dim=256, activation_function="gelu"),
    semantic_encoder_adapter_kwargs=dict(input_dim=128, output_dim=128, d_model=128, max_source_positions=1500,
                                         encoder_layers=1, encoder_attention_heads=2, encoder_ffn_dim=256),
    acoustic_encoder_kwargs=dict(num_mel_bins=80, sampling_rate=16000, hop_length=160, stride_size=2, kernel_size=3,
                                 d_model=128, scale_embedding=False, max_audio_seconds=30, encoder_layers=1,
                                 encoder_attention_heads=2, encoder_ffn_dim=256, activation_function="gelu"),
    pre_rvq_adapter_kwargs=dict(input_dim=256, output_dim=128, d_model=128, max_source_positions=1500, encoder_layers=1,
                                encoder_attention_heads=2, encoder_ffn_dim=256),
    downsample_kwargs=dict(d_model=128, avg_pooler=4),
    quantizer_kwargs=dict(input_dim=512, rvq_dim=64, output_dim=512, num_quantizers=8, codebook_size=128,
                          codebook_dim=64, quantizer_dropout=0.0, commitment=1),
    post_rvq_adapter_kwargs=dict(input_dim=512, output_dim=512, d_model=128, max_source_positions=375, encoder_layers=2,
                                 encoder_attention_heads=2, encoder_ffn_dim=256),
    upsample_kwargs=dict(d_model=128, stride=4),
    acoustic_decoder_kwargs=dict(num_mel_bins=80, sampling_rate=16000, hop_length=160, stride_size=2, kernel_size=3,
                                 d_model=128, scale_embedding=False, max_audio_seconds=30, decoder_layers=2,
                                 decoder_attention_heads=2, decoder_ffn_dim=256, activation_function="gelu"),
    vocos_kwargs=dict(input_channels=80, dim=128, intermediate_dim=256, num_layers=3, n_fft=960, hop_size=240,
                      padding="same"),
)


def full_codec_params():
    """xy_tokenizer_config.yaml (XY_Tokenizer/config/xy_tokenizer_config.yaml) restated verbatim as a dict."""
    enc = dict(num_mel_bins=80, sampling_rate=16000, hop_length=160, stride_size=2, kernel_size=3, d_model=768,
               scale_embedding=False, max_audio_seconds=30, encoder_layers=12, encoder_attention_heads=12,
               encoder_ffn_dim=3072, activation_function="gelu")
    return dict(
        input_sample_rate=16000, output_sample_rate=24000,
        feature_extractor_kwargs=dict(TINY_CODEC["feature_extractor_kwargs"]),
        semantic_encoder_kwargs=dict(enc),
        semantic_encoder_adapter_kwargs=dict(input_dim=768, output_dim=768, d_model=768, max_source_positions=1500,
                                             encoder_layers=4, encoder_attention_heads=12, encoder_ffn_dim=3072),
        acoustic_encoder_kwargs=dict(enc),
        pre_rvq_adapter_kwargs=dict(input_dim=1536, output_dim=768, d_model=768, max_source_positions=1500,
                                    encoder_layers=4, encoder_attention_heads=12, encoder_ffn_dim=3072),
        downsample_kwargs=dict(d_model=768, avg_pooler=4),
        quantizer_kwargs=dict(input_dim=3072, rvq_dim=512, output_dim=3072, num_quantizers=8, codebook_size=1024,
                              codebook_dim=512, quantizer_dropout=0.0, commitment=1),
        post_rvq_adapter_kwargs=dict(input_dim=3072, output_dim=3072, d_model=768, max_source_positions=375,
                                     encoder_layers=4, encoder_attention_heads=12, encoder_ffn_dim=3072),
        upsample_kwargs=dict(d_model=768, stride=4),
        acoustic_decoder_kwargs=dict(num_mel_bins=80, sampling_rate=16000, hop_length=160, stride_size=2, kernel_size=3,
                                     d_model=768, scale_embedding=False, max_audio_seconds=30, decoder_layers=12,
                                     decoder_attention_heads=12, decoder_ffn_dim=3072, activation_function="gelu"),
        vocos_kwargs=dict(input_channels=80, dim=512, intermediate_dim=4096, num_layers=30, n_fft=960, hop_size=240,
                          padding="same"),
    )


# ------------------------------------------------------------------------------------------------ encode side
def encoder_param_specs(gp):
    """(name, shape, kind) of the encode-side parameters (semantic / acoustic OmniAudioEncoder, the two adapter
    Transformers, ResidualDownConv) under the reference's state-dict names (model.py:26-38, nn/modules.py:208-326,
    426-477, 519-640)."""
    specs = []
    for name in ("semantic_encoder", "acoustic_encoder"):
        kw = gp[f"{name}_kwargs"]
        d, k = kw["d_model"], kw["kernel_size"]
        specs += [(f"{name}.conv1.weight", (d, kw["num_mel_bins"], k), "conv_out"), (f"{name}.conv1.bias", (d,), "bias"),
                  (f"{name}.conv2.weight", (d, d, k), "conv_out"), (f"{name}.conv2.bias", (d,), "bias")]
        for l in range(kw["encoder_layers"]):
            specs += _transformer_layer_specs(f"{name}.layers.{l}.", d, kw["encoder_ffn_dim"])
        specs += [(f"{name}.layer_norm.weight", (d,), "ln_w"), (f"{name}.layer_norm.bias", (d,), "bias")]
    for name in ("semantic_encoder_adapter", "pre_rvq_adapter"):
        kw = gp[f"{name}_kwargs"]
        d = kw["d_model"]
        if kw["input_dim"] != d:
            specs += [(f"{name}.proj.weight", (d, kw["input_dim"]), "lin"), (f"{name}.proj.bias", (d,), "bias")]
        for l in range(kw["encoder_layers"]):
            specs += _transformer_layer_specs(f"{name}.layers.{l}.", d, kw["encoder_ffn_dim"])
        specs += [(f"{name}.layer_norm.weight", (d,), "ln_w"), (f"{name}.layer_norm.bias", (d,), "bias")]
        if kw["output_dim"] != d:
            specs += [(f"{name}.out_proj.weight", (kw["output_dim"], d), "lin"), (f"{name}.out_proj.bias", (kw["output_dim"],), "bias")]
    dk = gp["downsample_kwargs"]
    d, p = dk["d_model"], dk["avg_pooler"]
    specs += [("downsample.gate_proj.weight", (d * p, d, p), "conv_out"), ("downsample.up_proj.weight", (d * p, d, p), "conv_out"),
              ("downsample.down_proj.weight", (d * p, d * p), "lin"),
              ("downsample.layer_norm.weight", (d * p,), "ln_w"), ("downsample.layer_norm.bias", (d * p,), "bias")]
    return specs


def make_encoder_weights(gp, seed):
    rng = np.random.default_rng(seed)
    f = np.float32
    sd = {}
    for name, shape, kind in encoder_param_specs(gp):
        if kind == "lin":
            w = rng.standard_normal(shape, dtype=f) * f(1.0 / np.sqrt(shape[1]))
        elif kind == "conv_out":
            w = rng.standard_normal(shape, dtype=f) * f(1.0 / np.sqrt(shape[1] * shape[2]))
        elif kind == "ln_w":
            w = (1.0 + 0.1 * rng.standard_normal(shape)).astype(f)
        else:
            w = (0.02 * rng.standard_normal(shape)).astype(f)
        sd[name] = w.astype(f)
    return sd
