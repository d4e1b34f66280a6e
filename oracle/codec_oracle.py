"""Oracle (TEST INFRASTRUCTURE, never shipped): CPU restatement in plain torch of the XY_Tokenizer decode path.

Follows /root/reference/XY_Tokenizer/xy_tokenizer:
  decode / chunking          <- model.py:194-256
  inference_detokenize       <- model.py:103-128
  quantizer.decode_codes     <- nn/quantizer.py:345-364 (+ old-style weight_norm output_proj :224-225)
  transformer stack          <- nn/modules.py:58-205 (VarLenAttention + OmniWhisperTransformerLayer), :519-640 (Transformer)
  UpConv, OmniAudioDecoder   <- nn/modules.py:480-515, 329-423
  Vocos / ConvNeXt / ISTFT   <- nn/modules.py:1096-1154, 1347-1410, 1451-1479, 939-988, 709-792
Parity status: PINNED against the reference's own `decode` outputs (tests/golden/codec_*.npz from
oracle/gen_golden_codec.py), checked in tests/test_oracle_pin.py.
"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle.codec_weights import weight_norm_weight


def sinusoids(length, channels, max_timescale=10000):
    log_timescale_increment = np.log(max_timescale) / (channels // 2 - 1)
    inv_timescales = torch.exp(-log_timescale_increment * torch.arange(channels // 2))
    scaled_time = torch.arange(length)[:, np.newaxis] * inv_timescales[np.newaxis, :]
    return torch.cat([torch.sin(scaled_time), torch.cos(scaled_time)], dim=1)


class CodecOracle:
    def __init__(self, gp: dict, sd_np: dict, device="cpu"):
        """`device="cuda"`: the same eager torch ops on the GPU (bench.py's `gpu_eager_baseline` leg, decode side only)."""
        self.gp = gp
        self.dev = torch.device(device)
        cpu = {k: (v.detach().cpu().float() if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v)).float())
               for k, v in sd_np.items()}
        if "quantizer.output_proj.weight_v" in cpu:
            self.w_out = torch.from_numpy(weight_norm_weight(cpu["quantizer.output_proj.weight_v"].numpy(),
                                                             cpu["quantizer.output_proj.weight_g"].numpy())).to(self.dev)
        self.sd = {k: v.to(self.dev) for k, v in cpu.items()}
        self.nq = gp["quantizer_kwargs"]["num_quantizers"]

    # -------------------------------------------------------------- pieces
    def _attn_mask(self, seq_len, max_len):
        valid = torch.arange(max_len, device=self.dev)[None, :] < seq_len[:, None]
        m = (valid[:, None, :, None] & valid[:, None, None, :]).float()
        return m + (1.0 - m) * torch.finfo(torch.float32).min

    def _layer(self, h, p, heads, seq_len):
        sd = self.sd
        B, T, E = h.shape
        hd = E // heads
        res = h
        x = F.layer_norm(h, (E,), sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"])
        q = F.linear(x, sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"]) * (hd ** -0.5)
        k = F.linear(x, sd[p + "self_attn.k_proj.weight"])
        v = F.linear(x, sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"])
        q, k, v = (t.view(B, T, heads, hd).transpose(1, 2) for t in (q, k, v))
        s = torch.matmul(q, k.transpose(-1, -2)) + self._attn_mask(seq_len, T)
        a = torch.matmul(F.softmax(s, dim=-1), v).transpose(1, 2).contiguous().view(B, T, E)
        h = res + F.linear(a, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
        res = h
        x = F.layer_norm(h, (E,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"])
        x = F.linear(F.gelu(F.linear(x, sd[p + "fc1.weight"], sd[p + "fc1.bias"])), sd[p + "fc2.weight"], sd[p + "fc2.bias"])
        return res + x

    def _stack(self, h, prefix, n_layers, heads, seq_len, pos):
        B, T, E = h.shape
        h = h + (pos[:T] if T < pos.shape[0] else pos).to(h.device)
        for l in range(n_layers):
            h = self._layer(h, f"{prefix}layers.{l}.", heads, seq_len)
        h = F.layer_norm(h, (E,), self.sd[prefix + "layer_norm.weight"], self.sd[prefix + "layer_norm.bias"])
        mask = (torch.arange(T, device=self.dev)[None, :] < seq_len[:, None])[..., None]
        return torch.where(mask, h, torch.zeros((), dtype=h.dtype, device=self.dev))

    def decode_codes(self, codes):
        sd = self.sd
        nq, B, T = codes.shape
        emb = torch.zeros(B, T, sd["quantizer.quantizers.0.codebook"].shape[1], device=self.dev)
        for i in range(nq):
            emb += F.embedding(codes[i], sd[f"quantizer.quantizers.{i}.codebook"])
        return F.linear(emb, self.w_out[:, :, 0], sd["quantizer.output_proj.bias"])  # (B, T, 3072) token-major

    def detokenize(self, codes, lengths):
        gp, sd = self.gp, self.sd
        pk, ak, vk = gp["post_rvq_adapter_kwargs"], gp["acoustic_decoder_kwargs"], gp["vocos_kwargs"]
        lengths = lengths.long()
        z = self.decode_codes(codes)
        h = F.linear(z, sd["post_rvq_adapter.proj.weight"], sd["post_rvq_adapter.proj.bias"])
        h = self._stack(h, "post_rvq_adapter.", pk["encoder_layers"], pk["encoder_attention_heads"], lengths,
                        sinusoids(pk["max_source_positions"], pk["d_model"]))
        z = F.linear(h, sd["post_rvq_adapter.out_proj.weight"], sd["post_rvq_adapter.out_proj.bias"])
        s = gp["upsample_kwargs"]["stride"]
        u = F.conv_transpose1d(z.transpose(1, 2), sd["upsample.up_conv.weight"], stride=s)       # (B, 768, 4T)
        max_pos = (ak["max_audio_seconds"] * ak["sampling_rate"] // ak["hop_length"]) // ak["stride_size"]
        h = self._stack(u.transpose(1, 2), "acoustic_decoder.", ak["decoder_layers"], ak["decoder_attention_heads"],
                        lengths * s, sinusoids(max_pos, ak["d_model"]))
        tgt = h.shape[1]
        x = F.gelu(F.conv_transpose1d(h.permute(0, 2, 1), sd["acoustic_decoder.deconv1.weight"], sd["acoustic_decoder.deconv1.bias"],
                                      stride=ak["stride_size"]))
        x = F.gelu(F.conv_transpose1d(x, sd["acoustic_decoder.deconv2.weight"], sd["acoustic_decoder.deconv2.bias"], stride=1))
        x = x[:, :, :tgt * ak["stride_size"]]
        # Vocos
        p = "enhanced_vocos.backbone."
        dim = vk["dim"]
        x = F.conv1d(x, sd[p + "embed.weight"], sd[p + "embed.bias"], padding=3)
        x = F.layer_norm(x.transpose(1, 2), (dim,), sd[p + "norm.weight"], sd[p + "norm.bias"], eps=1e-6).transpose(1, 2)
        for i in range(vk["num_layers"]):
            q = f"{p}convnext.{i}."
            r = x
            y = F.conv1d(x, sd[q + "dwconv.weight"], sd[q + "dwconv.bias"], padding=3, groups=dim).transpose(1, 2)
            y = F.layer_norm(y, (dim,), sd[q + "norm.weight"], sd[q + "norm.bias"], eps=1e-6)
            y = F.linear(F.gelu(F.linear(y, sd[q + "pwconv1.weight"], sd[q + "pwconv1.bias"])), sd[q + "pwconv2.weight"], sd[q + "pwconv2.bias"])
            x = r + (sd[q + "gamma"] * y).transpose(1, 2)
        x = F.layer_norm(x.transpose(1, 2), (dim,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], eps=1e-6)
        # ISTFT head
        n_fft, hop = vk["n_fft"], vk["hop_size"]
        o = F.linear(x, sd["enhanced_vocos.head.out.weight"], sd["enhanced_vocos.head.out.bias"]).transpose(1, 2)
        mag, ph = o.chunk(2, dim=1)
        mag = torch.clip(torch.exp(mag), max=1e2)
        S = mag * (torch.cos(ph) + 1j * torch.sin(ph))
        win = torch.hann_window(n_fft, device=self.dev)
        pad = (n_fft - hop) // 2
        Bn, N, T = S.shape
        ifft = torch.fft.irfft(S, n_fft, dim=1, norm="backward") * win[None, :, None]
        out_size = (T - 1) * hop + n_fft
        y = F.fold(ifft, output_size=(1, out_size), kernel_size=(1, n_fft), stride=(1, hop))[:, 0, 0, pad:-pad]
        env = F.fold(win.square().expand(1, T, -1).transpose(1, 2), output_size=(1, out_size), kernel_size=(1, n_fft),
                     stride=(1, hop)).squeeze()[pad:-pad]
        return y / env

    def decode(self, codes_list, overlap_seconds=10):
        in_sr, down, up = self.gp["input_sample_rate"], 1280, 1920
        chunk_len = int(30 * in_sr // down)
        dur_len = int((30 - overlap_seconds) * in_sr // down)
        dur_wav = dur_len * up
        B = len(codes_list)
        lens = torch.tensor([c.shape[-1] for c in codes_list], device=self.dev)
        Tm = int(lens.max())
        codes = torch.zeros(self.nq, B, Tm, dtype=torch.long, device=self.dev)
        for i, c in enumerate(codes_list):
            codes[:, i, :c.shape[-1]] = c.to(self.dev)
        chunks = []
        for ci in range((Tm + dur_len - 1) // dur_len):
            start = ci * dur_len
            end = min(start + chunk_len, Tm)
            cl = torch.clamp(lens - start, 0, end - start)
            if cl.max() == 0:
                continue
            wav = self.detokenize(codes[:, :, start:end], cl)
            valid = torch.zeros(B, dur_wav, device=self.dev)
            cl_h = cl.tolist()
            for b in range(B):
                n = min(int(cl_h[b]) * up, dur_wav)
                if n > 0:
                    valid[b, :n] = wav[b, :n]
            chunks.append(valid)
        full = torch.cat(chunks, -1)
        return [full[i, :int(lens[i]) * up] for i in range(B)]


    # -------------------------------------------------------------- encode side (model.py:54-101,130-192)
    def log_mel(self, wav_list):
        """MelFeatureExtractor (nn/feature_extractor.py:78-104,128-237, torch path): pad to 30 s, torch.stft(400, 160, hann),
        power, Slaney mel, log10, per-item max-8 clamp, (x+4)/4. -> (B, 80, 3000), frame counts."""
        from transformers.audio_utils import mel_filter_bank
        fk = self.gp["feature_extractor_kwargs"]
        n_fft, hop, n_samples = fk["n_fft"], fk["hop_length"], fk["chunk_length"] * fk["sampling_rate"]
        B = len(wav_list)
        wav = torch.zeros(B, n_samples)
        lens = []
        for i, w in enumerate(wav_list):
            n = min(len(w), n_samples)
            wav[i, :n] = w[:n]
            lens.append(n)
        stft = torch.stft(wav, n_fft, hop, window=torch.hann_window(n_fft), return_complex=True)
        mag = stft[..., :-1].abs() ** 2
        filt = torch.from_numpy(mel_filter_bank(num_frequency_bins=1 + n_fft // 2, num_mel_filters=fk["feature_size"],
                                                min_frequency=0.0, max_frequency=fk["sampling_rate"] / 2,
                                                sampling_rate=fk["sampling_rate"], norm="slaney", mel_scale="slaney")).float()
        log_spec = torch.clamp(filt.T @ mag, min=1e-10).log10()
        mx = log_spec.max(dim=2, keepdim=True)[0].max(dim=1, keepdim=True)[0]
        log_spec = (torch.maximum(log_spec, mx - 8.0) + 4.0) / 4.0
        frames = torch.tensor([(n + hop - 1) // hop for n in lens])
        return log_spec, frames

    def _audio_encoder(self, mel, mel_len, name):
        sd, kw = self.sd, self.gp[f"{name}_kwargs"]
        x = F.gelu(F.conv1d(mel, sd[f"{name}.conv1.weight"], sd[f"{name}.conv1.bias"], padding=1))
        x = F.gelu(F.conv1d(x, sd[f"{name}.conv2.weight"], sd[f"{name}.conv2.bias"], stride=kw["stride_size"], padding=1))
        out_len = (mel_len // kw["stride_size"]).long()
        max_pos = (kw["max_audio_seconds"] * kw["sampling_rate"] // kw["hop_length"]) // kw["stride_size"]
        h = self._stack(x.permute(0, 2, 1), f"{name}.", kw["encoder_layers"], kw["encoder_attention_heads"], out_len,
                        sinusoids(max_pos, kw["d_model"]))
        return h, out_len          # (B, T, D) token-major

    def _transformer(self, x, lens, name):
        sd, kw = self.sd, self.gp[f"{name}_kwargs"]
        if kw["input_dim"] != kw["d_model"]:
            x = F.linear(x, sd[f"{name}.proj.weight"], sd[f"{name}.proj.bias"])
        h = self._stack(x, f"{name}.", kw["encoder_layers"], kw["encoder_attention_heads"], lens,
                        sinusoids(kw["max_source_positions"], kw["d_model"]))
        if kw["output_dim"] != kw["d_model"]:
            h = F.linear(h, sd[f"{name}.out_proj.weight"], sd[f"{name}.out_proj.bias"])
        return h

    def tokenize(self, wav_list):
        """inference_tokenize (model.py:54-101) on one chunk -> pre-RVQ features (B, T/4.., 3072), codes (nq, B, T), lengths."""
        from oracle import rvq_np
        sd = self.sd
        mel, frames = self.log_mel(wav_list)
        sem, l2 = self._audio_encoder(mel, frames, "semantic_encoder")
        sem = self._transformer(sem, l2, "semantic_encoder_adapter")
        aco, _ = self._audio_encoder(mel, frames, "acoustic_encoder")
        h = self._transformer(torch.cat([sem, aco], dim=-1), l2, "pre_rvq_adapter")
        p = self.gp["downsample_kwargs"]["avg_pooler"]
        B, T, D = h.shape
        xt = h.permute(0, 2, 1)
        g = F.conv1d(xt, sd["downsample.gate_proj.weight"], stride=p).permute(0, 2, 1)
        u = F.conv1d(xt, sd["downsample.up_proj.weight"], stride=p).permute(0, 2, 1)
        xr = h.reshape(B, -1, D * p)
        c = F.linear(F.silu(g) * u, sd["downsample.down_proj.weight"])
        z = F.layer_norm(c + xr, (D * p,), sd["downsample.layer_norm.weight"], sd["downsample.layer_norm.bias"])
        l3 = l2 // p
        w_in = torch.from_numpy(weight_norm_weight(sd["quantizer.input_proj.weight_v"].numpy(), sd["quantizer.input_proj.weight_g"].numpy()))
        zt = F.linear(z, w_in[:, :, 0], sd["quantizer.input_proj.bias"])
        T3 = z.shape[1]
        valid = (torch.arange(T3)[None, :] < l3[:, None]).reshape(-1).numpy()
        cbs = np.stack([sd[f"quantizer.quantizers.{i}.codebook"].numpy() for i in range(self.nq)])
        codes, _, _, _ = rvq_np.rvq_forward(zt.reshape(B * T3, -1).numpy(), cbs, valid)
        return z, torch.from_numpy(codes).view(self.nq, B, T3), l3

    def encode(self, wav_list, overlap_seconds=10):
        sr, down = self.gp["input_sample_rate"], 1280
        chunk, dur = int(30 * sr), int((30 - overlap_seconds) * sr)
        code_dur = dur // down
        lens = [len(w) for w in wav_list]
        Lm = max(lens)
        B = len(wav_list)
        chunks = []
        for ci in range((Lm + dur - 1) // dur):
            start, end = ci * dur, min(ci * dur + chunk, Lm)
            cl = [min(max(l - start, 0), end - start) for l in lens]
            if max(cl) == 0:
                continue
            _, codes, l3 = self.tokenize([w[start:start + n] for w, n in zip(wav_list, cl)])
            valid = torch.zeros(self.nq, B, code_dur, dtype=torch.long)
            for b in range(B):
                n = min(int(l3[b]), code_dur)
                if n > 0:
                    valid[:, b, :n] = codes[:, b, :n]
            chunks.append(valid)
        full = torch.cat(chunks, -1)
        return [full[:, i, :lens[i] // down] for i in range(B)]
