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
