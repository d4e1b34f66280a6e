"""Shared fixtures for the parity tests (oracle side only; nothing here is product code)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TINY = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
            head_dim=128, rms_norm_eps=1e-6, rope_theta=1e6, vocab_size=152697, speech_vocab_size=1025, channels=8,
            speech_token_range=[151665, 152689])
TINY_SEED = 1234


def gold(name):
    return np.load(os.path.join(GOLD, name))


def tiny_model(device="cuda", shape=TINY, seed=TINY_SEED, tied=False):
    """The CUDA drop-in model loaded with the oracle's seeded weights."""
    from moss_ttsd_b200.modeling_asteroid import AsteroidTTSConfig, AsteroidTTSInstruct
    from oracle import lm_oracle
    cfg = AsteroidTTSConfig(**{k: v for k, v in shape.items()}, eos_token_id=152694, pad_token_id=151643,
                            tie_word_embeddings=tied)
    m = AsteroidTTSInstruct(cfg, device=device)
    sd = lm_oracle.make_weights(shape, seed, tied=tied)
    m.load_state_dict(sd, tie_word_embeddings=tied)
    return m, sd
