"""The deterministic test signals of the encode goldens (same generator as oracle/gen_golden_codec.gen_encode)."""
import numpy as np


def make_signals():
    rng = np.random.default_rng(9)

    def sig(n):
        t = np.arange(n) / 16000.0
        x = 0.3 * np.sin(2 * np.pi * 220 * t) + 0.2 * np.sin(2 * np.pi * 1330 * t + 1.0) + 0.05 * rng.standard_normal(n)
        return (x * (0.5 + 0.5 * np.sin(2 * np.pi * 0.7 * t))).astype(np.float32)

    return [sig(35 * 16000), sig(3 * 16000)]
