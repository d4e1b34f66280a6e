"""CPU statements of summation orders that CUDA kernels rely on for bit-exactness (fp32 addition is commutative but not
associative, so a re-laid-out kernel must rebuild the same tree). No GPU needed; the kernels themselves are compared on
the device by tests/test_attention_gpu.py (fused decode prologue == mtts_qknorm_rope_kvappend, same bits)."""
import numpy as np

f32 = np.float32


def _lane_partials(x):
    """qknorm_rope_kv_kernel (lm_ops.cu): lane v of the warp owns elements (2v, 2v+1, 64+2v, 65+2v) of a 128-wide head and
    chains x0*x0, fma(x1,x1,.), fma(x2,x2,.), fma(x3,x3,.). fma = one rounding: emulate in float64 and round once."""
    p = np.empty(32, dtype=f32)
    for v in range(32):
        e = [x[2 * v], x[2 * v + 1], x[64 + 2 * v], x[65 + 2 * v]]
        s = f32(e[0] * e[0])                                   # fp32 multiply
        for t in e[1:]:
            s = f32(np.float64(t) * np.float64(t) + np.float64(s))   # fmaf: exact product + add, rounded once
        p[v] = s
    return p


def _butterfly(p):
    """warp_sum (common.cuh): v += shfl_xor(v, o) for o = 16, 8, 4, 2, 1; every lane ends with the same bits."""
    v = p.copy()
    for o in (16, 8, 4, 2, 1):
        v = np.array([f32(v[l] + v[l ^ o]) for l in range(32)], dtype=f32)
    assert len({x.tobytes() for x in v}) == 1
    return v[0]


def _wide(p):
    """qknorm_rope_kv_wide_kernel: physical lane j (8 per head) holds the partials of virtual lanes 4j..4j+3; levels
    xor 16 / 8 / 4 are shuffles over physical lanes xor 4 / 2 / 1 applied to each of the four values, levels xor 2 / 1
    are (s0 + s2) + (s1 + s3) inside the lane."""
    s = p.reshape(8, 4).copy()          # [physical lane j][m]
    for o in (4, 2, 1):
        s = np.array([[f32(s[j][m] + s[j ^ o][m]) for m in range(4)] for j in range(8)], dtype=f32)
    out = np.array([f32(f32(s[j][0] + s[j][2]) + f32(s[j][1] + s[j][3])) for j in range(8)], dtype=f32)
    assert len({x.tobytes() for x in out}) == 1
    return out[0]


def test_wide_qknorm_kernel_rebuilds_the_warp_butterfly_bit_for_bit():
    rng = np.random.default_rng(11)
    for trial in range(300):
        scale = f32(10.0 ** rng.uniform(-3, 3))
        x = (rng.standard_normal(128).astype(f32) * scale)
        # the kernels see bf16 inputs: drop the low 16 mantissa bits
        x = (x.view(np.uint32) & np.uint32(0xFFFF0000)).view(f32)
        p = _lane_partials(x)
        a, b = _butterfly(p), _wide(p)
        assert a.tobytes() == b.tobytes(), (trial, a, b)


def test_a_different_association_is_not_bit_exact():
    """The guard has teeth: summing the 32 partials left to right differs from the butterfly on some input."""
    rng = np.random.default_rng(5)
    diff = 0
    for _ in range(200):
        x = rng.standard_normal(128).astype(f32)
        p = _lane_partials(x)
        seq = f32(0)
        for v in p:
            seq = f32(seq + v)
        diff += seq.tobytes() != _butterfly(p).tobytes()
    assert diff > 0
