"""Continuous batching: rows evicted and refilled inside a running batch must produce exactly the tokens
of solo runs. Uses the planted-margin model (tests/golden/lm_margin.npz's weights), whose greedy decisions have a top-2
gap far above bf16 noise, so 'equal' means equal."""
import numpy as np
import pytest
import torch

from oracle import lm_oracle


def test_page_pool_reserve_take_release():
    from moss_ttsd_b200.continuous import PagePool
    pool = PagePool(10)
    assert pool.available == 9
    assert pool.reserve(6) and pool.available == 3
    assert not pool.reserve(4)
    a = pool.take_reserved(2)
    assert len(set(a)) == 2 and 0 not in a and pool.available == 3 and pool.reserved == 4
    b = pool.take_reserved(4)
    assert not set(a) & set(b)
    pool.release(a + b)
    assert pool.available == 9 and pool.reserved == 0
    assert pool.reserve(9) and not pool.reserve(1)
    pool.release([], unreserve=9)
    assert pool.available == 9
    with pytest.raises(ValueError):
        PagePool(1)


def _requests(rng, n, shape):
    """Unpadded delay-shifted prompt grids of different lengths + a per-request frame budget."""
    from oracle.gen_golden import make_prompt
    grids, budgets = [], []
    for i in range(n):
        ids, _ = make_prompt(rng, 1, [int(rng.integers(3, 12))], [int(rng.integers(9, 20))], shape)
        grids.append(torch.from_numpy(ids[0]))
        budgets.append(int(rng.integers(4, 30)))
    return grids, budgets


@pytest.fixture(scope="module")
def margin_model():
    from moss_ttsd_b200.modeling_asteroid import AsteroidTTSConfig, AsteroidTTSInstruct
    shape = lm_oracle.MARGIN_SHAPE
    cfg = AsteroidTTSConfig(**shape, eos_token_id=152694, pad_token_id=151643, tie_word_embeddings=False)
    m = AsteroidTTSInstruct(cfg, device="cuda")
    m.load_state_dict(lm_oracle.make_planted_weights(shape, lm_oracle.MARGIN_SEED, emb_gain=lm_oracle.MARGIN_GAIN),
                      tie_word_embeddings=False)
    m.generation_config.eos_token_id = 152694
    return m


@pytest.mark.gpu
@pytest.mark.parametrize("slots,pool_pages", [(2, None), (3, 9), (8, None)])
def test_evicted_and_refilled_rows_equal_solo_runs(margin_model, slots, pool_pages):
    m = margin_model
    rng = np.random.default_rng(5)
    grids, budgets = _requests(rng, 9, lm_oracle.MARGIN_SHAPE)
    C = 8
    solo = []
    for g, nb in zip(grids, budgets):
        T = g.shape[0]
        out = m.generate(input_ids=g[None].cuda(), attention_mask=torch.ones(1, T).cuda(), max_new_tokens=40, do_sample=False,
                         eos_at=[T - (C - 1) + nb])
        solo.append(out[0].cpu())
    # the budget really shortens the rows: budget rows, then the EOS row and 6 wind-down rows, not the
    # 47 rows of max_new_tokens
    for s, g, nb in zip(solo, grids, budgets):
        P = g.shape[0] - 7
        assert s.shape[0] == P + nb + 7, (s.shape, P, nb)
        assert int(s[P + nb, 0]) == 152694 and (s[-1, 1:7] == 1024).all()
    m.kv_page_size = 16                       # many pages per request -> growth and recycling are exercised
    try:
        m._continuous = None
        outs = m.generate_continuous(grids, max_new_tokens=40, max_batch=slots, do_sample=False, pool_pages=pool_pages,
                                     eos_at=[g.shape[0] - 7 + nb for g, nb in zip(grids, budgets)], sync_every=4)
        cd = m._continuous[1]
    finally:
        m.kv_page_size = 64
    assert len(outs) == len(grids)
    for i, (o, s) in enumerate(zip(outs, solo)):
        assert tuple(o.shape) == tuple(s.shape), (i, o.shape, s.shape)
        assert torch.equal(o.cpu(), s), i
    assert cd.admitted == len(grids) and cd.pool.available == cd.pool.num_pages - 1 and cd.pool.reserved == 0
    if slots < len(grids):
        assert cd.decode_steps < sum(s.shape[0] for s in solo)      # rows really shared steps
        # longest-first admission: same rows per request (results come back in prompt order), no more steps than FIFO
        fifo_steps = cd.decode_steps
        m.kv_page_size = 16
        try:
            outs2 = m.generate_continuous(grids, max_new_tokens=40, max_batch=slots, do_sample=False, pool_pages=pool_pages,
                                          eos_at=[g.shape[0] - 7 + nb for g, nb in zip(grids, budgets)], sync_every=4,
                                          queue_order="longest_first")
        finally:
            m.kv_page_size = 64
        assert all(torch.equal(a.cpu(), b) for a, b in zip(outs2, solo))
        assert m._continuous[1].decode_steps - fifo_steps <= fifo_steps


@pytest.mark.gpu
def test_static_batch_with_budgets_equals_solo_runs(margin_model):
    """The same requests as one left-padded static batch (run-to-longest, the reference's scheme) give the same rows."""
    m = margin_model
    rng = np.random.default_rng(6)
    grids, budgets = _requests(rng, 5, lm_oracle.MARGIN_SHAPE)
    T = max(g.shape[0] for g in grids)
    ids = torch.full((len(grids), T, 8), 1024, dtype=torch.int64)
    ids[:, :, 0] = 151643
    mask = torch.zeros(len(grids), T)
    for b, g in enumerate(grids):
        ids[b, T - g.shape[0]:] = g
        mask[b, T - g.shape[0]:] = 1
    out = m.generate(input_ids=ids.cuda(), attention_mask=mask.cuda(), max_new_tokens=40, do_sample=False,
                     eos_at=[T - 7 + nb for nb in budgets]).cpu()
    cont = m.generate_continuous(grids, max_new_tokens=40, max_batch=2, do_sample=False,
                                 eos_at=[g.shape[0] - 7 + nb for g, nb in zip(grids, budgets)])
    for b, (g, nb) in enumerate(zip(grids, budgets)):
        pad = T - g.shape[0]
        n = g.shape[0] - 7 + nb + 7
        assert torch.equal(out[b, pad:pad + n], cont[b].cpu()), b
        assert (out[b, pad + n:, 0] == 152694).all() and (out[b, pad + n:, 1:] == 1024).all()
