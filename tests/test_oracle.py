"""Pins the CPU oracle (oracle/) to the reference's own outputs (tests/golden/)."""
import numpy as np
import pytest

from conftest import load_cases, case_args
from oracle import oracle

LOSS_CASES = load_cases('loss_cases.npz')
TARGET_CASES = load_cases('target_cases.npz')

ATOL = 1e-5   # per element (north star: 1e-5 fp32)
RTOL = 1e-5   # reduced scalars


def split(case):
    batch = {k[3:]: v for k, v in case.items() if k.startswith('in.')}
    outs = {k[4:]: v for k, v in case.items() if k.startswith('out.')}
    grads = {k[5:]: v for k, v in case.items() if k.startswith('grad.')}
    losses = {k[5:]: float(v) for k, v in case.items() if k.startswith('loss.')}
    return batch, outs, grads, losses


@pytest.mark.parametrize('dtype', [np.float32, np.float64], ids=['f32', 'f64'])
@pytest.mark.parametrize('name', sorted(LOSS_CASES))
def test_loss_oracle_matches_reference(name, dtype):
    case = LOSS_CASES[name]
    batch, outs, grads, losses = split(case)
    res = oracle.loss(batch, outs, case_args(case['meta']), dtype=dtype)
    for k, ref in losses.items():
        got = res['loss'][k]
        assert abs(got - ref) <= RTOL * abs(ref) + 1e-5, (name, k, got, ref)
    for k in ('v', 'r'):   # a missing head reports no loss in the reference, 0 here
        if k not in losses:
            assert res['loss'][k] == 0.0
    np.testing.assert_allclose(res['dpolicy_raw'], grads['policy'], rtol=0, atol=ATOL)
    if 'value' in grads:
        np.testing.assert_allclose(res['dvalue_raw'], grads['value'], rtol=0, atol=ATOL)
    if 'return' in grads:
        np.testing.assert_allclose(res['dreturn_raw'], grads['return'], rtol=0, atol=ATOL)


@pytest.mark.parametrize('name', sorted(TARGET_CASES))
def test_target_oracle_matches_reference(name):
    c = TARGET_CASES[name]
    algo = name.split('_')[0]
    gamma = 1.0 if name.endswith('outcome') else 0.9
    tg, ad = oracle.compute_target(algo, c['values'], c['returns'], c.get('rewards'), 0.7, gamma,
                                   c['rhos'], c['cs'], c['masks'])
    np.testing.assert_allclose(tg, np.broadcast_to(c['targets'], tg.shape), rtol=0, atol=1e-6)
    np.testing.assert_allclose(ad, c['advantages'], rtol=0, atol=1e-6)


def test_clip_adam_oracle_matches_torch():
    """torch.optim.Adam + clip_grad_norm_ are what the reference calls (train.py:331, 370-371)."""
    import torch
    rng = np.random.default_rng(0)
    n = 5000
    p0 = rng.standard_normal(n).astype(np.float32)
    tp = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.Adam([tp], lr=3e-4, weight_decay=1e-5)
    p, m, v = p0.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for step in range(4):
        g = (rng.standard_normal(n) * (0.2 if step % 2 else 0.01)).astype(np.float32)
        tp.grad = torch.from_numpy(g.copy())
        ref_norm = float(torch.nn.utils.clip_grad_norm_([tp], 4.0))
        opt.step()
        norm = oracle.clip_adam(p, g, m, v, 3e-4, step)
        assert abs(norm - ref_norm) <= 1e-5 * ref_norm
        np.testing.assert_allclose(p, tp.detach().numpy(), rtol=0, atol=1e-7)
