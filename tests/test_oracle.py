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


@pytest.mark.parametrize('name', sorted(LOSS_CASES))
def test_torch_port_matches_reference(name):
    """The eager-PyTorch CPU port (bench.py's cpu_baseline / --impl reference) vs the reference."""
    import torch
    from oracle import torch_learner
    case = LOSS_CASES[name]
    batch, outs, grads, losses = split(case)
    tb = {k: torch.from_numpy(v) for k, v in batch.items()}
    leaves = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in outs.items()}
    got, dcnt = torch_learner.loss_from_raw(leaves, tb, case_args(case['meta']))
    got['total'].backward()
    assert float(dcnt) == losses['dcnt']
    for k, ref in losses.items():
        if k != 'dcnt':
            assert abs(float(got[k]) - ref) <= RTOL * abs(ref) + 1e-5, (name, k, float(got[k]), ref)
    for k, ref in grads.items():
        np.testing.assert_allclose(leaves[k].grad.numpy(), ref, rtol=0, atol=ATOL, err_msg=k)


def test_torch_port_full_steps_match_reference():
    """CpuLearner (model + loss + clip + Adam) reproduces the reference's three optimiser steps."""
    import os
    import pickle
    import torch
    from conftest import GOLDEN
    from oracle.torch_learner import CpuLearner
    from handyrl_b200.nets import tictactoe_net, load_state_by_order
    from handyrl_b200.synthetic import synthetic_batch
    with open(os.path.join(GOLDEN, 'step_cases.pkl'), 'rb') as f:
        cases = pickle.load(f)
    for name, c in cases.items():
        B, T, P, A = c['dims']
        args = c['args']
        net = load_state_by_order(tictactoe_net(), c['state0'])
        lrn = CpuLearner(net, args, lr=c['lr'])
        for s, ref in enumerate(c['steps']):
            batch = synthetic_batch(B, T, P, A, turn_based=args['turn_based_training'], observation=args['observation'], seed=40 + s)
            losses, dcnt = lrn.step(batch)
            assert dcnt == ref['dcnt']
            for k, v in ref['losses'].items():
                assert abs(losses[k] - v) <= 1e-5 * abs(v) + 1e-5, (name, s, k)
        for (k, v), (kr, vr) in zip(net.state_dict().items(), c['state3'].items()):
            np.testing.assert_allclose(v.numpy(), vr, rtol=1e-5, atol=1e-6, err_msg=k)


@pytest.mark.parametrize('name', ['geister_obs', 'geister_alt', 'geese'])
def test_torch_port_and_net_standins_match_reference_nets(name):
    """The architecture stand-ins (DRC ConvLSTM = GeisterNet, torus tower = GeeseNet) driven by the CPU port reproduce three
    optimiser steps of the reference's own networks, recurrent path (burn-in, hidden masking) included."""
    import os
    import pickle
    from conftest import GOLDEN, net_case_setup, noise_driven
    from oracle.torch_learner import CpuLearner
    with open(os.path.join(GOLDEN, 'net_step_cases.pkl'), 'rb') as f:
        c = pickle.load(f)[name]
    net, batches = net_case_setup(c)
    lrn = CpuLearner(net, c['args'], lr=c['lr'])
    for s, (batch, ref) in enumerate(zip(batches, c['steps'])):
        losses, dcnt = lrn.step(batch)
        assert dcnt == ref['dcnt']
        scale = max(abs(v) for v in ref['losses'].values())     # `total` is a difference of the larger terms
        for k, v in ref['losses'].items():
            assert abs(losses[k] - v) <= 1e-5 * scale + 1e-5, (name, s, k, losses[k], v)
        assert abs(lrn.grad_norm - ref['grad_norm']) <= 1e-3 * ref['grad_norm']
    for (k, v), (kr, vr) in zip(net.state_dict().items(), c['state3'].items()):
        if noise_driven(c, k):
            continue
        np.testing.assert_allclose(v.numpy(), vr, rtol=1e-4, atol=2e-5, err_msg='%s/%s' % (k, kr))   # Adam turns 1e-9 gradient noise into ~1e-6 weight noise
