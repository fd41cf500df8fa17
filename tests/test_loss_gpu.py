"""GPU parity: the fused CUDA loss (through the C ABI) vs the reference's golden vectors and the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_cases, case_args

pytestmark = pytest.mark.gpu

LOSS_CASES = load_cases('loss_cases.npz')
TARGET_CASES = load_cases('target_cases.npz')
ATOL = 1e-5
RTOL = 1e-5


def to_dev(d):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in d.items()}


def split(case):
    batch = {k[3:]: v for k, v in case.items() if k.startswith('in.')}
    outs = {k[4:]: v for k, v in case.items() if k.startswith('out.')}
    grads = {k[5:]: v for k, v in case.items() if k.startswith('grad.')}
    losses = {k[5:]: float(v) for k, v in case.items() if k.startswith('loss.')}
    return batch, outs, grads, losses


@pytest.mark.parametrize('name', sorted(LOSS_CASES))
def test_fused_loss_matches_reference_golden(name):
    from handyrl_b200 import ops
    from oracle import oracle
    case = LOSS_CASES[name]
    batch, outs, grads, losses = split(case)
    args = case_args(case['meta'])
    res = ops.loss_fwd_bwd(to_dev(outs), to_dev(batch), args, taps=True)
    torch.cuda.synchronize()
    got = dict(zip(ops.LOSS_KEYS, res.losses.cpu().tolist()))
    for k, ref in losses.items():
        assert abs(got[k] - ref) <= RTOL * abs(ref) + 1e-5, (name, k, got[k], ref)
    np.testing.assert_allclose(res.dpolicy.cpu().numpy(), grads['policy'], rtol=0, atol=ATOL)
    if 'value' in grads:
        np.testing.assert_allclose(res.dvalue.cpu().numpy(), grads['value'], rtol=0, atol=ATOL)
    if 'return' in grads:
        np.testing.assert_allclose(res.dreturn.cpu().numpy(), grads['return'], rtol=0, atol=ATOL)
    # per-element intermediates against the (reference-pinned) oracle
    orc = oracle.loss(batch, outs, args, dtype=np.float64)
    for k in ('target_value', 'target_return', 'advantage', 'logp', 'rho', 'entropy'):
        np.testing.assert_allclose(res.taps[k].cpu().numpy(), orc[k], rtol=0, atol=ATOL, err_msg=k)
    # masks / indices are bit-exact: dcnt is an integer count
    assert got['dcnt'] == losses['dcnt']


@pytest.mark.parametrize('name', sorted(TARGET_CASES))
def test_compute_target_matches_reference_golden(name):
    from handyrl_b200 import ops
    c = TARGET_CASES[name]
    algo = name.split('_')[0]
    gamma = 1.0 if name.endswith('outcome') else 0.9
    d = to_dev({k: v for k, v in c.items()})
    tg, ad = ops.compute_target(algo, d['values'], d['returns'], d.get('rewards'), 0.7, gamma, d['rhos'], d['cs'], d['masks'])
    np.testing.assert_allclose(tg.cpu().numpy(), np.broadcast_to(c['targets'], tg.shape), rtol=0, atol=1e-6)
    np.testing.assert_allclose(ad.cpu().numpy(), c['advantages'], rtol=0, atol=1e-6)


def test_compute_target_unknown_algorithm_raises():
    from handyrl_b200 import ops
    x = torch.zeros(2, 3, 2, 1, device='cuda')
    with pytest.raises(ValueError):
        ops.compute_target('SARSA', x, x, x, 0.7, 0.9, x, x, x)


FULL = [  # BASELINE.json configs at full size (loss pass only)
    dict(id='cfg2', B=512, T=32, P=2, A=9, turn_based=True, observation=False, has_return=False,
         policy_target='UPGO', value_target='VTRACE', reward_kind='zero', burn_in=0),
    dict(id='cfg2_sim', B=512, T=32, P=2, A=9, turn_based=False, observation=False, has_return=False,
         policy_target='UPGO', value_target='VTRACE', reward_kind='zero', burn_in=0),
    dict(id='cfg3_geister', B=256, T=20, P=2, A=214, turn_based=True, observation=True, has_return=True,
         policy_target='TD', value_target='TD', reward_kind='step', burn_in=4),
    dict(id='cfg4_geese', B=256, T=32, P=4, A=4, turn_based=False, observation=False, has_return=False,
         policy_target='VTRACE', value_target='VTRACE', reward_kind='zero', burn_in=0),
    dict(id='cfg5_shard', B=512, T=64, P=2, A=512, turn_based=True, observation=False, has_return=False,
         policy_target='UPGO', value_target='VTRACE', reward_kind='zero', burn_in=0),
]


@pytest.mark.parametrize('cfg', FULL, ids=[c['id'] for c in FULL])
def test_full_size_against_oracle_and_properties(cfg):
    """Full BASELINE sizes: per-element parity with the C oracle (fast enough at these sizes) and
    size-independent properties: determinism, additivity over a split of B (what the multi-GPU shard
    relies on), zero gradient where masks are zero."""
    from handyrl_b200 import ops
    from handyrl_b200.synthetic import synthetic_batch, synthetic_outputs
    from oracle import oracle
    cfg = dict(cfg)
    cfg.pop('id')
    has_return = cfg.pop('has_return')
    args = {'turn_based_training': cfg['turn_based'], 'observation': cfg['observation'], 'gamma': 0.8, 'lambda': 0.7,
            'burn_in_steps': cfg['burn_in'], 'entropy_regularization': 0.1, 'entropy_regularization_decay': 0.1,
            'policy_target': cfg.pop('policy_target'), 'value_target': cfg.pop('value_target')}
    batch = synthetic_batch(cfg['B'], cfg['T'], cfg['P'], cfg['A'], turn_based=cfg['turn_based'],
                            observation=cfg['observation'], reward_kind=cfg['reward_kind'], burn_in=cfg['burn_in'],
                            seed=0, with_obs=False)
    outs = synthetic_outputs(batch, has_value=True, has_return=has_return, seed=1)
    dbatch = {k: v.cuda() for k, v in batch.items()}
    douts = {k: v.cuda() for k, v in outs.items()}
    res = ops.loss_fwd_bwd(douts, dbatch, args)
    torch.cuda.synchronize()
    losses = res.losses.cpu().numpy().astype(np.float64)
    dpol = res.dpolicy.cpu().numpy()

    orc = oracle.loss({k: v.numpy() for k, v in batch.items()}, {k: v.numpy() for k, v in outs.items()}, args,
                      dtype=np.float64)
    np.testing.assert_allclose(losses, orc['losses'], rtol=RTOL, atol=1e-5)
    np.testing.assert_allclose(dpol, orc['dpolicy_raw'], rtol=0, atol=ATOL)
    np.testing.assert_allclose(res.dvalue.cpu().numpy(), orc['dvalue_raw'], rtol=0, atol=ATOL)
    if has_return:
        np.testing.assert_allclose(res.dreturn.cpu().numpy(), orc['dreturn_raw'], rtol=0, atol=ATOL)

    # determinism: a second launch gives bit-identical results
    res2 = ops.loss_fwd_bwd(douts, dbatch, args)
    torch.cuda.synchronize()
    assert torch.equal(res2.losses.cpu(), res.losses.cpu())
    assert torch.equal(res2.dpolicy, res.dpolicy)

    # additivity over a split of the batch dimension (the multi-GPU sharding contract)
    h = cfg['B'] // 2
    parts = []
    for sl in (slice(0, h), slice(h, None)):
        r = ops.loss_fwd_bwd({k: v[sl].contiguous() for k, v in douts.items()},
                             {k: v[sl].contiguous() for k, v in dbatch.items()}, args)
        torch.cuda.synchronize()
        parts.append((r.losses.cpu().numpy().astype(np.float64), r.dpolicy.cpu().numpy()))
    np.testing.assert_allclose(parts[0][0] + parts[1][0], losses, rtol=1e-6, atol=1e-5)
    assert np.array_equal(np.concatenate([parts[0][1], parts[1][1]]), dpol)  # per-episode results do not depend on B

    # rows whose turn mask is zero get exactly zero policy gradient; burn-in steps too
    tm = batch['turn_mask'].numpy()[..., 0]
    scale = tm if batch['action_mask'].shape[2] == tm.shape[2] else tm.sum(-1, keepdims=True)
    assert np.all(dpol[scale == 0] == 0)
    if cfg['burn_in']:
        assert np.all(dpol[:, :cfg['burn_in']] == 0)


def test_bad_arguments_fail_loudly():
    from handyrl_b200 import ops, _capi
    from handyrl_b200.synthetic import synthetic_batch, synthetic_outputs
    batch = synthetic_batch(4, 6, 2, 9, with_obs=False)
    outs = synthetic_outputs(batch)
    args = {'turn_based_training': True, 'gamma': 0.8, 'lambda': 0.7, 'burn_in_steps': 0,
            'entropy_regularization': 0.1, 'entropy_regularization_decay': 0.1,
            'policy_target': 'UPGO', 'value_target': 'VTRACE'}
    with pytest.raises(_capi.HrlError):   # CPU tensors: no CPU path exists
        ops.loss_fwd_bwd(outs, batch, args)
    with pytest.raises(ValueError):       # the reference prints and returns None; we raise
        ops.loss_fwd_bwd({k: v.cuda() for k, v in outs.items()}, {k: v.cuda() for k, v in batch.items()},
                         dict(args, policy_target='NOPE'))
    with pytest.raises(_capi.HrlError):   # burn-in swallowing the whole window
        ops.loss_fwd_bwd({k: v.cuda() for k, v in outs.items()}, {k: v.cuda() for k, v in batch.items()},
                         dict(args, burn_in_steps=6))


@pytest.mark.parametrize('scan', ['serial', 'scan'])
@pytest.mark.parametrize('name', ['alt_UPGO_VTRACE', 'sim_TD_UPGO', 'obs_VTRACE_TD', 'alt_MC_UPGO', 'long', 'burnin_alt', 'alt4', 'wide512'])
def test_both_recurrence_forms_match_reference(name, scan):
    """The serial per-column loops (default for short windows) and the parallel suffix scan of max-affine maps
    (default for T >= 96) must both reproduce the reference."""
    from handyrl_b200 import ops
    case = LOSS_CASES[name]
    batch, outs, grads, losses = split(case)
    res = ops.loss_fwd_bwd(to_dev(outs), to_dev(batch), case_args(case['meta']), tuning={'recurrence': scan})
    torch.cuda.synchronize()
    got = dict(zip(ops.LOSS_KEYS, res.losses.cpu().tolist()))
    for k, ref in losses.items():
        assert abs(got[k] - ref) <= RTOL * abs(ref) + 1e-5, (name, scan, k, got[k], ref)
    np.testing.assert_allclose(res.dpolicy.cpu().numpy(), grads['policy'], rtol=0, atol=ATOL)
    if 'value' in grads:
        np.testing.assert_allclose(res.dvalue.cpu().numpy(), grads['value'], rtol=0, atol=ATOL)
    if 'return' in grads:
        np.testing.assert_allclose(res.dreturn.cpu().numpy(), grads['return'], rtol=0, atol=ATOL)


MODE_CASES = ['alt_UPGO_VTRACE', 'obs_TD_TD', 'geese4', 'wide', 'wide512', 'odd33', 'burnin_obs', 'novalue_ret', 'solo1']


@pytest.mark.parametrize('mode', ['rows-direct', 'rows-staged', 'bulk', 'element', 'group'])
@pytest.mark.parametrize('name', MODE_CASES)
def test_every_kernel_variant_matches_reference(name, mode):
    """HrlLossArgs.tuning.variant forces one data-movement variant of the fused kernel (falling back to the direct rows kernel
    where a variant does not apply to the shape): all of them must reproduce the reference."""
    from handyrl_b200 import ops
    case = LOSS_CASES[name]
    batch, outs, grads, losses = split(case)
    res = ops.loss_fwd_bwd(to_dev(outs), to_dev(batch), case_args(case['meta']), tuning={'variant': mode})
    torch.cuda.synchronize()
    got = dict(zip(ops.LOSS_KEYS, res.losses.cpu().tolist()))
    for k, ref in losses.items():
        assert abs(got[k] - ref) <= RTOL * abs(ref) + 1e-5, (name, mode, k, got[k], ref)
    np.testing.assert_allclose(res.dpolicy.cpu().numpy(), grads['policy'], rtol=0, atol=ATOL)
    if 'value' in grads:
        np.testing.assert_allclose(res.dvalue.cpu().numpy(), grads['value'], rtol=0, atol=ATOL)
    if 'return' in grads:
        np.testing.assert_allclose(res.dreturn.cpu().numpy(), grads['return'], rtol=0, atol=ATOL)


def test_wide_rows_without_staging_and_cluster_forms():
    """Full-size wide-row shape through the non-default forms: 1-CTA bulk, rows-direct with and without z staging."""
    from handyrl_b200 import ops
    from handyrl_b200.synthetic import synthetic_batch, synthetic_outputs
    args = {'turn_based_training': True, 'observation': False, 'gamma': 0.8, 'lambda': 0.7, 'burn_in_steps': 0,
            'entropy_regularization': 0.1, 'entropy_regularization_decay': 0.1, 'policy_target': 'UPGO', 'value_target': 'VTRACE'}
    batch = synthetic_batch(64, 64, 2, 512, seed=3, with_obs=False)
    outs = synthetic_outputs(batch, seed=4)
    db, do = {k: v.cuda() for k, v in batch.items()}, {k: v.cuda() for k, v in outs.items()}
    ref = ops.loss_fwd_bwd(do, db, args)            # default: 2-CTA cluster bulk kernel (checked against the oracle elsewhere)
    torch.cuda.synchronize()
    for tuning in ({'cluster': 1}, {'variant': 'rows-direct'}, {'variant': 'rows-direct', 'unstaged': 1}):
        res = ops.loss_fwd_bwd(do, db, args, tuning=tuning)
        torch.cuda.synchronize()
        np.testing.assert_allclose(res.losses.cpu().numpy(), ref.losses.cpu().numpy(), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(res.dpolicy.cpu().numpy(), ref.dpolicy.cpu().numpy(), rtol=0, atol=2e-6)
        np.testing.assert_allclose(res.dvalue.cpu().numpy(), ref.dvalue.cpu().numpy(), rtol=0, atol=2e-6)


@pytest.mark.parametrize('A', [640, 1000, 1024])
def test_action_spaces_up_to_1024_against_oracle(A):
    from handyrl_b200 import ops
    from handyrl_b200.synthetic import synthetic_batch, synthetic_outputs
    from oracle import oracle
    args = {'turn_based_training': True, 'observation': False, 'gamma': 0.8, 'lambda': 0.7, 'burn_in_steps': 0,
            'entropy_regularization': 0.1, 'entropy_regularization_decay': 0.1, 'policy_target': 'UPGO', 'value_target': 'VTRACE'}
    batch = synthetic_batch(6, 10, 2, A, seed=A, with_obs=False)
    outs = synthetic_outputs(batch, seed=A + 1)
    res = ops.loss_fwd_bwd({k: v.cuda() for k, v in outs.items()}, {k: v.cuda() for k, v in batch.items()}, args)
    torch.cuda.synchronize()
    orc = oracle.loss({k: v.numpy() for k, v in batch.items()}, {k: v.numpy() for k, v in outs.items()}, args, dtype=np.float64)
    np.testing.assert_allclose(res.losses.cpu().numpy(), orc['losses'], rtol=RTOL, atol=1e-5)
    np.testing.assert_allclose(res.dpolicy.cpu().numpy(), orc['dpolicy_raw'], rtol=0, atol=ATOL)
    np.testing.assert_allclose(res.dvalue.cpu().numpy(), orc['dvalue_raw'], rtol=0, atol=ATOL)


def test_action_space_beyond_the_built_range_is_refused():
    from handyrl_b200 import ops, _capi
    from handyrl_b200.synthetic import synthetic_batch, synthetic_outputs
    args = {'turn_based_training': True, 'observation': False, 'gamma': 0.8, 'lambda': 0.7, 'burn_in_steps': 0,
            'entropy_regularization': 0.1, 'entropy_regularization_decay': 0.1, 'policy_target': 'TD', 'value_target': 'TD'}
    batch = synthetic_batch(2, 4, 2, 1500, seed=1, with_obs=False)
    outs = synthetic_outputs(batch, seed=2)
    with pytest.raises(_capi.HrlError, match='not built'):
        ops.loss_fwd_bwd({k: v.cuda() for k, v in outs.items()}, {k: v.cuda() for k, v in batch.items()}, args)


@pytest.mark.parametrize('B,T,P,A,turn_based,cluster', [(64, 64, 2, 512, True, 0), (33, 20, 2, 320, True, 1), (16, 32, 4, 264, False, 2)])
def test_bf16_logit_io_equals_the_fp32_pass_on_the_same_logits(B, T, P, A, turn_based, cluster):
    """HrlLossArgs.io_bf16 (wide rows): the logits are read and the policy gradient is written as bf16, everything in between is
    the fp32 arithmetic of the default path.  So against the fp32 call on the SAME (bf16-representable) logits: the six losses
    and the value gradient are bit-identical, and the policy gradient is the fp32 one rounded to nearest-even bf16 -- tolerance
    of the flag: half a bf16 ulp (2^-9 relative) of each gradient element, nothing else."""
    from handyrl_b200 import ops
    from handyrl_b200.synthetic import synthetic_batch, synthetic_outputs
    args = {'turn_based_training': turn_based, 'observation': False, 'gamma': 0.8, 'lambda': 0.7, 'burn_in_steps': 0,
            'entropy_regularization': 0.1, 'entropy_regularization_decay': 0.1, 'policy_target': 'UPGO', 'value_target': 'VTRACE'}
    batch = synthetic_batch(B, T, P, A, turn_based=turn_based, seed=5, with_obs=False)
    outs = synthetic_outputs(batch, seed=6)
    db, do = {k: v.cuda() for k, v in batch.items()}, {k: v.cuda() for k, v in outs.items()}
    half = dict(do, policy=do['policy'].to(torch.bfloat16))
    widened = dict(do, policy=half['policy'].float())
    tuning = {'variant': 'bulk', 'cluster': cluster} if cluster else {'variant': 'bulk'}
    ref = ops.loss_fwd_bwd(widened, db, args, tuning=tuning)
    res = ops.loss_fwd_bwd(half, db, args, tuning=tuning)
    torch.cuda.synchronize()
    assert res.dpolicy.dtype == torch.bfloat16
    assert torch.equal(res.losses, ref.losses)
    assert torch.equal(res.dvalue, ref.dvalue)
    assert torch.equal(res.dpolicy, ref.dpolicy.to(torch.bfloat16))
    # and the flag refuses shapes the wide-row kernel does not cover instead of silently widening
    small = synthetic_batch(8, 8, 2, 9, seed=1, with_obs=False)
    so = synthetic_outputs(small, seed=2)
    with pytest.raises(Exception, match='bf16'):
        ops.loss_fwd_bwd(dict({k: v.cuda() for k, v in so.items()}, policy=so['policy'].cuda().to(torch.bfloat16)),
                         {k: v.cuda() for k, v in small.items()}, args)
