"""The fused tower engine (handyrl_b200/tower.py on hrl_gemm_fused) against the same nets.BoardNet run (a) module by module on
the same tensor-core products (fastnet: must agree closely -- same arithmetic, fused differently) and (b) by PyTorch in
float64 (outputs to 3e-5; gradients within the accuracy of 3xTF32 products, whose fp32 accumulator truncates toward zero
-- a systematic ~1e-5 relative shrink per product that batch sums amplify).  Also BatchNorm
running buffers."""
import copy
import zlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = {
    'tictactoe': dict(kw=dict(planes=3, board=(3, 3), width=32, depth=3, actions=9), M=16384 // 8),
    'ragged_rows': dict(kw=dict(planes=3, board=(3, 3), width=32, depth=3, actions=9), M=300),
    'return_head': dict(kw=dict(planes=2, board=(3, 3), width=16, depth=2, actions=7, return_head=True), M=515),
    'board4x4': dict(kw=dict(planes=5, board=(4, 4), width=8, depth=2, actions=12, policy_maps=3), M=260),
}


@pytest.mark.parametrize('name', sorted(CASES))
def test_fused_tower_matches_float64_modules(name, seed=None):
    from handyrl_b200 import nets, tower
    case = CASES[name]
    torch.manual_seed(zlib.crc32(name.encode()) % 1000 if seed is None else seed)      # (hash() of a str changes per process)
    ref = nets.BoardNet(**case['kw']).double().cuda().train()
    for blk in ref.tower:          # non-trivial affine parameters and running statistics
        blk[1].weight.data.uniform_(0.5, 1.5)
        blk[1].bias.data.normal_(0, 0.3)
        blk[1].running_mean.normal_(0, 0.1)
        blk[1].running_var.uniform_(0.5, 2.0)
    fast = copy.deepcopy(ref).float()
    modular = copy.deepcopy(ref).float()
    from handyrl_b200 import fastnet
    fastnet.optimize_small_boards(modular)
    assert tower.supports(fast)
    M = case['M']
    board, planes = case['kw']['board'], case['kw']['planes']
    x = (torch.rand(M, planes, *board, device='cuda') < 0.4).float()
    eng = tower.FusedBoardNet(fast, M, torch.device('cuda'))
    for p in fast.parameters():
        p.grad = torch.full_like(p, 7.0)           # backward must overwrite, not accumulate
    out = eng.forward(x)
    want = ref(x.double())
    for k in want:
        np.testing.assert_allclose(out[k].double().cpu().numpy(), want[k].detach().cpu().numpy(), rtol=0, atol=3e-5, err_msg=k)
    g = torch.Generator().manual_seed(5)
    dout = {k: torch.randn(v.shape, generator=g).cuda() for k, v in out.items()}
    sum((want[k] * dout[k].double()).sum() for k in want).backward()
    eng.backward(dout['policy'], dout['value'], dout.get('return'))
    mout = modular(x)
    sum((mout[k] * dout[k]).sum() for k in mout).backward()
    torch.cuda.synchronize()
    for (k, pr), (_, pf), (_, pm) in zip(ref.named_parameters(), fast.named_parameters(), modular.named_parameters()):
        scale = pr.grad.abs().max().item() + 1e-6
        assert (pf.grad - pm.grad).abs().max().item() <= 1e-3 * scale, (k, 'fused vs module-by-module')
        assert (pf.grad.double() - pr.grad).abs().max().item() <= 5e-2 * scale, (k, 'fused vs float64')
    for (k, br), (_, bf) in zip(ref.named_buffers(), fast.named_buffers()):
        if br.dtype.is_floating_point:
            np.testing.assert_allclose(bf.double().cpu().numpy(), br.cpu().numpy(), rtol=1e-5, atol=1e-6, err_msg=k)
        else:
            assert int(bf) == int(br), k


def test_learner_step_uses_the_engine_and_matches_the_module_path():
    """LearnerStep picks the fused engine for nets.BoardNet; three optimiser steps land on the same weights as the
    module-by-module path (which the reference's goldens pin)."""
    from handyrl_b200.nets import tictactoe_net
    from handyrl_b200.synthetic import synthetic_batch
    from handyrl_b200.train import LearnerStep
    args = {'turn_based_training': True, 'observation': False, 'gamma': 0.8, 'lambda': 0.7, 'burn_in_steps': 0, 'forward_steps': 8,
            'entropy_regularization': 0.1, 'entropy_regularization_decay': 0.1, 'policy_target': 'UPGO', 'value_target': 'VTRACE'}
    res = {}
    for fused in (True, False):
        torch.manual_seed(3)
        stepper = LearnerStep(tictactoe_net(), args, synthetic_batch(64, 8, 2, 9, seed=1), lr=1e-4, fused_tower=fused)
        assert (stepper.engine is not None) == fused
        losses = []
        for s in range(3):
            stepper.step(stepper.new_packed().fill(synthetic_batch(64, 8, 2, 9, seed=10 + s)))
            losses.append(stepper.read_losses())
        res[fused] = (losses, stepper.cpu_state_dict(), stepper.launches_per_step)
    for a, b in zip(res[True][0], res[False][0]):
        scale = max(abs(v) for v in b.values())
        for k in b:
            assert abs(a[k] - b[k]) <= 1e-4 * scale + 1e-4, (k, a[k], b[k])
    for (k, va), (_, vb) in zip(res[True][1].items(), res[False][1].items()):
        if va.dtype.is_floating_point:
            bad = (va - vb).abs() > 2e-5 + 1e-3 * vb.abs()
            assert bad.float().mean() <= 2e-3, (k, int(bad.sum()))
            assert (va - vb).abs().max() <= 1e-3, k
        else:
            assert int(va) == int(vb), k
    assert res[True][2] < res[False][2]
