"""GPU parity of the WHOLE learner step (net fwd -> fused loss -> net bwd -> clip+Adam) against the
reference's own three optimiser steps (golden step_cases.pkl) and its recurrent path (rnn_cases.pkl)."""
import os
import pickle

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

with open(os.path.join(GOLDEN, 'step_cases.pkl'), 'rb') as f:
    STEP_CASES = pickle.load(f)
with open(os.path.join(GOLDEN, 'rnn_cases.pkl'), 'rb') as f:
    RNN_CASES = pickle.load(f)
with open(os.path.join(GOLDEN, 'net_step_cases.pkl'), 'rb') as f:
    NET_CASES = pickle.load(f)


@pytest.fixture(autouse=True)
def default_precision_flags():
    """PyTorch's defaults (cuDNN TF32 allowed): LearnerStep itself must switch to full fp32, not the harness."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize('use_graph', [False, True], ids=['eager', 'graph'])
@pytest.mark.parametrize('name', sorted(NET_CASES))
def test_three_learner_steps_match_reference_geister_and_geese_nets(name, use_graph):
    """BASELINE configs[2]/[3] architectures (DRC ConvLSTM with burn-in and hidden masking; 12-block torus tower) through
    LearnerStep: losses, gradient norm and final weights of three optimiser steps of the reference's own GeisterNet /
    GeeseNet (weights loaded by order into the stand-ins)."""
    from conftest import net_case_setup, noise_driven
    from handyrl_b200.train import LearnerStep
    c = NET_CASES[name]
    net, batches = net_case_setup(c)
    stepper = LearnerStep(net, c['args'], batches[0], lr=c['lr'], use_graph=use_graph)
    assert torch.backends.cudnn.allow_tf32 is False
    for s, (batch, ref) in enumerate(zip(batches, c['steps'])):
        stepper.step(stepper.new_packed().fill(batch))
        got = stepper.read_losses()
        scale = max(abs(v) for v in ref['losses'].values())
        for k, v in ref['losses'].items():
            assert abs(got[k] - v) <= 1e-4 * scale + 1e-4, (s, k, got[k], v)
        assert got['dcnt'] == ref['dcnt']
        assert abs(float(stepper.opt.grad_norm) - ref['grad_norm']) <= 2e-3 * ref['grad_norm']
    final = stepper.cpu_state_dict()
    for (k, v), (kr, vr) in zip(final.items(), c['state3'].items()):
        if noise_driven(c, k):
            continue
        if v.dtype.is_floating_point:
            np.testing.assert_allclose(v.numpy(), vr, rtol=1e-3, atol=5e-5, err_msg='%s/%s' % (k, kr))
        else:
            assert int(v) == int(vr)


@pytest.mark.parametrize('use_graph', [False, True], ids=['eager', 'graph'])
@pytest.mark.parametrize('name', sorted(STEP_CASES))
def test_three_learner_steps_match_reference(name, use_graph):
    from handyrl_b200.nets import tictactoe_net, load_state_by_order
    from handyrl_b200.synthetic import synthetic_batch
    from handyrl_b200.train import LearnerStep
    c = STEP_CASES[name]
    B, T, P, A = c['dims']
    args = c['args']
    net = load_state_by_order(tictactoe_net(), c['state0'])
    mk = lambda s: synthetic_batch(B, T, P, A, turn_based=args['turn_based_training'], observation=args['observation'], seed=40 + s)
    stepper = LearnerStep(net, args, mk(0), lr=c['lr'], use_graph=use_graph)
    for s, ref in enumerate(c['steps']):
        pk = stepper.new_packed().fill(mk(s))
        stepper.step(pk)
        got = stepper.read_losses()
        for k, v in ref['losses'].items():
            assert abs(got[k] - v) <= 2e-4 * abs(v) + 1e-4, (s, k, got[k], v)
        assert got['dcnt'] == ref['dcnt']
        assert abs(float(stepper.opt.grad_norm) - ref['grad_norm']) <= 1e-3 * ref['grad_norm']
    final = stepper.cpu_state_dict()
    for (k, v), (kr, vr) in zip(final.items(), c['state3'].items()):
        if v.dtype.is_floating_point:
            np.testing.assert_allclose(v.numpy(), vr, rtol=1e-4, atol=2e-5, err_msg='%s/%s' % (k, kr))
        else:
            assert int(v) == int(vr)     # num_batches_tracked


@pytest.mark.parametrize('name', sorted(RNN_CASES))
def test_recurrent_compute_loss_and_param_grads(name):
    from handyrl_b200.batch import tree_map
    from handyrl_b200.nets import GatedBoardNet
    from handyrl_b200.train import compute_loss
    c = RNN_CASES[name]
    torch.backends.cudnn.allow_tf32 = False      # the bare drop-in function leaves backend flags to its caller
    net = GatedBoardNet()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in c['state0'].items()})
    net = net.cuda().train()
    batch = tree_map(lambda a: torch.from_numpy(a).cuda(), c['batch'])
    B, T, P, A = c['dims']
    hidden = tree_map(lambda h: h.cuda(), net.init_hidden([B, P]))
    losses, dcnt = compute_loss(batch, net, hidden, c['args'])
    losses['total'].backward()
    assert dcnt == c['dcnt']
    for k, v in c['losses'].items():
        assert abs(float(losses[k]) - v) <= 1e-4 * abs(v) + 1e-4, (k, float(losses[k]), v)
    for k, p in net.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), c['param_grads'][k], rtol=1e-3, atol=2e-5, err_msg=k)


@pytest.mark.parametrize('gpu_replay', [True, False], ids=['gpu_replay', 'host_batcher'])
def test_trainer_thread_protocol(gpu_replay):
    """Trainer.run() as the Learner drives it (train.py:389-400, 342-345): feed episodes, call update()."""
    import threading
    from handyrl_b200.train import Trainer
    from handyrl_b200.nets import tictactoe_net
    with open(os.path.join(GOLDEN, 'batch_cases.pkl'), 'rb') as f:
        case = pickle.load(f)['tictactoe']
    args = dict(case['args'], batch_size=8, minimum_episodes=4, num_batchers=1, **{'lambda': 0.7},
                entropy_regularization=0.1, entropy_regularization_decay=0.1, policy_target='UPGO', value_target='VTRACE',
                gpu_replay=gpu_replay)
    tr = Trainer(args, tictactoe_net())
    tr.episodes.extend(case['episodes'])
    th = threading.Thread(target=tr.run, daemon=True)
    th.start()
    model, steps = tr.update()
    assert steps >= 1 and not model.training and next(model.parameters()).device.type == 'cpu'
    model2, steps2 = tr.update()
    assert steps2 > steps
    assert any(not torch.equal(a, b) for a, b in zip(model.state_dict().values(), model2.state_dict().values()))
    tr.stop()
    th.join(timeout=10)
    assert not th.is_alive()


@pytest.mark.parametrize('name', sorted(RNN_CASES))
def test_recurrent_learner_step_graph_equals_eager_and_reference_losses(name):
    """The recurrent path (hidden masking, burn-in in eval mode, dict observations) inside LearnerStep: the first
    step's loss sums match the reference's, and the CUDA-graph step is identical to the eager one over 3 steps."""
    from handyrl_b200.batch import tree_map
    from handyrl_b200.nets import GatedBoardNet
    from handyrl_b200.train import LearnerStep
    c = RNN_CASES[name]
    batch = tree_map(lambda a: torch.from_numpy(a), c['batch'])
    results = {}
    for use_graph in (False, True):
        net = GatedBoardNet()
        net.load_state_dict({k: torch.from_numpy(v) for k, v in c['state0'].items()})
        stepper = LearnerStep(net, c['args'], batch, lr=1e-3, use_graph=use_graph)
        losses = []
        for _ in range(3):
            stepper.step(stepper.new_packed().fill(batch))
            losses.append(stepper.read_losses())
        results[use_graph] = (losses, stepper.cpu_state_dict())
    first = results[True][0][0]
    for k, v in c['losses'].items():
        assert abs(first[k] - v) <= 1e-4 * abs(v) + 1e-4, (k, first[k], v)
    assert first['dcnt'] == c['dcnt']
    for a, b in zip(results[False][0], results[True][0]):
        for k in a:
            assert abs(a[k] - b[k]) <= 1e-5 * abs(a[k]) + 1e-5, (k, a[k], b[k])
    for (k, va), (_, vb) in zip(results[False][1].items(), results[True][1].items()):
        torch.testing.assert_close(vb.float(), va.float(), rtol=1e-4, atol=1e-5, msg=k)


def test_trainer_on_geister_episodes_with_gpu_replay():
    """Dict observations + burn-in + recurrent net + GPU-resident replay through the Trainer thread protocol."""
    import threading
    from handyrl_b200.train import Trainer
    from handyrl_b200.nets import GatedBoardNet
    with open(os.path.join(GOLDEN, 'batch_cases.pkl'), 'rb') as f:
        case = pickle.load(f)['geister_burnin']
    args = dict(case['args'], batch_size=6, minimum_episodes=4, num_batchers=1, **{'lambda': 0.7},
                entropy_regularization=0.1, entropy_regularization_decay=0.1, policy_target='TD', value_target='TD',
                gpu_replay=True)
    net = GatedBoardNet(scalars=18, planes=7, board=(6, 6), width=8, actions=214)
    tr = Trainer(args, net)
    tr.episodes.extend(case['episodes'])
    th = threading.Thread(target=tr.run, daemon=True)
    th.start()
    model, steps = tr.update()
    assert steps >= 1 and not model.training
    assert all(torch.isfinite(p).all() for p in model.parameters())
    model2, steps2 = tr.update()
    assert steps2 > steps
    tr.stop()
    th.join(timeout=10)
    assert not th.is_alive()
