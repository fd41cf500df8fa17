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
            # step 0 sees the reference's weights; later steps see weights that went through Adam's sign-like first updates,
            # which turn rounding differences of the gradient into lr-sized weight differences (cuDNN's autotuner picks
            # convolution algorithms with different summation orders from run to run: 1.3e-4 * scale was seen at step 2)
            assert abs(got[k] - v) <= (1 + s) * 1e-4 * scale + 1e-4, (s, k, got[k], v)
        assert got['dcnt'] == ref['dcnt']
        assert abs(float(stepper.opt.grad_norm) - ref['grad_norm']) <= 2e-3 * ref['grad_norm']
    final = stepper.cpu_state_dict()
    for (k, v), (kr, vr) in zip(final.items(), c['state3'].items()):
        if noise_driven(c, k):
            continue
        if v.dtype.is_floating_point:
            # Adam's first steps move a weight by ~lr*sign(g): an element whose gradient is at rounding-noise level may
            # take the other sign in an equally valid fp32 run -> allow <= 0.1% of a tensor to differ by up to 2*lr per step
            bad = np.abs(v.numpy() - vr) > 5e-5 + 1e-3 * np.abs(vr)
            assert bad.mean() <= 1e-3, '%s/%s: %d of %d elements differ' % (k, kr, bad.sum(), bad.size)
            np.testing.assert_allclose(v.numpy(), vr, rtol=1e-3, atol=2 * c['lr'] * len(c['steps']) + 5e-5, err_msg='%s/%s' % (k, kr))
        else:
            assert int(v) == int(vr)


@pytest.mark.parametrize('name', sorted(STEP_CASES))
def test_three_learner_steps_strict_fp32_mode(name):
    """train_args['tensor_cores'] = False: the net's products stay on fp32 SIMT kernels; the reference's three optimiser steps
    are reproduced to the tolerances of plain fp32 reordering."""
    from handyrl_b200.nets import tictactoe_net, load_state_by_order
    from handyrl_b200.synthetic import synthetic_batch
    from handyrl_b200.train import LearnerStep
    c = STEP_CASES[name]
    B, T, P, A = c['dims']
    args = dict(c['args'], tensor_cores=False)
    net = load_state_by_order(tictactoe_net(), c['state0'])
    mk = lambda s: synthetic_batch(B, T, P, A, turn_based=args['turn_based_training'], observation=args['observation'], seed=40 + s)
    stepper = LearnerStep(net, args, mk(0), lr=c['lr'])
    assert stepper.engine is None and not stepper.tensor_cores
    for s, ref in enumerate(c['steps']):
        stepper.step(stepper.new_packed().fill(mk(s)))
        got = stepper.read_losses()
        for k, v in ref['losses'].items():
            assert abs(got[k] - v) <= 2e-5 * abs(v) + 2e-5, (s, k, got[k], v)
        assert abs(float(stepper.opt.grad_norm) - ref['grad_norm']) <= 1e-4 * ref['grad_norm']
    for (k, v), (kr, vr) in zip(stepper.cpu_state_dict().items(), c['state3'].items()):
        if v.dtype.is_floating_point:
            np.testing.assert_allclose(v.numpy(), vr, rtol=1e-4, atol=2e-5, err_msg='%s/%s' % (k, kr))


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
                gpu_replay=gpu_replay, num_gpus=1)
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
                gpu_replay=True, num_gpus=1)
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


def test_epoch_hand_off_never_synchronises_the_step_stream_and_workers_can_unpickle_it():
    """f-4 (train.py:385-387, 605-615): update() may not stall the step stream; the model it returns pickles from cached
    bytes and unpickles, in a process WITHOUT CUDA, to the plain module class in eval mode on the CPU."""
    import pickle as pk
    import subprocess
    import sys
    import threading
    import time
    from handyrl_b200.train import Trainer
    from handyrl_b200.nets import tictactoe_net, BoardNet
    with open(os.path.join(GOLDEN, 'batch_cases.pkl'), 'rb') as f:
        case = pickle.load(f)['tictactoe']
    args = dict(case['args'], batch_size=8, minimum_episodes=4, num_batchers=1, **{'lambda': 0.7}, num_gpus=1,
                entropy_regularization=0.1, entropy_regularization_decay=0.1, policy_target='UPGO', value_target='VTRACE')
    tr = Trainer(args, tictactoe_net())
    tr.episodes.extend(case['episodes'])
    th = threading.Thread(target=tr.run, daemon=True)
    th.start()
    tr.update()                                   # first epoch: builds and captures the step
    stream_cls = type(tr.stepper.stream)
    calls = []
    real = stream_cls.synchronize
    stream_cls.synchronize = lambda self: (calls.append(self), real(self))[1]
    try:
        steps_before = tr.steps
        model, steps = tr.update()
        time.sleep(0.05)
    finally:
        stream_cls.synchronize = real
    assert not any(s is tr.stepper.stream for s in calls)          # the step stream was never synchronised
    assert steps >= steps_before and type(model) is BoardNet and not model.training
    assert all(p.device.type == 'cpu' for p in model.parameters())
    assert tr.steps >= steps                                       # and training went on while we resolved the model
    t0 = time.perf_counter()
    blobs = [pk.dumps(model) for _ in range(200)]                  # Learner.server: one pickle per worker request
    per_pickle = (time.perf_counter() - t0) / 200
    assert per_pickle < 2e-3, per_pickle
    code = ('import sys, pickle, torch; sys.path.insert(0, %r); m = pickle.loads(sys.stdin.buffer.read()); '
            'assert not torch.cuda.is_available() and type(m).__name__ == "BoardNet" and not m.training; '
            'print("OUT", float(m(torch.ones(1, 3, 3, 3))["value"]))' % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    res = subprocess.run([sys.executable, '-c', code], input=blobs[0], capture_output=True, timeout=300,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=''))
    assert res.returncode == 0, res.stderr.decode()[-2000:]
    assert abs(float(res.stdout.decode().split('OUT')[1]) - float(model(torch.ones(1, 3, 3, 3))['value'])) < 1e-6
    # the weights handed over are exactly the learner's weights at that step boundary or later ones, never torn
    model2, _ = tr.update()
    assert any(not torch.equal(a, b) for a, b in zip(model.state_dict().values(), model2.state_dict().values()))
    tr.stop()
    th.join(timeout=10)
    assert not th.is_alive()


def test_flat_wire_episodes_through_the_gpu_replay_feeder_under_a_live_learner():
    """f-2 on the GPU + the Learner's side of the protocol: episodes in the flat wire format (moments dropped) arrive from
    another thread while the trainer runs, the deque is trimmed as Learner.feed_episodes does (train.py:476-483), and
    several epochs are handed off."""
    import threading
    import time
    from handyrl_b200.train import Trainer
    from handyrl_b200.nets import tictactoe_net
    from handyrl_b200.wire import pack_episode
    with open(os.path.join(GOLDEN, 'batch_cases.pkl'), 'rb') as f:
        case = pickle.load(f)['tictactoe']
    flat = [pack_episode(ep, drop_moments=True) for ep in case['episodes']]
    assert all(ep['moment'] == [] and 'flat' in ep for ep in flat)
    args = dict(case['args'], batch_size=8, minimum_episodes=4, maximum_episodes=12, num_batchers=1, **{'lambda': 0.7}, num_gpus=1,
                entropy_regularization=0.1, entropy_regularization_decay=0.1, policy_target='UPGO', value_target='VTRACE')
    tr = Trainer(args, tictactoe_net())
    tr.episodes.extend(flat[:5])
    stop = threading.Event()

    def learner_side():
        i = 0
        while not stop.is_set():
            tr.episodes.extend([dict(flat[i % len(flat)])])
            while len(tr.episodes) > args['maximum_episodes']:
                tr.episodes.popleft()
            i += 1
            time.sleep(0.002)

    th = threading.Thread(target=tr.run, daemon=True)
    th.start()
    feeder = threading.Thread(target=learner_side, daemon=True)
    feeder.start()
    seen = 0
    for _ in range(4):
        model, steps = tr.update()
        assert steps > seen and all(torch.isfinite(p).all() for p in model.parameters())
        seen = steps
    assert len(tr.gpu_batcher.replay) <= args['maximum_episodes'] and tr.gpu_batcher.fed > 5
    stop.set()
    feeder.join(timeout=5)
    tr.stop()
    th.join(timeout=10)
    assert not th.is_alive()
