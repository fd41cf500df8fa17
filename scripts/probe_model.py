"""Scratch probe: net fwd+bwd time at the cfg2 shape under different cuDNN settings, and the loss kernel alone."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from handyrl_b200 import nets, ops
from handyrl_b200.synthetic import synthetic_batch, synthetic_outputs

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
N = 512 * 32
x = (torch.rand(N, 3, 3, 3, device='cuda') < 0.3).float()


def run(tag, net, inp, iters=20):
    net.train()
    def step():
        o = net(inp, None)
        (o['policy'].sum() + o['value'].sum()).backward()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    print('%-40s %8.3f ms / fwd+bwd' % (tag, e0.elapsed_time(e1) / iters), flush=True)


torch.manual_seed(0)
run('default heuristics', nets.tictactoe_net().cuda(), x)
torch.backends.cudnn.benchmark = True
run('cudnn.benchmark', nets.tictactoe_net().cuda(), x)
run('benchmark + channels_last', nets.tictactoe_net().cuda().to(memory_format=torch.channels_last), x.contiguous(memory_format=torch.channels_last))
torch.backends.cudnn.enabled = False
run('cudnn disabled (native kernels)', nets.tictactoe_net().cuda(), x)
torch.backends.cudnn.enabled = True

# loss kernel alone
args = {'turn_based_training': True, 'observation': False, 'gamma': 0.8, 'lambda': 0.7, 'burn_in_steps': 0,
        'entropy_regularization': 0.1, 'entropy_regularization_decay': 0.1, 'policy_target': 'UPGO', 'value_target': 'VTRACE'}
for (B, T, P, A, tb) in [(512, 32, 2, 9, True), (512, 32, 2, 9, False), (512, 64, 2, 512, True), (256, 20, 2, 214, True), (1024, 32, 4, 4, False)]:
    n = 8 if A > 100 else 64
    sets = []
    for i in range(n):
        b = synthetic_batch(B, T, P, A, turn_based=tb, seed=i, with_obs=False)
        o = synthetic_outputs(b, seed=100 + i)
        sets.append(({k: v.cuda() for k, v in o.items()}, {k: v.cuda() for k, v in b.items()},
                     ops.LossBuffers(B, T, P, b['action_mask'].shape[2], A, True, False, 'cuda')))
    for o, b, buf in sets:
        ops.loss_fwd_bwd(o, b, args, buffers=buf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 200
    e0.record()
    for i in range(reps):
        o, b, buf = sets[i % n]
        ops.loss_fwd_bwd(o, b, args, buffers=buf)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    from handyrl_b200.synthetic import bytes_per_cell
    by = bytes_per_cell(P, b['action_mask'].shape[2], A, T, 0) * B * T
    print('loss kernel B=%d T=%d P=%d A=%d Pa=%d: %8.2f us  %7.1f GB/s algorithmic (%.2f MB)' % (B, T, P, A, b['action_mask'].shape[2], us, by / us / 1e3, by / 1e6), flush=True)
