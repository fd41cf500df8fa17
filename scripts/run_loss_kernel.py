"""Launch the fused loss kernel a few times at a named shape (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from handyrl_b200 import ops
from handyrl_b200.synthetic import synthetic_batch, synthetic_outputs

SHAPES = {'cfg2': (512, 32, 2, 9, True, False), 'cfg2sim': (512, 32, 2, 9, False, False), 'cfg5': (512, 64, 2, 512, True, False),
          'cfg3': (256, 20, 2, 214, True, True), 'cfg4': (1024, 32, 4, 4, False, False)}
name = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, T, P, A, tb, obs = SHAPES[name]
args = {'turn_based_training': tb, 'observation': obs, 'gamma': 0.8, 'lambda': 0.7, 'burn_in_steps': 0,
        'entropy_regularization': 0.1, 'entropy_regularization_decay': 0.1, 'policy_target': 'UPGO', 'value_target': 'VTRACE'}
b = synthetic_batch(B, T, P, A, turn_based=tb, observation=obs, seed=0, with_obs=False)
o = synthetic_outputs(b, seed=1)
o = {k: v.cuda() for k, v in o.items()}
b = {k: v.cuda() for k, v in b.items()}
buf = ops.LossBuffers(B, T, P, b['action_mask'].shape[2], A, True, False, 'cuda')
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
trace = torch.zeros(32, dtype=torch.int64, device='cuda')
tuning = {'trace': trace} if os.environ.get('TRACE') else None
if os.environ.get('VARIANT'):
    tuning = dict(tuning or {}, variant=os.environ['VARIANT'])
mode = os.environ.get('FLUSH', 'read')
flush.fill_(1)
torch.cuda.synchronize()
for _ in range(reps):
    if mode == 'write':
        flush.fill_(1)          # evict L2 with dirty lines (their write-back lands inside the next kernel)
    elif mode == 'read':
        flush.view(torch.int32).sum()   # evict L2 with clean lines
    ops.loss_fwd_bwd(o, b, args, buffers=buf, tuning=tuning)
torch.cuda.synchronize()
print(name, buf.losses.tolist())
if os.environ.get('TRACE'):
    t = trace.cpu().tolist()
    print('trace cycles since start:', [x - t[0] for x in t[1:7]], 'consumer (wait-begin, data-ready) per chunk:', [x - t[0] for x in t[7:15]], 'phase2 (2a.2, 2b, 2c starts):', [x - t[0] for x in t[15:18]], 'phase3 (3b start, 3b end):', [x - t[0] for x in t[18:20]])
