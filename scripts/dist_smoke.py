"""2-rank smoke of LearnerStep with NCCL (prints a line per stage so a hang can be located)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from handyrl_b200 import multigpu as hdist
from handyrl_b200.nets import tictactoe_net
from handyrl_b200.synthetic import synthetic_batch
from handyrl_b200.train import LearnerStep

def log(*a):
    print('[rank %s %.1fs]' % (os.environ.get('RANK'), time.time() - T0), *a, flush=True)

T0 = time.time()
rank, world, local = hdist.init_from_env()
log('init done', world)
torch.backends.cudnn.allow_tf32 = False
args = {'turn_based_training': True, 'observation': False, 'gamma': 0.8, 'lambda': 0.7, 'burn_in_steps': 0, 'forward_steps': 32,
        'entropy_regularization': 0.1, 'entropy_regularization_decay': 0.1, 'policy_target': 'UPGO', 'value_target': 'VTRACE'}
torch.manual_seed(0)
mode = sys.argv[1] if len(sys.argv) > 1 else 'graph'
use_graph = mode != 'eager'
split = mode == 'split'
stepper = LearnerStep(tictactoe_net(), args, synthetic_batch(64, 32, 2, 9, seed=rank), lr=1e-4, device=torch.device('cuda', local),
                      process_group=dist.group.WORLD if world > 1 else None, use_graph=use_graph, time_loss_kernel=split)
log('stepper built')
t = torch.ones(1, device='cuda'); dist.all_reduce(t); torch.cuda.synchronize(); log('eager allreduce ok', float(t))
pk = stepper.new_packed().fill(synthetic_batch(64, 32, 2, 9, seed=10 + rank))
for i in range(5):
    stepper.step(pk)
    log('step', i, 'enqueued')
    print(stepper.read_losses(), flush=True)
dist.barrier(); torch.cuda.synchronize()
log('done')
stepper.close()
log('closed')
dist.destroy_process_group()
log('destroyed')
