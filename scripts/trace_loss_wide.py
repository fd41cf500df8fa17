"""Phase stamps (clock64 of CTA 0) of the wide-row loss kernel at the cfg5shard shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from handyrl_b200 import ops
from handyrl_b200.synthetic import synthetic_batch, synthetic_outputs
w = bench.WORKLOADS['cfg5shard']; args = bench.train_args(w)
b = synthetic_batch(w['B'], w['T'], w['P'], w['A'], turn_based=True, observation=False, seed=1, with_obs=False)
o = synthetic_outputs(b, seed=2)
o = {k: v.cuda() for k, v in o.items()}; b = {k: v.cuda() for k, v in b.items()}
for bf16 in (False, True):
    oo = dict(o, policy=o['policy'].to(torch.bfloat16)) if bf16 else o
    tr = torch.zeros(32, dtype=torch.int64, device='cuda')
    for _ in range(3):
        ops.loss_fwd_bwd(oo, b, args, tuning={'variant': 'bulk', 'trace': tr})
    torch.cuda.synchronize()
    t = tr.cpu().tolist()
    base = t[0]
    names = {0: 'start', 1: 'small tensors staged + baselines', 7: 'chunk0 wait', 8: 'chunk0 landed', 9: 'chunk1 wait', 10: 'chunk1 landed',
             11: 'chunk2 wait', 12: 'chunk2 landed', 13: 'chunk3 wait', 14: 'chunk3 landed', 2: 'statistics pass done', 3: 'cluster sync + row epilogue',
             4: 'targets and losses', 5: 'partials published', 6: 'gradients stored, end'}
    print('bf16 =', bf16)
    for k in sorted(names, key=lambda k: t[k]):
        if t[k]:
            print('  %-36s %7d cycles' % (names[k], t[k] - base))
