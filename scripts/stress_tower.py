"""Determinism stress: the fused tower engine run repeatedly on the same inputs must give bit-identical outputs and gradients."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from handyrl_b200 import nets, tower
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_tower_gpu import CASES
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for name, case in sorted(CASES.items()):
    torch.manual_seed(1)
    net = nets.BoardNet(**case['kw']).cuda().train()
    M = case['M']
    x = (torch.rand(M, case['kw']['planes'], *case['kw']['board'], device='cuda') < 0.4).float()
    eng = tower.FusedBoardNet(net, M, torch.device('cuda'))
    for p in net.parameters():
        p.grad = torch.zeros_like(p)
    first, bad = None, {}
    for i in range(reps):
        out = eng.forward(x)
        dout = {k: torch.ones_like(v) * 0.5 for k, v in out.items()}
        eng.backward(dout['policy'], dout['value'], dout.get('return'))
        torch.cuda.synchronize()
        snap = {('out', k): v.clone() for k, v in out.items()}
        snap.update({('grad', k): p.grad.clone() for k, p in net.named_parameters()})
        if first is None:
            first = snap
        else:
            for k in snap:
                if not torch.equal(snap[k], first[k]):
                    bad[k] = bad.get(k, 0) + 1
    print(name, 'reps', reps, 'MISMATCHES' if bad else 'deterministic', bad)
