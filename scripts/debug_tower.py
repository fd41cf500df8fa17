import copy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from handyrl_b200 import nets, tower
for M, kw in ((300, dict(planes=3, board=(3, 3), width=32, depth=3, actions=9)), (2048, dict(planes=3, board=(3, 3), width=32, depth=3, actions=9)),
              (300, dict(planes=3, board=(3, 3), width=32, depth=1, actions=9)), (300, dict(planes=3, board=(3, 3), width=16, depth=3, actions=9))):
    torch.manual_seed(1)
    ref = nets.BoardNet(**kw).double().cuda().train()
    fast = copy.deepcopy(ref).float()
    x = (torch.rand(M, kw['planes'], *kw['board'], device='cuda') < 0.4).float()
    eng = tower.FusedBoardNet(fast, M, torch.device('cuda'))
    for p in fast.parameters():
        p.grad = torch.zeros_like(p)
    out = eng.forward(x)
    want = ref(x.double())
    g = torch.Generator().manual_seed(5)
    dout = {k: torch.randn(v.shape, generator=g).cuda() for k, v in out.items()}
    sum((want[k] * dout[k].double()).sum() for k in want).backward()
    eng.backward(dout['policy'], dout['value'], dout.get('return'))
    torch.cuda.synchronize()
    print('M', M, kw)
    for (k, pr), (_, pf) in zip(ref.named_parameters(), fast.named_parameters()):
        print('  %-28s rel err %.2e  (max |g| %.3g)' % (k, (pf.grad.double() - pr.grad).abs().max().item() / (pr.grad.abs().max().item() + 1e-12), pr.grad.abs().max().item()))
