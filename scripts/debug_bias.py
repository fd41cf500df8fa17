import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from handyrl_b200 import ops, nets, fastnet
torch.manual_seed(0)
a = torch.rand(4096, 288, device='cuda'); b = torch.rand(288, 288, device='cuda')
for sa, sb, tag in ((1, 1, '+ +'), (-1, 1, '- +'), (1, -1, '+ -')):
    got = ops.gemm_tf32x3(sa * a, sb * b).double()
    want = (sa * a).double() @ (sb * b).double().t()
    err = (got - want)
    print(tag, 'mean signed err / |want|: %.3e   mean |err|/|want| %.3e' % ((err / want.abs()).mean().item(), (err.abs() / want.abs()).mean().item()))
a = torch.randn(4096, 288, device='cuda'); b = torch.randn(288, 288, device='cuda')
got = ops.gemm_tf32x3(a, b).double(); want = a.double() @ b.double().t()
err = got - want
print('randn: mean signed err %.3e, mean |err| %.3e, mean |want| %.3e; corr with sign(want): %.3e' % (err.mean().item(), err.abs().mean().item(), want.abs().mean().item(), (err * want.sign()).mean().item()))
torch.backends.cuda.matmul.allow_tf32 = False
got32 = (a @ b.t()).double(); e32 = got32 - want
print('fp32 cublas: mean signed err %.3e, mean |err| %.3e' % (e32.mean().item(), e32.abs().mean().item()))
# module path (fastnet on tensor cores) vs float64, same config as the failing tower case
kw = dict(planes=3, board=(3, 3), width=32, depth=3, actions=9)
torch.manual_seed(1)
ref = nets.BoardNet(**kw).double().cuda().train()
for use_tc in (True, False):
    fast = copy.deepcopy(ref).float()
    if use_tc:
        fastnet.optimize_small_boards(fast)
    x = (torch.rand(300, 3, 3, 3, device='cuda') < 0.4).float()
    out = fast(x); want_o = ref(x.double())
    g = torch.Generator().manual_seed(5)
    dout = {k: torch.randn(v.shape, generator=g).cuda() for k, v in out.items()}
    ref.zero_grad(); sum((want_o[k] * dout[k].double()).sum() for k in want_o).backward()
    sum((out[k] * dout[k]).sum() for k in out).backward()
    print('module path, tensor cores' if use_tc else 'module path, stock PyTorch fp32 (cuDNN)')
    for (k, pr), (_, pf) in zip(ref.named_parameters(), fast.named_parameters()):
        print('  %-28s rel err %.2e' % (k, (pf.grad.double() - pr.grad).abs().max().item() / (pr.grad.abs().max().item() + 1e-12)))
