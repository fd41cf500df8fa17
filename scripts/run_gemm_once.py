"""Launch hrl_gemm_fused a few times at the tower shape (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from handyrl_b200 import nets, tower
M = 16384
eng = tower.FusedBoardNet(nets.tictactoe_net().cuda(), M, torch.device('cuda'))
D = eng.D
X = torch.randn(M, D, device='cuda'); Y = torch.randn(M, D, device='cuda'); W = torch.randn(D, D, device='cuda') * 0.1
out = torch.empty(M, D, device='cuda'); c = [torch.rand(D, device='cuda') for _ in range(5)]
mode = sys.argv[1] if len(sys.argv) > 1 else 'plain'
for _ in range(4):
    if mode == 'plain':
        eng._gemm(dict(t=X), dict(t=W), out, K=D, N=D)
    else:
        eng._gemm(dict(t=X, t2=Y, consts=(c[0], c[1], c[2])), dict(t=W, kmajor=False), out, K=D, N=D, epilogue='mask_stats',
                  ep=dict(y=Y, scale=c[0], shift=c[1], mean=c[2], rstd=c[3]))
torch.cuda.synchronize()
