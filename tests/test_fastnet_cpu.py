"""The small-board rewrite pass computes the same function as the stock modules (float64 on CPU)."""
import copy

import pytest
import torch

from handyrl_b200 import fastnet
from handyrl_b200.nets import BoardNet, tictactoe_net


@pytest.mark.parametrize('board', [(3, 3), (6, 6), (4, 5)])
def test_rewritten_net_matches_stock_modules(board):
    torch.manual_seed(0)
    ref = BoardNet(planes=3, board=board, width=8, depth=2, actions=7, return_head=True).double().train()
    fast = copy.deepcopy(ref)
    assert fastnet.optimize_small_boards(fast) == 1 + 2 * 2 + 3      # stem, 2 x (conv + bn), 3 squeeze convs
    x = torch.randn(37, 3, *board, dtype=torch.float64)
    outs_r, outs_f = ref(x), fast(x)
    for k in outs_r:
        torch.testing.assert_close(outs_f[k], outs_r[k], rtol=1e-10, atol=1e-10)
    sum(o.square().sum() for o in outs_r.values()).backward()
    sum(o.square().sum() for o in outs_f.values()).backward()
    for (k, p), (_, q) in zip(ref.named_parameters(), fast.named_parameters()):
        torch.testing.assert_close(q.grad, p.grad, rtol=1e-8, atol=1e-10, msg=k)
    for (k, b), (_, c) in zip(ref.named_buffers(), fast.named_buffers()):
        torch.testing.assert_close(c, b, rtol=1e-10, atol=1e-12, msg=k)       # running stats, num_batches_tracked
    assert list(ref.state_dict()) == list(fast.state_dict())


def test_eval_mode_and_restore():
    torch.manual_seed(1)
    net = tictactoe_net().double()
    x = torch.randn(5, 3, 3, 3, dtype=torch.float64)
    net.eval()
    want = net(x)
    fastnet.optimize_small_boards(net)
    got = net(x)
    for k in want:
        torch.testing.assert_close(got[k], want[k], rtol=1e-10, atol=1e-10)
    fastnet.restore(net)
    assert all(type(m) in (torch.nn.Conv2d, torch.nn.BatchNorm2d) for m in net.modules()
               if isinstance(m, (torch.nn.Conv2d, torch.nn.BatchNorm2d)))


def test_large_inputs_fall_through_to_cudnn_path():
    conv = torch.nn.Conv2d(2, 3, 3, padding=1).double()
    x = torch.randn(2, 2, 16, 16, dtype=torch.float64)
    want = conv(x)
    m = torch.nn.Sequential(conv)
    fastnet.optimize_small_boards(m)
    torch.testing.assert_close(m(x), want)
