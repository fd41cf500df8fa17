"""Kernels around the user's net (csrc/net_kernel.cu, csrc/gemm_kernel.cu through autograd) against plain PyTorch."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('shape', [(32, 32, 3, 3, 3, 3), (5, 7, 3, 3, 4, 4), (2, 32, 1, 1, 3, 3), (4, 3, 5, 3, 3, 4)])
def test_board_dense_is_the_convolution_and_fold_is_its_adjoint(shape):
    from handyrl_b200 import ops
    Cout, Cin, kh, kw, H, W = shape
    g = torch.Generator().manual_seed(sum(shape))
    w = torch.randn((Cout, Cin, kh, kw), generator=g).cuda().requires_grad_(True)
    x = torch.randn((6, Cin, H, W), generator=g).cuda()
    dense = ops.board_dense(w, H, W)
    y = (x.reshape(6, -1).double() @ dense.double().t()).reshape(6, Cout, H, W)
    want = F.conv2d(x.double(), w.detach().double(), padding=(kh // 2, kw // 2))
    torch.testing.assert_close(y, want, rtol=1e-12, atol=1e-12)          # a pure re-indexing: exact
    probe = torch.randn(dense.shape, generator=torch.Generator().manual_seed(1)).cuda()
    (dense * probe).sum().backward()
    w2 = w.detach().clone().requires_grad_(True)
    cols = F.unfold(torch.eye(Cin * H * W, device='cuda').reshape(-1, Cin, H, W), (kh, kw), padding=(kh // 2, kw // 2))
    dense_ref = torch.einsum('ok,pkq->oqp', w2.reshape(Cout, -1), cols).reshape(Cout * H * W, Cin * H * W)
    torch.testing.assert_close(dense_ref, dense.detach(), rtol=0, atol=0)
    (dense_ref * probe).sum().backward()
    torch.testing.assert_close(w.grad, w2.grad, rtol=1e-5, atol=1e-5)


def test_linear_tc_autograd_matches_float64():
    from handyrl_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.randn((1500, 288), generator=g).cuda().requires_grad_(True)
    w = (0.1 * torch.randn((288, 288), generator=g)).cuda().requires_grad_(True)
    dy = torch.randn((1500, 288), generator=g).cuda()
    ops.linear_tc(x, w).backward(dy)
    xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    (xd @ wd.t()).backward(dy.double())
    for got, want, a, b in ((x.grad, xd.grad, dy, w.detach().t()), (w.grad, wd.grad, dy.t(), x.detach().t())):
        scale = a.double().abs() @ b.double().abs().t()
        assert ((got.double() - want).abs() / scale).max().item() < 3e-6


@pytest.mark.parametrize('board,chans', [((3, 3), (3, 32)), ((3, 3), (32, 32)), ((4, 4), (11, 8)), ((3, 3), (32, 2))])
def test_board_conv_on_tensor_cores_matches_conv2d(board, chans):
    from handyrl_b200 import fastnet
    torch.manual_seed(0)
    k = 3 if chans[1] > 2 else 1
    ref = nn.Conv2d(chans[0], chans[1], k, padding=k // 2).cuda()
    fast = nn.Conv2d(chans[0], chans[1], k, padding=k // 2).cuda()
    fast.load_state_dict(ref.state_dict())
    fastnet.optimize_small_boards(nn.Sequential(fast))
    ref = ref.double()          # float64 reference (cuDNN's autotuned fp32 algorithms are themselves only ~1e-3 accurate)
    x = torch.randn(700, chans[0], *board, device='cuda')
    xr, xf = x.double().requires_grad_(True), x.clone().requires_grad_(True)
    before = fastnet.BoardConv2d.dense_calls
    yr, yf = ref(xr), fast(xf)
    assert fastnet.BoardConv2d.dense_calls == before + 1
    dy = torch.randn_like(yf)
    yr.backward(dy.double())
    yf.backward(dy)
    torch.testing.assert_close(yf.double(), yr, rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(xf.grad.double(), xr.grad, rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(fast.weight.grad.double(), ref.weight.grad, rtol=1e-5, atol=1e-4)      # sums of 700 x cells terms
    torch.testing.assert_close(fast.bias.grad.double(), ref.bias.grad, rtol=1e-5, atol=1e-4)


def test_lstm_gates_kernel_matches_torch():
    from handyrl_b200 import ops
    g = torch.Generator().manual_seed(9)
    gates = torch.randn((37, 4 * 8, 6, 6), generator=g).cuda()
    c0 = torch.randn((37, 8, 6, 6), generator=g).cuda()
    dh, dc = torch.randn((37, 8, 6, 6), generator=g).cuda(), torch.randn((37, 8, 6, 6), generator=g).cuda()

    def torch_cell(gt, c):
        i, f, o, gg = gt.chunk(4, 1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        return torch.sigmoid(o) * torch.tanh(c), c

    for use_dc in (True, False):
        a, b = gates.clone().requires_grad_(True), c0.clone().requires_grad_(True)
        a2, b2 = gates.clone().requires_grad_(True), c0.clone().requires_grad_(True)
        h, c = ops.lstm_gates(a, b)
        h2, c2 = torch_cell(a2, b2)
        torch.testing.assert_close(h, h2, rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(c, c2, rtol=1e-6, atol=1e-6)
        ((h * dh).sum() + ((c * dc).sum() if use_dc else 0)).backward()
        ((h2 * dh).sum() + ((c2 * dc).sum() if use_dc else 0)).backward()
        torch.testing.assert_close(a.grad, a2.grad, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(b.grad, b2.grad, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('alternating', [True, False])
def test_hidden_mask_and_blend_kernels_match_reference_arithmetic(alternating):
    """train.py:152-158, 173 on one hidden leaf, mask taken as the strided slice observation_mask[:, t]."""
    from handyrl_b200 import ops
    g = torch.Generator().manual_seed(4)
    B, T, P = 9, 5, 2
    omask = (torch.rand((B, T, P, 1), generator=g) < 0.6).float().cuda()
    om = omask[:, 2]
    h = torch.randn((B, P, 4, 3, 3), generator=g).cuda()
    Pn = 1 if alternating else P
    nh = torch.randn((B, Pn, 4, 3, 3), generator=g).cuda()
    d1 = torch.randn((B, 4, 3, 3) if alternating else (B, P, 4, 3, 3), generator=g).cuda()
    d2 = torch.randn((B, P, 4, 3, 3), generator=g).cuda()
    gate = om.view(B, P, 1, 1, 1)
    ha, na = h.clone().requires_grad_(True), nh.clone().requires_grad_(True)
    hb, nb = h.clone().requires_grad_(True), nh.clone().requires_grad_(True)
    va = ops.hidden_visible(ha, om, alternating)
    vb = (hb * gate).sum(1) if alternating else hb * gate
    torch.testing.assert_close(va, vb, rtol=0, atol=0)
    ka = ops.hidden_blend(ha, na, om)
    kb = hb * (1 - gate) + nb * gate
    torch.testing.assert_close(ka, kb, rtol=0, atol=0)
    ((va * d1).sum() + (ka * d2).sum()).backward()
    ((vb * d1).sum() + (kb * d2).sum()).backward()
    torch.testing.assert_close(ha.grad, hb.grad, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(na.grad, nb.grad, rtol=1e-6, atol=1e-6)


def test_fused_conv_lstm_cell_rewrite_covers_reference_style_cells():
    """The rewrite pass recognises a ConvLSTM cell by structure (reference geister.py:18-56 has `hidden_dim`, ours
    `state_maps`) and keeps its function; restore() gives the original class back."""
    from handyrl_b200 import fastnet, nets

    class ReferenceStyleCell(nn.Module):
        def __init__(self):
            super().__init__()
            self.input_dim, self.hidden_dim = 5, 6
            self.conv = nn.Conv2d(11, 24, 3, padding=1)

        def forward(self, x, state):
            h, c = state
            i, f, o, g = torch.split(self.conv(torch.cat([x, h], dim=-3)), self.hidden_dim, dim=-3)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            return torch.sigmoid(o) * torch.tanh(c), c

    torch.manual_seed(2)
    torch.backends.cudnn.allow_tf32 = False
    for cell in (ReferenceStyleCell().cuda(), nets.ConvLstmCell(5, 6).cuda()):
        x = torch.randn(10, 5, 6, 6, device='cuda')
        st = (torch.randn(10, 6, 6, 6, device='cuda'), torch.randn(10, 6, 6, 6, device='cuda'))
        want = cell(x, st)
        holder = nn.Sequential(cell)
        assert fastnet.optimize_small_boards(holder) == 2             # the cell and its convolution
        assert type(cell).__name__.startswith('Fused')
        got = cell(x, st)
        torch.testing.assert_close(got[0], want[0], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(got[1], want[1], rtol=1e-5, atol=1e-6)
        fastnet.restore(holder)
        assert not type(cell).__name__.startswith('Fused')


def test_gate_and_hidden_kernels_take_channels_last_blocks_as_they_lie():
    """Channels-last gates (what the tensor-core convolution produces) and hidden leaves whose (C,H,W) blocks are channels-last go
    through the same kernels without a layout copy: results equal the NCHW call, outputs keep the layout, gradients too."""
    from handyrl_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(11)
    N, C, H, W = 10, 8, 6, 6
    gates = torch.randn(N, 4 * C, H, W, device='cuda', generator=g)
    c0 = torch.randn(N, C, H, W, device='cuda', generator=g)
    dh, dc = torch.randn(N, C, H, W, device='cuda', generator=g), torch.randn(N, C, H, W, device='cuda', generator=g)
    res = {}
    for cl in (False, True):
        fmt = torch.channels_last if cl else torch.contiguous_format
        gi = gates.clone(memory_format=fmt).requires_grad_(True)
        ci = c0.clone(memory_format=fmt).requires_grad_(True)
        h, c = ops.lstm_gates(gi, ci)
        assert h.is_contiguous(memory_format=fmt) and c.is_contiguous(memory_format=fmt)
        torch.autograd.backward([h, c], [dh, dc])
        res[cl] = (h, c, gi.grad, ci.grad)
    for a, b in zip(res[False], res[True]):
        assert torch.equal(a.contiguous(), b.contiguous())

    B, P = 5, 2
    hstate = torch.randn(B, P, C, H, W, device='cuda', generator=g)
    nh = torch.randn(B, P, C, H, W, device='cuda', generator=g)
    om = (torch.rand(B, P, 1, device='cuda', generator=g) < 0.6).float()
    dout = torch.randn(B, P, C, H, W, device='cuda', generator=g)
    as_cl = lambda t: t.flatten(0, 1).contiguous(memory_format=torch.channels_last).unflatten(0, (B, P))
    out = {}
    for cl in (False, True):
        hi = (as_cl(hstate) if cl else hstate.clone()).detach().requires_grad_(True)
        ni = (as_cl(nh) if cl else nh.clone()).detach().requires_grad_(True)
        vis = ops.hidden_visible(hi, om, False)
        mix = ops.hidden_blend(hi, ni, om)
        if cl:
            assert ops._block_layout(vis) == 'cl' and ops._block_layout(mix) == 'cl'
        (vis * dout).sum().backward(retain_graph=True)
        (mix * dout.flip(0)).sum().backward()
        out[cl] = (vis, mix, hi.grad, ni.grad)
    for a, b in zip(out[False], out[True]):
        assert torch.equal(a.contiguous(), b.contiguous())
