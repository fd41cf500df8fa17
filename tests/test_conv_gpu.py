"""Convolutions over a board as implicit tensor-core products (ops.conv_implicit: hrl_gemm_fused conv_mode 1 / 2, hrl_conv_pack,
hrl_conv_wgrad_reduce) against float64 F.conv2d: zero `same` padding (Geister's ConvLSTM cells, reference geister.py:18-56) and
wrap-around padding (Hungry Geese's TorusConv2d, reference hungry_geese.py:20-37); output, input gradient, weight and bias gradient."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    # N, Cin, Cout, H, W, kh, kw, wrap, bias
    (37, 64, 128, 6, 6, 3, 3, False, True),       # Geister ConvLSTM gates: [x, h] 64 maps -> 4 x 32 gate maps
    (21, 32, 32, 7, 11, 3, 3, True, True),        # Hungry Geese torus block
    (9, 36, 20, 6, 6, 3, 3, False, False),        # channel counts that are not multiples of 32: padded chunks
    (130, 8, 12, 3, 3, 3, 3, False, True),        # fewer cells (9) than a chunk has elements
    (5, 16, 288, 5, 4, 1, 3, True, False),        # one-row kernel, widest operand tile
    (3, 288, 8, 4, 4, 3, 1, False, True),
]


@pytest.mark.parametrize('N,Cin,Cout,H,W,kh,kw,wrap,bias', CASES)
@pytest.mark.parametrize('channels_last', [True, False])
def test_conv_implicit_matches_float64(N, Cin, Cout, H, W, kh, kw, wrap, bias, channels_last):
    from handyrl_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(N * 1000 + Cin)
    x = torch.randn(N, Cin, H, W, device='cuda', generator=g)
    w = torch.randn(Cout, Cin, kh, kw, device='cuda', generator=g) * 0.2
    b = torch.randn(Cout, device='cuda', generator=g) if bias else None
    dy = torch.randn(N, Cout, H, W, device='cuda', generator=g)
    if channels_last:
        x, dy = x.contiguous(memory_format=torch.channels_last), dy.contiguous(memory_format=torch.channels_last)
    assert ops.conv_implicit_supported(x, w)
    xs, ws = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    bs = b.clone().requires_grad_(True) if bias else None
    ops.conv_weights_changed()
    y = ops.conv_implicit(xs, ws, bs, wrap)
    y.backward(dy)

    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    bd = b.double().requires_grad_(True) if bias else None
    if wrap:
        xp = F.pad(xd, (kw // 2, kw // 2, kh // 2, kh // 2), mode='circular')
        yd = F.conv2d(xp, wd, bd)
    else:
        yd = F.conv2d(xd, wd, bd, padding=(kh // 2, kw // 2))
    yd.backward(dy.double())

    def close(got, want, what, tol=3e-6, red=Cin * kh * kw):
        # 3xTF32 products: ~1e-6 of sum |a||b|; the bound below is in units of the result's largest magnitude
        scale = want.abs().max().item() + 1e-12
        err = (got.double() - want).abs().max().item()
        assert err <= tol * scale * (1 + red ** 0.5 / 8), (what, err, scale)

    assert y.shape == yd.shape
    close(y, yd, 'output')
    close(xs.grad, xd.grad, 'input gradient', red=Cout * kh * kw)
    close(ws.grad, wd.grad, 'weight gradient', red=N * H * W)
    if bias:
        close(bs.grad, bd.grad, 'bias gradient', tol=1e-5, red=1)


def test_rewritten_modules_use_the_implicit_products():
    """fastnet routes nn.Conv2d (zeros / circular `same` padding) over boards too large for the dense form to conv_implicit."""
    from handyrl_b200 import fastnet, ops
    torch.manual_seed(0)
    for mode in ('zeros', 'circular'):
        conv = torch.nn.Conv2d(16, 24, 3, padding=1, padding_mode=mode).cuda()
        ref = torch.nn.Conv2d(16, 24, 3, padding=1, padding_mode=mode).cuda().double()
        ref.load_state_dict({k: v.double() for k, v in conv.state_dict().items()})
        net = torch.nn.Sequential(conv)
        assert fastnet.optimize_small_boards(net) == 1
        x = torch.randn(11, 16, 6, 6, device='cuda')
        before = ops.LAUNCHES['n']
        fastnet.new_step()
        y = net(x)
        assert ops.LAUNCHES['n'] > before
        assert (y.double() - ref(x.double())).abs().max().item() < 1e-4


def test_reference_style_torus_convolution_is_recognised():
    """The reference writes the wrap-around convolution as edge concatenation + an unpadded convolution (hungry_geese.py:24-37);
    fastnet recognises the module by structure and runs it as the wrap-around implicit product, forward and backward."""
    import torch.nn as nn
    from handyrl_b200 import fastnet, ops

    class TorusConv2d(nn.Module):          # restated from the reference's forward (not imported: no kaggle_environments here)
        def __init__(self, input_dim, output_dim, kernel_size, bn):
            super().__init__()
            self.edge_size = (kernel_size[0] // 2, kernel_size[1] // 2)
            self.conv = nn.Conv2d(input_dim, output_dim, kernel_size=kernel_size)
            self.bn = nn.BatchNorm2d(output_dim) if bn else None

        def forward(self, x):
            h = torch.cat([x[:, :, :, -self.edge_size[1]:], x, x[:, :, :, :self.edge_size[1]]], dim=3)
            h = torch.cat([h[:, :, -self.edge_size[0]:], h, h[:, :, :self.edge_size[0]]], dim=2)
            h = self.conv(h)
            return self.bn(h) if self.bn is not None else h

    torch.manual_seed(3)
    ref = TorusConv2d(32, 32, (3, 3), False).cuda().double()
    net = TorusConv2d(32, 32, (3, 3), False).cuda()
    net.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    assert fastnet.optimize_small_boards(net) >= 1
    x = torch.randn(9, 32, 7, 11, device='cuda')
    xs, xd = x.clone().requires_grad_(True), x.double().requires_grad_(True)
    dy = torch.randn(9, 32, 7, 11, device='cuda')
    before = ops.LAUNCHES['n']
    fastnet.new_step()
    y = net(xs)
    y.backward(dy)
    assert ops.LAUNCHES['n'] >= before + 3
    yd = ref(xd)
    yd.backward(dy.double())
    assert (y.double() - yd).abs().max().item() < 1e-4
    assert (xs.grad.double() - xd.grad).abs().max().item() < 1e-4
    assert (net.conv.weight.grad.double() - ref.conv.weight.grad).abs().max().item() < 2e-3 * ref.conv.weight.grad.abs().max().item()
    fastnet.restore(net)
    assert type(net) is TorusConv2d


@pytest.mark.parametrize('wrap,bias,steps', [(False, True, 5), (True, False, 5), (False, True, 70)])
def test_deferred_weight_gradients_of_a_shared_convolution(wrap, bias, steps):
    """A recurrent cell applies ONE convolution several times per backward pass.  Inside ops.deferred_weight_gradients() the
    applications only record their (dy, x) pairs; one segmented product per weight (+ its ones row for the bias) then adds the
    summed gradient into weight.grad / bias.grad: same result as per-application products, and as float64 autograd."""
    from handyrl_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(21)
    N, Cin, Cout, H, W = 23, 16, 16, 6, 6
    w = torch.nn.Parameter(torch.randn(Cout, Cin, 3, 3, device='cuda', generator=g) * 0.2)
    b = torch.nn.Parameter(torch.randn(Cout, device='cuda', generator=g)) if bias else None
    x0 = torch.randn(N, Cin, H, W, device='cuda', generator=g)

    def run(conv, x):
        if steps > 10:        # many applications side by side (a long chain of them is chaotic: nothing to compare)
            return sum(torch.tanh(conv(x * (0.5 + i / steps))) for i in range(steps))
        for _ in range(steps):                       # the output of one application feeds the next (as h does in a ConvLSTM)
            x = torch.tanh(conv(x))
        return x

    dy = torch.randn(N, Cout, H, W, device='cuda', generator=g)
    grads = {}
    for mode in ('immediate', 'deferred'):
        w.grad = torch.full_like(w, 0.5)                         # the flush must ADD to what is there (other uses of the weight)
        if bias:
            b.grad = None
        ops.conv_weights_changed()
        y = run(lambda t: ops.conv_implicit(t, w, b, wrap), x0)
        before = ops.LAUNCHES['n']
        if mode == 'deferred':
            with ops.deferred_weight_gradients():
                y.backward(dy)
        else:
            y.backward(dy)
        grads[mode] = (w.grad.clone(), None if not bias else b.grad.clone(), ops.LAUNCHES['n'] - before)
    assert grads['deferred'][2] < grads['immediate'][2]          # `steps` products + reductions became 1 + 1 (2 + 2 beyond 64 pairs)
    wd = w.detach().double().requires_grad_(True)
    bd = b.detach().double().requires_grad_(True) if bias else None

    def conv64(t):
        if wrap:
            return torch.nn.functional.conv2d(torch.nn.functional.pad(t, (1, 1, 1, 1), mode='circular'), wd, bd)
        return torch.nn.functional.conv2d(t, wd, bd, padding=1)
    run(conv64, x0.double()).backward(dy.double())
    scale = wd.grad.abs().max().item()
    for mode in grads:
        assert (grads[mode][0].double() - 0.5 - wd.grad).abs().max().item() <= 3e-5 * scale, mode
        if bias:
            assert (grads[mode][1].double() - bd.grad).abs().max().item() <= 3e-5 * bd.grad.abs().max().item(), mode
    assert (grads['deferred'][0] - grads['immediate'][0]).abs().max().item() <= 2e-5 * scale
