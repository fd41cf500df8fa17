"""GPU parity of the fused clip + Adam step (train.py:370-371) against torch.optim.Adam and the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('n', [29006, 231604, 1000003])
def test_clip_adam_matches_torch_and_oracle(n):
    from handyrl_b200 import ops
    from oracle import oracle
    rng = np.random.default_rng(n)
    p0 = rng.standard_normal(n).astype(np.float32)
    # torch reference (what the reference learner runs)
    tp = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    topt = torch.optim.Adam([tp], lr=3e-4, weight_decay=1e-5)
    # oracle state
    op, om, ov = p0.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    # ours
    dp = torch.nn.Parameter(torch.from_numpy(p0.copy()).cuda())
    opt = ops.FlatAdam([dp], lr=3e-4, weight_decay=1e-5, max_norm=4.0)
    for step in range(5):
        g = (rng.standard_normal(n) * (0.2 if step % 2 else 0.001)).astype(np.float32)
        tp.grad = torch.from_numpy(g.copy())
        ref_norm = float(torch.nn.utils.clip_grad_norm_([tp], 4.0))
        topt.step()
        onorm = oracle.clip_adam(op, g, om, ov, 3e-4, step)
        opt.zero_grad()
        dp.grad.copy_(torch.from_numpy(g).cuda())
        opt.step()
        torch.cuda.synchronize()
        # the fp64 norm is the truth; torch's single-thread fp32 accumulation is itself ~1e-5 off at n = 1e6
        true_norm = float(np.sqrt((g.astype(np.float64) ** 2).sum()))
        assert abs(float(opt.grad_norm) - true_norm) <= 1e-6 * true_norm
        assert abs(onorm - true_norm) <= 1e-6 * true_norm
        assert abs(float(opt.grad_norm) - ref_norm) <= 5e-5 * ref_norm
        np.testing.assert_allclose(dp.detach().cpu().numpy(), tp.detach().numpy(), rtol=0, atol=1e-6)
        np.testing.assert_allclose(dp.detach().cpu().numpy(), op, rtol=0, atol=1e-6)
    assert int(opt.step_count) == 5


def test_lr_is_read_from_device():
    from handyrl_b200 import ops
    p = torch.nn.Parameter(torch.ones(1024, device='cuda'))
    opt = ops.FlatAdam([p], lr=0.0)
    p.grad.fill_(1.0)
    opt.step()
    torch.cuda.synchronize()
    assert torch.all(p == 1.0)          # lr 0: nothing moves
    opt.set_lr(1e-2)
    opt.step()
    torch.cuda.synchronize()
    assert torch.all(p < 1.0)


def test_peer_allreduce_kernel_world1_matches_sumsq():
    """hrl_peer_allreduce_sumsq with a single rank: the flag protocol must not dead-lock, the reduced bucket equals
    the input, the partials equal hrl_grad_sumsq's over the first n_norm floats, and repeated calls keep working
    (epochs advance).  The multi-rank form: tests/test_multi_gpu.py (world >= 2) and bench.py --gpus N."""
    import ctypes as C
    from handyrl_b200 import ops
    from handyrl_b200._capi import lib, check
    n, n_norm, world = 4096 + 8, 4096, 1
    bucket = torch.zeros(n + 2 * world + 64, device='cuda')
    peer_ptrs = torch.tensor([bucket.data_ptr()], dtype=torch.int64, device='cuda')
    out = torch.zeros(n, device='cuda')
    npart = lib().hrl_sumsq_num_partials()
    partials, ref_partials = torch.zeros(npart, device='cuda'), torch.zeros(npart, device='cuda')
    epoch = torch.zeros(1, dtype=torch.int32, device='cuda')
    ticket = torch.zeros(1, dtype=torch.int32, device='cuda')
    status = torch.zeros(1, dtype=torch.int32, device='cuda')
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    for it in range(3):
        bucket[:n].normal_()
        check(lib().hrl_peer_allreduce_sumsq(p(out), p(peer_ptrs), n, world, 0, n, n_norm, p(partials), p(epoch), p(ticket), p(status), stream))
        torch.cuda.synchronize()
        assert torch.equal(out, bucket[:n])
        assert int(epoch) == it + 1 and int(ticket) == 0 and int(status) == 0
        want = float((bucket[:n_norm].double() ** 2).sum())
        assert abs(float(partials.double().sum()) - want) <= 1e-5 * want


@pytest.mark.parametrize('channels_last', [False, True], ids=['nchw', 'nhwc'])
@pytest.mark.parametrize('shape', [(16384, 32, 3, 3), (1000, 8, 6, 6), (37, 5, 4, 5), (520, 32, 7, 11)])
def test_fused_batchnorm_matches_torch(shape, channels_last):
    """hrl_bn_train_fwd / _bwd vs nn.BatchNorm2d in training mode: output, input/affine gradients, running statistics,
    for NCHW and channels-last activations (the layout between cuDNN's NHWC convolutions; 7x11 = the Hungry Geese board)."""
    from handyrl_b200 import fastnet
    torch.manual_seed(0)
    N, C, H, W = shape
    ref = torch.nn.BatchNorm2d(C).cuda().train()
    with torch.no_grad():
        ref.weight.uniform_(0.5, 1.5)
        ref.bias.uniform_(-0.5, 0.5)
    import copy
    fast = torch.nn.Sequential(copy.deepcopy(ref))
    assert fastnet.optimize_small_boards(fast) == 1
    old = torch.backends.cudnn.enabled
    for it in range(2):
        x = (torch.randn(shape, device='cuda') * 2 + 0.5)
        g = torch.randn(shape, device='cuda')
        if channels_last:
            x, g = x.contiguous(memory_format=torch.channels_last), g.contiguous(memory_format=torch.channels_last)
        xr, xf = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        yr = ref(xr)
        yf = fast(xf)
        assert yf.is_contiguous(memory_format=torch.channels_last if channels_last else torch.contiguous_format)
        yr.backward(g)
        yf.backward(g)
        torch.testing.assert_close(yf, yr, rtol=1e-5, atol=2e-5)
        torch.testing.assert_close(xf.grad, xr.grad, rtol=1e-4, atol=2e-5)
        torch.testing.assert_close(fast[0].weight.grad, ref.weight.grad, rtol=1e-4, atol=1e-3)
        torch.testing.assert_close(fast[0].bias.grad, ref.bias.grad, rtol=1e-4, atol=1e-3)
        torch.testing.assert_close(fast[0].running_mean, ref.running_mean, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(fast[0].running_var, ref.running_var, rtol=1e-5, atol=1e-6)
        assert int(fast[0].num_batches_tracked) == int(ref.num_batches_tracked) == it + 1
    torch.backends.cudnn.enabled = old
