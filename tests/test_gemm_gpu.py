"""hrl_gemm_tf32x3 (tcgen05 3xTF32) against float64: every operand layout, ragged sizes, bias, split-K, N tiling."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, b, a_k, b_k):
    A = a.double() if a_k else a.double().t()
    B = b.double() if b_k else b.double().t()
    return A @ B.t(), A.abs() @ B.abs().t()


CASES = [
    # M, N, K, a_kmajor, b_kmajor, bias, splits
    (1000, 288, 288, True, True, False, 1),        # forward of a 32-channel 3x3-board layer
    (16384, 288, 288, True, True, True, 1),
    (777, 288, 288, True, False, False, 1),        # input gradient: B stored (K, N)
    (288, 288, 5000, False, False, False, 1),      # weight gradient: reduce over samples, both operands transposed
    (288, 288, 16384, False, False, False, 37),    # ... split over K slices
    (300, 288, 27, True, True, True, 1),           # stem: K = 27 (unaligned rows)
    (515, 27, 288, True, True, True, 1),           # heads: N = 27
    (130, 600, 96, True, True, False, 1),          # N tiled over several CTAs
    (64, 16, 8, True, True, False, 1),
    (129, 272, 40, False, True, False, 2),
]


@pytest.mark.parametrize('M,N,K,a_k,b_k,with_bias,splits', CASES)
def test_gemm_tf32x3_matches_float64(M, N, K, a_k, b_k, with_bias, splits):
    from handyrl_b200 import ops
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.randn((M, K) if a_k else (K, M), generator=g).cuda()
    b = torch.randn((N, K) if b_k else (K, N), generator=g).cuda()
    bias = torch.randn(N, generator=g).cuda() if with_bias else None
    got = ops.gemm_tf32x3(a, b, bias, a_kmajor=a_k, b_kmajor=b_k, splits=splits)
    torch.cuda.synchronize()
    want, scale = _ref(a, b, a_k, b_k)
    if with_bias:
        want = want + bias.double()
    err = ((got.double() - want).abs() / (scale + 1e-30)).max().item()
    # the tensor core adds every product into the fp32 accumulator with truncation: error ~ 0.5 sqrt(K_slice) ulp of
    # sum|a||b| (measured 10 ulp at K=288, 36 ulp at K=5000); single-pass TF32 would be ~1e-3
    k_slice = -(-K // splits)
    assert err < 1.2e-7 * (0.8 * k_slice ** 0.5 + 4), (err, k_slice)
    torch.backends.cuda.matmul.allow_tf32 = False
    A32 = a if a_k else a.t()
    B32 = b if b_k else b.t()
    fp32 = A32 @ B32.t() + (bias if with_bias else 0)
    err32 = ((fp32.double() - want).abs() / (scale + 1e-30)).max().item()
    assert err < 8 * err32 + 1.2e-7 * 0.8 * k_slice ** 0.5, (err, err32)      # same class as plain fp32 summation


def test_gemm_non_contiguous_leading_dimensions_and_output_view():
    from handyrl_b200 import ops
    g = torch.Generator().manual_seed(5)
    big_a = torch.randn((200, 320), generator=g).cuda()
    big_b = torch.randn((288, 300), generator=g).cuda()
    a, b = big_a[:, :288], big_b[:, :288]                     # lda = 320, ldb = 300
    out_big = torch.zeros((200, 400), device='cuda')
    out = out_big[:, 16:16 + 288]                             # ldc = 400
    ops.gemm_tf32x3(a, b, out=out)
    torch.cuda.synchronize()
    want = a.double() @ b.double().t()
    assert (out.double() - want).abs().max().item() < 2e-6 * (a.double().abs() @ b.double().abs().t()).max().item()
    assert out_big[:, :16].abs().max().item() == 0 and out_big[:, 16 + 288:].abs().max().item() == 0


@pytest.mark.parametrize('Cout,Cin,H,W,ksz,M', [(32, 32, 3, 3, 3, 1000), (32, 3, 3, 3, 3, 515), (2, 32, 3, 3, 1, 300), (8, 5, 4, 4, 3, 129)])
def test_packed_weight_images_forward_and_input_gradient(Cout, Cin, H, W, ksz, M):
    """hrl_board_pack writes a convolution over the board as the GEMM's packed B operand (pre-split, pre-swizzled, one bulk
    copy per stage): both images (forward / input gradient) must give the products of the float64 convolution."""
    import ctypes as C
    from handyrl_b200._capi import HrlGemmArgs, check, lib
    from handyrl_b200.ops import _ptr, _stream_ptr
    g = torch.Generator(device='cuda').manual_seed(Cout * 100 + Cin)
    w = torch.randn(Cout, Cin, ksz, ksz, device='cuda', generator=g)
    x = torch.randn(M, Cin, H, W, device='cuda', generator=g)
    dy = torch.randn(M, Cout, H, W, device='cuda', generator=g)
    row0 = 7 if Cout * H * W + 7 <= 288 else 0                  # forward operand shared with another (absent) convolution
    rows_f, rows_b = Cout * H * W + row0, Cin * H * W
    fwd = torch.zeros(lib().hrl_board_pack_floats(rows_f, rows_b), device='cuda')
    bwd = torch.zeros(lib().hrl_board_pack_floats(rows_b, rows_f), device='cuda')
    check(lib().hrl_board_pack(_ptr(w), Cout, Cin, ksz, ksz, H, W, _ptr(fwd), rows_f, row0, _ptr(bwd), rows_b, row0, _stream_ptr()))

    def product(a, image, N, K):
        out = torch.empty(M, N, device='cuda')
        args = HrlGemmArgs()
        args.a.ptr, args.a.ld, args.a.kmajor = _ptr(a), a.stride(0), 1
        args.b.ptr, args.b.kmajor, args.b.packed = _ptr(image), 1, 1
        args.C, args.ldc, args.M, args.N, args.K, args.splits = _ptr(out), N, M, N, K, 1
        check(lib().hrl_gemm_fused(C.byref(args), _stream_ptr()))
        return out

    xd, wd, dyd = x.double().requires_grad_(True), w.double(), dy.double()
    ref = torch.nn.functional.conv2d(xd, wd, padding=ksz // 2)
    ref.backward(dyd)
    y = product(x.reshape(M, -1), fwd, rows_f, rows_b)
    assert (y[:, :row0] == 0).all()
    scale = torch.nn.functional.conv2d(x.double().abs(), wd.abs(), padding=ksz // 2).max().item()
    assert (y[:, row0:].double() - ref.reshape(M, -1)).abs().max().item() <= 2e-6 * scale
    dyp = torch.cat([torch.randn(M, row0, device='cuda', generator=g), dy.reshape(M, -1)], 1).contiguous()      # garbage under the absent rows
    dx = product(dyp, bwd, rows_b, rows_f)
    assert (dx.double() - xd.grad.reshape(M, -1)).abs().max().item() <= 2e-6 * xd.grad.abs().max().item() * 9 * Cout


@pytest.mark.parametrize('M,N,K', [(300, 288, 288), (515, 144, 144), (515, 36, 144), (515, 144, 36), (2048, 27, 288)])
def test_a_operand_staged_through_shared_memory_equals_direct_reads(M, N, K):
    """With a packed B image and 16-byte aligned A rows the A tiles are copied to shared memory by cp.async, one chunk ahead,
    instead of being read row-per-lane from global memory: same arithmetic, so every variant (plain, affine + ReLU with the
    statistics epilogue, two sources with the masked epilogue) must agree bit for bit with the direct path (hook +64)."""
    import ctypes as C
    from handyrl_b200._capi import GEMM_EPILOGUES, HrlGemmArgs, check, lib
    from handyrl_b200.ops import _ptr, _stream_ptr
    g = torch.Generator(device='cuda').manual_seed(M + N)
    x, x2 = torch.randn(M, K, device='cuda', generator=g), torch.randn(M, K, device='cuda', generator=g)
    y = torch.randn(M, N, device='cuda', generator=g)
    c3 = [torch.rand(K, device='cuda', generator=g) for _ in range(3)]
    img = torch.randn(lib().hrl_board_pack_floats(N, K), device='cuda', generator=g)
    cp = torch.zeros(((M + 127) // 128) * 2 * N, device='cuda')

    def run(x2_, consts, relu, ep):
        out = torch.empty(M, N, device='cuda')
        a = HrlGemmArgs()
        a.a.ptr, a.a.ptr2, a.a.ld, a.a.kmajor, a.a.relu = _ptr(x), _ptr(x2_), K, 1, int(relu)
        if consts is not None:
            a.a.p, a.a.r = _ptr(consts[0]), _ptr(consts[-1])
            a.a.q = _ptr(consts[1]) if len(consts) == 3 else None
        a.b.ptr, a.b.kmajor, a.b.packed = _ptr(img), 1, 1
        a.C, a.ldc, a.M, a.N, a.K, a.splits, a.epilogue = _ptr(out), N, M, N, K, 1, GEMM_EPILOGUES[ep]
        if ep in ('stats', 'mask_stats'):
            a.col_partials = _ptr(cp)
        if ep == 'mask_stats':
            a.ep_y, a.ep_ldy = _ptr(y), N
        check(lib().hrl_gemm_fused(C.byref(a), _stream_ptr()))
        torch.cuda.synchronize()
        return out, cp.clone()

    vec = N % 4 == 0
    try:
        for x2_, consts, relu, ep in ((None, None, False, 'store'), (None, c3[:2], True, 'stats' if vec else 'relu'),
                                      (x2, c3, False, 'mask_stats' if vec else 'store')):
            lib().hrl_gemm_set_debug(64)
            direct = run(x2_, consts, relu, ep)
            lib().hrl_gemm_set_debug(0)
            staged = run(x2_, consts, relu, ep)
            assert torch.equal(direct[0], staged[0]) and torch.equal(direct[1], staged[1]), ep
    finally:
        lib().hrl_gemm_set_debug(0)
