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
