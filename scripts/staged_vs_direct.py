"""Packed-B products with the A operand staged through shared memory (cp.async) vs read directly: same arithmetic, so the
results must be bit-identical.  Prints where they are not."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from handyrl_b200._capi import HrlGemmArgs, GEMM_EPILOGUES, check, lib
from handyrl_b200.ops import _ptr, _stream_ptr
dbg = lib().hrl_gemm_set_debug
def run(x, x2, consts, relu, img, N, K, ep, y, cp):
    M = x.shape[0]
    out = torch.empty(M, N, device='cuda')
    a = HrlGemmArgs()
    a.a.ptr, a.a.ptr2, a.a.ld, a.a.kmajor, a.a.relu = _ptr(x), _ptr(x2), x.stride(0), 1, int(relu)
    if consts is not None:
        a.a.p, a.a.r = _ptr(consts[0]), _ptr(consts[-1])
        a.a.q = _ptr(consts[1]) if len(consts) == 3 else None
    a.b.ptr, a.b.kmajor, a.b.packed = _ptr(img), 1, 1
    a.C, a.ldc, a.M, a.N, a.K, a.splits = _ptr(out), N, M, N, K, 1
    a.epilogue = GEMM_EPILOGUES[ep]
    if ep != 'store' and ep != 'relu':
        a.col_partials = _ptr(cp)
    if ep == 'mask_stats':
        a.ep_y, a.ep_ldy = _ptr(y), y.stride(0)
    check(lib().hrl_gemm_fused(C.byref(a), _stream_ptr()))
    return out
bad = 0
for seed in range(12):
    for (M, N, K) in ((300, 288, 288), (515, 144, 144), (2048, 288, 288), (515, 36, 144), (515, 144, 36), (260, 128, 128), (2048, 27, 288)):
        g = torch.Generator(device='cuda').manual_seed(seed * 100 + M)
        x = torch.randn(M, K, device='cuda', generator=g); x2 = torch.randn(M, K, device='cuda', generator=g)
        y = torch.randn(M, N, device='cuda', generator=g)
        c3 = [torch.rand(K, device='cuda', generator=g) for _ in range(3)]
        img = torch.randn(lib().hrl_board_pack_floats(N, K), device='cuda', generator=g)      # any image will do
        cp = torch.empty(((M + 127) // 128) * 2 * N, device='cuda')
        for name, kw in (('plain', dict(x2=None, consts=None, relu=False, ep='store')),
                         ('affine+relu+stats', dict(x2=None, consts=c3[:2], relu=True, ep='stats' if N % 4 == 0 else 'store')),
                         ('2src+mask', dict(x2=x2, consts=c3, relu=False, ep='mask_stats' if N % 4 == 0 else 'store'))):
            res = []
            for mode in (64, 0):
                dbg(mode)
                res.append(run(x, kw['x2'], kw['consts'], kw['relu'], img, N, K, kw['ep'], y, cp))
                torch.cuda.synchronize()
            dbg(0)
            if not torch.equal(res[0], res[1]):
                bad += 1
                d = (res[0] != res[1])
                rows = d.any(1).nonzero().flatten().tolist(); cols = d.any(0).nonzero().flatten().tolist()
                print('MISMATCH seed %d M=%d N=%d K=%d %s: %d elements, rows %s.. (%d) cols %s.. (%d) max diff %g' % (
                    seed, M, N, K, name, int(d.sum()), rows[:6], len(rows), cols[:6], len(cols), (res[0] - res[1]).abs().max().item()))
print('mismatching products:', bad)
