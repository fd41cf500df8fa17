"""Where the time of a packed-B hrl_gemm_fused launch goes at the tower shape: K sweep (prologue + epilogue = the K -> 0
intercept, per-chunk cost = the slope) and the profiling knob (1 = no MMAs, 2 = no operand loads / stores)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from handyrl_b200 import nets, tower
from handyrl_b200._capi import lib
from bench_gemm_fused import timeit, eng, X, Y, out, c, ep_full, D, M   # noqa (runs that script's table first)

print('--- breakdown')
dbg = lib().hrl_gemm_set_debug
Wf, Wb = eng.Wf[0], eng.Wb[0]
fwd = lambda K: (lambda: eng._gemm(dict(t=X), dict(t=Wf, packed=True), out, K=K, N=D))
fwdT = lambda K: (lambda: eng._gemm(dict(t=X, consts=(c[0], c[1]), relu=True), dict(t=Wf, packed=True), out, K=K, N=D, epilogue='stats'))
dgr = lambda K: (lambda: eng._gemm(dict(t=X, t2=Y, consts=(c[0], c[1], c[2])), dict(t=Wb, packed=True), out, K=K, N=D,
                                   epilogue='mask_stats', ep=ep_full))
for name, f in (('fwd packed', fwd), ('fwd packed + transform + stats', fwdT), ('dgrad packed 2-source + mask_stats', dgr)):
    for mode in (0, 1, 2):
        dbg(mode)
        print('%-36s debug %d  ' % (name, mode) + '  '.join('K=%d: %5.1f' % (K, timeit(f(K))) for K in (32, 96, 160, 288)))
dbg(0)
print('--- other kernels')
from bench_gemm_fused import cases
for k in ('fwd plain', 'fwd stats epilogue', 'dgrad plain + mask_stats', 'wgrad plain (48 slices, partials)', 'wgrad both transformed'):
    print('%-40s %6.1f us' % (k, timeit(cases[k])))
