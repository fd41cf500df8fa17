"""Which operand transform costs what in the weight-gradient product (288 x 288 over 16384 samples in 48 slices)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from handyrl_b200._capi import lib
from bench_gemm_fused import timeit, eng, X, Y, c, D, M
sp = eng.splits['tower']
A0 = dict(t=X, kmajor=False)
A1 = dict(t=X, consts=(c[0], c[1]), kmajor=False, by_row=True)
A2 = dict(t=X, t2=Y, consts=(c[0], c[1], c[2]), kmajor=False, by_row=True)
B0 = dict(t=Y, kmajor=False)
B1 = dict(t=Y, consts=(c[3], c[4]), relu=True, kmajor=False, by_row=True)
dbg = lib().hrl_gemm_set_debug
for mode in (0, 1, 2):
    dbg(mode)
    for an, a in (('A plain', A0), ('A affine', A1), ('A 2-source', A2)):
        for bn, b in (('B plain', B0), ('B affine+relu', B1)):
            print('debug %d  %-12s %-14s %6.1f us' % (mode, an, bn, timeit(lambda: eng._gemm(a, b, None, K=M, N=D, M=D, splits=sp, partial=True, ws=eng.ws))))
dbg(0)
