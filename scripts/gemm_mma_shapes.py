"""EXPERIMENT: tcgen05.mma time per 32-element chunk at N = 288 cut into instructions in different ways (no operand loads)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from handyrl_b200._capi import lib
from bench_gemm_fused import timeit, eng, X, out, D
dbg = lib().hrl_gemm_set_debug
Wf = eng.Wf[0]
f = lambda K: (lambda: eng._gemm(dict(t=X), dict(t=Wf, packed=True), out, K=K, N=D))
for mode, name in ((0, '144+144'), (1, '256+32'), (2, '96+96+96'), (3, '192+96'), (4, '128+128+32')):
    for sub in (2, 0):
        dbg(mode * 4 + sub)
        t = [timeit(f(K)) for K in (32, 288)]
        print('%-12s %s  K=32 %5.1f  K=288 %5.1f  -> %.2f us/chunk' % (name, 'MMA only' if sub else 'full    ', t[0], t[1], (t[1] - t[0]) / 8))
dbg(0)
