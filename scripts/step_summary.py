"""Cut one steady-state learner step out of an ncu launch list (--metrics gpu__time_duration.sum --csv) and print
per-kernel totals and shares: the launches between the last two clip_adam_kernel launches.

    python scripts/step_summary.py gpurun_out/launches.csv [title]
"""
import csv
import sys
from collections import OrderedDict

path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else path
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
h = rows[0]
ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
launches = [(r[ki], float(r[vi].replace(',', '')) * (1e-3 if r[ui] == 'ns' else 1.0)) for r in rows[1:] if r[h.index('Metric Name')] == 'gpu__time_duration.sum']
ends = [i for i, (k, _) in enumerate(launches) if 'clip_adam_kernel' in k]
assert len(ends) >= 2, 'need two optimiser steps in the list'
step = launches[ends[-2] + 1:ends[-1] + 1]
# the step counter bump belongs to the optimiser of the same step
if ends[-1] + 1 < len(launches) and 'bump_step' in launches[ends[-1] + 1][0]:
    step = launches[ends[-2] + 2:ends[-1] + 2]
agg = OrderedDict()
for k, us in step:
    a = agg.setdefault(k, [0.0, 0])
    a[0] += us
    a[1] += 1
total = sum(us for _, us in step)
print('# %s' % title)
print('# one steady-state learner step: the %d launches between two consecutive optimiser steps.  ncu serialises launches and' % len(step))
print('# flushes caches: compare SHARES, not absolutes.')
print('  total_us  count     avg_us  share  kernel')
for k, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print('%10.1f %6d %10.2f %5.1f%%  %s' % (us, n, us / n, 100 * us / total, k[:150]))
ours = sum(us for k, us in step if k.startswith('hrl::') or 'hrl::' in k)
print('\nstep total (serialised) %.1f us over %d launches; hrl:: kernels %.1f%% of it' % (total, len(step), 100 * ours / total))
