#!/bin/bash
# usage: ncu_durations.sh <shape> [reps]   -- prints per-launch gpu__time_duration of the loss kernel (cold L2)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:loss_ --csv python scripts/run_loss_kernel.py $1 ${2:-5} 2>/dev/null | python -c "
import csv,sys
rows=[r for r in csv.reader(l for l in sys.stdin if l.startswith('\"'))]
h=rows[0]; i=h.index('Metric Value'); u=h.index('Metric Unit'); b=h.index('Block Size'); g=h.index('Grid Size')
vals=[float(r[i].replace(',','')) * (1e-3 if r[u]=='ns' else 1.0) for r in rows[1:]]
print('$1 FLUSH=%s STAGE=%s THREADS=%s' % ('${FLUSH:-read}', '${HRL_LOSS_STAGE:-}', '${HRL_LOSS_THREADS:-}'), 'block', rows[1][b], 'grid', rows[1][g], 'us:', ' '.join('%.2f'%v for v in vals))
"
