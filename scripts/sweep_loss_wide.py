"""Wide-row loss kernel (cfg5shard shape) alone, over the bulk kernel's tuning knobs: CTAs per window x row-reducing warps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json, torch
import bench
from handyrl_b200 import ops
from handyrl_b200.synthetic import synthetic_batch, synthetic_outputs, bytes_per_cell

w = bench.WORKLOADS['cfg5shard']
args = bench.train_args(w)
B, T, P, A = w['B'], w['T'], w['P'], w['A']
Pa = 1
dev = torch.device('cuda')
peak = json.load(open(os.path.join(bench.ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(bench.ROOT, 'MEASURED_PEAKS.json')) else 6586.7
n = 6
sets = []
for i in range(n):
    b = synthetic_batch(B, T, P, A, turn_based=True, observation=False, seed=300 + i, with_obs=False)
    o = synthetic_outputs(b, seed=400 + i)
    sets.append(({k: v.to(dev) for k, v in o.items()}, {k: v.to(dev) for k, v in b.items()}))

def run(tuning, bf16=False):
    bufs = [ops.LossBuffers(B, T, P, Pa, A, True, False, dev, policy_dtype=torch.bfloat16 if bf16 else torch.float32) for _ in sets]
    ins = [(dict(o, policy=o['policy'].to(torch.bfloat16)) if bf16 else o, b) for o, b in sets]
    for (o, b), buf in zip(ins, bufs):
        ops.loss_fwd_bwd(o, b, args, buffers=buf, tuning=tuning)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for (o, b), buf in zip(ins, bufs):
                ops.loss_fwd_bwd(o, b, args, buffers=buf, tuning=tuning)
        g.replay(); side.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(6):
            g.replay()
        e1.record(side); side.synchronize()
    return e0.elapsed_time(e1) / (6 * n) * 1e3

per_set = bytes_per_cell(P, Pa, A, T, 0) * B * T
for bf16 in (False, True):
    by = per_set - (4 * Pa * A * B * T if bf16 else 0)
    for cl in (2,):
        for nc in (16,):
            try:
                us = run({'variant': 'bulk', 'cluster': cl, 'consumers': nc}, bf16)
                print('bf16=%d cluster %d consumers %2d: %6.1f us  %5.0f GB/s  frac %.3f' % (bf16, cl, nc, us, by / us / 1e3, by / us / 1e3 / peak))
            except Exception as e:
                print('bf16=%d cluster %d consumers %2d: %s' % (bf16, cl, nc, str(e)[:80]))
print('--- recurrence form')
for rec in ('serial', 'scan'):
    for nc in (16, 17):
        us = run({'variant': 'bulk', 'cluster': 2, 'consumers': nc, 'recurrence': rec})
        print('recurrence %-6s consumers %d: %6.1f us  frac %.3f' % (rec, nc, us, per_set / us / 1e3 / peak))
