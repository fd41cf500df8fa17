"""Kernel-time table of one learner step (torch profiler / CUPTI; cheap alternative to an ncu launch list).
    python scripts/profile_step.py <workload> [rows]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from handyrl_b200.train import LearnerStep

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'cfg2']
args = bench.train_args(w)
stepper = LearnerStep(bench.make_net(w), args, bench.make_batch(w, 1), lr=1e-4, use_graph=False)
pk = stepper.new_packed().fill(bench.make_batch(w, 2))
for _ in range(3):
    stepper.step(pk)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        stepper.step(pk)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=int(sys.argv[2]) if len(sys.argv) > 2 else 25, max_name_column_width=90))
