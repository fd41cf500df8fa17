#!/usr/bin/env python
"""Benchmark of the HandyRL learner hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one learner step on one replay batch: Batcher output -> net forward -> fused loss
fwd+bwd kernel -> net backward -> [NCCL all-reduce SUM] -> clip + Adam.
metric = learner samples/s = B*T*steps/s over all GPUs (BASELINE.json).

  value     inputs already resident in HBM (a ring of distinct batches larger than L2)
  e2e       the same step through LearnerStep.step() with HOST (pinned) batches: one H2D copy per
            step and a D2H read of the step's loss sums inside the timed region
  roofline  the fused loss kernel: algorithmic bytes / CUDA-event duration measured live in the
            timed region, against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference: the eager-PyTorch CPU port of the reference learner step
            (oracle/torch_learner.py, pinned to the reference's golden vectors) on the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]: TicTacToe net, synthetic replay (T=32,B=512,P=2), V-Trace + UPGO
    'cfg2': dict(B=512, T=32, P=2, A=9, turn_based=True, observation=False, obs_shape=(3, 3, 3), net='tictactoe',
                 policy_target='UPGO', value_target='VTRACE', reward_kind='zero',
                 desc='configs[1]: TicTacToe net (29,006 params), synthetic replay T=32 B=512/GPU P=2 Pa=1 A=9, '
                      'policy UPGO + value V-Trace'),
    # per-GPU shard of BASELINE.json configs[4]: 64x64 obs, 512 actions, T=64, B=4096/8
    'cfg5shard': dict(B=512, T=64, P=2, A=512, turn_based=True, observation=False, obs_shape=(1, 64, 64), net='wide',
                      policy_target='UPGO', value_target='VTRACE', reward_kind='zero',
                      desc='configs[4] per-GPU shard: 64x64 obs / 512 actions, T=64 B=512/GPU P=2 Pa=1'),
    # BASELINE.json configs[2]: Geister-shaped recurrent net (dict observation, policy/value/return heads), TD(lambda),
    # batch 256, burn-in 4 + 16 forward steps
    'cfg3': dict(B=256, T=20, P=2, A=214, turn_based=True, observation=True, obs_shape=None, net='geister', burn_in=4,
                 policy_target='TD', value_target='TD', reward_kind='step',
                 desc='configs[2]: Geister-shaped recurrent net (conv-gated memory, 3 heads), TD(lambda), B=256/GPU, '
                      'T=4 burn-in + 16, P=Pa=2, A=214, dict observation {scalar 18, board 7x6x6}'),
}
L2_BYTES = 126e6


def train_args(w):
    return {'turn_based_training': w['turn_based'], 'observation': w['observation'], 'gamma': 0.8, 'lambda': 0.7,
            'burn_in_steps': w.get('burn_in', 0), 'forward_steps': w['T'] - w.get('burn_in', 0), 'entropy_regularization': 0.1,
            'entropy_regularization_decay': 0.1, 'policy_target': w['policy_target'], 'value_target': w['value_target'],
            'batch_size': w['B']}


def make_net(w):
    from handyrl_b200 import nets
    torch.manual_seed(0)
    if w['net'] == 'geister':
        return nets.GatedBoardNet(scalars=18, planes=7, board=(6, 6), width=32, actions=w['A'])
    return nets.tictactoe_net() if w['net'] == 'tictactoe' else nets.WideActionNet()


def make_batch(w, seed, B=None):
    from handyrl_b200.synthetic import synthetic_batch
    if w['net'] == 'geister':
        b = synthetic_batch(B or w['B'], w['T'], w['P'], w['A'], turn_based=w['turn_based'], observation=w['observation'],
                            reward_kind=w['reward_kind'], seed=seed, burn_in=w.get('burn_in', 0), with_obs=False)
        g = torch.Generator().manual_seed(seed + 7)
        Bn, T, Pa = b['action'].shape[:3]
        b['observation'] = {'scalar': torch.rand((Bn, T, Pa, 18), generator=g),
                            'board': (torch.rand((Bn, T, Pa, 7, 6, 6), generator=g) < 0.3).float()}
        return b
    return synthetic_batch(B or w['B'], w['T'], w['P'], w['A'], turn_based=w['turn_based'], observation=w['observation'],
                           reward_kind=w['reward_kind'], seed=seed, obs_shape=w['obs_shape'])


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU while the timed region runs (NVML, the
    same source nvidia-smi reads)."""

    def __init__(self, index, period=0.05):
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.stop_flag = threading.Event()
        self.sm_max = None
        self.th = None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4): 'sw_power_cap',
                getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8): 'hw_slowdown',
                getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
                getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
                getattr(nv, 'nvmlClocksThrottleReasonHwPowerBrakeSlowdown', 0x80): 'hw_power_brake',
            }
            get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
                getattr(nv, 'nvmlDeviceGetCurrentClocksThrottleReasons')
            while not self.stop_flag.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = get_reasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(self.period)
        except Exception as e:  # noqa: BLE001
            self.reasons.add('sampler_error:%s' % type(e).__name__)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *exc):
        self.stop_flag.set()
        self.th.join(timeout=2)

    def summary(self):
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.sm_max, 'reasons': sorted(self.reasons),
                'samples': len(s)}


def physical_device_index(local):
    vis = os.environ.get('CUDA_VISIBLE_DEVICES')
    if vis:
        try:
            return int(vis.split(',')[local])
        except Exception:
            return local
    return local


# ------------------------------------------------------------------------------ CPU arm

def run_cpu_port(w, steps, warmup, budget_s, threads=None):
    """The eager-PyTorch CPU port of the reference learner step on a bounded sample."""
    from oracle.torch_learner import CpuLearner
    if threads:
        torch.set_num_threads(threads)
    args = train_args(w)
    lrn = CpuLearner(make_net(w), args, lr=3e-8 * w['B'] * w['T'])
    B = w['B']
    probe = make_batch(w, 1000, B=min(B, 64))
    t0 = time.perf_counter()
    lrn.step(probe)
    lrn.step(probe)
    t_probe = (time.perf_counter() - t0) / 2 * (B / probe['action'].shape[0])    # estimated full-batch step
    if t_probe * (steps + warmup) > budget_s:
        Bs = max(16, int(B * budget_s / (t_probe * (steps + warmup))))
        Bs = 1 << (Bs.bit_length() - 1)
    else:
        Bs = B
    batches = [make_batch(w, 2000 + i, B=Bs) for i in range(4)]
    for i in range(warmup):
        lrn.step(batches[i % 4])
    t0 = time.perf_counter()
    for i in range(steps):
        lrn.step(batches[i % 4])
    dt = time.perf_counter() - t0
    return {'value': Bs * w['T'] * steps / dt, 'ms_per_step': dt / steps * 1e3, 'B_sample': Bs,
            'cores': torch.get_num_threads(), 'steps': steps}


def reference_arm(opt, w):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return      # the host has one set of cores: rank 0 alone measures it
    if w['net'] == 'geister':
        print(json.dumps({'impl': 'reference', 'unavailable': 'the CPU port (oracle/torch_learner.py) covers feed-forward nets only'}), flush=True)
        return
    # torchrun exports OMP_NUM_THREADS=1; the baseline gets the host's physical cores regardless (capped at 64)
    r = run_cpu_port(w, opt.steps, opt.warmup, budget_s=150.0, threads=max(1, min(64, (os.cpu_count() or 2) // 2)))
    sample = '%d steps of a B=%d x T=%d batch (workload B=%d)' % (opt.steps, r['B_sample'], w['T'], w['B'])
    line = {
        'impl': 'reference', 'metric': 'learner_samples_per_sec', 'value': r['value'], 'unit': 'samples/s',
        'n_gpus': opt.gpus, 'steps': opt.steps, 'warmup': opt.warmup, 'ms_per_step': r['ms_per_step'],
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
        'config': {'workload': w['desc'], 'global_batch': r['B_sample'], 'seq_len': w['T'], 'parallelism': 'cpu'},
        'cpu_baseline': {'value': r['value'], 'unit': 'samples/s', 'cores': r['cores'], 'kind': 'port', 'sample': sample},
        'e2e': {'value': r['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ B200 arm

def loss_kernel_name(A):
    """Which variant hrl_loss_fwd_bwd dispatches to for this action count (csrc/loss_kernel.cu)."""
    return 'hrl::loss_group_kernel' if A <= 32 else ('hrl::loss_bulk_kernel' if (A > 256 and A % 4 == 0) else 'hrl::loss_rows_kernel')


def time_loss_alone(B, T, P, A, turn_based, observation, args, device, reps):
    """Average device time (ms) of hrl_loss_fwd_bwd launched back to back over input sets that together exceed L2."""
    from handyrl_b200 import ops
    from handyrl_b200.synthetic import synthetic_batch, synthetic_outputs, bytes_per_cell
    Pa = 1 if (turn_based and not observation) else P
    per_set = bytes_per_cell(P, Pa, A, T, 0) * B * T
    n = max(2, min(64, int(2 * L2_BYTES / per_set) + 1))
    sets = []
    for i in range(n):
        b = synthetic_batch(B, T, P, A, turn_based=turn_based, observation=observation, seed=300 + i, with_obs=False)
        o = synthetic_outputs(b, seed=400 + i)
        sets.append(({k: v.to(device) for k, v in o.items()}, {k: v.to(device) for k, v in b.items()},
                     ops.LossBuffers(B, T, P, Pa, A, True, False, device)))
    for o, b, buf in sets:
        ops.loss_fwd_bwd(o, b, args, buffers=buf)
    torch.cuda.synchronize()
    # one launch per input set captured in a CUDA graph: no host launch overhead between kernels
    side = torch.cuda.Stream(device=device)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for o, b, buf in sets:
                ops.loss_fwd_bwd(o, b, args, buffers=buf)
        rounds = max(1, reps // n)
        graph.replay()
        side.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(rounds):
            graph.replay()
        e1.record(side)
        side.synchronize()
    return {'ms': e0.elapsed_time(e1) / (rounds * n), 'bytes': per_set, 'sets': n}


def b200_arm(opt, w):
    import torch.distributed as dist
    from handyrl_b200 import ops
    from handyrl_b200.synthetic import bytes_per_cell
    from handyrl_b200.train import LearnerStep, PackedBatch

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    pg = None
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
        pg = dist.group.WORLD

    args = train_args(w)
    B, T, P, A = w['B'], w['T'], w['P'], w['A']
    example = make_batch(w, 10_000 + rank)
    stepper = LearnerStep(make_net(w), args, example, lr=3e-8 * B * T * world, device=device, process_group=pg,
                          use_graph=True, time_loss_kernel=True)
    nbytes = stepper.layout.nbytes
    R = max(8, int(2 * L2_BYTES / nbytes) + 1)
    R = min(R, 96)
    host_ring = [PackedBatch(stepper.layout).fill(make_batch(w, 20_000 + rank * 1000 + i)) for i in range(R)]
    dev_ring = torch.empty((R, nbytes), dtype=torch.uint8, device=device)
    for i, pk in enumerate(host_ring):
        dev_ring[i].copy_(pk.buffer)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n_warm, n_steps):
        for i in range(n_warm):
            fn(i)
        stepper.loss_kernel_ms()            # drop warm-up kernel timings
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        with torch.cuda.stream(stepper.stream):
            e0.record()
        for i in range(n_steps):
            fn(n_warm + i)
        with torch.cuda.stream(stepper.stream):
            e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, wall

    # ---- value: inputs resident in HBM
    with ClockSampler(physical_device_index(local)) as clocks:
        ms, wall = timed(lambda i: stepper.step_resident(dev_ring[i % R]), max(3, opt.warmup), opt.steps)
    kernel_ms, n_k = stepper.loss_kernel_ms()
    value = B * T * world * opt.steps / (ms * 1e-3)

    # ---- e2e: host batches, H2D inside, loss read back every step (lagged by one step)
    pending = []

    def e2e_step(i):
        stepper.step(host_ring[i % R])
        pending.append(stepper.fetch_losses_async())
        if len(pending) > 1:
            pending.pop(0)()

    ms_e2e, wall_e2e = timed(e2e_step, max(3, opt.warmup), opt.steps)
    last = pending.pop()()
    stepper.loss_kernel_ms()
    e2e_value = B * T * world * opt.steps / (ms_e2e * 1e-3)

    # ---- the loss kernel alone on cold inputs (distinct input sets larger than L2), at the bench shape and at the
    #      wide-row shape of configs[4]'s per-GPU shard (where an HBM roofline is physically meaningful)
    alone, wide = None, None
    if rank == 0:
        alone = time_loss_alone(B, T, P, A, w['turn_based'], w['observation'], args, device, reps=200)
        if not opt.no_wide:
            ww = WORKLOADS['cfg5shard']
            wide = time_loss_alone(ww['B'], ww['T'], ww['P'], ww['A'], ww['turn_based'], ww['observation'], train_args(ww),
                                   device, reps=40)

    if rank != 0:
        shutdown(stepper, world)
        return

    from handyrl_b200 import fastnet
    n_fused_bn = sum(isinstance(m, fastnet.BoardBatchNorm2d) for m in stepper.model.modules()) if w['net'] == 'tictactoe' else 0
    peak, peak_src = measured_peak()
    Pa = example['action_mask'].shape[2]
    alg_bytes = bytes_per_cell(P, Pa, A, T, 1 if w['net'] == 'geister' else 0) * B * T
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_all = None, None
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            traffic_all = json.load(f)
            traffic = traffic_all.get(opt.workload)
    except Exception:
        pass
    line = {
        'metric': 'learner_samples_per_sec', 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': opt.steps,
        'warmup': max(3, opt.warmup), 'ms_per_step': ms / opt.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
        'config': {'workload': w['desc'], 'global_batch': B * world, 'seq_len': T, 'parallelism': 'dp%d' % world,
                   'tf32': False, 'cuda_graph': True,
                   'l2': 'inputs rotate over a ring of %d distinct resident batches (%.0f MB > 126 MB L2)' % (R, R * nbytes / 1e6),
                   'batchnorm': 'per-shard statistics (as the reference DataParallel)'},
        'clocks': clocks.summary(),
        'e2e': {'value': e2e_value, 'unit': 'samples/s', 'ms_per_step': ms_e2e / opt.steps,
                'h2d_bytes_per_step': nbytes * world, 'd2h_bytes_per_step': 24 * world,
                'wall_s': wall_e2e, 'last_losses': last},
        # our kernels per step: loss fwd+bwd, [peer all-reduce+]grad sum-of-squares, clip+Adam, step counter,
        # and 3 forward + 3 backward launches per fused BatchNorm layer of the (rewritten) net
        'gpu_launches': (4 + 6 * n_fused_bn) * opt.steps,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': traffic, 'kernel': loss_kernel_name(A) + ' (hrl_loss_fwd_bwd)', 'kernel_us': kernel_ms * 1e3,
                     'launches_timed': n_k, 'algorithmic_bytes': alg_bytes, 'peak_source': peak_src,
                     'alone_cold_us': None if alone is None else alone['ms'] * 1e3,
                     'alone_cold_gbs': None if alone is None else alg_bytes / (alone['ms'] * 1e-3) / 1e9,
                     'note': 'event-bracketed single launch inside the step; %.2f MB per launch = %.2f us at peak%s, see '
                             'DESIGN.md section 4' % (alg_bytes / 1e6, alg_bytes / peak / 1e3,
                                                      ' (latency-bound: below one DRAM round trip + launch)' if alg_bytes < 2e7 else '')},
        'wall_s': wall,
    }
    if wide is not None:
        gbs = wide['bytes'] / (wide['ms'] * 1e-3) / 1e9
        line['roofline_wide_rows'] = {
            'workload': WORKLOADS['cfg5shard']['desc'] + ' (loss kernel alone, %d cold input sets)' % wide['sets'],
            'bound': 'hbm', 'achieved': gbs, 'peak': peak, 'unit': 'GB/s', 'frac': gbs / peak,
            'kernel': loss_kernel_name(WORKLOADS['cfg5shard']['A']) + ' (hrl_loss_fwd_bwd)', 'kernel_us': wide['ms'] * 1e3,
            'algorithmic_bytes': wide['bytes'], 'traffic': None if traffic_all is None else traffic_all.get('cfg5shard')}
    if world == 1 and not opt.no_cpu and w['net'] != 'geister':     # the CPU port is feed-forward only
        r = run_cpu_port(w, steps=8, warmup=1, budget_s=25.0, threads=max(1, min(64, (os.cpu_count() or 2) // 2)))
        r1 = run_cpu_port(w, steps=4, warmup=1, budget_s=12.0, threads=1)
        line['cpu_baseline'] = {
            'value': r['value'], 'unit': 'samples/s', 'cores': r['cores'], 'kind': 'port',
            'sample': '%d steps of B=%d x T=%d (eager PyTorch CPU port of the reference step, oracle/torch_learner.py)'
                      % (r['steps'], r['B_sample'], T),
            'as_shipped_1_thread': r1['value'], 'host_cores': os.cpu_count()}
    print(json.dumps(line), flush=True)
    shutdown(stepper, world)


def shutdown(stepper, world):
    """Tear down NCCL after the captured graphs are gone; never let a teardown hang eat the run."""
    if world <= 1:
        return
    import torch.distributed as dist
    sys.stdout.flush()
    killer = threading.Timer(20.0, lambda: os._exit(0))
    killer.daemon = True
    killer.start()
    stepper.close()
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()
    killer.cancel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='cfg2', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-wide', action='store_true', help='skip the wide-row loss-kernel measurement')
    opt = ap.parse_args()
    w = WORKLOADS[opt.workload]
    if opt.impl == 'reference':
        reference_arm(opt, w)
    else:
        if not torch.cuda.is_available():
            raise SystemExit('bench.py: no CUDA device; the learner hot path has no CPU fallback '
                             '(use --impl reference for the CPU port)')
        b200_arm(opt, w)


if __name__ == '__main__':
    main()
