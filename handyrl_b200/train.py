"""B200-native learner hot path behind the reference's Python surface.

Same public names as handyrl/train.py for the path
    Batcher.batch -> forward_prediction -> compute_loss -> backward -> optimizer.step
(reference train.py:127-400), different machinery:

  * the net runs ONCE per step and returns raw outputs; the mask epilogue, the whole of
    compute_loss / compose_losses / losses.py AND their backward run in one CUDA kernel
    (ops.loss_fwd_bwd -> csrc/loss_kernel.cu) that hands autograd closed-form gradients;
  * parameters and gradients live in one flat bucket: one NCCL all-reduce(SUM) per step when
    sharded over GPUs, then clip+Adam in two launches (csrc/optim_kernel.cu);
  * the whole step (H2D copies, net forward, loss kernel, net backward, all-reduce, optimiser)
    is captured in a CUDA graph and replayed; the host never synchronises inside an epoch
    (the reference does 4-6 .item() syncs per step, train.py:200, 375-376).
"""
import collections
import copy
import os
import pickle
import queue
import threading
import time
import traceback
from collections import deque

import numpy as np
import torch

from . import ops, fastnet
from .batch import tree_map, tree_leaves, make_batch, gather_windows, sample_window
from ._capi import LOSS_KEYS, NUM_LOSS


# --------------------------------------------------------------------------- forward

def _call_model(model, obs, hidden):
    return model(obs, hidden)


def _as_format(o, memory_format):
    if memory_format is not None and o.dim() == 4:
        return o.contiguous(memory_format=memory_format)
    return o


def forward_raw(model, hidden, batch, args, memory_format=None):
    """Run the net over a batch and return its RAW outputs shaped (B, T, Pa, ...).

    Feed-forward nets see all B*T*Pa observations at once; recurrent nets are stepped over T
    with the hidden state masked by observation_mask, burn-in steps without gradient and in
    eval mode, exactly as the reference does (train.py:142-174) -- but WITHOUT the mask
    epilogue of train.py:176-184, which the fused loss kernel applies on the fly.
    """
    observations = batch['observation']
    B, T, Pa = batch['action'].shape[:3]

    if hidden is None:
        flat = tree_map(lambda o: _as_format(o.flatten(0, 2), memory_format), observations)
        outs = _call_model(model, flat, None)
        return {k: v.unflatten(0, (B, T, Pa)) for k, v in outs.items() if k != 'hidden' and v is not None}

    alternating = args['turn_based_training'] and not args['observation']
    burn_in = args['burn_in_steps']
    per_step = {}
    omask_all = batch['observation_mask']
    for t in range(T):
        obs_t = tree_map(lambda o: _as_format(o[:, t].flatten(0, 1), memory_format), observations)
        om = omask_all[:, t]                                                   # (B, P, 1)

        def gate(h):
            return om.view(*h.shape[:2], *([1] * (h.dim() - 2)))

        fused = om.is_cuda and om.dtype == torch.float32          # one kernel per hidden leaf instead of mul + sum / blend chains
        if fused:
            visible = tree_map(lambda h: ops.hidden_visible(h, om, alternating), hidden)
            if not alternating:
                visible = tree_map(lambda h: h.flatten(0, 1), visible)
        else:
            visible = tree_map(lambda h: h * gate(h), hidden)
            if alternating:
                visible = tree_map(lambda h: h.sum(1), visible)                # only the turn player observes
            else:
                visible = tree_map(lambda h: h.flatten(0, 1), visible)
        if t < burn_in:
            model.eval()
            with torch.no_grad():
                out_t = _call_model(model, obs_t, visible)
        else:
            if not model.training:
                model.train()
            out_t = _call_model(model, obs_t, visible)
        new_hidden = out_t.pop('hidden', None)
        for k, v in out_t.items():
            if v is not None:
                per_step.setdefault(k, []).append(v.unflatten(0, (B, Pa)))
        new_hidden = tree_map(lambda h: h.unflatten(0, (B, Pa)), new_hidden)
        if fused:
            hidden = tree_map(lambda h, nh: ops.hidden_blend(h, nh, om), hidden, new_hidden)
        else:
            hidden = tree_map(lambda h, nh: h * (1 - gate(h)) + nh * gate(h), hidden, new_hidden)
    return {k: torch.stack(v, dim=1) for k, v in per_step.items()}


def forward_prediction(model, hidden, batch, args):
    """API-compatible with handyrl.train.forward_prediction: raw outputs + the mask epilogue
    (train.py:176-184) as torch ops.  NOT used by the learner below (the epilogue is fused into
    the loss kernel); kept so code written against the reference keeps importing."""
    outs = forward_raw(model, hidden, batch, args)
    Pa = batch['action'].shape[2]
    masked = {}
    for k, o in outs.items():
        if k == 'policy':
            o = o * batch['turn_mask']
            if o.size(2) > 1 and Pa == 1:
                o = o.sum(2, keepdim=True)
            masked[k] = o - batch['action_mask']
        else:
            masked[k] = o * batch['observation_mask']
    return masked


# --------------------------------------------------------------------------- loss

class _FusedLoss(torch.autograd.Function):
    """total = fused_loss(policy_raw, value_raw?, return_raw?); backward returns the closed-form
    gradients the kernel already produced, scaled by grad_output."""

    @staticmethod
    def forward(ctx, batch, args, keys, *heads):
        outs = dict(zip(keys, [h.detach() for h in heads]))
        buf = ops.loss_fwd_bwd(outs, batch, args)
        grads = {'policy': buf.dpolicy, 'value': buf.dvalue, 'return': buf.dreturn}
        ctx.save_for_backward(*[grads[k] for k in keys])
        ctx.mark_non_differentiable(buf.losses)
        total = buf.losses[4].clone()
        return total, buf.losses

    @staticmethod
    def backward(ctx, g_total, _g_losses):
        return (None, None, None) + tuple(g * g_total for g in ctx.saved_tensors)


def compute_loss(batch, model, hidden, args):
    """Drop-in for handyrl.train.compute_loss (train.py:218-267): returns
    ({'p','v','r','ent','total'} 0-d tensors, dcnt float).  `total` is differentiable."""
    outs = forward_raw(model, hidden, batch, args)
    keys = [k for k in ('policy', 'value', 'return') if k in outs]
    total, vec = _FusedLoss.apply(batch, args, keys, *[outs[k] for k in keys])
    losses = {'p': vec[0]}
    if 'value' in outs:
        losses['v'] = vec[1]
    if 'return' in outs:
        losses['r'] = vec[2]
    losses['ent'] = vec[3]
    losses['total'] = total
    return losses, float(vec[5].item())     # the reference also synchronises here (train.py:200)


# --------------------------------------------------------------------------- one learner step

def obs_rows(obs):
    return tree_leaves(obs)[0].shape[0]


def _align(x, a=256):
    return (x + a - 1) // a * a


class BatchLayout:
    """Byte layout of one replay batch packed into a single buffer (one H2D copy per step).
    `value` (the behaviour value, train.py:117) is left out: nothing on the hot path reads it."""

    SKIP = ('value',)

    def __init__(self, example_batch):
        self.entries = []      # (key path tuple, shape, dtype, offset)
        off = 0

        def walk(tree, path):
            nonlocal off
            if isinstance(tree, dict):
                for k, v in tree.items():
                    walk(v, path + (k,))
            elif isinstance(tree, (list, tuple)):
                for i, v in enumerate(tree):
                    walk(v, path + (i,))
            else:
                self.entries.append((path, tuple(tree.shape), tree.dtype, off))
                off = _align(off + tree.numel() * tree.element_size())

        for k, v in example_batch.items():
            if k not in self.SKIP:
                walk(v, (k,))
        self.template = {k: v for k, v in example_batch.items() if k not in self.SKIP}
        self.nbytes = off
        self.payload_bytes = sum(int(torch.tensor(sh).prod()) * torch.empty(0, dtype=dt).element_size()
                                 for _, sh, dt, _ in self.entries)

    def views(self, buf):
        """Tree of tensors (same nesting as the batch) aliasing the flat uint8 buffer `buf`."""
        flat = {}
        for path, shape, dtype, off in self.entries:
            n = int(torch.tensor(shape).prod()) * torch.empty(0, dtype=dtype).element_size()
            flat[path] = buf[off:off + n].view(dtype).view(shape)

        def build(tree, path):
            if isinstance(tree, dict):
                return type(tree)((k, build(v, path + (k,))) for k, v in tree.items())
            if isinstance(tree, (list, tuple)):
                return type(tree)(build(v, path + (i,)) for i, v in enumerate(tree))
            return flat[path]

        return build(self.template, ())


class PackedBatch:
    """A replay batch in ONE pinned host buffer; `tensors` alias it in the reference's dict layout."""

    def __init__(self, layout):
        self.layout = layout
        self.buffer = torch.empty(layout.nbytes, dtype=torch.uint8).pin_memory()
        self.tensors = layout.views(self.buffer)
        self.in_flight = None      # CUDA event: the H2D copy reading this buffer has completed

    def fill(self, host_batch):
        for d, s in zip(tree_leaves(self.tensors), tree_leaves({k: host_batch[k] for k in self.tensors})):
            d.copy_(s)
        return self

    def wait_reusable(self):
        if self.in_flight is not None:
            self.in_flight.synchronize()
            self.in_flight = None


class StateStore:
    """Everything `model.state_dict()` holds, in ONE device allocation: [fp32 parameters, padded to 4 | fp32 buffers
    (BatchNorm running statistics) | int64 buffers (num_batches_tracked)].  Parameters and buffers are re-pointed
    into it, so the per-epoch model hand-off to the workers (train.py:385-387) is one device-to-device snapshot on
    the step stream plus one device-to-host copy on a side stream instead of a `model.cpu()` of every tensor."""

    def __init__(self, model, device):
        params = list(model.parameters())
        self.n = sum(p.numel() for p in params)
        self.n_pad = (self.n + 3) // 4 * 4
        named = list(model.named_buffers())
        fbufs = [(k, b) for k, b in named if b.dtype == torch.float32]
        ibufs = [(k, b) for k, b in named if b.dtype == torch.int64]
        self.loose = [(k, b) for k, b in named if b.dtype not in (torch.float32, torch.int64)]   # copied one by one
        nf = sum(b.numel() for _, b in fbufs)
        ni = sum(b.numel() for _, b in ibufs)
        self.f_off = 4 * self.n_pad
        self.i_off = (self.f_off + 4 * nf + 7) // 8 * 8
        self.nbytes = self.i_off + 8 * ni
        self.bytes = torch.zeros(max(self.nbytes, 16), dtype=torch.uint8, device=device)
        self.flat_param = self.bytes[:4 * self.n_pad].view(torch.float32)
        fview = self.bytes[self.f_off:self.f_off + 4 * nf].view(torch.float32)
        iview = self.bytes[self.i_off:self.i_off + 8 * ni].view(torch.int64)
        self.where = {}           # state_dict key -> (view name, offset in elements, shape)
        with torch.no_grad():
            for view, tag, bufs in ((fview, 'f', fbufs), (iview, 'i', ibufs)):
                off = 0
                for k, b in bufs:
                    n = b.numel()
                    view[off:off + n].copy_(b.reshape(-1))
                    b.data = view[off:off + n].view(b.shape)
                    self.where[k] = (tag, off, tuple(b.shape))
                    off += n
        self._params = params

    def index_params(self, model):
        """Call after FlatAdam re-pointed the parameters into flat_param (same order as model.parameters())."""
        off = 0
        for k, p in model.named_parameters():
            self.where[k] = ('p', off, tuple(p.shape))
            off += p.numel()

    def state_dict_from(self, host_bytes, keys):
        """Rebuild {key: CPU tensor} from a host copy of `bytes` (fresh storage per tensor)."""
        views = {'p': host_bytes[:4 * self.n_pad].view(torch.float32),
                 'f': host_bytes[self.f_off:self.i_off].view(torch.float32) if self.i_off > self.f_off else None,
                 'i': host_bytes[self.i_off:self.nbytes].view(torch.int64) if self.nbytes > self.i_off else None}
        out = {}
        for k in keys:
            if k in self.where:
                tag, off, shape = self.where[k]
                n = 1
                for d in shape:
                    n *= d
                out[k] = views[tag][off:off + n].clone().view(shape)
        return out


class PendingModel:
    """The model of a finished epoch, still on its way to the host.  `resolve()` (called by Trainer.update() on the
    Learner's thread) waits for the side-stream copy only, prints the epoch's loss line, rebuilds the CPU model in
    eval mode and caches its pickled bytes on it (the Learner pickles the model for every worker request,
    train.py:605-615)."""

    def __init__(self, stepper, done_event, host_state, host_losses, heads, template):
        self.stepper, self.done, self.host_state, self.host_losses = stepper, done_event, host_state, host_losses
        self.heads, self.template = heads, template

    def resolve(self):
        self.done.synchronize()
        sums = dict(zip(LOSS_KEYS, self.host_losses.tolist()))
        dcnt = sums['dcnt']
        if dcnt > 0:
            print('loss = %s' % ' '.join([k + ':' + '%.3f' % (sums[k] / dcnt) for k in self.heads]))
        tpl = self.template
        state = self.stepper.state.state_dict_from(self.host_state, tpl.state_dict().keys())
        for k, b in self.stepper.state.loose:
            state[k] = b.detach().cpu()
        tpl.load_state_dict(state)
        tpl.eval()
        blob = pickle.dumps(tpl)
        model = pickle.loads(blob)                        # == copy.deepcopy(tpl), and leaves the bytes for the workers
        attach_pickle_cache(model, blob)
        return model, sums


def attach_pickle_cache(model, blob):
    """pickle.dumps(model) / copy.deepcopy(model) of this instance replay `blob` instead of walking the module tree:
    the instance-level __reduce_ex__ makes every later pickle of the (immutable until the next epoch) model a memcpy.
    What the workers unpickle is the plain nn.Module that `blob` holds."""
    object.__setattr__(model, '__reduce_ex__', lambda protocol, _b=blob: (pickle.loads, (_b,)))
    return model


class LearnerStep:
    """One replay batch -> one optimiser step.

    step(packed):  ONE H2D copy of the packed pinned batch -> [CUDA graph: net forward ->
    fused loss fwd+bwd kernel -> net backward -> (all-reduce SUM) -> clip + Adam].
    The six loss sums land in `last_losses` / `loss_accum` on the device; nothing in here
    synchronises the host.
    """

    def __init__(self, model, args, example_batch, lr, device=None, process_group=None, use_graph=True,
                 max_norm=4.0, weight_decay=1e-5, time_loss_kernel=False, channels_last=True, cudnn_benchmark=True,
                 small_boards=True, peer_allreduce=None, allow_tf32=None, fused_tower=True, tensor_cores=None):
        self.device = torch.device(device if device is not None else 'cuda')
        self.args = args
        # the learner owns its precision contract (1e-5 of the reference's fp32 arithmetic): PyTorch's default lets
        # cuDNN convolutions run on TF32 tensor cores (10-bit mantissa).  train_args['allow_tf32'] = True opts out.
        if allow_tf32 is None:
            allow_tf32 = bool(args.get('allow_tf32', False))
        self.allow_tf32 = allow_tf32
        torch.backends.cudnn.allow_tf32 = allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = allow_tf32
        self.model = model.to(self.device)
        # cuDNN's default heuristics pick FFT / NCHW-spatial kernels that are 5x slower than its NHWC
        # implicit-GEMM kernels on the tiny boards of these games; NHWC + autotune is a pure layout choice
        # tiny boards: convolutions as one SGEMM, BatchNorm as fused reductions (fastnet.py); NCHW stays as is
        self.engine = None          # hand-scheduled fused forward/backward for recognised architectures (tower.py)
        # train_args['tensor_cores'] = False: the small-board dense products stay on fp32 SIMT kernels (strict fp32 summation)
        if tensor_cores is None:
            tensor_cores = bool(args.get('tensor_cores', True))
        self.tensor_cores = tensor_cores
        fused_tower = fused_tower and tensor_cores
        self.rewritten = fastnet.optimize_small_boards(self.model, tensor_cores=tensor_cores) if small_boards else 0
        if self.rewritten and channels_last:
            channels_last = not self._uses_dense_convs(example_batch)       # dense products want NCHW-flat activations
        self.memory_format = torch.channels_last if channels_last else None
        if channels_last:
            self.model = self.model.to(memory_format=torch.channels_last)
        if cudnn_benchmark:
            torch.backends.cudnn.benchmark = True
        self.model.train()
        self.time_loss_kernel = time_loss_kernel
        self.kernel_events = []
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        params = [p for p in self.model.parameters()]
        # gradient exchange: fused one-shot all-reduce over NVLink peer memory (default when sharded), or NCCL
        if peer_allreduce is None:
            peer_allreduce = self.world > 1 and os.environ.get('HRL_PEER_ALLREDUCE', '1') != '0'
        self.peer = ops.PeerAllReduce(self.pg, self.device) if (peer_allreduce and self.world > 1) else None
        self.state = StateStore(self.model, self.device)
        self.opt = ops.FlatAdam(params, lr=lr, weight_decay=weight_decay, max_norm=max_norm, extra=NUM_LOSS,
                                grad_alloc=self.peer.alloc if self.peer is not None else None,
                                param_storage=self.state.flat_param)
        self.state.index_params(self.model)
        self.state_snap = torch.empty_like(self.state.bytes)
        self.acc_snap = torch.zeros(NUM_LOSS, dtype=torch.float64, device=self.device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._handoff_slots = None
        self._handoff_i = 0
        self._staging = None
        self._staging_i = 0
        self.launches_per_step = 0
        self.ema = torch.full((1,), float(example_batch['action'].shape[0] * args.get('forward_steps', 1)) * self.world,
                              dtype=torch.float32, device=self.device)

        self.layout = BatchLayout(example_batch)
        self.dev_buffer = torch.zeros(self.layout.nbytes, dtype=torch.uint8, device=self.device)
        self.dev = self.layout.views(self.dev_buffer)
        self.h2d_bytes = self.layout.nbytes
        B, T, Pa, A = example_batch['action_mask'].shape
        P = example_batch['turn_mask'].shape[2]
        self.dims = (B, T, P, Pa, A)
        self.hidden0 = None
        if hasattr(self.model, 'init_hidden'):
            self.hidden0 = tree_map(lambda h: h.to(self.device), self.model.init_hidden([B, P]))
        from . import tower
        if fused_tower and small_boards and self.hidden0 is None and tower.supports(self.model) and \
                torch.is_tensor(example_batch['observation']) and example_batch['observation'].shape[-2] * example_batch['observation'].shape[-1] <= 16:
            self.engine = tower.FusedBoardNet(self.model, B * T * Pa, self.device)
        self.loss_buf = None
        self.last_losses = torch.zeros(NUM_LOSS, device=self.device)
        self.loss_accum = torch.zeros(NUM_LOSS, dtype=torch.float64, device=self.device)
        self.host_slots = torch.zeros((8, NUM_LOSS)).pin_memory()
        self._slot = 0
        self.graph = self.graph_fwd = self.graph_bwd = None
        self.use_graph = use_graph
        self.steps = 0
        self.stream = torch.cuda.Stream(device=self.device)
        self._warm = PackedBatch(self.layout).fill(example_batch)

    def _uses_dense_convs(self, example_batch):
        """One no-grad probe call of the net on a two-sample slice: do its convolutions take the dense small-board path
        (then activations stay NCHW) or cuDNN (then NHWC, whose implicit-GEMM kernels are the fast ones)?"""
        before = fastnet.BoardConv2d.dense_calls
        obs = tree_map(lambda o: o[:1, :1].flatten(0, 2).to(self.device), example_batch['observation'])
        hidden = None
        if hasattr(self.model, 'init_hidden'):
            hidden = tree_map(lambda h: h.to(self.device), self.model.init_hidden([obs_rows(obs)]))
        was_training = self.model.training
        self.model.eval()
        with torch.no_grad():
            self.model(obs, hidden)
        self.model.train(was_training)
        return fastnet.BoardConv2d.dense_calls > before

    # -- the device work of one step, on the current stream (inputs already in self.dev)
    def _part_forward(self):
        fastnet.new_step()          # adjoint-weight copies of the previous step are stale: the optimiser has run
        self.opt.zero_grad()
        if self.engine is not None:
            B, T, P, Pa, A = self.dims
            flat = self.engine.forward(self.dev['observation'].flatten(0, 2))
            self._outs = {k: v.unflatten(0, (B, T, Pa)) for k, v in flat.items()}
        else:
            self._outs = forward_raw(self.model, self.hidden0, self.dev, self.args, self.memory_format)
        if self.loss_buf is None:
            B, T, P, Pa, A = self.dims
            self.loss_buf = ops.LossBuffers(B, T, P, Pa, A, 'value' in self._outs, 'return' in self._outs, self.device)

    def _part_loss(self):
        outs = self._outs
        ops.loss_fwd_bwd({k: outs[k] for k in ('policy', 'value', 'return') if k in outs}, self.dev, self.args,
                         buffers=self.loss_buf)

    def _part_backward(self):
        outs, buf = self._outs, self.loss_buf
        if self.engine is not None:
            self.engine.backward(buf.dpolicy.flatten(0, 2), buf.dvalue.flatten(0, 2),
                                 buf.dreturn.flatten(0, 2) if buf.dreturn is not None else None)
        else:
            heads, grads = [outs['policy']], [buf.dpolicy]
            if 'value' in outs:
                heads.append(outs['value'])
                grads.append(buf.dvalue)
            if 'return' in outs:
                heads.append(outs['return'])
                grads.append(buf.dreturn)
            with ops.deferred_weight_gradients():       # shared (recurrent) convolution weights: one product per weight, at the end
                torch.autograd.backward(heads, grads)
        self.opt.extra_slots[:NUM_LOSS].copy_(buf.losses)     # the loss sums ride the gradient bucket
        if self.peer is not None:
            reduced = self.peer(self.opt.n_pad, self.opt.partials)      # all-reduce + norm partials, one kernel
            self.opt.step_reduced(reduced)
            self.last_losses.copy_(reduced[self.opt.n_pad:self.opt.n_pad + NUM_LOSS])
        else:
            if self.world > 1:
                torch.distributed.all_reduce(self.opt.flat_grad, op=torch.distributed.ReduceOp.SUM, group=self.pg)
            self.opt.step()
            self.last_losses.copy_(self.opt.extra_slots[:NUM_LOSS])
        self.loss_accum.add_(self.last_losses)

    def _device_step(self):
        self._part_forward()
        self._part_loss()
        self._part_backward()

    def _capture(self):
        # warm-up on the step stream (cuDNN heuristics, lazy inits, NCCL communicator), then capture
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.dev_buffer.copy_(self._warm.buffer, non_blocking=True)
            state = self._snapshot()
            for i in range(3):
                before = ops.LAUNCHES['n']
                self._device_step()
                self.launches_per_step = ops.LAUNCHES['n'] - before      # this library's kernels in one step
            self.stream.synchronize()
            self._restore(state)
            if self.use_graph and not self.time_loss_kernel:
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=self.stream):
                    self._device_step()
                self._restore(state)
            elif self.use_graph:
                # the loss kernel is launched between two graphs so that CUDA events can bracket it
                self.graph_fwd = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_fwd, stream=self.stream):
                    self._part_forward()
                self._part_loss()
                self.graph_bwd = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_bwd, pool=self.graph_fwd.pool(), stream=self.stream):
                    self._part_backward()
                self._restore(state)
            self.stream.synchronize()
        self._captured = True

    def _snapshot(self):
        bufs = {k: v.clone() for k, v in self.model.state_dict().items()}
        return (bufs, self.opt.exp_avg.clone(), self.opt.exp_avg_sq.clone(), self.opt.step_count.clone(),
                self.loss_accum.clone())

    def _restore(self, state):
        bufs, m, v, sc, acc = state
        with torch.no_grad():
            for k, t in self.model.state_dict().items():
                t.copy_(bufs[k])
            self.opt.exp_avg.copy_(m)
            self.opt.exp_avg_sq.copy_(v)
            self.opt.step_count.copy_(sc)
            self.loss_accum.copy_(acc)

    def new_packed(self):
        return PackedBatch(self.layout)

    def warm_up(self):
        """Run the warm-up steps and capture the graphs now (otherwise done lazily by the first step)."""
        if not getattr(self, '_captured', False):
            self._capture()

    def step(self, packed):
        """Enqueue H2D + one learner step for a PackedBatch; returns without waiting for the GPU.
        `packed.in_flight` tells when its host buffer may be refilled."""
        self._enqueue(packed.buffer, packed)

    def step_in_place(self):
        """Same step when the inputs were written directly into self.dev (GPU replay gather) on the step stream."""
        self._enqueue(None, None)

    def step_resident(self, dev_bytes):
        """Same step with the packed batch already in HBM (e.g. produced by the replay gather kernel)."""
        self._enqueue(dev_bytes, None)

    def _enqueue(self, src_bytes, packed):
        if not getattr(self, '_captured', False):
            self._capture()
        if packed is not None:
            # host batch: the H2D copy runs on the copy stream into one of two staging buffers, i.e. while the previous
            # step still computes; the step stream only waits for it and moves the bytes device-to-device (~1 us)
            if self._staging is None:
                self._staging = [torch.empty_like(self.dev_buffer) for _ in range(2)]
                self._staging_free = [None, None]
            k = self._staging_i % 2
            self._staging_i += 1
            with torch.cuda.stream(self.copy_stream):
                if self._staging_free[k] is not None:
                    self.copy_stream.wait_event(self._staging_free[k])
                self._staging[k].copy_(src_bytes, non_blocking=True)
                arrived = torch.cuda.Event()
                arrived.record(self.copy_stream)
            packed.in_flight = arrived
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(arrived)
                self.dev_buffer.copy_(self._staging[k], non_blocking=True)
                freed = torch.cuda.Event()
                freed.record(self.stream)
                self._staging_free[k] = freed
            src_bytes = None
        with torch.cuda.stream(self.stream):
            if src_bytes is not None:
                self.dev_buffer.copy_(src_bytes, non_blocking=True)
            if not self.use_graph:
                self._device_step()
            elif not self.time_loss_kernel:
                self.graph.replay()
            else:
                self.graph_fwd.replay()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(self.stream)
                self._part_loss()
                e1.record(self.stream)
                self.kernel_events.append((e0, e1))
                self.graph_bwd.replay()
        self.steps += 1

    def loss_kernel_ms(self):
        """Average device time of the fused loss kernel over the steps since the last call
        (needs time_loss_kernel=True)."""
        self.stream.synchronize()
        ts = [a.elapsed_time(b) for a, b in self.kernel_events]
        self.kernel_events = []
        return sum(ts) / max(1, len(ts)), len(ts)

    def fetch_losses_async(self):
        """Enqueue the D2H copy of the last step's six loss sums (24 bytes) behind that step and
        return a handle; handle() waits for just that copy and returns the dict."""
        slot = self.host_slots[self._slot % self.host_slots.shape[0]]
        self._slot += 1
        with torch.cuda.stream(self.stream):
            slot.copy_(self.last_losses, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)

        def get():
            ev.synchronize()
            return dict(zip(LOSS_KEYS, slot.tolist()))
        return get

    def read_losses(self):
        """Blocking read of the last step's six loss sums."""
        return self.fetch_losses_async()()

    def pop_accumulated(self):
        """Loss sums accumulated since the last call (ONE host sync per epoch)."""
        self.stream.synchronize()
        vals = self.loss_accum.cpu().tolist()
        self.loss_accum.zero_()
        return dict(zip(LOSS_KEYS, vals))

    def close(self):
        """Release the captured graphs (they pin NCCL kernels: destroy them before the process group)."""
        self.stream.synchronize()
        self.graph = self.graph_fwd = self.graph_bwd = None
        self._outs = None
        self._captured = False
        import gc
        gc.collect()
        torch.cuda.synchronize(self.device)

    def cpu_state_dict(self):
        """Blocking copy of the model's state_dict to the host (tests, checkpoints)."""
        self.stream.synchronize()
        host = self.state.bytes.cpu()
        keys = [k for k in self.model.state_dict().keys() if not k.endswith('_sel_cache')]
        out = self.state.state_dict_from(host, keys)
        for k, b in self.state.loose:
            out[k] = b.detach().cpu().clone()
        return {k: out[k] for k in keys}

    def epoch_schedule(self, batch_cnt, steps, default_lr):
        """Device half of the epoch boundary, on the current stream: move the epoch's loss sums aside and apply the
        learning-rate schedule from the (all-reduced, hence global) data count.  Every rank of a sharded learner
        enqueues this at the same step, so all ranks keep identical learning rates without a broadcast."""
        self.acc_snap.copy_(self.loss_accum)
        self.loss_accum.zero_()
        dcnt = self.acc_snap[5:6].float()
        fresh = self.ema * 0.8 + dcnt * (0.2 / (1e-2 + batch_cnt))
        self.ema.copy_(torch.where(dcnt > 0, fresh, self.ema))
        self.opt.lr.copy_(self.ema * (default_lr / (1 + steps * 1e-5)))

    def end_epoch(self, batch_cnt, steps, default_lr, template, heads):
        """Epoch boundary without a host synchronisation (train.py:378-387): on the step stream, snapshot the loss
        sums and the whole model state device-to-device, apply the learning-rate schedule ON THE DEVICE
        (data_cnt_ema <- 0.8 ema + 0.2 dcnt/(0.01+batch_cnt); lr <- 3e-8 ema/(1+steps 1e-5), train.py:382-384 -- dcnt is
        the all-reduced global count) and let a side stream carry the snapshot to pinned host memory while the
        next epoch's steps already run.  Returns a PendingModel."""
        if self._handoff_slots is None:
            self._handoff_slots = [(torch.empty(self.state.bytes.numel(), dtype=torch.uint8).pin_memory(),
                                    torch.zeros(NUM_LOSS, dtype=torch.float64).pin_memory()) for _ in range(2)]
        host_state, host_losses = self._handoff_slots[self._handoff_i % 2]
        self._handoff_i += 1
        with torch.cuda.stream(self.stream):
            self.epoch_schedule(batch_cnt, steps, default_lr)
            if getattr(self, '_last_done', None) is not None:
                self.stream.wait_event(self._last_done)          # the previous snapshot has left the device buffer
            self.state_snap.copy_(self.state.bytes)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            host_state.copy_(self.state_snap, non_blocking=True)
            host_losses.copy_(self.acc_snap, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        self._last_done = done
        return PendingModel(self, done, host_state, host_losses, heads, template)


# --------------------------------------------------------------------------- batcher + trainer

class _FlatCache:
    """Decoded episodes for the host batcher, least-recently-used first, bounded in bytes (the reference keeps only
    the bz2 blocks, train.py:54; an unbounded cache of decoded arrays would grow to tens of GB at
    maximum_episodes = 100000 and trip the Learner's memory guard)."""

    def __init__(self, budget_bytes):
        self.budget, self.used = int(budget_bytes), 0
        self.items = collections.OrderedDict()       # id(episode) -> (episode, FlatEpisode, bytes)
        self.lock = threading.Lock()

    @staticmethod
    def _size(fe):
        return sum(a.nbytes for a in tree_leaves(fe.obs)) + fe.amask.nbytes + fe.value.nbytes + 24 * fe.prob.size

    def get(self, ep):
        from .wire import episode_to_flat
        key = id(ep)
        with self.lock:
            hit = self.items.get(key)
            if hit is not None and hit[0] is ep:
                self.items.move_to_end(key)
                return hit[1]
        fe = episode_to_flat(ep)          # flat wire format if the worker sent it, else decode the moments
        size = self._size(fe)
        with self.lock:
            self.items[key] = (ep, fe, size)
            self.used += size
            while self.used > self.budget and len(self.items) > 1:
                _, (_, _, sz) = self.items.popitem(last=False)
                self.used -= sz
        return fe


class Batcher:
    """Feeds the trainer: recency-biased window sampling (train.py:291-315) and collation.

    Unlike the reference there is no process pool shipping pickled batches through pipes
    (train.py:274, connection.py:133-173): episodes are decoded once into FlatEpisode arrays
    (a byte-bounded LRU cache) and batches are built by `num_batchers` threads with numpy
    gathers that release the GIL.
    """

    def __init__(self, args, episodes):
        self.args = args
        self.episodes = episodes
        self.out = queue.Queue(maxsize=8)
        self.threads = []
        self.started = False
        self.stop_event = threading.Event()
        self.cache = _FlatCache(args.get('host_cache_bytes', 2 << 30))

    def _fetch(self, idx):
        ep = self.episodes[idx]           # ONE read: the deque may shift under us (train.py:298-302)
        return ep['steps'], ep

    def select_episode(self):
        idx, st, ed, tst, ep = sample_window(lambda: len(self.episodes), self._fetch, self.args)
        cs = self.args['compress_steps']
        b0, b1 = st // cs, (ed - 1) // cs + 1
        return {'args': ep['args'], 'outcome': ep['outcome'], 'moment': ep['moment'][b0:b1], 'base': b0 * cs,
                'start': st, 'end': ed, 'train_start': tst, 'total': ep['steps'], '_episode': ep}

    def _make(self):
        windows = []
        for _ in range(self.args['batch_size']):
            sel = self.select_episode()
            fe = self.cache.get(sel['_episode'])
            windows.append((fe, sel['start'], sel['end'], sel['start'], sel['train_start'], sel['total']))
        nb = gather_windows(windows, self.args)
        return tree_map(lambda a: torch.from_numpy(a), nb)

    def _worker(self, bid):
        print('started batcher %d' % bid)
        while not self.stop_event.is_set():
            try:
                item = self._make()
            except Exception:              # a batcher thread must never die silently: Batcher.batch() would spin forever
                traceback.print_exc()
                time.sleep(0.05)
                continue
            while not self.stop_event.is_set():
                try:
                    self.out.put(item, timeout=0.2)
                    break
                except queue.Full:
                    continue
        print('finished batcher %d' % bid)

    def run(self):
        if self.started:
            return
        self.started = True
        for i in range(self.args['num_batchers']):
            th = threading.Thread(target=self._worker, args=(i,), daemon=True)
            th.start()
            self.threads.append(th)

    def batch(self):
        while True:
            try:
                return self.out.get(timeout=0.2)
            except queue.Empty:
                if self.stop_event.is_set():
                    return None

    def stop(self):
        self.stop_event.set()
        for th in self.threads:
            th.join(timeout=5)


class EpisodeDeque(deque):
    """The `Trainer.episodes` deque the Learner appends to (train.py:472, 482-483), with a tap: every appended
    episode is also handed to a listener (the GPU replay feeder) exactly once."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.listener = None

    def append(self, ep):
        super().append(ep)
        if self.listener is not None:
            self.listener(ep)

    def extend(self, eps):
        for ep in eps:
            self.append(ep)


class GpuBatcher:
    """Batcher on the GPU-resident replay (replay.py + the gather/pad kernel): arriving episodes are decoded once
    by a feeder thread and uploaded into the device ring; a batch is B window descriptors (drawn with array
    operations, same sampling law as Batcher.select_episode) + ONE kernel that writes straight into the learner
    step's input buffer.

    Feeder and learner never wait for each other's GPU work on the host: the upload stream waits (on the device) for
    the last gather that may read rows it overwrites, the step stream waits for the last upload; the only shared
    host lock covers "sample + enqueue gather" on one side and "update directory + enqueue copies" on the other
    (microseconds each: staging into pinned memory happens outside it)."""

    DESC_SLOTS = 4        # pinned descriptor buffers in rotation: bounds how far the host runs ahead of the GPU

    def __init__(self, args, episodes, device, seed=None, forward=None):
        from .replay import DeviceReplay
        from .wire import episode_to_flat
        self.args = args
        self.device = device
        self.forward = forward              # multi-GPU: callable(list of episodes) that ships them to the other ranks
        self.pending = queue.Queue()
        self.upload_stream = torch.cuda.Stream(device=device)
        self.last_upload = None
        self.last_gather = None
        self.order_lock = threading.Lock()
        self.stop_event = threading.Event()
        self.rng = np.random.default_rng(seed if seed is not None else args.get('seed', 0) * 7919 + 17)
        self.fed = 0
        self._slots = None
        self._slot_i = 0

        # tap first, then the backlog: nothing is missed; what the tap saw meanwhile is not enqueued twice, and the
        # backlog keeps its (recency) order ahead of it
        tapped, tap_lock = [], threading.Lock()

        def tap(ep):
            with tap_lock:
                if tapped is not None and self._backlog_open:
                    tapped.append(ep)
                    return
            self.pending.put(ep)

        self._backlog_open = True
        episodes.listener = tap
        backlog = list(episodes)
        with tap_lock:
            seen = {id(e) for e in tapped}
            backlog = [e for e in backlog if id(e) not in seen] + tapped
            self._backlog_open = False
        for ep in backlog:
            self.pending.put(ep)
        self.backlog_n = len(backlog)

        # ring capacity from the observed episode lengths and the free HBM (the reference bounds episodes, not steps)
        fe0 = episode_to_flat(backlog[0]) if backlog else None
        cap = int(args.get('replay_capacity_steps', 0))
        if not cap:
            lens = [e['steps'] for e in backlog] or [64]
            mean_len, max_len = sum(lens) / len(lens), max(lens)
            want = int(args['maximum_episodes'] * mean_len * 1.25) + 4 * max_len
            budget = want
            if fe0 is not None and torch.device(device).type == 'cuda':
                free, _ = torch.cuda.mem_get_info(device)
                budget = int(args.get('replay_memory_fraction', 0.5) * free / DeviceReplay.bytes_per_step(fe0))
            cap = max(4 * max_len, min(want, budget))
            est = int(cap / max(mean_len, 1))
            if est < args['maximum_episodes']:
                print('handyrl_b200: the GPU replay holds about %d episodes (%d steps), fewer than maximum_episodes=%d'
                      % (est, cap, args['maximum_episodes']))
        self.replay = DeviceReplay(cap, args['maximum_episodes'], device=device)
        self.thread = threading.Thread(target=self._feed, daemon=True)

    def run(self):
        if not self.thread.is_alive():
            self.thread.start()

    def _feed(self):
        from .wire import episode_to_flat
        while not self.stop_event.is_set():
            try:
                eps = [self.pending.get(timeout=0.2)]
            except queue.Empty:
                continue
            while len(eps) < 256:            # everything that has queued up goes in one upload
                try:
                    eps.append(self.pending.get_nowait())
                except queue.Empty:
                    break
            try:
                if self.forward is not None:
                    self.forward(eps)
                staged = self.replay.stage([episode_to_flat(ep) for ep in eps])
                with self.order_lock:
                    if self.last_gather is not None:   # device-side: never overwrite rows an enqueued gather reads
                        self.upload_stream.wait_event(self.last_gather)
                    with torch.cuda.stream(self.upload_stream):
                        self.replay.commit(staged)
                        ev = torch.cuda.Event()
                        ev.record(self.upload_stream)
                        self.last_upload = ev
            except Exception:
                traceback.print_exc()
            self.fed += len(eps)

    def ready(self):
        """True once the whole backlog the trainer started with is resident (the reference samples from at least
        `minimum_episodes` episodes from its first step on)."""
        return self.fed >= self.backlog_n and len(self.replay) > 0

    def _descriptor_slot(self, B):
        from .replay import WINDOW_DTYPE
        if self._slots is None:
            nb = B * WINDOW_DTYPE.itemsize
            self._slots = [{'host': torch.empty(nb, dtype=torch.uint8).pin_memory(),
                            'dev': torch.empty((B, WINDOW_DTYPE.itemsize), dtype=torch.uint8, device=self.device),
                            'event': None} for _ in range(self.DESC_SLOTS)]
        slot = self._slots[self._slot_i % self.DESC_SLOTS]
        self._slot_i += 1
        if slot['event'] is not None:
            slot['event'].synchronize()      # the gather that read this slot DESC_SLOTS steps ago
        return slot

    def fill(self, stepper):
        """Sample a batch and gather it into stepper.dev (on the step stream)."""
        B = stepper.dims[0]
        slot = self._descriptor_slot(B)
        with self.order_lock:
            win = self.replay.sample_windows(B, self.args, self.rng)
            slot['host'].numpy()[:] = win.view(np.uint8)
            with torch.cuda.stream(stepper.stream):
                if self.last_upload is not None:
                    stepper.stream.wait_event(self.last_upload)
                slot['dev'].view(-1).copy_(slot['host'], non_blocking=True)
                out = dict(stepper.dev)
                single_leaf = torch.is_tensor(stepper.dev['observation'])
                if single_leaf:
                    out['observation'] = stepper.dev['observation'].view(*stepper.dev['observation'].shape[:3], -1)
                else:
                    out['observation'] = self._flat_obs(stepper)
                out['value'] = self._value_sink(stepper)
                self.replay.gather(slot['dev'], self.args, out=out)
                if not single_leaf:
                    nested = self.replay.split_observation(out['observation'])
                    for d, s_ in zip(tree_leaves(stepper.dev['observation']), tree_leaves(nested)):
                        d.copy_(s_)
                ev = torch.cuda.Event()
                ev.record(stepper.stream)
                self.last_gather = ev
                slot['event'] = ev

    def _flat_obs(self, stepper):
        if not hasattr(self, '_obs_buf'):
            B, T, Pa = stepper.dev['action'].shape[:3]
            self._obs_buf = torch.empty((B, T, Pa, self.replay.OE), device=self.device)
        return self._obs_buf

    def _value_sink(self, stepper):
        if not hasattr(self, '_val_buf'):
            B, T, P = stepper.dev['turn_mask'].shape[:3]
            self._val_buf = torch.empty((B, T, P, 1), device=self.device)    # behaviour value: unused by the loss
        return self._val_buf

    def stop(self):
        self.stop_event.set()
        if self.thread.is_alive():
            self.thread.join(timeout=5)


class Trainer:
    """Drop-in for handyrl.train.Trainer (train.py:321-400): same constructor, attributes
    (`episodes`, `steps`), `run()` thread body and `update()` hand-off, same printed lines.

    Multi-GPU (the reference wraps its model in nn.DataParallel when it sees several GPUs, train.py:325, 339-340):
    with train_args['num_gpus'] > 1 (default: every visible GPU, as the reference) this process is rank 0 and spawns
    one helper process per further GPU (multigpu.py); every rank holds the whole replay, samples batch_size/num_gpus
    windows per step and the gradient bucket is all-reduced (SUM) inside the captured step."""

    def __init__(self, args, model):
        self.episodes = EpisodeDeque()
        self.args = args
        self.model = model
        self.gpu_replay = bool(args.get('gpu_replay', True))
        self.gpu_batcher = None
        self.default_lr = 3e-8
        self.data_cnt_ema = self.args['batch_size'] * self.args['forward_steps']
        self.params = list(self.model.parameters())
        self.lr = self.default_lr * self.data_cnt_ema
        self.steps = 0
        self.batcher = Batcher(self.args, self.episodes)
        self.update_flag = False
        self.update_queue = queue.Queue(maxsize=1)
        self.stepper = None
        self.fleet = None
        self.stop_event = threading.Event()
        if len(self.params) > 0 and not torch.cuda.is_available():
            raise RuntimeError('handyrl_b200.Trainer needs a CUDA device; there is no CPU learner in this package')
        self.world = 1
        if len(self.params) > 0:
            want = args.get('num_gpus')
            self.world = max(1, min(int(want), torch.cuda.device_count()) if want else torch.cuda.device_count())
            if self.world > 1 and (not self.gpu_replay or args['batch_size'] % self.world != 0):
                print('handyrl_b200: multi-GPU needs gpu_replay and batch_size %% num_gpus == 0; using one GPU')
                self.world = 1

    def update(self):
        """Called by the Learner (train.py:342-345, 533): ends the running epoch and returns (CPU model in eval
        mode, steps).  The trainer thread only enqueues the hand-off; the wait for the side-stream copy, the
        state_dict rebuild and the pickling happen here, on the caller's thread."""
        while True:
            self.update_flag = True
            item, steps = self.update_queue.get()
            if not isinstance(item, PendingModel):
                return item, steps
            model, sums = item.resolve()
            if sums['dcnt'] > 0:           # train.py:357: an epoch needs at least one sample with a turn in it
                break
        # host mirrors of the schedule the device applied (train.py:382-384), for inspection / logging
        self.data_cnt_ema = self.data_cnt_ema * 0.8 + sums['dcnt'] / (1e-2 + item.batch_cnt) * 0.2
        self.lr = self.default_lr * self.data_cnt_ema / (1 + steps * 1e-5)
        return model, steps

    def _start_stepper(self):
        # the first batch is built on the host: it fixes every shape of the captured step
        batch = self._first_host_batch()
        self.cpu_template = copy.deepcopy(self.model)
        self.cpu_template.eval()
        pg = None
        if self.world > 1:
            from . import multigpu
            batch = tree_map(lambda t: t[:t.shape[0] // self.world].contiguous(), batch)
            self.fleet = multigpu.Fleet(self.world, self.args, self.cpu_template, list(self.episodes), self.lr)
            pg = self.fleet.process_group
        self.stepper = LearnerStep(self.model, self.args, batch, self.lr, process_group=pg)
        self.stepper.warm_up()
        self.batcher.pool = [self.stepper.new_packed() for _ in range(4)]
        if self.gpu_replay:
            self.gpu_batcher = GpuBatcher(self.args, self.episodes, self.stepper.device,
                                          forward=self.fleet.send_episodes if self.fleet is not None else None)
            self.gpu_batcher.run()
        loss_buf = self.stepper.loss_buf
        self.heads = ['p'] + (['v'] if loss_buf.dvalue is not None else []) + \
            (['r'] if loss_buf.dreturn is not None else []) + ['ent', 'total']

    def train(self):
        if len(self.params) == 0:          # non-parametric model (train.py:348-350)
            time.sleep(0.1)
            return self.model
        batch_cnt = 0
        chunk = int(self.args.get('multi_gpu_chunk', 8))
        while True:
            if self.stop_event.is_set() and (self.fleet is None or batch_cnt % chunk == 0):
                return None                 # (sharded: only between chunks -- the other ranks run whole chunks)
            if self.stepper is None:
                self._start_stepper()
            if self.gpu_batcher is not None:
                while not (self.gpu_batcher.ready() and (self.fleet is None or self.fleet.all_ready())):
                    if self.stop_event.is_set():
                        return None
                    time.sleep(0.01)
                if self.fleet is not None and batch_cnt % chunk == 0:
                    self.fleet.run_steps(chunk)          # every rank runs exactly the steps rank 0 runs
                self.gpu_batcher.fill(self.stepper)
                self.stepper.step_in_place()
            else:
                batch = self.batcher.batch()
                if batch is None:           # stop() was called
                    return None
                packed = self.batcher.pool[batch_cnt % len(self.batcher.pool)]
                packed.wait_reusable()
                packed.fill(batch)
                self.stepper.step(packed)
            batch_cnt += 1
            self.steps += 1
            if self.update_flag and (self.fleet is None or batch_cnt % chunk == 0):
                break
        # epoch boundary: nothing here waits for the GPU (LearnerStep.end_epoch)
        if self.fleet is not None:
            self.fleet.end_epoch(batch_cnt, self.steps, self.default_lr)
        pending = self.stepper.end_epoch(batch_cnt, self.steps, self.default_lr, self.cpu_template, self.heads)
        pending.batch_cnt = batch_cnt
        return pending

    def _first_host_batch(self):
        return self.batcher._make()

    def run(self):
        print('waiting training')
        while len(self.episodes) < self.args['minimum_episodes']:
            if self.stop_event.is_set():
                return
            time.sleep(1)
        if len(self.params) > 0:
            if not self.gpu_replay:
                self.batcher.run()
            print('started training')
        while not self.stop_event.is_set():
            model = self.train()
            if model is None:
                break
            self.update_flag = False
            while not self.stop_event.is_set():
                try:
                    self.update_queue.put((model, self.steps), timeout=0.2)
                    break
                except queue.Full:
                    continue
        print('finished training')

    def stop(self):
        """Not in the reference (its threads die with the process): lets tests and embedders shut the
        trainer down cleanly -- stops the batcher threads, the helper ranks, and ends run()."""
        self.stop_event.set()
        self.batcher.stop()
        if self.gpu_batcher is not None:
            self.gpu_batcher.stop()
        if self.stepper is not None:
            self.stepper.stream.synchronize()
        if self.fleet is not None:
            self.fleet.stop()
            self.stepper.close()
            self.fleet.destroy()


def install(flat_episodes=False):
    """Swap the reference's learner hot path for this one in an importable `handyrl` package
    (flat_episodes=True additionally makes workers forked afterwards ship the flat wire format of wire.py):
    after `handyrl_b200.train.install()`, `python main.py --train` runs the reference's Learner,
    workers and server unchanged on top of this Trainer."""
    import handyrl.train as ref
    ref.Trainer = Trainer
    ref.Batcher = Batcher
    ref.make_batch = make_batch
    ref.forward_prediction = forward_prediction
    ref.compute_loss = compute_loss
    import handyrl.losses as ref_losses
    ref_losses.compute_target = ops.compute_target
    if flat_episodes:
        from .wire import install_worker_hook
        install_worker_hook()
    return ref
