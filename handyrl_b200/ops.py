"""Thin Python front end of the C ABI: torch tensors in, torch tensors out.

torch is used for device memory and streams only; every computation happens in the
CUDA library (handyrl_b200/libhrl_b200.so, sources in handyrl_b200/csrc/).  There is no
fallback path: a CPU tensor or a missing library raises.
"""
import ctypes as C

import torch

from . import _capi
from ._capi import ALGO_ID, LOSS_KEYS, NUM_LOSS, HrlLossArgs, check, lib

# kernels of THIS library launched through the wrappers below (bench.py reports them as `gpu_launches`; a CUDA graph
# replays the launches counted while it was captured)
LAUNCHES = {'n': 0}


def _count(n=1):
    LAUNCHES['n'] += n


_BATCH_KEYS = ('action_mask', 'action', 'selected_prob', 'reward', 'return', 'turn_mask', 'observation_mask',
               'episode_mask', 'progress', 'outcome')


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_f32(t, name):
    if t is None:
        return None
    if not t.is_cuda:
        raise _capi.HrlError('handyrl_b200: %s must be a CUDA tensor (no CPU path exists)' % name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def algo_id(name):
    try:
        return ALGO_ID[name]
    except KeyError:
        # the reference prints 'No algorithm named ...' and returns None (losses.py:79-80)
        raise ValueError('No algorithm named %s' % name)


class LossBuffers:
    """Pre-allocated outputs of one fused loss launch (re-used across steps / graph replays)."""

    def __init__(self, B, T, P, Pa, A, has_value, has_return, device, taps=False, policy_dtype=torch.float32):
        f = dict(dtype=torch.float32, device=device)
        self.dims = (B, T, P, Pa, A)
        self.dpolicy = torch.empty((B, T, Pa, A), dtype=policy_dtype, device=device)      # bf16 logits get bf16 gradients
        self.dvalue = torch.empty((B, T, Pa, 1), **f) if has_value else None
        self.dreturn = torch.empty((B, T, Pa, 1), **f) if has_return else None
        self.losses = torch.zeros(NUM_LOSS, **f)
        self.taps = None
        if taps:
            self.taps = {
                'target_value': torch.zeros((B, T, P, 1), **f), 'target_return': torch.zeros((B, T, P, 1), **f),
                'advantage': torch.zeros((B, T, P, 1), **f), 'logp': torch.zeros((B, T, Pa, 1), **f),
                'rho': torch.zeros((B, T, Pa, 1), **f), 'entropy': torch.zeros((B, T, Pa), **f),
            }
        self.workspace = torch.zeros(lib().hrl_loss_workspace_bytes(B, T, P, Pa, A), dtype=torch.uint8, device=device)


def loss_fwd_bwd(outputs, batch, args, buffers=None, taps=False, tuning=None):
    """Fused mask epilogue + compute_loss + closed-form backward (reference train.py:176-267).

    outputs: raw net outputs {'policy': (B,T,Pa,A), 'value': (B,T,Pa,1)?, 'return': (B,T,Pa,1)?}
    batch:   the make_batch dict (device tensors)
    args:    train_args (lambda, gamma, entropy_regularization[_decay], policy_target, value_target,
             turn_based_training, burn_in_steps)
    tuning:  None (the library chooses), or a dict for tests / profiling with any of
             variant ('rows-direct'|'rows-staged'|'bulk'|'element'|'group'), recurrence ('serial'|'scan'),
             cluster, consumers, threads, unstaged, trace (an int64 CUDA tensor of >= 32 elements)
    returns  LossBuffers with .losses = [p, v, r, ent, total, dcnt] and the gradients.
    """
    policy = outputs['policy']
    io_bf16 = policy.dtype == torch.bfloat16       # wide rows only: 8 instead of 12 bytes per action through HBM (include/hrl_b200.h)
    if io_bf16:
        if not (policy.is_cuda and policy.is_contiguous()):
            raise _capi.HrlError('handyrl_b200: bf16 policy logits must be a contiguous CUDA tensor')
    else:
        policy = _dev_f32(policy, 'policy')
    B, T, Pa, A = policy.shape
    P = batch['turn_mask'].shape[2]
    value = _dev_f32(outputs.get('value'), 'value')
    ret_head = _dev_f32(outputs.get('return'), 'return')
    if buffers is None:
        buffers = LossBuffers(B, T, P, Pa, A, value is not None, ret_head is not None, policy.device, taps=taps, policy_dtype=policy.dtype)
    assert buffers.dims == (B, T, P, Pa, A) and buffers.dpolicy.dtype == policy.dtype

    # static buffers (CUDA-graph replays): the argument block of the previous call is still valid
    key = (policy.data_ptr(), 0 if value is None else value.data_ptr(), 0 if ret_head is None else ret_head.data_ptr(),
           args['value_target'], args['policy_target'], bool(args['turn_based_training']), args.get('burn_in_steps', 0),
           args['lambda'], args['gamma'], args['entropy_regularization'], args['entropy_regularization_decay'],
           buffers.taps is not None, None if tuning is None else tuple(sorted((k, v if not torch.is_tensor(v) else v.data_ptr())
                                                                                 for k, v in tuning.items()))) + \
        tuple(batch[k].data_ptr() for k in _BATCH_KEYS)
    cached = getattr(buffers, '_cached', None)
    if cached is not None and cached[0] == key:
        check(lib().hrl_loss_fwd_bwd(C.byref(cached[1]), _stream_ptr()))
        _count()
        return buffers

    a = HrlLossArgs()
    a.B, a.T, a.P, a.Pa, a.A = B, T, P, Pa, A
    a.burn_in = int(args.get('burn_in_steps', 0))
    a.value_target = algo_id(args['value_target'])
    a.policy_target = algo_id(args['policy_target'])
    a.two_player_zero_sum = int(bool(args['turn_based_training']) and P == 2)
    a.lambda_ = float(args['lambda'])
    a.gamma = float(args['gamma'])
    a.entropy_regularization = float(args['entropy_regularization'])
    a.entropy_regularization_decay = float(args['entropy_regularization_decay'])

    action = batch['action']
    if not action.is_cuda:
        raise _capi.HrlError('handyrl_b200: the batch must live on the GPU')
    keep = [policy, value, ret_head,
            _dev_f32(batch['action_mask'], 'action_mask'), action.long().contiguous(),
            _dev_f32(batch['selected_prob'], 'selected_prob'), _dev_f32(batch['reward'], 'reward'),
            _dev_f32(batch['return'], 'return'), _dev_f32(batch['turn_mask'], 'turn_mask'),
            _dev_f32(batch['observation_mask'], 'observation_mask'), _dev_f32(batch['episode_mask'], 'episode_mask'),
            _dev_f32(batch['progress'], 'progress'), _dev_f32(batch['outcome'], 'outcome')]
    (a.policy_raw, a.value_raw, a.return_raw, a.action_mask, a.action, a.selected_prob, a.reward, a.ret,
     a.turn_mask, a.observation_mask, a.episode_mask, a.progress, a.outcome) = [_ptr(t) for t in keep]
    a.dpolicy_raw, a.dvalue_raw, a.dreturn_raw = _ptr(buffers.dpolicy), _ptr(buffers.dvalue), _ptr(buffers.dreturn)
    a.losses = _ptr(buffers.losses)
    if buffers.taps is not None:
        t = buffers.taps
        a.tap_target_value, a.tap_target_return, a.tap_advantage = _ptr(t['target_value']), _ptr(t['target_return']), _ptr(t['advantage'])
        a.tap_logp, a.tap_rho, a.tap_entropy = _ptr(t['logp']), _ptr(t['rho']), _ptr(t['entropy'])
    a.io_bf16 = int(io_bf16)
    a.workspace = _ptr(buffers.workspace)
    a.workspace_bytes = buffers.workspace.numel()
    if tuning:
        t = dict(tuning)
        a.tuning.variant = _capi.LOSS_VARIANTS[t.pop('variant', 'auto')]
        a.tuning.recurrence = _capi.LOSS_RECURRENCES[t.pop('recurrence', 'auto')]
        trace = t.pop('trace', None)
        if trace is not None:
            assert trace.is_cuda and trace.dtype == torch.int64 and trace.numel() >= 32
            a.tuning.trace = trace.data_ptr()
        for k in ('cluster', 'consumers', 'threads', 'unstaged'):
            setattr(a.tuning, k, int(t.pop(k, 0)))
        if t:
            raise ValueError('unknown loss tuning keys: %s' % sorted(t))
    check(lib().hrl_loss_fwd_bwd(C.byref(a), _stream_ptr()))
    _count()
    buffers._keep = keep  # the launch is asynchronous: keep temporaries alive
    if all(k is None or k is o for k, o in zip(keep[3:], (batch['action_mask'], batch['action'], batch['selected_prob'],
                                                          batch['reward'], batch['return'], batch['turn_mask'],
                                                          batch['observation_mask'], batch['episode_mask'],
                                                          batch['progress'], batch['outcome']))):
        buffers._cached = (key, a)   # only when no temporary copies were made
    return buffers


def compute_target(algorithm, values, returns, rewards, lmb, gamma, rhos, cs, masks):
    """Drop-in for handyrl.losses.compute_target on CUDA tensors of shape (B,T,P,1)."""
    aid = algo_id(algorithm)
    returns = _dev_f32(returns, 'returns')
    if values is None:  # losses.py:64-66
        return returns, returns
    values = _dev_f32(values, 'values')
    B, T, P = values.shape[:3]
    rewards, rhos, cs, masks = (_dev_f32(x, n) for x, n in ((rewards, 'rewards'), (rhos, 'rhos'), (cs, 'cs'), (masks, 'masks')))
    if masks is not None and masks.shape[2] != P:
        masks = masks.expand(B, T, P, 1).contiguous()
    Tr = returns.shape[1]
    if returns.shape[2] != P:
        returns = returns.expand(B, Tr, P, 1).contiguous()
    Pr = rhos.shape[2] if rhos is not None else 1
    targets = torch.empty_like(values)
    advantages = torch.empty_like(values)
    check(lib().hrl_compute_target(aid, B, T, P, Tr, Pr, _ptr(values), _ptr(returns), _ptr(rewards), float(lmb),
                                   float(gamma), _ptr(rhos), _ptr(cs), _ptr(masks), _ptr(targets), _ptr(advantages),
                                   _stream_ptr()))
    _count()
    return targets, advantages


class FlatAdam:
    """clip_grad_norm_(max_norm) + Adam(weight_decay) on one flat fp32 bucket (train.py:331, 370-371).

    The model's parameters are re-pointed into one contiguous buffer and their .grad into a
    second one, so that (a) the multi-GPU gradient exchange is ONE all-reduce(SUM) and
    (b) the whole optimiser step is two kernel launches.  `extra` floats are appended to the
    gradient bucket (the learner puts the six loss scalars there so they ride the same
    all-reduce).
    """

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5, max_norm=4.0, extra=0, grad_alloc=None,
                 param_storage=None):
        params = [p for p in params]
        assert len(params) > 0 and all(p.is_cuda and p.dtype == torch.float32 for p in params)
        self.params = params
        device = params[0].device
        self.n = sum(p.numel() for p in params)
        n_pad = (self.n + 3) // 4 * 4
        self.extra = extra
        # param_storage lets the caller place the weights inside a larger state buffer (LearnerStep: weights + BatchNorm
        # buffers in one allocation, so the per-epoch model hand-off is ONE device-to-host copy)
        if param_storage is not None:
            assert param_storage.numel() == n_pad and param_storage.dtype == torch.float32 and param_storage.data_ptr() % 16 == 0
        self.flat_param = param_storage if param_storage is not None else torch.zeros(n_pad, dtype=torch.float32, device=device)
        extra = (extra + 3) // 4 * 4
        self.extra = extra
        # grad_alloc(numel) lets the caller place the bucket in NVLink-symmetric memory (peer all-reduce)
        self.grad_storage = grad_alloc(n_pad + extra) if grad_alloc is not None else None
        self.flat_grad = (self.grad_storage[:n_pad + extra] if self.grad_storage is not None
                          else torch.zeros(n_pad + extra, dtype=torch.float32, device=device))
        off = 0
        with torch.no_grad():
            for p in params:
                k = p.numel()
                self.flat_param[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat_param[off:off + k].view_as(p)
                p.grad = self.flat_grad[off:off + k].view_as(p)
                off += k
        self.n_pad = n_pad
        self.exp_avg = torch.zeros(n_pad, dtype=torch.float32, device=device)
        self.exp_avg_sq = torch.zeros(n_pad, dtype=torch.float32, device=device)
        self.partials = torch.zeros(lib().hrl_sumsq_num_partials(), dtype=torch.float32, device=device)
        self.lr = torch.full((1,), float(lr), dtype=torch.float32, device=device)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=device)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=device)
        self.betas, self.eps, self.weight_decay, self.max_norm = betas, eps, weight_decay, max_norm

    @property
    def extra_slots(self):
        return self.flat_grad[self.n_pad:]

    def set_lr(self, lr):
        self.lr.fill_(float(lr))

    def zero_grad(self):
        self.flat_grad.zero_()

    def step_reduced(self, reduced):
        """clip + Adam on an already all-reduced bucket whose sum-of-squares partials are in self.partials
        (hrl_peer_allreduce_sumsq)."""
        check(lib().hrl_clip_adam_step(_ptr(self.flat_param), _ptr(reduced), _ptr(self.exp_avg), _ptr(self.exp_avg_sq),
                                       self.n_pad, _ptr(self.partials), _ptr(self.lr), _ptr(self.step_count), self.max_norm,
                                       self.betas[0], self.betas[1], self.eps, self.weight_decay, _ptr(self.grad_norm),
                                       _stream_ptr()))
        _count(2)

    def step(self):
        s = _stream_ptr()
        check(lib().hrl_grad_sumsq(_ptr(self.flat_grad), self.n_pad, _ptr(self.partials), s))
        check(lib().hrl_clip_adam_step(_ptr(self.flat_param), _ptr(self.flat_grad), _ptr(self.exp_avg),
                                       _ptr(self.exp_avg_sq), self.n_pad, _ptr(self.partials), _ptr(self.lr),
                                       _ptr(self.step_count), self.max_norm, self.betas[0], self.betas[1], self.eps,
                                       self.weight_decay, _ptr(self.grad_norm), s))
        _count(3)
        from . import fastnet
        fastnet.new_step()          # cached adjoint weights of the convolutions are stale now


def gemm_tf32x3(a, b, bias=None, a_kmajor=True, b_kmajor=True, splits=1, out=None):
    """C = A_op @ B_op^T (+ bias) on the tensor cores with fp32-class accuracy (hrl_gemm_tf32x3, csrc/gemm_kernel.cu).

    a: (M, K) if a_kmajor else (K, M) -- the operand as it lies in memory; b likewise (N, K) / (K, N).
    """
    assert a.is_cuda and b.is_cuda and a.dtype == torch.float32 and b.dtype == torch.float32
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[0], a.shape[1]) if a_kmajor else (a.shape[1], a.shape[0])
    N, Kb = (b.shape[0], b.shape[1]) if b_kmajor else (b.shape[1], b.shape[0])
    assert K == Kb, (a.shape, b.shape, a_kmajor, b_kmajor)
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    ws = None
    if splits > 1:
        ws = torch.empty(lib().hrl_gemm_workspace_floats(M, N, K, splits), dtype=torch.float32, device=a.device)
    check(lib().hrl_gemm_tf32x3(_ptr(a), a.stride(0), int(a_kmajor), _ptr(b), b.stride(0), int(b_kmajor), _ptr(bias), _ptr(out),
                                out.stride(0), M, N, K, splits, _ptr(ws), _stream_ptr()))
    _count(2 if splits > 1 else 1)
    return out


class _LinearTC(torch.autograd.Function):
    """y = x @ w^T with all three products (forward, input gradient, weight gradient) on the tensor cores at fp32-class
    accuracy (hrl_gemm_tf32x3).  The weight gradient reduces over the rows of x (samples): both operands are read
    transposed on the fly and the reduction is split over enough K slices to fill the GPU."""

    @staticmethod
    def forward(ctx, x, w):
        x, w = x.contiguous(), w.contiguous()
        ctx.save_for_backward(x, w)
        return gemm_tf32x3(x, w)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = gemm_tf32x3(dy, w, b_kmajor=False)                       # (M,N) x (N,K): w is read as stored
        if ctx.needs_input_grad[1]:
            M = x.shape[0]
            tiles = ((w.shape[0] + 127) // 128) * ((w.shape[1] + 287) // 288)
            splits = max(1, min(M // 64, 148 // tiles))
            dw = gemm_tf32x3(dy, x, a_kmajor=False, b_kmajor=False, splits=splits)
        return dx, dw


def linear_tc(x, w):
    return _LinearTC.apply(x, w)


class _BoardDense(torch.autograd.Function):
    """Dense matrix of a small-board convolution from its weight (hrl_board_expand) and the adjoint (hrl_board_fold)."""

    @staticmethod
    def forward(ctx, weight, H, W):
        weight = weight.contiguous()
        Cout, Cin, kh, kw = weight.shape
        ctx.dims = (Cout, Cin, kh, kw, H, W)
        dense = torch.empty((Cout * H * W, Cin * H * W), dtype=torch.float32, device=weight.device)
        check(lib().hrl_board_expand(_ptr(weight), _ptr(dense), Cout, Cin, kh, kw, H, W, _stream_ptr()))
        _count()
        return dense

    @staticmethod
    def backward(ctx, ddense):
        Cout, Cin, kh, kw, H, W = ctx.dims
        ddense = ddense.contiguous()
        dw = torch.empty((Cout, Cin, kh, kw), dtype=torch.float32, device=ddense.device)
        check(lib().hrl_board_fold(_ptr(ddense), 1, 0, _ptr(dw), Cout, Cin, kh, kw, H, W, _stream_ptr()))
        _count()
        return dw, None, None


def board_dense(weight, H, W):
    return _BoardDense.apply(weight, H, W)


class _BoardConv(torch.autograd.Function):
    """A stride-1 "same" convolution over a tiny board, NCHW in and out, as dense products on the tensor cores:
    forward   y = x2d @ dense(w)^T            (hrl_board_expand + hrl_gemm_tf32x3)
    backward  dx = dy2d @ dense(w)            (hrl_gemm_tf32x3, the dense matrix read as stored)
              dw = fold(sum_s dy2d_s^T x2d_s) (split-K hrl_gemm_tf32x3 leaving its slice partials, folded AND summed by
                                               one hrl_board_fold launch: no dense gradient is ever materialised)"""

    @staticmethod
    def forward(ctx, x, weight):
        x = x.contiguous()
        weight = weight.contiguous()
        N, Cin, H, W = x.shape
        Cout, _, kh, kw = weight.shape
        dense = torch.empty((Cout * H * W, Cin * H * W), dtype=torch.float32, device=x.device)
        check(lib().hrl_board_expand(_ptr(weight), _ptr(dense), Cout, Cin, kh, kw, H, W, _stream_ptr()))
        _count()
        y = gemm_tf32x3(x.view(N, Cin * H * W), dense)
        ctx.save_for_backward(x, dense)
        ctx.dims = (N, Cin, H, W, Cout, kh, kw)
        return y.view(N, Cout, H, W)

    @staticmethod
    def backward(ctx, dy):
        x, dense = ctx.saved_tensors
        N, Cin, H, W, Cout, kh, kw = ctx.dims
        dy2 = dy.contiguous().view(N, Cout * H * W)
        x2 = x.view(N, Cin * H * W)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = gemm_tf32x3(dy2, dense, b_kmajor=False).view(N, Cin, H, W)
        if ctx.needs_input_grad[1]:
            rows, cols = Cout * H * W, Cin * H * W
            tiles = ((rows + 127) // 128) * ((cols + 287) // 288)
            splits = lib().hrl_gemm_effective_splits(N, max(1, min(N // 64, 148 // tiles)))
            dw = torch.empty((Cout, Cin, kh, kw), dtype=torch.float32, device=dy.device)
            if splits > 1:
                ws = torch.empty(splits * rows * cols, dtype=torch.float32, device=dy.device)
                check(lib().hrl_gemm_tf32x3(_ptr(dy2), dy2.stride(0), 0, _ptr(x2), x2.stride(0), 0, None, None, cols, rows, cols, N,
                                            splits, _ptr(ws), _stream_ptr()))
            else:
                ws = gemm_tf32x3(dy2, x2, a_kmajor=False, b_kmajor=False).view(-1)
                _count(-1)
            check(lib().hrl_board_fold(_ptr(ws), splits, rows * cols, _ptr(dw), Cout, Cin, kh, kw, H, W, _stream_ptr()))
            _count(2)
        return dx, dw


def board_conv(x, weight):
    return _BoardConv.apply(x, weight)


# ---- stride-1 "same" / wrap-around convolutions as implicit tensor-core products (hrl_gemm_fused conv_mode 1 / 2) -----------
_CONV_TABLES = {}        # (H, W, kh, kw, wrap, device) -> int16 neighbour-offset table on the device
_CONV_IMAGES = {}        # (weight ptr, shape) -> [forward image, adjoint image, generation they were packed in]
_CONV_GENERATION = [0]   # bumped whenever the parameters may have changed (fastnet.new_step)


def conv_weights_changed():
    _CONV_GENERATION[0] += 1


def _conv_table(H, W, kh, kw, wrap, device):
    key = (H, W, kh, kw, bool(wrap), str(device))
    t = _CONV_TABLES.get(key)
    if t is None:
        import numpy as np
        host = np.empty(H * W * kh * kw, dtype=np.int16)
        check(lib().hrl_conv_geometry(H, W, kh, kw, int(bool(wrap)), host.ctypes.data))
        t = _CONV_TABLES[key] = torch.from_numpy(host).to(device)
    return t


def _conv_images(w):
    """The weight's packed forward / adjoint operand images, re-packed once per parameter generation.  (Keyed by address AND
    identity: a freed model's weight address can be handed to another tensor of the same shape.)"""
    import weakref
    Cout, Cin, kh, kw = w.shape
    key = (w.data_ptr(), tuple(w.shape))
    ent = _CONV_IMAGES.get(key)
    if ent is None or ent[3]() is not w:
        if ent is None:
            z = lambda rows, ch: torch.zeros(lib().hrl_conv_pack_floats(rows, ch, kh * kw), dtype=torch.float32, device=w.device)
            ent = _CONV_IMAGES[key] = [z(Cout, Cin), z(Cin, Cout), None, None]
        ent[2], ent[3] = None, weakref.ref(w)
    gen = (_CONV_GENERATION[0], w._version)
    if ent[2] != gen:
        check(lib().hrl_conv_pack(_ptr(w), Cout, Cin, kh, kw, _ptr(ent[0]), _ptr(ent[1]), _stream_ptr()))
        _count()
        ent[2] = gen
    return ent[0], ent[1]


def conv_implicit_supported(x, w):
    """Shapes the implicit products cover: fp32 CUDA, at most 256 cells and 9 taps, channel counts that are multiples of 4 (16-byte
    pixel rows) and at most 288 (one operand tile)."""
    if not (x.is_cuda and x.dtype == torch.float32 and w.dtype == torch.float32 and x.dim() == 4):
        return False
    Cout, Cin, kh, kw = w.shape
    return (x.shape[2] * x.shape[3] <= 256 and kh * kw <= 9 and kh % 2 == 1 and kw % 2 == 1 and Cin % 4 == 0 and Cout % 4 == 0
            and Cin <= 288 and Cout <= 288 and x.shape[1] == Cin)


def _pixels(t):
    """(N, C, H, W) tensor -> its channels-last (N*H*W, C) view (copying only if it is not channels-last already)."""
    t = t.contiguous(memory_format=torch.channels_last)
    return t, t.permute(0, 2, 3, 1).reshape(-1, t.shape[1])


def _conv_product(pix, image, rows, cin, taps, table, hw, bias=None):
    """out[pixel][row] = sum over (tap, channel) of pix[neighbour(pixel, tap)][channel] * image[row][tap, channel]"""
    M = pix.shape[0]
    out = torch.empty((M, rows), dtype=torch.float32, device=pix.device)
    g = _capi.HrlGemmArgs()
    g.a.ptr, g.a.ld, g.a.kmajor = _ptr(pix), pix.stride(0), 1
    g.b.ptr, g.b.kmajor, g.b.packed = _ptr(image), 1, 1
    g.bias, g.C, g.ldc = _ptr(bias), _ptr(out), rows
    g.M, g.N, g.K, g.splits = M, rows, taps * ((cin + 31) // 32 * 32), 1
    g.conv_off, g.conv_mode, g.conv_hw, g.conv_taps, g.conv_cin = _ptr(table), 1, hw, taps, cin
    check(lib().hrl_gemm_fused(C.byref(g), _stream_ptr()))
    _count()
    return out


# Deferred weight gradients.  A recurrent net applies the same convolution at every time step (and DRC cells several times per step):
# computing its weight gradient per application means T x repeats small products, slice reductions and gradient accumulations --
# ~1,000 launches of a 4,400-launch Geister step, in a chain that is launch-bound.  Inside `deferred_weight_gradients()` the backward of
# conv_implicit only records its (dy, x) pair; on exit ONE segmented product per weight reduces over all its pairs (hrl_gemm_fused with
# `segments`), its ones row yields the bias gradient, and hrl_conv_wgrad_reduce2 adds the result into weight.grad / bias.grad.
_DEFER = {'on': False, 'pending': {}}


class deferred_weight_gradients:
    def __enter__(self):
        assert not _DEFER['on']
        _DEFER['on'], _DEFER['pending'] = True, {}
        return self

    def __exit__(self, exc_type, exc, tb):
        pending, _DEFER['pending'], _DEFER['on'] = _DEFER['pending'], {}, False
        if exc_type is None:
            for job in pending.values():
                _flush_weight_gradient(job)
        return False


def _flush_weight_gradient(job):
    w, b, pairs = job['w'], job['b'], job['pairs']
    Cout, Cin, kh, kw = w.shape
    taps, (table, hw) = kh * kw, job['geom']
    # the ones row (bias gradient as one more column) is free unless that column opens a new 288-wide tile AND the extra tile's CTAs
    # take K slices away from the others (a weight used once: 1 tile x 148 slices vs 2 tiles x 74): then a column sum does it
    cols, pixels0 = taps * Cin, pairs[0][0].shape[0]
    n0 = min(len(pairs), 64)

    def slices(ncols_):
        t_ = ((Cout + 127) // 128) * ((ncols_ + 287) // 288)
        return lib().hrl_gemm_effective_splits(pixels0, max(1, min(pixels0 // 64, 148 // (t_ * n0))))
    ones = b is not None and slices(cols + 1) >= slices(cols)
    if b is not None and not ones:
        if b.grad is None:
            b.grad = torch.zeros_like(b)
        for dy2, _ in pairs:
            b.grad.add_(dy2.sum(0))
    ncols = cols + (1 if ones else 0)
    pixels = pairs[0][0].shape[0]
    tiles = ((Cout + 127) // 128) * ((ncols + 287) // 288)
    for start in range(0, len(pairs), 64):
        chunk = pairs[start:start + 64]
        n = len(chunk)
        per = lib().hrl_gemm_effective_splits(pixels, max(1, min(pixels // 64, 148 // (tiles * n))))
        ws = torch.empty((n * per, Cout, ncols), dtype=torch.float32, device=w.device)
        seg_a = (C.c_void_p * n)(*[dy2.data_ptr() for dy2, _ in chunk])
        seg_b = (C.c_void_p * n)(*[x2.data_ptr() for _, x2 in chunk])
        g = _capi.HrlGemmArgs()
        g.a.ptr, g.a.ld, g.a.kmajor = _ptr(chunk[0][0]), Cout, 0
        g.b.ptr, g.b.ld, g.b.kmajor = _ptr(chunk[0][1]), Cin, 0
        g.C, g.ldc = None, ncols
        g.M, g.N, g.K, g.splits, g.workspace = Cout, ncols, pixels, per, _ptr(ws)
        g.conv_off, g.conv_mode, g.conv_hw, g.conv_taps, g.conv_cin = _ptr(table), 2, hw, taps, Cin
        g.seg_a, g.seg_b = C.cast(seg_a, C.c_void_p), C.cast(seg_b, C.c_void_p)
        g.segments, g.conv_ones_row = n, int(ones)
        check(lib().hrl_gemm_fused(C.byref(g), _stream_ptr()))
        for t, shape in ((w, w.shape), (b, None)):
            if t is not None and t.grad is None:
                t.grad = torch.zeros_like(t, memory_format=torch.contiguous_format)
        assert w.grad.is_contiguous()
        check(lib().hrl_conv_wgrad_reduce2(_ptr(ws), n * per, ncols, _ptr(w.grad), _ptr(b.grad) if ones else None, Cout, Cin, taps, 1,
                                           _stream_ptr()))
        _count(2)


class _ConvImplicit(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, wrap):
        N, Cin, H, W = x.shape
        Cout, _, kh, kw = w.shape
        table = _conv_table(H, W, kh, kw, wrap, x.device)
        fwd, _ = _conv_images(w)
        xl, x2 = _pixels(x)
        y2 = _conv_product(x2, fwd, Cout, Cin, kh * kw, table, H * W, bias=b)
        ctx.save_for_backward(xl, w)
        ctx.wrap, ctx.has_bias = wrap, b is not None
        ctx.params = (w, b)             # the Parameter objects themselves (their .grad is what a deferred flush accumulates into)
        return y2.view(N, H, W, Cout).permute(0, 3, 1, 2)           # logical NCHW, channels-last in memory

    @staticmethod
    def backward(ctx, dy):
        xl, w = ctx.saved_tensors
        N, Cin, H, W = xl.shape
        Cout, _, kh, kw = w.shape
        taps = kh * kw
        table = _conv_table(H, W, kh, kw, ctx.wrap, xl.device)
        dyl, dy2 = _pixels(dy)
        dx = dw = db = None
        pw, pb = ctx.params
        if ctx.needs_input_grad[0]:
            _, adj = _conv_images(pw)           # (the object forward saw: the image cache checks identity)
            dx = _conv_product(dy2, adj, Cin, Cout, taps, table, H * W).view(N, H, W, Cin).permute(0, 3, 1, 2)
        if (ctx.needs_input_grad[1] and _DEFER['on'] and pw.is_leaf and (pb is None or (pb.is_leaf and ctx.needs_input_grad[2]))
                and (pw.grad is None or pw.grad.is_contiguous())):
            job = _DEFER['pending'].setdefault(pw.data_ptr(), {'w': pw, 'b': pb, 'pairs': [], 'geom': (table, H * W)})
            if not job['pairs'] or job['pairs'][0][0].shape[0] == dy2.shape[0]:      # (pairs of one product cover the same pixels)
                job['pairs'].append((dy2, xl.permute(0, 2, 3, 1).reshape(-1, Cin)))
                return dx, None, None, None
        if ctx.needs_input_grad[1]:
            x2 = xl.permute(0, 2, 3, 1).reshape(-1, Cin)
            pixels, cols = x2.shape[0], taps * Cin
            tiles = ((Cout + 127) // 128) * ((cols + 287) // 288)
            s = lib().hrl_gemm_effective_splits(pixels, max(1, min(pixels // 64, 148 // tiles)))
            ws = torch.empty((s, Cout, cols), dtype=torch.float32, device=xl.device)
            g = _capi.HrlGemmArgs()
            g.a.ptr, g.a.ld, g.a.kmajor = _ptr(dy2), Cout, 0
            g.b.ptr, g.b.ld, g.b.kmajor = _ptr(x2), Cin, 0
            g.C, g.ldc = (None if s > 1 else _ptr(ws)), cols
            g.M, g.N, g.K, g.splits, g.workspace = Cout, cols, pixels, s, (_ptr(ws) if s > 1 else None)
            g.conv_off, g.conv_mode, g.conv_hw, g.conv_taps, g.conv_cin = _ptr(table), 2, H * W, taps, Cin
            check(lib().hrl_gemm_fused(C.byref(g), _stream_ptr()))
            dw = torch.empty_like(w, memory_format=torch.contiguous_format)
            check(lib().hrl_conv_wgrad_reduce(_ptr(ws), s, _ptr(dw), Cout, Cin, taps, _stream_ptr()))
            _count(2)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy2.sum(0)
        return dx, dw, db, None


def conv_implicit(x, w, b=None, wrap=False):
    """Stride-1 convolution with `same` zero padding (wrap=False) or wrap-around padding on both axes (wrap=True) of a board of at most
    256 cells, forward / input gradient / weight gradient on the tensor cores at fp32-class accuracy (3xTF32)."""
    return _ConvImplicit.apply(x, w, b, bool(wrap))


def _is_channels_last(t):
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous()


class _LstmGates(torch.autograd.Function):
    """The kernels index (sample, 4 gates x CS, ...) blocks.  A channels-last tensor is exactly that with sample := pixel and
    CS := C (the gate maps of one pixel are contiguous), so channels-last gates (what the tensor-core convolution produces) are
    processed in place, no layout copy, and h', c' come out channels-last for the next convolution."""

    @staticmethod
    def forward(ctx, gates, c_prev):
        cl = _is_channels_last(gates)
        fmt = torch.channels_last if cl else torch.contiguous_format
        gates, c_prev = gates.contiguous(memory_format=fmt), c_prev.contiguous(memory_format=fmt)
        N, C4 = gates.shape[:2]
        S = gates[0, 0].numel()
        h, c = torch.empty_like(c_prev), torch.empty_like(c_prev)          # (preserve_format: channels-last stays channels-last)
        dims = (N * S, C4 // 4, 1) if cl else (N, C4 // 4, S)
        check(lib().hrl_lstm_gates_fwd(_ptr(gates), _ptr(c_prev), _ptr(h), _ptr(c), *dims, _stream_ptr()))
        _count()
        ctx.save_for_backward(gates, c_prev)
        ctx.dims, ctx.fmt = dims, fmt
        return h, c

    @staticmethod
    def backward(ctx, dh, dc):
        gates, c_prev = ctx.saved_tensors
        dgates, dc_prev = torch.empty_like(gates), torch.empty_like(c_prev)
        # (named, not temporaries: a converted copy freed before the launch is enqueued could be handed to the next conversion)
        dh_ = None if dh is None else dh.contiguous(memory_format=ctx.fmt)
        dc_ = None if dc is None else dc.contiguous(memory_format=ctx.fmt)
        check(lib().hrl_lstm_gates_bwd(_ptr(gates), _ptr(c_prev), _ptr(dh_), _ptr(dc_), _ptr(dgates), _ptr(dc_prev), *ctx.dims,
                                       _stream_ptr()))
        _count()
        return dgates, dc_prev


def lstm_gates(gates, c_prev):
    """(h', c') of a convolutional LSTM cell from its gate pre-activations (N,4C,...) in the order i, f, o, g."""
    return _LstmGates.apply(gates, c_prev)


def _mask_view(om):
    """observation_mask[:, t] (B,P,1) as (pointer tensor, batch stride): element (b,p) at base + b*stride + p."""
    assert om.dim() == 3 and om.shape[2] == 1 and (om.stride(1) == 1 or om.shape[1] == 1) and om.dtype == torch.float32
    return om, om.stride(0)


# The hidden-state kernels treat a leaf as (B, P, R) blocks and are elementwise inside a block, so any dense ordering of the
# R elements works as long as every operand of a call shares it.  A leaf (B, P, C, H, W) whose (C, H, W) block is channels-last
# (what a net built on the tensor-core convolutions hands back) is therefore used as it lies: no layout copies in the time loop.
def _block_layout(t):
    if t.is_contiguous():
        return 'std'
    if t.dim() == 5 and t.stride(0) == t.shape[1] * t.stride(1) and _is_channels_last(t.flatten(0, 1)):
        return 'cl'
    if _is_channels_last(t):
        return 'cl'
    return None


def _as_block_layout(t, layout):
    if layout == 'cl' and t.dim() in (4, 5):
        if _block_layout(t) == 'cl':
            return t
        lead = t.shape[:-3]
        return t.reshape(-1, *t.shape[-3:]).contiguous(memory_format=torch.channels_last).view(*lead, *t.shape[-3:])
    return t.contiguous()


def _empty_block_layout(shape, layout, device):
    if layout == 'cl' and len(shape) in (4, 5):
        lead = tuple(shape[:-3])
        n = 1
        for d in lead:
            n *= d
        return torch.empty((n,) + tuple(shape[-3:]), dtype=torch.float32, device=device,
                           memory_format=torch.channels_last).view(*lead, *shape[-3:])
    return torch.empty(tuple(shape), dtype=torch.float32, device=device)


class _HiddenVisible(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, om, sum_players):
        layout = _block_layout(h) or 'std'
        h = _as_block_layout(h, layout)
        B, P = h.shape[:2]
        R = h[0, 0].numel()
        om, stride = _mask_view(om)
        out = _empty_block_layout((B,) + tuple(h.shape[2:]) if sum_players else tuple(h.shape), layout, h.device)
        check(lib().hrl_hidden_visible_fwd(_ptr(h), _ptr(om), stride, _ptr(out), B, P, R, int(sum_players), _stream_ptr()))
        _count()
        ctx.save_for_backward(om)
        ctx.meta = (B, P, R, stride, bool(sum_players), tuple(h.shape), layout)
        return out

    @staticmethod
    def backward(ctx, dout):
        om, = ctx.saved_tensors
        B, P, R, stride, sum_players, shape, layout = ctx.meta
        dh = _empty_block_layout(shape, layout, dout.device)
        dout = _as_block_layout(dout, layout)
        check(lib().hrl_hidden_visible_bwd(_ptr(dout), _ptr(om), stride, _ptr(dh), B, P, R, int(sum_players), _stream_ptr()))
        _count()
        return dh, None, None


class _HiddenBlend(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, nh, om):
        layout = _block_layout(nh) or 'std'          # the net's output decides: it keeps coming back in that layout
        h, nh = _as_block_layout(h, layout), _as_block_layout(nh, layout)
        B, P = h.shape[:2]
        Pn = nh.shape[1]
        R = h[0, 0].numel()
        om, stride = _mask_view(om)
        out = _empty_block_layout(tuple(h.shape), layout, h.device)
        check(lib().hrl_hidden_blend_fwd(_ptr(h), _ptr(nh), _ptr(om), stride, _ptr(out), B, P, Pn, R, _stream_ptr()))
        _count()
        ctx.save_for_backward(om)
        ctx.meta = (B, P, Pn, R, stride, tuple(h.shape), tuple(nh.shape), layout)
        return out

    @staticmethod
    def backward(ctx, dout):
        om, = ctx.saved_tensors
        B, P, Pn, R, stride, hshape, nshape, layout = ctx.meta
        dh = _empty_block_layout(hshape, layout, dout.device) if ctx.needs_input_grad[0] else None
        dnh = _empty_block_layout(nshape, layout, dout.device)
        dout = _as_block_layout(dout, layout)
        check(lib().hrl_hidden_blend_bwd(_ptr(dout), _ptr(om), stride, _ptr(dh), _ptr(dnh), B, P, Pn, R, _stream_ptr()))
        _count()
        return dh, dnh, None


def hidden_visible(h, om, sum_players):
    """The hidden leaf a recurrent net sees at one step (train.py:152-158): h (B,P,...) masked by om (B,P,1), summed
    over players when sum_players (turn-alternating batches) -- one kernel instead of mul + sum."""
    return _HiddenVisible.apply(h, om, sum_players)


def hidden_blend(h, nh, om):
    """h (1 - om) + nh om (train.py:173), nh (B,Pa,...) broadcast over players when Pa == 1 -- one kernel."""
    return _HiddenBlend.apply(h, nh, om)


def _nchw_or_nhwc(x):
    """(tensor in one of the two dense layouts, channels_last flag)"""
    if x.is_contiguous():
        return x, 0
    if x.is_contiguous(memory_format=torch.channels_last):
        return x, 1
    return x.contiguous(), 0


class _BatchNormTrain(torch.autograd.Function):
    """nn.BatchNorm2d (training mode) on (N, C, H, W) activations in NCHW or channels-last memory, through hrl_bn_train_fwd / _bwd."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum):
        x, cl = _nchw_or_nhwc(x)
        N, Cn, H, W = x.shape
        y = torch.empty_like(x)          # keeps the memory format
        mean = torch.empty(Cn, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        ws = torch.empty(lib().hrl_bn_workspace_floats(N, Cn, H * W, cl), dtype=torch.float32, device=x.device)
        check(lib().hrl_bn_train_fwd(_ptr(x), _ptr(weight), _ptr(bias), _ptr(y), _ptr(mean), _ptr(rstd), _ptr(running_mean),
                                     _ptr(running_var), N, Cn, H * W, cl, float(eps), float(momentum), _ptr(ws), _stream_ptr()))
        _count(3)
        ctx.save_for_backward(x, weight, mean, rstd)
        ctx.cl = cl
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, mean, rstd = ctx.saved_tensors
        cl = ctx.cl
        dy = dy.contiguous(memory_format=torch.channels_last) if cl else dy.contiguous()
        N, Cn, H, W = x.shape
        dx = torch.empty_like(x)
        dgamma = torch.empty(Cn, dtype=torch.float32, device=x.device)
        dbeta = torch.empty_like(dgamma)
        ws = torch.empty(lib().hrl_bn_workspace_floats(N, Cn, H * W, cl), dtype=torch.float32, device=x.device)
        check(lib().hrl_bn_train_bwd(_ptr(x), _ptr(dy), _ptr(weight), _ptr(mean), _ptr(rstd), _ptr(dx), _ptr(dgamma), _ptr(dbeta),
                                     N, Cn, H * W, cl, _ptr(ws), _stream_ptr()))
        _count(3)
        return dx, (dgamma if weight is not None else None), (dbeta if ctx.needs_input_grad[2] else None), None, None, None, None


def batch_norm_train(x, weight, bias, running_mean, running_var, eps, momentum):
    """Fused train-mode BatchNorm for small boards (updates the running statistics in place)."""
    return _BatchNormTrain.apply(x, weight, bias, running_mean, running_var, eps, momentum)


class PeerAllReduce:
    """One-shot all-reduce(SUM) of the flat gradient bucket over NVLink peer memory, fused with the gradient-norm
    partials (hrl_peer_allreduce_sumsq).  The bucket lives in torch symmetric memory so that every rank holds a
    mapping of every other rank's bucket; ranks synchronise inside the kernel with system-scope flags stored
    behind the bucket.  Bit-identical results on all ranks (fixed summation order), CUDA-graph capturable,
    no NCCL call on the step."""

    def __init__(self, group, device):
        import torch.distributed as dist
        self.group, self.device = group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.numel = None

    def alloc(self, numel):
        """Bucket of `numel` floats (+ 2*world flag words) in symmetric memory; rendezvous with the peers."""
        import torch.distributed._symmetric_memory as symm
        self.numel = numel
        words = numel + 2 * self.world + 64
        self.storage = symm.empty(words, dtype=torch.float32, device=self.device)
        self.storage.zero_()
        try:
            self.handle = symm.rendezvous(self.storage, self.group)
        except Exception:
            symm.enable_symm_mem_for_group(self.group.group_name)
            self.handle = symm.rendezvous(self.storage, self.group.group_name)
        torch.cuda.synchronize(self.device)
        self.handle.barrier()
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        assert len(ptrs) == self.world and ptrs[self.rank] == self.storage.data_ptr()
        self.peer_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        self.reduced = torch.zeros(numel, dtype=torch.float32, device=self.device)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        return self.storage

    def check(self):
        """Raise if a rank ever failed to arrive in the in-kernel rendezvous (the kernel gives up after 20 s instead of
        spinning forever; its sums are then garbage).  Synchronises the current stream."""
        if int(self.status.item()) != 0:
            raise _capi.HrlError('hrl_peer_allreduce_sumsq: a peer rank did not arrive within the timeout')

    def __call__(self, n_norm, partials):
        """Enqueue the fused reduce on the current stream; returns the reduced bucket."""
        check(lib().hrl_peer_allreduce_sumsq(_ptr(self.reduced), _ptr(self.peer_ptrs), self.numel, self.world, self.rank,
                                             self.numel, n_norm, _ptr(partials), _ptr(self.epoch), _ptr(self.ticket),
                                             _ptr(self.status), _stream_ptr()))
        _count(1)
        return self.reduced
