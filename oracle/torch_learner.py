"""TEST INFRASTRUCTURE -- eager-PyTorch CPU restatement of the reference learner step.

This is the "port" that bench.py times as `cpu_baseline` and as `--impl reference`:
the reference itself (pure Python on PyTorch) cannot travel to the GPU box, so its learner
step is restated here with the same kind of work the reference does on a CPU -- eager ATen
ops, a Python loop over T for every recurrence, autograd for the backward, then
clip_grad_norm_ + torch.optim.Adam (reference handyrl/train.py:127-267, 366-371;
handyrl/losses.py:16-80).  Only tests/, __graft_entry__.smoke() and bench.py may import it.

Parity status: PINNED -- tests/test_oracle.py::test_torch_port_* checks losses and autograd
gradients against tests/golden/loss_cases.npz (outputs of the reference).
"""
import torch
import torch.nn.functional as F


def _reverse_scan(kind, base, boot, rew, lam, gamma, rho, c):
    """One column-wise target recurrence; all tensors (B,T,P).  Returns (target, advantage).
    losses.py:20-60."""
    T = base.shape[1]
    if kind in ('TD', 'UPGO'):
        carry = boot
        rows = [carry]
        for t in range(T - 2, -1, -1):
            nxt = base[:, t + 1]
            blend = (1 - lam[:, t + 1]) * nxt + lam[:, t + 1] * carry
            if kind == 'UPGO':
                blend = torch.maximum(nxt, blend)
            carry = rew[:, t] + gamma * blend
            rows.append(carry)
        tgt = torch.stack(rows[::-1], 1)
        return tgt, tgt - base
    # V-Trace
    nxt = torch.cat([base[:, 1:], boot.unsqueeze(1)], 1)
    delta = rho * (rew + gamma * nxt - base)
    carry = delta[:, -1]
    rows = [carry]
    for t in range(T - 2, -1, -1):
        carry = delta[:, t] + gamma * lam[:, t + 1] * c[:, t] * carry
        rows.append(carry)
    vs = torch.stack(rows[::-1], 1) + base
    vs_nxt = torch.cat([vs[:, 1:], boot.unsqueeze(1)], 1)
    return vs, rew + gamma * vs_nxt - base


def _target(kind, base, all_returns, rew, lmb, gamma, rho, c, mask):
    """losses.py:63-80 on (B,T,P) tensors; `all_returns` is (B,T,P) or (B,1,P)."""
    if base is None:
        r = all_returns.expand_as(mask)
        return r, r
    if kind == 'MC':
        return all_returns.expand_as(base), all_returns - base
    lam = lmb + (1 - lmb) * (1 - mask)
    zero = torch.zeros_like(base) if rew is None else rew
    return _reverse_scan(kind, base, all_returns[:, -1], zero, lam, gamma, rho.expand_as(base), c.expand_as(base))


def loss_from_raw(raw, batch, args):
    """Mask epilogue + compute_loss + compose_losses from the net's RAW outputs (B,T,Pa,...)
    (train.py:176-184, 189-267).  Returns ({p,v,r,ent,total} tensors, dcnt tensor)."""
    bi = args.get('burn_in_steps', 0)
    tm4, om4 = batch['turn_mask'], batch['observation_mask']
    pol = raw['policy'] * tm4
    if pol.size(2) > 1 and batch['action'].size(2) == 1:
        pol = pol.sum(2, keepdim=True)
    pol = pol - batch['action_mask']
    heads = {k: raw[k] * om4 for k in ('value', 'return') if k in raw}

    def cut(x):
        return x[:, bi:] if (bi > 0 and x.size(1) > 1) else x

    pol = cut(pol)
    heads = {k: cut(v) for k, v in heads.items()}
    tm, om = cut(tm4).squeeze(-1), cut(om4).squeeze(-1)                    # (B,T,P)
    em = cut(batch['episode_mask']).squeeze(-1)                            # (B,T,1)
    act, mu = cut(batch['action']), cut(batch['selected_prob'])
    oc = batch['outcome'].squeeze(-1)                                      # (B,1,P)
    rew, ret = cut(batch['reward']).squeeze(-1), cut(batch['return']).squeeze(-1)
    prog = cut(batch['progress'])                                          # (B,T,1)

    logp_all = F.log_softmax(pol, -1)
    logp = logp_all.gather(-1, act).squeeze(-1) * em                       # (B,T,Pa)
    logmu = torch.log(mu.clamp(1e-16, 1)).squeeze(-1) * em
    rho = torch.exp(logp.detach() - logmu).clamp(0, 1)                     # rho-bar == c-bar (both thresholds 1)

    v_base = r_base = None
    v_mask = om
    if 'value' in heads:
        v = heads['value'].detach().squeeze(-1)
        if args['turn_based_training'] and v.size(2) == 2:
            v_opp, om_opp = -v.flip(2), om.flip(2)
            v = (v * om + v_opp * om_opp) / (om + om_opp + 1e-8)
            v_mask = (om + om_opp).clamp(0, 1)
        v_base = v * em + oc * (1 - em)
    if 'return' in heads:
        r_base = heads['return'].detach().squeeze(-1)

    vt, pt = args['value_target'], args['policy_target']
    tg_v, adv_v = _target(vt, v_base, oc, None, args['lambda'], 1.0, rho, rho, v_mask)
    tg_r, adv_r = _target(vt, r_base, ret, rew, args['lambda'], args['gamma'], rho, rho, om)
    if pt != vt:
        _, adv_v = _target(pt, v_base, oc, None, args['lambda'], 1.0, rho, rho, v_mask)
        _, adv_r = _target(pt, r_base, ret, rew, args['lambda'], args['gamma'], rho, rho, om)
    adv = rho * (adv_v + adv_r)

    out = {'p': (-logp * adv * tm).sum()}
    if 'value' in heads:
        out['v'] = ((heads['value'].squeeze(-1) - tg_v) ** 2 * om).sum() / 2
    if 'return' in heads:
        out['r'] = (F.smooth_l1_loss(heads['return'].squeeze(-1), tg_r.expand_as(om), reduction='none') * om).sum()
    probs = logp_all.exp()
    ent = -(probs * logp_all.clamp(min=torch.finfo(logp_all.dtype).min)).sum(-1) * tm
    out['ent'] = ent.sum()
    reg = (ent * (1 - prog * (1 - args['entropy_regularization_decay']))).sum()
    out['total'] = out['p'] + out.get('v', 0) + out.get('r', 0) - args['entropy_regularization'] * reg
    return out, tm.sum()


def _walk(fn, *trees):
    t = trees[0]
    if isinstance(t, dict):
        return {k: _walk(fn, *[x[k] for x in trees]) for k in t}
    if isinstance(t, (list, tuple)):
        return type(t)(_walk(fn, *[x[i] for x in trees]) for i in range(len(t)))
    return fn(*trees)


def recurrent_raw_outputs(net, hidden, batch, args):
    """The reference's sequential pass over T for nets with a hidden state (train.py:147-174): the hidden state is
    masked by observation_mask before each call (summed over players in the turn-alternating layout), burn-in steps run
    in eval mode without gradient, and the new hidden state replaces the old one only where the player observed."""
    B, T, Pa = batch['action'].shape[:3]
    alternating = args['turn_based_training'] and not args['observation']
    rows = {}
    for t in range(T):
        obs = _walk(lambda o: o[:, t].flatten(0, 1), batch['observation'])
        seen = batch['observation_mask'][:, t]
        shaped = _walk(lambda h: seen.view(*h.shape[:2], *([1] * (h.dim() - 2))), hidden)
        visible = _walk(lambda h, m: h * m, hidden, shaped)
        visible = _walk((lambda h: h.sum(1)) if alternating else (lambda h: h.flatten(0, 1)), visible)
        if t < args['burn_in_steps']:
            net.eval()
            with torch.no_grad():
                out = net(obs, visible)
        else:
            if not net.training:
                net.train()
            out = net(obs, visible)
        fresh = _walk(lambda h: h.unflatten(0, (B, Pa)), out.pop('hidden'))
        for k, v in out.items():
            if v is not None:
                rows.setdefault(k, []).append(v.unflatten(0, (B, Pa)))
        hidden = _walk(lambda h, n, m: h * (1 - m) + n * m, hidden, fresh, shaped)
    return {k: torch.stack(v, 1) for k, v in rows.items()}


class CpuLearner:
    """Learner step on the host: net forward (one call, or the sequential pass for recurrent nets) -> loss ->
    autograd -> clip -> Adam."""

    def __init__(self, net, args, lr, weight_decay=1e-5, max_norm=4.0):
        self.net, self.args, self.max_norm = net, args, max_norm
        self.params = list(net.parameters())
        self.opt = torch.optim.Adam(self.params, lr=lr, weight_decay=weight_decay)
        self.net.train()

    def step(self, batch):
        B, T, Pa = batch['action'].shape[:3]
        if hasattr(self.net, 'init_hidden'):
            P = batch['turn_mask'].shape[2]
            raw = recurrent_raw_outputs(self.net, self.net.init_hidden([B, P]), batch, self.args)
        else:
            outs = self.net(_walk(lambda o: o.flatten(0, 2), batch['observation']), None)
            raw = {k: v.unflatten(0, (B, T, Pa)) for k, v in outs.items() if v is not None and k != 'hidden'}
        losses, dcnt = loss_from_raw(raw, batch, self.args)
        self.opt.zero_grad()
        losses['total'].backward()
        self.grad_norm = float(torch.nn.utils.clip_grad_norm_(self.params, self.max_norm))
        self.opt.step()
        return {k: float(v.detach()) for k, v in losses.items()}, float(dcnt)
