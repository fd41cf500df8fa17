"""Nets used by bench.py and the tests where the reference's env plug-ins cannot travel.

The learner itself takes ANY `nn.Module` honouring the reference's model contract
(forward(x, hidden) -> {'policy', 'value'?, 'return'?, 'hidden'?}); these are
only stand-ins with the architectures BASELINE.json's configs name, random-initialised.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class BoardNet(nn.Module):
    """Conv tower + policy/value heads over a small board.

    With planes=3, board=(3,3), width=32, depth=3, actions=9 this is the architecture of the
    reference's TicTacToe net (handyrl/envs/tictactoe.py:52-69: 3x3 conv stem with bias,
    `depth` x [3x3 conv without bias + BatchNorm], 1x1-conv heads with LeakyReLU(0.1) and a
    bias-free Linear; tanh value) -- 29,006 parameters.
    """

    def __init__(self, planes=3, board=(3, 3), width=32, depth=3, actions=9, policy_maps=2, value_maps=1,
                 return_head=False, norm=True):
        super().__init__()
        cells = board[0] * board[1]
        self.stem = nn.Conv2d(planes, width, 3, padding=1)
        self.tower = nn.ModuleList()
        for _ in range(depth):
            if norm:
                self.tower.append(nn.Sequential(nn.Conv2d(width, width, 3, padding=1, bias=False), nn.BatchNorm2d(width)))
            else:       # no batch statistics: a sharded step is then exactly the full-batch step (multi-GPU parity tests)
                self.tower.append(nn.Sequential(nn.Conv2d(width, width, 3, padding=1, bias=True)))
        self.p_squeeze = nn.Conv2d(width, policy_maps, 1)
        self.p_out = nn.Linear(cells * policy_maps, actions, bias=False)
        self.v_squeeze = nn.Conv2d(width, value_maps, 1)
        self.v_out = nn.Linear(cells * value_maps, 1, bias=False)
        self.r_squeeze = self.r_out = None
        if return_head:
            self.r_squeeze = nn.Conv2d(width, value_maps, 1)
            self.r_out = nn.Linear(cells * value_maps, 1, bias=False)

    def forward(self, x, hidden=None):
        h = F.relu(self.stem(x))
        for blk in self.tower:
            h = F.relu(blk(h))
        out = {
            'policy': self.p_out(F.leaky_relu(self.p_squeeze(h), 0.1).flatten(1)),
            'value': torch.tanh(self.v_out(F.leaky_relu(self.v_squeeze(h), 0.1).flatten(1))),
        }
        if self.r_out is not None:
            out['return'] = self.r_out(F.leaky_relu(self.r_squeeze(h), 0.1).flatten(1))
        return out


def load_state_by_order(module, state):
    """Load a state dict whose entries come in the same order and shapes but under other names
    (e.g. the reference SimpleConv2dModel's) into `module`."""
    own = module.state_dict()
    assert len(own) == len(state), (len(own), len(state))
    mapped = {}
    for (k_own, v_own), (k_src, v_src) in zip(own.items(), state.items()):
        v_src = torch.as_tensor(v_src)
        assert tuple(v_own.shape) == tuple(v_src.shape), (k_own, k_src, v_own.shape, v_src.shape)
        mapped[k_own] = v_src
    module.load_state_dict(mapped)
    return module


class GatedBoardNet(nn.Module):
    """Small recurrent net (conv-gated memory over the board) with policy / value / return heads
    and a dict observation {'scalar': (S,), 'board': (C,H,W)} -- the SHAPE of the reference's
    Geister net interface (geister.py:101-167: dict obs, init_hidden, 3 heads), much smaller.
    Used to exercise the recurrent path (burn-in, hidden masking) end to end."""

    def __init__(self, scalars=4, planes=3, board=(4, 4), width=8, actions=12):
        super().__init__()
        self.board, self.width = board, width
        cells = board[0] * board[1]
        self.embed = nn.Linear(scalars, width)
        self.inp = nn.Conv2d(planes + width, width, 3, padding=1)
        self.norm = nn.BatchNorm2d(width)
        self.gates = nn.Conv2d(2 * width, 4 * width, 3, padding=1)
        self.p_out = nn.Linear(cells * width, actions)
        self.v_out = nn.Linear(cells * width, 1)
        self.r_out = nn.Linear(cells * width, 1)

    def init_hidden(self, batch_size=None):
        shape = tuple(batch_size or []) + (self.width,) + tuple(self.board)
        return (torch.zeros(shape), torch.zeros(shape))

    def forward(self, x, hidden):
        s = self.embed(x['scalar'])[:, :, None, None].expand(-1, -1, *self.board)
        h_in = F.relu(self.norm(self.inp(torch.cat([x['board'], s], 1))))
        h, c = hidden
        i, f, o, g = self.gates(torch.cat([h_in, h], 1)).chunk(4, 1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        flat = h.flatten(1)
        return {'policy': self.p_out(flat), 'value': torch.tanh(self.v_out(flat)), 'return': self.r_out(flat),
                'hidden': (h, c)}


class WideActionNet(nn.Module):
    """Stand-in for BASELINE config 5 (64x64 one-plane observation, 512 actions): strided conv
    stack down to 4x4, then Linear -> 512 logits and a tanh value."""

    def __init__(self, planes=1, actions=512, width=32):
        super().__init__()
        chans = [planes, width, width, 2 * width, 2 * width]
        self.convs = nn.ModuleList(
            nn.Sequential(nn.Conv2d(chans[i], chans[i + 1], 3, stride=2, padding=1, bias=False), nn.BatchNorm2d(chans[i + 1]))
            for i in range(4))
        self.p_out = nn.Linear(2 * width * 16, actions)
        self.v_out = nn.Linear(2 * width * 16, 1)

    def forward(self, x, hidden=None):
        h = x
        for c in self.convs:
            h = F.relu(c(h))
        h = h.flatten(1)
        return {'policy': self.p_out(h), 'value': torch.tanh(self.v_out(h))}


def tictactoe_net():
    return BoardNet(planes=3, board=(3, 3), width=32, depth=3, actions=9)


class ConvLstmCell(nn.Module):
    """Convolutional LSTM cell: one 3x3 convolution over [input, h] produces the four gate maps
    (order i, f, o, g along the channel axis, as in the reference's cell, geister.py:43-56)."""

    def __init__(self, in_maps, state_maps, ksize=3):
        super().__init__()
        self.state_maps = state_maps
        self.conv = nn.Conv2d(in_maps + state_maps, 4 * state_maps, ksize, padding=ksize // 2, bias=True)

    def forward(self, x, state):
        h, c = state
        i, f, o, g = self.conv(torch.cat([x, h], dim=1)).chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        return torch.sigmoid(o) * torch.tanh(c), c


class _MoveHead(nn.Module):
    def __init__(self, maps, mid, out_maps):
        super().__init__()
        self.reduce = nn.Conv2d(maps, mid, 3, padding=1, bias=False)
        self.norm = nn.BatchNorm2d(mid)
        self.project = nn.Conv2d(mid, out_maps, 1, bias=False)

    def forward(self, h):
        return self.project(F.relu(self.norm(self.reduce(h)))).flatten(1)


class _ScalarHead(nn.Module):
    def __init__(self, maps, mid, cells, outputs):
        super().__init__()
        self.reduce = nn.Conv2d(maps, mid, 1, bias=False)
        self.norm = nn.BatchNorm2d(mid)
        self.out = nn.Linear(cells * mid, outputs, bias=False)

    def forward(self, h):
        return self.out(F.relu(self.norm(self.reduce(h))).flatten(1))


class DrcBoardNet(nn.Module):
    """Deep-repeated ConvLSTM net over a 6x6 board with a dict observation {'scalar': (18,), 'board': (7,6,6)}:
    the architecture of the reference's GeisterNet (geister.py:66-98, 101-167) -- stem conv+BN over the scalar planes
    stacked on the board planes, `depth` ConvLSTM cells applied `repeats` times per step (cell i > 0 reads the fresh
    h of cell i-1), a 3x3-conv/BN/1x1-conv move head whose 4x36 logits are followed by 70 "set" logits computed
    from the turn colour, and conv/BN/Linear value (tanh) and return heads.  231,604 parameters; its state_dict
    has the reference's order and shapes, so reference weights load with `load_state_by_order`."""

    def __init__(self, scalars=18, planes=7, board=(6, 6), width=32, depth=3, repeats=3, move_maps=4, set_actions=70):
        super().__init__()
        self.board, self.width, self.depth, self.repeats = tuple(board), width, depth, repeats
        cells = board[0] * board[1]
        self.stem = nn.Conv2d(scalars + planes, width, 3, padding=1, bias=False)
        self.stem_norm = nn.BatchNorm2d(width)
        self.cells = nn.ModuleList(ConvLstmCell(width, width) for _ in range(depth))
        self.move_head = _MoveHead(width, 8, move_maps)
        self.set_head = nn.Linear(1, set_actions, bias=True)
        self.value_head = _ScalarHead(width, 2, cells, 1)
        self.return_head = _ScalarHead(width, 2, cells, 1)

    def init_hidden(self, batch_size=None):
        shape = tuple(batch_size or []) + (self.width,) + self.board
        return ([torch.zeros(shape) for _ in range(self.depth)], [torch.zeros(shape) for _ in range(self.depth)])

    def forward(self, x, hidden):
        board, scalar = x['board'], x['scalar']
        planes = scalar[:, :, None, None].expand(-1, -1, *self.board)
        e = F.relu(self.stem_norm(self.stem(torch.cat([planes, board], dim=1))))
        if hidden is None:
            hidden = self.init_hidden([board.shape[0]])
            hidden = tuple([t.to(board.device) for t in part] for part in hidden)
        hs, cs = list(hidden[0]), list(hidden[1])
        for _ in range(self.repeats):
            for i, cell in enumerate(self.cells):
                hs[i], cs[i] = cell(hs[i - 1] if i > 0 else e, (hs[i], cs[i]))
        top = hs[-1]
        policy = torch.cat([self.move_head(top), self.set_head(scalar[:, :1])], dim=1)
        return {'policy': policy, 'value': torch.tanh(self.value_head(top)), 'return': self.return_head(top),
                'hidden': (hs, cs)}


class _TorusBlock(nn.Module):
    def __init__(self, in_maps, out_maps):
        super().__init__()
        # wrap-around padding on both board axes == the reference's explicit edge concatenation (hungry_geese.py:30-32)
        self.conv = nn.Conv2d(in_maps, out_maps, 3, padding=1, padding_mode='circular')
        self.bn = nn.BatchNorm2d(out_maps)

    def forward(self, x):
        return self.bn(self.conv(x))


class TorusNet(nn.Module):
    """Residual tower of wrap-around 3x3 convolutions over a 7x11 torus, policy from the features at the marked
    head cell (observation plane 0), value from [head features, mean features]: the architecture of the reference's
    GeeseNet (hungry_geese.py:23-57).  Same state_dict order and shapes, so reference weights load by order."""

    def __init__(self, planes=17, width=32, depth=12, actions=4):
        super().__init__()
        self.stem = _TorusBlock(planes, width)
        self.tower = nn.ModuleList(_TorusBlock(width, width) for _ in range(depth))
        self.p_out = nn.Linear(width, actions, bias=False)
        self.v_out = nn.Linear(2 * width, 1, bias=False)

    def forward(self, x, hidden=None):
        h = F.relu(self.stem(x))
        for blk in self.tower:
            h = F.relu(h + blk(h))
        flat = h.flatten(2)
        at_head = (flat * x[:, :1].flatten(2)).sum(-1)
        return {'policy': self.p_out(at_head), 'value': torch.tanh(self.v_out(torch.cat([at_head, flat.mean(-1)], dim=1)))}


def geister_net():
    return DrcBoardNet()


def geese_net():
    return TorusNet()
