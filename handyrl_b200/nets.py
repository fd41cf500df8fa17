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
                 return_head=False):
        super().__init__()
        cells = board[0] * board[1]
        self.stem = nn.Conv2d(planes, width, 3, padding=1)
        self.tower = nn.ModuleList()
        for _ in range(depth):
            self.tower.append(nn.Sequential(nn.Conv2d(width, width, 3, padding=1, bias=False), nn.BatchNorm2d(width)))
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
