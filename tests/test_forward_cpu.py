"""Host logic of the recurrent / feed-forward net driver vs the reference's forward_prediction
(golden rnn_cases.pkl; the torch ops run on CPU here -- no loss kernel involved)."""
import os
import pickle

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from handyrl_b200.batch import tree_map
from handyrl_b200.nets import GatedBoardNet, BoardNet, load_state_by_order, tictactoe_net
from handyrl_b200.train import forward_prediction, forward_raw, BatchLayout, PackedBatch

with open(os.path.join(GOLDEN, 'rnn_cases.pkl'), 'rb') as f:
    RNN_CASES = pickle.load(f)


@pytest.mark.parametrize('name', sorted(RNN_CASES))
def test_recurrent_forward_matches_reference(name):
    c = RNN_CASES[name]
    net = GatedBoardNet()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in c['state0'].items()})
    net.train()
    batch = tree_map(lambda a: torch.from_numpy(a), c['batch'])
    B, T, P, A = c['dims']
    got = forward_prediction(net, net.init_hidden([B, P]), batch, c['args'])
    assert set(got) == set(c['masked_outputs'])
    for k, ref in c['masked_outputs'].items():
        np.testing.assert_allclose(got[k].detach().numpy(), ref, rtol=1e-5, atol=1e-5, err_msg=k)


def test_tictactoe_architecture_parameter_count():
    net = tictactoe_net()
    assert sum(p.numel() for p in net.parameters()) == 29006


def test_feed_forward_raw_shapes():
    from handyrl_b200.synthetic import synthetic_batch
    batch = synthetic_batch(3, 5, 2, 9, seed=2)
    outs = forward_raw(tictactoe_net(), None, batch, {'turn_based_training': True, 'observation': False, 'burn_in_steps': 0})
    assert outs['policy'].shape == (3, 5, 1, 9) and outs['value'].shape == (3, 5, 1, 1)


def test_packed_batch_round_trip():
    from handyrl_b200.synthetic import synthetic_batch
    if not torch.cuda.is_available():
        pytest.skip('pinned memory needs a CUDA runtime')
    batch = synthetic_batch(3, 5, 2, 9, seed=2)
    layout = BatchLayout(batch)
    pk = PackedBatch(layout).fill(batch)
    for k, v in pk.tensors.items():
        assert torch.equal(v, batch[k]), k
    assert 'value' not in pk.tensors and layout.nbytes % 256 == 0


def test_batch_layout_offsets_are_aligned_and_disjoint():
    from handyrl_b200.synthetic import synthetic_batch
    batch = synthetic_batch(3, 5, 2, 9, seed=2)
    batch['observation'] = {'scalar': torch.zeros(3, 5, 1, 4), 'board': torch.zeros(3, 5, 1, 3, 4, 4)}
    layout = BatchLayout(batch)
    spans = sorted((off, off + int(np.prod(sh)) * torch.empty(0, dtype=dt).element_size()) for _, sh, dt, off in layout.entries)
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0 and b0 % 256 == 0
    views = layout.views(torch.zeros(layout.nbytes, dtype=torch.uint8))
    assert views['observation']['board'].shape == (3, 5, 1, 3, 4, 4) and views['action'].dtype == torch.int64
