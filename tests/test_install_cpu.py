"""install() swaps the learner hot path into an importable `handyrl` (only checkable where the reference is mounted)."""
import os
import sys

import pytest

REF = os.environ.get('HANDYRL_REFERENCE', '/root/reference')


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'handyrl')), reason='reference checkout not mounted')
def test_install_swaps_reference_symbols():
    sys.path.insert(0, REF)
    try:
        import handyrl.train as ref_train
        import handyrl.losses as ref_losses
        originals = {k: getattr(ref_train, k) for k in ('Trainer', 'Batcher', 'make_batch', 'forward_prediction', 'compute_loss')}
        orig_target = ref_losses.compute_target
        import handyrl_b200.train as b200
        from handyrl_b200 import ops
        try:
            b200.install()
            assert ref_train.Trainer is b200.Trainer and ref_train.Batcher is b200.Batcher
            assert ref_train.make_batch is b200.make_batch and ref_train.compute_loss is b200.compute_loss
            assert ref_losses.compute_target is ops.compute_target
            # the reference Learner builds its trainer through the module attribute (train.py:439)
            assert 'Trainer(args, copy.deepcopy(self.model))' in open(os.path.join(REF, 'handyrl', 'train.py')).read()
        finally:
            for k, v in originals.items():
                setattr(ref_train, k, v)
            ref_losses.compute_target = orig_target
    finally:
        sys.path.remove(REF)
        for m in [m for m in sys.modules if m == 'handyrl' or m.startswith('handyrl.')]:
            del sys.modules[m]


def test_trainer_constructor_mirrors_reference_attributes():
    """Attributes the reference Learner touches on its Trainer (train.py:439, 472, 482-483, 533)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('covered by the GPU suite')
    from handyrl_b200.train import Trainer

    class NoParams(torch.nn.Module):
        def forward(self, x, h=None):
            return {'policy': x}

    args = {'batch_size': 4, 'forward_steps': 8, 'burn_in_steps': 0, 'num_batchers': 1, 'maximum_episodes': 10}
    tr = Trainer(args, NoParams())          # non-parametric model: allowed without a GPU (train.py:348-350)
    assert len(tr.episodes) == 0 and tr.steps == 0 and hasattr(tr, 'update') and hasattr(tr, 'run')
    tr.episodes.append({'steps': 3})
    tr.episodes.extend([{'steps': 4}])
    assert len(tr.episodes) == 2 and tr.episodes.popleft()['steps'] == 3
    assert tr.train() is tr.model            # sleeps 0.1 s and hands the model back
