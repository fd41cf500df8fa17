"""End to end: the REFERENCE's Learner, worker processes and server loop on top of handyrl_b200's Trainer.

    python tests/e2e_reference_learner.py [--uniform-net] [--epochs N]

Needs the reference checkout (HANDYRL_REFERENCE, default /root/reference) -- it is NOT copied into this repo, so
this script only runs where it is mounted.  With a CUDA device it trains the reference's own TicTacToe net through
the GPU learner; with `--uniform-net` (no GPU needed) the environment's net is replaced by a parameter-free model, which
exercises everything around the optimiser step: install(), Learner.feed_episodes -> Trainer.episodes, the trainer
thread protocol, update() hand-offs, pickling the returned model for the workers and the workers unpickling it.
Prints E2E_OK on success.
"""
import argparse
import os
import sys
import tempfile

REF = os.environ.get('HANDYRL_REFERENCE', '/root/reference')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)


import torch  # noqa: E402


class UniformNet(torch.nn.Module):
    """Parameter-free stand-in for the environment's net (module level: the Learner pickles it for the workers)."""

    def forward(self, x, hidden=None):
        return {'policy': torch.zeros(x.shape[0], 9), 'value': torch.zeros(x.shape[0], 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--uniform-net', action='store_true')
    ap.add_argument('--epochs', type=int, default=2)
    opt = ap.parse_args()
    os.chdir(tempfile.mkdtemp(prefix='hrl_e2e_'))           # the Learner writes models/<epoch>.pth into the cwd

    import handyrl_b200.train as b200
    ref = b200.install()                                       # the three lines INTEGRATION.md adds to main.py
    assert ref.Trainer is b200.Trainer

    if opt.uniform_net:
        import handyrl.envs.tictactoe as ttt
        ttt.Environment.net = lambda self: UniformNet()

    args = {
        'env_args': {'env': 'TicTacToe'},
        'train_args': {
            'turn_based_training': True, 'observation': False, 'gamma': 0.8, 'forward_steps': 8, 'burn_in_steps': 0,
            'compress_steps': 4, 'entropy_regularization': 0.1, 'entropy_regularization_decay': 0.1,
            'update_episodes': 40, 'batch_size': 16, 'minimum_episodes': 40, 'maximum_episodes': 500,
            'epochs': opt.epochs, 'num_batchers': 1, 'eval_rate': 0.1, 'worker': {'num_parallel': 2}, 'lambda': 0.7,
            'policy_target': 'UPGO', 'value_target': 'VTRACE', 'eval': {'opponent': ['random']}, 'seed': 0,
            'restart_epoch': 0,
        },
        'worker_args': {'server_address': '', 'num_parallel': 2},
    }
    from handyrl.environment import prepare_env
    prepare_env(args['env_args'])
    learner = ref.Learner(args=args)
    assert isinstance(learner.trainer, b200.Trainer)
    learner.run()
    assert learner.model_epoch >= opt.epochs, learner.model_epoch
    assert os.path.exists(os.path.join('models', '%d.pth' % opt.epochs))
    if not opt.uniform_net:
        assert learner.trainer.steps > 0
        first = torch.load(os.path.join('models', '1.pth'))
        last = torch.load(os.path.join('models', '%d.pth' % opt.epochs))
        assert any(not torch.equal(first[k], last[k]) for k in first)
    print('E2E_OK epochs=%d steps=%d episodes=%d' % (learner.model_epoch, learner.trainer.steps, learner.num_returned_episodes))
    os._exit(0)          # the reference's daemon threads / worker pipes have no shutdown path


if __name__ == '__main__':
    main()
