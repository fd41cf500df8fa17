"""B200-native learner hot path for HandyRL (see DESIGN.md / INTEGRATION.md).

    handyrl_b200.train     Trainer / Batcher / make_batch / compute_loss drop-ins, LearnerStep, install()
    handyrl_b200.ops       tensor-level wrappers of the C ABI (include/hrl_b200.h), FlatAdam, PeerAllReduce
    handyrl_b200.replay    GPU-resident replay + gather/pad kernel
    handyrl_b200.batch     host-side episode decoding and window sampling
    handyrl_b200.wire      flat episode wire format for workers
    handyrl_b200.fastnet   small-board rewrite pass for user nets
    handyrl_b200.multigpu  multi-GPU learner: helper ranks behind Trainer, sharding helpers

The CUDA library (handyrl_b200/libhrl_b200.so) is built by `__graft_entry__.build()`; there is no CPU fallback.
"""
__version__ = '0.1.0'
