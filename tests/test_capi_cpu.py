"""CPU-only checks of the boundary: the library loads and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'hrl_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(hrl_[a-z0-9_]+)\s*\(', text)))


def test_header_and_binding_agree():
    from handyrl_b200 import _capi
    assert header_symbols() == sorted(_capi.SYMBOLS)


def test_library_exports_every_symbol():
    import __graft_entry__ as g
    g.build()
    from handyrl_b200 import _capi
    h = ctypes.CDLL(_capi.LIB_PATH)
    for name in header_symbols():
        assert hasattr(h, name), name
    lib = _capi.lib()
    assert lib.hrl_abi_version() == _capi.HRL_ABI_VERSION
    assert lib.hrl_loss_workspace_bytes(512, 32, 2, 1, 9) >= 256 + 512 * 32
    assert lib.hrl_sumsq_num_partials() > 0
    assert lib.hrl_last_error() == b''


def test_struct_layout_matches_header():
    """ctypes mirrors must have the C struct sizes (checked against a tiny C program)."""
    import subprocess
    import tempfile
    from handyrl_b200 import _capi
    src = '#include <stdio.h>\n#include "hrl_b200.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(HrlLossArgs), sizeof(HrlWindow), sizeof(HrlGatherArgs));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, 's.c')
        open(c, 'w').write(src)
        exe = os.path.join(d, 's')
        subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), c, '-o', exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(_capi.HrlLossArgs), ctypes.sizeof(_capi.HrlWindow), ctypes.sizeof(_capi.HrlGatherArgs)]


def test_ops_refuse_cpu_tensors():
    import torch
    from handyrl_b200 import ops, _capi
    from handyrl_b200.synthetic import synthetic_batch, synthetic_outputs
    batch = synthetic_batch(2, 4, 2, 9, with_obs=False)
    args = {'turn_based_training': True, 'gamma': 0.8, 'lambda': 0.7, 'burn_in_steps': 0,
            'entropy_regularization': 0.1, 'entropy_regularization_decay': 0.1,
            'policy_target': 'UPGO', 'value_target': 'VTRACE'}
    with pytest.raises(_capi.HrlError):
        ops.loss_fwd_bwd(synthetic_outputs(batch), batch, args)
