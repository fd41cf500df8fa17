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
    structs = ['HrlLossArgs', 'HrlWindow', 'HrlGatherArgs', 'HrlLossTuning', 'HrlGemmOperand', 'HrlGemmArgs', 'HrlPackJob', 'HrlFoldJob']
    fields = [('HrlLossArgs', 'io_bf16'), ('HrlLossArgs', 'tuning'), ('HrlGemmArgs', 'col_partials'), ('HrlGemmArgs', 'conv_off'),
              ('HrlGemmArgs', 'conv_cin'), ('HrlGemmArgs', 'seg_a'), ('HrlGemmArgs', 'conv_ones_row'), ('HrlGemmOperand', 'packed'),
              ('HrlPackJob', 'bias_cells'), ('HrlFoldJob', 'dw')]
    src = ('#include <stdio.h>\n#include <stddef.h>\n#include "hrl_b200.h"\nint main(){' +
           ''.join('printf("%%zu\\n", sizeof(%s));' % n for n in structs) +
           ''.join('printf("%%zu\\n", offsetof(%s, %s));' % f for f in fields) + 'return 0;}\n')
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, 's.c')
        open(c, 'w').write(src)
        exe = os.path.join(d, 's')
        subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), c, '-o', exe], check=True)
        got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(getattr(_capi, n)) for n in structs] + [getattr(getattr(_capi, n), f).offset for n, f in fields]
    assert got == want, list(zip(structs + ['%s.%s' % f for f in fields], got, want))


def test_convolution_geometry_table_and_packed_image_sizes():
    """Host-side helpers of the implicit convolutions: hrl_conv_geometry (neighbour offsets, zero padding = outside, torus = wrap),
    hrl_gemm_padded_rows / hrl_board_pack_floats / hrl_conv_pack_floats (operand image sizes)."""
    import numpy as np
    from handyrl_b200._capi import check, lib
    OUT = -32768
    for H, W, kh, kw, wrap in ((6, 6, 3, 3, 0), (7, 11, 3, 3, 1), (3, 3, 1, 3, 0), (4, 5, 3, 1, 1)):
        tab = np.empty(H * W * kh * kw, dtype=np.int16)
        check(lib().hrl_conv_geometry(H, W, kh, kw, wrap, tab.ctypes.data))
        tab = tab.reshape(H * W, kh * kw)
        for y in range(H):
            for x in range(W):
                for a in range(kh):
                    for b in range(kw):
                        yy, xx = y + a - kh // 2, x + b - kw // 2
                        if wrap:
                            want = (yy % H) * W + (xx % W) - (y * W + x)
                        else:
                            want = (yy * W + xx) - (y * W + x) if (0 <= yy < H and 0 <= xx < W) else OUT
                        assert tab[y * W + x, a * kw + b] == want
        assert (tab != OUT).all() if wrap else (tab == OUT).any()
    assert [lib().hrl_gemm_padded_rows(n) for n in (1, 16, 27, 128, 256, 257, 288)] == [16, 16, 32, 128, 256, 288, 288]
    assert lib().hrl_board_pack_floats(288, 288) == 9 * 2 * 288 * 32          # 9 chunks x (hi | lo) x 288 rows x 32 elements
    assert lib().hrl_conv_pack_floats(128, 64, 9) == lib().hrl_board_pack_floats(128, 9 * 64)
    assert lib().hrl_conv_pack_floats(32, 17, 9) == lib().hrl_board_pack_floats(32, 9 * 32)      # channels padded to whole chunks
    assert lib().hrl_conv_geometry(20, 20, 3, 3, 0, tab.ctypes.data) != 0                          # more than 256 cells: refused


def test_block_layout_helpers_keep_channels_last_blocks():
    """ops._block_layout / _as_block_layout / _empty_block_layout (host logic of the layout-agnostic hidden-state kernels)."""
    import torch
    from handyrl_b200 import ops
    t = torch.randn(6, 8, 3, 4).contiguous(memory_format=torch.channels_last)
    u = t.unflatten(0, (3, 2))
    assert ops._block_layout(t) == 'cl' and ops._block_layout(u) == 'cl' and ops._block_layout(torch.randn(3, 2, 8, 3, 4)) == 'std'
    assert ops._block_layout(u.transpose(3, 4)) is None
    e = ops._empty_block_layout((3, 2, 8, 3, 4), 'cl', 'cpu')
    assert e.shape == (3, 2, 8, 3, 4) and ops._block_layout(e) == 'cl'
    x = torch.randn(3, 2, 8, 3, 4)
    a = ops._as_block_layout(x, 'cl')
    assert ops._block_layout(a) == 'cl' and torch.equal(a, x) and ops._as_block_layout(a, 'cl') is a
    assert ops._as_block_layout(a, 'std').is_contiguous()
    assert ops._empty_block_layout((5, 7), 'cl', 'cpu').is_contiguous()          # not a (C,H,W) block: plain layout


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
