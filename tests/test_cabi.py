"""The C-ABI shared library loads on a machine without a GPU and exports every symbol that include/b200diff.h
declares (no compute calls here).  Also: the product path refuses to run without CUDA instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'b200diff.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(b200_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported():
    import b200diff
    lib = ctypes.CDLL(b200diff.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 13
    for sym in declared:
        assert hasattr(lib, sym), f'{sym} declared in include/b200diff.h but not exported'
    assert set(b200diff.EXPORTED_SYMBOLS) == set(declared)


def test_version_and_error_string():
    import b200diff
    lib = b200diff.lib()
    assert lib.b200_version() == 200
    assert isinstance(lib.b200_last_error(), bytes)
    assert b200diff.direct_launch_count() >= 0


def test_struct_layout_matches_header(tmp_path):
    """The ctypes mirrors of the descriptor structs agree with the C compiler's layout of include/b200diff.h
    (offsetof / sizeof printed by a tiny program built with gcc)."""
    import shutil
    import subprocess
    import b200diff
    if shutil.which('gcc') is None:
        pytest.skip('gcc not available')
    fields = {'b200_conv_desc': b200diff.ConvDesc, 'b200_sampler_desc': b200diff.SamplerDesc,
              'b200_gemm_desc': b200diff.GemmDesc, 'b200_wgrad_desc': b200diff.WgradDesc,
              'b200_gn_bwd_desc': b200diff.GnBwdDesc, 'b200_optim_desc': b200diff.OptimDesc,
              'b200_ode_desc': b200diff.OdeDesc, 'b200_gn_fuse_desc': b200diff.GnFuseDesc,
              'b200_attn_block_desc': b200diff.AttnBlockDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200diff.h"', 'int main(void) {']
    for cname, cls in fields.items():
        for fname, ftype in cls._fields_:
            if isinstance(ftype, type) and issubclass(ftype, (ctypes.Structure, ctypes.Array)):
                continue      # nested structs (gemm operands) / arrays: covered by the offsets of the members that follow them
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
        lines.append(f'  printf("{cname}.sizeof %zu\\n", sizeof({cname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in fields.items():
        for fname, ftype in cls._fields_:
            if isinstance(ftype, type) and issubclass(ftype, (ctypes.Structure, ctypes.Array)):
                continue
            assert int(out[f'{cname}.{fname}']) == getattr(cls, fname).offset, (cname, fname)
        assert int(out[f'{cname}.sizeof']) == ctypes.sizeof(cls), cname
    # the pack / optimizer tables are serialised with struct.pack: their C sizes are part of the contract too
    src.write_text('#include <stdio.h>\n#include "b200diff.h"\nint main(void) { printf("%zu %zu\\n", '
                   'sizeof(b200_pack_entry), sizeof(b200_optim_chunk)); return 0; }')
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    import struct
    sz = subprocess.check_output([str(exe)], text=True).split()
    assert int(sz[0]) == struct.calcsize('<3Q8i') and int(sz[1]) == struct.calcsize('<5Q2i')


def test_no_cpu_fallback():
    import diffusions
    import models
    torch.manual_seed(0)
    m = models.UNet(in_channels=1, out_channels=1, dim=32).eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match='CUDA'):
        m(torch.zeros(1, 1, 32, 32), torch.zeros(1, dtype=torch.long))
    d = diffusions.DDIM(respace_type='uniform', respace_steps=10)
    with pytest.raises(RuntimeError, match='CUDA'):
        d.denoise(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4, 4), 900, 800)
    with pytest.raises(RuntimeError):
        m.train()
        m(torch.zeros(1, 1, 32, 32), torch.zeros(1, dtype=torch.long))
