"""The C-ABI library loads on a machine without a GPU and exports exactly the
symbols that include/gradpath.h declares (no compute calls here)."""
import os
import re
import subprocess

from chainer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, 'include', 'gradpath.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return set(re.findall(r'\b(gp_[a-z0-9_]+)\s*\(', text))


def test_header_prototypes_and_binding_agree():
    declared = _declared()
    assert declared == set(_lib.PROTOTYPES), (declared ^ set(_lib.PROTOTYPES))


def test_library_loads_and_exports_every_symbol():
    path = _lib.library_path()
    assert os.path.exists(path), 'run `python -c "import __graft_entry__ as g; g.build()"` first'
    out = subprocess.run(['nm', '-D', '--defined-only', path], stdout=subprocess.PIPE, text=True,
                         check=True).stdout
    exported = set(re.findall(r'\bT (gp_[a-z0-9_]+)\b', out))
    assert _declared() <= exported, _declared() - exported
    lib = _lib._Lib(path)                      # resolves every prototype through ctypes
    assert lib.gp_abi_version() == 1


def test_seg_struct_layout_matches_header():
    assert _lib.SEG_DTYPE.itemsize == 64
    assert _lib.SEG_DTYPE.fields['buf_off'][1] == 40
    assert _lib.SEG_DTYPE.fields['dtype0'][1] == 48
    assert _lib.SEG_DTYPE.fields['flags'][1] == 56


def test_errors_are_reported_without_a_gpu():
    import ctypes
    import pytest
    lib = _lib._Lib(_lib.library_path())
    with pytest.raises(_lib.GradpathError) as e:
        lib.gp_set_tuning(b'no_such_knob', 1)
    assert e.value.code == -22 and 'no_such_knob' in str(e.value)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        n = ctypes.c_int()
        with pytest.raises(_lib.GradpathError):
            lib.gp_device_count(ctypes.byref(n))      # CUDA error surfaces, no crash


def test_bn_workspace_header_does_not_depend_on_the_channel_count():
    """One statistics workspace serves every BN layer of a communicator.  The words the kernels
    need ZERO between launches (channels-done counter, channel tickets) must therefore sit at
    offsets that do not move with C, in front of the scratch that is left holding junk (split
    partials, staged local statistics) -- otherwise a C = 64 layer's partials land on a C = 128
    layer's tickets.  Host-side query only, no launch."""
    import ctypes
    lib = _lib._Lib(_lib.library_path())
    seen = None
    for C in (1, 3, 64, 128, 256, 512, 1024, 2048, 4096, 20000):
        out = (ctypes.c_int64 * 4)()
        lib.gp_bn_workspace_layout(C, out)
        h0, h1, s0, s1 = list(out)
        assert (h0, h1) == (seen or (h0, h1)), 'header moved with C'
        seen = (h0, h1)
        assert h0 == 0 and h1 >= 256 + 4 * min(C, 16384)
        assert s0 >= h1 and s1 == lib.gp_bn_workspace_bytes(C)
        assert s1 - s0 >= C * 8          # room for the 2C staged float statistics at least
