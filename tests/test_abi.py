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
